#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -q -m gpu --tb=line > gpurun_out/test_gpu_train.log 2>&1; echo "train tests exit $?" > gpurun_out/summary.txt
tail -2 gpurun_out/test_gpu_train.log >> gpurun_out/summary.txt
python bench.py --steps 80 --warmup 40 --skip-extras > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 80 --warmup 40 --skip-extras > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches exit $?" >> gpurun_out/summary.txt
python tools/prof_target.py 3 > gpurun_out/plain_prof.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'preprocess_kernel|conv_tc_fwd_kernel|conv_fp32_kernel|adam_kernel' -c 15 \
    -o gpurun_out/prof_r1 -f python tools/prof_target.py 3 > gpurun_out/ncu_prof.log 2>&1; echo "ncu full exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
