#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_train.py -q -m gpu --tb=short -x > gpurun_out/test_gpu_train.log 2>&1; echo "train tests exit $?" >> gpurun_out/summary.txt
tail -2 gpurun_out/test_gpu_train.log >> gpurun_out/summary.txt
python tools/prof_target.py 3 > gpurun_out/plain_prof.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'preprocess_kernel|conv_tc_tile_kernel|adam_kernel' -c 12 \
    -o gpurun_out/prof_r1b -f python tools/prof_target.py 3 > gpurun_out/ncu_prof.log 2>&1; echo "ncu full exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
