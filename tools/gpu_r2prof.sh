#!/bin/bash
# round 2 profiles (1 GPU): ncu launch list of the bench command, ncu --set full of every hot kernel
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
S=gpurun_out/summary.txt
python bench.py --steps 20 --warmup 5 --skip-extras > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 20 --warmup 5 --skip-extras > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches exit $?" >> $S
python tools/prof_target.py 2 > gpurun_out/plain_prof.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'preprocess_kernel|conv_tc_tile_kernel|conv_tc_wide|adam_gp_wide' -c 40 \
    -o gpurun_out/prof_r2 -f python tools/prof_target.py 2 > gpurun_out/ncu_prof.log 2>&1; echo "ncu full exit $?" >> $S
ls -la gpurun_out/*.ncu-rep >> $S
tail -3 gpurun_out/plain_prof.log >> $S
cat $S
