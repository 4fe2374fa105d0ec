#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
S=gpurun_out/summary.txt
python tools/k0_prof.py 216000 30 > gpurun_out/k0_plain.log 2>&1; echo "plain exit $?" >> $S; tail -2 gpurun_out/k0_plain.log >> $S
ncu --set full --clock-control none --import-source on -k regex:preprocess_kernel -c 2 -o gpurun_out/prof_k0 -f python tools/k0_prof.py 108000 2 > gpurun_out/ncu_k0.log 2>&1; echo "ncu exit $?" >> $S
cat $S
