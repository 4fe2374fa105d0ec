#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
S=gpurun_out/summary.txt
for C in 1 2 3 4; do B2H_K0_CTAS=$C python tools/k0_prof.py time 108000 144000 180000 216000 432000 864000 2>&1 | grep "us/launch" >> $S; done
cat $S
