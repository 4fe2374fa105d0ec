#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for f in test_gpu_forward test_gpu_train; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu --tb=short -x > gpurun_out/$f.log 2>&1; echo "$f exit $?" >> gpurun_out/summary.txt
  tail -3 gpurun_out/$f.log >> gpurun_out/summary.txt
done
timeout 300 python tools/phase_timing.py > gpurun_out/phase_timing.log 2>&1; cat gpurun_out/phase_timing.log >> gpurun_out/summary.txt
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
