#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for f in test_gpu_preprocess test_gpu_forward test_gpu_train; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu --tb=short > gpurun_out/$f.log 2>&1; echo "$f exit $?" >> gpurun_out/summary.txt
  tail -3 gpurun_out/$f.log >> gpurun_out/summary.txt
done
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
