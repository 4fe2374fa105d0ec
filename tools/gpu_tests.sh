#!/bin/bash
# usage: tools/gpu_tests.sh [pytest -k expression]
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for f in test_gpu_forward test_gpu_train; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu --tb=short -x ${1:+-k "$1"} > gpurun_out/$f.log 2>&1; echo "$f exit $?" >> gpurun_out/summary.txt
  tail -3 gpurun_out/$f.log >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt
