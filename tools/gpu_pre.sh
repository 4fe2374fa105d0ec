#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_preprocess.py -q -m gpu --tb=short > gpurun_out/test_gpu_preprocess.log 2>&1; echo "pre tests exit $?" >> gpurun_out/summary.txt
tail -3 gpurun_out/test_gpu_preprocess.log >> gpurun_out/summary.txt
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
