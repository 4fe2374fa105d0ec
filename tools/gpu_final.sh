#!/bin/bash
# Round-end style run on one GPU: all parity tests, smoke, both bench arms, ncu launch list + full capture.
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests -q -m gpu --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu exit $?" >> gpurun_out/summary.txt
tail -4 gpurun_out/pytest_gpu.log >> gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
timeout 300 python bench.py --impl reference --steps 50 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref exit $?" >> gpurun_out/summary.txt
python bench.py --steps 80 --warmup 40 --skip-extras > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 80 --warmup 40 --skip-extras > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches exit $?" >> gpurun_out/summary.txt
python tools/prof_target.py 3 > gpurun_out/plain_prof.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'preprocess_kernel|conv_tc_tile_kernel|conv_tc_wide_fwd_kernel|adam_kernel' -c 15 \
    -o gpurun_out/prof_final -f python tools/prof_target.py 3 > gpurun_out/ncu_prof.log 2>&1; echo "ncu full exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
