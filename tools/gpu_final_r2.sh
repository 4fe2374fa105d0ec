#!/bin/bash
# Round-end style run on one GPU: all parity tests, smoke, both bench arms, ncu launch list + full capture.
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
S=gpurun_out/summary.txt
timeout 1500 python -m pytest tests -q -m gpu --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu exit $?" >> $S
tail -4 gpurun_out/pytest_gpu.log >> $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> $S
tail -3 gpurun_out/smoke.log >> $S
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> $S
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref exit $?" >> $S
timeout 300 python tools/phase_timing.py > gpurun_out/phase_timing.log 2>&1; echo "phase exit $?" >> $S
python bench.py --steps 20 --warmup 5 --skip-extras > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 20 --warmup 5 --skip-extras > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches exit $?" >> $S
python tools/prof_target.py 2 > gpurun_out/plain_prof.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'preprocess_kernel|conv_tc_tile_kernel|conv_tc_wide|adam_gp' -c 44 \
    -o gpurun_out/prof_r2 -f python tools/prof_target.py 2 > gpurun_out/ncu_prof.log 2>&1; echo "ncu full exit $?" >> $S
ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 40 --csv --log-file gpurun_out/launches_wide_train.csv python tools/wide_train_prof.py 256 > gpurun_out/ncu_wide.log 2>&1; echo "ncu wide exit $?" >> $S
ls -la gpurun_out/*.ncu-rep >> $S
cat $S
