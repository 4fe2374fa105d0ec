#!/bin/bash
# round 2, run E (1 GPU): full tests, wide-training launch list, bench
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
S=gpurun_out/summary.txt
timeout 1500 python -m pytest tests -q -m gpu --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu exit $?" >> $S
tail -6 gpurun_out/pytest_gpu.log >> $S
for C in 256 128 64; do
  python tools/wide_train_prof.py $C >> $S 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 40 --csv --log-file gpurun_out/launches_wide_train.csv python tools/wide_train_prof.py 256 > gpurun_out/ncu_wide.log 2>&1; echo "ncu wide exit $?" >> $S
python - <<'PY' >> $S 2>&1
import csv
rows = [r for r in csv.reader(open('gpurun_out/launches_wide_train.csv')) if len(r) > 5 and r[0].isdigit()]
from collections import defaultdict
d = defaultdict(list)
for r in rows:
    name = r[4][:60]; val = float(r[-1].replace(',', ''))
    d[name].append(val)
for k, v in d.items():
    print(k, len(v), 'avg', sum(v) / len(v), r[-2] if rows else '')
PY
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> $S
tail -5 gpurun_out/bench.err >> $S
python - <<'PY' >> $S 2>&1
import json
d = json.load(open('gpurun_out/bench.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'tc_status', 'status_ok')})
for k in ('train_wide', 'train_wide_c128', 'train_c64', 'train_fp32', 'train_fp32_ffma', 'fwd_fp32', 'fwd_fp32_ffma'):
    print(k, json.dumps(d.get(k))[:500])
PY
cat $S
