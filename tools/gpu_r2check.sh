#!/bin/bash
# quick check of bench.py on N GPUs (default 1): full line at N=1, DP line at N>1
N=${1:-1}
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
S=gpurun_out/summary.txt
if [ "$N" = 1 ]; then
  timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> $S
  python - <<PY >> $S 2>&1
import json
d = json.load(open('gpurun_out/bench.json'))
print('us/step', d['ms_per_step'] * 1e3, 'e2e', d['e2e']['value'] / 1e6, d['e2e']['segment_us_per_step'])
print('pre', d['preprocess']['ms_per_launch'], d['preprocess']['roofline']['frac'], 'h5', d.get('preprocess_h5'))
PY
else
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N exit $?" >> $S
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; echo "bench ref n$N exit $?" >> $S
  python - <<PY >> $S 2>&1
import json
d = json.load(open('gpurun_out/bench_n$N.json'))
print('us/step', d['ms_per_step'] * 1e3, 'e2e', d['e2e']['value'] / 1e6, d['e2e']['segment_us_per_step'], d['roofline'])
print(open('gpurun_out/bench_ref_n$N.json').read()[:400])
PY
fi
cat $S
