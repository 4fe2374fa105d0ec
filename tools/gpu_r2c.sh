#!/bin/bash
# round 2, run C (8 GPUs): scaling check -- bench N=1 and N=8 (multicast / unicast push), DP phase stamps at N=8
N=${1:-8}
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
S=gpurun_out/summary.txt
nvidia-smi -L | wc -l >> $S
timeout 200 python bench.py --steps 20 --warmup 5 --skip-extras > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench n1 exit $?" >> $S
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N exit $?" >> $S
tail -3 gpurun_out/bench_n$N.err >> $S
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 20 --warmup 5 --no-multicast > gpurun_out/bench_n${N}_nomc.json 2> gpurun_out/bench_n${N}_nomc.err; echo "bench n$N nomc exit $?" >> $S
B2H_MULTICAST=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/phase_timing.py > gpurun_out/phase_n${N}_mc1.log 2>&1; echo "phase n$N exit $?" >> $S
grep -E "\[r0\]|\[r5\]" gpurun_out/phase_n${N}_mc1.log | grep -E "exchange=|rep 2|us/step" >> $S
grep -A1 "\[r0\] train, steady state (graph replay, rep 2): 3" gpurun_out/phase_n${N}_mc1.log | tail -1 >> $S
python - <<PY >> $S 2>&1
import json
for f in ('bench_n1', 'bench_n$N', 'bench_n${N}_nomc'):
    try:
        d = json.load(open('gpurun_out/%s.json' % f))
        print(f, d['n_gpus'], 'value', d['value'], 'us/step', d['ms_per_step'] * 1e3, 'e2e', d['e2e']['value'], 'e2e us', d['e2e']['us_per_step'], d.get('dp_parity'), d.get('tc_status'), d.get('status_ok'))
    except Exception as e:
        print(f, 'ERR', e)
PY
cat $S
