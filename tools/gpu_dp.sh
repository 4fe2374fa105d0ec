#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 240 python -m pytest tests/test_gpu_dp.py -q -m gpu --tb=short -x > gpurun_out/test_gpu_dp.log 2>&1; echo "dp tests exit $?" >> gpurun_out/summary.txt
tail -15 gpurun_out/test_gpu_dp.log >> gpurun_out/summary.txt
timeout 120 python bench.py --skip-extras > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench n1 exit $?" >> gpurun_out/summary.txt
timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N exit $?" >> gpurun_out/summary.txt
tail -3 gpurun_out/bench_n$N.err >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
python -c "
import json
for n in (1, $N):
    d = json.load(open('gpurun_out/bench_n%d.json' % n)); print(n, d['value'], d['ms_per_step'], d['config']['cuda_graph'], d['e2e']['value'])
"
