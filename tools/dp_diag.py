"""Multi-GPU bring-up diagnostic (run under torchrun with a short timeout): prints each stage as it passes."""
import faulthandler, os, sys, time
faulthandler.dump_traceback_later(45, exit=True)
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
def say(*a):
    print(f"[r{rank} {time.time()%1000:7.2f}]", *a, flush=True)
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
say("init pg")
dist.init_process_group("nccl", device_id=dev)
say("pg ok")
x = torch.ones(19032, device=dev) * (rank + 1)
dist.all_reduce(x); torch.cuda.synchronize(); say("allreduce ok", float(x[0]))
import hand_pose_sl_b200 as b2h
from hand_pose_sl_b200 import parallel, synthetic
torch.manual_seed(0)
model = b2h.ConvModel(30, "ReLU", False, precision="bf16").to(dev)
opt = b2h.FusedAdam(model.parameters(), lr=2e-4)
say("model ok")
tr = parallel.DataParallelTrainer(model, opt, 256, 64, "L1", n_slots=4, exchange=os.environ.get("B2H_EXCHANGE", "auto"))
say("exchange =", tr.exchange)
say("trainer ok")
for s in range(4):
    tr.load(synthetic.model_batch(256, 64, seed=rank * 10 + s), slot=s, non_blocking=False)
for s in range(3):
    l = tr.step(s)
torch.cuda.synchronize(); say("3 eager steps ok, loss", float(l))
t0 = time.time()
for s in range(50):
    tr.step(s % 4)
torch.cuda.synchronize(); say("50 eager steps", (time.time() - t0) / 50 * 1e6, "us/step")
if os.environ.get("B2H_DIAG_GRAPH", "1") == "1":
    say("capturing graph")
    tr.capture(4)
    say("captured")
    tr.replay(); torch.cuda.synchronize(); say("replay ok")
    t0 = time.time()
    for _ in range(25):
        tr.replay()
    torch.cuda.synchronize(); say("graph steps", (time.time() - t0) / 100 * 1e6, "us/step")
from hand_pose_sl_b200 import _lib
lib = _lib.load()
buf = torch.zeros(1024, dtype=torch.int64, device=dev)
for rep in range(2):
    torch.cuda.synchronize(); dist.barrier()
    buf.zero_(); lib.b2h_debug_timing(_lib.ptr(buf))
    tr.step(0); torch.cuda.synchronize()
    lib.b2h_debug_timing(None)
    st = [int(v) for v in buf.cpu().tolist() if v != 0]
    say("stamps", len(st), "total", st[-1] - st[0], "tail deltas", [st[i + 1] - st[i] for i in range(len(st) - 8, len(st) - 1)])
say("dp_status", _lib.load().b2h_dp_status(), "loss", float(tr.loss[0]))
dist.barrier(); say("done")
sys.stdout.flush(); os._exit(0)
