"""Bring-up: per-phase clock64 stamps of CTA 0 of the tensor-core tile kernels (fwd B=512, train B=256)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hand_pose_sl_b200 as b2h
from hand_pose_sl_b200 import _lib, synthetic
from hand_pose_sl_b200.runner import ForwardRunner, TrainStepRunner
dev = torch.device("cuda:0")
lib = _lib.load()
buf = torch.zeros(128, dtype=torch.int64, device=dev)
torch.manual_seed(0)
m = b2h.ConvModel(30, "ReLU", False, precision="bf16").to(dev)
fr = ForwardRunner(m, 512, 64, x_dtype=torch.bfloat16)
fr.x[0].copy_(synthetic.model_batch(512, 64, seed=99)["input_kp"])
opt = b2h.FusedAdam(m.parameters(), lr=2e-4)
tr = TrainStepRunner(m, opt, 256, 64)
tr.load(synthetic.model_batch(256, 64, seed=1234), non_blocking=False)
for name, fn in (("fwd", lambda: fr.run(0)), ("train", lambda: tr.step(0))):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    buf.zero_(); lib.b2h_debug_timing(_lib.ptr(buf))
    fn(); torch.cuda.synchronize()
    lib.b2h_debug_timing(None)
    st = [int(v) for v in buf.cpu().tolist() if v != 0]
    print(name, "stamps:", len(st), "total cycles", st[-1] - st[0])
    print("  deltas:", [st[i + 1] - st[i] for i in range(len(st) - 1)])
