"""Bring-up: clock64 phase stamps of CTA 0 of the wide forward kernel (MMA thread: slots 0..59, epilogue thread 0: 64..)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hand_pose_sl_b200 as b2h
from hand_pose_sl_b200 import _lib, synthetic
from hand_pose_sl_b200.runner import ForwardRunner
dev = torch.device("cuda:0")
lib = _lib.load()
C = int(sys.argv[1]) if len(sys.argv) > 1 else 256
B, T = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (444, 64)
buf = torch.zeros(256, dtype=torch.int64, device=dev)
m = b2h.ConvModel(C, "ReLU", False, precision="bf16").to(dev)
fr = ForwardRunner(m, B, T, x_dtype=torch.bfloat16)
fr.x[0].copy_(synthetic.model_batch(B, T, seed=99)["input_kp"])
for _ in range(3): fr.run(0)
torch.cuda.synchronize()
buf.zero_(); lib.b2h_debug_timing(_lib.ptr(buf))
fr.run(0); torch.cuda.synchronize()
lib.b2h_debug_timing(None)
st = buf.cpu().tolist()
mma = [v for v in st[:8] if v]; epi = [v for v in st[64:128] if v]
t0 = mma[0]
print("MMA thread  (ready, issued) per layer:", [(mma[i] - t0, mma[i + 1] - t0) for i in range(0, len(mma) - 1, 2)])
print("epilogue t0 (acc_full, published) per layer:", [(epi[i] - t0, epi[i + 1] - t0) for i in range(0, len(epi) - 1, 2)])
