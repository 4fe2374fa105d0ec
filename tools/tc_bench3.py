"""MMA cost for the wgrad shapes: M=128 / 64, N = 32..256, K-major vs MN-major operands, rotating accumulators."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hand_pose_sl_b200 import _lib
lib = _lib.load()
out = torch.zeros(2, dtype=torch.int64, device="cuda")
print("M N mn nacc | cyc per MMA")
for mn in (0, 1):
    for M, N, nacc in ((128, 32, 5), (128, 64, 5), (128, 96, 5), (128, 128, 4), (128, 256, 2), (64, 64, 5), (64, 32, 6), (128, 64, 1), (128, 128, 1)):
        reps = 80
        for _ in range(2):
            _lib.check(lib.b2h_tc_bench(_lib.ptr(out), M, N, reps, nacc | (1 << 8), mn | 16, _lib.stream_ptr()))
        torch.cuda.synchronize()
        a, b = out.cpu().tolist()
        print(f"{M:4d} {N:4d} {mn} {nacc} | {a/reps:8.1f}")
