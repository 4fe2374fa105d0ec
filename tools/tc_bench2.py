"""MMA pacing at wide N: cycles per tcgen05.mma (M=128, K=16) for N=256/128/48, no-swizzle vs 128B-swizzle descriptors
(timing only), one CTA vs one CTA on every SM."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hand_pose_sl_b200 import _lib
lib = _lib.load()
out = torch.zeros(2, dtype=torch.int64, device="cuda")
print("flags: 32 = B chunk stride 4096 B, 64 = A chunk stride 4096 B, 128 = B chunk stride 4224 B")
print("flags: 4 = vary A/B addresses, 8 = commit every 2nd MMA, 16 = non-zero operand data")
print("M N flags grid nacc reps | cyc per MMA | issue-loop cyc/MMA")
for N, nacc in ((256, 2), (48, 2)):
    for swz in (0, 32, 64, 96, 128, 32 + 12, 128 + 12):      # bit 2: vary operands, bit 3: commit every 2 MMAs, bit 4: non-zero data
        for grid in (1,):
            reps = 400
            for _ in range(2):
                _lib.check(lib.b2h_tc_bench(_lib.ptr(out), 128, N, reps, nacc | (1 << 8), swz | (grid << 8), _lib.stream_ptr()))
            torch.cuda.synchronize()
            a, b = out.cpu().tolist()
            print(f"128 {N:4d} {swz} {grid:4d} {nacc} {reps} | {a/reps:8.1f} | {b/reps:8.1f}", flush=True)
