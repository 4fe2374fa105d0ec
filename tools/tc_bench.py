import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hand_pose_sl_b200 import _lib
lib = _lib.load()
out = torch.zeros(2, dtype=torch.int64, device="cuda")
print("M N mn nissue nacc reps | total cycles | cyc per MMA (all warps) | issue-loop cyc/MMA (warp0)")
for mn in (0, 1):
    for M, N in ((64, 32), (128, 32), (128, 64), (128, 128), (64, 48)):
        for nissue in (1, 2, 4):
            nacc = 1
            if N * nacc * nissue > 512: continue
            reps = 64
            for _ in range(2):
                _lib.check(lib.b2h_tc_bench(_lib.ptr(out), M, N, reps, nacc | (nissue << 8), mn, _lib.stream_ptr()))
            torch.cuda.synchronize()
            a, b = out.cpu().tolist()
            print(f"{M:4d} {N:4d} {mn} {nissue} {nacc} {reps} | {a:7d} | {a/(reps*nissue):8.1f} | {b/reps:8.1f}")
