"""GPU bring-up helper: runs b2h_tc_probe for every (mode, variant) in its own subprocess (a faulting
descriptor then cannot poison the others) and prints the max relative error vs a CPU matmul."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r)
from hand_pose_sl_b200 import _lib
mode, variant, n, ks, shift = map(int, sys.argv[1:6])
lib = _lib.load()
rng = np.random.default_rng(0)
K = 16 * ks
if mode == 1:
    a = rng.normal(size=(K, 128)).astype(np.float32); b = rng.normal(size=(K + 8, n)).astype(np.float32)
else:
    a = rng.normal(size=(136, K)).astype(np.float32); b = rng.normal(size=(n, K)).astype(np.float32)
A = torch.from_numpy(a).cuda().to(torch.bfloat16).contiguous(); B = torch.from_numpy(b).cuda().to(torch.bfloat16).contiguous()
out = torch.zeros((128, n), dtype=torch.float32, device="cuda")
rc = lib.b2h_tc_probe(_lib.ptr(A), _lib.ptr(B), _lib.ptr(out), n, ks, shift, (mode << 4) | variant, _lib.stream_ptr())
torch.cuda.synchronize()
st = lib.b2h_tc_status()
Af, Bf, o = A.float().cpu().numpy(), B.float().cpu().numpy(), out.cpu().numpy()
if mode == 1:
    want = Af.T @ Bf[shift:shift + K]
else:
    want = Af[shift:shift + 128] @ Bf.T
if mode == 2:
    lanes = np.zeros_like(want)
    for j in range(2):
        for q in range(4):
            lanes[32*q+16*j:32*q+16*j+16] = want[64*j+16*q:64*j+16*q+16]
    err = np.abs(o - lanes).max() / np.abs(want).max()
    # describe where each expected row actually landed
    where = []
    for r in range(0, 128, 16):
        d = np.abs(o[:, None, :] - want[None, r:r+1, :]).max(axis=2)[:, 0]
        where.append((r, int(d.argmin()), float(d.min())))
    print("rows->lanes", where)
else:
    err = np.abs(o - want).max() / np.abs(want).max()
print(f"mode={mode} variant={variant} n={n} ks={ks} shift={shift} rc={rc} status={st} relerr={err:.3e}")
''' % ROOT

if __name__ == "__main__":
    for mode in (0, 1, 2):
        for variant in (0, 1, 2, 3):
            for (n, ks, shift) in ((32, 2, 0), (32, 2, 3)):
                try:
                    r = subprocess.run([sys.executable, "-c", CHILD, str(mode), str(variant), str(n), str(ks), str(shift)],
                                       capture_output=True, text=True, timeout=120)
                    tail = (r.stdout.strip().splitlines() or ["<no output>"])
                    print("\n".join(tail[-2:]), "| exit", r.returncode, "|", r.stderr.strip().splitlines()[-1] if r.returncode else "")
                except subprocess.TimeoutExpired:
                    print(f"mode={mode} variant={variant} n={n} ks={ks} shift={shift} TIMEOUT")
                sys.stdout.flush()
