"""GPU tier: tcgen05 descriptor / layout self-test (b2h_tc_probe) against a CPU matmul.  Validates the
no-swizzle K-major layout, the row-shift (implicit im2col) trick, MN-major operands and the M=64 TMEM layout."""
import numpy as np
import pytest
import torch

from hand_pose_sl_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _probe(a, b, n, ksteps, shift, variant):
    lib = _lib.load()
    out = torch.zeros((128, n), dtype=torch.float32, device=DEV)
    A = torch.from_numpy(a).to(DEV).to(torch.bfloat16).contiguous()
    B = torch.from_numpy(b).to(DEV).to(torch.bfloat16).contiguous()
    _lib.check(lib.b2h_tc_probe(_lib.ptr(A), _lib.ptr(B), _lib.ptr(out), n, ksteps, shift, variant, _lib.stream_ptr()))
    torch.cuda.synchronize()
    assert lib.b2h_tc_status() == 0
    return out.cpu().numpy(), A.float().cpu().numpy(), B.float().cpu().numpy()


@pytest.mark.parametrize("n,ksteps,shift", [(32, 1, 0), (32, 2, 0), (32, 2, 3), (48, 2, 4), (64, 4, 1), (16, 1, 2)])
def test_k_major_with_row_shift(n, ksteps, shift):
    rng = np.random.default_rng(n + ksteps + shift)
    K = 16 * ksteps
    a = rng.normal(size=(136, K)).astype(np.float32)
    b = rng.normal(size=(n, K)).astype(np.float32)
    out, A, B = _probe(a, b, n, ksteps, shift, 0)
    want = A[shift:shift + 128] @ B.T
    assert np.abs(out - want).max() <= 1e-3 * np.abs(want).max()


@pytest.mark.parametrize("n,ksteps,shift", [(32, 1, 0), (32, 4, 2), (48, 8, 4)])
def test_mn_major_operands(n, ksteps, shift):
    rng = np.random.default_rng(100 + n + ksteps + shift)
    K = 16 * ksteps
    a = rng.normal(size=(K, 128)).astype(np.float32)
    b = rng.normal(size=(K + 8, n)).astype(np.float32)
    out, A, B = _probe(a, b, n, ksteps, shift, 16)
    want = A.T @ B[shift:shift + K]
    assert np.abs(out - want).max() <= 1e-3 * np.abs(want).max()


def test_m64_tmem_layout():
    """Two M=64 MMAs into lanes 0 and 16: row 16q+i of MMA j lands in TMEM lane 32q+16j+i."""
    rng = np.random.default_rng(7)
    a = rng.normal(size=(136, 32)).astype(np.float32)
    b = rng.normal(size=(32, 32)).astype(np.float32)
    out, A, B = _probe(a, b, 32, 2, 0, 32)
    want = A[:128] @ B.T
    lanes = np.empty_like(want)
    for j in range(2):
        for q in range(4):
            lanes[32 * q + 16 * j: 32 * q + 16 * j + 16] = want[64 * j + 16 * q: 64 * j + 16 * q + 16]
    assert np.abs(out - lanes).max() <= 1e-3 * np.abs(want).max()
