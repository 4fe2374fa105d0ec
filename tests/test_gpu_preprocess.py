"""GPU tier: K0 preprocessing through the C-ABI vs the oracle and the reference's golden vectors.
Bit-exact (indexing, windowing, sub.rn, div.rn)."""
import numpy as np
import pytest
import torch

import b2h_oracle as oracle
import hand_pose_sl_b200 as b2h
from conftest import load_golden
from hand_pose_sl_b200 import _lib, synthetic

pytestmark = pytest.mark.gpu
KEYS = ("input_kp", "input_conf", "target_kp", "target_conf", "left_hand_kp", "left_hand_conf", "n_frames")


def _run(pose, lh, rh, starts, T, **kw):
    dev = torch.device("cuda:0")
    pre = b2h.PreprocessRightHand(**kw)
    out = pre(torch.from_numpy(pose).to(dev), torch.from_numpy(lh).to(dev), torch.from_numpy(rh).to(dev), starts, T)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items() if v.dtype != torch.bfloat16}


@pytest.mark.parametrize("tag,dif", [("dif", True), ("nodif", False)])
def test_golden_reference_vectors(tag, dif):
    g = load_golden("preprocess.npz")
    out = _run(g["pose25"], g["hand_left"], g["hand_right"], g["win_start"], int(g["T"]), dif_encoding=dif)
    for k in KEYS:
        assert np.array_equal(out[k], g[f"{tag}_{k}"]), k
    assert out["body_kp"] is not None and np.array_equal(out["body_kp"], out["input_kp"])   # BuildRightHandItem aliases


@pytest.mark.parametrize("F,T,stride,pad", [(1000, 64, 64, "repeat_first"), (1000, 64, 16, "repeat_first"),
                                            (333, 100, 50, "zeros"), (61, 64, 64, "repeat_first"), (257, 63, 7, "zeros"),
                                            (1, 64, 64, "repeat_first"), (4096, 200, 200, "repeat_first")])
def test_against_oracle_windows(F, T, stride, pad):
    pose, lh, rh = synthetic.synthetic_clip(F, seed=F + T)
    starts = b2h.sliding_window_starts(F, T, stride)
    out = _run(pose, lh, rh, starts, T, pad_mode=pad)
    want = oracle.preprocess_windows(pose, lh, rh, starts, T, oracle.PAD_REPEAT_FIRST if pad == "repeat_first" else oracle.PAD_ZEROS)
    for k in KEYS:
        assert np.array_equal(out[k], want[k]), (k, F, T, stride)


def test_random_crops_and_unaligned_starts():
    pose, lh, rh = synthetic.synthetic_clip(500, seed=3)
    rng = np.random.default_rng(0)
    starts = rng.integers(0, 500, size=37).astype(np.int64)
    for T in (64, 30):
        out = _run(pose, lh, rh, starts, T)
        want = oracle.preprocess_windows(pose, lh, rh, starts, T)
        for k in KEYS:
            assert np.array_equal(out[k], want[k]), (k, T)


def test_no_normalize_no_left_hand_and_bf16_copy():
    pose, lh, rh = synthetic.synthetic_clip(256, seed=9)
    starts = np.array([0, 64, 128, 192], dtype=np.int64)
    out = _run(pose, lh, rh, starts, 64, normalize=False, with_left_hand=False, emit_bf16=True)
    want = oracle.preprocess_windows(pose, lh, rh, starts, 64, normalize=False)
    for k in ("input_kp", "input_conf", "target_kp", "target_conf", "n_frames"):
        assert np.array_equal(out[k], want[k]), k
    assert "left_hand_kp" not in out
    dev = torch.device("cuda:0")
    pre = b2h.PreprocessRightHand(emit_bf16=True)
    o = pre(torch.from_numpy(pose).to(dev), torch.from_numpy(lh).to(dev), torch.from_numpy(rh).to(dev), starts, 64)
    assert torch.equal(o["input_kp_bf16"].float().cpu(), o["input_kp"].cpu().to(torch.bfloat16).float())


def test_special_values_take_the_ieee_division_path():
    """-0, denormals, tiny / huge magnitudes, inf and NaN inside a 4-frame group: the kernel's range test sends the whole
    float4 through div.rn, everything else through the verified fast division -- both must equal the CPU's x / 1280."""
    F = 256
    pose, lh, rh = synthetic.synthetic_clip(F, seed=21)
    specials = np.array([-0.0, 0.0, 1e-45, -1e-45, 1e-38, 3e-31, -9e-31, 1.1e-30, 9.9e29, -1.1e30, 3e38, -3.4e38,
                         np.inf, -np.inf, np.nan, 1280.0, -1280.0, 1e-30, 1e30], dtype=np.float32)
    rng = np.random.default_rng(5)
    for arr in (pose, lh, rh):
        flat = arr.reshape(F, -1)
        for f in range(0, F, 3):                       # a few specials in most groups, on x / y / confidence alike
            cols = rng.integers(0, flat.shape[1], size=4)
            flat[f, cols] = rng.choice(specials, size=4)
    starts = np.array([0, 3, 64, 129, 190], dtype=np.int64)
    for dif in (True, False):
        out = _run(pose, lh, rh, starts, 64, dif_encoding=dif)
        want = oracle.preprocess_windows(pose, lh, rh, starts, 64, dif_encoding=dif)
        for k in KEYS:
            a, b = out[k], want[k]
            same = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b)) if a.dtype == np.float32 else (a == b)
            assert same.all(), (k, dif, int((~same).sum()))


@pytest.mark.parametrize("T", [1, 2, 3, 5, 7])
def test_tiny_windows_cross_group_boundaries(T):
    """Windows shorter than / not a multiple of the 4-frame group: every group mixes windows (per-slot walk)."""
    pose, lh, rh = synthetic.synthetic_clip(97, seed=T)
    starts = np.array([0, 4, 8, 93, 96, 40, 41, 42, 43], dtype=np.int64)
    for pad in ("repeat_first", "zeros"):
        out = _run(pose, lh, rh, starts, T, pad_mode=pad)
        want = oracle.preprocess_windows(pose, lh, rh, starts, T, oracle.PAD_REPEAT_FIRST if pad == "repeat_first" else oracle.PAD_ZEROS)
        for k in KEYS:
            assert np.array_equal(out[k], want[k]), (k, T, pad)


@pytest.mark.parametrize("tag", ["h5short", "h5long"])
def test_h5_rows_golden(tag):
    g = load_golden("preprocess.npz")
    T = int(g["T"])
    dev = torch.device("cuda:0")
    pre = b2h.PreprocessRightHand(pad_mode="zeros")
    out = pre.from_h5_rows(torch.from_numpy(g[f"{tag}_array"]).to(dev), np.array([0], dtype=np.int64), T)
    assert int(out["n_frames"][0]) == int(g[f"{tag}_n_frames"])
    for k in ("input_kp", "input_conf", "target_kp", "target_conf", "left_hand_kp", "left_hand_conf"):
        assert np.array_equal(out[k][0].cpu().numpy(), g[f"{tag}_{k}"]), k


def test_full_size_stream_properties():
    """One hour of 30 fps frames (BASELINE config 5): size-independent properties instead of a full oracle pass --
    windows are pure gathers of the per-frame transform (stride-16 windows == slices of the stride-1 stream),
    the chest row is exactly 0, confidences pass through untouched."""
    F = 108000
    pose, lh, rh = synthetic.synthetic_clip(F, seed=1234)
    dev = torch.device("cuda:0")
    tp, tl, tr = (torch.from_numpy(a).to(dev) for a in (pose, lh, rh))
    pre = b2h.PreprocessRightHand(with_left_hand=False)
    frames = pre(tp, tl, tr, np.array([0], dtype=np.int64), F)              # T = F: the per-frame stream
    starts = b2h.sliding_window_starts(F - 64, 64, 16)
    wins = pre(tp, tl, tr, starts, 64)
    idx = torch.from_numpy(starts).to(dev)[:, None] + torch.arange(64, device=dev)[None, :]
    assert torch.equal(wins["input_kp"], frames["input_kp"][0][idx])
    assert torch.equal(wins["target_kp"], frames["target_kp"][0][idx])
    assert not frames["input_kp"][0, :, 1, :].any()
    assert torch.equal(frames["target_conf"][0], tr[:, :, 2])
    body = torch.tensor(b2h.BODY_HEAD_KEYPOINTS, device=dev)
    assert torch.equal(frames["input_conf"][0], tp[:, body, 2])
    # spot-check 2000 random frames against the oracle
    rng = np.random.default_rng(1)
    pick = np.sort(rng.choice(F, size=2000, replace=False))
    want = oracle.preprocess_windows(pose[pick], lh[pick], rh[pick], np.array([0]), 2000)
    assert np.array_equal(frames["input_kp"][0][torch.from_numpy(pick).to(dev)].cpu().numpy(), want["input_kp"][0])
    assert np.array_equal(frames["target_kp"][0][torch.from_numpy(pick).to(dev)].cpu().numpy(), want["target_kp"][0])


def test_fast_division_is_exact_for_all_floats():
    """The kernel divides by 1280 with reciprocal + two FMAs instead of the generic div.rn routine; the identity is
    checked against IEEE div.rn for every one of the 2^32 float bit patterns (NaNs compare equal)."""
    from hand_pose_sl_b200 import _lib
    lib = _lib.load()
    bad = torch.zeros(1, dtype=torch.int64, device="cuda:0")
    _lib.check(lib.b2h_verify_fastdiv(1280.0, _lib.ptr(bad), _lib.stream_ptr()))
    torch.cuda.synchronize()
    assert int(bad.item()) == 0


def test_preprocess_status_word_is_clean_after_the_suite():
    """The staged-group wait of K0 is a bounded spin that records a timeout instead of breaking silently."""
    import torch
    torch.cuda.synchronize()
    assert _lib.load().b2h_preprocess_status() == 0
