"""GPU tier: ConvModel forward through the C-ABI vs the oracle and the reference's golden vectors.
fp32 mode <= 1e-4 relative, bf16 mode <= 2e-2 relative (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

import b2h_oracle as oracle
import hand_pose_sl_b200 as b2h
from conftest import golden_sd, load_golden
from hand_pose_sl_b200 import _lib, synthetic

pytestmark = pytest.mark.gpu
TOL = {"fp32": 1e-4, "fp32-ffma": 1e-4, "bf16": 2e-2}
DEV = "cuda:0"


def _model(sd, C, pe, prec):
    m = b2h.ConvModel(C, "ReLU", pe, precision=prec)
    m.load_state_dict(sd)
    return m.to(DEV)


def _tc_clean():
    torch.cuda.synchronize()
    assert _lib.load().b2h_tc_status() == 0, "tensor-core kernel hit a wait timeout"


@pytest.mark.parametrize("prec", ["fp32", "fp32-ffma", "bf16"])
@pytest.mark.parametrize("name", ["convmodel_c30_b1.npz", "convmodel_c30.npz", "convmodel_c30_posemb.npz",
                                  "convmodel_c64.npz", "convmodel_c30_t200.npz"])
def test_forward_golden(name, prec):
    g = load_golden(name)
    m = _model(golden_sd(g), int(g["C"]), bool(g["pos_emb"]), prec)
    x = torch.from_numpy(g["input_kp"]).to(DEV)
    with torch.no_grad():
        y = m(x)
    _tc_clean()
    assert y.shape == g["pred"].shape and y.dtype == torch.float32
    assert oracle.rel_err(y.cpu().numpy(), g["pred"]) <= TOL[prec]
    # fused mask_output epilogue == reference mask_output on the reference prediction
    ym = m.predict(x, lengths=torch.from_numpy(g["lengths"]))
    _tc_clean()
    assert oracle.rel_err(ym.cpu().numpy(), g["pred_masked"]) <= TOL[prec]
    for i, ln in enumerate(g["lengths"]):
        assert not ym[i, int(ln):].any()                 # exact zeros (index work: bit-exact)
    # fused de-normalise (traintest.py:270-271)
    yd = m.predict(x, denormalize=1280)
    assert torch.equal(yd, y * 1280)


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("B,T,C", [(1, 64, 30), (7, 64, 30), (5, 1, 30), (3, 5, 30), (2, 130, 30), (9, 33, 16), (2, 64, 64),
                                   (300, 64, 30), (3, 127, 30), (2, 128, 30), (3, 129, 30), (5, 200, 30), (2, 256, 30),
                                   (2, 257, 30), (2, 200, 64), (150, 200, 30)])
def test_forward_vs_oracle_shapes(B, T, C, prec):
    sd = oracle.init_params(C, False, seed=B + T)
    batch = synthetic.model_batch(B, T, seed=B * 1000 + T)
    m = _model(sd, C, False, prec)
    with torch.no_grad():
        y = m(batch["input_kp"].to(DEV))
    _tc_clean()
    ref = oracle.conv_model_forward(sd, batch["input_kp"]).contiguous().numpy()
    assert oracle.rel_err(y.cpu().numpy(), ref) <= TOL[prec]


@pytest.mark.parametrize("B,T,C", [(4, 64, 256), (7, 64, 128), (5, 64, 80), (3, 126, 256), (2, 200, 256), (3, 256, 96),
                                   (450, 64, 256)])
def test_wide_forward_vs_oracle(B, T, C):
    """conv_channels > 64 in bf16 mode: the streamed-weight kernel (256-row tiles, weight ring; 450 windows of 64 frames
    = 150 tiles > 148 CTAs, so the ring and the phases wrap across tiles).  Also with mask_output + de-normalisation."""
    sd = oracle.init_params(C, False, seed=B + T)
    batch = synthetic.model_batch(B, T, seed=B * 1000 + T, ragged=True, len_seed=C)
    m = _model(sd, C, False, "bf16")
    with torch.no_grad():
        y = m(batch["input_kp"].to(DEV))
        ym = m.predict(batch["input_kp"].to(DEV), lengths=batch["n_frames"], denormalize=1280.0)
    _tc_clean()
    ref = oracle.conv_model_forward(sd, batch["input_kp"]).contiguous()
    assert oracle.rel_err(y.cpu().numpy(), ref.numpy()) <= TOL["bf16"]
    refm = oracle.mask_output(ref.clone(), batch["n_frames"]) * 1280.0
    assert oracle.rel_err(ym.cpu().numpy(), refm.numpy()) <= TOL["bf16"]
    lens = batch["n_frames"]
    for i in range(min(B, 8)):
        assert float(ym[i, int(lens[i]):].abs().max() if int(lens[i]) < T else 0.0) == 0.0


def test_wide_forward_golden():
    """conv_channels = 256 against the reference's own outputs (tests/golden/convmodel_c256_fwd.npz; the weights are
    the reference's seeded default init, reproduced by oracle.init_params and pinned by checksums on the CPU tier)."""
    g = load_golden("convmodel_c256_fwd.npz")
    sd = oracle.init_params(int(g["C"]), False, seed=int(g["seed"]))
    m = _model(sd, int(g["C"]), False, "bf16")
    x = torch.from_numpy(g["input_kp"]).to(DEV)
    with torch.no_grad():
        y = m(x)
        ym = m.predict(x, lengths=torch.from_numpy(g["lengths"]))
    _tc_clean()
    assert oracle.rel_err(y.cpu().numpy(), g["pred"]) <= TOL["bf16"]
    assert oracle.rel_err(ym.cpu().numpy(), g["pred_masked"]) <= TOL["bf16"]


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_bf16_input_tensor(prec):
    sd = oracle.init_params(30, False, seed=0)
    batch = synthetic.model_batch(8, 64, seed=5)
    m = _model(sd, 30, False, prec)
    xb = batch["input_kp"].to(DEV).to(torch.bfloat16)
    with torch.no_grad():
        y = m(xb)
    _tc_clean()
    ref = oracle.conv_model_forward(sd, xb.float().cpu()).contiguous().numpy()
    assert oracle.rel_err(y.cpu().numpy(), ref) <= TOL[prec]


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_config2_full_size_properties(prec):
    """BASELINE config 2 (B=512 x 64): window independence (any sub-batch gives the same rows), determinism,
    and a seeded sample of windows against the oracle."""
    sd = oracle.init_params(30, False, seed=0)
    batch = synthetic.model_batch(512, 64, seed=1234)
    m = _model(sd, 30, False, prec)
    x = batch["input_kp"].to(DEV)
    with torch.no_grad():
        y = m(x)
        y2 = m(x)
        part = m(x[100:229])
        perm = torch.randperm(512, generator=torch.Generator().manual_seed(0)).to(DEV)
        yp = m(x[perm])
    _tc_clean()
    assert torch.equal(y, y2)
    assert torch.equal(part, y[100:229])
    assert torch.equal(yp, y[perm])
    pick = [0, 1, 63, 64, 255, 256, 300, 510, 511]
    ref = oracle.conv_model_forward(sd, batch["input_kp"][pick]).contiguous().numpy()
    assert oracle.rel_err(y[pick].cpu().numpy(), ref) <= TOL[prec]


def test_errors_mirror_reference():
    m = b2h.ConvModel(30, "ReLU", False).to(DEV)
    with pytest.raises(RuntimeError):                       # 8 keypoints -> 16 != 24 channels
        m(torch.zeros(1, 64, 8, 2, device=DEV))
    mp = b2h.ConvModel(30, "ReLU", True).to(DEV)
    with pytest.raises(RuntimeError):                       # pos_emb only works for T == 100 (HandPoseModels.py:82)
        mp(torch.zeros(1, 64, 12, 2, device=DEV))


@pytest.mark.parametrize("stride", [64, 16])
def test_streaming_pipeline_k0_to_k1(stride):
    """BASELINE config 5: raw OpenPose clip -> K0 (bit-exact preprocessing, bf16 copy for the net) -> K1 forward over
    sliding windows, against the oracle pipeline (reference transforms + reference ConvModel) on the same clip."""
    F, T = 1000, 64
    pose, lh, rh = synthetic.synthetic_clip(F, seed=21)
    starts = b2h.sliding_window_starts(F, T, stride)
    want_item = oracle.preprocess_windows(pose, lh, rh, starts, T)
    sd = oracle.init_params(30, False, seed=0)
    ref = oracle.conv_model_forward(sd, torch.from_numpy(want_item["input_kp"])).contiguous().numpy() * np.float32(1280)
    pre = b2h.PreprocessRightHand(emit_bf16=True)
    out = pre(torch.from_numpy(pose).to(DEV), torch.from_numpy(lh).to(DEV), torch.from_numpy(rh).to(DEV), starts, T)
    assert np.array_equal(out["input_kp"].cpu().numpy(), want_item["input_kp"])          # K0 stays bit-exact
    for prec, x in (("bf16", out["input_kp_bf16"]), ("fp32", out["input_kp"])):
        m = _model(sd, 30, False, prec)
        y = m.predict(x, denormalize=1280)                                                 # traintest.py:270-271
        _tc_clean()
        assert oracle.rel_err(y.cpu().numpy(), ref) <= TOL[prec]


def test_inference_output_formats():
    """de-normalised prediction -> OpenPose rows / packed H5 rows, bit-exact vs the reference writers' layout."""
    g = load_golden("convmodel_c30.npz")
    m = _model(golden_sd(g), 30, False, "fp32")
    y = m.predict(torch.from_numpy(g["input_kp"]).to(DEV), denormalize=1280)
    op = b2h.format_prediction(y, "openpose").cpu().numpy()
    h5 = b2h.format_prediction(y, "h5").cpu().numpy()
    yc = y.cpu().numpy()
    for b in range(yc.shape[0]):
        assert np.array_equal(h5[b], oracle.order_and_reshape_toh5(yc[b]))
        for t_ in (0, 17, yc.shape[1] - 1):
            assert np.array_equal(op[b, t_], oracle.array2open_pose(yc[b, t_]))


def test_linear_positional_embedding_forward_matches_reference_semantics():
    """LinearPositionalEmbedding.forward(inp, lengths) (HandPoseModels.py:78-84): cat([t/100, inp], dim=1), bit-exact;
    T != max_len raises like the reference's torch.cat."""
    import hand_pose_sl_b200 as b2h
    pe = b2h.LinearPositionalEmbedding(max_len=100)
    inp = torch.randn(3, 24, 100, generator=torch.Generator().manual_seed(0))
    want = torch.cat([torch.cat(3 * [(torch.arange(100)[None, None, :].float() / 100)], dim=0), inp], dim=1)
    got = pe(inp.to("cuda:0"), None)
    assert got.shape == (3, 25, 100) and torch.equal(got.cpu(), want)
    with pytest.raises(RuntimeError):
        pe(torch.zeros(2, 24, 64, device="cuda:0"), None)


@pytest.mark.parametrize("pad_mode", ["repeat_first", "zeros"])
@pytest.mark.parametrize("stride,T", [(64, 64), (16, 64), (7, 37), (50, 200)])
def test_streaming_window_views_equal_materialised_windows(stride, T, pad_mode):
    """BASELINE config 5 without re-materialisation: K0 once per unique frame -> (F,12,2) stream; the forward reads
    sliding windows as VIEWS of it (crop + pad rule on the fly).  Bit-identical to the forward over the windows K0
    materialises, and within tolerance of the oracle pipeline; the stream itself is the bf16 rounding of the bit-exact
    reference input."""
    F = 1000
    pose, lh, rh = synthetic.synthetic_clip(F, seed=21)
    starts = b2h.sliding_window_starts(F, T, stride)
    tp, tl, tr = (torch.from_numpy(a).to(DEV) for a in (pose, lh, rh))
    pre = b2h.PreprocessRightHand(emit_bf16=True, pad_mode=pad_mode)
    mat = pre(tp, tl, tr, starts, T)                                                        # materialised windows (reference item)
    stream = pre.frame_stream(tp, tl, tr)                                                   # (F,12,2) bf16, one row per unique frame
    one = pre(tp, tl, tr, np.zeros(1, dtype=np.int64), F)
    assert torch.equal(stream, one["input_kp_bf16"].view(F, 12, 2))
    assert torch.equal(pre.frame_stream(tp, tl, tr, dtype=torch.float32), one["input_kp"].view(F, 12, 2))
    sd = oracle.init_params(30, False, seed=0)
    m = _model(sd, 30, False, "bf16")
    y_view = m.predict_windows(stream, starts, T, pad_mode=pad_mode, denormalize=1280)
    y_mat = m.predict(mat["input_kp_bf16"], denormalize=1280)
    _tc_clean()
    assert torch.equal(y_view, y_mat)
    want_item = oracle.preprocess_windows(pose, lh, rh, starts, T,
                                          pad_mode=oracle.PAD_REPEAT_FIRST if pad_mode == "repeat_first" else oracle.PAD_ZEROS)
    ref = oracle.conv_model_forward(sd, torch.from_numpy(want_item["input_kp"])).contiguous().numpy() * np.float32(1280)
    assert oracle.rel_err(y_view.cpu().numpy(), ref) <= TOL["bf16"]
    # utterance ends inside the stream: windows cut at win_end
    ends = np.minimum(starts + 45, F).astype(np.int64)
    y_cut = m.predict_windows(stream, starts, T, win_end=ends, pad_mode=pad_mode)
    mat_cut = pre(tp, tl, tr, starts, T, win_end=ends)
    assert torch.equal(y_cut, m.predict(mat_cut["input_kp_bf16"]))
    # fp32 mode (split-operand tensor-core kernel) serves window views too; shapes it does not cover must be materialised
    m32 = _model(sd, 30, False, "fp32")
    s32 = pre.frame_stream(tp, tl, tr, dtype=torch.float32)
    assert torch.equal(m32.predict_windows(s32, starts, T, pad_mode=pad_mode), m32.predict(mat["input_kp"]))
    with pytest.raises(_lib.B2HError):
        _model(oracle.init_params(64, False, seed=0), 64, False, "fp32").predict_windows(s32, starts, T)
