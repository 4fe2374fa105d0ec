// extern "C" surface of libb2h.so (see include/b2h.h).  Argument checking, error reporting and
// dispatch between the fp32 (FFMA) and bf16 (tcgen05) kernels.  No CPU fallback exists anywhere.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <cmath>

#include "b2h_common.cuh"

namespace b2h {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return B2H_ECUDA;
  }
  return B2H_OK;
}
void count_launch(int n) { g_launches.fetch_add(n); }

static bool geo_ok(int n_in, int C, int pos_emb, const char* who) {
  if (n_in < 1 || n_in > 64 || C < 1 || C > B2H_MAX_C || pos_emb < 0 || pos_emb > 65536) {
    set_error("%s: unsupported geometry n_in=%d C=%d pos_emb=%d", who, n_in, C, pos_emb);
    return false;
  }
  return true;
}

// precision -> is it a tensor-core mode, and with bf16 high/low operand pairs (fp32 mode) or plain bf16 operands
static bool is_split(int precision) { return precision == B2H_FP32; }
static bool prec_ok(int precision) { return precision == B2H_FP32 || precision == B2H_BF16 || precision == B2H_FP32_FFMA; }
static bool use_tc_train(const Geo& g, int T, int precision) {
  return (precision == B2H_BF16 || precision == B2H_FP32) && tc_tile_ok(g, T, true, is_split(precision));
}

// Which kernel b2h_conv_forward launches (the ONE place that decides; b2h_kernel_choice reports it).
static int forward_choice(const Geo& g, int T, int precision) {
  const int ffma = fp32_smem_bytes(g, T, false) <= (size_t)226 * 1024 ? B2H_KERNEL_FFMA : B2H_KERNEL_NONE;
  if (precision == B2H_FP32_FFMA) return ffma;
  if (precision == B2H_FP32) return tc_tile_ok(g, T, false, true) ? B2H_KERNEL_TC_TILE : ffma;   // split tile kernel: C <= 32, T <= 256
  if (precision != B2H_BF16) return B2H_KERNEL_NONE;
  if (tc_tile_ok(g, T, false, false)) return B2H_KERNEL_TC_TILE; // independent 128/256-row tiles (T <= 256, C <= 64)
  if (tc_fwd_supported(g, T)) return B2H_KERNEL_TC_ROWSPACE;     // long windows (T <= 1024, C <= 64): layer-major row space
  if (tc_wide_supported(g, T)) return B2H_KERNEL_TC_WIDE;        // wide models (C <= 256, T <= 256): streamed weights
  return B2H_KERNEL_NONE;
}

// Which training path serves (B, T, C, precision) and how its workspace is laid out: the ONE place that decides.
//   [header B2H_WS_HEADER][gradient partials: nparts x stride floats][loss partials: n_loss floats][scratch (wide path)]
struct TrainPlan {
  int kernel;            // B2H_KERNEL_TC_TILE | B2H_KERNEL_TC_WIDE_TRAIN | B2H_KERNEL_FFMA | B2H_KERNEL_NONE
  int nparts, gp_layout, n_loss;
  int64_t stride, scratch_off, total;
};
static TrainPlan train_plan(const Geo& g, int B, int T, int precision) {
  TrainPlan t{};
  t.kernel = B2H_KERNEL_NONE;
  int64_t scratch = 0;
  if (use_tc_train(g, T, precision)) {
    t.kernel = B2H_KERNEL_TC_TILE; t.nparts = tc_train_grid(g, B, T, is_split(precision)); t.stride = gp_total(g); t.gp_layout = 1; t.n_loss = t.nparts;
  } else if (precision == B2H_BF16 && tc_wide_supported(g, T)) {
    // 32 < conv_channels <= 256: forward+criterion / dgrad chain / split-K wgrad kernels over scratch dumps
    t.kernel = B2H_KERNEL_TC_WIDE_TRAIN; t.nparts = tc_wide_train_ksplit(g, B, T); t.stride = gp_total(g); t.gp_layout = 1;
    t.n_loss = tc_wide_train_loss_parts(g, B, T);
    scratch = tc_wide_train_scratch_bytes(g, B, T);
  } else if (fp32_smem_bytes(g, T, true) <= (size_t)226 * 1024) {
    t.kernel = B2H_KERNEL_FFMA; t.nparts = fp32_train_grid(g, B, T); t.stride = g.P; t.gp_layout = 0; t.n_loss = t.nparts;
  }
  int64_t o = B2H_WS_HEADER + ((int64_t)t.nparts * t.stride + t.n_loss) * 4;
  o = (o + 255) / 256 * 256;
  t.scratch_off = o;
  t.total = o + scratch + 256;
  return t;
}

}  // namespace b2h

using namespace b2h;

extern "C" const char* b2h_last_error(void) { return g_err; }
extern "C" int b2h_version(void) { return 100; }
extern "C" int64_t b2h_launch_count(void) { return g_launches.load(); }

extern "C" int b2h_device_ok(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  return major == 10 ? 1 : 0;
}

extern "C" int64_t b2h_param_count(int n_in, int C, int pos_emb) {
  if (!geo_ok(n_in, C, pos_emb, "b2h_param_count")) return B2H_ESHAPE;
  return make_geo(n_in, C, pos_emb).P;
}
extern "C" int64_t b2h_param_offset(int n_in, int C, int pos_emb, int layer, int is_bias) {
  if (!geo_ok(n_in, C, pos_emb, "b2h_param_offset")) return B2H_ESHAPE;
  if (layer < 1 || layer > 4) { set_error("b2h_param_offset: layer %d", layer); return B2H_EINVAL; }
  Geo g = make_geo(n_in, C, pos_emb);
  return is_bias ? g.b_off[layer - 1] : g.w_off[layer - 1];
}
/* Host-side self-check of the gradient-partial slot layout (no GPU): every flat parameter index maps to a distinct
 * slot and back, every other slot decodes as padding.  Returns the number of violations (0 = consistent). */
extern "C" int64_t b2h_gp_layout_check(int n_in, int C, int pos_emb) {
  if (!geo_ok(n_in, C, pos_emb, "b2h_gp_layout_check")) return B2H_ESHAPE;
  Geo g = make_geo(n_in, C, pos_emb);
  const int nj = gp_total(g);
  int64_t bad = (nj & 3) ? 1 : 0;
  for (int l = 0; l < 4; ++l) bad += (gp_layer_off(g, l) & 3) ? 1 : 0;
  int64_t real = 0;
  for (int j = 0; j < nj; ++j) {
    const int i = flat_index_of_gp(g, j);
    if (i < 0) continue;
    ++real;
    if (i >= g.P || gp_index_of_flat(g, i) != j) ++bad;
  }
  if (real != g.P) ++bad;
  for (int i = 0; i < g.P; ++i) {
    const int j = gp_index_of_flat(g, i);
    if (j < 0 || j >= nj || flat_index_of_gp(g, j) != i) ++bad;
  }
  return bad;
}
extern "C" int64_t b2h_packed_bytes(int n_in, int C, int pos_emb) {
  if (!geo_ok(n_in, C, pos_emb, "b2h_packed_bytes")) return B2H_ESHAPE;
  return make_geo(n_in, C, pos_emb).packed_bytes;
}
extern "C" int b2h_supported(int T, int n_in, int C, int pos_emb, int precision) {
  if (n_in < 1 || n_in > 64 || C < 1 || C > B2H_MAX_C || T < 1 || pos_emb < 0 || pos_emb > 65536) return 0;
  Geo g = make_geo(n_in, C, pos_emb);
  if (!prec_ok(precision)) return 0;
  return (forward_choice(g, T, precision) != B2H_KERNEL_NONE && train_plan(g, 1, T, precision).kernel != B2H_KERNEL_NONE) ? 1 : 0;
}
extern "C" int b2h_forward_supported(int T, int n_in, int C, int pos_emb, int precision) {
  if (n_in < 1 || n_in > 64 || C < 1 || C > B2H_MAX_C || T < 1 || pos_emb < 0 || pos_emb > 65536) return 0;
  return forward_choice(make_geo(n_in, C, pos_emb), T, precision) != B2H_KERNEL_NONE ? 1 : 0;
}
extern "C" int b2h_kernel_choice(int T, int n_in, int C, int pos_emb, int precision, int train) {
  if (n_in < 1 || n_in > 64 || C < 1 || C > B2H_MAX_C || T < 1 || pos_emb < 0 || pos_emb > 65536) return B2H_KERNEL_NONE;
  Geo g = make_geo(n_in, C, pos_emb);
  if (!train) return forward_choice(g, T, precision);
  if (!prec_ok(precision)) return B2H_KERNEL_NONE;
  return train_plan(g, 1, T, precision).kernel;
}
/* How the tile kernel's training path cuts a window of T frames (host-only query): returns n_sub (1 = whole windows) and,
 * for sub-window i, writes {first frame, first core row, end of the core rows (sub-window coordinates), sub-window length}. */
extern "C" int b2h_train_subwindows(int T, int n_in, int C, int pos_emb, int precision, int i, int* out4) {
  if (!geo_ok(n_in, C, pos_emb, "b2h_train_subwindows")) return B2H_ESHAPE;
  if (T < 1) { set_error("b2h_train_subwindows: bad T"); return B2H_ESHAPE; }
  Geo g = make_geo(n_in, C, pos_emb);
  const int n = use_tc_train(g, T, precision) ? tc_train_nsub(T, is_split(precision)) : 1;
  if (out4) {
    if (i < 0 || i >= n) { set_error("b2h_train_subwindows: sub-window %d of %d", i, n); return B2H_EINVAL; }
    if (n == 1) { out4[0] = 0; out4[1] = 0; out4[2] = T; out4[3] = T; }
    else {
      const int Ts = tc_train_sub_len(is_split(precision));
      const SubWindow s = sub_window(T, Ts, n, i);
      out4[0] = s.start; out4[1] = s.clo; out4[2] = s.chi; out4[3] = Ts;
    }
  }
  return n;
}

extern "C" int64_t b2h_workspace_bytes(int B, int T, int n_in, int C, int pos_emb, int precision) {
  if (!geo_ok(n_in, C, pos_emb, "b2h_workspace_bytes")) return B2H_ESHAPE;
  if (B < 1 || T < 1) return B2H_WS_HEADER + 256;
  return train_plan(make_geo(n_in, C, pos_emb), B, T, precision).total;
}

extern "C" int b2h_pack_weights(const float* params, void* packed, int n_in, int C, int pos_emb, void* stream) {
  if (!params || !packed) { set_error("b2h_pack_weights: null pointer"); return B2H_EINVAL; }
  if (!geo_ok(n_in, C, pos_emb, "b2h_pack_weights")) return B2H_ESHAPE;
  return launch_pack(params, packed, make_geo(n_in, C, pos_emb), (cudaStream_t)stream);
}

extern "C" int b2h_conv_forward(const void* x, int x_dtype, const float* params, const void* packed, const int32_t* lengths,
                                float* y, int B, int T, int n_in, int C, int pos_emb, int precision, int apply_mask,
                                float out_scale, void* stream) {
  if (!x || !params || !packed || !y) { set_error("b2h_conv_forward: null pointer"); return B2H_EINVAL; }
  if (x_dtype != B2H_DT_F32 && x_dtype != B2H_DT_BF16) { set_error("b2h_conv_forward: bad x_dtype %d", x_dtype); return B2H_EINVAL; }
  if (!geo_ok(n_in, C, pos_emb, "b2h_conv_forward")) return B2H_ESHAPE;
  if (B < 0 || T < 1) { set_error("b2h_conv_forward: bad B=%d T=%d", B, T); return B2H_ESHAPE; }
  if (apply_mask && !lengths) { set_error("b2h_conv_forward: apply_mask needs lengths"); return B2H_EINVAL; }
  if (B == 0) return B2H_OK;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 15) || (reinterpret_cast<uintptr_t>(packed) & 15)) {
    set_error("b2h_conv_forward: x, y and packed must be 16-byte aligned");
    return B2H_EALIGN;
  }
  Geo g = make_geo(n_in, C, pos_emb);
  if (!prec_ok(precision)) { set_error("b2h_conv_forward: bad precision %d", precision); return B2H_EINVAL; }
  const int choice = forward_choice(g, T, precision);
  if (precision == B2H_FP32_FFMA || (precision == B2H_FP32 && choice != B2H_KERNEL_TC_TILE)) {
    Fp32Args a{};
    a.x = x; a.x_dtype = x_dtype; a.lengths = lengths; a.params = params; a.packed = reinterpret_cast<const char*>(packed);
    a.y = y; a.B = B; a.T = T; a.apply_mask = apply_mask; a.mode = 0; a.out_scale = out_scale; a.geo = g;
    return launch_fp32(a, false, (cudaStream_t)stream, 0);
  } else {
    if (choice == B2H_KERNEL_TC_TILE)
      return launch_tc_tile_fwd(x, x_dtype, params, reinterpret_cast<const char*>(packed), lengths, y, B, T, apply_mask, out_scale, g,
                                (cudaStream_t)stream, nullptr, is_split(precision));
    if (choice == B2H_KERNEL_TC_WIDE)
      return launch_tc_wide_fwd(x, x_dtype, params, reinterpret_cast<const char*>(packed), lengths, y, B, T, apply_mask, out_scale, g,
                                (cudaStream_t)stream);
    TcFwdArgs a{};                 // long windows: layer-major row-space kernel (reports the limits itself when unsupported)
    a.x = x; a.x_dtype = x_dtype; a.params = params; a.packed = reinterpret_cast<const char*>(packed); a.lengths = lengths;
    a.y = y; a.B = B; a.T = T; a.apply_mask = apply_mask; a.out_scale = out_scale; a.geo = g;
    return launch_tc_fwd(a, (cudaStream_t)stream);
  }
}

extern "C" int b2h_conv_forward_windows(const void* frames, int x_dtype, int64_t n_frames, const int64_t* win_start,
                                        const int64_t* win_end, int pad_mode, const float* params, const void* packed,
                                        const int32_t* lengths, float* y, int n_win, int T, int n_in, int C, int pos_emb,
                                        int precision, int apply_mask, float out_scale, void* stream) {
  if (!frames || !win_start || !params || !packed || !y) { set_error("b2h_conv_forward_windows: null pointer"); return B2H_EINVAL; }
  if (x_dtype != B2H_DT_F32 && x_dtype != B2H_DT_BF16) { set_error("b2h_conv_forward_windows: bad x_dtype %d", x_dtype); return B2H_EINVAL; }
  if (pad_mode != B2H_PAD_REPEAT_FIRST && pad_mode != B2H_PAD_ZEROS) { set_error("b2h_conv_forward_windows: bad pad_mode %d", pad_mode); return B2H_EINVAL; }
  if (!geo_ok(n_in, C, pos_emb, "b2h_conv_forward_windows")) return B2H_ESHAPE;
  if (n_win < 0 || T < 1 || n_frames < 0) { set_error("b2h_conv_forward_windows: bad n_win=%d T=%d", n_win, T); return B2H_ESHAPE; }
  if (apply_mask && !lengths) { set_error("b2h_conv_forward_windows: apply_mask needs lengths"); return B2H_EINVAL; }
  if (n_win == 0) return B2H_OK;
  if ((reinterpret_cast<uintptr_t>(frames) & 15) || (reinterpret_cast<uintptr_t>(y) & 15) || (reinterpret_cast<uintptr_t>(packed) & 15)) {
    set_error("b2h_conv_forward_windows: frames, y and packed must be 16-byte aligned");
    return B2H_EALIGN;
  }
  Geo g = make_geo(n_in, C, pos_emb);
  if ((precision != B2H_BF16 && precision != B2H_FP32) || forward_choice(g, T, precision) != B2H_KERNEL_TC_TILE) {
    set_error("b2h_conv_forward_windows: window views are served by the tcgen05 tile kernel (conv_channels <= 64 in bf16 mode, "
              "<= 32 in fp32 mode, T <= 256); materialise the windows (b2h_preprocess) for C=%d, T=%d, precision=%d", C, T, precision);
    return B2H_ESHAPE;
  }
  WindowView wv{reinterpret_cast<const long long*>(win_start), reinterpret_cast<const long long*>(win_end), (long long)n_frames, pad_mode};
  return launch_tc_tile_fwd(frames, x_dtype, params, reinterpret_cast<const char*>(packed), lengths, y, n_win, T, apply_mask, out_scale, g,
                            (cudaStream_t)stream, &wv, is_split(precision));
}

static int train_common(const void* x, int x_dtype, const float* target, const float* conf, const float* d_y,
                        const int32_t* lengths, const float* params, const void* packed, float* pred_out, int B, int T,
                        int n_in, int C, int pos_emb, int loss_kind, int precision, int mode, void* workspace,
                        int64_t workspace_bytes, cudaStream_t stream, Geo& g, TrainPlan& plan, float*& partials,
                        float*& loss_partials, const char* who, long long* step_dev = nullptr, long long* epoch_dev = nullptr,
                        FuseAdam* fuse = nullptr) {
  if (!x || !params || !packed || !workspace) { set_error("%s: null pointer", who); return B2H_EINVAL; }
  if (x_dtype != B2H_DT_F32 && x_dtype != B2H_DT_BF16) { set_error("%s: bad x_dtype %d", who, x_dtype); return B2H_EINVAL; }
  if (!geo_ok(n_in, C, pos_emb, who)) return B2H_ESHAPE;
  if (B < 1 || T < 1) { set_error("%s: bad B=%d T=%d", who, B, T); return B2H_ESHAPE; }
  if (!prec_ok(precision)) { set_error("%s: bad precision %d", who, precision); return B2H_EINVAL; }
  if (mode == 1) {
    if (!target || !lengths) { set_error("%s: null target/lengths", who); return B2H_EINVAL; }
    if (loss_kind != B2H_LOSS_L1 && loss_kind != B2H_LOSS_CONFL1) { set_error("%s: bad loss_kind %d", who, loss_kind); return B2H_EINVAL; }
    if (loss_kind == B2H_LOSS_CONFL1 && !conf) { set_error("%s: confL1 needs scores", who); return B2H_EINVAL; }
  } else if (!d_y) { set_error("%s: null d_y", who); return B2H_EINVAL; }
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(target) & 15) || (reinterpret_cast<uintptr_t>(packed) & 15)) {
    set_error("%s: x, target and packed must be 16-byte aligned", who);
    return B2H_EALIGN;
  }
  g = make_geo(n_in, C, pos_emb);
  plan = train_plan(g, B, T, precision);
  if (plan.kernel == B2H_KERNEL_NONE) {
    set_error("%s: no training kernel for conv_channels=%d, T=%d, precision=%d (tcgen05 training: C <= 32 in the one-launch tile "
              "kernel, C <= 256 in bf16 mode through the wide kernels, T <= 256; the FFMA kernel would need %zu B of shared memory)",
              who, C, T, precision, fp32_smem_bytes(g, T, true));
    return B2H_ESHAPE;
  }
  const int nparts = plan.nparts;
  const int64_t stride = plan.stride;
  if (workspace_bytes < plan.total - 256) { set_error("%s: workspace %lld B < %lld B", who, (long long)workspace_bytes, (long long)plan.total); return B2H_EWORKSPACE; }
  if (reinterpret_cast<uintptr_t>(workspace) & 15) { set_error("%s: workspace must be 16-byte aligned", who); return B2H_EALIGN; }
  // [header: launch sequence + per-CTA arrival flags of the fused kernel, FIXED offset][gradient partials][loss partials][scratch]
  partials = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + B2H_WS_HEADER);
  loss_partials = partials + (size_t)nparts * stride;
  Fp32Args a{};
  a.x = x; a.x_dtype = x_dtype; a.target = target; a.conf = conf; a.d_y = d_y; a.lengths = lengths; a.params = params;
  a.packed = reinterpret_cast<const char*>(packed); a.y = pred_out; a.partials = partials; a.loss_partials = loss_partials;
  a.B = B; a.T = T; a.loss_kind = loss_kind; a.apply_mask = 1; a.mode = mode; a.out_scale = 1.0f; a.geo = g;
  a.step_dev = step_dev;
  a.epoch_dev = epoch_dev;
  if (plan.kernel == B2H_KERNEL_TC_TILE) {
    if (fuse) {
      if (nparts > B2H_WS_MAX_CTA) { set_error("%s: %d CTAs exceed the %d arrival flags of the workspace header", who, nparts, B2H_WS_MAX_CTA); return B2H_ESHAPE; }
      fuse->enabled = 1;
      fuse->hdr = reinterpret_cast<unsigned*>(workspace);
      a.fuse = *fuse;
    }
    return launch_tc_tile_train(a, stream, is_split(precision));
  }
  if (plan.kernel == B2H_KERNEL_TC_WIDE_TRAIN) {
    // (the forward+criterion kernel's CTA 0 bumps the device-side step / epoch counters the Adam kernels read)
    return launch_tc_wide_train(a, reinterpret_cast<unsigned char*>(workspace) + plan.scratch_off, stream);
  }
  return launch_fp32(a, true, stream, nparts);
}

extern "C" int b2h_train_forward_backward(const void* x, int x_dtype, const float* target, const float* conf,
                                          const int32_t* lengths, const float* params, const void* packed,
                                          float* grads_out, float* loss_out, float* pred_out, int B, int T, int n_in,
                                          int C, int pos_emb, int loss_kind, int precision, int64_t* step_dev,
                                          void* workspace, int64_t workspace_bytes, void* stream) {
  if (grads_out && !loss_out) { set_error("b2h_train_forward_backward: null loss_out"); return B2H_EINVAL; }
  Geo g; TrainPlan plan; float *partials, *loss_partials;
  int rc = train_common(x, x_dtype, target, conf, nullptr, lengths, params, packed, pred_out, B, T, n_in, C, pos_emb,
                        loss_kind, precision, 1, workspace, workspace_bytes, (cudaStream_t)stream, g, plan, partials,
                        loss_partials, "b2h_train_forward_backward", reinterpret_cast<long long*>(step_dev));
  if (rc || !grads_out) return rc;   // grads_out == NULL: only the fused kernel runs, partials stay in the workspace
  return launch_reduce(partials, plan.nparts, plan.gp_layout, g, grads_out, loss_partials, loss_out, (cudaStream_t)stream, nullptr, plan.n_loss);
}

extern "C" int b2h_train_forward_backward_dp(const void* x, int x_dtype, const float* target, const float* conf,
                                             const int32_t* lengths, const float* params, const void* packed,
                                             float* sym_grads, float* loss_out, int B, int T, int n_in, int C, int pos_emb,
                                             int loss_kind, int precision, int64_t* step_dev, int64_t* epoch_dev,
                                             void* workspace, int64_t workspace_bytes, void* stream) {
  if (!sym_grads || !loss_out || !step_dev || !epoch_dev) { set_error("b2h_train_forward_backward_dp: null pointer"); return B2H_EINVAL; }
  Geo g; TrainPlan plan; float *partials, *loss_partials;
  int rc = train_common(x, x_dtype, target, conf, nullptr, lengths, params, packed, nullptr, B, T, n_in, C, pos_emb,
                        loss_kind, precision, 1, workspace, workspace_bytes, (cudaStream_t)stream, g, plan, partials,
                        loss_partials, "b2h_train_forward_backward_dp", reinterpret_cast<long long*>(step_dev),
                        reinterpret_cast<long long*>(epoch_dev));
  if (rc) return rc;
  return launch_reduce(partials, plan.nparts, plan.gp_layout, g, sym_grads, loss_partials, loss_out,
                       (cudaStream_t)stream, reinterpret_cast<const long long*>(epoch_dev), plan.n_loss);
}

extern "C" int b2h_adam_step_dp(float* params, const void* peer_bufs_dev, int rank, int world, float* exp_avg, float* exp_avg_sq,
                                int64_t n, double lr, double beta1, double beta2, double eps, const int64_t* step_dev,
                                const int64_t* epoch_dev, const double* lr_dev, float grad_scale, void* packed, int n_in, int C,
                                int pos_emb, void* stream) {
  if (!params || !peer_bufs_dev || !exp_avg || !exp_avg_sq || !step_dev || !epoch_dev) { set_error("b2h_adam_step_dp: null pointer"); return B2H_EINVAL; }
  if (world < 1 || world > 32 || rank < 0 || rank >= world) { set_error("b2h_adam_step_dp: bad rank/world"); return B2H_EINVAL; }
  if (!geo_ok(n_in, C, pos_emb, "b2h_adam_step_dp")) return B2H_ESHAPE;
  Geo g = make_geo(n_in, C, pos_emb);
  if (g.P != n) { set_error("b2h_adam_step_dp: n=%lld does not match geometry (%d)", (long long)n, g.P); return B2H_ESHAPE; }
  return launch_adam_dp(params, reinterpret_cast<const float* const*>(peer_bufs_dev), rank, world, exp_avg, exp_avg_sq, n, lr, beta1,
                        beta2, eps, reinterpret_cast<const long long*>(step_dev), reinterpret_cast<const long long*>(epoch_dev),
                        lr_dev, grad_scale, packed, g, (cudaStream_t)stream);
}

extern "C" int64_t b2h_dp_exchange_floats(int n_in, int C, int pos_emb, int world) {
  if (!geo_ok(n_in, C, pos_emb, "b2h_dp_exchange_floats")) return B2H_ESHAPE;
  if (world < 1 || world > 32) { set_error("b2h_dp_exchange_floats: bad world"); return B2H_EINVAL; }
  Geo g = make_geo(n_in, C, pos_emb);
  // tensor-core path: uint64 words [2][world][slots] (= 4*world*slots floats); the three-launch path needs
  // [2][P] fp32 + [world] int64 flags, which is smaller
  return 4 * (int64_t)world * gp_total(g) + 64;
}

extern "C" int64_t b2h_dp_exchange_fill(int T, int n_in, int C, int pos_emb, int precision) {
  if (!geo_ok(n_in, C, pos_emb, "b2h_dp_exchange_fill")) return B2H_ESHAPE;
  // fused tile kernel: every 32-bit word starts as the "not arrived" sentinel; three-launch path: zeroed gradients + flags
  return train_plan(make_geo(n_in, C, pos_emb), 1, T, precision).kernel == B2H_KERNEL_TC_TILE ? 0xFFFFFFFFll : 0ll;
}

extern "C" int b2h_dp_status(void) { return dp_status_and_clear(); }

extern "C" int b2h_conv_backward(const void* x, int x_dtype, const float* d_y, const float* params, const void* packed,
                                 float* grads_out, int B, int T, int n_in, int C, int pos_emb, int precision,
                                 void* workspace, int64_t workspace_bytes, void* stream) {
  if (!grads_out) { set_error("b2h_conv_backward: null output"); return B2H_EINVAL; }
  Geo g; TrainPlan plan; float *partials, *loss_partials;
  int rc = train_common(x, x_dtype, nullptr, nullptr, d_y, nullptr, params, packed, nullptr, B, T, n_in, C, pos_emb, 0,
                        precision, 2, workspace, workspace_bytes, (cudaStream_t)stream, g, plan, partials,
                        loss_partials, "b2h_conv_backward");
  if (rc) return rc;
  return launch_reduce(partials, plan.nparts, plan.gp_layout, g, grads_out, nullptr, nullptr, (cudaStream_t)stream, nullptr, plan.n_loss);
}

extern "C" int b2h_train_step(const void* x, int x_dtype, const float* target, const float* conf, const int32_t* lengths,
                              float* params, void* packed, float* exp_avg, float* exp_avg_sq, float* loss_out, int B,
                              int T, int n_in, int C, int pos_emb, int loss_kind, int precision, double lr, double beta1,
                              double beta2, double eps, int64_t step, int64_t* step_dev, const double* lr_dev, void* workspace,
                              int64_t workspace_bytes, void* stream) {
  if (!exp_avg || !exp_avg_sq || !loss_out) { set_error("b2h_train_step: null pointer"); return B2H_EINVAL; }
  if (step < 1 && !step_dev) { set_error("b2h_train_step: step must be >= 1"); return B2H_EINVAL; }
  Geo g; TrainPlan plan; float *partials, *loss_partials;
  // bf16 tile kernel + device-side step counter: ONE cooperative launch (reduction + Adam + re-pack in its tail)
  FuseAdam fuse{};
  fuse.params = params; fuse.m = exp_avg; fuse.v = exp_avg_sq; fuse.packed = reinterpret_cast<char*>(packed);
  fuse.lr = lr; fuse.beta1 = beta1; fuse.beta2 = beta2; fuse.eps = (float)eps; fuse.grad_scale = 1.0f;
  fuse.lr_dev = lr_dev;
  fuse.step_dev = reinterpret_cast<const long long*>(step_dev); fuse.loss_out = loss_out; fuse.world = 1;
  int rc = train_common(x, x_dtype, target, conf, nullptr, lengths, params, packed, nullptr, B, T, n_in, C, pos_emb,
                        loss_kind, precision, 1, workspace, workspace_bytes, (cudaStream_t)stream, g, plan, partials,
                        loss_partials, "b2h_train_step", reinterpret_cast<long long*>(step_dev), nullptr,
                        step_dev ? &fuse : nullptr);
  if (rc) return rc;
  if (step_dev && plan.kernel == B2H_KERNEL_TC_TILE) return B2H_OK;
  return launch_adam(params, partials, plan.nparts, plan.gp_layout, exp_avg, exp_avg_sq, g.P, lr, beta1, beta2, eps, step < 1 ? 1 : step,
                     reinterpret_cast<const long long*>(step_dev), lr_dev, 1.0f, packed, g,
                     loss_partials, loss_out, (cudaStream_t)stream, plan.n_loss);
}

extern "C" int b2h_train_step_dp(const void* x, int x_dtype, const float* target, const float* conf, const int32_t* lengths,
                                 float* params, void* packed, float* exp_avg, float* exp_avg_sq, float* loss_out, int B, int T,
                                 int n_in, int C, int pos_emb, int loss_kind, int precision, double lr, double beta1,
                                 double beta2, double eps, int64_t* step_dev, int64_t* epoch_dev, const double* lr_dev,
                                 float* sym_grads, const void* peer_bufs_dev, void* multicast_buf, int rank, int world,
                                 float grad_scale, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!exp_avg || !exp_avg_sq || !loss_out || !step_dev || !epoch_dev || !sym_grads || !peer_bufs_dev) {
    set_error("b2h_train_step_dp: null pointer");
    return B2H_EINVAL;
  }
  if (world < 1 || world > 32 || rank < 0 || rank >= world) { set_error("b2h_train_step_dp: bad rank/world"); return B2H_EINVAL; }
  Geo g; TrainPlan plan; float *partials, *loss_partials;
  FuseAdam fuse{};
  fuse.params = params; fuse.m = exp_avg; fuse.v = exp_avg_sq; fuse.packed = reinterpret_cast<char*>(packed);
  fuse.lr = lr; fuse.beta1 = beta1; fuse.beta2 = beta2; fuse.eps = (float)eps; fuse.grad_scale = grad_scale;
  fuse.lr_dev = lr_dev;
  fuse.step_dev = reinterpret_cast<const long long*>(step_dev); fuse.loss_out = loss_out;
  fuse.peer_bufs = reinterpret_cast<const float* const*>(peer_bufs_dev); fuse.sym_grads = sym_grads;
  fuse.mc_buf = reinterpret_cast<unsigned long long*>(multicast_buf);
  fuse.epoch_dev = reinterpret_cast<const long long*>(epoch_dev); fuse.rank = rank; fuse.world = world;
  int rc = train_common(x, x_dtype, target, conf, nullptr, lengths, params, packed, nullptr, B, T, n_in, C, pos_emb,
                        loss_kind, precision, 1, workspace, workspace_bytes, (cudaStream_t)stream, g, plan, partials,
                        loss_partials, "b2h_train_step_dp", reinterpret_cast<long long*>(step_dev),
                        reinterpret_cast<long long*>(epoch_dev), &fuse);
  if (rc) return rc;
  if (plan.kernel == B2H_KERNEL_TC_TILE) return B2H_OK;    // everything happened inside the one cooperative launch
  rc = launch_reduce(partials, plan.nparts, plan.gp_layout, g, sym_grads, loss_partials, loss_out, (cudaStream_t)stream,
                     reinterpret_cast<const long long*>(epoch_dev), plan.n_loss);
  if (rc) return rc;
  return launch_adam_dp(params, reinterpret_cast<const float* const*>(peer_bufs_dev), rank, world, exp_avg, exp_avg_sq, g.P, lr,
                        beta1, beta2, eps, reinterpret_cast<const long long*>(step_dev),
                        reinterpret_cast<const long long*>(epoch_dev), lr_dev, grad_scale, packed, g, (cudaStream_t)stream);
}

extern "C" int b2h_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, double lr,
                             double beta1, double beta2, double eps, int64_t step, const int64_t* step_dev,
                             const double* lr_dev, float grad_scale, void* packed, int n_in, int C, int pos_emb, void* stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq) { set_error("b2h_adam_step: null pointer"); return B2H_EINVAL; }
  if ((step < 1 && !step_dev) || n < 0) { set_error("b2h_adam_step: bad step/n"); return B2H_EINVAL; }
  Geo g{};
  if (packed) {
    if (!geo_ok(n_in, C, pos_emb, "b2h_adam_step")) return B2H_ESHAPE;
    g = make_geo(n_in, C, pos_emb);
    if (g.P != n) { set_error("b2h_adam_step: n=%lld does not match geometry (%d)", (long long)n, g.P); return B2H_ESHAPE; }
  }
  if (n == 0) return B2H_OK;
  return launch_adam(params, grads, 1, 0, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, step < 1 ? 1 : step,
                     reinterpret_cast<const long long*>(step_dev), lr_dev, grad_scale, packed, g, nullptr, nullptr, (cudaStream_t)stream);
}

extern "C" int b2h_mask_output(float* y, const int32_t* lengths, int B, int T, int row_elems, void* stream) {
  if (!y || !lengths) { set_error("b2h_mask_output: null pointer"); return B2H_EINVAL; }
  if (B < 0 || T < 0 || row_elems < 1) { set_error("b2h_mask_output: bad shape"); return B2H_ESHAPE; }
  return launch_mask_output(y, lengths, B, T, row_elems, (cudaStream_t)stream);
}

extern "C" int b2h_pose_l1(const float* pred, const float* target, const float* scores, const int32_t* lengths, int B, int T,
                           int row_elems, int loss_kind, float* loss_out, float* d_pred, float* row_scratch, void* stream) {
  if (!pred || !target || !lengths || !loss_out || !row_scratch) { set_error("b2h_pose_l1: null pointer"); return B2H_EINVAL; }
  if (loss_kind != B2H_LOSS_L1 && loss_kind != B2H_LOSS_CONFL1) { set_error("b2h_pose_l1: bad loss_kind"); return B2H_EINVAL; }
  if (loss_kind == B2H_LOSS_CONFL1 && (!scores || (row_elems & 1))) { set_error("b2h_pose_l1: confL1 needs scores and even row_elems"); return B2H_EINVAL; }
  if (B < 1 || T < 1 || row_elems < 1) { set_error("b2h_pose_l1: bad shape"); return B2H_ESHAPE; }
  return launch_pose_l1(pred, target, scores, lengths, B, T, row_elems, loss_kind, loss_out, d_pred, row_scratch, (cudaStream_t)stream);
}

extern "C" int b2h_format_prediction(const float* pred, float* out, int64_t rows, int mode, void* stream) {
  if (!pred || !out) { set_error("b2h_format_prediction: null pointer"); return B2H_EINVAL; }
  if (rows < 0 || (mode != 0 && mode != 1)) { set_error("b2h_format_prediction: bad rows/mode"); return B2H_EINVAL; }
  return launch_format_prediction(pred, out, rows, mode, (cudaStream_t)stream);
}

extern "C" int b2h_pos_emb_concat(const float* inp, float* out, int B, int channels, int T, int max_len, void* stream) {
  if (!inp || !out) { set_error("b2h_pos_emb_concat: null pointer"); return B2H_EINVAL; }
  if (B < 0 || channels < 1 || T < 1 || max_len < 1) { set_error("b2h_pos_emb_concat: bad shape"); return B2H_ESHAPE; }
  if (T != max_len) {   // torch.cat of (B,1,max_len) with (B,C,T) fails in the reference (HandPoseModels.py:82)
    set_error("b2h_pos_emb_concat: Sizes of tensors must match except in dimension 1 (T=%d, max_len=%d)", T, max_len);
    return B2H_ESHAPE;
  }
  return launch_pos_emb_concat(inp, out, B, channels, T, max_len, (cudaStream_t)stream);
}

extern "C" int b2h_tc_probe(const void* a_bf16, const void* b_bf16, float* out, int n, int ksteps, int shift, int variant,
                            void* stream) {
  if (!a_bf16 || !b_bf16 || !out) { set_error("b2h_tc_probe: null pointer"); return B2H_EINVAL; }
  return launch_tc_probe(a_bf16, b_bf16, out, n, ksteps, shift, variant, (cudaStream_t)stream);
}

extern "C" int b2h_tc_bench(void* out_i64x2, int M, int N, int reps, int nacc, int mn_major, void* stream) {
  if (!out_i64x2) { set_error("b2h_tc_bench: null pointer"); return B2H_EINVAL; }
  return launch_tc_bench(reinterpret_cast<long long*>(out_i64x2), M, N, reps, nacc, mn_major, (cudaStream_t)stream);
}

extern "C" void b2h_debug_timing(void* dev_i64x1024) { set_debug_timing(reinterpret_cast<long long*>(dev_i64x1024)); }

extern "C" int b2h_tc_status(void) { return tc_status_and_clear(); }

namespace {
constexpr int kMaxDev = 64;
std::mutex g_cache_mu;
int g_sms[kMaxDev] = {0};
struct SmemAttr { const void* fn; size_t bytes[kMaxDev]; };
SmemAttr g_attr[32] = {};
int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
  return dev < 0 ? 0 : dev;
}
}  // namespace

int b2h::num_sms() {
  const int dev = current_device();
  std::lock_guard<std::mutex> lk(g_cache_mu);
  if (dev < kMaxDev && g_sms[dev]) return g_sms[dev];
  int n = 0;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (n <= 0) n = 148;
  if (dev < kMaxDev) g_sms[dev] = n;
  return n;
}

int b2h::ensure_dyn_smem(const void* fn, size_t bytes) {
  const int dev = current_device();
  std::lock_guard<std::mutex> lk(g_cache_mu);
  SmemAttr* slot = nullptr;
  for (auto& a : g_attr) {
    if (a.fn == fn) { slot = &a; break; }
    if (!a.fn) { a.fn = fn; slot = &a; break; }
  }
  if (slot && dev < kMaxDev && slot->bytes[dev] >= bytes) return B2H_OK;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) { cudaGetLastError(); set_error("cudaFuncSetAttribute(%zu B): %s", bytes, cudaGetErrorString(e)); return B2H_ECUDA; }
  if (slot && dev < kMaxDev) slot->bytes[dev] = bytes;
  return B2H_OK;
}
