// Shared geometry / helpers for libb2h.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/b2h.h"

#define B2H_KW 5          // kernel_size of every Conv1d (HandPoseModels.py:24-32)
#define B2H_PADW 2        // padding=2
#define B2H_COUT 42       // 2*21 output channels (HandPoseModels.py:32)
#define B2H_MAX_C 256

namespace b2h {

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// Layout of the flat fp32 parameter buffer (state_dict order) and of the packed weight buffer.
struct Geo {
  int n_in, C, pos_emb;
  int cin[4], cout[4];
  int w_off[4], b_off[4];   // float offsets into the flat parameter buffer
  int P;                    // total parameter count
  // packed buffer (byte offsets, each 128-B aligned)
  int64_t wf_off[4];        // fp32 forward taps   Wf[k][ci][co]            = W[co][ci][k]
  int64_t wd_off[4];        // fp32 dgrad taps     Wd[k'][co][ci]           = W[co][ci][4-k']   (layers 2..4)
  int64_t tf_off[4];        // bf16 UMMA B operand, forward  (blocks of [2][N][8])
  int64_t td_off[4];        // bf16 UMMA B operand, dgrad    (layers 2..4)
  int64_t bias_off;         // fp32 [4][64] zero-padded biases (one 1-KB bulk copy into smem)
  int kp[4], np_[4];        // UMMA padded reduction (cin -> mult of 16) and N (cout -> mult of 16), forward
  int64_t packed_bytes;
};

__host__ __device__ inline Geo make_geo(int n_in, int C, int pos_emb) {
  Geo g;
  g.n_in = n_in; g.C = C; g.pos_emb = pos_emb;
  g.cin[0] = n_in + (pos_emb ? 1 : 0); g.cout[0] = C;
  g.cin[1] = C; g.cout[1] = C;
  g.cin[2] = C; g.cout[2] = C;
  g.cin[3] = C; g.cout[3] = B2H_COUT;
  int off = 0;
  for (int l = 0; l < 4; ++l) {
    g.w_off[l] = off; off += g.cout[l] * g.cin[l] * B2H_KW;
    g.b_off[l] = off; off += g.cout[l];
  }
  g.P = off;
  int64_t b = 0;
  for (int l = 0; l < 4; ++l) { g.wf_off[l] = b; b += (int64_t)B2H_KW * g.cin[l] * g.cout[l] * 4; b = (b + 127) / 128 * 128; }
  for (int l = 0; l < 4; ++l) { g.wd_off[l] = b; if (l > 0) { b += (int64_t)B2H_KW * g.cin[l] * g.cout[l] * 4; b = (b + 127) / 128 * 128; } }
  for (int l = 0; l < 4; ++l) {
    g.kp[l] = round_up(g.cin[l], 16); g.np_[l] = round_up(g.cout[l], 16);
    g.tf_off[l] = b; b += (int64_t)B2H_KW * g.kp[l] * g.np_[l] * 2; b = (b + 127) / 128 * 128;
  }
  for (int l = 0; l < 4; ++l) {
    g.td_off[l] = b;
    if (l > 0) { b += (int64_t)B2H_KW * round_up(g.cout[l], 16) * round_up(g.cin[l], 16) * 2; b = (b + 127) / 128 * 128; }
  }
  g.bias_off = b; b += 4 * 64 * 4;
  g.packed_bytes = b;
  return g;
}

// Byte offset inside one UMMA B-operand section of element (n, kk) of tap k.
// Section = blocks q = k*(KP/16)+s of [2 k-chunks][N rows][8 bf16] (no-swizzle K-major canonical
// layout: core matrix = 8 rows x 16 B, SBO = 128 B between 8-row groups, LBO = N*16 B between chunks).
__host__ __device__ inline int64_t umma_b_offset(int k, int kk, int n, int KP, int N) {
  int s = kk >> 4, chunk = (kk >> 3) & 1, within = kk & 7;
  int64_t q = (int64_t)k * (KP >> 4) + s;
  return q * ((int64_t)N * 32) + (int64_t)chunk * (N * 16) + (int64_t)n * 16 + within * 2;
}

// Gradient-partial layout written by the tensor-core train kernel (coalesced 128-B rows):
//   per layer l: W part [k][co][kp_l] then bias part [cout_l]
__host__ __device__ inline int gp_layer_size(const Geo& g, int l) { return B2H_KW * g.cout[l] * g.kp[l] + round_up(g.cout[l], 4); }
__host__ __device__ inline int gp_layer_off(const Geo& g, int l) {
  int o = 0;
  for (int q = 0; q < l; ++q) o += gp_layer_size(g, q);
  return o;
}
__host__ __device__ inline int gp_total(const Geo& g) { return gp_layer_off(g, 4); }
// flat parameter index -> index in the GP layout
__host__ __device__ inline int gp_index_of_flat(const Geo& g, int i) {
  int l = 0;
  for (int q = 1; q < 4; ++q)
    if (i >= g.w_off[q]) l = q;
  const int base = gp_layer_off(g, l);
  if (i >= g.b_off[l]) return base + B2H_KW * g.cout[l] * g.kp[l] + (i - g.b_off[l]);
  const int cin = g.cin[l];
  const int rel = i - g.w_off[l];
  const int co = rel / (cin * B2H_KW);
  const int rem = rel - co * cin * B2H_KW;
  const int ci = rem / B2H_KW, k = rem - ci * B2H_KW;
  return base + (k * g.cout[l] + co) * g.kp[l] + ci;
}

// GP index -> flat parameter index (-1 for the padding slots of the GP layout)
__host__ __device__ inline int flat_index_of_gp(const Geo& g, int j) {
  int l = 0, base = 0;
  for (int q = 0; q < 4; ++q) {
    const int sz = gp_layer_size(g, q);
    if (j >= base + sz && q < 3) { base += sz; l = q + 1; }
    else break;
  }
  const int rel = j - base;
  const int wsz = B2H_KW * g.cout[l] * g.kp[l];
  if (rel >= wsz) { const int b = rel - wsz; return b < g.cout[l] ? g.b_off[l] + b : -1; }
  const int k = rel / (g.cout[l] * g.kp[l]);
  const int r2 = rel - k * g.cout[l] * g.kp[l];
  const int co = r2 / g.kp[l], ci = r2 - co * g.kp[l];
  return ci < g.cin[l] ? g.w_off[l] + (co * g.cin[l] + ci) * B2H_KW + k : -1;
}

// Scatter one fp32 parameter (flat index i) into the packed operand layouts.
__device__ __forceinline__ void scatter_packed(const Geo& g, char* packed, int i, float v) {
  int l = 0;
#pragma unroll
  for (int q = 1; q < 4; ++q)
    if (i >= g.w_off[q]) l = q;
  if (i >= g.b_off[l]) {        // bias: also kept zero-padded [4][64] for the tensor-core kernels' smem copy
    if (i - g.b_off[l] < 64) reinterpret_cast<float*>(packed + g.bias_off)[l * 64 + (i - g.b_off[l])] = v;
    return;
  }
  const int cin = g.cin[l], cout = g.cout[l];
  const int rel = i - g.w_off[l];
  const int co = rel / (cin * B2H_KW);
  const int rem = rel - co * cin * B2H_KW;
  const int ci = rem / B2H_KW, k = rem - ci * B2H_KW;
  reinterpret_cast<float*>(packed + g.wf_off[l])[(k * cin + ci) * cout + co] = v;
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  *reinterpret_cast<__nv_bfloat16*>(packed + g.tf_off[l] + umma_b_offset(k, ci, co, g.kp[l], g.np_[l])) = h;
  if (l > 0) {
    reinterpret_cast<float*>(packed + g.wd_off[l])[((B2H_KW - 1 - k) * cout + co) * cin + ci] = v;
    *reinterpret_cast<__nv_bfloat16*>(packed + g.td_off[l] +
                                      umma_b_offset(B2H_KW - 1 - k, co, ci, round_up(cout, 16), round_up(cin, 16))) = h;
  }
}


// beta^t for an integer step count by binary exponentiation (~2*log2(t) DMULs instead of a double-precision pow())
__host__ __device__ inline double ipow(double b, long long t) {
  double r = 1.0;
  while (t > 0) {
    if (t & 1) r *= b;
    b *= b;
    t >>= 1;
  }
  return r;
}

// Optional tail of the tensor-core train kernel: cross-CTA gradient reduction, (data-parallel) gradient exchange
// over peer memory and Adam, all inside the SAME cooperative launch (grid barriers through `sync`).
struct FuseAdam {
  int enabled;
  float* params; float* m; float* v; char* packed;
  double lr, beta1, beta2;
  float eps, grad_scale;
  const long long* step_dev;
  float* loss_out;
  unsigned* sync;                    // [2] zero-initialised words: arrival counter, generation
  // data parallel (world > 1)
  const float* const* peer_bufs;     // device array [world] of peer exchange buffers ([2][P] fp32 + [world] int64 flags)
  float* sym_grads;                  // this rank's exchange buffer
  const long long* epoch_dev;
  int rank, world;
};

void set_error(const char* fmt, ...);
int check_launch(const char* what);
void count_launch(int n = 1);


// ---- kernel argument blocks and launchers shared between translation units ----
struct Fp32Args {
  const void* x; int x_dtype;
  const float* target; const float* conf; const float* d_y; const int32_t* lengths;
  const float* params; const char* packed;
  float* y;                 // forward output / masked prediction (nullable in train mode)
  float* partials;          // [grid][P]
  float* loss_partials;     // [grid]
  long long* step_dev;      // nullable: device step counter, incremented by block 0 (read by the Adam kernel)
  long long* epoch_dev;     // nullable: data-parallel exchange epoch, incremented the same way (never rewound)
  FuseAdam fuse;            // tensor-core train kernel only: reduce + exchange + Adam in the same launch
  int B, T, loss_kind, apply_mask, mode;   // mode 0 = forward, 1 = train (loss inside), 2 = backward of given d_y
  float out_scale;
  Geo geo;
};
int launch_fp32(Fp32Args& p, bool train, cudaStream_t stream, int grid_override);
int fp32_train_grid(const Geo& g, int B, int T);
size_t fp32_smem_bytes(const Geo& g, int T, bool train);

struct TcFwdArgs {
  const void* x; int x_dtype;
  const float* params; const char* packed; const int32_t* lengths;
  float* y;
  int B, T, G, NT, RBUF, apply_mask;
  float out_scale;
  Geo geo;
};
int launch_tc_fwd(TcFwdArgs& p, cudaStream_t stream);
bool tc_wide_supported(const Geo& g, int T);
int launch_tc_wide_fwd(const void* x, int x_dtype, const float* params, const char* packed, const int32_t* lengths, float* y,
                       int B, int T, int apply_mask, float out_scale, const Geo& g, cudaStream_t stream);
bool tc_fwd_supported(const Geo& g, int T);
int launch_tc_probe(const void* a, const void* b, float* out, int n, int ksteps, int shift, int variant, cudaStream_t stream);
int tc_status_and_clear();

struct TcTileArgs;
int tc_train_grid(const Geo& g, int B, int T);
bool tc_tile_ok(const Geo& g, int T, bool train);
int launch_tc_tile_fwd(const void* x, int x_dtype, const float* params, const char* packed, const int32_t* lengths, float* y,
                       int B, int T, int apply_mask, float out_scale, const Geo& g, cudaStream_t stream);
int launch_tc_tile_train(const Fp32Args& a, cudaStream_t stream);

void set_debug_timing(long long* p);
int launch_tc_bench(long long* out, int M, int N, int reps, int nacc, int mn_major, cudaStream_t stream);
int launch_format_prediction(const float* pred, float* out, int64_t rows, int mode, cudaStream_t stream);
int launch_pack(const float* params, void* packed, const Geo& g, cudaStream_t stream);
int launch_reduce(const float* partials, int nparts, int gp_layout, const Geo& g, float* grads, const float* loss_partials,
                  float* loss_out, cudaStream_t stream, const long long* epoch_dev = nullptr);
int launch_adam_dp(float* params, const float* const* peer_bufs, int rank, int world, float* m, float* v, int64_t n, double lr,
                   double beta1, double beta2, double eps, const long long* step_dev, const long long* epoch_dev, float grad_scale,
                   void* packed, const Geo& g, cudaStream_t stream);
int dp_status_and_clear();
int launch_adam(float* params, const float* grads, int nparts, int gp_layout, float* m, float* v, int64_t n, double lr, double beta1, double beta2,
                double eps, int64_t step, const long long* step_dev, float grad_scale, void* packed, const Geo& g,
                const float* loss_partials, float* loss_out, cudaStream_t stream);
int launch_mask_output(float* y, const int32_t* lengths, int B, int T, int row, cudaStream_t stream);
int launch_pose_l1(const float* pred, const float* target, const float* scores, const int32_t* lengths, int B, int T, int row,
                   int loss_kind, float* loss_out, float* d_pred, float* row_scratch, cudaStream_t stream);
int num_sms();

}  // namespace b2h
