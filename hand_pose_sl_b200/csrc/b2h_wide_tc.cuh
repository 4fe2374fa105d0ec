// bf16 mode, wide forward kernel (64 < conv_channels <= 256): ConvModel.forward on tcgen05 / TMEM with the weights
// STREAMED through shared memory.  Included by b2h_conv_tc.cu (same translation unit as the PTX layer).
//
// Reference semantics: ConvModel.forward  body2hand/src/models/HandPoseModels.py:40-64 (`--conv-channels`, run.py),
//                      mask_output        body2hand/src/steps/utils.py:309-312 (optional epilogue).
//
// At C = 256 one layer's weights are 5 x 256 x 256 bf16 = 640 KB: they cannot live in shared memory like the C <= 64
// tile kernel's (b2h_train_tc.cuh), and an N = 256 MMA is datapath-bound (128 x 256 x 16 MACs = 128 tensor cycles)
// instead of instruction-bound.  So the kernel is organised like a GEMM main loop:
//   * tile = 256 output rows = two M=128 accumulators of N <= 256 fp32 columns = all 512 TMEM columns; the rows hold
//     floor(258 / (T+2)) whole windows separated by 2 shared zero rows (T=64: 3 windows, T=126: 2, T<=256: 1);
//   * ONE activation buffer [channel/8][264 rows][8 ch] bf16 (no-swizzle K-major UMMA layout, conv tap k = +k rows in
//     the descriptor start address) that every layer reads and then overwrites IN PLACE: all MMAs of a layer have
//     completed (commit -> acc_full) before its epilogue writes, and inference keeps no activations;
//   * the packed UMMA B blocks (b2h_common.cuh umma_b_offset: one [2][N][8] block per (tap, 16-channel k-step),
//     contiguous in global memory) stream through a ring of 8-KB stages by 1-D TMA bulk copies issued by a dedicated
//     producer warp that runs ahead across layers and tiles; full[]/empty[] mbarriers, the empty side armed by
//     tcgen05.commit of the MMAs that read the stage;
//   * warp roles: 8 epilogue warps (two per TMEM lane quadrant, alternating 32-column chunks: tcgen05.ld -> bias/ReLU
//     -> bf16 -> st.shared, or the fp32 prediction rows of layer 4), 1 MMA-issuing warp, 1 weight-producer warp.
#pragma once

namespace b2h {
using namespace tc;

struct WideArgs {
  const void* x; int x_dtype;
  const float* params; const char* packed; const int32_t* lengths;
  float* y;
  long long* dbg;
  int B, T, apply_mask;
  float out_scale;
  int n_tiles, gh, nstage, a_bytes;
  Geo geo;
};

constexpr int kWideThreads = 320;
constexpr int kWideEpiThreads = 256;
constexpr int kWideRows = 264;            // 2 zero rows + 256 output rows + 6 zero rows
constexpr int kWideStage = 8192;          // bytes per ring stage (one N=256 block, or several narrower ones)
constexpr int kWideMaxStages = 16;

struct WideSched { int KS, N, blk_bytes, nblk, bps, nst; };
__host__ __device__ inline WideSched wide_sched(const Geo& g, int l) {
  WideSched s;
  s.KS = g.kp[l] >> 4; s.N = g.np_[l];
  s.blk_bytes = s.N * 32;
  s.nblk = B2H_KW * s.KS;
  s.bps = kWideStage / s.blk_bytes; if (s.bps < 1) s.bps = 1;
  s.nst = (s.nblk + s.bps - 1) / s.bps;
  return s;
}

__global__ void __launch_bounds__(kWideThreads, 1) conv_tc_wide_fwd_kernel(WideArgs p) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t full_bar[kWideMaxStages];    // stage landed (TMA complete_tx)
  __shared__ __align__(8) uint64_t empty_bar[kWideMaxStages];   // stage consumed (tcgen05.commit)
  __shared__ __align__(8) uint64_t acc_full;                    // a layer's accumulators are complete
  __shared__ __align__(8) uint64_t act_ready;                   // activation buffer written + accumulators drained
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float bias_s[4][256];
  const Geo& g = p.geo;
  const int T = p.T, gh = p.gh;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  constexpr int rows = kWideRows;
  constexpr int CH = rows * 16;
  unsigned char* A = smem;
  unsigned char* RING = smem + p.a_bytes;
  const int S = p.nstage;

  // ---- setup: biases, zeroed activation buffer, barriers, TMEM ----
  for (int i = tid; i < 4 * 256; i += kWideThreads) {
    const int l = i >> 8, c = i & 255;
    bias_s[l][c] = c < g.cout[l] ? __ldg(p.params + g.b_off[l] + c) : 0.0f;
  }
  {
    uint4* z = reinterpret_cast<uint4*>(A);
    const uint4 zero = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < p.a_bytes / 16; i += kWideThreads) z[i] = zero;
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&acc_full, 1);
    mbar_init(&act_ready, kWideEpiThreads);
    fence_barrier_init();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;

  if (warp == 9) {
    // ===================== weight producer: one thread, runs ahead by the ring depth =====================
    if (elect_one()) {
      uint32_t gs = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x)
        for (int l = 0; l < 4; ++l) {
          const WideSched sc = wide_sched(g, l);
          const char* src = p.packed + g.tf_off[l];
          for (int st = 0; st < sc.nst; ++st, ++gs) {
            const uint32_t slot = gs % S, use = gs / S;
            if (use > 0 && !mbar_wait(&empty_bar[slot], (use - 1) & 1, 60)) return;
            int nb = sc.nblk - st * sc.bps; nb = nb > sc.bps ? sc.bps : nb;
            const uint32_t bytes = (uint32_t)nb * sc.blk_bytes;
            mbar_arrive_expect_tx(&full_bar[slot], bytes);
            bulk_g2s(RING + (size_t)slot * kWideStage, src + (size_t)st * sc.bps * sc.blk_bytes, bytes, &full_bar[slot]);
          }
        }
    }
    __syncwarp();
  } else if (warp == 8) {
    // ===================== MMA issuer: one thread =====================
    if (elect_one()) {
      uint32_t gs = 0, act_phase = 0;
      const uint32_t hi_k = desc_hi(128);
      const uint32_t a_lo0 = desc_lo(smem_u32(A), (uint32_t)CH);
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x)
        for (int l = 0; l < 4; ++l) {
          const WideSched sc = wide_sched(g, l);
          const uint32_t idesc = make_idesc_bf16(128, sc.N, 0, 0);
          if (!mbar_wait(&act_ready, act_phase, 61)) return;
          act_phase ^= 1;
          tc_fence_after();
          int k = 0, s = 0;                           // block q = k*KS + s
          uint32_t acc = 0;
          for (int st = 0; st < sc.nst; ++st, ++gs) {
            const uint32_t slot = gs % S, use = gs / S;
            if (!mbar_wait(&full_bar[slot], use & 1, 62)) return;
            tc_fence_after();
            int nb = sc.nblk - st * sc.bps; nb = nb > sc.bps ? sc.bps : nb;
            const uint32_t ring = smem_u32(RING + (size_t)slot * kWideStage);
            for (int b = 0; b < nb; ++b) {
              const uint64_t bd = desc64(desc_lo(ring + (uint32_t)b * sc.blk_bytes, (uint32_t)sc.N * 16), hi_k);
              const uint32_t a_lo = a_lo0 + 2 * s * rows + k;   // output row 2+m reads input row m+k
              umma_bf16(tbase, desc64(a_lo, hi_k), bd, idesc, acc);
              umma_bf16(tbase + 256, desc64(a_lo + 128, hi_k), bd, idesc, acc);
              acc = 1;
              if (++s == sc.KS) { s = 0; ++k; }
            }
            umma_commit(&empty_bar[slot]);            // stage reusable once these MMAs have read it
          }
          umma_commit(&acc_full);
        }
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps (256 threads): staging, bias/ReLU, prediction rows =====================
    const int r128 = tid & 127;
    const int ch = warp >> 2;                                  // column half (alternating 32-column chunks)
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    const int n_in = g.n_in, pe = g.pos_emb;
    const int nch0 = g.kp[0] / 8;
    const bool vec_in = (pe == 0) && ((n_in & 7) == 0);
    struct RowCtx { int row, t, gw; bool valid; };
    auto rowctx = [&](int j, int tile) {
      const int mm = r128 + 128 * j;
      const int wjj = mm / (T + 2);
      RowCtx r;
      r.t = mm - wjj * (T + 2);
      r.row = 2 + mm;
      r.gw = tile * gh + wjj;
      r.valid = (r.t < T) && (wjj < gh) && (r.gw < p.B) && tile < p.n_tiles;
      return r;
    };
    auto stage_inputs = [&](int tile) {      // (n_in) channels NWC -> chunks [0, kp0/8) of the activation buffer
      for (int j = 0; j < 2; ++j) {
        const RowCtx rc = rowctx(j, tile);
        for (int c8 = ch; c8 < nch0; c8 += 2) {
          uint4 q = make_uint4(0, 0, 0, 0);
          if (rc.valid) {
            const size_t base = ((size_t)rc.gw * T + rc.t) * n_in;
            if (vec_in && c8 * 8 < n_in) {
              if (p.x_dtype == B2H_DT_F32) {
                const float4* src = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.x) + base) + 2 * c8;
                const float4 lo = __ldg(src), hi = __ldg(src + 1);
                q.x = pack_bf16x2(lo.x, lo.y); q.y = pack_bf16x2(lo.z, lo.w);
                q.z = pack_bf16x2(hi.x, hi.y); q.w = pack_bf16x2(hi.z, hi.w);
              } else {
                q = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.x) + base) + c8);
              }
            } else if (!vec_in) {
              float v[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const int cc = c8 * 8 + e;          // channel in the conv1 input (pos-emb row first)
                float val = 0.0f;
                if (pe && cc == 0) val = __fdiv_rn((float)rc.t, 100.0f);             // HandPoseModels.py:70-82
                else if (cc - pe < n_in && cc - pe >= 0)
                  val = (p.x_dtype == B2H_DT_F32) ? __ldg(reinterpret_cast<const float*>(p.x) + base + (cc - pe))
                                                  : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.x)[base + (cc - pe)]);
                v[e] = val;
              }
              q.x = pack_bf16x2(v[0], v[1]); q.y = pack_bf16x2(v[2], v[3]);
              q.z = pack_bf16x2(v[4], v[5]); q.w = pack_bf16x2(v[6], v[7]);
            }
          }
          *reinterpret_cast<uint4*>(A + (size_t)c8 * CH + (size_t)rc.row * 16) = q;
        }
      }
    };
    auto publish = [&]() {                    // my smem writes -> async proxy, my TMEM reads are done
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&act_ready);
    };

    stage_inputs(blockIdx.x);
    publish();
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      for (int l = 0; l < 4; ++l) {
        const int N = g.np_[l];
        if (!mbar_wait(&acc_full, acc_phase, 63 + l)) return;
        acc_phase ^= 1;
        tc_fence_after();
        for (int j = 0; j < 2; ++j) {
          const RowCtx rc = rowctx(j, tile);
          const uint32_t taddr = tbase + lane_addr + 256 * j;
          if (l < 3) {
            for (int c0 = 32 * ch; c0 < N; c0 += 64) {
              uint32_t v[32];
              tmem_ld32(taddr + c0, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                if (c0 + 8 * i >= N) continue;        // N is a multiple of 16: never write a chunk past the layer's columns
                float f[8];
#pragma unroll
                for (int q = 0; q < 8; ++q)
                  f[q] = rc.valid ? fmaxf(__uint_as_float(v[8 * i + q]) + bias_s[l][c0 + 8 * i + q], 0.0f) : 0.0f;
                store8_bf16(A, CH, rc.row, (c0 >> 3) + i, f);
              }
            }
          } else {
            // layer 4: prediction rows (+ mask_output utils.py:309-312, de-normalisation) straight to global memory
            int len = T;
            if (rc.valid && p.lengths) { len = p.lengths[rc.gw]; len = len < 0 ? 0 : (len > T ? T : len); }
            float* yrow = rc.valid ? p.y + ((size_t)rc.gw * T + rc.t) * B2H_COUT : nullptr;
            const bool masked = p.apply_mask && (rc.t >= len);
            for (int c0 = 16 * ch; c0 < N; c0 += 32) {
              uint32_t v[16];
              tmem_ld16(taddr + c0, v);
              tmem_ld_wait();
#pragma unroll
              for (int q = 0; q < 16; q += 2) {
                const int c = c0 + q;
                if (yrow && c < B2H_COUT) {
                  float a = __uint_as_float(v[q]) + bias_s[3][c];
                  float b = __uint_as_float(v[q + 1]) + bias_s[3][c + 1];
                  if (masked) { a = 0.0f; b = 0.0f; }
                  else if (p.out_scale != 1.0f) { a *= p.out_scale; b *= p.out_scale; }
                  *reinterpret_cast<float2*>(yrow + c) = make_float2(a, b);
                }
              }
            }
          }
        }
        if (l == 3) stage_inputs(tile + gridDim.x);   // the buffer is dead after layer 4's MMAs: next tile's rows go in
        publish();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

// ---- host side ----
inline int wide_a_bytes(const Geo& g) {
  int kmax = g.kp[0];
  for (int l = 1; l < 4; ++l) kmax = kmax > g.kp[l] ? kmax : g.kp[l];
  for (int l = 0; l < 3; ++l) kmax = kmax > g.np_[l] ? kmax : g.np_[l];
  return (kmax / 8) * kWideRows * 16;
}

bool tc_wide_supported(const Geo& g, int T) {
  if (T < 1 || T > 256 || g.C > 256 || g.cin[0] > 64) return false;
  return (size_t)wide_a_bytes(g) + 4 * kWideStage <= (size_t)220 * 1024;
}

int launch_tc_wide_fwd(const void* x, int x_dtype, const float* params, const char* packed, const int32_t* lengths, float* y,
                       int B, int T, int apply_mask, float out_scale, const Geo& g, cudaStream_t stream) {
  WideArgs p{};
  p.x = x; p.x_dtype = x_dtype; p.params = params; p.packed = packed; p.lengths = lengths; p.y = y;
  p.B = B; p.T = T; p.apply_mask = apply_mask; p.out_scale = out_scale; p.geo = g;
  p.gh = 258 / (T + 2);
  p.n_tiles = (B + p.gh - 1) / p.gh;
  p.a_bytes = wide_a_bytes(g);
  int S = (int)(((size_t)220 * 1024 - p.a_bytes) / kWideStage);
  if (S > kWideMaxStages) S = kWideMaxStages;
  if (S < 4) { set_error("wide tensor-core forward: C=%d leaves no room for the weight ring", g.C); return B2H_ESHAPE; }
  p.nstage = S;
  const size_t smem = (size_t)p.a_bytes + (size_t)S * kWideStage;
  static size_t attr_bytes = 0;
  if (smem > attr_bytes) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_wide_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { cudaGetLastError(); set_error("cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e)); return B2H_ECUDA; }
    attr_bytes = smem;
  }
  int grid = num_sms();
  if (grid > p.n_tiles) grid = p.n_tiles;
  conv_tc_wide_fwd_kernel<<<grid, kWideThreads, smem, stream>>>(p);
  count_launch();
  return check_launch("conv_tc_wide_fwd_kernel");
}

}  // namespace b2h
