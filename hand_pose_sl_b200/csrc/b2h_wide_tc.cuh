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
//     contiguous in global memory) stream through a ring of 16-KB stages by TMA tensor-map loads (the packed sections
//     viewed as a 2-D tensor of 128-byte rows, box = 128 rows; measured here: 1-D cp.async.bulk copies top out at
//     ~23 B/cycle/SM, a third of what two N=256 MMAs per stage consume) issued by a dedicated producer warp that
//     runs ahead across layers and tiles; full[]/empty[] mbarriers, the empty side armed by tcgen05.commit of the
//     MMAs that read the stage;
//   * warp roles: 8 epilogue warps (two per TMEM lane quadrant, alternating 32-column chunks: tcgen05.ld -> bias/ReLU
//     -> bf16 -> st.shared, or the fp32 prediction rows of layer 4), 1 MMA-issuing warp, 1 weight-producer warp.
#pragma once
#include <cuda.h>   // CUtensorMap + enums only; the encoder is fetched with cudaGetDriverEntryPoint (no libcuda link)

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
constexpr int kWideStage = 16384;         // bytes per ring stage (two N=256 blocks, or several narrower ones)
constexpr int kWideBoxRows = kWideStage / 128;
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

// raw shared-address variants of the barrier / TMA / commit helpers for the two single-thread loops
__device__ __forceinline__ void tma_load_2d_addr(uint32_t smem_dst, const CUtensorMap* tmap, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void expect_tx_addr(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void commit_addr(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool wait_addr(uint32_t bar, uint32_t parity, int site) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  if (ok) return true;                               // fast path: no clock reads
  const long long t0 = clock64();
  for (;;) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return true;
    if (clock64() - t0 > 4000000000LL) { atomicExch(&g_tc_status, site); return false; }
  }
}

__global__ void __launch_bounds__(kWideThreads, 1) conv_tc_wide_fwd_kernel(WideArgs p, const __grid_constant__ CUtensorMap wmap) {
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

  // Both single-thread loops below are latency-bound scalar code (every instruction waits for the previous one): a
  // first version that recomputed slot = stage % S, the barrier addresses and the descriptors per stage needed ~540
  // cycles per stage -- twice the 258 tensor cycles of the two N=256 MMAs it feeds.  So: raw shared-memory addresses
  // computed once, ring position and parity carried as counters, descriptors advanced by one add.
  const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]), ring0 = smem_u32(RING);
  if (warp == 9) {
    // ===================== weight producer: one thread, runs ahead by the ring depth =====================
    if (elect_one()) {
      uint32_t slot = 0, ph = 1;                     // parity 1 on a fresh barrier = "already free" (first lap)
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x)
        for (int l = 0; l < 4; ++l) {
          const WideSched sc = wide_sched(g, l);
          int row = (int)((g.tf_off[l] - g.tf_off[0]) >> 7);            // 128-byte rows of the weight tensor map
          const int rows_per_stage = (sc.bps * sc.blk_bytes) >> 7;
          for (int st = 0; st < sc.nst; ++st) {
            if (!wait_addr(empty0 + slot * 8, ph, 60)) return;
            // one 128-row box = 16 KB whatever the stage's payload is (rows past it belong to the next stage / are
            // zero-filled past the tensor): the transaction count is always the full box
            expect_tx_addr(full0 + slot * 8, kWideStage);
            tma_load_2d_addr(ring0 + slot * kWideStage, &wmap, 0, row, full0 + slot * 8);
            row += rows_per_stage;
            if (++slot == (uint32_t)S) { slot = 0; ph ^= 1; }
          }
        }
    }
    __syncwarp();
  } else if (warp == 8) {
    // ===================== MMA issuer: one thread =====================
    if (elect_one()) {
      uint32_t slot = 0, ph = 0, act_phase = 0;
      int dn = 0;
      const uint32_t hi_k = desc_hi(128);
      const uint32_t a_lo0 = desc_lo(smem_u32(A), (uint32_t)CH);
      const uint32_t acc_addr = smem_u32(&acc_full);
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x)
        for (int l = 0; l < 4; ++l) {
          const WideSched sc = wide_sched(g, l);
          const uint32_t idesc = make_idesc_bf16(128, sc.N, 0, 0);
          const uint32_t b_lo0 = desc_lo(ring0, (uint32_t)sc.N * 16);   // + slot * stage + block * blk, in 16-B units
          const uint32_t blk16 = (uint32_t)sc.blk_bytes >> 4;
          if (!mbar_wait(&act_ready, act_phase, 61)) return;
          act_phase ^= 1;
          tc_fence_after();
          if (p.dbg && blockIdx.x == 0 && dn < 8) p.dbg[dn++] = clock64();   // layer inputs ready
          int s = 0, left = sc.nblk;
          uint32_t a_tap = a_lo0, a_cur = a_lo0;       // output row 2+m reads input row m+k: tap k = +k rows
          uint32_t acc = 0;
          for (int st = 0; st < sc.nst; ++st) {
            // no tcgen05.fence after this wait: the mbarrier orders the TMA writes before the MMAs' reads
            if (!wait_addr(full0 + slot * 8, ph, 62)) return;
            uint32_t b_cur = b_lo0 + slot * (kWideStage >> 4);
            const int nb = left < sc.bps ? left : sc.bps;
            left -= nb;
            for (int b = 0; b < nb; ++b) {
              const uint64_t bd = desc64(b_cur, hi_k);
              umma_bf16(tbase, desc64(a_cur, hi_k), bd, idesc, acc);
              umma_bf16(tbase + 256, desc64(a_cur + 128, hi_k), bd, idesc, acc);
              acc = 1;
              b_cur += blk16;
              a_cur += 2 * rows;                       // next 16-channel k-step: two chunks further
              if (++s == sc.KS) { s = 0; a_tap += 1; a_cur = a_tap; }
            }
            commit_addr(empty0 + slot * 8);            // stage reusable once these MMAs have read it
            if (++slot == (uint32_t)S) { slot = 0; ph ^= 1; }
          }
          commit_addr(acc_addr);
          if (p.dbg && blockIdx.x == 0 && dn < 8) p.dbg[dn++] = clock64();   // layer MMAs issued
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
    int en = 64;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      for (int l = 0; l < 4; ++l) {
        const int N = g.np_[l];
        if (!mbar_wait(&acc_full, acc_phase, 63 + l)) return;
        acc_phase ^= 1;
        tc_fence_after();
        if (p.dbg && blockIdx.x == 0 && tid == 0 && en < 124) p.dbg[en++] = clock64();   // accumulators complete
        if (l < 3) {
          // both 128-row accumulators' chunks are fetched before one wait: two TMEM loads in flight per thread
          const RowCtx rc0 = rowctx(0, tile), rc1 = rowctx(1, tile);
          const uint32_t taddr = tbase + lane_addr;
          for (int c0 = 32 * ch; c0 < N; c0 += 64) {
            uint32_t v0[32], v1[32];
            tmem_ld32(taddr + c0, v0);
            tmem_ld32(taddr + 256 + c0, v1);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (c0 + 8 * i >= N) continue;          // N is a multiple of 16: never write a chunk past the layer's columns
              const float4 b0 = *reinterpret_cast<const float4*>(&bias_s[l][c0 + 8 * i]);
              const float4 b1 = *reinterpret_cast<const float4*>(&bias_s[l][c0 + 8 * i + 4]);
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
              float f0[8], f1[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                f0[q] = rc0.valid ? fmaxf(__uint_as_float(v0[8 * i + q]) + bb[q], 0.0f) : 0.0f;
                f1[q] = rc1.valid ? fmaxf(__uint_as_float(v1[8 * i + q]) + bb[q], 0.0f) : 0.0f;
              }
              store8_bf16(A, CH, rc0.row, (c0 >> 3) + i, f0);
              store8_bf16(A, CH, rc1.row, (c0 >> 3) + i, f1);
            }
          }
        } else {
          for (int j = 0; j < 2; ++j) {
            const RowCtx rc = rowctx(j, tile);
            const uint32_t taddr = tbase + lane_addr + 256 * j;
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
        if (p.dbg && blockIdx.x == 0 && tid == 0 && en < 124) p.dbg[en++] = clock64();   // epilogue done
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
  return (size_t)wide_a_bytes(g) + 3 * kWideStage <= (size_t)220 * 1024;
}

int launch_tc_wide_fwd(const void* x, int x_dtype, const float* params, const char* packed, const int32_t* lengths, float* y,
                       int B, int T, int apply_mask, float out_scale, const Geo& g, cudaStream_t stream) {
  WideArgs p{};
  p.x = x; p.x_dtype = x_dtype; p.params = params; p.packed = packed; p.lengths = lengths; p.y = y;
  p.B = B; p.T = T; p.apply_mask = apply_mask; p.out_scale = out_scale; p.geo = g;
  p.dbg = g_dbg_timing;
  p.gh = 258 / (T + 2);
  p.n_tiles = (B + p.gh - 1) / p.gh;
  p.a_bytes = wide_a_bytes(g);
  int S = (int)(((size_t)220 * 1024 - p.a_bytes) / kWideStage);
  if (S > kWideMaxStages) S = kWideMaxStages;
  if (S < 3) { set_error("wide tensor-core forward: C=%d leaves no room for the weight ring", g.C); return B2H_ESHAPE; }
  p.nstage = S;
  const size_t smem = (size_t)p.a_bytes + (size_t)S * kWideStage;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(conv_tc_wide_fwd_kernel), smem)) return rc;
  int grid = num_sms();
  if (grid > p.n_tiles) grid = p.n_tiles;
  // weight tensor map: the forward UMMA sections [tf_off[0], td_off[0]) as rows of 64 bf16 (128 B), box = 64 rows
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || !fn) { cudaGetLastError(); set_error("cuTensorMapEncodeTiled entry point unavailable"); return B2H_ECUDA; }
    encode = (EncodeTiledFn)fn;
  }
  CUtensorMap wmap;
  {
    const cuuint64_t gdim[2] = {64, (cuuint64_t)((g.td_off[0] - g.tf_off[0]) >> 7)};
    const cuuint64_t gstride[1] = {128};
    const cuuint32_t box[2] = {64, kWideBoxRows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&wmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<char*>(packed) + g.tf_off[0], gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return B2H_ECUDA; }
  }
  conv_tc_wide_fwd_kernel<<<grid, kWideThreads, smem, stream>>>(p, wmap);
  count_launch();
  return check_launch("conv_tc_wide_fwd_kernel");
}

}  // namespace b2h
