// conv_tc.cu -- horizontally-circular implicit-GEMM convolution on tcgen05 tensor cores.
//
// Replaces `Conv2d._conv_forward` (reference `ldm/utils.py:40-58`, twin
// `vae/sgm/modules/diffusionmodules/model.py:93-108`): F.pad(circular, W) + F.pad(zeros, H) +
// F.conv2d(pad 0) -- two materialised padded copies and a cuDNN call -- with one kernel:
//
//   D[m, n] = sum_{tap, c} A_tap[m, c] * Wt[tap][n][c]      m = output pixel, n = output channel
//
// * activations are channels-last fp16 (B, W, H, C); a 128-pixel M tile is 128/Ho whole azimuth
//   columns.  For each (tap, 64-channel chunk) the producer warp issues one TMA box per column:
//   the wrap on W is a modular column coordinate, the zero pad on H is TMA out-of-bounds fill, and
//   stride 2 is the tensor map's element stride -- no padded copy ever exists;
// * weights [tap][Cout][Cin] fp16 arrive by TMA as the K-major B operand;
// * both land SWIZZLE_128B in a multi-stage mbarrier ring; one thread issues tcgen05.mma
//   (M=128, N=BLOCK_N, K=16) accumulating fp32 in TMEM; tcgen05.commit frees the stage;
// * 4 epilogue warps read TMEM (tcgen05.ld), add bias + time-embedding + residual and store fp32
//   channels-last (or atomically accumulate when the K loop is split across CTAs).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace rldm {

static long long* g_conv_dbg = nullptr;

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                       // fp16 elements = one 128 B swizzle row
constexpr int kABytes = kBlockM * kBlockK * 2;    // 16 KB

struct ConvParams {
  const float* bias;
  const float* temb;
  const float* residual;
  float* out;
  int temb_stride;
  int M_total;      // B*Wo*Ho
  int Wo, Ho, W_in;
  int pix_per_img;  // Wo*Ho
  int Cout;
  int ks, stride, pad_lo, circular;
  int total_iters;  // main_iters + shortcut chunks
  int main_iters;   // (Cin/64) * ks*ks: K steps of the convolution proper; the rest are the fused 1x1 shortcut's
  double* stats;    // optional GroupNorm moments of the output: [B][stats_G][2]
  int stats_cpg, stats_G;
  // halo-reuse 3x3 kernel only
  int a_part_bytes; // bytes of one operand part of an A stage: (MT*128 + 2*Ho) rows x 128 B
  int nb_stages;    // depth of the weight (B) ring
  int units;        // (Cin/64) * 3 : one unit = (channel chunk, kernel column tj) = 3 taps
  long long* dbg;   // optional: 8 clock64 timestamps written by CTA (0,0,0) (profiling aid, normally NULL)
  float* ws;        // optional split-K workspace in global memory: [tile][rank][128][BLOCK_N] fp32 partial tiles
};

// Tensor maps of one launch.  a/alo/b: activation (hi, lo) and weights of the convolution.  a2/a2lo/b2: operand and
// weights of an optional 1x1 convolution over a second tensor with the same spatial grid whose product is
// accumulated into the same tile (ResnetBlock2D's conv_shortcut folded into conv2's K loop).
struct ConvMaps {
  CUtensorMap a, alo, b, a2, a2lo, b2;
};

__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------
// Epilogue of one 128 x BLOCK_N accumulator tile, called by ALL 192 threads (warps 0/1 only take part in the
// barriers).  tmem_acc = TMEM address of the accumulator (lane 0, first column); m0 = first output pixel.
template <int BLOCK_N>
__device__ __forceinline__ void epilogue_tile(uint8_t* smem, uint32_t tmem_acc, int m0, int n0, const ConvParams& p,
                                              int warp, int lane) {
  constexpr int kStagePitch = BLOCK_N + 4;                     // floats per row of the epilogue staging tile
  if (warp >= 2) {
    // ===================== epilogue phase 1: TMEM -> registers -> shared staging tile =========
    // The pipeline stages are dead once tmem_full fires (all TMA writes consumed, all MMA reads done), so the
    // fp32 accumulator tile [128][BLOCK_N] is staged over them (row pitch +4 floats: conflict-free float4).
    const int q = warp & 3;                 // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;
    float* stage_row = reinterpret_cast<float*>(smem) + row * kStagePitch;
#pragma unroll 1
    for (int nc = 0; nc < BLOCK_N / 32; ++nc) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_acc + (static_cast<uint32_t>(q * 32) << 16) + nc * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint4*>(stage_row + nc * 32 + 4 * j) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
    }
  }
  // ======================= epilogue phase 2: (cluster) reduce + bias/temb/residual + stats + store ============
  // split-K: the `nsplit` CTAs of a cluster (same tile, different K slices) each staged a partial tile; CTA `rank`
  // now owns rows [rank*128/nsplit, ...) and sums them over all ranks through distributed shared memory in a
  // fixed order (deterministic, no atomics, no zero-fill).  nsplit == 1: same code on the local tile.
  const int nsplit = gridDim.z;
  const bool dbg = p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 64;
  if (dbg) p.dbg[6] = clock64();
  tc_fence_before();
  if (nsplit > 1) {
    cluster_sync_all();
  } else {
    __syncthreads();
  }
  if (dbg) p.dbg[7] = clock64();
  if (warp >= 2) {
    const int ew = warp - 2;                              // 0..3
    constexpr int kLanesPerRow = BLOCK_N / 4;             // 32 (BN=128) or 16 (BN=64)
    constexpr int kRowsPerIter = 32 / kLanesPerRow;       // 1 or 2
    const int rows_cta = kBlockM / nsplit;                // rows this CTA finalises
    const int rows_warp = rows_cta / 4;                   // contiguous rows per epilogue warp (>= 4)
    const int r_begin = blockIdx.z * rows_cta + ew * rows_warp;
    const int col = (lane % kLanesPerRow) * 4;
    const int rsub = lane / kLanesPerRow;
    const uint32_t stage_u32 = smem_u32(smem);
    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.bias) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + col));
    // all rows of one warp lie in one image (pix_per_img is a power of two >= 64 >= rows_warp*kRowsPerIter... see host)
    const int m_first = m0 + r_begin;
    const int bimg = min(m_first, p.M_total - 1) / p.pix_per_img;
    if (p.temb) {
      const float4 t4 = __ldg(reinterpret_cast<const float4*>(p.temb + static_cast<size_t>(bimg) * p.temb_stride + n0 + col));
      bias4.x += t4.x; bias4.y += t4.y; bias4.z += t4.z; bias4.w += t4.w;
    }
    float s01 = 0.f, q01 = 0.f, s23 = 0.f, q23 = 0.f;     // moments of channel pairs (0,1) and (2,3)
    // The residual rows come from L2 (~800 cycles): issue up to kU row loads per lane before consuming any.
    constexpr int kU = 8;
    for (int rr0 = rsub; rr0 < rows_warp; rr0 += kRowsPerIter * kU) {
      float4 res[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int rr = rr0 + u * kRowsPerIter;
        const int m = m0 + r_begin + rr;
        res[u] = bias4;
        if (p.residual && rr < rows_warp && m < p.M_total) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(p.residual + static_cast<size_t>(m) * p.Cout + n0 + col));
          res[u].x += t.x; res[u].y += t.y; res[u].z += t.z; res[u].w += t.w;
        }
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int rr = rr0 + u * kRowsPerIter;
        const int m = m0 + r_begin + rr;
        if (rr < rows_warp) {
          const int r = r_begin + rr;
          float4 v = res[u];
          if (nsplit > 1) {
            const uint32_t local = stage_u32 + (r * kStagePitch + col) * 4;
            for (int sidx = 0; sidx < nsplit; ++sidx) {
              const float4 t = ld_dsmem_f4(mapa_u32(local, sidx));
              v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
            }
          } else {
            const float4 t = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(smem) + r * kStagePitch + col);
            v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
          }
          if (m < p.M_total) {
            *reinterpret_cast<float4*>(p.out + static_cast<size_t>(m) * p.Cout + n0 + col) = v;
            s01 += v.x + v.y; q01 += v.x * v.x + v.y * v.y;
            s23 += v.z + v.w; q23 += v.z * v.z + v.w * v.w;
          }
        }
      }
    }
    if (dbg) p.dbg[8] = clock64();
    if (p.stats) {
      // Moments of the finished output per (image, channel PAIR): lanes -> the 4 epilogue warps (shared memory) ->
      // ONE double atomic per (pair, moment) and image for the whole CTA.  Pairs are the finest granularity any
      // consumer GroupNorm needs (its groups, also over a skip concat, are unions of whole pairs).
      __shared__ float red_s[4][BLOCK_N / 2], red_q[4][BLOCK_N / 2];
      __shared__ int red_b[4];
      if (kRowsPerIter == 2) {
        s01 += __shfl_xor_sync(0xffffffffu, s01, 16); q01 += __shfl_xor_sync(0xffffffffu, q01, 16);
        s23 += __shfl_xor_sync(0xffffffffu, s23, 16); q23 += __shfl_xor_sync(0xffffffffu, q23, 16);
      }
      const bool writer_row = (kRowsPerIter == 1) || rsub == 0;
      constexpr int nslots = BLOCK_N / 2;                 // one slot per channel PAIR of the tile
      if (writer_row) {
        red_s[ew][col / 2] = s01; red_q[ew][col / 2] = q01;
        red_s[ew][col / 2 + 1] = s23; red_q[ew][col / 2 + 1] = q23;
      }
      if (lane == 0) red_b[ew] = m_first < p.M_total ? bimg : -1;
      asm volatile("bar.sync 1, 128;" ::: "memory");      // the 4 epilogue warps only
      const int t = ew * 32 + lane;
      if (t < nslots) {
        const int g = n0 / 2 + t;
        double ds = 0.0, dq = 0.0;
        int cur = red_b[0];
        for (int e = 0; e < 4; ++e) {
          const int be = red_b[e];
          if (be != cur) {
            if (cur >= 0) {
              double* st = p.stats + (static_cast<size_t>(cur) * p.stats_G + g) * 2;
              atomicAdd(st, ds); atomicAdd(st + 1, dq);
            }
            cur = be; ds = 0.0; dq = 0.0;
          }
          ds += static_cast<double>(red_s[e][t]); dq += static_cast<double>(red_q[e][t]);
        }
        if (cur >= 0) {
          double* st = p.stats + (static_cast<size_t>(cur) * p.stats_G + g) * 2;
          atomicAdd(st, ds); atomicAdd(st + 1, dq);
        }
      }
    }
  }
  if (nsplit > 1) cluster_sync_all();     // nobody moves on while a peer may still read its staging tile
  else if (warp >= 2) asm volatile("bar.sync 1, 128;" ::: "memory");   // staging tile may be reused by a next pass
}

// TERMS = 1: D += A*W with fp16 operands (11-bit significands).
// TERMS = 3: split-fp16 ("fp16x3"): A = Ah + Al, W = Wh + Wl (each part fp16), D += Ah*Wh + Al*Wh + Ah*Wl --
//            ~22-bit operand significands on the fp16 tensor pipe, fp32 accumulation in TMEM.  This is the
//            default: it keeps 20-step trajectories within the 1e-3 parity tolerance with >100x margin.
// NSPLIT   : CTAs of the cluster that share one output tile (split K); == gridDim.z.
//
// Latency structure (these layers are small: the whole kernel is a handful of microseconds, so every serial
// round trip counts).  The four epilogue warps are idle during the K loop, so they fetch bias, time embedding and
// ALL residual rows they will need into registers right after griddepcontrol.wait -- the L2 latency of the
// residual hides under the mainloop.  The split-K reduction pulls the partial tiles of the peer CTAs through
// distributed shared memory in batches of 16 independent 16 B loads per lane (fixed summation order:
// deterministic), instead of one dependent load at a time.  DSMEM moves only ~17-21 B/clk per SM, so a cluster of 8
// needs ~3000 cycles to pull its 56 KB; when the caller provides a global workspace (p.ws) the partial tiles go
// through L2 instead (written straight from TMEM, read back at ~64 B/clk per SM after the cluster barrier; the
// barrier's release/acquire at cluster scope orders the global stores) -- same fixed summation order.  (Measured
// slower than DSMEM on the C3 shapes; the engine leaves p.ws NULL unless RLDM_SPLITK_VIA_L2=1.)
template <int BLOCK_N, int STAGES, int TERMS, int NSPLIT>
__global__ void __launch_bounds__(192, 1)
conv_tc_kernel(const __grid_constant__ ConvMaps tm, const ConvParams p) {
  constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  constexpr int kParts = TERMS == 1 ? 1 : 2;
  constexpr int kStageBytes = kParts * (kABytes + kBBytes);   // [A_hi][A_lo][B_hi][B_lo]
  constexpr int kBOff = kParts * kABytes;
  constexpr int kStagePitch = BLOCK_N + 4;                     // floats per row of the epilogue staging tile
  static_assert(kBlockM * kStagePitch * 4 <= STAGES * kStageBytes, "staging tile must fit in the pipeline stages");
  // epilogue geometry: a warp instruction covers kRowsPerIter rows of BLOCK_N floats (float4 per lane)
  constexpr int kLanesPerRow = BLOCK_N / 4;               // 32 (BN=128) or 16 (BN=64)
  constexpr int kRowsPerIter = 32 / kLanesPerRow;         // 1 or 2
  constexpr int kRowsCta = kBlockM / NSPLIT;              // rows this CTA finalises
  constexpr int kRowsWarp = kRowsCta / 4;                 // contiguous rows per epilogue warp
  constexpr int kPerLane = kRowsWarp / kRowsPerIter;      // float4 per lane: 32, 16, 8, 4 (BN=128); half for BN=64
  static_assert(kPerLane >= 1, "tile too small for this split");
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B atoms need 1024 B alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * kStageBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float* red_s = reinterpret_cast<float*>(tmem_ptr + 4);       // [4][BLOCK_N/2]
  float* red_q = red_s + 4 * (BLOCK_N / 2);
  int* red_b = reinterpret_cast<int*>(red_q + 4 * (BLOCK_N / 2));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kBlockM;
  const int n0 = blockIdx.y * BLOCK_N;
  const int it0 = static_cast<int>(static_cast<long long>(blockIdx.z) * p.total_iters / NSPLIT);
  const int it1 = static_cast<int>(static_cast<long long>(blockIdx.z + 1) * p.total_iters / NSPLIT);
  const int n_it = it1 - it0;                    // >= 1: the host keeps NSPLIT <= total_iters

  const bool dbg = p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  if (dbg && threadIdx.x == 0) p.dbg[0] = clock64();
  pdl_trigger_conv_early();     // let the next kernel's CTAs launch and run their prologue while this grid drains
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm.a);
    if (TERMS > 1) tma_prefetch_desc(&tm.alo);
    tma_prefetch_desc(&tm.b);
    if (p.total_iters > p.main_iters) {
      tma_prefetch_desc(&tm.a2);
      if (TERMS > 1) tma_prefetch_desc(&tm.a2lo);
      tma_prefetch_desc(&tm.b2);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<BLOCK_N>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);   // shfl: tells the compiler it is warp-uniform (UR, no per-MMA R2UR loop)
  // Weights never depend on the previous kernel: the producer arms the first STAGES barriers and starts their
  // weight tiles BEFORE griddepcontrol.wait, so they land while the previous grid is still draining.
  // weight tiles of K step `it` (hi plane, then lo plane: rows [taps*Cout, 2*taps*Cout) of the packed tensor)
  auto load_weights = [&](int it, uint32_t b_dst, uint64_t* bar) {
    const int taps = p.ks * p.ks;
    if (it < p.main_iters) {
      const int chunk = it / taps;
      const int tap = it - chunk * taps;
      tma_load_2d(b_dst, &tm.b, bar, chunk * kBlockK, tap * p.Cout + n0);
      if (TERMS > 1) tma_load_2d(b_dst + kBBytes, &tm.b, bar, chunk * kBlockK, (taps + tap) * p.Cout + n0);
    } else {
      const int chunk = it - p.main_iters;
      tma_load_2d(b_dst, &tm.b2, bar, chunk * kBlockK, n0);
      if (TERMS > 1) tma_load_2d(b_dst + kBBytes, &tm.b2, bar, chunk * kBlockK, p.Cout + n0);
    }
  };
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < n_it && i < STAGES; ++i) {
      mbar_arrive_expect_tx(&full_bar[i], kStageBytes);
      load_weights(it0 + i, smem_u32(smem + i * kStageBytes) + kBOff, &full_bar[i]);
    }
  }
  pdl_wait();        // everything above overlapped the previous kernel; below we touch its outputs
  if (dbg && threadIdx.x == 0) p.dbg[1] = clock64();

  // epilogue coordinates (meaningful for warps >= 2)
  const int ew = (warp - 2) & 3;
  const int r_begin = blockIdx.z * kRowsCta + ew * kRowsWarp;     // first tile row this warp finalises
  const int col = (lane % kLanesPerRow) * 4;
  const int rsub = lane / kLanesPerRow;
  const int m_first = m0 + r_begin;
  // all rows of one warp lie in one image (pix_per_img is a power of two >= 64, or whole images per tile row group)
  const int bimg = min(m_first, p.M_total - 1) / p.pix_per_img;
  float4 res[kPerLane];

  if (warp == 0) {
    // ===================== TMA producer (whole warp: lanes share the column loads) =============
    const int taps = p.ks * p.ks;
    // tile = `ncols` whole azimuth columns of `nb` consecutive images (nb > 1 only when an image has < 128 px)
    const int q0 = m0 / p.Ho;                     // global column index of the tile's first column
    const int b0 = q0 / p.Wo;
    const int wo0 = q0 - b0 * p.Wo;
    for (int i = 0; i < n_it; ++i) {
      const int s = i % STAGES;
      const uint32_t ph = (i / STAGES) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1);
      const int it = it0 + i;
      const bool main = it < p.main_iters;
      const int chunk = main ? it / taps : it - p.main_iters;
      const int tap = it - chunk * taps;
      const int ti = tap / p.ks, tj = tap - ti * p.ks;
      const uint32_t a_dst = smem_u32(smem + s * kStageBytes);
      if (lane == 0 && i >= STAGES) {   // (the first STAGES weight tiles were issued before griddepcontrol.wait)
        mbar_arrive_expect_tx(&full_bar[s], kStageBytes);
        load_weights(it, a_dst + kBOff, &full_bar[s]);
      }
      // ONE box per operand part: the activation tensor is W-padded (halo columns hold the circular wrap,
      // written by rldm_prep), H zero padding is TMA out-of-bounds fill, stride 2 is the map's element stride.
      // Shortcut K steps read the centre tap of the second tensor (1x1, stride 1, same grid as the output).
      const int h_in = main ? tj - p.pad_lo : 0;
      const int w_in = main ? p.stride * wo0 + ti - p.pad_lo + 1 : wo0 + 1;
      if (lane == (TERMS == 1 ? 0 : 1)) tma_load_4d(a_dst, main ? &tm.a : &tm.a2, &full_bar[s], chunk * kBlockK, h_in, w_in, b0);
      if (TERMS > 1 && lane == 2)
        tma_load_4d(a_dst + kABytes, main ? &tm.alo : &tm.a2lo, &full_bar[s], chunk * kBlockK, h_in, w_in, b0);
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) ============================================
    if (elect_one()) {   // elect.sync: ptxas knows exactly one lane is active (no per-MMA waterfall loop)
      constexpr uint32_t idesc = umma_idesc_f16(kBlockM, BLOCK_N);
      for (int i = 0; i < n_it; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (dbg && i == 0) p.dbg[2] = clock64();
        const uint32_t a_addr = smem_u32(smem + s * kStageBytes);
        const uint64_t a_desc = umma_desc_sw128(a_addr);
        const uint64_t b_desc = umma_desc_sw128(a_addr + kBOff);
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k) {
          // advancing K by 16 fp16 = 32 B inside the 128 B swizzle row: +2 in the (addr>>4) field
          umma_f16(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc, (i | k) != 0);
        }
        if (TERMS > 1) {
          const uint64_t al_desc = umma_desc_sw128(a_addr + kABytes);
          const uint64_t bl_desc = umma_desc_sw128(a_addr + kBOff + kBBytes);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            umma_f16(tmem_base, al_desc + 2 * k, b_desc + 2 * k, idesc, 1u);   // A_lo * W_hi
            umma_f16(tmem_base, a_desc + 2 * k, bl_desc + 2 * k, idesc, 1u);   // A_hi * W_lo
          }
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(tmem_full_bar);
      if (dbg) p.dbg[3] = clock64();
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps: prefetch, then TMEM -> shared staging tile ===========
    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.bias) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + col));
    if (p.temb) {
      const float4 t4 = __ldg(reinterpret_cast<const float4*>(p.temb + static_cast<size_t>(bimg) * p.temb_stride + n0 + col));
      bias4.x += t4.x; bias4.y += t4.y; bias4.z += t4.z; bias4.w += t4.w;
    }
#pragma unroll
    for (int u = 0; u < kPerLane; ++u) {
      const int m = m_first + u * kRowsPerIter + rsub;
      res[u] = bias4;
      if (p.residual && m < p.M_total) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p.residual + static_cast<size_t>(m) * p.Cout + n0 + col));
        res[u].x += t.x; res[u].y += t.y; res[u].z += t.z; res[u].w += t.w;
      }
    }
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    if (dbg && threadIdx.x == 64) p.dbg[4] = clock64();
    // The pipeline stages are dead once tmem_full fires (all TMA writes consumed, all MMA reads done), so the
    // fp32 accumulator tile [128][BLOCK_N] is staged over them (row pitch +4 floats: conflict-free float4).
    const int q = warp & 3;                 // TMEM lane quadrant this warp may access
    const bool via_l2 = NSPLIT > 1 && p.ws != nullptr;
    float* stage_row = via_l2
        ? p.ws + ((static_cast<size_t>(blockIdx.y * gridDim.x + blockIdx.x) * NSPLIT + blockIdx.z) * kBlockM + q * 32 + lane) * BLOCK_N
        : reinterpret_cast<float*>(smem) + (q * 32 + lane) * kStagePitch;
#pragma unroll 1
    for (int nc = 0; nc < BLOCK_N / 32; ++nc) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + nc * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint4*>(stage_row + nc * 32 + 4 * j) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
    }
    if (dbg && threadIdx.x == 64) p.dbg[6] = clock64();
  }
  pdl_trigger_conv_late();      // mainloop issued / accumulator staged: dependents may start their prologue
  // ======================= (cluster) reduce + stats + store =====================================================
  // split-K: the NSPLIT CTAs of a cluster (same tile, different K slices) each staged a partial tile; CTA `rank`
  // owns rows [rank*128/NSPLIT, ...) and sums them over all ranks through distributed shared memory in a fixed
  // order (deterministic, no atomics, no zero-fill).
  tc_fence_before();
  if (NSPLIT > 1) {
    cluster_sync_all();
  } else if (warp >= 2) {
    asm volatile("bar.sync 1, 128;" ::: "memory");        // only the epilogue warps touch the staging tile
  }
  if (warp >= 2) {
    if (dbg && threadIdx.x == 64) p.dbg[7] = clock64();
    float s01 = 0.f, q01 = 0.f, s23 = 0.f, q23 = 0.f;     // moments of channel pairs (0,1) and (2,3)
    const uint32_t stage_u32 = smem_u32(smem);
    if (NSPLIT == 1) {
#pragma unroll
      for (int u = 0; u < kPerLane; ++u) {
        const int r = r_begin + u * kRowsPerIter + rsub;
        const int m = m0 + r;
        const float4 t = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(smem) + r * kStagePitch + col);
        float4 v = res[u];
        v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
        if (m < p.M_total) {
          *reinterpret_cast<float4*>(p.out + static_cast<size_t>(m) * p.Cout + n0 + col) = v;
          s01 += v.x + v.y; q01 += v.x * v.x + v.y * v.y;
          s23 += v.z + v.w; q23 += v.z * v.z + v.w * v.w;
        }
      }
    } else {
      constexpr int kUB = (16 / NSPLIT) < kPerLane ? (16 / NSPLIT) : kPerLane;   // rows per batch of <= 16 loads
      uint32_t rbase[NSPLIT];
#pragma unroll
      for (int sidx = 0; sidx < NSPLIT; ++sidx) rbase[sidx] = mapa_u32(stage_u32, sidx);
#pragma unroll
      for (int u0 = 0; u0 < kPerLane; u0 += kUB) {
        float4 part[kUB][NSPLIT];
#pragma unroll
        for (int ub = 0; ub < kUB; ++ub) {
          const int r = r_begin + (u0 + ub) * kRowsPerIter + rsub;
          if (p.ws != nullptr) {             // partial tiles through L2 (written by the peers before the cluster barrier)
            const float* wrow = p.ws + (static_cast<size_t>(blockIdx.y * gridDim.x + blockIdx.x) * NSPLIT * kBlockM + r) * BLOCK_N + col;
#pragma unroll
            for (int sidx = 0; sidx < NSPLIT; ++sidx)
              part[ub][sidx] = __ldcg(reinterpret_cast<const float4*>(wrow + static_cast<size_t>(sidx) * kBlockM * BLOCK_N));
          } else {
            const uint32_t off = static_cast<uint32_t>(r * kStagePitch + col) * 4u;
#pragma unroll
            for (int sidx = 0; sidx < NSPLIT; ++sidx) part[ub][sidx] = ld_dsmem_f4(rbase[sidx] + off);
          }
        }
#pragma unroll
        for (int ub = 0; ub < kUB; ++ub) {
          const int m = m0 + r_begin + (u0 + ub) * kRowsPerIter + rsub;
          float4 v = res[u0 + ub];
#pragma unroll
          for (int sidx = 0; sidx < NSPLIT; ++sidx) {
            v.x += part[ub][sidx].x; v.y += part[ub][sidx].y; v.z += part[ub][sidx].z; v.w += part[ub][sidx].w;
          }
          if (m < p.M_total) {
            *reinterpret_cast<float4*>(p.out + static_cast<size_t>(m) * p.Cout + n0 + col) = v;
            s01 += v.x + v.y; q01 += v.x * v.x + v.y * v.y;
            s23 += v.z + v.w; q23 += v.z * v.z + v.w * v.w;
          }
        }
      }
    }
    if (dbg && threadIdx.x == 64) p.dbg[8] = clock64();
    if (p.stats) {
      // Moments of the finished output per (image, channel PAIR): lanes -> the 4 epilogue warps (shared memory) ->
      // ONE double atomic per (pair, moment) and image for the whole CTA.  Pairs are the finest granularity any
      // consumer GroupNorm needs (its groups, also over a skip concat, are unions of whole pairs).
      if (kRowsPerIter == 2) {
        s01 += __shfl_xor_sync(0xffffffffu, s01, 16); q01 += __shfl_xor_sync(0xffffffffu, q01, 16);
        s23 += __shfl_xor_sync(0xffffffffu, s23, 16); q23 += __shfl_xor_sync(0xffffffffu, q23, 16);
      }
      constexpr int nslots = BLOCK_N / 2;                 // one slot per channel PAIR of the tile
      if (rsub == 0) {
        red_s[ew * nslots + col / 2] = s01; red_q[ew * nslots + col / 2] = q01;
        red_s[ew * nslots + col / 2 + 1] = s23; red_q[ew * nslots + col / 2 + 1] = q23;
      }
      if (lane == 0) red_b[ew] = m_first < p.M_total ? bimg : -1;
      asm volatile("bar.sync 1, 128;" ::: "memory");      // the 4 epilogue warps only
      const int t = ew * 32 + lane;
      if (t < nslots) {
        const int g = n0 / 2 + t;
        double ds = 0.0, dq = 0.0;
        int cur = red_b[0];
        for (int e = 0; e < 4; ++e) {
          const int be = red_b[e];
          if (be != cur) {
            if (cur >= 0) {
              double* st = p.stats + (static_cast<size_t>(cur) * p.stats_G + g) * 2;
              atomicAdd(st, ds); atomicAdd(st + 1, dq);
            }
            cur = be; ds = 0.0; dq = 0.0;
          }
          ds += static_cast<double>(red_s[e * nslots + t]); dq += static_cast<double>(red_q[e * nslots + t]);
        }
        if (cur >= 0) {
          double* st = p.stats + (static_cast<size_t>(cur) * p.stats_G + g) * 2;
          atomicAdd(st, ds); atomicAdd(st + 1, dq);
        }
      }
    }
    if (dbg && threadIdx.x == 64) p.dbg[5] = clock64();
  }
  if (NSPLIT > 1) cluster_sync_all();     // nobody exits while a peer may still read its staging tile
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<BLOCK_N>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// Persistent variant of the per-tap kernel for layers with more tiles than SMs (no K split).
//
// Timeline stamps of the kernel above show the 128x128 epilogue (64 KB of residual reads + 64 KB of stores per
// CTA, issued by all CTAs of a wave at once) costing ~45 % of the mainloop time.  Here one CTA per SM walks over
// tiles; the fp32 accumulator is DOUBLE-BUFFERED in TMEM (2 x BLOCK_N columns), so the four epilogue warps drain
// tile k (TMEM -> 32-column shared staging slab -> coalesced global, bias/temb/residual/GroupNorm moments) while
// the producer and MMA warps are already running the K loop of tile k+1.  The staging slab is private to the
// epilogue (not aliased with the pipeline stages).
//
// MT = 2: a work unit is TWO consecutive 128-pixel M tiles that share every weight tile: a stage holds
// [A0_hi A0_lo A1_hi A1_lo B_hi B_lo] and feeds 2 x 12 MMAs.  The K loop of these kernels is bound by the operand
// stream into the SM (~64 B/clk: a 64 KB stage lands in ~1000 cycles while its 12 MMAs need 768); sharing B over two
// tiles cuts the bytes per MMA by 25 % (48 KB per 12 MMAs), which makes the loop MMA-bound, and it halves the number
// of units (256 tiles -> 128 units: one wave on 148 SMs instead of 1.73).  TMEM: 2 units x MT x BLOCK_N columns.
// KB = 32 (opt-in experiment): a stage carries half a 64-channel chunk (SWIZZLE_64B operand tiles, two K steps).  With
// MT = 2 a 64-wide stage is 96 KB and only two fit; 48 KB half-stages give a 4-deep ring, but measured slower.
//
// HALO = true (3x3, stride 1, symmetric pad, MT = 2, KB = 64): the A operand of the three taps ti = 0,1,2 of one
// kernel column tj is ONE shared-memory window of (2*128 + 2*Ho) pixels (the unit's columns plus one halo column on
// each side; tap ti = the window shifted by ti*Ho rows, a whole number of 1024 B swizzle atoms for Ho >= 8), so the
// A stream from L2 drops 3x.  Two rings: two A windows ([hi][lo], one per (64-channel chunk, tj)) and p.nb_stages
// weight entries of ONE operand part each (W_hi of a tap feeds A_hi W_hi + A_lo W_hi of both tiles and is released
// before W_lo is needed).  Per tap 72/3 + 32 = 56 KB instead of 96 KB per 24 MMAs: the K loop turns MMA-bound.
template <int BLOCK_N, int STAGES, int TERMS, int MT, int KB, bool HALO = false>
__global__ void __launch_bounds__(192, 1)
conv_tc_persistent_kernel(const __grid_constant__ ConvMaps tm, const ConvParams p) {
  static_assert(KB == 64 || KB == 32, "stage depth along K: one 128 B swizzle row or half of it");
  static_assert(!HALO || (MT == 2 && KB == 64), "halo windows: two M tiles per unit, 64-channel stages");
  constexpr int kMaxNB = 8;                                     // HALO: upper bound of the weight ring depth
  constexpr int kRingBars = HALO ? 4 + 2 * kMaxNB : 2 * STAGES;
  constexpr int kSub = kBlockK / KB;                            // stages per 64-channel chunk
  constexpr int kAB = kBlockM * KB * 2;                         // one operand part of one M tile
  constexpr int kBBytes = BLOCK_N * KB * 2;
  constexpr int kParts = TERMS == 1 ? 1 : 2;
  constexpr int kATile = kParts * kAB;                          // one M tile of a stage: [A_hi][A_lo]
  constexpr int kStageBytes = MT * kATile + kParts * kBBytes;
  constexpr int kBOff = MT * kATile;
  constexpr int kSlabPitch = 36;                                // floats per row of the 32-column staging slab
  constexpr int kSlabBytes = kBlockM * kSlabPitch * 4;          // 18 KB
  constexpr int kChunks = BLOCK_N / 32;
  constexpr int kAccCols = MT * BLOCK_N;                        // TMEM columns of one unit's accumulators
  static_assert(2 * kAccCols <= 512, "double-buffered accumulators must fit TMEM");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int a_stage = HALO ? kParts * p.a_part_bytes : 0;       // HALO: one A window stage, [hi][lo]
  const int NB = HALO ? p.nb_stages : 0;
  uint8_t* b_ring = smem + 2 * a_stage;                          // HALO: NB entries of kBBytes
  uint8_t* tail = HALO ? b_ring + NB * kBBytes : smem + STAGES * kStageBytes;
  float* slab = reinterpret_cast<float*>(tail);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail + kSlabBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* a_full = full_bar;                   // HALO: [2] [2] [kMaxNB] [kMaxNB]
  uint64_t* a_empty = a_full + 2;
  uint64_t* b_full = a_empty + 2;
  uint64_t* b_empty = b_full + kMaxNB;
  uint64_t* tmem_full = full_bar + kRingBars;    // [2]
  uint64_t* tmem_empty = tmem_full + 2;          // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* red_s = reinterpret_cast<float*>(tmem_ptr + 4);       // [4][BLOCK_N/2]
  float* red_q = red_s + 4 * (BLOCK_N / 2);
  int* red_b = reinterpret_cast<int*>(red_q + 4 * (BLOCK_N / 2));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_n = p.Cout / BLOCK_N;
  const int tiles_m = (p.M_total + kBlockM - 1) / kBlockM;      // a multiple of MT (host)
  const int total_units = (tiles_m / MT) * tiles_n;
  const int taps = p.ks * p.ks;
  const int n_it = p.total_iters * kSub;

  pdl_trigger_conv_early();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm.a);
    if (TERMS > 1) tma_prefetch_desc(&tm.alo);
    tma_prefetch_desc(&tm.b);
    if (p.total_iters > p.main_iters) {
      tma_prefetch_desc(&tm.a2);
      if (TERMS > 1) tma_prefetch_desc(&tm.a2lo);
      tma_prefetch_desc(&tm.b2);
    }
    for (int s = 0; s < kRingBars; ++s) mbar_init(&full_bar[s], 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4);          // one arrival per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<2 * kAccCols>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);   // shfl: tells the compiler it is warp-uniform (UR, no per-MMA R2UR loop)
  pdl_wait();

  if (HALO && warp == 0) {
    // ===================== TMA producer (halo windows): A ring of 2, weight ring of NB ============
    int ga = 0, gb = 0;
    const int chunks = p.units / 3;
    for (int t = blockIdx.x; t < total_units; t += gridDim.x) {
      if (t + static_cast<int>(gridDim.x) >= total_units) pdl_trigger_conv_late();
      const int um = t / tiles_n, tn = t - um * tiles_n;
      const int n0 = tn * BLOCK_N;
      const int q0 = um * (MT * kBlockM) / p.Ho;            // global column index of the unit's first column
      const int b0 = q0 / p.Wo, wo0 = q0 - b0 * p.Wo;       // wo0 == padded column of tap ti = 0
      for (int chunk = 0; chunk < chunks; ++chunk) {
        for (int tj = 0; tj < 3; ++tj, ++ga) {
          const int sa = ga & 1;
          mbar_wait(&a_empty[sa], ((ga >> 1) & 1) ^ 1);
          const uint32_t a_dst = smem_u32(smem + sa * a_stage);
          if (lane == 0) {
            mbar_arrive_expect_tx(&a_full[sa], a_stage);
            tma_load_4d(a_dst, &tm.a, &a_full[sa], chunk * kBlockK, tj - 1, wo0, b0);
            if (TERMS > 1) tma_load_4d(a_dst + p.a_part_bytes, &tm.alo, &a_full[sa], chunk * kBlockK, tj - 1, wo0, b0);
          }
          for (int e = 0; e < 3 * kParts; ++e, ++gb) {      // (ti, part): W_hi then W_lo of each tap
            const int sb = gb % NB;
            mbar_wait(&b_empty[sb], ((gb / NB) & 1) ^ 1);
            if (lane == 0) {
              const int ti = e / kParts, part = e - ti * kParts;
              const int tap = ti * 3 + tj;
              mbar_arrive_expect_tx(&b_full[sb], kBBytes);
              tma_load_2d(smem_u32(b_ring + sb * kBBytes), &tm.b, &b_full[sb], chunk * kBlockK,
                          (part * 9 + tap) * p.Cout + n0);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (HALO && warp == 1) {
    // ===================== MMA issuer (halo windows) ============================================
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_f16(kBlockM, BLOCK_N);
      int ga = 0, gb = 0, k = 0;
      for (int t = blockIdx.x; t < total_units; t += gridDim.x, ++k) {
        if (t + static_cast<int>(gridDim.x) >= total_units) pdl_trigger_conv_late();
        const int acc = k & 1;
        mbar_wait(&tmem_empty[acc], ((k >> 1) & 1) ^ 1);       // epilogue has drained this accumulator set
        tc_fence_after();
        for (int u = 0; u < p.units; ++u, ++ga) {
          const int sa = ga & 1;
          mbar_wait(&a_full[sa], (ga >> 1) & 1);
          const uint32_t a_base = smem_u32(smem + sa * a_stage);
          for (int ti = 0; ti < 3; ++ti) {
            // tap ti of this kernel column = the window shifted by ti columns (ti*Ho rows); tile mt starts mt*128
            // rows further down.  Both offsets are whole 1024 B swizzle atoms.
            const uint32_t a_tap = a_base + ti * p.Ho * 128;
            {
              const int sb = gb % NB;
              mbar_wait(&b_full[sb], (gb / NB) & 1);
              tc_fence_after();
              const uint64_t b_desc = umma_desc_sw128(smem_u32(b_ring + sb * kBBytes));
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
                const uint32_t d_tmem = tmem_base + acc * kAccCols + mt * BLOCK_N;
                const uint64_t a_desc = umma_desc_sw128(a_tap + mt * kBlockM * 128);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  umma_f16(d_tmem, a_desc + 2 * kk, b_desc + 2 * kk, idesc, (u | ti | kk) != 0);
                if (TERMS > 1) {
                  const uint64_t al_desc = umma_desc_sw128(a_tap + mt * kBlockM * 128 + p.a_part_bytes);
#pragma unroll
                  for (int kk = 0; kk < 4; ++kk) umma_f16(d_tmem, al_desc + 2 * kk, b_desc + 2 * kk, idesc, 1u);
                }
              }
              umma_commit(&b_empty[sb]);
              ++gb;
            }
            if (TERMS > 1) {
              const int sb = gb % NB;
              mbar_wait(&b_full[sb], (gb / NB) & 1);
              tc_fence_after();
              const uint64_t bl_desc = umma_desc_sw128(smem_u32(b_ring + sb * kBBytes));
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
                const uint32_t d_tmem = tmem_base + acc * kAccCols + mt * BLOCK_N;
                const uint64_t a_desc = umma_desc_sw128(a_tap + mt * kBlockM * 128);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) umma_f16(d_tmem, a_desc + 2 * kk, bl_desc + 2 * kk, idesc, 1u);
              }
              umma_commit(&b_empty[sb]);
              ++gb;
            }
          }
          umma_commit(&a_empty[sa]);
        }
        umma_commit(&tmem_full[acc]);
      }
    }
    __syncwarp();
  } else if (warp == 0) {
    // ===================== TMA producer: one continuous stage ring across units =================
    int g = 0;
    for (int t = blockIdx.x; t < total_units; t += gridDim.x) {
      if (t + static_cast<int>(gridDim.x) >= total_units) pdl_trigger_conv_late();
      const int um = t / tiles_n, tn = t - um * tiles_n;
      const int m0 = um * (MT * kBlockM), n0 = tn * BLOCK_N;
      int b0[MT], wo0[MT];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int q0 = (m0 + mt * kBlockM) / p.Ho;          // global column index of the tile's first column
        b0[mt] = q0 / p.Wo;
        wo0[mt] = q0 - b0[mt] * p.Wo;
      }
      for (int it = 0; it < n_it; ++it, ++g) {
        const int s = g % STAGES;
        mbar_wait(&empty_bar[s], ((g / STAGES) & 1) ^ 1);
        const int it64 = it / kSub, sub = it - it64 * kSub;      // 64-channel K step and the KB-wide part of it
        const bool main = it64 < p.main_iters;
        const int chunk = main ? it64 / taps : it64 - p.main_iters;
        const int tap = it64 - chunk * taps;
        const int ti = tap / p.ks, tj = tap - ti * p.ks;
        const int kc = chunk * kBlockK + sub * KB;                 // first channel of this stage
        const uint32_t a_dst = smem_u32(smem + s * kStageBytes);
        // shortcut K steps: centre tap of the second tensor (1x1, stride 1, same grid as the output)
        const int h_in = main ? tj - p.pad_lo : 0;
        const int w_off = main ? ti - p.pad_lo + 1 : 1;
        const int w_mul = main ? p.stride : 1;
        if (lane == 0) {
          mbar_arrive_expect_tx(&full_bar[s], kStageBytes);
          if (main) {
            tma_load_2d(a_dst + kBOff, &tm.b, &full_bar[s], kc, tap * p.Cout + n0);
            if (TERMS > 1)
              tma_load_2d(a_dst + kBOff + kBBytes, &tm.b, &full_bar[s], kc, (taps + tap) * p.Cout + n0);
          } else {
            tma_load_2d(a_dst + kBOff, &tm.b2, &full_bar[s], kc, n0);
            if (TERMS > 1) tma_load_2d(a_dst + kBOff + kBBytes, &tm.b2, &full_bar[s], kc, p.Cout + n0);
          }
        }
        // lanes 1.. : one box per (M tile, operand part)
        if (lane >= 1 && lane <= MT * kParts) {
          const int mt = (lane - 1) / kParts, part = (lane - 1) % kParts;
          const CUtensorMap* map = main ? (part ? &tm.alo : &tm.a) : (part ? &tm.a2lo : &tm.a2);
          tma_load_4d(a_dst + mt * kATile + part * kAB, map, &full_bar[s], kc, h_in, w_mul * wo0[mt] + w_off, b0[mt]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: alternates between the two TMEM accumulator sets ===========
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_f16(kBlockM, BLOCK_N);
      int g = 0, k = 0;
      for (int t = blockIdx.x; t < total_units; t += gridDim.x, ++k) {
        if (t + static_cast<int>(gridDim.x) >= total_units) pdl_trigger_conv_late();
        const int acc = k & 1;
        mbar_wait(&tmem_empty[acc], ((k >> 1) & 1) ^ 1);       // epilogue has drained this accumulator set
        tc_fence_after();
        for (int it = 0; it < n_it; ++it, ++g) {
          const int s = g % STAGES;
          mbar_wait(&full_bar[s], (g / STAGES) & 1);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * kStageBytes);
          auto desc = [](uint32_t addr) { return KB == 64 ? umma_desc_sw128(addr) : umma_desc_sw64(addr); };
          const uint64_t b_desc = desc(a_addr + kBOff);
          const uint64_t bl_desc = desc(a_addr + kBOff + kBBytes);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const uint32_t d_tmem = tmem_base + acc * kAccCols + mt * BLOCK_N;
            const uint64_t a_desc = desc(a_addr + mt * kATile);
#pragma unroll
            for (int kk = 0; kk < KB / 16; ++kk)
              umma_f16(d_tmem, a_desc + 2 * kk, b_desc + 2 * kk, idesc, (it | kk) != 0);
            if (TERMS > 1) {
              const uint64_t al_desc = desc(a_addr + mt * kATile + kAB);
#pragma unroll
              for (int kk = 0; kk < KB / 16; ++kk) {
                umma_f16(d_tmem, al_desc + 2 * kk, b_desc + 2 * kk, idesc, 1u);
                umma_f16(d_tmem, a_desc + 2 * kk, bl_desc + 2 * kk, idesc, 1u);
              }
            }
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tmem_full[acc]);
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps: drain unit k while unit k+1 is being computed =========
    const int ew = warp - 2;                  // rows ew*32 .. ew*32+31 of the tile in the coalesced phase
    const int q = warp & 3;                   // TMEM lane quadrant readable by this warp
    const int rsub = lane >> 3;               // 4 rows per warp instruction in the coalesced phase
    const int col = (lane & 7) * 4;           // float4 column inside the 32-column slab
    int k = 0;
    for (int t = blockIdx.x; t < total_units; t += gridDim.x, ++k) {
      const int acc = k & 1;
      const int um = t / tiles_n, tn = t - um * tiles_n;
      const int n0 = tn * BLOCK_N;
      mbar_wait(&tmem_full[acc], (k >> 1) & 1);
      if (t + static_cast<int>(gridDim.x) >= total_units) pdl_trigger_conv_late();
      tc_fence_after();
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
      const int m0 = um * (MT * kBlockM) + mt * kBlockM;
      const int m_first = m0 + ew * 32;
      const int bimg = min(m_first, p.M_total - 1) / p.pix_per_img;   // a warp's 32 rows lie in one image
      float sg[kChunks], qg[kChunks];         // per-chunk moments of this lane's first channel pair
      float s23c[kChunks], q23c[kChunks];     // second channel pair
      // residual rows of chunk nc+1 are fetched while chunk nc is drained (two chunks of loads in flight per lane)
      float4 res_nx[8];
      auto fetch_res = [&](int nc, float4 (&dst)[8]) {
        const int c = n0 + nc * 32 + col;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int m = m_first + u * 4 + rsub;
          dst[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (m < p.M_total) dst[u] = __ldg(reinterpret_cast<const float4*>(p.residual + static_cast<size_t>(m) * p.Cout + c));
        }
      };
      if (p.residual) fetch_res(0, res_nx);
#pragma unroll
      for (int nc = 0; nc < kChunks; ++nc) {
        // (1) accumulator slab -> registers -> shared (row = TMEM lane)
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + acc * kAccCols + mt * BLOCK_N + (static_cast<uint32_t>(q * 32) << 16) + nc * 32, r);
        tmem_ld_wait();
        float* srow = slab + (q * 32 + lane) * kSlabPitch;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(srow + 4 * j) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
        if (mt == MT - 1 && nc == kChunks - 1) {   // last TMEM read of this unit: hand the accumulators back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty[acc])) : "memory");
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        // (2) coalesced phase: this warp finalises rows ew*32.. of the tile, 4 rows x 128 B per instruction
        const int c = n0 + nc * 32 + col;
        float4 add4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias) add4 = __ldg(reinterpret_cast<const float4*>(p.bias + c));
        if (p.temb) {
          const float4 t4 = __ldg(reinterpret_cast<const float4*>(p.temb + static_cast<size_t>(bimg) * p.temb_stride + c));
          add4.x += t4.x; add4.y += t4.y; add4.z += t4.z; add4.w += t4.w;
        }
        float4 res[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          res[u] = add4;
          if (p.residual) { res[u].x += res_nx[u].x; res[u].y += res_nx[u].y; res[u].z += res_nx[u].z; res[u].w += res_nx[u].w; }
        }
        if (p.residual && nc + 1 < kChunks) fetch_res(nc + 1, res_nx);
        float s01 = 0.f, q01 = 0.f, s23 = 0.f, q23 = 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int rr = ew * 32 + u * 4 + rsub;
          const int m = m0 + rr;
          const float4 a4 = *reinterpret_cast<const float4*>(slab + rr * kSlabPitch + col);
          float4 v = res[u];
          v.x += a4.x; v.y += a4.y; v.z += a4.z; v.w += a4.w;
          if (m < p.M_total) {
            *reinterpret_cast<float4*>(p.out + static_cast<size_t>(m) * p.Cout + c) = v;
            s01 += v.x + v.y; q01 += v.x * v.x + v.y * v.y;
            s23 += v.z + v.w; q23 += v.z * v.z + v.w * v.w;
          }
        }
        // fold the 4 row-groups of the warp (lanes l, l+8, l+16, l+24 hold the same columns)
        s01 += __shfl_xor_sync(0xffffffffu, s01, 8); q01 += __shfl_xor_sync(0xffffffffu, q01, 8);
        s23 += __shfl_xor_sync(0xffffffffu, s23, 8); q23 += __shfl_xor_sync(0xffffffffu, q23, 8);
        s01 += __shfl_xor_sync(0xffffffffu, s01, 16); q01 += __shfl_xor_sync(0xffffffffu, q01, 16);
        s23 += __shfl_xor_sync(0xffffffffu, s23, 16); q23 += __shfl_xor_sync(0xffffffffu, q23, 16);
        sg[nc] = s01; qg[nc] = q01; s23c[nc] = s23; q23c[nc] = q23;
        asm volatile("bar.sync 1, 128;" ::: "memory");     // slab free for the next chunk / next tile
      }
      if (p.stats) {
        // per-CTA fold of the channel-pair moments, then one double atomic per (pair, moment) and image
        constexpr int nslots = BLOCK_N / 2;
        if (lane < 8) {
#pragma unroll
          for (int nc = 0; nc < kChunks; ++nc) {
            const int cc = nc * 32 + col;                  // column of this lane's float4 inside the tile
            red_s[ew * (BLOCK_N / 2) + cc / 2] = sg[nc];       red_q[ew * (BLOCK_N / 2) + cc / 2] = qg[nc];
            red_s[ew * (BLOCK_N / 2) + cc / 2 + 1] = s23c[nc]; red_q[ew * (BLOCK_N / 2) + cc / 2 + 1] = q23c[nc];
          }
        }
        if (lane == 0) red_b[ew] = m_first < p.M_total ? bimg : -1;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int tix = ew * 32 + lane;
        if (tix < nslots) {
          const int gidx = n0 / 2 + tix;
          double ds = 0.0, dq = 0.0;
          int cur = red_b[0];
          for (int e = 0; e < 4; ++e) {
            const int be = red_b[e];
            if (be != cur) {
              if (cur >= 0) {
                double* st = p.stats + (static_cast<size_t>(cur) * p.stats_G + gidx) * 2;
                atomicAdd(st, ds); atomicAdd(st + 1, dq);
              }
              cur = be; ds = 0.0; dq = 0.0;
            }
            ds += static_cast<double>(red_s[e * (BLOCK_N / 2) + tix]);
            dq += static_cast<double>(red_q[e * (BLOCK_N / 2) + tix]);
          }
          if (cur >= 0) {
            double* st = p.stats + (static_cast<size_t>(cur) * p.stats_G + gidx) * 2;
            atomicAdd(st, ds); atomicAdd(st + 1, dq);
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");     // red_* reusable by the next tile
      }
      }   // mt
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<2 * kAccCols>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// Persistent kernel with the operands' ROLES SWAPPED: D^T[Cout tile, pixels] = W[Cout tile, K] X^T[K, pixels].
//
// ncu on the kernel above: a 128x128x16 SS-form MMA reads 4 KB of A and 4 KB of B from shared memory in its 64
// clocks, i.e. 64 wavefronts of 128 B = the WHOLE shared-memory data pipe (l1tex__data_pipe_tc_wavefronts: 64 per
// MMA), so every TMA fill of the ring steals tensor cycles: 1.0 (reads) + 0.5 (fills) wavefronts per MMA clock =
// the measured ~2200 instead of 1536 cycles per K step, tensor pipe 54-69 % active.  Sharing a weight tile between
// two M tiles (MT = 2) saves L2 traffic but not shared-memory reads (each MMA re-reads its B tile).
// Here the 128 output CHANNELS are the M side (A operand = weight tile, 128 rows x 64 channels, K-major) and 256
// PIXELS the N side (B operand = the two 128-pixel activation tiles stored back to back = one 256-row K-major tile):
// one 128x256x16 MMA reads 4 + 8 KB in 128 clocks = 0.75 wavefronts per clock for the same FLOPs.  The accumulator
// is [channel lane][pixel column] (2 x 256 TMEM columns, double-buffered), which also makes the epilogue simpler:
// a thread owns ONE output channel, `tcgen05.ld` hands it 32 pixels, lanes = 32 consecutive channels of one pixel,
// so stores and residual loads are coalesced 128 B rows straight from registers (no shared staging slab, no
// bar.sync), bias/temb are per-thread scalars and the GroupNorm channel-pair moments are per-thread sums plus one
// shuffle.  Stage = [X_hi 256 rows][X_lo 256 rows][W_hi][W_lo] = 96 KB, two stages.
//
// HALO = true (3x3, stride 1, symmetric pad, Ho in 8..32, no fused shortcut): the pixel operand of the three taps
// ti = 0,1,2 of one kernel column tj is ONE window of (256 + 2*Ho) rows (the unit's columns plus one halo column on
// each side; tap ti = the window shifted by ti*Ho rows = whole 1024 B swizzle atoms), so the ring fills drop from
// 96 KB to 56 KB per tap (0.5 -> 0.29 wavefronts per MMA clock next to 0.75 of operand reads).  Two rings: two pixel
// windows ([hi][lo], one per (64-channel chunk, tj)) and p.nb_stages weight entries of ONE operand part each: W_hi of
// a tap feeds W_hi X_hi + W_hi X_lo and is released before W_lo is needed.
constexpr int kWtEpiWarps = 8;                  // two warps per TMEM lane quadrant, four 32-pixel chunks each
constexpr int kWtThreads = 64 + 32 * kWtEpiWarps;
template <int TERMS, bool HALO = false>
__global__ void __launch_bounds__(kWtThreads, 1)
conv_tc_wt_kernel(const __grid_constant__ ConvMaps tm, const ConvParams p) {
  constexpr int kCoutTile = 128, kPix = 256, STAGES = 2;
  constexpr int kMaxNW = 8;                                     // HALO: upper bound of the weight ring depth
  constexpr int kRingBars = HALO ? 4 + 2 * kMaxNW : 2 * STAGES;
  constexpr int kParts = TERMS == 1 ? 1 : 2;
  constexpr int kXPart = kPix * kBlockK * 2;                    // 32 KB: 256 pixel rows x 128 B
  constexpr int kWPart = kCoutTile * kBlockK * 2;               // 16 KB
  constexpr int kWOff = kParts * kXPart;
  constexpr int kStageBytes = kParts * (kXPart + kWPart);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int x_stage = HALO ? kParts * p.a_part_bytes : 0;       // HALO: one pixel-window stage, [hi][lo]
  const int NW = HALO ? p.nb_stages : 0;
  uint8_t* w_ring = smem + 2 * x_stage;                          // HALO: NW entries of kWPart
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(HALO ? w_ring + NW * kWPart : smem + STAGES * kStageBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* x_full = full_bar;                   // HALO: [2] [2] [kMaxNW] [kMaxNW]
  uint64_t* x_empty = x_full + 2;
  uint64_t* w_full = x_empty + 2;
  uint64_t* w_empty = w_full + kMaxNW;
  uint64_t* tmem_full = full_bar + kRingBars;    // [2]
  uint64_t* tmem_empty = tmem_full + 2;          // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_n = p.Cout / kCoutTile;
  const int total_units = (p.M_total / kPix) * tiles_n;
  const int taps = p.ks * p.ks;
  const int n_it = p.total_iters;

  // profiling aid (normally NULL): clock64 stamps of CTA 0 -- [0] entry, [1] setup done, [2] first stage landed,
  // [3] last MMA issued, [4] accumulators complete, [6] this warp's first chunk in registers, [8] stores issued, [5] done
  const bool dbg = p.dbg != nullptr && blockIdx.x == 0;
  if (dbg && threadIdx.x == 0) p.dbg[0] = clock64();
  pdl_trigger_conv_early();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm.a);
    if (TERMS > 1) tma_prefetch_desc(&tm.alo);
    tma_prefetch_desc(&tm.b);
    if (p.total_iters > p.main_iters) {
      tma_prefetch_desc(&tm.a2);
      if (TERMS > 1) tma_prefetch_desc(&tm.a2lo);
      tma_prefetch_desc(&tm.b2);
    }
    for (int s = 0; s < kRingBars; ++s) mbar_init(&full_bar[s], 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], kWtEpiWarps);   // one arrival per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<2 * kPix>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
  pdl_wait();
  if (dbg && threadIdx.x == 0) p.dbg[1] = clock64();

  if (HALO && warp == 0) {
    // ===================== TMA producer (pixel windows): window ring of 2, weight ring of NW =====
    int gx = 0, gw = 0;
    const int chunks = p.units / 3;
    for (int t = blockIdx.x; t < total_units; t += gridDim.x) {
      if (t + static_cast<int>(gridDim.x) >= total_units) pdl_trigger_conv_late();
      const int um = t / tiles_n, tn = t - um * tiles_n;
      const int n0 = tn * kCoutTile;
      const int q0 = um * kPix / p.Ho;                      // global column index of the unit's first column
      const int b0 = q0 / p.Wo, wo0 = q0 - b0 * p.Wo;       // wo0 == padded column of tap ti = 0
      for (int chunk = 0; chunk < chunks; ++chunk) {
        for (int tj = 0; tj < 3; ++tj, ++gx) {
          const int sx = gx & 1;
          mbar_wait(&x_empty[sx], ((gx >> 1) & 1) ^ 1);
          const uint32_t x_dst = smem_u32(smem + sx * x_stage);
          if (lane == 0) {
            mbar_arrive_expect_tx(&x_full[sx], x_stage);
            tma_load_4d(x_dst, &tm.a, &x_full[sx], chunk * kBlockK, tj - 1, wo0, b0);
            if (TERMS > 1) tma_load_4d(x_dst + p.a_part_bytes, &tm.alo, &x_full[sx], chunk * kBlockK, tj - 1, wo0, b0);
          }
          for (int e = 0; e < 3 * kParts; ++e, ++gw) {      // (ti, part): W_hi then W_lo of each tap
            const int sw = gw % NW;
            mbar_wait(&w_empty[sw], ((gw / NW) & 1) ^ 1);
            if (lane == 0) {
              const int ti = e / kParts, part = e - ti * kParts;
              const int tap = ti * 3 + tj;
              mbar_arrive_expect_tx(&w_full[sw], kWPart);
              tma_load_2d(smem_u32(w_ring + sw * kWPart), &tm.b, &w_full[sw], chunk * kBlockK,
                          (part * 9 + tap) * p.Cout + n0);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (HALO && warp == 1) {
    // ===================== MMA issuer (pixel windows) ===========================================
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_f16(kCoutTile, kPix);
      int gx = 0, gw = 0, k = 0;
      for (int t = blockIdx.x; t < total_units; t += gridDim.x, ++k) {
        if (t + static_cast<int>(gridDim.x) >= total_units) pdl_trigger_conv_late();
        const int acc = k & 1;
        mbar_wait(&tmem_empty[acc], ((k >> 1) & 1) ^ 1);       // epilogue has drained this accumulator set
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kPix;
        for (int u = 0; u < p.units; ++u, ++gx) {
          const int sx = gx & 1;
          mbar_wait(&x_full[sx], (gx >> 1) & 1);
          if (dbg && gx == 0) p.dbg[2] = clock64();
          const uint32_t x_base = smem_u32(smem + sx * x_stage);
          for (int ti = 0; ti < 3; ++ti) {
            // tap ti of this kernel column = the window shifted by ti columns (ti*Ho rows): whole swizzle atoms
            const uint64_t x_desc = umma_desc_sw128(x_base + ti * p.Ho * 128);
            {
              const int sw = gw % NW;
              mbar_wait(&w_full[sw], (gw / NW) & 1);
              tc_fence_after();
              const uint64_t w_desc = umma_desc_sw128(smem_u32(w_ring + sw * kWPart));
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) umma_f16(d_tmem, w_desc + 2 * kk, x_desc + 2 * kk, idesc, (u | ti | kk) != 0);
              if (TERMS > 1) {
                const uint64_t xl_desc = umma_desc_sw128(x_base + ti * p.Ho * 128 + p.a_part_bytes);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) umma_f16(d_tmem, w_desc + 2 * kk, xl_desc + 2 * kk, idesc, 1u);   // W_hi X_lo
              }
              umma_commit(&w_empty[sw]);
              ++gw;
            }
            if (TERMS > 1) {
              const int sw = gw % NW;
              mbar_wait(&w_full[sw], (gw / NW) & 1);
              tc_fence_after();
              const uint64_t wl_desc = umma_desc_sw128(smem_u32(w_ring + sw * kWPart));
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) umma_f16(d_tmem, wl_desc + 2 * kk, x_desc + 2 * kk, idesc, 1u);     // W_lo X_hi
              umma_commit(&w_empty[sw]);
              ++gw;
            }
          }
          umma_commit(&x_empty[sx]);
        }
        umma_commit(&tmem_full[acc]);
        if (dbg && k == 0) p.dbg[3] = clock64();
      }
    }
    __syncwarp();
  } else if (warp == 0) {
    // ===================== TMA producer: one continuous stage ring across units =================
    int g = 0;
    for (int t = blockIdx.x; t < total_units; t += gridDim.x) {
      if (t + static_cast<int>(gridDim.x) >= total_units) pdl_trigger_conv_late();
      const int um = t / tiles_n, tn = t - um * tiles_n;
      const int m0 = um * kPix, n0 = tn * kCoutTile;
      int b0[2], wo0[2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int q0 = (m0 + mt * kBlockM) / p.Ho;          // global column index of the 128-pixel tile's first column
        b0[mt] = q0 / p.Wo;
        wo0[mt] = q0 - b0[mt] * p.Wo;
      }
      for (int it = 0; it < n_it; ++it, ++g) {
        const int s = g % STAGES;
        mbar_wait(&empty_bar[s], ((g / STAGES) & 1) ^ 1);
        const bool main = it < p.main_iters;
        const int chunk = main ? it / taps : it - p.main_iters;
        const int tap = it - chunk * taps;
        const int ti = tap / p.ks, tj = tap - ti * p.ks;
        const int kc = chunk * kBlockK;
        const uint32_t dst = smem_u32(smem + s * kStageBytes);
        // shortcut K steps: centre tap of the second tensor (1x1, stride 1, same grid as the output)
        const int h_in = main ? tj - p.pad_lo : 0;
        const int w_off = main ? ti - p.pad_lo + 1 : 1;
        const int w_mul = main ? p.stride : 1;
        if (lane == 0) {
          mbar_arrive_expect_tx(&full_bar[s], kStageBytes);
          if (main) {
            tma_load_2d(dst + kWOff, &tm.b, &full_bar[s], kc, tap * p.Cout + n0);
            if (TERMS > 1) tma_load_2d(dst + kWOff + kWPart, &tm.b, &full_bar[s], kc, (taps + tap) * p.Cout + n0);
          } else {
            tma_load_2d(dst + kWOff, &tm.b2, &full_bar[s], kc, n0);
            if (TERMS > 1) tma_load_2d(dst + kWOff + kWPart, &tm.b2, &full_bar[s], kc, p.Cout + n0);
          }
        }
        // lanes 1.. : one 128-pixel box per (tile, operand part); the two tiles of a part are adjacent = 256 rows
        if (lane >= 1 && lane <= 2 * kParts) {
          const int mt = (lane - 1) & 1, part = (lane - 1) >> 1;
          const CUtensorMap* map = main ? (part ? &tm.alo : &tm.a) : (part ? &tm.a2lo : &tm.a2);
          tma_load_4d(dst + part * kXPart + mt * kABytes, map, &full_bar[s], kc, h_in, w_mul * wo0[mt] + w_off, b0[mt]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: alternates between the two TMEM accumulator sets ===========
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_f16(kCoutTile, kPix);
      int g = 0, k = 0;
      for (int t = blockIdx.x; t < total_units; t += gridDim.x, ++k) {
        if (t + static_cast<int>(gridDim.x) >= total_units) pdl_trigger_conv_late();
        const int acc = k & 1;
        mbar_wait(&tmem_empty[acc], ((k >> 1) & 1) ^ 1);       // epilogue has drained this accumulator set
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kPix;
        for (int it = 0; it < n_it; ++it, ++g) {
          const int s = g % STAGES;
          mbar_wait(&full_bar[s], (g / STAGES) & 1);
          if (dbg && g == 0) p.dbg[2] = clock64();
          tc_fence_after();
          const uint32_t base = smem_u32(smem + s * kStageBytes);
          const uint64_t x_desc = umma_desc_sw128(base), w_desc = umma_desc_sw128(base + kWOff);
#pragma unroll
          for (int kk = 0; kk < kBlockK / 16; ++kk)
            umma_f16(d_tmem, w_desc + 2 * kk, x_desc + 2 * kk, idesc, (it | kk) != 0);
          if (TERMS > 1) {
            const uint64_t xl_desc = umma_desc_sw128(base + kXPart), wl_desc = umma_desc_sw128(base + kWOff + kWPart);
#pragma unroll
            for (int kk = 0; kk < kBlockK / 16; ++kk) {
              umma_f16(d_tmem, w_desc + 2 * kk, xl_desc + 2 * kk, idesc, 1u);     // W_hi X_lo
              umma_f16(d_tmem, wl_desc + 2 * kk, x_desc + 2 * kk, idesc, 1u);     // W_lo X_hi
            }
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tmem_full[acc]);
        if (dbg && k == 0) p.dbg[3] = clock64();
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps: thread = output channel, drain unit k under unit k+1 ===
    const int q = warp & 3;                   // TMEM lane quadrant readable by this warp = channels 32q..32q+31
    constexpr int kChunksPerWarp = (kPix / 32) / (kWtEpiWarps / 4);
    const int ch0 = ((warp - 2) >> 2) * kChunksPerWarp;       // this warp's first 32-pixel chunk of every unit
    int k = 0;
    for (int t = blockIdx.x; t < total_units; t += gridDim.x, ++k) {
      const int acc = k & 1;
      const int um = t / tiles_n, tn = t - um * tiles_n;
      const int m0 = um * kPix;
      const int c = tn * kCoutTile + q * 32 + lane;             // this thread's output channel
      const int bimg = m0 / p.pix_per_img;                      // a unit lies inside one image (host)
      float add = p.bias ? __ldg(p.bias + c) : 0.f;
      if (p.temb) add += __ldg(p.temb + static_cast<size_t>(bimg) * p.temb_stride + c);
      float* outp = p.out + static_cast<size_t>(m0) * p.Cout + c;
      const float* resp = p.residual ? p.residual + static_cast<size_t>(m0) * p.Cout + c : nullptr;
      float rs[32];
      auto fetch_res = [&](int ch) {                            // 32 coalesced 128 B rows
#pragma unroll
        for (int j = 0; j < 32; ++j) rs[j] = __ldg(resp + static_cast<size_t>(ch * 32 + j) * p.Cout);
      };
      if (resp) fetch_res(ch0);
      mbar_wait(&tmem_full[acc], (k >> 1) & 1);
      const bool dbg_e = dbg && k == 0 && threadIdx.x == 64;
      if (dbg_e) p.dbg[4] = clock64();
      if (t + static_cast<int>(gridDim.x) >= total_units) pdl_trigger_conv_late();
      tc_fence_after();
      float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
      for (int ch = ch0; ch < ch0 + kChunksPerWarp; ++ch) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + acc * kPix + (static_cast<uint32_t>(q * 32) << 16) + ch * 32, r);
        tmem_ld_wait();
        if (dbg_e && ch == ch0) p.dbg[6] = p.dbg[7] = clock64();
        if (ch == ch0 + kChunksPerWarp - 1) {  // this warp's last TMEM read of the unit: hand the accumulators back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty[acc])) : "memory");
        }
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + add + (resp ? rs[j] : 0.f);
        if (resp && ch + 1 < ch0 + kChunksPerWarp) fetch_res(ch + 1);   // next chunk's residual rows fly under these stores
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          outp[static_cast<size_t>(ch * 32 + j) * p.Cout] = v[j];
          s1 += v[j];
          s2 = fmaf(v[j], v[j], s2);
        }
      }
      if (dbg_e) p.dbg[8] = clock64();
      if (p.stats) {                          // channel-pair moments of the finished output: (c, c+1) -> one slot
        s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
        s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
        if ((lane & 1) == 0) {
          double* st = p.stats + (static_cast<size_t>(bimg) * p.stats_G + c / 2) * 2;
          atomicAdd(st, static_cast<double>(s1));
          atomicAdd(st + 1, static_cast<double>(s2));
        }
      }
      if (dbg_e) { __threadfence(); p.dbg[5] = clock64(); }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<2 * kPix>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// Halo-reuse kernel for the 3x3 / stride-1 convolutions that carry ~95 % of the FLOPs.
//
// In the per-tap kernel above every one of the 9 taps re-reads its A tile from L2 and every 128-pixel tile
// re-reads all weights: 64 KB per K step, which makes the kernel operand-ingest bound.  Here
//   * one TMA box per (64-channel chunk, kernel column tj) brings MT*128 output pixels PLUS one halo column on
//     each side: (MT*ncols + 2) columns x Ho rows.  The three taps ti = 0,1,2 of that kernel column are the same
//     shared-memory tile shifted by ti*Ho rows (a multiple of the 1024 B swizzle atom for Ho >= 8), i.e. just
//     three UMMA descriptors -- A traffic drops 3x;
//   * MT = 2 accumulators (2 x 128 pixels, TMEM columns [0,BN) and [BN,2BN)) share every weight tile -- B traffic
//     per output halves;
//   * A stages (2) and weight stages (p.nb_stages) live in separate mbarrier rings.
template <int BLOCK_N, int MT, int TERMS>
__global__ void __launch_bounds__(192, 1)
conv3x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAlo,
               const __grid_constant__ CUtensorMap tmB, const ConvParams p) {
  constexpr int kParts = TERMS == 1 ? 1 : 2;
  constexpr int kBBytes = BLOCK_N * kBlockK * 2;      // one part of one weight tile
  constexpr int kBStage = kParts * kBBytes;
  constexpr int kMaxNB = 4;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int a_part = p.a_part_bytes;
  const int a_stage = kParts * a_part;
  const int NB = p.nb_stages;
  uint8_t* a_ring = smem;
  uint8_t* b_ring = smem + 2 * a_stage;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(b_ring + NB * kBStage);
  uint64_t* a_empty = a_full + 2;
  uint64_t* b_full = a_empty + 2;
  uint64_t* b_empty = b_full + kMaxNB;
  uint64_t* tmem_full_bar = b_empty + kMaxNB;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * (MT * kBlockM);
  const int n0 = blockIdx.y * BLOCK_N;
  const int u0 = static_cast<int>(static_cast<long long>(blockIdx.z) * p.units / gridDim.z);
  const int u1 = static_cast<int>(static_cast<long long>(blockIdx.z + 1) * p.units / gridDim.z);
  const int n_units = u1 - u0;                   // >= 1 (host keeps gridDim.z <= units)

  pdl_trigger();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    if (TERMS > 1) tma_prefetch_desc(&tmAlo);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < kMaxNB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    mbar_init(tmem_full_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<MT * BLOCK_N>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);   // shfl: tells the compiler it is warp-uniform (UR, no per-MMA R2UR loop)
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer =========================================================
    const int q0 = m0 / p.Ho;                     // first output column of the tile (global column index)
    const int b0 = q0 / p.Wo;
    const int wo0 = q0 - b0 * p.Wo;               // == padded column of tap ti = 0
    int bi = 0;
    for (int ui = 0; ui < n_units; ++ui) {
      const int u = u0 + ui;
      const int chunk = u / 3, tj = u - chunk * 3;
      const int sa = ui & 1;
      mbar_wait(&a_empty[sa], ((ui >> 1) & 1) ^ 1);
      const uint32_t a_dst = smem_u32(a_ring + sa * a_stage);
      if (lane == 0) {
        mbar_arrive_expect_tx(&a_full[sa], a_stage);
        tma_load_4d(a_dst, &tmA, &a_full[sa], chunk * kBlockK, tj - 1, wo0, b0);
      }
      if (TERMS > 1 && lane == 1) tma_load_4d(a_dst + a_part, &tmAlo, &a_full[sa], chunk * kBlockK, tj - 1, wo0, b0);
      for (int ti = 0; ti < 3; ++ti, ++bi) {
        const int sb = bi % NB;
        mbar_wait(&b_empty[sb], ((bi / NB) & 1) ^ 1);
        if (lane == 0) {
          const int tap = ti * 3 + tj;
          const uint32_t b_dst = smem_u32(b_ring + sb * kBStage);
          mbar_arrive_expect_tx(&b_full[sb], kBStage);
          tma_load_2d(b_dst, &tmB, &b_full[sb], chunk * kBlockK, tap * p.Cout + n0);
          if (TERMS > 1) tma_load_2d(b_dst + kBBytes, &tmB, &b_full[sb], chunk * kBlockK, (9 + tap) * p.Cout + n0);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer ===========================================================
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_f16(kBlockM, BLOCK_N);
      int bi = 0;
      for (int ui = 0; ui < n_units; ++ui) {
        const int sa = ui & 1;
        mbar_wait(&a_full[sa], (ui >> 1) & 1);
        const uint32_t a_base = smem_u32(a_ring + sa * a_stage);
        for (int ti = 0; ti < 3; ++ti, ++bi) {
          const int sb = bi % NB;
          mbar_wait(&b_full[sb], (bi / NB) & 1);
          tc_fence_after();
          const uint32_t b_addr = smem_u32(b_ring + sb * kBStage);
          const uint64_t b_desc = umma_desc_sw128(b_addr);
          const uint64_t bl_desc = umma_desc_sw128(b_addr + kBBytes);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            // tap ti of this kernel column = the staged tile shifted by ti columns (ti*Ho rows); sub-tile mt
            // starts mt*128 rows further down.  Both offsets are whole 1024 B swizzle atoms.
            const uint32_t a_addr = a_base + (ti * p.Ho + mt * kBlockM) * 128;
            const uint64_t a_desc = umma_desc_sw128(a_addr);
            const uint64_t al_desc = umma_desc_sw128(a_addr + a_part);
            const uint32_t acc = tmem_base + mt * BLOCK_N;
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)
              umma_f16(acc, a_desc + 2 * k, b_desc + 2 * k, idesc, (ui | ti | k) != 0);
            if (TERMS > 1) {
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k) {
                umma_f16(acc, al_desc + 2 * k, b_desc + 2 * k, idesc, 1u);   // A_lo * W_hi
                umma_f16(acc, a_desc + 2 * k, bl_desc + 2 * k, idesc, 1u);   // A_hi * W_lo
              }
            }
          }
          umma_commit(&b_empty[sb]);
        }
        umma_commit(&a_empty[sa]);
      }
      umma_commit(tmem_full_bar);
    }
    __syncwarp();
  } else {
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
  }
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
    epilogue_tile<BLOCK_N>(smem, tmem_base + mt * BLOCK_N, m0 + mt * kBlockM, n0, p, warp, lane);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<MT * BLOCK_N>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

template <int BLOCK_N, int STAGES, int TERMS, int NSPLIT>
static int launch_conv_n(const ConvMaps& tm, const ConvParams& p, cudaStream_t st) {
  // pipeline stages (+ alignment slack) + barriers/TMEM pointer + the per-warp GroupNorm-moment scratch
  constexpr int smem = STAGES * (TERMS == 1 ? 1 : 2) * (kABytes + BLOCK_N * kBlockK * 2) + 1024 + 256 +
                       2 * 4 * (BLOCK_N / 2) * 4 + 64;
  static bool attr_set = false;
  if (!attr_set) {
    RLDM_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BLOCK_N, STAGES, TERMS, NSPLIT>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  dim3 grid((p.M_total + kBlockM - 1) / kBlockM, p.Cout / BLOCK_N, NSPLIT);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl_enabled_small()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (NSPLIT > 1) {   // the K splits of one tile form a thread-block cluster (DSMEM reduction in the epilogue)
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 1;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = NSPLIT;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  RLDM_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<BLOCK_N, STAGES, TERMS, NSPLIT>, tm, p));
  return 0;
}

template <int BLOCK_N, int STAGES, int TERMS>
static int launch_conv(const ConvMaps& tm, const ConvParams& p, int split, cudaStream_t st) {
  switch (split) {
    case 1: return launch_conv_n<BLOCK_N, STAGES, TERMS, 1>(tm, p, st);
    case 2: return launch_conv_n<BLOCK_N, STAGES, TERMS, 2>(tm, p, st);
    case 4: return launch_conv_n<BLOCK_N, STAGES, TERMS, 4>(tm, p, st);
    default: return launch_conv_n<BLOCK_N, STAGES, TERMS, 8>(tm, p, st);
  }
}

template <int BLOCK_N, int STAGES, int TERMS, int MT, int KB>
static int launch_conv_persistent(const ConvMaps& tm, const ConvParams& p, int n_ctas, cudaStream_t st) {
  constexpr int smem = STAGES * (TERMS == 1 ? 1 : 2) * (MT * kBlockM + BLOCK_N) * KB * 2 + kBlockM * 36 * 4 + 256 +
                       2 * 4 * (BLOCK_N / 2) * 4 + 64 + 1024;
  static_assert(smem <= 232448, "persistent conv: shared memory budget exceeded");
  static bool attr_set = false;
  if (!attr_set) {
    RLDM_CUDA(cudaFuncSetAttribute(conv_tc_persistent_kernel<BLOCK_N, STAGES, TERMS, MT, KB>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  RLDM_CUDA(launch_pdl(conv_tc_persistent_kernel<BLOCK_N, STAGES, TERMS, MT, KB>, dim3(n_ctas), dim3(192), smem, st, tm, p));
  return 0;
}

template <int TERMS>
static int launch_conv_wt(const ConvMaps& tm, const ConvParams& p, int n_ctas, cudaStream_t st) {
  constexpr int smem = 2 * (TERMS == 1 ? 1 : 2) * (256 + 128) * 64 * 2 + 256 + 1024;
  static_assert(smem <= 232448, "conv_tc_wt: shared memory budget exceeded");
  static bool attr_set = false;
  if (!attr_set) {
    RLDM_CUDA(cudaFuncSetAttribute(conv_tc_wt_kernel<TERMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  RLDM_CUDA(launch_pdl(conv_tc_wt_kernel<TERMS>, dim3(n_ctas), dim3(kWtThreads), smem, st, tm, p));
  return 0;
}

template <int TERMS>
static int launch_conv_wt_halo(const ConvMaps& tm, const ConvParams& p, int n_ctas, size_t smem, cudaStream_t st) {
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    RLDM_CUDA(cudaFuncSetAttribute(conv_tc_wt_kernel<TERMS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(smem)));
    attr_smem = smem;
  }
  // single-wave launches (one unit per CTA: the top-level UNet layers) may start under the tail of the producing
  // pass: setup, TMEM allocation and descriptor prefetch overlap it.  UNet forward 1951-1953 -> 1944-1945 us
  // (RLDM_WT_PDL=0 switches it off).
  static int wt_pdl = -1;
  if (wt_pdl < 0) { const char* e = getenv("RLDM_WT_PDL"); wt_pdl = e ? atoi(e) : 1; }
  const int units = (p.M_total / 256) * (p.Cout / 128);
  if (wt_pdl && units <= n_ctas)
    RLDM_CUDA(launch_pdl_small(conv_tc_wt_kernel<TERMS, true>, dim3(n_ctas), dim3(kWtThreads), smem, st, tm, p));
  else
    RLDM_CUDA(launch_pdl(conv_tc_wt_kernel<TERMS, true>, dim3(n_ctas), dim3(kWtThreads), smem, st, tm, p));
  return 0;
}

// Persistent halo-window variant: shared memory = 2 A windows + nb weight entries + slab + barriers (sized by the caller)
template <int BLOCK_N, int TERMS>
static int launch_conv_persistent_halo(const ConvMaps& tm, const ConvParams& p, int n_ctas, size_t smem, cudaStream_t st) {
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    RLDM_CUDA(cudaFuncSetAttribute(conv_tc_persistent_kernel<BLOCK_N, 2, TERMS, 2, 64, true>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr_smem = smem;
  }
  RLDM_CUDA(launch_pdl(conv_tc_persistent_kernel<BLOCK_N, 2, TERMS, 2, 64, true>, dim3(n_ctas), dim3(192), smem, st, tm, p));
  return 0;
}

template <int BLOCK_N, int MT, int TERMS>
static int launch_conv3x3(const CUtensorMap& tmA, const CUtensorMap& tmAlo, const CUtensorMap& tmB,
                          const ConvParams& p, int split, size_t smem, cudaStream_t st) {
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    RLDM_CUDA(cudaFuncSetAttribute(conv3x3_kernel<BLOCK_N, MT, TERMS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(smem)));
    attr_smem = smem;
  }
  dim3 grid((p.M_total + MT * kBlockM - 1) / (MT * kBlockM), p.Cout / BLOCK_N, split);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (split > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 1;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = split;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  RLDM_CUDA(cudaLaunchKernelEx(&cfg, conv3x3_kernel<BLOCK_N, MT, TERMS>, tmA, tmAlo, tmB, p));
  return 0;
}

}  // namespace rldm

using namespace rldm;

// profiling aid (not part of include/rldm.h): device buffer of 16 int64 that CTA 0 of the per-tap kernel fills with
// clock64() stamps: [0] entry, [1] prologue done, [2] first stage landed, [3] last MMA issued, [4] accumulator
// complete, [5] epilogue done.
extern "C" void rldm_debug_conv_timestamps(long long* dev_buf) { g_conv_dbg = dev_buf; }

static int conv_tc_impl(const uint16_t* x, const uint16_t* x_lo, const uint16_t* wgt, const float* bias,
                        const float* temb, int temb_stride, const float* residual, float* out,
                        int B, int W, int H, int Cin, int Cout, int ks, int stride, int pad_lo,
                        int circular, int split_k, double* stats, const uint16_t* sc_x, const uint16_t* sc_x_lo,
                        const uint16_t* sc_wgt, int sc_cin, float* splitk_ws, size_t splitk_ws_bytes, void* stream) {
  RLDM_CHECK(ks == 1 || ks == 3, "conv_tc: ks must be 1 or 3 (got %d)", ks);
  RLDM_CHECK(!sc_x || (sc_wgt && sc_cin > 0 && sc_cin % 64 == 0 && stride == 1 && (!x_lo == !sc_x_lo)),
             "conv_tc: fused shortcut needs weights, Cin2 %% 64 == 0 (got %d), stride 1 and the same operand precision",
             sc_cin);
  RLDM_CHECK(stride == 1 || stride == 2, "conv_tc: stride must be 1 or 2 (got %d)", stride);
  RLDM_CHECK(Cin % 64 == 0, "conv_tc: Cin %% 64 != 0 (got %d)", Cin);
  RLDM_CHECK(Cout % 64 == 0, "conv_tc: Cout %% 64 != 0 (got %d)", Cout);
  RLDM_CHECK(W % stride == 0 && H % stride == 0, "conv_tc: W,H must divide by stride");
  const int Wo = W / stride, Ho = H / stride;
  RLDM_CHECK(Ho >= 1 && Ho <= 128 && (Ho & (Ho - 1)) == 0, "conv_tc: Ho must be a power of two <= 128 (got %d)", Ho);
  RLDM_CHECK((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(x_lo) & 15) == 0 &&
             (reinterpret_cast<uintptr_t>(wgt) & 15) == 0 &&
             (reinterpret_cast<uintptr_t>(out) & 15) == 0, "conv_tc: pointers must be 16 B aligned");
  EncodeTiledFn encode = get_encode();
  RLDM_CHECK(encode != nullptr, "conv_tc: cuTensorMapEncodeTiled unavailable (no CUDA driver?)");

  const int BN = (Cout % 128 == 0) ? 128 : 64;
  // M tile = 128 output pixels = ncols whole columns x nb images
  const int pix = Wo * Ho;
  RLDM_CHECK(pix % 128 == 0 || 128 % pix == 0, "conv_tc: Wo*Ho=%d must divide or be a multiple of 128", pix);
  const int nb = pix >= 128 ? 1 : 128 / pix;
  const int ncols = pix >= 128 ? 128 / Ho : Wo;
  RLDM_CHECK(ncols * stride <= 256, "conv_tc: tile of %d columns exceeds the TMA box limit", ncols);
  const int parts = x_lo ? 2 : 1;
  CUtensorMap tmA, tmAlo, tmB;
  // ---- persistent halo-window path (opt-in, RLDM_HALO_P=1 | nores): 3x3, stride 1, symmetric pad, split-fp16, more
  //      128x128 tiles than SMs, two whole M tiles per unit, room for two A windows plus >= 3 weight entries.
  //      Measured on B200 (C3, batch 8): bit-for-bit the same contract, L2->SM bytes -42 %, but 8-10 % SLOWER than the
  //      per-tap persistent kernel (UNet 128->128 @256x16: 30.7 -> 33.6 us; decoder 256->256 @256x16: 90 -> 99 us).
  //      The K loop is paced by shared-memory bandwidth (a 128x128x16 SS MMA reads 8 KB in 64 clk = the whole
  //      128 B/clk, TMA fills compete for the rest) and the 3 x 16 KB weight ring that fits next to two 72 KB windows
  //      is too shallow; fewer operand bytes from L2 do not help.  The fix is fewer shared-memory reads per MMA
  //      (cta_group::2 / N = 256), see DESIGN.md "Next". ----
  {
    static int n_sms_h = 0;
    if (n_sms_h == 0) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&n_sms_h, cudaDevAttrMultiProcessorCount, dev);
      if (n_sms_h <= 0) n_sms_h = 148;
    }
    const char* hp = getenv("RLDM_HALO_P");              // unset / "0": off, "nores": only layers without a residual
    const int tiles_h = (B * pix / kBlockM) * (Cout / BN);
    const size_t a_stage = static_cast<size_t>(2) * (2 * kBlockM + 2 * Ho) * 128;
    const size_t b_entry = static_cast<size_t>(BN) * 128;
    const size_t fixed = 2 * a_stage + kBlockM * 36 * 4 + (4 + 2 * 8 + 4) * 8 + 16 + 2 * 4 * (BN / 2) * 4 + 64 + 1024;
    int nbs = fixed + 2 * b_entry <= 232448 ? static_cast<int>((232448 - fixed) / b_entry) : 0;
    if (nbs > 8) nbs = 8;
    static int min_nb = 0;
    if (min_nb == 0) { const char* e = getenv("RLDM_HALO_P_MINNB"); min_nb = e ? atoi(e) : 3; }
    const bool halo_geom = ks == 3 && stride == 1 && pad_lo == 1 && Ho >= 8 && pix % (2 * kBlockM) == 0 && !sc_x && parts == 2 &&
                           split_k <= 1 && tiles_h > n_sms_h && !getenv("RLDM_NO_PERSISTENT") && !getenv("RLDM_HALO");
    const bool halo_p = halo_geom && nbs >= min_nb && hp && hp[0] != '0' && !(hp[0] == 'n' && residual);
    // role-swapped kernel with pixel windows (default where it applies): no staging slab, so the weight ring is 4-5 deep
    const char* wh = getenv("RLDM_CONV_WT_HALO");         // "0": off, "nores": only layers without a residual operand
    const char* wt_env_h = getenv("RLDM_CONV_WT");
    const size_t fixed_wt = 2 * a_stage + (4 + 2 * 8 + 4) * 8 + 16 + 1024;
    int nws = fixed_wt + 2 * 16384 <= 232448 ? static_cast<int>((232448 - fixed_wt) / 16384) : 0;
    if (nws > 8) nws = 8;
    const bool halo_wt = halo_geom && !halo_p && BN == 128 && nws >= min_nb && !(wh && wh[0] == '0') &&
                         !(wh && wh[0] == 'n' && residual) && !(wt_env_h && wt_env_h[0] == '0');
    if (halo_p || halo_wt) {
      ConvMaps tmh;
      const int cols = 2 * (kBlockM / Ho);
      for (int part = 0; part < 2; ++part) {
        cuuint64_t gdim[4] = {(cuuint64_t)Cin, (cuuint64_t)H, (cuuint64_t)(W + 2), (cuuint64_t)B};
        cuuint64_t gstr[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)H * Cin * 2, (cuuint64_t)(W + 2) * H * Cin * 2};
        cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)Ho, (cuuint32_t)(cols + 2), 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(part ? &tmh.alo : &tmh.a, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4,
                            const_cast<uint16_t*>(part ? x_lo : x), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        RLDM_CHECK(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled(A window) failed: %d", (int)r);
      }
      {
        cuuint64_t gdim[2] = {(cuuint64_t)Cin, (cuuint64_t)2 * 9 * Cout};
        cuuint64_t gstr[1] = {(cuuint64_t)Cin * 2};
        cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)BN};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&tmh.b, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<uint16_t*>(wgt), gdim, gstr,
                            box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        RLDM_CHECK(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled(B) failed: %d", (int)r);
      }
      tmh.a2 = tmh.a; tmh.a2lo = tmh.alo; tmh.b2 = tmh.b;
      ConvParams p;
      p.bias = bias; p.temb = temb; p.residual = residual; p.out = out;
      p.temb_stride = temb_stride;
      p.M_total = B * pix;
      p.Wo = Wo; p.Ho = Ho; p.W_in = W;
      p.pix_per_img = pix;
      p.Cout = Cout;
      p.ks = 3; p.stride = 1; p.pad_lo = 1; p.circular = circular;
      p.total_iters = p.main_iters = (Cin / kBlockK) * 9;
      p.units = (Cin / kBlockK) * 3;
      p.a_part_bytes = static_cast<int>(a_stage / 2);
      p.nb_stages = nbs;
      p.dbg = g_conv_dbg;
      p.ws = nullptr;
      p.stats = stats;
      p.stats_G = Cout / 2;
      p.stats_cpg = 2;
      const int units = tiles_h / 2;
      const int ctas = units < n_sms_h ? units : n_sms_h;
      cudaStream_t st = as_stream(stream);
      if (halo_wt) {
        p.nb_stages = nws;
        return launch_conv_wt_halo<3>(tmh, p, ctas, fixed_wt + static_cast<size_t>(nws) * 16384, st);
      }
      const size_t smem = fixed + nbs * b_entry;
      if (BN == 128) return launch_conv_persistent_halo<128, 3>(tmh, p, ctas, smem, st);
      return launch_conv_persistent_halo<64, 3>(tmh, p, ctas, smem, st);
    }
  }
  // ---- halo-reuse path: 3x3, stride 1, symmetric pad, column pitch a whole number of swizzle atoms ----
  // (measured on B200: correct but not faster than the per-tap kernel, whose limiter is per-CTA latency rather
  //  than operand bytes -- kept opt-in with RLDM_HALO=1 until it is made persistent)
  const bool halo_ok = ks == 3 && stride == 1 && pad_lo == 1 && Ho >= 8 && pix >= 128 && !sc_x && getenv("RLDM_HALO");
  if (halo_ok) {
    const size_t limit = 232448 - 1024 - 3328;          // 227 KB minus alignment slack and static shared memory
    const size_t b_stage = static_cast<size_t>(parts) * BN * 128;
    const size_t bars = 256;
    auto a_stage_of = [&](int mt) { return static_cast<size_t>(parts) * (mt * 128 + 2 * Ho) * 128; };
    int MT = (pix % 256 == 0 && 2 * a_stage_of(2) + 2 * b_stage + bars <= limit) ? 2 : 1;
    if (getenv("RLDM_HALO_MT1")) MT = 1;
    const size_t a_stage = a_stage_of(MT);
    if (2 * a_stage + 2 * b_stage + bars <= limit) {
      int nbs = static_cast<int>((limit - bars - 2 * a_stage) / b_stage);
      if (nbs > 4) nbs = 4;
      size_t smem = 2 * a_stage + nbs * b_stage + bars;
      const size_t stage_tile = static_cast<size_t>(128) * (BN + 4) * 4;
      if (smem < stage_tile + bars) smem = stage_tile + bars;
      smem += 1024;
      const int cols = MT * (128 / Ho);
      for (int part = 0; part < parts; ++part) {
        cuuint64_t gdim[4] = {(cuuint64_t)Cin, (cuuint64_t)H, (cuuint64_t)(W + 2), (cuuint64_t)B};
        cuuint64_t gstr[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)H * Cin * 2, (cuuint64_t)(W + 2) * H * Cin * 2};
        cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)Ho, (cuuint32_t)(cols + 2), 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(part ? &tmAlo : &tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4,
                            const_cast<uint16_t*>(part ? x_lo : x), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        RLDM_CHECK(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled(A halo) failed: %d", (int)r);
      }
      if (parts == 1) tmAlo = tmA;
      {
        cuuint64_t gdim[2] = {(cuuint64_t)Cin, (cuuint64_t)parts * 9 * Cout};
        cuuint64_t gstr[1] = {(cuuint64_t)Cin * 2};
        cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)BN};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<uint16_t*>(wgt), gdim, gstr,
                            box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        RLDM_CHECK(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled(B) failed: %d", (int)r);
      }
      ConvParams p;
      p.bias = bias; p.temb = temb; p.residual = residual; p.out = out;
      p.temb_stride = temb_stride;
      p.M_total = B * pix;
      p.Wo = Wo; p.Ho = Ho; p.W_in = W;
      p.pix_per_img = pix;
      p.Cout = Cout;
      p.ks = 3; p.stride = 1; p.pad_lo = 1; p.circular = circular;
      p.total_iters = p.main_iters = (Cin / kBlockK) * 9;
      p.units = (Cin / kBlockK) * 3;
      p.a_part_bytes = static_cast<int>(a_stage / parts);
      p.nb_stages = nbs;
      p.dbg = nullptr;
      p.ws = nullptr;
      p.stats = stats;
      p.stats_G = Cout / 2;
      p.stats_cpg = 2;
      const int tiles = ((p.M_total + MT * 128 - 1) / (MT * 128)) * (Cout / BN);
      int split = split_k;
      if (split <= 0) {
        split = 1;
        while (tiles * split * 2 <= 160 && p.units / (split * 2) >= 2 && split < 8) split *= 2;
      }
      RLDM_CHECK(split == 1 || split == 2 || split == 4 || split == 8, "conv_tc: split_k must be 1, 2, 4 or 8 (got %d)", split);
      while (split > p.units) split /= 2;
      cudaStream_t st = as_stream(stream);
      if (parts == 2) {
        if (BN == 128) return MT == 2 ? launch_conv3x3<128, 2, 3>(tmA, tmAlo, tmB, p, split, smem, st)
                                      : launch_conv3x3<128, 1, 3>(tmA, tmAlo, tmB, p, split, smem, st);
        return MT == 2 ? launch_conv3x3<64, 2, 3>(tmA, tmAlo, tmB, p, split, smem, st)
                       : launch_conv3x3<64, 1, 3>(tmA, tmAlo, tmB, p, split, smem, st);
      }
      if (BN == 128) return MT == 2 ? launch_conv3x3<128, 2, 1>(tmA, tmAlo, tmB, p, split, smem, st)
                                    : launch_conv3x3<128, 1, 1>(tmA, tmAlo, tmB, p, split, smem, st);
      return MT == 2 ? launch_conv3x3<64, 2, 1>(tmA, tmAlo, tmB, p, split, smem, st)
                     : launch_conv3x3<64, 1, 1>(tmA, tmAlo, tmB, p, split, smem, st);
    }
  }
  ConvMaps tm;
  // activation maps: (C, H, W+2, B) fp16, box = (64 channels, Ho*stride rows, ncols*stride columns, nb images)
  int KB = kBlockK;     // channels per pipeline stage: 64 (SWIZZLE_128B rows) or 32 (SWIZZLE_64B), chosen below
  auto encode_act = [&](CUtensorMap* m, const uint16_t* ptr, int C, int st) -> CUresult {
    cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)H, (cuuint64_t)(W + 2), (cuuint64_t)B};
    cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)H * C * 2, (cuuint64_t)(W + 2) * H * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)KB, (cuuint32_t)(Ho * st), (cuuint32_t)(ncols * st), (cuuint32_t)nb};
    cuuint32_t estr[4] = {1, (cuuint32_t)st, (cuuint32_t)st, 1};
    return encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<uint16_t*>(ptr), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, KB == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  // weight maps: [planes*taps*Cout][C] fp16, box = (64 channels, BN rows)
  auto encode_wgt = [&](CUtensorMap* m, const uint16_t* ptr, int C, int rows) -> CUresult {
    cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {(cuuint32_t)KB, (cuuint32_t)BN};
    cuuint32_t estr[2] = {1, 1};
    return encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<uint16_t*>(ptr), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, KB == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  auto build_maps = [&]() -> int {
    CUresult r = encode_act(&tm.a, x, Cin, stride);
    RLDM_CHECK(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled(A) failed: %d", (int)r);
    if (parts == 2) {
      r = encode_act(&tm.alo, x_lo, Cin, stride);
      RLDM_CHECK(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled(A lo) failed: %d", (int)r);
    } else {
      tm.alo = tm.a;
    }
    r = encode_wgt(&tm.b, wgt, Cin, parts * ks * ks * Cout);
    RLDM_CHECK(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled(B) failed: %d", (int)r);
    tm.a2 = tm.a; tm.a2lo = tm.alo; tm.b2 = tm.b;
    if (sc_x) {
      r = encode_act(&tm.a2, sc_x, sc_cin, 1);
      RLDM_CHECK(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled(shortcut A) failed: %d", (int)r);
      tm.a2lo = tm.a2;
      if (parts == 2) {
        r = encode_act(&tm.a2lo, sc_x_lo, sc_cin, 1);
        RLDM_CHECK(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled(shortcut A lo) failed: %d", (int)r);
      }
      r = encode_wgt(&tm.b2, sc_wgt, sc_cin, parts * Cout);
      RLDM_CHECK(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled(shortcut B) failed: %d", (int)r);
    }
    return 0;
  };
  ConvParams p;
  p.bias = bias; p.temb = temb; p.residual = residual; p.out = out;
  p.temb_stride = temb_stride;
  p.M_total = B * Wo * Ho;
  p.Wo = Wo; p.Ho = Ho; p.W_in = W;
  p.pix_per_img = Wo * Ho;
  p.Cout = Cout;
  p.ks = ks; p.stride = stride; p.pad_lo = pad_lo; p.circular = circular;
  p.main_iters = (Cin / kBlockK) * ks * ks;
  p.total_iters = p.main_iters + (sc_x ? sc_cin / kBlockK : 0);
  p.units = 0; p.a_part_bytes = 0; p.nb_stages = 0;
  p.dbg = g_conv_dbg;
  p.ws = nullptr;
  p.stats = stats;
  p.stats_G = Cout / 2;       // channel pairs per image
  p.stats_cpg = 2;
  RLDM_CHECK(pix >= 64 || !stats, "conv_tc: fused statistics need >= 64 pixels per image");
  const int tiles = ((p.M_total + kBlockM - 1) / kBlockM) * (Cout / BN);
  int split = split_k;
  if (split <= 0) {  // auto: fill the 148 SMs when the tile grid is small (clusters of <= 8 CTAs along K)
    split = 1;
    const int cap = parts == 2 ? 160 : 296;   // resident CTAs: 1 per SM in split-fp16 mode, 2 otherwise
    while (tiles * split * 2 <= cap && p.total_iters / (split * 2) >= 4 && split < 8) split *= 2;
  }
  RLDM_CHECK(split == 1 || split == 2 || split == 4 || split == 8, "conv_tc: split_k must be 1, 2, 4 or 8 (got %d)", split);
  while (split > p.total_iters) split /= 2;
  if (split > 1 && splitk_ws != nullptr && (reinterpret_cast<uintptr_t>(splitk_ws) & 15) == 0 &&
      static_cast<size_t>(tiles) * split * kBlockM * BN * sizeof(float) <= splitk_ws_bytes && !getenv("RLDM_SPLITK_DSMEM"))
    p.ws = splitk_ws;        // partial tiles through L2 instead of DSMEM
  cudaStream_t st = as_stream(stream);
  // more tiles than SMs and no K split: persistent CTAs with a double-buffered TMEM accumulator
  static int n_sms = 0;
  if (n_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev);
    if (n_sms <= 0) n_sms = 148;
  }
  if (split == 1 && tiles > n_sms && parts == 2 && !getenv("RLDM_NO_PERSISTENT")) {
    // two M tiles per unit share the weight tiles when the tile count allows it (RLDM_CONV_MT1=1: one tile per unit)
    const int tiles_m = (p.M_total + kBlockM - 1) / kBlockM;
    // (measured: -5..-20 % on layers without a residual operand and on 64-channel layers; 128-wide layers WITH a
    //  residual are paced by the drain of two tiles, not the K loop: they keep one tile per unit and three stages)
    const bool mt2 = tiles_m % 2 == 0 && p.M_total % kBlockM == 0 && (residual == nullptr || BN == 64 || getenv("RLDM_CONV_MT2_RES")) &&
                     !getenv("RLDM_CONV_MT1");
    // Cout tiles of 128 and whole 256-pixel units inside one image: roles swapped (weights = M side, N = 256 pixels),
    // 25 % fewer shared-memory operand reads per FLOP.  RLDM_CONV_WT=0 switches it off, =nores keeps layers with a
    // residual operand on the kernels above.
    {
      const char* wt_env = getenv("RLDM_CONV_WT");
      const bool wt = BN == 128 && p.M_total % 256 == 0 && pix % 256 == 0 && !(wt_env && wt_env[0] == '0') &&
                      !(wt_env && wt_env[0] == 'n' && residual);
      if (wt) {
        if (int rc = build_maps()) return rc;
        const int units_wt = (p.M_total / 256) * (Cout / 128);
        return launch_conv_wt<3>(tm, p, units_wt < n_sms ? units_wt : n_sms, st);
      }
    }
    const int units = mt2 ? tiles / 2 : tiles;
    const int ctas = units < n_sms ? units : n_sms;
    if (mt2) {
      // RLDM_CONV_KB32=1: half-chunk stages (SWIZZLE_64B, 48 KB) in a 4-deep ring instead of two 96 KB stages.
      // Measured slower on B200 (decoder convs 3.24 -> 3.62 ms): 64 B TMA rows move the same bytes less efficiently
      // than 128 B rows, which costs more than the deeper ring gains.  Kept as an experiment switch.
      if (getenv("RLDM_CONV_KB32")) {
        KB = 32;
        if (int rc = build_maps()) return rc;
        if (BN == 128) return launch_conv_persistent<128, 4, 3, 2, 32>(tm, p, ctas, st);
        return launch_conv_persistent<64, 4, 3, 2, 32>(tm, p, ctas, st);
      }
      if (int rc = build_maps()) return rc;
      if (BN == 128) return launch_conv_persistent<128, 2, 3, 2, 64>(tm, p, ctas, st);
      return launch_conv_persistent<64, 2, 3, 2, 64>(tm, p, ctas, st);
    }
    if (int rc = build_maps()) return rc;
    if (BN == 128) return launch_conv_persistent<128, 3, 3, 1, 64>(tm, p, ctas, st);
    return launch_conv_persistent<64, 4, 3, 1, 64>(tm, p, ctas, st);
  }
  if (int rc = build_maps()) return rc;
  if (parts == 2) {
    if (BN == 128) return launch_conv<128, 3, 3>(tm, p, split, st);
    return launch_conv<64, 4, 3>(tm, p, split, st);
  }
  if (BN == 128) return launch_conv<128, 3, 1>(tm, p, split, st);
  return launch_conv<64, 4, 1>(tm, p, split, st);
}

extern "C" int rldm_conv_tc(const uint16_t* x, const uint16_t* x_lo, const uint16_t* wgt, const float* bias,
                            const float* temb, int temb_stride, const float* residual, float* out,
                            int B, int W, int H, int Cin, int Cout, int ks, int stride, int pad_lo,
                            int circular, int split_k, double* stats, void* stream) {
  return conv_tc_impl(x, x_lo, wgt, bias, temb, temb_stride, residual, out, B, W, H, Cin, Cout, ks, stride, pad_lo,
                      circular, split_k, stats, nullptr, nullptr, nullptr, 0, nullptr, 0, stream);
}

extern "C" int rldm_conv_tc_shortcut(const uint16_t* x, const uint16_t* x_lo, const uint16_t* wgt, const float* bias,
                                     const float* temb, int temb_stride, const float* residual, float* out,
                                     int B, int W, int H, int Cin, int Cout, int ks, int stride, int pad_lo,
                                     int circular, int split_k, double* stats, const uint16_t* sc_x,
                                     const uint16_t* sc_x_lo, const uint16_t* sc_wgt, int sc_cin, void* stream) {
  return conv_tc_impl(x, x_lo, wgt, bias, temb, temb_stride, residual, out, B, W, H, Cin, Cout, ks, stride, pad_lo,
                      circular, split_k, stats, sc_x, sc_x_lo, sc_wgt, sc_cin, nullptr, 0, stream);
}

extern "C" int rldm_conv_tc_ws(const uint16_t* x, const uint16_t* x_lo, const uint16_t* wgt, const float* bias,
                               const float* temb, int temb_stride, const float* residual, float* out,
                               int B, int W, int H, int Cin, int Cout, int ks, int stride, int pad_lo,
                               int circular, int split_k, double* stats, const uint16_t* sc_x,
                               const uint16_t* sc_x_lo, const uint16_t* sc_wgt, int sc_cin, float* splitk_ws,
                               long long splitk_ws_bytes, void* stream) {
  return conv_tc_impl(x, x_lo, wgt, bias, temb, temb_stride, residual, out, B, W, H, Cin, Cout, ks, stride, pad_lo,
                      circular, split_k, stats, sc_x, sc_x_lo, sc_wgt, sc_cin, splitk_ws,
                      splitk_ws_bytes > 0 ? static_cast<size_t>(splitk_ws_bytes) : 0, stream);
}
