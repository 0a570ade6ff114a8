// conv_tc.cu -- horizontally-circular implicit-GEMM convolution on tcgen05 tensor cores.
//
// Replaces `Conv2d._conv_forward` (reference `ldm/utils.py:40-58`, twin
// `vae/sgm/modules/diffusionmodules/model.py:93-108`): F.pad(circular, W) + F.pad(zeros, H) +
// F.conv2d(pad 0) -- two materialised padded copies and a cuDNN call -- with one kernel:
//
//   D[m, n] = sum_{tap, c} A_tap[m, c] * Wt[tap][n][c]      m = output pixel, n = output channel
//
// * activations are channels-last fp16 (B, W+2, H, C), W-padded by the producer (the halo columns hold the circular
//   wrap); a 128-pixel M tile is 128/Ho whole azimuth columns.  For each (tap, 64-channel chunk) the producer warp
//   issues one TMA box: the zero pad on H is TMA out-of-bounds fill, stride 2 is the tensor map's element stride --
//   no padded copy ever exists;
// * weights [plane][tap][Cout][Cin] fp16 arrive by TMA as the K-major B operand;
// * both land SWIZZLE_128B in a multi-stage mbarrier ring; one thread issues tcgen05.mma (K=16) accumulating fp32
//   in TMEM; tcgen05.commit frees the stage;
// * epilogue warps read TMEM (tcgen05.ld), add bias + time-embedding + residual, store fp32 channels-last and
//   accumulate the GroupNorm channel-pair moments of the finished output.
//
// Operand precision (TERMS, chosen per layer by the engine):
//   3: split-fp16 ("fp16x3"): X = Xh + Xl, W = Wh + Wl (each part fp16), D += Xh*Wh + Xl*Wh + Xh*Wl -- ~22-bit operand
//      significands on the fp16 tensor pipe, fp32 accumulation in TMEM;
//   2: activations single fp16, weights split: D += Xh*Wh + Xh*Wl -- no X_lo plane is ever written or read (half the
//      operand bytes of the producer pass and of the ring), weights stay ~22 bit;
//   1: plain fp16 operands.
// Three kernel families: `conv_tc_kernel` (small layers: one tile per CTA, K split over a thread-block cluster),
// `conv_tc_persistent_kernel` (pixel-M tiles, more tiles than SMs), `conv_tc_wt_kernel` (roles swapped: weights are
// the M operand, N = 256 pixels; with pixel windows for 3x3 stride-1 layers).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "blocks.cuh"

namespace rldm {

static long long* g_conv_dbg = nullptr;

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                       // fp16 elements = one 128 B swizzle row
constexpr int kABytes = kBlockM * kBlockK * 2;    // 16 KB

__host__ __device__ constexpr int x_parts(int terms) { return terms == 3 ? 2 : 1; }
__host__ __device__ constexpr int w_parts(int terms) { return terms >= 2 ? 2 : 1; }

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
  int pad_h;        // low-side pad along H (beams); == pad_lo except for the phase convolutions of a folded nearest-2x upsampling
  int out_up;       // 1, or 2: the output pixel (w, h) is written to (2w + out_a, 2h + out_b) of a (B, 2Wo, 2Ho, Cout) tensor
  int out_a, out_b; // (role-swapped kernel only)
  int total_iters;  // main_iters + shortcut chunks
  int main_iters;   // (Cin/64) * ks*ks: K steps of the convolution proper; the rest are the fused 1x1 shortcut's
  double* stats;    // optional GroupNorm moments of the output: [B][stats_G][2]
  int stats_G;
  // role-swapped kernel with pixel windows only
  int a_part_bytes; // bytes of one operand part of a pixel window: (256 + 2*Ho) rows x 128 B
  int nb_stages;    // depth of the weight ring
  int units;        // (Cin/64) * 3 : one unit = (channel chunk, kernel column tj) = 3 taps
  long long* dbg;   // optional: clock64 timestamps written by CTA 0 (profiling aid, normally NULL)
  // small-layer kernel only: the epilogue also EMITS the next GroupNorm's operand (see conv_tc_kernel)
  __half* emit_out;           // fp16 W-padded operand (B, Wo+2, Ho, Cout) of the consumer, or NULL
  const float* emit_gamma;    // [Cout] affine of the consumer's GroupNorm
  const float* emit_beta;
  float emit_eps;
  float emit_inv_n;           // 1 / (pixels per image x channels per group)
  int emit_cpg;               // channels per group of that GroupNorm (multiple of 4, divides BLOCK_N)
  int emit_silu;
  int emit_circular;          // halo columns of the emitted operand: wrap (1) or zeros (0)
  int clm;                    // M tiles per image that share a cluster (1 unless emitting for images of > 128 pixels)
};

// Tensor maps of one launch.  a/alo/b: activation (hi, lo) and weights of the convolution.  a2/a2lo/b2: operand and
// weights of an optional 1x1 convolution over a second tensor with the same spatial grid whose product is
// accumulated into the same tile (ResnetBlock2D's conv_shortcut folded into conv2's K loop).
struct ConvMaps {
  CUtensorMap a, alo, b, a2, a2lo, b2;
};

// Operand production inside the small-layer kernel (replaces a rldm_prep launch in front of it): the CTA turns the fp32
// source(s) into exactly the part of the fp16 W-padded operand it will read -- the input window of its 128-pixel tile
// (halo columns included) x the 64-channel chunks of its own K slice -- writes it to the operand buffer and loads it
// back by TMA like any other operand.  Neighbouring CTAs write the same values to the halo columns they share; nothing
// crosses CTAs, so no grid-wide dependency is created.  `main`: the convolution's operand (GroupNorm-apply + SiLU +
// concat + nearest-2x + circular halo, as rldm_prep); `sc`: the raw operand of a folded 1x1 conv_shortcut.
struct ConvFused {
  PrepArgs main, sc;
  int main_on, sc_on;
};
constexpr int kFusedTabFloats = 1024;           // scale/shift of up to 512 channels
// small-layer kernel: TMA producer warp, MMA warp and eight epilogue warps (two per TMEM lane quadrant).  The epilogue
// of these launches is latency-bound code that one warp per scheduler runs at ~4 cycles per instruction: with two warps
// per scheduler the 1x1 projections lose ~1 us and the stride-2 convolutions 2-3 us per launch (UNet forward 1647 ->
// 1583 us).  Four warps for the 4- and 8-way K splits (16-32 rows per CTA) measured the same within noise.
// Sixteen warps (four per quadrant) for the un-split and two-way split launches measured no better (224.1 vs 224.4
// images/s): compile with -DRLDM_CONV_EPI16_MAXSPLIT=2 to re-check.
#ifndef RLDM_CONV_EPI16_MAXSPLIT
#define RLDM_CONV_EPI16_MAXSPLIT 0
#endif
__host__ __device__ constexpr int conv_epi_warps(int nsplit, int bn = 128) { return (bn == 128 && nsplit <= RLDM_CONV_EPI16_MAXSPLIT) ? 16 : 8; }
__host__ __device__ constexpr int conv_threads(int nsplit, int bn = 128) { return 64 + 32 * conv_epi_warps(nsplit, bn); }
constexpr int kConvEpiWarpsMax = 16;
// in-kernel operand production: items per thread and step / software pipelining of prep_range (the body runs ONCE per
// launch on one warp per scheduler: a long unrolled body is bound by instruction fetch, `stall_no_inst` in ncu)
#ifndef RLDM_OWN_U
#define RLDM_OWN_U 2
#endif
#ifndef RLDM_OWN_PIPE
#define RLDM_OWN_PIPE true
#endif
constexpr int kOwnU = RLDM_OWN_U;
constexpr bool kOwnPipe = RLDM_OWN_PIPE;

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

// One K step (64 channels) of the accumulation D += X W for the pixel-M kernels: a_addr = [X_hi][X_lo] tiles of
// 128 rows, b_addr = [W_hi][W_lo] tiles of BLOCK_N rows (lo parts present as TERMS says).
template <int BLOCK_N, int TERMS>
__device__ __forceinline__ void issue_kstep(uint32_t d_tmem, uint32_t a_addr, uint32_t b_addr, uint32_t first) {
  constexpr uint32_t idesc = umma_idesc_f16(kBlockM, BLOCK_N);
  constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  const uint64_t a_desc = umma_desc_sw128(a_addr);
  const uint64_t b_desc = umma_desc_sw128(b_addr);
#pragma unroll
  for (int k = 0; k < kBlockK / 16; ++k)     // advancing K by 16 fp16 = 32 B inside the 128 B swizzle row: +2 in (addr>>4)
    umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (first == 0u || k != 0) ? 1u : 0u);
  if (TERMS == 3) {
    const uint64_t al_desc = umma_desc_sw128(a_addr + kABytes);
#pragma unroll
    for (int k = 0; k < kBlockK / 16; ++k) umma_f16(d_tmem, al_desc + 2 * k, b_desc + 2 * k, idesc, 1u);   // X_lo * W_hi
  }
  if (TERMS >= 2) {
    const uint64_t bl_desc = umma_desc_sw128(b_addr + kBBytes);
#pragma unroll
    for (int k = 0; k < kBlockK / 16; ++k) umma_f16(d_tmem, a_desc + 2 * k, bl_desc + 2 * k, idesc, 1u);   // X_hi * W_lo
  }
}

// NSPLIT   : CTAs of the cluster that share one output tile (split K); == gridDim.z.
//
// Latency structure (these layers are small: the whole kernel is a handful of microseconds, so every serial
// round trip counts).  The four epilogue warps are idle during the K loop, so they fetch bias, time embedding and
// ALL residual rows they will need into registers right after griddepcontrol.wait -- the L2 latency of the
// residual hides under the mainloop.  The split-K reduction pulls the partial tiles of the peer CTAs through
// distributed shared memory in batches of 16 independent 16 B loads per lane (fixed summation order:
// deterministic), instead of one dependent load at a time.
//
// EMIT (p.emit_out != NULL): the consumer of this convolution's output is a GroupNorm (+ SiLU) in front of another
// convolution, i.e. a rldm_prep launch that could only start once every CTA of this grid has added its moments.  Here
// the cluster is shaped so that it covers WHOLE IMAGES for its output-channel tile -- (p.clm M tiles) x (NSPLIT K
// slices), <= 8 CTAs -- so the (image, group) moments are complete inside the cluster: every epilogue warp publishes
// the moments of its finished rows in shared memory, one cluster barrier later each thread sums the <= 32 partials
// of its own (image, group) through DSMEM, normalises the values it still holds in registers and writes the fp16
// W-padded operand (halo columns included) the next convolution reads.  One prep launch and one kernel boundary less
// per GroupNorm; the fp32 output (p.out) is still written when the residual stream needs it.
template <int BLOCK_N, int STAGES, int TERMS, int NSPLIT>
__global__ void __launch_bounds__(conv_threads(NSPLIT, BLOCK_N), 1)
conv_tc_kernel(const __grid_constant__ ConvMaps tm, const ConvParams p, const __grid_constant__ ConvFused fz) {
  constexpr int kEpiWarps = conv_epi_warps(NSPLIT, BLOCK_N), kThreads = conv_threads(NSPLIT, BLOCK_N);
  constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  constexpr int XP = x_parts(TERMS), WP = w_parts(TERMS);
  constexpr int kStageBytes = XP * kABytes + WP * kBBytes;    // [X_hi][X_lo][W_hi][W_lo]
  constexpr int kBOff = XP * kABytes;
  constexpr int kStagePitch = BLOCK_N + 4;                     // floats per row of the epilogue staging tile
  static_assert(kBlockM * kStagePitch * 4 <= STAGES * kStageBytes, "staging tile must fit in the pipeline stages");
  static_assert(NSPLIT == 1 || (kBlockM + kBlockM / NSPLIT) * kStagePitch * 4 <= STAGES * kStageBytes,
                "the kept rows of an emitting launch must fit behind the staging tile");
  // epilogue geometry: a warp instruction covers kRowsPerIter rows of BLOCK_N floats (float4 per lane)
  constexpr int kLanesPerRow = BLOCK_N / 4;               // 32 (BN=128) or 16 (BN=64)
  constexpr int kRowsPerIter = 32 / kLanesPerRow;         // 1 or 2
  constexpr int kRowsCta = kBlockM / NSPLIT;              // rows this CTA finalises
  constexpr int kRowsWarp = kRowsCta / kEpiWarps;     // contiguous rows per epilogue warp
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
  float* red_s = reinterpret_cast<float*>(tmem_ptr + 4);       // [kEpiWarps][BLOCK_N/2]
  float* red_q = red_s + kEpiWarps * (BLOCK_N / 2);
  int* red_b = reinterpret_cast<int*>(red_q + kEpiWarps * (BLOCK_N / 2));
  float* fused_tab = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(red_b + kEpiWarps) + 15) & ~static_cast<uintptr_t>(15));   // [kFusedTabFloats], float4 reads

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
    if (XP > 1) tma_prefetch_desc(&tm.alo);
    tma_prefetch_desc(&tm.b);
    if (p.total_iters > p.main_iters) {
      tma_prefetch_desc(&tm.a2);
      if (XP > 1) tma_prefetch_desc(&tm.a2lo);
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
      if (WP > 1) tma_load_2d(b_dst + kBBytes, &tm.b, bar, chunk * kBlockK, (taps + tap) * p.Cout + n0);
    } else {
      const int chunk = it - p.main_iters;
      tma_load_2d(b_dst, &tm.b2, bar, chunk * kBlockK, n0);
      if (WP > 1) tma_load_2d(b_dst + kBBytes, &tm.b2, bar, chunk * kBlockK, p.Cout + n0);
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

  if (fz.main_on | fz.sc_on) {
    // ---- in-kernel operand production (see ConvFused): all threads -- nobody has anything else to do until the
    // operand exists (the first weight tiles are already in flight) ----
    {
      const int tid = threadIdx.x;
      const int taps = p.ks * p.ks;
      const int q0 = m0 / p.Ho, b0 = q0 / p.Wo, wo0 = q0 - b0 * p.Wo;
      const int nb = p.pix_per_img >= kBlockM ? 1 : kBlockM / p.pix_per_img;
      const int ncols = p.pix_per_img >= kBlockM ? kBlockM / p.Ho : p.Wo;
      const int n_img = p.M_total / p.pix_per_img;
      const int Hop = p.Ho * p.stride;                      // operand rows per column
      const int b_end = min(b0 + nb, n_img);
      const int main_end = min(it1, p.main_iters);
      if (fz.main_on && it0 < main_end) {
        const int ch_lo = (it0 / taps) * kBlockK, ch_hi = ((main_end - 1) / taps + 1) * kBlockK;
        const int first = p.stride * wo0 - p.pad_lo + 1;    // padded column of tap ti = 0
        const int col_lo = max(first, 0), col_hi = min(first + (p.ks - 1) + ncols * p.stride, p.W_in + 2);
        for (int b = b0; b < b_end; ++b) {
          prep_range<false, kOwnU, kOwnPipe>(fz.main, b, col_lo * Hop, col_hi * Hop, ch_lo, ch_hi, fused_tab, tid, kThreads, 0);
          __syncthreads();                                  // scale / shift table reusable
        }
      }
      const int sc_begin = max(it0, p.main_iters);
      if (fz.sc_on && sc_begin < it1) {
        const int ch_lo = (sc_begin - p.main_iters) * kBlockK, ch_hi = (it1 - p.main_iters) * kBlockK;
        for (int b = b0; b < b_end; ++b)
          prep_range<false, kOwnU, kOwnPipe>(fz.sc, b, (wo0 + 1) * p.Ho, (wo0 + 1 + ncols) * p.Ho, ch_lo, ch_hi, fused_tab, tid, kThreads, 0);
      }
      asm volatile("fence.proxy.async;" ::: "memory");     // generic-proxy stores -> TMA (async proxy) reads below
    }
    __syncthreads();
    if (warp == 0) asm volatile("fence.proxy.async;" ::: "memory");
    if (dbg && threadIdx.x == 0) p.dbg[9] = clock64();
  }

  // epilogue coordinates (meaningful for warps >= 2)
  const int ew = (warp - 2) & (kEpiWarps - 1);
  const int r_begin = blockIdx.z * kRowsCta + ew * kRowsWarp;     // first tile row this warp finalises
  const int col = (lane % kLanesPerRow) * 4;
  const int rsub = lane / kLanesPerRow;
  const int m_first = m0 + r_begin;
  // all rows of one warp lie in one image (pix_per_img is a power of two >= 64, or whole images per tile row group)
  const int bimg = min(m_first, p.M_total - 1) / p.pix_per_img;
  float4 res[kPerLane];
  float4 emit_g = make_float4(0.f, 0.f, 0.f, 0.f), emit_b = emit_g;      // consumer GroupNorm affine of this thread's 4 columns
  float es = 0.f, eq = 0.f;                                              // this thread's share of its group's moments
  const int cl_x = p.clm > 1 ? static_cast<int>(blockIdx.x) % p.clm : 0;   // position among the image's M tiles in the cluster

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
      const int h_in = main ? tj - p.pad_h : 0;
      const int w_in = main ? p.stride * wo0 + ti - p.pad_lo + 1 : wo0 + 1;
      if (lane == (XP == 1 ? 0 : 1)) tma_load_4d(a_dst, main ? &tm.a : &tm.a2, &full_bar[s], chunk * kBlockK, h_in, w_in, b0);
      if (XP > 1 && lane == 2)
        tma_load_4d(a_dst + kABytes, main ? &tm.alo : &tm.a2lo, &full_bar[s], chunk * kBlockK, h_in, w_in, b0);
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) ============================================
    if (elect_one()) {   // elect.sync: ptxas knows exactly one lane is active (no per-MMA waterfall loop)
      for (int i = 0; i < n_it; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (dbg && i == 0) p.dbg[2] = clock64();
        const uint32_t a_addr = smem_u32(smem + s * kStageBytes);
        issue_kstep<BLOCK_N, TERMS>(tmem_base, a_addr, a_addr + kBOff, i == 0 ? 1u : 0u);
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
    if (p.emit_out) {
      emit_g = __ldg(reinterpret_cast<const float4*>(p.emit_gamma + n0 + col));
      emit_b = __ldg(reinterpret_cast<const float4*>(p.emit_beta + n0 + col));
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
    float* stage_row = reinterpret_cast<float*>(smem) + (q * 32 + lane) * kStagePitch;
    // kEpiWarps / 4 warps share a lane quadrant; each takes its share of the 32-column chunks
    constexpr int kChunksPerWarp = (BLOCK_N / 32) / (kEpiWarps / 4);
    static_assert(kChunksPerWarp >= 1, "more epilogue warps than 32-column chunks per quadrant");
    const int nc0 = ((warp - 2) >> 2) * kChunksPerWarp;
#pragma unroll 1
    for (int nc = nc0; nc < nc0 + kChunksPerWarp; ++nc) {
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
    asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");        // only the epilogue warps touch the staging tile
  }
  if (warp >= 2) {
    if (dbg && threadIdx.x == 64) p.dbg[7] = clock64();
    float s01 = 0.f, q01 = 0.f, s23 = 0.f, q23 = 0.f;     // moments of channel pairs (0,1) and (2,3)
    const uint32_t stage_u32 = smem_u32(smem);
    if (NSPLIT == 1) {
      // four rows per step: the shared-memory loads first, then the arithmetic, then the stores (one warp per scheduler
      // runs this: a row at a time is a chain of dependent latencies)
      constexpr int kRB = kPerLane < 4 ? kPerLane : 4;
      const bool keep = p.emit_out != nullptr;
#pragma unroll
      for (int u0 = 0; u0 < kPerLane; u0 += kRB) {
        float4 t[kRB];
#pragma unroll
        for (int j = 0; j < kRB; ++j)
          t[j] = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(smem) + (r_begin + (u0 + j) * kRowsPerIter + rsub) * kStagePitch + col);
#pragma unroll
        for (int j = 0; j < kRB; ++j) {
          const int r = r_begin + (u0 + j) * kRowsPerIter + rsub;
          const int m = m0 + r;
          float4 v = res[u0 + j];
          v.x += t[j].x; v.y += t[j].y; v.z += t[j].z; v.w += t[j].w;
          if (keep)                                          // kept for the emit pass (in place: nobody else reads this tile)
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(smem) + r * kStagePitch + col) = v;
          const bool live = m < p.M_total;
          if (live && p.out) *reinterpret_cast<float4*>(p.out + static_cast<size_t>(m) * p.Cout + n0 + col) = v;
          const float lm = live ? 1.f : 0.f;
          s01 += lm * (v.x + v.y); q01 += lm * (v.x * v.x + v.y * v.y);
          s23 += lm * (v.z + v.w); q23 += lm * (v.z * v.z + v.w * v.w);
        }
      }
    } else {
      constexpr int kUB = (16 / NSPLIT) < kPerLane ? (16 / NSPLIT) : kPerLane;   // rows per batch of <= 16 loads
      uint32_t rbase[NSPLIT];
#pragma unroll
      for (int sidx = 0; sidx < NSPLIT; ++sidx) rbase[sidx] = mapa_u32(stage_u32, cl_x + p.clm * sidx);   // rank = x + clm * z
#pragma unroll
      for (int u0 = 0; u0 < kPerLane; u0 += kUB) {
        float4 part[kUB][NSPLIT];
#pragma unroll
        for (int ub = 0; ub < kUB; ++ub) {
          const int r = r_begin + (u0 + ub) * kRowsPerIter + rsub;
          const uint32_t off = static_cast<uint32_t>(r * kStagePitch + col) * 4u;
#pragma unroll
          for (int sidx = 0; sidx < NSPLIT; ++sidx) part[ub][sidx] = ld_dsmem_f4(rbase[sidx] + off);
        }
#pragma unroll
        for (int ub = 0; ub < kUB; ++ub) {
          const int m = m0 + r_begin + (u0 + ub) * kRowsPerIter + rsub;
          float4 v = res[u0 + ub];
#pragma unroll
          for (int sidx = 0; sidx < NSPLIT; ++sidx) {
            v.x += part[ub][sidx].x; v.y += part[ub][sidx].y; v.z += part[ub][sidx].z; v.w += part[ub][sidx].w;
          }
          if (p.emit_out)                                    // kept for the emit pass, behind the staging tile the peers still read
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(smem) + (kBlockM + r_begin - blockIdx.z * kRowsCta + (u0 + ub) * kRowsPerIter + rsub) * kStagePitch + col) = v;
          if (m < p.M_total) {
            if (p.out) *reinterpret_cast<float4*>(p.out + static_cast<size_t>(m) * p.Cout + n0 + col) = v;
            s01 += v.x + v.y; q01 += v.x * v.x + v.y * v.y;
            s23 += v.z + v.w; q23 += v.z * v.z + v.w * v.w;
          }
        }
      }
    }
    if (dbg && threadIdx.x == 64) p.dbg[8] = clock64();
    es = s01 + s23; eq = q01 + q23;
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
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");      // the epilogue warps only
      const int t = ew * 32 + lane;
      if (t < nslots) {
        const int g = n0 / 2 + t;
        double ds = 0.0, dq = 0.0;
        int cur = red_b[0];
        for (int e = 0; e < kEpiWarps; ++e) {
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
  const bool clustered = NSPLIT > 1 || p.clm > 1;
  if (p.emit_out) {
    // ---- EMIT: (image, group) moments complete inside the cluster -> normalise own rows -> fp16 operand ----
    // per epilogue warp: [kEpiWarps][32] (sum, sum of squares) of the warp's rows per group of the tile; then folded
    // inside the CTA into [2 image slots][32] (a CTA's rows touch at most two images: slot = image - image of its first
    // row), which is what the other CTAs of the cluster read: ONE DSMEM round trip of <= 8 loads per thread
    float2* pub = reinterpret_cast<float2*>(fused_tab);                   // [kEpiWarps][32]
    float2* pub2 = pub + kEpiWarps * 32;                                  // [2][32]
    static_assert((kEpiWarps + 2) * 32 * 2 <= kFusedTabFloats, "moment exchange tables exceed the scratch area");
    const int lpg = p.emit_cpg >> 2;                        // lanes (column quads) per group
    const int gl = col / p.emit_cpg;                        // group of this thread's columns inside the tile
    const int cta_first = m0 + static_cast<int>(blockIdx.z) * kRowsCta;   // first row this CTA finalises
    const int pix_sh = 31 - __clz(p.pix_per_img);                         // pixels per image: a power of two (host check)
    if (warp >= 2) {
      for (int o = 1; o < lpg; o <<= 1) {
        es += __shfl_xor_sync(0xffffffffu, es, o); eq += __shfl_xor_sync(0xffffffffu, eq, o);
      }
      if (kRowsPerIter == 2) { es += __shfl_xor_sync(0xffffffffu, es, 16); eq += __shfl_xor_sync(0xffffffffu, eq, 16); }
      if (rsub == 0 && ((lane % kLanesPerRow) % lpg) == 0) pub[ew * 32 + gl] = make_float2(es, eq);
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
      const int t = ew * 32 + lane;
      if (t < 64) {                                          // (slot, group): fold the warps of that image
        const int slot = t >> 5, g = t & 31;
        const int img = (cta_first >> pix_sh) + slot;
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int w = 0; w < kEpiWarps; ++w) {
          const int mf = cta_first + w * kRowsWarp;          // all rows of a warp lie in one image
          if (mf < p.M_total && (mf >> pix_sh) == img) { const float2 v = pub[w * 32 + g]; acc.x += v.x; acc.y += v.y; }
        }
        pub2[slot * 32 + g] = acc;
      }
    }
    if (clustered) cluster_sync_all();
    else if (warp >= 2) asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
    if (dbg && threadIdx.x == 64) p.dbg[10] = clock64();
    if (warp >= 2) {
      const int csize = clustered ? NSPLIT * p.clm : 1;
      const uint32_t pub2_u32 = smem_u32(pub2);
      // the image slot of this thread's image in cluster rank rk = (x, z) follows from the geometry: the first row that
      // CTA finalises is (tile blockIdx.x - cl_x + x) * 128 + z * kRowsCta
      const int tile0 = static_cast<int>(blockIdx.x) - cl_x;
      const int clm_sh = 31 - __clz(p.clm);
      float2 pq[8];
#pragma unroll
      for (int rk = 0; rk < 8; ++rk) {
        pq[rk] = make_float2(0.f, 0.f);
        if (rk < csize) {
          const int x = rk & (p.clm - 1), z = rk >> clm_sh;       // clm is a power of two (host check)
          const int first = (tile0 + x) * kBlockM + z * kRowsCta;
          const int slot = bimg - (first >> pix_sh);
          if (first < p.M_total && slot >= 0 && slot < 2) {
            const uint32_t base = clustered ? mapa_u32(pub2_u32, rk) : pub2_u32;
            asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];"
                         : "=f"(pq[rk].x), "=f"(pq[rk].y) : "r"(base + static_cast<uint32_t>(slot * 32 + gl) * 8u) : "memory");
          }
        }
      }
      // (fp32 is ample next to the fp16 operand: <= 16 partials of <= 1024 values each)
      const float S = ((pq[0].x + pq[1].x) + (pq[2].x + pq[3].x)) + ((pq[4].x + pq[5].x) + (pq[6].x + pq[7].x));
      const float Q = ((pq[0].y + pq[1].y) + (pq[2].y + pq[3].y)) + ((pq[4].y + pq[5].y) + (pq[6].y + pq[7].y));
      const float mu = S * p.emit_inv_n;
      const float var = fmaxf(fmaf(-mu, mu, Q * p.emit_inv_n), 0.f);
      const float rstd = rsqrtf(var + p.emit_eps);
      const float4 sc = make_float4(rstd * emit_g.x, rstd * emit_g.y, rstd * emit_g.z, rstd * emit_g.w);
      const float4 sf = make_float4(emit_b.x - mu * sc.x, emit_b.y - mu * sc.y, emit_b.z - mu * sc.z, emit_b.w - mu * sc.w);
      if (dbg && threadIdx.x == 64) p.dbg[11] = clock64();
      // W-padded operand (B, Wo+2, Ho, Cout): pixel pin = w*Ho + h of the image sits at padded pixel pin + Ho; padded
      // column 0 holds image column Wo-1 and padded column Wo+1 image column 0 (zeros when the consumer does not wrap)
      __half* obase = p.emit_out + static_cast<size_t>(bimg) * (p.Wo + 2) * p.Ho * p.Cout + n0 + col;
      const int pin0 = m_first - bimg * p.pix_per_img + rsub;
      const int rows_left = p.M_total - m_first - rsub;          // rows u with u * kRowsPerIter < rows_left exist
      const int last_col0 = p.pix_per_img - p.Ho;                 // first pixel of image column Wo-1
      // rows of this warp in the kept tile: NSPLIT == 1 in place (tile row r), else local row behind the staging tile
      const float* kept = reinterpret_cast<const float*>(smem) +
                          (NSPLIT == 1 ? r_begin : kBlockM + r_begin - static_cast<int>(blockIdx.z) * kRowsCta) * kStagePitch +
                          rsub * kStagePitch + col;
      // four rows per step: loads, then the arithmetic of all sixteen values, then the stores -- one warp per scheduler
      // runs this, a row at a time would be a chain of dependent MUFU latencies
      constexpr int kEB = kPerLane < 4 ? kPerLane : 4;
#pragma unroll 1
      for (int u0 = 0; u0 < kPerLane; u0 += kEB) {
        float4 y[kEB];
#pragma unroll
        for (int j = 0; j < kEB; ++j) y[j] = *reinterpret_cast<const float4*>(kept + (u0 + j) * kRowsPerIter * kStagePitch);
#pragma unroll
        for (int j = 0; j < kEB; ++j) {
          y[j].x = fmaf(y[j].x, sc.x, sf.x); y[j].y = fmaf(y[j].y, sc.y, sf.y);
          y[j].z = fmaf(y[j].z, sc.z, sf.z); y[j].w = fmaf(y[j].w, sc.w, sf.w);
        }
        if (p.emit_silu) {
#pragma unroll
          for (int j = 0; j < kEB; ++j) { y[j].x = silu_f(y[j].x); y[j].y = silu_f(y[j].y); y[j].z = silu_f(y[j].z); y[j].w = silu_f(y[j].w); }
        }
#pragma unroll
        for (int j = 0; j < kEB; ++j) {
          const int pin = pin0 + (u0 + j) * kRowsPerIter;
          const __half2 h01 = __floats2half2_rn(y[j].x, y[j].y), h23 = __floats2half2_rn(y[j].z, y[j].w);
          uint2 pk = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
          const bool live = (u0 + j) * kRowsPerIter < rows_left;
          if (live) *reinterpret_cast<uint2*>(obase + static_cast<size_t>(pin + p.Ho) * p.Cout) = pk;
          if (!p.emit_circular) pk = make_uint2(0u, 0u);
          if (live && pin < p.Ho) *reinterpret_cast<uint2*>(obase + static_cast<size_t>(p.pix_per_img + p.Ho + pin) * p.Cout) = pk;
          if (live && pin >= last_col0) *reinterpret_cast<uint2*>(obase + static_cast<size_t>(pin - last_col0) * p.Cout) = pk;
        }
      }
      if (dbg && threadIdx.x == 64) p.dbg[12] = clock64();
    }
  }
  if (clustered) cluster_sync_all();      // nobody exits while a peer may still read its staging tile / published moments
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<BLOCK_N>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// Persistent variant of the per-tap kernel for layers with more tiles than SMs (no K split).
//
// One CTA per SM walks over tiles; the fp32 accumulator is DOUBLE-BUFFERED in TMEM (2 x BLOCK_N columns), so the four
// epilogue warps drain tile k (TMEM -> 32-column shared staging slab -> coalesced global, bias/temb/residual/GroupNorm
// moments) while the producer and MMA warps are already running the K loop of tile k+1.  The staging slab is private
// to the epilogue (not aliased with the pipeline stages).
//
// MT = 2: a work unit is TWO consecutive 128-pixel M tiles that share every weight tile: a stage holds
// [X0_hi X0_lo X1_hi X1_lo W_hi W_lo].  The K loop of these kernels is bound by the operand stream into the SM
// (~64 B/clk); sharing W over two tiles cuts the bytes per MMA by 25 %, and it halves the number of units.
// TMEM: 2 units x MT x BLOCK_N columns.  (Used for the layers the role-swapped kernel below does not take: Cout = 64,
// stride 2, ragged pixel counts.)
template <int BLOCK_N, int STAGES, int TERMS, int MT>
__global__ void __launch_bounds__(192, 1)
conv_tc_persistent_kernel(const __grid_constant__ ConvMaps tm, const ConvParams p) {
  constexpr int XP = x_parts(TERMS), WP = w_parts(TERMS);
  constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  constexpr int kATile = XP * kABytes;                          // one M tile of a stage: [X_hi][X_lo]
  constexpr int kStageBytes = MT * kATile + WP * kBBytes;
  constexpr int kBOff = MT * kATile;
  constexpr int kSlabPitch = 36;                                // floats per row of the 32-column staging slab
  constexpr int kSlabBytes = kBlockM * kSlabPitch * 4;          // 18 KB
  constexpr int kChunks = BLOCK_N / 32;
  constexpr int kAccCols = MT * BLOCK_N;                        // TMEM columns of one unit's accumulators
  static_assert(2 * kAccCols <= 512, "double-buffered accumulators must fit TMEM");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* tail = smem + STAGES * kStageBytes;
  float* slab = reinterpret_cast<float*>(tail);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail + kSlabBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = full_bar + 2 * STAGES;   // [2]
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
  const int n_it = p.total_iters;

  pdl_trigger_conv_early();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm.a);
    if (XP > 1) tma_prefetch_desc(&tm.alo);
    tma_prefetch_desc(&tm.b);
    if (p.total_iters > p.main_iters) {
      tma_prefetch_desc(&tm.a2);
      if (XP > 1) tma_prefetch_desc(&tm.a2lo);
      tma_prefetch_desc(&tm.b2);
    }
    for (int s = 0; s < 2 * STAGES; ++s) mbar_init(&full_bar[s], 1);
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

  if (warp == 0) {
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
        const bool main = it < p.main_iters;
        const int chunk = main ? it / taps : it - p.main_iters;
        const int tap = it - chunk * taps;
        const int ti = tap / p.ks, tj = tap - ti * p.ks;
        const int kc = chunk * kBlockK;                            // first channel of this stage
        const uint32_t a_dst = smem_u32(smem + s * kStageBytes);
        // shortcut K steps: centre tap of the second tensor (1x1, stride 1, same grid as the output)
        const int h_in = main ? tj - p.pad_h : 0;
        const int w_off = main ? ti - p.pad_lo + 1 : 1;
        const int w_mul = main ? p.stride : 1;
        if (lane == 0) {
          mbar_arrive_expect_tx(&full_bar[s], kStageBytes);
          if (main) {
            tma_load_2d(a_dst + kBOff, &tm.b, &full_bar[s], kc, tap * p.Cout + n0);
            if (WP > 1) tma_load_2d(a_dst + kBOff + kBBytes, &tm.b, &full_bar[s], kc, (taps + tap) * p.Cout + n0);
          } else {
            tma_load_2d(a_dst + kBOff, &tm.b2, &full_bar[s], kc, n0);
            if (WP > 1) tma_load_2d(a_dst + kBOff + kBBytes, &tm.b2, &full_bar[s], kc, p.Cout + n0);
          }
        }
        // lanes 1.. : one box per (M tile, operand part)
        if (lane >= 1 && lane <= MT * XP) {
          const int mt = (lane - 1) / XP, part = (lane - 1) % XP;
          const CUtensorMap* map = main ? (part ? &tm.alo : &tm.a) : (part ? &tm.a2lo : &tm.a2);
          tma_load_4d(a_dst + mt * kATile + part * kABytes, map, &full_bar[s], kc, h_in, w_mul * wo0[mt] + w_off, b0[mt]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: alternates between the two TMEM accumulator sets ===========
    if (elect_one()) {
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
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
            issue_kstep<BLOCK_N, TERMS>(tmem_base + acc * kAccCols + mt * BLOCK_N, a_addr + mt * kATile, a_addr + kBOff,
                                        it == 0 ? 1u : 0u);
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
// MMA), so every TMA fill of the ring steals tensor cycles.  Here the 128 output CHANNELS are the M side (A operand =
// weight tile, 128 rows x 64 channels, K-major) and 256 PIXELS the N side (B operand = the two 128-pixel activation
// tiles stored back to back = one 256-row K-major tile): one 128x256x16 MMA reads 4 + 8 KB in 128 clocks = 0.75
// wavefronts per clock for the same FLOPs.  The accumulator is [channel lane][pixel column] (2 x 256 TMEM columns,
// double-buffered), which also makes the epilogue simpler: a thread owns ONE output channel, `tcgen05.ld` hands it 32
// pixels, lanes = 32 consecutive channels of one pixel, so stores and residual loads are coalesced 128 B rows straight
// from registers (no shared staging slab, no bar.sync), bias/temb are per-thread scalars and the GroupNorm
// channel-pair moments are per-thread sums plus one shuffle.  Stage = [X_hi 256 rows][X_lo][W_hi][W_lo] (parts as
// TERMS says): 96 / 64 / 48 KB, 2 / 3 / 4 stages.
//
// HALO = true (3x3, stride 1, symmetric pad, Ho in 8..32, no fused shortcut): the pixel operand of the three taps
// ti = 0,1,2 of one kernel column tj is ONE window of (256 + 2*Ho) rows (the unit's columns plus one halo column on
// each side; tap ti = the window shifted by ti*Ho rows = whole 1024 B swizzle atoms), so the ring fills drop from
// 96 KB to 56 KB per tap (TERMS = 3).  Two rings: two pixel windows ([hi][lo], one per (64-channel chunk, tj)) and
// p.nb_stages weight entries of ONE operand part each: W_hi of a tap feeds W_hi X_hi (+ W_hi X_lo) and is released
// before W_lo is needed.
constexpr int kWtEpiWarps = 8;                  // two warps per TMEM lane quadrant, four 32-pixel chunks each
constexpr int kWtThreads = 64 + 32 * kWtEpiWarps;
__host__ __device__ constexpr int wt_stages(int terms) { return terms == 3 ? 2 : (terms == 2 ? 3 : 4); }
template <int TERMS, bool HALO = false>
__global__ void __launch_bounds__(kWtThreads, 1)
conv_tc_wt_kernel(const __grid_constant__ ConvMaps tm, const ConvParams p) {
  constexpr int XP = x_parts(TERMS), WP = w_parts(TERMS);
  constexpr int kCoutTile = 128, kPix = 256, STAGES = wt_stages(TERMS);
  constexpr int kMaxNW = 8;                                     // HALO: upper bound of the weight ring depth
  constexpr int kRingBars = HALO ? 4 + 2 * kMaxNW : 2 * STAGES;
  constexpr int kXPart = kPix * kBlockK * 2;                    // 32 KB: 256 pixel rows x 128 B
  constexpr int kWPart = kCoutTile * kBlockK * 2;               // 16 KB
  constexpr int kWOff = XP * kXPart;
  constexpr int kStageBytes = XP * kXPart + WP * kWPart;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int x_stage = HALO ? XP * p.a_part_bytes : 0;           // HALO: one pixel-window stage, [hi][lo]
  const int NW = HALO ? p.nb_stages : 1;                          // (1: the HALO branches are dead code otherwise)
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
    if (XP > 1) tma_prefetch_desc(&tm.alo);
    tma_prefetch_desc(&tm.b);
    if (p.total_iters > p.main_iters) {
      tma_prefetch_desc(&tm.a2);
      if (XP > 1) tma_prefetch_desc(&tm.a2lo);
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
            if (XP > 1) tma_load_4d(x_dst + p.a_part_bytes, &tm.alo, &x_full[sx], chunk * kBlockK, tj - 1, wo0, b0);
          }
          for (int e = 0; e < 3 * WP; ++e, ++gw) {          // (ti, part): W_hi then W_lo of each tap
            const int sw = gw % NW;
            mbar_wait(&w_empty[sw], ((gw / NW) & 1) ^ 1);
            if (lane == 0) {
              const int ti = e / WP, part = e - ti * WP;
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
              // (the low-order plane first: with the planes issued the other way round compute-sanitizer's memcheck
              // flags the fourth K slice of the second group as an out-of-range shared address, although both groups
              // read the same rows of two adjacent windows -- the report follows the instruction slot, not the address)
              if (XP > 1) {
                const uint64_t xl_desc = umma_desc_sw128(x_base + ti * p.Ho * 128 + p.a_part_bytes);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) umma_f16(d_tmem, w_desc + 2 * kk, xl_desc + 2 * kk, idesc, (u | ti | kk) != 0);   // W_hi X_lo
              }
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) umma_f16(d_tmem, w_desc + 2 * kk, x_desc + 2 * kk, idesc, XP > 1 ? 1u : ((u | ti | kk) != 0));
              umma_commit(&w_empty[sw]);
              ++gw;
            }
            if (WP > 1) {
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
        const int h_in = main ? tj - p.pad_h : 0;
        const int w_off = main ? ti - p.pad_lo + 1 : 1;
        const int w_mul = main ? p.stride : 1;
        if (lane == 0) {
          mbar_arrive_expect_tx(&full_bar[s], kStageBytes);
          if (main) {
            tma_load_2d(dst + kWOff, &tm.b, &full_bar[s], kc, tap * p.Cout + n0);
            if (WP > 1) tma_load_2d(dst + kWOff + kWPart, &tm.b, &full_bar[s], kc, (taps + tap) * p.Cout + n0);
          } else {
            tma_load_2d(dst + kWOff, &tm.b2, &full_bar[s], kc, n0);
            if (WP > 1) tma_load_2d(dst + kWOff + kWPart, &tm.b2, &full_bar[s], kc, p.Cout + n0);
          }
        }
        // lanes 1.. : one 128-pixel box per (tile, operand part); the two tiles of a part are adjacent = 256 rows
        if (lane >= 1 && lane <= 2 * XP) {
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
          if (XP > 1) {
            const uint64_t xl_desc = umma_desc_sw128(base + kXPart);
#pragma unroll
            for (int kk = 0; kk < kBlockK / 16; ++kk) umma_f16(d_tmem, w_desc + 2 * kk, xl_desc + 2 * kk, idesc, 1u);     // W_hi X_lo
          }
          if (WP > 1) {
            const uint64_t wl_desc = umma_desc_sw128(base + kWOff + kWPart);
#pragma unroll
            for (int kk = 0; kk < kBlockK / 16; ++kk) umma_f16(d_tmem, wl_desc + 2 * kk, x_desc + 2 * kk, idesc, 1u);     // W_lo X_hi
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
      // folded nearest-2x upsampling (one of the four output phases): pixel pin = w*Ho + h of the image -> pixel
      // (2w + a)*2Ho + 2h + b of an image four times as large (Ho is a power of two)
      const int up_sh = 31 - __clz(p.Ho);
      const int pin0 = m0 - bimg * p.pix_per_img;
      float* outp_up = p.out + static_cast<size_t>(bimg) * 4 * p.pix_per_img * p.Cout + c;
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
        if (p.out_up == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) outp[static_cast<size_t>(ch * 32 + j) * p.Cout] = v[j];
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int pin = pin0 + ch * 32 + j;
            const int w = pin >> up_sh, h = pin & (p.Ho - 1);
            outp_up[static_cast<size_t>(((2 * w + p.out_a) << (up_sh + 1)) + 2 * h + p.out_b) * p.Cout] = v[j];
          }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
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

constexpr int kSmemLimit = 232448;      // 227 KB per CTA
constexpr int conv_stage_bytes(int bn, int terms) { return x_parts(terms) * kABytes + w_parts(terms) * bn * kBlockK * 2; }
constexpr int conv_stages(int bn, int terms) { return (bn == 128 && terms == 3) ? 3 : 4; }
constexpr int conv_smem(int bn, int terms) {
  // pipeline stages (+ alignment slack) + barriers/TMEM pointer + the per-warp GroupNorm-moment scratch
  return conv_stages(bn, terms) * conv_stage_bytes(bn, terms) + 1024 + 256 + 2 * kConvEpiWarpsMax * (bn / 2) * 4 + 64 + kFusedTabFloats * 4 + 16;
}
constexpr int pers_stage_bytes(int bn, int terms, int mt) { return mt * x_parts(terms) * kABytes + w_parts(terms) * bn * kBlockK * 2; }
constexpr int pers_fixed(int bn) { return kBlockM * 36 * 4 + 256 + 2 * 4 * (bn / 2) * 4 + 64 + 1024; }
constexpr int pers_stages(int bn, int terms, int mt) {
  return (kSmemLimit - pers_fixed(bn)) / pers_stage_bytes(bn, terms, mt) > 8 ? 8 : (kSmemLimit - pers_fixed(bn)) / pers_stage_bytes(bn, terms, mt);
}

template <int BLOCK_N, int TERMS, int NSPLIT>
static int launch_conv_n(const ConvMaps& tm, const ConvParams& p, const ConvFused& fz, cudaStream_t st) {
  constexpr int STAGES = conv_stages(BLOCK_N, TERMS);
  constexpr int smem = conv_smem(BLOCK_N, TERMS);
  static bool attr_set = false;
  if (!attr_set) {
    RLDM_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BLOCK_N, STAGES, TERMS, NSPLIT>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  dim3 grid((p.M_total + kBlockM - 1) / kBlockM, p.Cout / BLOCK_N, NSPLIT);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(conv_threads(NSPLIT, BLOCK_N));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl_enabled_small()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (NSPLIT > 1 || p.clm > 1) {   // the K splits of one tile (DSMEM reduction in the epilogue) x the M tiles of one image
    attr[na].id = cudaLaunchAttributeClusterDimension;    // (emitting launches: moments complete inside the cluster)
    attr[na].val.clusterDim.x = p.clm;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = NSPLIT;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  RLDM_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<BLOCK_N, STAGES, TERMS, NSPLIT>, tm, p, fz));
  return 0;
}

template <int BLOCK_N, int TERMS>
static int launch_conv(const ConvMaps& tm, const ConvParams& p, const ConvFused& fz, int split, cudaStream_t st) {
  switch (split) {
    case 1: return launch_conv_n<BLOCK_N, TERMS, 1>(tm, p, fz, st);
    case 2: return launch_conv_n<BLOCK_N, TERMS, 2>(tm, p, fz, st);
    case 4: return launch_conv_n<BLOCK_N, TERMS, 4>(tm, p, fz, st);
    default: return launch_conv_n<BLOCK_N, TERMS, 8>(tm, p, fz, st);
  }
}

template <int BLOCK_N, int TERMS, int MT>
static int launch_conv_persistent(const ConvMaps& tm, const ConvParams& p, int n_ctas, cudaStream_t st) {
  constexpr int STAGES = pers_stages(BLOCK_N, TERMS, MT);
  static_assert(STAGES >= 2, "persistent conv: at least two pipeline stages");
  constexpr int smem = STAGES * pers_stage_bytes(BLOCK_N, TERMS, MT) + pers_fixed(BLOCK_N);
  static_assert(smem <= kSmemLimit, "persistent conv: shared memory budget exceeded");
  static bool attr_set = false;
  if (!attr_set) {
    RLDM_CUDA(cudaFuncSetAttribute(conv_tc_persistent_kernel<BLOCK_N, STAGES, TERMS, MT>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  if (env().wt_pdl_all) RLDM_CUDA(launch_pdl_small(conv_tc_persistent_kernel<BLOCK_N, STAGES, TERMS, MT>, dim3(n_ctas), dim3(192), smem, st, tm, p));
  else RLDM_CUDA(launch_pdl(conv_tc_persistent_kernel<BLOCK_N, STAGES, TERMS, MT>, dim3(n_ctas), dim3(192), smem, st, tm, p));
  return 0;
}

template <int TERMS>
static int launch_conv_wt(const ConvMaps& tm, const ConvParams& p, int n_ctas, cudaStream_t st) {
  constexpr int smem = wt_stages(TERMS) * (x_parts(TERMS) * 256 + w_parts(TERMS) * 128) * 64 * 2 + 256 + 1024;
  static_assert(smem <= kSmemLimit, "conv_tc_wt: shared memory budget exceeded");
  static bool attr_set = false;
  if (!attr_set) {
    RLDM_CUDA(cudaFuncSetAttribute(conv_tc_wt_kernel<TERMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  if (env().wt_pdl_all) RLDM_CUDA(launch_pdl_small(conv_tc_wt_kernel<TERMS>, dim3(n_ctas), dim3(kWtThreads), smem, st, tm, p));
  else RLDM_CUDA(launch_pdl(conv_tc_wt_kernel<TERMS>, dim3(n_ctas), dim3(kWtThreads), smem, st, tm, p));
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
  // pass: setup, TMEM allocation and descriptor prefetch overlap it (RLDM_WT_PDL=0 switches it off).
  const int units = (p.M_total / 256) * (p.Cout / 128);
  if (env().wt_pdl && (units <= n_ctas || env().wt_pdl_all))
    RLDM_CUDA(launch_pdl_small(conv_tc_wt_kernel<TERMS, true>, dim3(n_ctas), dim3(kWtThreads), smem, st, tm, p));
  else
    RLDM_CUDA(launch_pdl(conv_tc_wt_kernel<TERMS, true>, dim3(n_ctas), dim3(kWtThreads), smem, st, tm, p));
  return 0;
}

#define RLDM_BY_TERMS(terms, CALL)            \
  switch (terms) {                            \
    case 3: { constexpr int T_ = 3; CALL; }   \
    case 2: { constexpr int T_ = 2; CALL; }   \
    default: { constexpr int T_ = 1; CALL; }  \
  }

}  // namespace rldm

using namespace rldm;

// profiling aid (not part of include/rldm.h): device buffer of 16 int64 that CTA 0 of the conv kernels fills with
// clock64() stamps: [0] entry, [1] prologue done, [2] first stage landed, [3] last MMA issued, [4] accumulator
// complete, [5] epilogue done.
extern "C" void rldm_debug_conv_timestamps(long long* dev_buf) { g_conv_dbg = dev_buf; }

// one output phase of a nearest-2x upsampling folded into the convolution (rldm_conv_tc_up2): a 2x2 convolution over the
// LOW-resolution operand with pads (pad_lo along W, pad_h along H) in {0, 1}, written to the pixels (2w + a, 2h + b)
struct ConvPhase { int pad_h, a, b; };

static int conv_tc_impl(const uint16_t* x, const uint16_t* x_lo, const uint16_t* wgt, const float* bias,
                        const float* temb, int temb_stride, const float* residual, float* out,
                        int B, int W, int H, int Cin, int Cout, int ks, int stride, int pad_lo,
                        int circular, int split_k, double* stats, const uint16_t* sc_x, const uint16_t* sc_x_lo,
                        const uint16_t* sc_wgt, int sc_cin, int terms, const rldm_conv_src* src, const rldm_conv_src* sc_src,
                        bool query_only, void* stream, const rldm_conv_emit* emit = nullptr, const ConvPhase* phase = nullptr) {
  // query_only: no launch; returns 0 when this layer would run on the small-layer kernel (the one that can produce its
  // own operand from `src` and emit the next GroupNorm's operand), 1 otherwise
  if (terms == 0) terms = x_lo ? 3 : 1;          // legacy entry points: the operand planes say it
  RLDM_CHECK(terms >= 1 && terms <= 3, "conv_tc: terms must be 1, 2 or 3 (got %d)", terms);
  RLDM_CHECK(terms != 3 || x_lo, "conv_tc: split-fp16 x3 needs the low-order activation plane");
  if (terms != 3) { x_lo = nullptr; sc_x_lo = nullptr; }
  RLDM_CHECK(ks == 1 || ks == 3 || (ks == 2 && phase), "conv_tc: ks must be 1 or 3 (got %d)", ks);
  RLDM_CHECK(!phase || (stride == 1 && !residual && !sc_x && !src && !sc_src && !emit), "conv_tc: a phase convolution is a plain stride-1 layer");
  RLDM_CHECK(!sc_x || (sc_wgt && sc_cin > 0 && sc_cin % 64 == 0 && stride == 1 && (terms != 3 || sc_x_lo)),
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
  EncodeTiledFn encode = query_only ? nullptr : get_encode();
  RLDM_CHECK(query_only || encode != nullptr, "conv_tc: cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  const EnvSwitches& sw = env();
  const int n_sms = sw.n_sms;
  auto allowed = [&](int mode) { return mode == 1 || (mode == 2 && !residual); };      // 1: on, 2: "nores", 0: off

  int BN = (Cout % 128 == 0) ? 128 : 64;
  // 1x1 projections with at most RLDM_SMALL_BN64 (default 128) 128 x 128 tiles run 64-wide tiles: they have 2-4 K steps
  // and no K split, so a CTA's time is its epilogue, and twice as many CTAs halve it (224.7 -> 226.4 images/s).  For the
  // 3x3 layers (RLDM_SMALL_BN64_ALL=1) the extra CTAs per cluster reduction cost more than they save (214 images/s).
  // (only while the 64-wide tiles still fit one wave of the small-layer kernel)
  if (BN == 128 && sw.small_bn64 > 0 && (ks == 1 || sw.small_bn64_all)) {
    const int tiles128 = ((B * (W / stride) * (H / stride) + kBlockM - 1) / kBlockM) * (Cout / 128);
    if (tiles128 <= sw.small_bn64 && 2 * tiles128 <= n_sms) BN = 64;
  }
  // M tile = 128 output pixels = ncols whole columns x nb images
  const int pix = Wo * Ho;
  RLDM_CHECK(pix % 128 == 0 || 128 % pix == 0, "conv_tc: Wo*Ho=%d must divide or be a multiple of 128", pix);
  const int nb = pix >= 128 ? 1 : 128 / pix;
  const int ncols = pix >= 128 ? 128 / Ho : Wo;
  RLDM_CHECK(ncols * stride <= 256, "conv_tc: tile of %d columns exceeds the TMA box limit", ncols);
  const int xp = terms == 3 ? 2 : 1, wp = terms >= 2 ? 2 : 1;
  cudaStream_t st = as_stream(stream);

  ConvParams p;
  p.bias = bias; p.temb = temb; p.residual = residual; p.out = out;
  p.temb_stride = temb_stride;
  p.M_total = B * pix;
  p.Wo = Wo; p.Ho = Ho; p.W_in = W;
  p.pix_per_img = pix;
  p.Cout = Cout;
  p.ks = ks; p.stride = stride; p.pad_lo = pad_lo; p.circular = circular;
  p.pad_h = phase ? phase->pad_h : pad_lo; p.out_up = phase ? 2 : 1; p.out_a = phase ? phase->a : 0; p.out_b = phase ? phase->b : 0;
  p.main_iters = (Cin / kBlockK) * ks * ks;
  p.total_iters = p.main_iters + (sc_x ? sc_cin / kBlockK : 0);
  p.units = 0; p.a_part_bytes = 0; p.nb_stages = 0;
  p.dbg = g_conv_dbg;
  p.stats = stats;
  p.stats_G = Cout / 2;       // channel pairs per image
  RLDM_CHECK(pix >= 64 || !stats, "conv_tc: fused statistics need >= 64 pixels per image");
  p.emit_out = nullptr; p.emit_gamma = nullptr; p.emit_beta = nullptr; p.emit_eps = 0.f;
  p.emit_cpg = 4; p.emit_silu = 0; p.emit_circular = 1; p.clm = 1; p.emit_inv_n = 0.f;
  if (emit) {
    // the cluster (clm M tiles of one image x the K slices) must hold whole images: <= 8 CTAs
    const int cpg = emit->G > 0 ? Cout / emit->G : 0;
    RLDM_CHECK(emit->out && emit->gamma && emit->beta && emit->G > 0 && Cout % emit->G == 0 && cpg % 4 == 0 && cpg <= 64 &&
               BN % cpg == 0 && (cpg & (cpg - 1)) == 0,
               "conv_tc: emit needs out/gamma/beta and a power-of-two group size that is a multiple of 4 and divides the %d-wide tile (Cout=%d G=%d)",
               BN, Cout, emit->G);
    RLDM_CHECK(pix >= 64 && pix <= 8 * kBlockM && (pix & (pix - 1)) == 0,
               "conv_tc: emit needs a power of two of 64..1024 output pixels per image (got %d)", pix);
    p.emit_out = reinterpret_cast<__half*>(emit->out); p.emit_gamma = emit->gamma; p.emit_beta = emit->beta;
    p.emit_eps = emit->eps; p.emit_cpg = cpg; p.emit_silu = emit->silu; p.emit_circular = emit->circular;
    p.clm = pix > kBlockM ? pix / kBlockM : 1;
    p.emit_inv_n = static_cast<float>(1.0 / (static_cast<double>(pix) * cpg));
  }

  // ---- role-swapped kernel with pixel windows: 3x3, stride 1, symmetric pad, more 128x128 tiles than SMs, whole
  //      256-pixel units, room for two pixel windows plus >= 3 weight entries ----
  {
    const int tiles_h = (B * pix / kBlockM) * (Cout / BN);
    const size_t a_stage = static_cast<size_t>(xp) * (2 * kBlockM + 2 * Ho) * 128;
    const size_t fixed_wt = 2 * a_stage + (4 + 2 * 8 + 4) * 8 + 16 + 1024;
    int nws = fixed_wt + 2 * 16384 <= static_cast<size_t>(kSmemLimit) ? static_cast<int>((kSmemLimit - fixed_wt) / 16384) : 0;
    if (nws > 8) nws = 8;
    const bool halo_wt = ks == 3 && stride == 1 && pad_lo == 1 && Ho >= 8 && pix % (2 * kBlockM) == 0 && !sc_x &&
                         split_k <= 1 && tiles_h > n_sms && sw.conv_persistent && BN == 128 && nws >= 3 &&
                         allowed(sw.conv_wt) && allowed(sw.conv_wt_halo);
    if (halo_wt) {
      if (query_only) return 1;
      RLDM_CHECK(!src && !sc_src && !emit, "conv_tc: in-kernel operand production / emission is a feature of the small-layer kernel only");
      ConvMaps tmh;
      const int cols = 2 * (kBlockM / Ho);
      for (int part = 0; part < xp; ++part) {
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
      if (xp == 1) tmh.alo = tmh.a;
      {
        cuuint64_t gdim[2] = {(cuuint64_t)Cin, (cuuint64_t)wp * 9 * Cout};
        cuuint64_t gstr[1] = {(cuuint64_t)Cin * 2};
        cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)BN};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&tmh.b, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<uint16_t*>(wgt), gdim, gstr,
                            box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        RLDM_CHECK(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled(B) failed: %d", (int)r);
      }
      tmh.a2 = tmh.a; tmh.a2lo = tmh.alo; tmh.b2 = tmh.b;
      p.units = (Cin / kBlockK) * 3;
      p.a_part_bytes = static_cast<int>(a_stage / xp);
      p.nb_stages = nws;
      const int units = tiles_h / 2;
      const int ctas = units < n_sms ? units : n_sms;
      const size_t smem = fixed_wt + static_cast<size_t>(nws) * 16384;
      RLDM_BY_TERMS(terms, return launch_conv_wt_halo<T_>(tmh, p, ctas, smem, st));
    }
  }
  ConvMaps tm;
  // activation maps: (C, H, W+2, B) fp16, box = (64 channels, Ho*stride rows, ncols*stride columns, nb images)
  auto encode_act = [&](CUtensorMap* m, const uint16_t* ptr, int C, int s) -> CUresult {
    cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)H, (cuuint64_t)(W + 2), (cuuint64_t)B};
    cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)H * C * 2, (cuuint64_t)(W + 2) * H * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)(Ho * s), (cuuint32_t)(ncols * s), (cuuint32_t)nb};
    cuuint32_t estr[4] = {1, (cuuint32_t)s, (cuuint32_t)s, 1};
    return encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<uint16_t*>(ptr), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  // weight maps: [planes*taps*Cout][C] fp16, box = (64 channels, BN rows)
  auto encode_wgt = [&](CUtensorMap* m, const uint16_t* ptr, int C, int rows) -> CUresult {
    cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)BN};
    cuuint32_t estr[2] = {1, 1};
    return encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<uint16_t*>(ptr), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  auto build_maps = [&]() -> int {
    CUresult r = encode_act(&tm.a, x, Cin, stride);
    RLDM_CHECK(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled(A) failed: %d", (int)r);
    tm.alo = tm.a;
    if (xp == 2) {
      r = encode_act(&tm.alo, x_lo, Cin, stride);
      RLDM_CHECK(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled(A lo) failed: %d", (int)r);
    }
    r = encode_wgt(&tm.b, wgt, Cin, wp * ks * ks * Cout);
    RLDM_CHECK(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled(B) failed: %d", (int)r);
    tm.a2 = tm.a; tm.a2lo = tm.alo; tm.b2 = tm.b;
    if (sc_x) {
      r = encode_act(&tm.a2, sc_x, sc_cin, 1);
      RLDM_CHECK(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled(shortcut A) failed: %d", (int)r);
      tm.a2lo = tm.a2;
      if (xp == 2) {
        r = encode_act(&tm.a2lo, sc_x_lo, sc_cin, 1);
        RLDM_CHECK(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled(shortcut A lo) failed: %d", (int)r);
      }
      r = encode_wgt(&tm.b2, sc_wgt, sc_cin, wp * Cout);
      RLDM_CHECK(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled(shortcut B) failed: %d", (int)r);
    }
    return 0;
  };
  const int tiles = ((p.M_total + kBlockM - 1) / kBlockM) * (Cout / BN);
  int split = split_k;
  if (split <= 0) {  // auto: fill the SMs when the tile grid is small (clusters of <= 8 CTAs along K)
    split = 1;
    const int cap = 2 * conv_smem(BN, terms) <= kSmemLimit ? 296 : 160;   // resident CTAs: one or two per SM
    while (tiles * split * 2 <= cap && p.total_iters / (split * 2) >= 4 && split < 8) split *= 2;
  }
  RLDM_CHECK(split == 1 || split == 2 || split == 4 || split == 8, "conv_tc: split_k must be 1, 2, 4 or 8 (got %d)", split);
  while (split > p.total_iters) split /= 2;
  // emitting launch: (M tiles of the image) x (K slices) <= 8 CTAs per cluster (RLDM_EMIT_MAXCL: a smaller limit for
  // clusters that span M tiles -- sixteen 8-CTA clusters do not all fit the GPCs at once)
  while (p.clm * split > (p.clm > 1 ? sw.emit_maxcl : 8) && split > 1) split /= 2;
  // more tiles than SMs and no K split: persistent CTAs with a double-buffered TMEM accumulator
  if (split == 1 && tiles > n_sms && sw.conv_persistent) {
    if (query_only) return 1;
    RLDM_CHECK(!src && !sc_src && !emit, "conv_tc: in-kernel operand production / emission is a feature of the small-layer kernel only");
    // Cout tiles of 128 and whole 256-pixel units inside one image: roles swapped (weights = M side, N = 256 pixels)
    if (BN == 128 && p.M_total % 256 == 0 && pix % 256 == 0 && allowed(sw.conv_wt)) {
      if (int rc = build_maps()) return rc;
      const int units_wt = (p.M_total / 256) * (Cout / 128);
      RLDM_BY_TERMS(terms, return launch_conv_wt<T_>(tm, p, units_wt < n_sms ? units_wt : n_sms, st));
    }
    RLDM_CHECK(!phase, "conv_tc: phase convolutions run on the role-swapped kernel only (rldm_conv_tc_up2_ok)");
    // two M tiles per unit share the weight tiles when the tile count allows it (measured: -5..-20 % on layers without
    // a residual operand and on 64-channel layers; 128-wide layers WITH a residual are paced by the drain of two
    // tiles, not the K loop: they keep one tile per unit)
    const int tiles_m = (p.M_total + kBlockM - 1) / kBlockM;
    const bool mt2 = tiles_m % 2 == 0 && p.M_total % kBlockM == 0 && (residual == nullptr || BN == 64 || sw.conv_mt2_res) && !sw.conv_mt1;
    const int units = mt2 ? tiles / 2 : tiles;
    const int ctas = units < n_sms ? units : n_sms;
    if (int rc = build_maps()) return rc;
    if (mt2) {
      if (BN == 128) { RLDM_BY_TERMS(terms, return (launch_conv_persistent<128, T_, 2>(tm, p, ctas, st))); }
      RLDM_BY_TERMS(terms, return (launch_conv_persistent<64, T_, 2>(tm, p, ctas, st)));
    }
    if (BN == 128) { RLDM_BY_TERMS(terms, return (launch_conv_persistent<128, T_, 1>(tm, p, ctas, st))); }
    RLDM_BY_TERMS(terms, return (launch_conv_persistent<64, T_, 1>(tm, p, ctas, st)));
  }
  if (query_only) return (Cin <= 512 && sc_cin <= 512) ? 0 : 1;
  RLDM_CHECK(!phase, "conv_tc: phase convolutions run on the role-swapped kernel only (rldm_conv_tc_up2_ok)");
  ConvFused fz;
  memset(&fz, 0, sizeof(fz));
  auto fill = [&](PrepArgs& a, const rldm_conv_src* s, const uint16_t* o_hi, const uint16_t* o_lo, int C, int Wop, int Hop) -> int {
    RLDM_CHECK(s->x0 != nullptr && s->c0 % 8 == 0 && s->c1 % 8 == 0 && s->c0 + s->c1 == C && (s->x1 != nullptr || s->c1 == 0),
               "conv_tc: source channels %d + %d do not make the operand's %d", s->c0, s->c1, C);
    RLDM_CHECK(s->up == 1 || s->up == 2, "conv_tc: source up must be 1 or 2");
    RLDM_CHECK(Wop % s->up == 0 && Hop % s->up == 0, "conv_tc: operand grid not divisible by the upsampling factor");
    RLDM_CHECK(!s->pairs0 || (s->gamma && s->beta && s->G > 0 && C % s->G == 0 && (C / s->G) % 2 == 0 && (s->c1 == 0 || s->pairs1)),
               "conv_tc: GroupNorm of the source needs gamma/beta/G, an even group size and moments for both concat halves");
    RLDM_CHECK(C <= kFusedTabFloats / 2, "conv_tc: in-kernel operand production supports up to %d channels", kFusedTabFloats / 2);
    a.x0 = s->x0; a.x1 = s->x1; a.sums = nullptr; a.pairs0 = s->pairs0; a.pairs1 = s->pairs1; a.gamma = s->gamma; a.beta = s->beta;
    a.out = reinterpret_cast<__half*>(const_cast<uint16_t*>(o_hi)); a.out_lo = reinterpret_cast<__half*>(const_cast<uint16_t*>(o_lo));
    a.raw = nullptr; a.raw_lo = nullptr;
    a.eps = s->eps; a.c0 = s->c0; a.c1 = s->c1; a.G = s->G; a.silu = s->silu; a.up = s->up; a.circular = s->circular;
    a.W = Wop / s->up; a.H = Hop / s->up; a.pix_per_block = 0;
    return 0;
  };
  if (src) {
    if (int rc = fill(fz.main, src, x, x_lo, Cin, W, H)) return rc;
    fz.main_on = 1;
  }
  if (sc_src) {
    RLDM_CHECK(sc_x != nullptr, "conv_tc: a shortcut source needs the shortcut operand buffers and weights");
    if (int rc = fill(fz.sc, sc_src, sc_x, sc_x_lo, sc_cin, W, H)) return rc;
    fz.sc_on = 1;
  }
  if (int rc = build_maps()) return rc;
  if (BN == 128) { RLDM_BY_TERMS(terms, return (launch_conv<128, T_>(tm, p, fz, split, st))); }
  RLDM_BY_TERMS(terms, return (launch_conv<64, T_>(tm, p, fz, split, st)));
}

extern "C" int rldm_conv_tc(const uint16_t* x, const uint16_t* x_lo, const uint16_t* wgt, const float* bias,
                            const float* temb, int temb_stride, const float* residual, float* out,
                            int B, int W, int H, int Cin, int Cout, int ks, int stride, int pad_lo,
                            int circular, int split_k, double* stats, void* stream) {
  return conv_tc_impl(x, x_lo, wgt, bias, temb, temb_stride, residual, out, B, W, H, Cin, Cout, ks, stride, pad_lo,
                      circular, split_k, stats, nullptr, nullptr, nullptr, 0, 0, nullptr, nullptr, false, stream);
}

extern "C" int rldm_conv_tc_shortcut(const uint16_t* x, const uint16_t* x_lo, const uint16_t* wgt, const float* bias,
                                     const float* temb, int temb_stride, const float* residual, float* out,
                                     int B, int W, int H, int Cin, int Cout, int ks, int stride, int pad_lo,
                                     int circular, int split_k, double* stats, const uint16_t* sc_x,
                                     const uint16_t* sc_x_lo, const uint16_t* sc_wgt, int sc_cin, void* stream) {
  return conv_tc_impl(x, x_lo, wgt, bias, temb, temb_stride, residual, out, B, W, H, Cin, Cout, ks, stride, pad_lo,
                      circular, split_k, stats, sc_x, sc_x_lo, sc_wgt, sc_cin, 0, nullptr, nullptr, false, stream);
}

extern "C" int rldm_conv_tc_ex(const uint16_t* x, const uint16_t* x_lo, const uint16_t* wgt, const float* bias,
                               const float* temb, int temb_stride, const float* residual, float* out,
                               int B, int W, int H, int Cin, int Cout, int ks, int stride, int pad_lo,
                               int circular, int split_k, double* stats, const uint16_t* sc_x,
                               const uint16_t* sc_x_lo, const uint16_t* sc_wgt, int sc_cin, int terms, void* stream) {
  return conv_tc_impl(x, x_lo, wgt, bias, temb, temb_stride, residual, out, B, W, H, Cin, Cout, ks, stride, pad_lo,
                      circular, split_k, stats, sc_x, sc_x_lo, sc_wgt, sc_cin, terms, nullptr, nullptr, false, stream);
}

extern "C" int rldm_conv_tc_fused(const rldm_conv_src* src, const rldm_conv_src* sc_src, const uint16_t* x, const uint16_t* x_lo,
                                  const uint16_t* wgt, const float* bias, const float* temb, int temb_stride,
                                  const float* residual, float* out, int B, int W, int H, int Cin, int Cout, int ks, int stride,
                                  int pad_lo, int circular, int split_k, double* stats, const uint16_t* sc_x,
                                  const uint16_t* sc_x_lo, const uint16_t* sc_wgt, int sc_cin, int terms, void* stream) {
  return conv_tc_impl(x, x_lo, wgt, bias, temb, temb_stride, residual, out, B, W, H, Cin, Cout, ks, stride, pad_lo,
                      circular, split_k, stats, sc_x, sc_x_lo, sc_wgt, sc_cin, terms, src, sc_src, false, stream);
}

extern "C" int rldm_conv_tc_up2_ok(int B, int W, int H, int Cin, int Cout) {
  const EnvSwitches& sw = env();
  const int pix = W * H;
  if (Cin % 64 != 0 || Cout % 128 != 0 || pix % 256 != 0 || (H & (H - 1)) != 0 || H > 128) return 0;
  if (!(sw.conv_wt == 1 || sw.conv_wt == 2) || !sw.conv_persistent) return 0;
  return (B * pix / kBlockM) * (Cout / 128) > sw.n_sms ? 1 : 0;
}

extern "C" int rldm_conv_tc_up2(const uint16_t* x, const uint16_t* x_lo, const uint16_t* wgt, const float* bias, float* out,
                                int B, int W, int H, int Cin, int Cout, int circular, double* stats, int terms, void* stream) {
  RLDM_CHECK(rldm_conv_tc_up2_ok(B, W, H, Cin, Cout), "conv_tc_up2: layer B=%d %dx%d %d->%d does not run on the role-swapped kernel",
             B, W, H, Cin, Cout);
  if (terms == 0) terms = x_lo ? 3 : 1;
  const size_t phase_elems = static_cast<size_t>(terms >= 2 ? 2 : 1) * 4 * Cout * Cin;
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) {
      const ConvPhase ph{1 - b, a, b};
      if (int rc = conv_tc_impl(x, x_lo, wgt + (2 * a + b) * phase_elems, bias, nullptr, 0, nullptr, out, B, W, H, Cin, Cout, 2, 1,
                                1 - a, circular, 1, stats, nullptr, nullptr, nullptr, 0, terms, nullptr, nullptr, false, stream,
                                nullptr, &ph))
        return rc;
    }
  return 0;
}

extern "C" int rldm_conv_tc_emit(const rldm_conv_emit* emit, const uint16_t* x, const uint16_t* x_lo, const uint16_t* wgt,
                                 const float* bias, const float* temb, int temb_stride, const float* residual, float* out,
                                 int B, int W, int H, int Cin, int Cout, int ks, int stride, int pad_lo, int circular,
                                 int split_k, double* stats, const uint16_t* sc_x, const uint16_t* sc_x_lo,
                                 const uint16_t* sc_wgt, int sc_cin, int terms, void* stream) {
  return conv_tc_impl(x, x_lo, wgt, bias, temb, temb_stride, residual, out, B, W, H, Cin, Cout, ks, stride, pad_lo,
                      circular, split_k, stats, sc_x, sc_x_lo, sc_wgt, sc_cin, terms, nullptr, nullptr, false, stream, emit);
}

extern "C" int rldm_conv_tc_emittable(int B, int W, int H, int Cin, int Cout, int ks, int stride, int pad_lo, int sc_cin,
                                      int has_residual, int G) {
  if (G <= 0 || Cout % G != 0) return 0;
  const int cpg = Cout / G, BN = (Cout % 128 == 0) ? 128 : 64;
  const int pix = (W / stride) * (H / stride);
  if (cpg % 4 != 0 || cpg > 64 || (cpg & (cpg - 1)) != 0 || BN % cpg != 0 || pix < 64 || pix > 8 * 128 || (pix & (pix - 1)) != 0) return 0;
  // RLDM_EMIT_MAXCLM (default 1: images of at most 128 pixels, i.e. clusters along K only -- level 3 of the C3 UNet;
  // measured 230.3 -> 232.7 images/s; 2 -> 218.7: clusters that also span M tiles are placed too late): largest number
  // of M tiles per image (= CTAs along M of the cluster) that still emits
  if ((pix > 128 ? pix / 128 : 1) > env().emit_maxclm) return 0;
  return rldm_conv_tc_fusable(B, W, H, Cin, Cout, ks, stride, pad_lo, sc_cin, has_residual);
}

extern "C" int rldm_conv_tc_fusable(int B, int W, int H, int Cin, int Cout, int ks, int stride, int pad_lo, int sc_cin,
                                    int has_residual) {
  // dummy 16 B-aligned, non-NULL pointers: the query path dereferences nothing
  static const uint16_t* kDummy = reinterpret_cast<const uint16_t*>(static_cast<uintptr_t>(256));
  const int rc = conv_tc_impl(kDummy, nullptr, kDummy, nullptr, nullptr, 0, has_residual ? reinterpret_cast<const float*>(kDummy) : nullptr,
                              reinterpret_cast<float*>(const_cast<uint16_t*>(kDummy)), B, W, H, Cin, Cout, ks, stride, pad_lo, 1, 0,
                              nullptr, sc_cin ? kDummy : nullptr, nullptr, sc_cin ? kDummy : nullptr, sc_cin, 1, nullptr, nullptr, true,
                              nullptr);
  return rc == 0 ? 1 : 0;
}
