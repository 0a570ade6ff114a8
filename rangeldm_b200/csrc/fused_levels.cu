// fused_levels.cu -- one persistent kernel for a whole run of small layers of the UNet.
//
// Levels 1..3 of the C3 UNet (128x8, 64x4 and 32x2 latents at batch 8) are a chain of ~140 tiny, strictly dependent
// ops per forward: GroupNorm-apply passes over a few hundred KB, 3x3 / 1x1 convolutions of 2-9 K steps per CTA and
// 64 / 256-token attention.  As separate graph nodes each costs 5-10 us of launch, prologue (barrier init, TMEM
// allocation, descriptor fetch), first-load latency and drain, although its arithmetic is worth 1-2 us
// (profiles/ablate_r2_start.json: 0.97 ms of the 1.87 ms forward).  Here a run of such ops becomes ONE launch of
// 148 co-resident CTAs that walk a phase list; the kernel boundary between two dependent ops is replaced by a grid-wide
// barrier (one atomic arrive + one polled generation word in L2, ~1 us), and barriers, TMEM and the TMA ring are set
// up once per run instead of once per op.
//
// Phases (a program op becomes one or two):
//   PREP       GroupNorm-apply + SiLU + concat + nearest-2x + circular halo + fp16 split (body shared with
//              rldm_prep: blocks.cuh), virtual blocks walked in a grid-stride loop;
//   CONV_MAIN  implicit-GEMM K loop of one (128-pixel x BLOCK_N, K slice) item per CTA on tcgen05 (TMA ring, one MMA
//              issuer thread, fp32 accumulator in TMEM, same operand layouts and tensor maps as rldm_conv_tc); the partial
//              tile goes to an L2-resident workspace;
//   CONV_FIN   sums the K slices in a fixed order (deterministic), adds bias / time embedding / residual, writes the
//              fp32 output and accumulates the GroupNorm channel-pair moments (8 rows x 128 channels per CTA pass);
//   ATTN       the mma.sync attention body (blocks.cuh), two 128-thread virtual blocks per CTA.
// Everything one phase writes and a later phase reads travels through L2: reads of such tensors use ld.global.cg
// (L1 is not coherent across SMs inside one launch), TMA reads are ordered behind the barrier by fence.proxy.async.
// Replaces, for these layers, the same reference arithmetic as the stand-alone kernels: ResnetBlock2D / Attention /
// Downsample2D / Upsample2D of diffusers' UNet2DModel over `ldm/utils.py:40-58` convolutions.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "blocks.cuh"

namespace rldm {

constexpr int kFThreads = 256;
constexpr int kFStages = 3;
constexpr int kFStageBytes = 65536;       // [X_hi 16K][X_lo 16K][W_hi 16K][W_lo 16K]; smaller layouts use a prefix
constexpr int kFABytes = 128 * 64 * 2;    // one 128-pixel x 64-channel operand part
constexpr int kFSmem = kFStages * kFStageBytes + 1024 /*align*/ + 128 /*barriers*/ + 8 * 32 * 4 * 4 /*moment scratch*/;

enum { F_PREP = 0, F_CONV_MAIN = 1, F_CONV_FIN = 2, F_ATTN = 3 };

struct FConv {
  float* ws;
  const float* bias; const float* temb; const float* residual; float* out; double* stats;
  int map;            // first of this convolution's six tensor maps (a, alo, b, a2, a2lo, b2)
  int M_total, Wo, Ho, pix_per_img, Cout, ks, stride, pad_lo;
  int total_iters, main_iters;
  int BN, terms, tiles_m, tiles_n, ksplit;
  int temb_stride, stats_G;
};
struct FAttn {
  const float* qkv; __half* out; __half* out_lo;
  int B, N, C, H;
};
struct FPrep {
  PrepArgs a;
  int vgx, vgy;
};
struct FPhase {
  int kind;
  int pad_;
  union {
    FPrep prep;
    FConv conv;
    FAttn attn;
  };
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// elect.sync that also reports the elected lane (uniform across the warp)
__device__ __forceinline__ bool elect_one_leader(int& leader) {
  uint32_t pred, lid;
  asm volatile(
      "{\n .reg .pred p;\n elect.sync %1|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(pred), "=r"(lid));
  leader = static_cast<int>(lid);
  return pred != 0;
}

// Grid-wide barrier of `nblocks` co-resident CTAs.  bar[0] = arrival count (reset by the last arriver), bar[1] =
// generation (only ever incremented, so the pair is consistent across launches and CUDA-graph replays).  A lost
// arrival traps after ~2 s instead of hanging the GPU.
// The arrival is ONE acq_rel atomic (its release half publishes the CTA's writes, ordered before it by bar.sync), the
// wait polls with RELAXED loads (an acquire load per poll would invalidate L1 on every iteration) and takes a single
// acquire fence once the generation has moved.  MODE 0 (RLDM_FUSE_BAR=0, for comparison): seq_cst __threadfence()
// around a relaxed atomic and acquire polling.
template <int MODE>
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned nblocks, unsigned& gen) {
  fence_proxy_async_all();            // this thread's generic writes -> later TMA (async proxy) reads of other CTAs
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned old;
    if (MODE == 0) {
      __threadfence();
      old = atomicAdd(&bar[0], 1u);
    } else {
      asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(bar) : "memory");
    }
    if (old == nblocks - 1) {
      if (MODE == 0) {
        bar[0] = 0u;
        __threadfence();
        atomicAdd(&bar[1], 1u);
      } else {
        asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(bar), "r"(0u) : "memory");
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar + 1) : "memory");
      }
    } else {
      const long long t0 = clock64();
      for (;;) {
        unsigned v;
        if (MODE == 0) v = ld_acquire_u32(&bar[1]);
        else asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar + 1) : "memory");
        if (v != gen) break;
        if (clock64() - t0 > 4000000000ll) {
          printf("rldm: grid barrier timeout block %d gen %u\n", blockIdx.x, gen);
          __trap();
        }
      }
    }
    if (MODE == 0) __threadfence();
    else asm volatile("fence.acq_rel.gpu;" ::: "memory");
  }
  ++gen;
  __syncthreads();
}

struct FRing {
  uint8_t* smem;
  uint64_t* full_bar; uint64_t* empty_bar; uint64_t* tmem_full; uint64_t* tmem_empty;
  uint32_t tmem;
};

// K loop of the items of one convolution; the partial 128 x BN tiles go to c.ws[item].
// g_ring: stages this warp role has walked so far (ring position / parity), n_acc: items it has finished (accumulator
// hand-over parity); both persist across phases.
__device__ __forceinline__ void conv_main_phase(const FConv& c, const CUtensorMap* tm, const FRing& R, int warp, int lane,
                                                uint32_t& g_ring, uint32_t& n_acc) {
  const int items = c.tiles_m * c.tiles_n * c.ksplit;
  const int XP = c.terms == 3 ? 2 : 1, WP = c.terms >= 2 ? 2 : 1;
  const int BN = c.BN;
  const int b_bytes = BN * 128;
  const uint32_t stage_bytes = XP * kFABytes + WP * b_bytes;
  const uint32_t b_off = XP * kFABytes;
  const int taps = c.ks * c.ks;
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) fence_proxy_async_all();
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const int ksi = item % c.ksplit, tile = item / c.ksplit;
      const int tn = tile % c.tiles_n, tmi = tile / c.tiles_n;
      const int m0 = tmi * 128, n0 = tn * BN;
      const int it0 = static_cast<int>(static_cast<long long>(ksi) * c.total_iters / c.ksplit);
      const int it1 = static_cast<int>(static_cast<long long>(ksi + 1) * c.total_iters / c.ksplit);
      const int q0 = m0 / c.Ho;                     // global column index of the tile's first column
      const int b0 = q0 / c.Wo, wo0 = q0 - b0 * c.Wo;
      for (int it = it0; it < it1; ++it, ++g_ring) {
        const uint32_t s = g_ring % kFStages, ph = (g_ring / kFStages) & 1;
        mbar_wait(&R.empty_bar[s], ph ^ 1);
        if (lane == 0) {
          const bool main = it < c.main_iters;
          const int chunk = main ? it / taps : it - c.main_iters;
          const int tap = it - chunk * taps;
          const int ti = tap / c.ks, tj = tap - ti * c.ks;
          const uint32_t dst = smem_u32(R.smem + s * kFStageBytes);
          mbar_arrive_expect_tx(&R.full_bar[s], stage_bytes);
          if (main) {
            tma_load_2d(dst + b_off, &tm[2], &R.full_bar[s], chunk * 64, tap * c.Cout + n0);
            if (WP > 1) tma_load_2d(dst + b_off + b_bytes, &tm[2], &R.full_bar[s], chunk * 64, (taps + tap) * c.Cout + n0);
          } else {
            tma_load_2d(dst + b_off, &tm[5], &R.full_bar[s], chunk * 64, n0);
            if (WP > 1) tma_load_2d(dst + b_off + b_bytes, &tm[5], &R.full_bar[s], chunk * 64, c.Cout + n0);
          }
          // shortcut K steps read the centre tap of the second tensor (1x1, stride 1, same grid as the output)
          const int h_in = main ? tj - c.pad_lo : 0;
          const int w_in = main ? c.stride * wo0 + ti - c.pad_lo + 1 : wo0 + 1;
          tma_load_4d(dst, main ? &tm[0] : &tm[3], &R.full_bar[s], chunk * 64, h_in, w_in, b0);
          if (XP > 1) tma_load_4d(dst + kFABytes, main ? &tm[1] : &tm[4], &R.full_bar[s], chunk * 64, h_in, w_in, b0);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    int leader = 0;
    if (elect_one_leader(leader)) {
      const uint32_t idesc = umma_idesc_f16(128, static_cast<uint32_t>(BN));
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int ksi = item % c.ksplit;
        const int it0 = static_cast<int>(static_cast<long long>(ksi) * c.total_iters / c.ksplit);
        const int it1 = static_cast<int>(static_cast<long long>(ksi + 1) * c.total_iters / c.ksplit);
        mbar_wait(R.tmem_empty, (n_acc & 1) ^ 1);          // the epilogue warps have drained the previous item
        tc_fence_after();
        for (int it = it0; it < it1; ++it, ++g_ring) {
          const uint32_t s = g_ring % kFStages, ph = (g_ring / kFStages) & 1;
          mbar_wait(&R.full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(R.smem + s * kFStageBytes);
          const uint64_t a_desc = umma_desc_sw128(a_addr), b_desc = umma_desc_sw128(a_addr + b_off);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(R.tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (it != it0 || k != 0) ? 1u : 0u);
          if (XP > 1) {
            const uint64_t al_desc = umma_desc_sw128(a_addr + kFABytes);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(R.tmem, al_desc + 2 * k, b_desc + 2 * k, idesc, 1u);   // X_lo * W_hi
          }
          if (WP > 1) {
            const uint64_t bl_desc = umma_desc_sw128(a_addr + b_off + b_bytes);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(R.tmem, a_desc + 2 * k, bl_desc + 2 * k, idesc, 1u);   // X_hi * W_lo
          }
          umma_commit(&R.empty_bar[s]);
        }
        umma_commit(R.tmem_full);
        ++n_acc;
      }
    }
    __syncwarp();
    g_ring = __shfl_sync(0xffffffffu, g_ring, leader);     // the counters live on across phases: every lane keeps them
    n_acc = __shfl_sync(0xffffffffu, n_acc, leader);
  } else if (warp < 6) {
    // ===================== accumulator -> workspace (thread = tile row) =====================
    const int q = warp & 3;                   // TMEM lane quadrant this warp may access
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      mbar_wait(R.tmem_full, n_acc & 1);
      tc_fence_after();
      float* dst = c.ws + (static_cast<size_t>(item) * 128 + q * 32 + lane) * BN;
      for (int nc = 0; nc < BN / 32; ++nc) {
        uint32_t r[32];
        tmem_ld_32x32(R.tmem + (static_cast<uint32_t>(q * 32) << 16) + nc * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(dst + nc * 32 + 4 * j) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(R.tmem_empty)) : "memory");
      ++n_acc;
    }
  }
}

// Reduce the K slices, finish the epilogue: unit = 8 output pixels (one per warp) x 128 channels (4 per lane).
__device__ __forceinline__ void conv_fin_phase(const FConv& c, float* red) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col_blocks = c.Cout >> 7;
  const int row_chunks = (c.M_total + 7) >> 3;
  const int units = row_chunks * col_blocks;
  const size_t slice = static_cast<size_t>(128) * c.BN;
  for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
    const int cb = unit % col_blocks, rc = unit / col_blocks;
    const int m = rc * 8 + warp;
    const int ch = cb * 128 + lane * 4;
    const bool live = m < c.M_total;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
      const int tmi = m >> 7, r = m & 127, tn = ch / c.BN, col = ch - tn * c.BN;
      const float* base = c.ws + ((static_cast<size_t>(tmi) * c.tiles_n + tn) * c.ksplit * 128 + r) * c.BN + col;
      if (c.bias) v = __ldg(reinterpret_cast<const float4*>(c.bias + ch));
      if (c.temb) {
        const float4 t4 = __ldg(reinterpret_cast<const float4*>(c.temb + static_cast<size_t>(m / c.pix_per_img) * c.temb_stride + ch));
        v.x += t4.x; v.y += t4.y; v.z += t4.z; v.w += t4.w;
      }
      if (c.residual) {
        const float4 t4 = __ldcg(reinterpret_cast<const float4*>(c.residual + static_cast<size_t>(m) * c.Cout + ch));
        v.x += t4.x; v.y += t4.y; v.z += t4.z; v.w += t4.w;
      }
#pragma unroll 4
      for (int s = 0; s < c.ksplit; ++s) {
        const float4 t4 = __ldcg(reinterpret_cast<const float4*>(base + s * slice));
        v.x += t4.x; v.y += t4.y; v.z += t4.z; v.w += t4.w;
      }
      *reinterpret_cast<float4*>(c.out + static_cast<size_t>(m) * c.Cout + ch) = v;
    }
    if (c.stats) {
      // channel-pair moments of the finished output: the unit's 8 rows lie in one image (pix_per_img % 8 == 0)
      float4 mo = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live) mo = make_float4(v.x + v.y, v.x * v.x + v.y * v.y, v.z + v.w, v.z * v.z + v.w * v.w);
      *reinterpret_cast<float4*>(red + (warp * 32 + lane) * 4) = mo;
      __syncthreads();
      if (threadIdx.x < 128) {
        const int l = threadIdx.x >> 2, comp = threadIdx.x & 3;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[(w * 32 + l) * 4 + comp];
        const int img = (rc * 8) / c.pix_per_img;
        const int pair = ((cb * 128 + l * 4) >> 1) + (comp >> 1);
        atomicAdd(c.stats + (static_cast<size_t>(img) * c.stats_G + pair) * 2 + (comp & 1), static_cast<double>(s));
      }
      __syncthreads();
    }
  }
}

__device__ __forceinline__ unsigned long long global_timer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// dbg (profiling aid, normally NULL): CTA 0 writes %globaltimer at [3p] phase start, [3p+1] its work done, [3p+2]
// grid barrier passed
template <int BAR_MODE>
__global__ void __launch_bounds__(kFThreads, 1)
fused_levels_kernel(const FPhase* __restrict__ phases, int n_phases, const CUtensorMap* __restrict__ maps, unsigned* bar,
                    unsigned long long* dbg) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kFStages * kFStageBytes);
  uint64_t* empty_bar = full_bar + kFStages;
  uint64_t* tmem_full = empty_bar + kFStages;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 1);
  float* red = reinterpret_cast<float*>(smem + kFStages * kFStageBytes + 128);
  __shared__ __align__(16) FPhase ph_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kFStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 4);              // one arrival per accumulator-draining warp
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<128>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  FRing R;
  R.smem = smem; R.full_bar = full_bar; R.empty_bar = empty_bar; R.tmem_full = tmem_full; R.tmem_empty = tmem_empty;
  R.tmem = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
  pdl_wait();                              // (no-op unless launched with the programmatic-serialization attribute)
  unsigned gen = 0;
  if (threadIdx.x == 0) gen = ld_acquire_u32(&bar[1]);   // cannot advance before every CTA has arrived once
  uint32_t g_ring = 0, n_acc = 0;

  for (int pi = 0; pi < n_phases; ++pi) {
    {   // phase descriptor -> shared memory
      const uint32_t* src = reinterpret_cast<const uint32_t*>(phases + pi);
      uint32_t* dst = reinterpret_cast<uint32_t*>(&ph_s);
      for (int i = threadIdx.x; i < static_cast<int>(sizeof(FPhase) / 4); i += kFThreads) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const bool stamp = dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
    if (stamp) dbg[3 * pi] = global_timer();
    const int kind = ph_s.kind;
    if (kind == F_PREP) {
      const int nvb = ph_s.prep.vgx * ph_s.prep.vgy;
      for (int vb = blockIdx.x; vb < nvb; vb += gridDim.x) {
        prep_block<true>(ph_s.prep.a, vb % ph_s.prep.vgx, vb / ph_s.prep.vgx, reinterpret_cast<float*>(smem));
        __syncthreads();                  // scale / shift table reusable by the next virtual block
      }
    } else if (kind == F_CONV_MAIN) {
      conv_main_phase(ph_s.conv, maps + ph_s.conv.map, R, warp, lane, g_ring, n_acc);
    } else if (kind == F_CONV_FIN) {
      conv_fin_phase(ph_s.conv, red);
    } else if (kind == F_ATTN) {
      const FAttn& a = ph_s.attn;
      const int nx = a.N / 64, ny = a.C / 8;
      const int nvb = nx * ny * a.B;
      const int half = threadIdx.x >> 7, t128 = threadIdx.x & 127;
      for (int vb0 = blockIdx.x * 2; vb0 < nvb; vb0 += gridDim.x * 2) {
        const int vb = vb0 + half;
        if (vb < nvb)
          attention_tc_block<true>(a.qkv, a.out, a.out_lo, a.N, a.C, a.H, vb % nx, (vb / nx) % ny, vb / (nx * ny),
                                   reinterpret_cast<__half*>(smem) + half * kAtSmemHalves, t128, 2 + half);
      }
    }
    if (dbg != nullptr) {
      __syncthreads();
      if (stamp) dbg[3 * pi + 1] = global_timer();
    }
    grid_barrier<BAR_MODE>(bar, gridDim.x, gen);
    if (stamp) dbg[3 * pi + 2] = global_timer();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<128>(R.tmem);
}

}  // namespace rldm

using namespace rldm;

// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn fused_get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

struct rldm_fused {
  unsigned long long* dbg = nullptr;
  std::vector<int> phase_kinds;
  FPhase* d_phases = nullptr;
  CUtensorMap* d_maps = nullptr;
  unsigned* d_bar = nullptr;
  int n_phases = 0;
  int n_ops = 0;
  int grid = 0;
};

// conv geometry shared by the support check and the compiler
struct FConvGeom {
  int B, W, H, Cin, Cout, ks, stride, pad_lo, sc_cin, terms, Wo, Ho, pix, BN, nb, ncols, tiles_m, tiles_n, total_iters;
};
static bool fused_conv_geom(const rldm_op& o, FConvGeom& g) {
  g.B = o.i[1]; g.W = o.i[2]; g.H = o.i[3]; g.Cin = o.i[4]; g.Cout = o.i[5]; g.ks = o.i[6]; g.stride = o.i[7];
  g.pad_lo = o.i[8]; g.sc_cin = o.i[11]; g.terms = o.i[12];
  if (o.i[10] != 0 || o.p[11] || o.p[17]) return false;             // explicit split_k / own operand production: stand-alone kernel
  if (g.terms < 1 || g.terms > 3) return false;
  if (g.terms == 3 && !o.p[6]) return false;
  if ((g.ks != 1 && g.ks != 3) || (g.stride != 1 && g.stride != 2)) return false;
  if (g.Cin % 64 || g.Cout % 128 || g.W % g.stride || g.H % g.stride) return false;
  if (o.p[8] && (g.sc_cin <= 0 || g.sc_cin % 64 || g.stride != 1)) return false;
  g.Wo = g.W / g.stride; g.Ho = g.H / g.stride;
  if (g.Ho < 1 || g.Ho > 128 || (g.Ho & (g.Ho - 1))) return false;
  g.pix = g.Wo * g.Ho;
  if (!(g.pix % 128 == 0 || 128 % g.pix == 0)) return false;
  if (g.pix % 8) return false;                                      // conv_fin: 8-row units never straddle images
  g.nb = g.pix >= 128 ? 1 : 128 / g.pix;
  g.ncols = g.pix >= 128 ? 128 / g.Ho : g.Wo;
  if (g.ncols * g.stride > 256) return false;
  g.BN = 128;
  g.tiles_m = (g.B * g.pix + 127) / 128;
  g.tiles_n = g.Cout / g.BN;
  g.total_iters = (g.Cin / 64) * g.ks * g.ks + (o.p[8] ? g.sc_cin / 64 : 0);
  // small layers only: at most ~1.3 items per CTA before the K split; the big ones keep their persistent kernels
  return g.tiles_m * g.tiles_n <= 192;
}

static int fused_attn_max() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("RLDM_FUSE_ATTN_MAX");     // longest attention sequence that runs inside a fused run
    v = e ? atoi(e) : 256;
  }
  return v;
}

extern "C" int rldm_fused_supported(const rldm_op* op) {
  switch (op->kind) {
    case RLDM_OP_PREP: {
      // scale/shift table in shared memory; passes over the full-resolution tensors keep their own launch (4 CTAs per
      // SM hide more latency than the 148 x 256 threads of this kernel)
      const long long elems = static_cast<long long>(op->i[5]) * (op->i[6] * op->i[4] + 2) * op->i[7] * op->i[4] * (op->i[0] + op->i[1]);
      return ((op->i[0] + op->i[1]) <= 2048 && elems <= (4ll << 20)) ? 1 : 0;
    }
    case RLDM_OP_CONV_TC: {
      if (op->p[19]) return 0;        // emitting convolutions (rldm_conv_tc_emit) keep their own launch
      FConvGeom g;
      return fused_conv_geom(*op, g) ? 1 : 0;
    }
    case RLDM_OP_ATTENTION: {
      // short sequences only (the mma.sync body); N >= 512 keeps the pipelined tcgen05 kernel and its own launch
      const int N = op->i[1], C = op->i[2];
      return (N % 64 == 0 && N <= fused_attn_max() && C % 8 == 0) ? 1 : 0;
    }
    default:
      return 0;
  }
}

extern "C" long long rldm_fused_ws_bytes(const rldm_op* ops, int n_ops) {
  long long need = 0;
  for (int k = 0; k < n_ops; ++k) {
    if (ops[k].kind != RLDM_OP_CONV_TC) continue;
    FConvGeom g;
    if (!fused_conv_geom(ops[k], g)) continue;
    const int tiles = g.tiles_m * g.tiles_n;
    int ksplit = env().n_sms / tiles;
    if (ksplit > g.total_iters / 3) ksplit = g.total_iters / 3;
    if (ksplit < 1) ksplit = 1;
    const long long b = static_cast<long long>(tiles) * ksplit * 128 * g.BN * 4;
    if (b > need) need = b;
  }
  return need;
}

extern "C" void rldm_fused_destroy(rldm_fused* h) {
  if (!h) return;
  cudaFree(h->d_phases);
  cudaFree(h->d_maps);
  cudaFree(h->d_bar);
  delete h;
}

extern "C" int rldm_fused_create(const rldm_op* ops, int n_ops, float* ws, long long ws_bytes, rldm_fused** out) {
  RLDM_CHECK(out != nullptr && n_ops > 0, "fused_create: bad arguments");
  *out = nullptr;
  EncodeTiledFn encode = fused_get_encode();
  RLDM_CHECK(encode != nullptr, "fused_create: cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  const int n_sms = env().n_sms;
  std::vector<FPhase> phases;
  std::vector<CUtensorMap> maps;
  for (int k = 0; k < n_ops; ++k) {
    const rldm_op& o = ops[k];
    RLDM_CHECK(rldm_fused_supported(&o), "fused_create: op %d (kind %d) cannot run inside a fused segment", k, o.kind);
    FPhase ph;
    memset(&ph, 0, sizeof(ph));
    if (o.kind == RLDM_OP_PREP) {
      ph.kind = F_PREP;
      PrepArgs& a = ph.prep.a;
      const int c0 = o.i[0], c1 = o.i[1], G = o.i[2], silu = o.i[3], up = o.i[4], B = o.i[5], W = o.i[6], H = o.i[7];
      a.x0 = (const float*)o.p[0]; a.x1 = (const float*)o.p[1]; a.sums = (const double*)o.p[2];
      a.pairs0 = (const double*)o.p[9]; a.pairs1 = (const double*)o.p[10];
      a.gamma = (const float*)o.p[3]; a.beta = (const float*)o.p[4];
      a.out = (__half*)o.p[5]; a.out_lo = (__half*)o.p[6]; a.raw = (__half*)o.p[7]; a.raw_lo = (__half*)o.p[8];
      a.eps = o.f[0]; a.c0 = c0; a.c1 = c1; a.G = G; a.silu = silu; a.up = up; a.circular = o.i[8]; a.W = W; a.H = H;
      RLDM_CHECK(c0 % 8 == 0 && c1 % 8 == 0 && (up == 1 || up == 2), "fused_create: prep channels / up");
      const int out_pix = (W * up + 2) * H * up;
      int chunks = n_sms / B;                         // all virtual blocks in ONE pass of the resident CTAs
      if (chunks < 1) chunks = 1;
      int ppb = (out_pix + chunks - 1) / chunks;
      if (ppb < 8) ppb = 8;
      chunks = (out_pix + ppb - 1) / ppb;
      a.pix_per_block = ppb;
      ph.prep.vgx = chunks; ph.prep.vgy = B;
      phases.push_back(ph);
    } else if (o.kind == RLDM_OP_ATTENTION) {
      ph.kind = F_ATTN;
      ph.attn.qkv = (const float*)o.p[0]; ph.attn.out = (__half*)o.p[1]; ph.attn.out_lo = (__half*)o.p[2];
      ph.attn.B = o.i[0]; ph.attn.N = o.i[1]; ph.attn.C = o.i[2]; ph.attn.H = o.i[3];
      phases.push_back(ph);
    } else {
      FConvGeom g;
      fused_conv_geom(o, g);
      const uint16_t* x = (const uint16_t*)o.p[0];
      const uint16_t* x_lo = g.terms == 3 ? (const uint16_t*)o.p[6] : nullptr;
      const uint16_t* wgt = (const uint16_t*)o.p[1];
      const uint16_t* sc_x = (const uint16_t*)o.p[8];
      const uint16_t* sc_x_lo = g.terms == 3 ? (const uint16_t*)o.p[9] : nullptr;
      const uint16_t* sc_wgt = (const uint16_t*)o.p[10];
      RLDM_CHECK(!sc_x || (sc_wgt && (g.terms != 3 || sc_x_lo)), "fused_create: shortcut operands of op %d", k);
      const int wp = g.terms >= 2 ? 2 : 1;
      auto encode_act = [&](CUtensorMap* m, const uint16_t* ptr, int C, int s) -> CUresult {
        cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)g.H, (cuuint64_t)(g.W + 2), (cuuint64_t)g.B};
        cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)g.H * C * 2, (cuuint64_t)(g.W + 2) * g.H * C * 2};
        cuuint32_t box[4] = {64u, (cuuint32_t)(g.Ho * s), (cuuint32_t)(g.ncols * s), (cuuint32_t)g.nb};
        cuuint32_t estr[4] = {1, (cuuint32_t)s, (cuuint32_t)s, 1};
        return encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<uint16_t*>(ptr), gdim, gstr, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      };
      auto encode_wgt = [&](CUtensorMap* m, const uint16_t* ptr, int C, int rows) -> CUresult {
        cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)rows};
        cuuint64_t gstr[1] = {(cuuint64_t)C * 2};
        cuuint32_t box[2] = {64u, (cuuint32_t)g.BN};
        cuuint32_t estr[2] = {1, 1};
        return encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<uint16_t*>(ptr), gdim, gstr, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      };
      CUtensorMap m6[6];
      CUresult r = encode_act(&m6[0], x, g.Cin, g.stride);
      RLDM_CHECK(r == CUDA_SUCCESS, "fused_create: cuTensorMapEncodeTiled(A) failed: %d", (int)r);
      m6[1] = m6[0];
      if (x_lo) {
        r = encode_act(&m6[1], x_lo, g.Cin, g.stride);
        RLDM_CHECK(r == CUDA_SUCCESS, "fused_create: cuTensorMapEncodeTiled(A lo) failed: %d", (int)r);
      }
      r = encode_wgt(&m6[2], wgt, g.Cin, wp * g.ks * g.ks * g.Cout);
      RLDM_CHECK(r == CUDA_SUCCESS, "fused_create: cuTensorMapEncodeTiled(B) failed: %d", (int)r);
      m6[3] = m6[0]; m6[4] = m6[1]; m6[5] = m6[2];
      if (sc_x) {
        r = encode_act(&m6[3], sc_x, g.sc_cin, 1);
        RLDM_CHECK(r == CUDA_SUCCESS, "fused_create: cuTensorMapEncodeTiled(shortcut A) failed: %d", (int)r);
        m6[4] = m6[3];
        if (sc_x_lo) {
          r = encode_act(&m6[4], sc_x_lo, g.sc_cin, 1);
          RLDM_CHECK(r == CUDA_SUCCESS, "fused_create: cuTensorMapEncodeTiled(shortcut A lo) failed: %d", (int)r);
        }
        r = encode_wgt(&m6[5], sc_wgt, g.sc_cin, wp * g.Cout);
        RLDM_CHECK(r == CUDA_SUCCESS, "fused_create: cuTensorMapEncodeTiled(shortcut B) failed: %d", (int)r);
      }
      FConv c;
      memset(&c, 0, sizeof(c));
      c.map = static_cast<int>(maps.size());
      for (int j = 0; j < 6; ++j) maps.push_back(m6[j]);
      c.bias = (const float*)o.p[2]; c.temb = (const float*)o.p[3]; c.residual = (const float*)o.p[4];
      c.out = (float*)o.p[5]; c.stats = (double*)o.p[7];
      c.temb_stride = o.i[0];
      c.M_total = g.B * g.pix; c.Wo = g.Wo; c.Ho = g.Ho; c.pix_per_img = g.pix; c.Cout = g.Cout;
      c.ks = g.ks; c.stride = g.stride; c.pad_lo = g.pad_lo;
      c.main_iters = (g.Cin / 64) * g.ks * g.ks;
      c.total_iters = g.total_iters;
      c.BN = g.BN; c.terms = g.terms; c.tiles_m = g.tiles_m; c.tiles_n = g.tiles_n;
      const int tiles = g.tiles_m * g.tiles_n;
      int ksplit = n_sms / tiles;                      // fill the machine ...
      if (ksplit > g.total_iters / 3) ksplit = g.total_iters / 3;      // ... with at least 3 K steps per item
      if (ksplit < 1) ksplit = 1;
      c.ksplit = ksplit;
      c.stats_G = g.Cout / 2;
      RLDM_CHECK(!c.stats || g.pix >= 64, "fused_create: fused statistics need >= 64 pixels per image");
      const long long need = static_cast<long long>(tiles) * ksplit * 128 * g.BN * 4;
      RLDM_CHECK(ws != nullptr && need <= ws_bytes && (reinterpret_cast<uintptr_t>(ws) & 15) == 0,
                 "fused_create: split-K workspace too small (%lld needed, %lld given)", need, ws_bytes);
      c.ws = ws;
      ph.kind = F_CONV_MAIN; ph.conv = c; phases.push_back(ph);
      ph.kind = F_CONV_FIN; phases.push_back(ph);
    }
  }
  rldm_fused* h = new rldm_fused();
  for (const FPhase& ph : phases) h->phase_kinds.push_back(ph.kind);
  h->n_phases = static_cast<int>(phases.size());
  h->n_ops = n_ops;
  h->grid = n_sms;
  cudaError_t e = cudaMalloc(&h->d_phases, phases.size() * sizeof(FPhase));
  if (e == cudaSuccess) e = cudaMalloc(&h->d_maps, (maps.empty() ? 1 : maps.size()) * sizeof(CUtensorMap));
  if (e == cudaSuccess) e = cudaMalloc(&h->d_bar, 2 * sizeof(unsigned));
  if (e == cudaSuccess) e = cudaMemcpy(h->d_phases, phases.data(), phases.size() * sizeof(FPhase), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && !maps.empty())
    e = cudaMemcpy(h->d_maps, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemset(h->d_bar, 0, 2 * sizeof(unsigned));
  if (e != cudaSuccess) {
    rldm_fused_destroy(h);
    set_error("fused_create: %s", cudaGetErrorString(e));
    return 2;
  }
  *out = h;
  return 0;
}

// profiling aids (not part of include/rldm.h): per-phase device timestamps of CTA 0 (3 per phase) and the phase kinds
extern "C" int rldm_fused_debug(rldm_fused* h, unsigned long long* dev_buf) { h->dbg = dev_buf; return h->n_phases; }
extern "C" int rldm_fused_phase_kind(rldm_fused* h, int k) { return (k >= 0 && k < h->n_phases) ? h->phase_kinds[k] : -1; }

extern "C" int rldm_fused_run(rldm_fused* h, void* stream) {
  RLDM_CHECK(h != nullptr, "fused_run: NULL handle");
  static int max_ctas = -1, bar_mode = 1;
  if (max_ctas < 0) {
    const char* e = getenv("RLDM_FUSE_BAR");
    bar_mode = e ? atoi(e) : 1;
    RLDM_CUDA(cudaFuncSetAttribute(fused_levels_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFSmem));
    RLDM_CUDA(cudaFuncSetAttribute(fused_levels_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFSmem));
    int per_sm = 0;
    RLDM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fused_levels_kernel<1>, kFThreads, kFSmem));
    max_ctas = per_sm * env().n_sms;
  }
  // the grid barrier needs every CTA resident at once
  RLDM_CHECK(h->grid <= max_ctas, "fused_run: %d CTAs cannot be co-resident (limit %d)", h->grid, max_ctas);
  if (bar_mode == 0)
    fused_levels_kernel<0><<<h->grid, kFThreads, kFSmem, as_stream(stream)>>>(h->d_phases, h->n_phases, h->d_maps, h->d_bar, h->dbg);
  else
    fused_levels_kernel<1><<<h->grid, kFThreads, kFSmem, as_stream(stream)>>>(h->d_phases, h->n_phases, h->d_maps, h->d_bar, h->dbg);
  RLDM_LAUNCH_CHECK();
  return 0;
}
