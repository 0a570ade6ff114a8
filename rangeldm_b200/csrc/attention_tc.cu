// attention_tc.cu -- softmax(Q K^T / sqrt(8)) V for heads of dimension 8 on tcgen05 tensor cores.
//
// Replaces F.scaled_dot_product_attention inside diffusers' AttnProcessor2_0 for the AttentionBlocks of
// UNet2DModel (call sites `ldm/pipelines.py:239,360`): token grids are short-H x long-W
// (1024 tokens at 128x8, 256 at 64x4), heads are only 8 wide.
//
// A work item = 128 queries of one (image, head); see attention_umma_pipelined_kernel below for the pipeline.
//   * S = Q K^T: head_dim 8 is padded to a K=32 contraction that carries the split-fp16 terms
//         A row = [q_hi | q_lo | q_hi | 0],  B row = [k_hi | k_hi | k_lo | 0]   (q pre-scaled by log2(e)/sqrt(8))
//     -> q_hi.k_hi + q_lo.k_hi + q_hi.k_lo in two tcgen05.mma (K=16 each) into a TMEM score buffer.
//   * softmax warps (thread = query row = TMEM lane): row max, P = exp2(S - max), and P goes BACK INTO THE SAME TMEM
//     COLUMNS as a split-fp16 pair with tcgen05.st (a chunk of 32 scores becomes 16 columns of P_hi and 16 of P_lo, two
//     keys per 32-bit column) -- P never touches shared memory.
//   * O = P_hi [V_hi | V_lo] + P_lo [V_hi | V_lo]: the A operand comes from TMEM (TS form of tcgen05.mma), the B
//     operand (V^T, hi and lo halves side by side) from shared memory.
//   * the result is written as the split-fp16 operand of the to_out projection (W-padded layout of rldm_conv_tc).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace rldm {

constexpr int kAtTile = 128;                    // queries per CTA and keys per tile
constexpr int kAtQBytes = kAtTile * 128;        // 16 KB: 128 rows x 128 B (first 64 B of a row used)
constexpr int kAtTmemCols = 256;

__device__ __forceinline__ uint32_t at_pack(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ void at_split2(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x, y);
  const float2 hf = __half22float2(h);
  hi = at_pack(h);
  lo = at_pack(__floats2half2_rn(x - hf.x, y - hf.y));
}
__device__ __forceinline__ float at_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void at_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void at_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void at_sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void at_sts64(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]: A is M x 16 fp16, two K-elements per 32-bit TMEM column (8 columns)
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
      " tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// max of the 32 scores of one TMEM chunk
__device__ __forceinline__ float at_max32(const uint32_t (&r)[32]) {
  float m0 = __uint_as_float(r[0]), m1 = __uint_as_float(r[1]), m2 = __uint_as_float(r[2]), m3 = __uint_as_float(r[3]);
#pragma unroll
  for (int e = 4; e < 32; e += 4) {
    m0 = fmaxf(m0, __uint_as_float(r[e])); m1 = fmaxf(m1, __uint_as_float(r[e + 1]));
    m2 = fmaxf(m2, __uint_as_float(r[e + 2])); m3 = fmaxf(m3, __uint_as_float(r[e + 3]));
  }
  return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}
// P = 2^(S - m) for the 32 keys of one chunk as split fp16, written over the chunk's own 32 columns (16 columns of
// P_hi followed by 16 of P_lo), with packed fp32 arithmetic (add.f32x2 / fma.f32x2: two lanes per issue slot -- the
// kernel is bound by instruction issue as much as by MUFU); the row sum comes from the MMA (ones row of V^T).
// PLO = false (the consuming projection runs below split-fp16 x3 precision): P is kept as ONE fp16 plane -- no
// residual arithmetic, half the TMEM stores and half the P V MMAs; a probability rounded to 11 bits is averaged over
// hundreds of keys.
// 2^x for a PAIR of arguments x <= 0 on the FMA pipe (no MUFU): Cody-Waite split x = j + f with the magic-number round
// (j = nearest integer, |f| <= 1/2), a degree-4 minimax polynomial of 2^f (relative error 2.7e-6, two hundred times
// below the fp16 rounding of P that follows) and j added into the exponent field.  4 packed fp32 + 2 integer issue
// slots per pair against 2 MUFU.EX2 -- the softmax warps of this kernel are bound by the 16-lane XU pipe, so a share
// of every chunk's exponentials (kAtPolyOf8 pairs out of 8) is moved over (the FlashAttention-4 trick).
// Arguments below -125 are clamped: 2^-125 vanishes in fp16 like the true value would.
__device__ __forceinline__ float2 at_ex2_poly2(float2 x) {
  const float kMagic = 12582912.f;                             // 1.5 * 2^23: (x + kMagic) holds round(x) in its low bits
  x.x = fmaxf(x.x, -125.f); x.y = fmaxf(x.y, -125.f);
  const float2 t = __fadd2_rn(x, make_float2(kMagic, kMagic));
  const float2 f = __fadd2_rn(x, __fadd2_rn(make_float2(kMagic, kMagic), make_float2(-t.x, -t.y)));   // x - round(x)
  float2 p = __ffma2_rn(make_float2(0.009570101276040077f, 0.009570101276040077f), f,
                        make_float2(0.05591785907745361f, 0.05591785907745361f));
  p = __ffma2_rn(p, f, make_float2(0.240247443318367f, 0.240247443318367f));
  p = __ffma2_rn(p, f, make_float2(0.6931217908859253f, 0.6931217908859253f));
  p = __ffma2_rn(p, f, make_float2(0.9999992847442627f, 0.9999992847442627f));
  // low 9 bits of the magic pattern 0x4B400000 are zero: (t_bits << 23) is exactly j << 23
  p.x = __int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23));
  p.y = __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23));
  return p;
}
// POLY pairs of every 8 take the polynomial (0 = all MUFU), spread evenly over the chunk so that MUFU and FMA work
// interleave in the instruction stream
__host__ __device__ constexpr bool at_use_poly(int e, int poly) { return ((e % 8 + 1) * poly) / 8 != ((e % 8) * poly) / 8; }

template <bool PLO, int POLY>
__device__ __forceinline__ void at_exp_store32_packed(const uint32_t (&r)[32], float m, uint32_t t_chunk) {
  uint32_t h[16], lo[16];
  const float2 nm = make_float2(-m, -m), neg1 = make_float2(-1.f, -1.f);
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    const float2 d = __fadd2_rn(make_float2(__uint_as_float(r[2 * e]), __uint_as_float(r[2 * e + 1])), nm);
    // split-fp16 P (PLO) needs all ~22 bits of the exponential: MUFU only there
    const float2 p = (!PLO && at_use_poly(e, POLY)) ? at_ex2_poly2(d) : make_float2(at_ex2(d.x), at_ex2(d.y));
    const __half2 hh = __floats2half2_rn(p.x, p.y);
    h[e] = at_pack(hh);
    if (PLO) {
      const float2 res = __ffma2_rn(__half22float2(hh), neg1, p);       // p - float(hi): exact
      lo[e] = at_pack(__floats2half2_rn(res.x, res.y));
    }
  }
  tmem_st_32x16(t_chunk, h);
  if (PLO) tmem_st_32x16(t_chunk + 16, lo);
}
__device__ __forceinline__ uint32_t tmem_ld_32x1(uint32_t taddr) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
  return r;
}

// ------------------------------------------------------------------------------------------------
// Persistent, pipelined kernel: 64-key tiles, THREE S/P buffers (64 TMEM columns each) and
// the per-tile outputs folded into registers, so Q K^T of tiles j+1..j+3 is already in TMEM while the softmax
// warps work on tile j -- their MUFU stream never waits for the MMA / barrier round trip.
//   * the key tiles of a CTA form ONE stream across its work items (Q is double-buffered), so Q K^T of the next
//     item's first tiles is issued while the softmax warps finish the current item.
//   * TMEM (256 columns per CTA, two CTAs per SM): S/P buffers at columns 0, 64, 128; two 32-column output slots at
//     192 and 224.
//   * two softmax GROUPS of four warps (thread = query row = TMEM lane in both): group 0 takes the even key tiles and
//     output slot 0, group 1 the odd tiles and slot 1.  Each group keeps its own running (max, sum, o[8]) in registers
//     (fold of the group's previous tile, with the usual 2^(m_old - m_new) rescale, right after P of the current tile
//     is written); the two partial results of a row meet ONCE PER ITEM in shared memory.  No per-tile exchange and
//     no per-tile state in shared memory.
//   * the V^T operand carries a row of ONES below [V_hi | V_lo] (N = 32: rows 0-7 v_hi, 8-15 v_lo, 16 ones, 17-31
//     zeros), so column 16 of an output slot is the tile's row sum, and the elementwise work uses packed fp32
//     (add.f32x2 / fma.f32x2): the softmax threads are bound by instruction issue as much as by MUFU.
//   * K/V ring of 6 tiles (8 KB + 4 KB each); loader warp w owns the global tiles g = w mod 3.
constexpr int kApKT = 64;                       // keys per tile
constexpr int kApRing = 6;
constexpr int kApKBytes = kApKT * 128;          // 8 KB
constexpr int kApVBytes = 32 * 128;             // 4 KB: one K-atom of 64 keys x 32 rows (16 written per tile)
constexpr int kApSlot0 = 192;
constexpr int kApSlot = 32;                     // TMEM columns per output slot
constexpr int kApXFloats = 10;                  // (max, sum, o[8]) of group 0 per row
constexpr int kApThreads = 12 * 32;            // 8 softmax warps, MMA warp, 3 loader warps

// 32-bit shared-window addresses for the per-tile barrier traffic (no generic-pointer arithmetic in the hot loops)
__device__ __forceinline__ void ap_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
    if (++spins > (1u << 24)) {
      printf("rldm: attention mbarrier timeout block %d thread %d\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void ap_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void ap_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
#define AP_BAR(field) (sb_a + static_cast<uint32_t>(offsetof(AttnPipeSmem, field)))

struct AttnPipeSmem {
  uint64_t q_full[2], q_empty[2];
  uint64_t k_full[kApRing], k_empty[kApRing], v_full[kApRing], v_empty[kApRing];
  uint64_t s_full[3], p_full[3], o_full[2];
  uint64_t x_full, x_empty;
  uint32_t tmem_ptr;
  uint32_t pad;
};

template <bool PLO, int POLY>
__global__ void __launch_bounds__(kApThreads, 2)
attention_umma_pipelined_kernel(const float* __restrict__ qkv, __half* __restrict__ out, __half* __restrict__ out_lo, int N,
                                int C, int H, int B) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;                                     // [2][16 KB]
  uint8_t* sK = sQ + 2 * kAtQBytes;                       // [6][8 KB]
  uint8_t* sV = sK + kApRing * kApKBytes;                 // [6][4 KB]
  float* sX = reinterpret_cast<float*>(sV + kApRing * kApVBytes);   // [10][128]: group 0's partial result of an item
  AttnPipeSmem* sb = reinterpret_cast<AttnPipeSmem*>(sX + kApXFloats * kAtTile);
  const uint32_t sQ_a = (smem_u32(smem_raw) + 1023u) & ~1023u;          // the same carve-up as 32-bit shared addresses
  const uint32_t sK_a = sQ_a + 2 * kAtQBytes, sV_a = sK_a + kApRing * kApKBytes;
  const uint32_t sX_a = sV_a + kApRing * kApVBytes, sb_a = sX_a + kApXFloats * kAtTile * 4;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t rowf = 3 * static_cast<size_t>(C);
  const int heads = C / 8;
  const int QB = N / kAtTile;                              // 128-query blocks per (image, head)
  const int TS = N / kApKT;                                // key tiles per item (even, >= 4)
  const int n_items = B * heads * QB;
  const int my_items = (n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int total_tiles = my_items * TS;
  auto item_base = [&](int item, int& q0) -> const float* {
    const int qb = item % QB, bh = item / QB;
    q0 = qb * kAtTile;
    return qkv + static_cast<size_t>(bh / heads) * N * rowf + (bh % heads) * 8;
  };

  pdl_trigger();
  if (threadIdx.x == 0) {
    // one arrival per WARP on the thread-filled barriers (lanes fence, __syncwarp, lane 0 arrives): hundreds of
    // per-thread arrivals on one shared-memory word would serialise every tile
    for (int i = 0; i < 2; ++i) { mbar_init(&sb->q_full[i], 3); mbar_init(&sb->q_empty[i], 1); }
    for (int i = 0; i < kApRing; ++i) {
      mbar_init(&sb->k_full[i], 1); mbar_init(&sb->k_empty[i], 1);
      mbar_init(&sb->v_full[i], 1); mbar_init(&sb->v_empty[i], 1);
    }
    for (int i = 0; i < 3; ++i) { mbar_init(&sb->s_full[i], 1); mbar_init(&sb->p_full[i], 4); }
    for (int i = 0; i < 2; ++i) mbar_init(&sb->o_full[i], 1);
    mbar_init(&sb->x_full, 4);
    mbar_init(&sb->x_empty, 4);
    mbar_fence_init();
  }
  if (warp == 8) tmem_alloc<kAtTmemCols>(&sb->tmem_ptr);
  // rows 16-31 of every V^T tile never change: row 16 = 1.0 (fp16) for all 64 keys, the rest zero
  for (int i = threadIdx.x; i < kApRing * 128; i += kApThreads) {
    const uint32_t v = (i & 127) < 8 ? 0x3C003C00u : 0u;
    at_sts128(smem_u32(sV + (i >> 7) * kApVBytes + 2048) + (i & 127) * 16, v, v, v, v);
  }
  at_fence_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, sb->tmem_ptr, 0);
  pdl_wait();

  if (warp >= 9) {
    // ===================== loaders ==============================================================
    const int w = warp - 9;
    const int qt = threadIdx.x - 9 * 32;                   // 0..95: Q rows qt and qt + 96
    // K: lane <-> rows lane, lane + 32;  V: lane <-> keys 2 lane, 2 lane + 1.  The next tile is always in registers.
    float4 ka[2], kb[2], va[2], vb[2];
    auto load_tile = [&](int g) {
      int q0;
      const float* base = item_base(static_cast<int>(blockIdx.x) + (g / TS) * static_cast<int>(gridDim.x), q0);
      const int t = g % TS;
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const float* kp = base + static_cast<size_t>(t * kApKT + lane + rr * 32) * rowf + C;
        ka[rr] = __ldg(reinterpret_cast<const float4*>(kp)); kb[rr] = __ldg(reinterpret_cast<const float4*>(kp) + 1);
        const float* vp = base + static_cast<size_t>(t * kApKT + 2 * lane + rr) * rowf + 2 * C;
        va[rr] = __ldg(reinterpret_cast<const float4*>(vp)); vb[rr] = __ldg(reinterpret_cast<const float4*>(vp) + 1);
      }
    };
    auto load_q = [&](int n) {                             // Q rows: [q_hi | q_lo | q_hi | 0], scale folded in
      int q0;
      const float* base = item_base(static_cast<int>(blockIdx.x) + n * static_cast<int>(gridDim.x), q0);
      float4 qa[2], qb[2];
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int i = qt + rr * 96;
        if (i < kAtTile) {
          const float* qp = base + static_cast<size_t>(q0 + i) * rowf;
          qa[rr] = __ldg(reinterpret_cast<const float4*>(qp)); qb[rr] = __ldg(reinterpret_cast<const float4*>(qp) + 1);
        }
      }
      mbar_wait(&sb->q_empty[n & 1], ((n >> 1) & 1) ^ 1);  // every Q K^T of item n - 2 has completed
      const float qs = 0.35355339059327373f * 1.4426950408889634f;
      const uint32_t q_base = smem_u32(sQ + (n & 1) * kAtQBytes);
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int i = qt + rr * 96, sw = i & 7;
        if (i < kAtTile) {
          uint32_t h[4], l[4];
          at_split2(qa[rr].x * qs, qa[rr].y * qs, h[0], l[0]); at_split2(qa[rr].z * qs, qa[rr].w * qs, h[1], l[1]);
          at_split2(qb[rr].x * qs, qb[rr].y * qs, h[2], l[2]); at_split2(qb[rr].z * qs, qb[rr].w * qs, h[3], l[3]);
          const uint32_t r = q_base + i * 128;
          at_sts128(r + ((0 ^ sw) << 4), h[0], h[1], h[2], h[3]);
          at_sts128(r + ((1 ^ sw) << 4), l[0], l[1], l[2], l[3]);
          at_sts128(r + ((2 ^ sw) << 4), h[0], h[1], h[2], h[3]);
          at_sts128(r + ((3 ^ sw) << 4), 0u, 0u, 0u, 0u);
        }
      }
      at_fence_async();
      __syncwarp();
      if (lane == 0) at_arrive(&sb->q_full[n & 1]);
    };
    if (w < total_tiles) load_tile(w);
    int q_next = 0;
    for (int g = w; g < total_tiles; g += 3) {
      while (q_next < my_items && q_next <= (g + 3) / TS) load_q(q_next++);   // Q K^T runs three tiles ahead
      const int slot = g % kApRing;
      const uint32_t ph = ((g / kApRing) & 1) ^ 1;
      mbar_wait(&sb->k_empty[slot], ph);
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int i = lane + rr * 32, sw = i & 7;
        uint32_t h[4], l[4];
        at_split2(ka[rr].x, ka[rr].y, h[0], l[0]); at_split2(ka[rr].z, ka[rr].w, h[1], l[1]);
        at_split2(kb[rr].x, kb[rr].y, h[2], l[2]); at_split2(kb[rr].z, kb[rr].w, h[3], l[3]);
        const uint32_t kr = smem_u32(sK + slot * kApKBytes) + i * 128;    // K row: [k_hi | k_hi | k_lo | 0]
        at_sts128(kr + ((0 ^ sw) << 4), h[0], h[1], h[2], h[3]);
        at_sts128(kr + ((1 ^ sw) << 4), h[0], h[1], h[2], h[3]);
        at_sts128(kr + ((2 ^ sw) << 4), l[0], l[1], l[2], l[3]);
        at_sts128(kr + ((3 ^ sw) << 4), 0u, 0u, 0u, 0u);
      }
      at_fence_async();
      __syncwarp();
      if (lane == 0) at_arrive(&sb->k_full[slot]);
      mbar_wait(&sb->v_empty[slot], ph);
      {
        // V^T: row d holds v_hi[d], row 8 + d holds v_lo[d]; this lane's 2 keys are 4 bytes of chunk lane / 4
        const uint32_t vbase = smem_u32(sV + slot * kApVBytes) + ((lane & 3) << 2);
        const int chunk = lane >> 2;
        const float v0[8] = {va[0].x, va[0].y, va[0].z, va[0].w, vb[0].x, vb[0].y, vb[0].z, vb[0].w};
        const float v1[8] = {va[1].x, va[1].y, va[1].z, va[1].w, vb[1].x, vb[1].y, vb[1].z, vb[1].w};
#pragma unroll
        for (int d = 0; d < 8; ++d) {
          uint32_t h01, l01;
          at_split2(v0[d], v1[d], h01, l01);
          const uint32_t off = static_cast<uint32_t>((chunk ^ d) << 4);   // rows d and 8 + d: (row & 7) == d
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(vbase + d * 128 + off), "r"(h01) : "memory");
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(vbase + (8 + d) * 128 + off), "r"(l01) : "memory");
        }
      }
      at_fence_async();
      __syncwarp();
      if (lane == 0) at_arrive(&sb->v_full[slot]);
      if (g + 3 < total_tiles) load_tile(g + 3);
    }
    while (q_next < my_items) load_q(q_next++);
  } else if (warp == 8) {
    // ===================== MMA issuer ===========================================================
    // (all ring positions and barrier parities advance incrementally: no division in the per-tile path)
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_f16(128, kApKT);
      constexpr uint32_t idesc_o = umma_idesc_f16(128, kApSlot);
      uint32_t qk_slot = 0, qk_ring_par = 0, qk_bs = 0, qk_buf = 0, qk_buf_par = 0;
      int qk_t = 0;                                        // tile of the next Q K^T inside its item
      auto issue_qk = [&]() {           // S buffer qk_bs is free: P V of the tile three back was issued before
        if (qk_t == 0) ap_wait(AP_BAR(q_full) + qk_buf * 8, qk_buf_par);
        ap_wait(AP_BAR(k_full) + qk_slot * 8, qk_ring_par);
        tc_fence_after();
        const uint64_t q_desc = umma_desc_sw128(sQ_a + qk_buf * kAtQBytes);
        const uint64_t k_desc = umma_desc_sw128(sK_a + qk_slot * kApKBytes);
        const uint32_t d = tmem + qk_bs * kApKT;
        umma_f16(d, q_desc, k_desc, idesc_s, 0u);
        umma_f16(d, q_desc + 2, k_desc + 2, idesc_s, 1u);
        ap_commit(AP_BAR(s_full) + qk_bs * 8);
        ap_commit(AP_BAR(k_empty) + qk_slot * 8);
        if (++qk_t == TS) {
          ap_commit(AP_BAR(q_empty) + qk_buf * 8);
          qk_t = 0;
          qk_buf ^= 1;
          if (qk_buf == 0) qk_buf_par ^= 1;
        }
        if (++qk_slot == kApRing) { qk_slot = 0; qk_ring_par ^= 1; }
        if (++qk_bs == 3) qk_bs = 0;
      };
      for (int g = 0; g < 3 && g < total_tiles; ++g) issue_qk();
      uint32_t slot = 0, ring_par = 0, bs = 0, bs_par = 0, bo = 0;
      for (int g = 0; g < total_tiles; ++g) {
        ap_wait(AP_BAR(v_full) + slot * 8, ring_par);
        // p_full(g): the owning group has written P(g) and, before that, folded output slot bo of tile g - 2
        ap_wait(AP_BAR(p_full) + bs * 8, bs_par);
        tc_fence_after();
        const uint32_t d = tmem + kApSlot0 + kApSlot * bo;
        const uint32_t p_tm = tmem + bs * kApKT;
        const uint64_t v_desc0 = umma_desc_sw128(sV_a + slot * kApVBytes);
#pragma unroll
        for (int part = 0; part < (PLO ? 2 : 1); ++part) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)                   // K step = 16 keys = 8 columns of chunk ks / 2
            umma_f16_ts(d, p_tm + (ks >> 1) * 32 + part * 16 + (ks & 1) * 8, v_desc0 + 2 * ks, idesc_o, (part | ks) != 0);
        }
        ap_commit(AP_BAR(o_full) + bo * 8);
        ap_commit(AP_BAR(v_empty) + slot * 8);
        if (g + 3 < total_tiles) issue_qk();
        if (++slot == kApRing) { slot = 0; ring_par ^= 1; }
        if (++bs == 3) { bs = 0; bs_par ^= 1; }
        bo ^= 1;
      }
    }
    __syncwarp();
  } else {
    // ===================== softmax warps ========================================================
    const int quad = warp & 3, grp = warp >> 2;
    const int row = quad * 32 + lane;                      // query row == TMEM lane
    const uint32_t t_lane = tmem + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t t_o = t_lane + kApSlot0 + kApSlot * grp; // this group's output slot
    const uint32_t x_row = sX_a + row * 4;
    float m_run = -INFINITY, l_run = 0.f;
    float2 o[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
    float m_prev = 0.f;                                    // max of the group's tile whose output is still in TMEM
    int k = 0;                                             // tiles this group has handed over so far
    auto fold = [&]() {                                    // tile k - 1 of this group -> running state
      const float m_new = fmaxf(m_run, m_prev);
      const float a = at_ex2(m_run - m_new), c = at_ex2(m_prev - m_new);
      m_run = m_new;
      ap_wait(AP_BAR(o_full) + grp * 8, (k - 1) & 1);
      tc_fence_after();
      uint32_t r[16];
      tmem_ld_32x16(t_o, r);
      const float l_tile = __uint_as_float(tmem_ld_32x1(t_o + 16));
      tmem_ld_wait();
      const float2 a2 = make_float2(a, a), c2 = make_float2(c, c);
      l_run = fmaf(l_run, a, l_tile * c);
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        const float2 v = __fadd2_rn(make_float2(__uint_as_float(r[2 * d]), __uint_as_float(r[2 * d + 1])),
                                    make_float2(__uint_as_float(r[8 + 2 * d]), __uint_as_float(r[9 + 2 * d])));
        o[d] = __ffma2_rn(o[d], a2, __fmul2_rn(v, c2));
      }
    };
    auto finalize = [&](int n) {                           // the two groups' partial results of item n meet
      // group 0 owns the even tiles, so it is done with an item one tile EARLIER than group 1: it deposits its partial
      // result and moves on; group 1 (which would otherwise be waited for) merges and writes the output
      if (grp == 0) {
        ap_wait(AP_BAR(x_empty), (n & 1) ^ 1);
        const float vals[kApXFloats] = {m_run, l_run, o[0].x, o[0].y, o[1].x, o[1].y, o[2].x, o[2].y, o[3].x, o[3].y};
#pragma unroll
        for (int j = 0; j < kApXFloats; ++j)
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(x_row + j * kAtTile * 4), "f"(vals[j]) : "memory");
        __syncwarp();
        if (lane == 0) ap_arrive(AP_BAR(x_full));
      } else {
        ap_wait(AP_BAR(x_full), n & 1);
        float v[kApXFloats];
#pragma unroll
        for (int j = 0; j < kApXFloats; ++j)
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[j]) : "r"(x_row + j * kAtTile * 4) : "memory");
        __syncwarp();
        if (lane == 0) ap_arrive(AP_BAR(x_empty));
        const float m = fmaxf(m_run, v[0]);
        const float a0 = at_ex2(m_run - m), a1 = at_ex2(v[0] - m);
        const float inv = 1.0f / fmaf(l_run, a0, v[1] * a1);
        const float s0 = a0 * inv, s1 = a1 * inv;
        uint32_t h[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
          at_split2(fmaf(o[e].x, s0, v[2 + 2 * e] * s1), fmaf(o[e].y, s0, v[3 + 2 * e] * s1), h[e], lo[e]);
        // W-padded operand layout (B, W+2, H, C): token n lands at padded pixel H + n
        const int item = static_cast<int>(blockIdx.x) + n * static_cast<int>(gridDim.x);
        const int qb = item % QB, bh = item / QB;
        const size_t oi = (static_cast<size_t>(bh / heads) * (N + 2 * H) + H + qb * kAtTile + row) * C + (bh % heads) * 8;
        *reinterpret_cast<uint4*>(out + oi) = make_uint4(h[0], h[1], h[2], h[3]);
        if (out_lo) *reinterpret_cast<uint4*>(out_lo + oi) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
      m_run = -INFINITY; l_run = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = make_float2(0.f, 0.f);
    };
    // own tiles g = grp, grp + 2, ...: one stream across items (TS is even); S buffer g % 3 and the parity of g / 3
    // advance incrementally
    const int own_total = total_tiles >> 1, own_per_item = TS >> 1;
    uint32_t bs = grp, s_par = 0;
    int left = own_per_item, n_done = 0;                   // own tiles left in the current item; finished items
    for (; k < own_total; ++k) {
      const uint32_t t_s = t_lane + bs * kApKT;
      ap_wait(AP_BAR(s_full) + bs * 8, s_par);
      tc_fence_after();
      uint32_t r[32];                                      // two cheap passes over the 64 scores of the row
      tmem_ld_32x32(t_s, r);
      tmem_ld_wait();
      float m = at_max32(r);
      tmem_ld_32x32(t_s + 32, r);
      tmem_ld_wait();
      m = fmaxf(m, at_max32(r));
      at_exp_store32_packed<PLO, POLY>(r, m, t_s + 32);               // P = 2^(S - m) in place: [hi 16 | lo 16] per 32-key chunk
      tmem_ld_32x32(t_s, r);
      tmem_ld_wait();
      at_exp_store32_packed<PLO, POLY>(r, m, t_s);
      tmem_st_wait();
      // the group's previous tile has long finished its P V; its slot is overwritten by P V of this tile, which is
      // issued only after the arrival below
      const bool item_done = k > 0 && left == own_per_item;   // that tile was the last one of the previous item
      if (k > 0) fold();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) ap_arrive(AP_BAR(p_full) + bs * 8);
      if (item_done) finalize(n_done++);
      m_prev = m;
      bs += 2;
      if (bs >= 3) { bs -= 3; s_par ^= 1; }
      if (--left == 0) left = own_per_item;
    }
    if (k > 0) {
      fold();
      tc_fence_before();
      finalize(n_done);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc<kAtTmemCols>(tmem);
}

}  // namespace rldm

using namespace rldm;

// Host entry used by rldm_attention (ops.cu).  Returns -1 when the shape is outside this kernel's range (the caller
// then takes the mma.sync / CUDA-core kernels), 0 on success, > 0 on error.
int rldm_attention_umma(const float* qkv, uint16_t* out, uint16_t* out_lo, int B, int N, int C, int H, void* stream) {
  // N a multiple of 128 and at least four 64-key tiles.  Shorter sequences (N = 128: one query block per head) are all
  // prologue and tail; the mma.sync kernel takes them.
  if (N % kAtTile != 0 || C % 8 != 0 || N < 4 * kApKT) return -1;
  const int T = N / kAtTile;
  static int n_sms_p = 0;
  if (n_sms_p == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sms_p, cudaDevAttrMultiProcessorCount, dev);
    if (n_sms_p <= 0) n_sms_p = 148;
  }
  const size_t smem_p = 1024 + 2 * kAtQBytes + kApRing * (kApKBytes + kApVBytes) + kApXFloats * kAtTile * 4 +
                        sizeof(AttnPipeSmem);
  const int items_p = B * (C / 8) * T;
  const int ctas_p = items_p < 2 * n_sms_p ? items_p : 2 * n_sms_p;      // two persistent CTAs per SM
  // out_lo == NULL: the consumer takes single-fp16 activations, so P is carried as one fp16 plane as well, and
  // env().attn_poly of every 8 exponential pairs run as a polynomial on the FMA pipe (RLDM_ATTN_POLY=0..4, default 2)
  using Kern = void (*)(const float*, __half*, __half*, int, int, int, int);
  static const Kern kerns[6] = {attention_umma_pipelined_kernel<true, 0>,  attention_umma_pipelined_kernel<false, 0>,
                                attention_umma_pipelined_kernel<false, 1>, attention_umma_pipelined_kernel<false, 2>,
                                attention_umma_pipelined_kernel<false, 3>, attention_umma_pipelined_kernel<false, 4>};
  static bool attr_p = false;
  if (!attr_p) {
    for (Kern k : kerns)
      RLDM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_p)));
    attr_p = true;
  }
  const int poly = env().attn_poly;
  const Kern kern = out_lo ? kerns[0] : kerns[1 + (poly < 0 || poly > 4 ? 2 : poly)];
  RLDM_CUDA(launch_pdl_cls(1, kern, dim3(ctas_p), dim3(kApThreads), smem_p, as_stream(stream), qkv,
                       reinterpret_cast<__half*>(out), reinterpret_cast<__half*>(out_lo), N, C, H, B));
  RLDM_LAUNCH_CHECK();
  return 0;
}
