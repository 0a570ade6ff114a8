// blocks.cuh -- CTA-level bodies shared by the stand-alone kernels (ops.cu) and the fused multi-layer kernel
// (fused_levels.cu).  A body processes ONE virtual block (vbx, vby, vbz); the stand-alone kernels call it with
// blockIdx, the fused kernel walks virtual blocks in a grid-stride loop.
// COH = true (fused kernel): every tensor the SAME kernel may have written earlier is read with ld.global.cg (L2
// only) -- L1 is not coherent across SMs inside one launch.  COH = false: the read-only path (__ldg).
#pragma once
#include "common.cuh"

namespace rldm {

template <bool COH> __device__ __forceinline__ float4 ld_act4(const float4* p) { return COH ? __ldcg(p) : __ldg(p); }
template <bool COH> __device__ __forceinline__ double ld_actd(const double* p) { return COH ? __ldcg(p) : __ldg(p); }

// 8 fp32 values -> fp16 hi (+ optional lo residual), one 16 B store each
__device__ __forceinline__ void store_split8(const float (&v)[8], __half* out, __half* out_lo, size_t o) {
  __align__(16) __half2 h[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
  *reinterpret_cast<uint4*>(out + o) = *reinterpret_cast<const uint4*>(h);
  if (out_lo) {   // residual of the fp16 rounding: x = hi + lo to ~22 bits
    __align__(16) __half2 l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 hf = __half22float2(h[j]);
      l[j] = __floats2half2_rn(v[2 * j] - hf.x, v[2 * j + 1] - hf.y);
    }
    *reinterpret_cast<uint4*>(out_lo + o) = *reinterpret_cast<const uint4*>(l);
  }
}

struct PrepArgs {
  const float* x0; const float* x1;
  const double* sums; const double* pairs0; const double* pairs1;
  const float* gamma; const float* beta;
  __half* out; __half* out_lo; __half* raw; __half* raw_lo;
  float eps;
  int c0, c1, G, silu, up, circular, W, H, pix_per_block;
};

// prep: y = [silu]([gn](concat(x0,x1))) cast to fp16, optionally nearest-2x upsampled.
// prep_range: `nthreads` threads (ids `tid`, synchronised by named barrier `bar_id`; 0 = __syncthreads of a CTA of
// exactly that many threads) produce the padded output pixels [p_begin, p_end) of image b, channels [ch_lo, ch_hi).
// Thread = 8 channels of one OUTPUT pixel (two float4 loads, one 16 B store).  shf: 2*(ch_hi-ch_lo) floats of shared
// memory (scale, shift).
// U items per thread and step; PIPE: the next step's loads are issued before the current step is consumed (the stand-alone
// kernel: 2 items, pipelined, 1024 threads per SM); the in-kernel producer of the small-layer convolution has only one
// CTA of 192 threads per SM and therefore issues 8 items (16 x 16 B loads) per thread at once.
template <bool COH, int U = 2, bool PIPE = true>
__device__ __forceinline__ void prep_range(const PrepArgs& a, int b, int p_begin, int p_end, int ch_lo, int ch_hi, float* shf,
                                           int tid, int nthreads, int bar_id) {
  const int c0 = a.c0, c1 = a.c1, W = a.W, H = a.H, up = a.up, G = a.G;
  const int C = c0 + c1;
  const int CR = ch_hi - ch_lo;
  float* sc = shf;
  float* sf = shf + CR;
  const bool norm = a.sums != nullptr || a.pairs0 != nullptr;
  // Output is W-PADDED: (B, Wo+2, Ho, C); padded column wp holds image column (wp-1) mod Wo, i.e. the
  // circular halo of `ldm/utils.py:47` is materialised here for free (zeros when !circular), so every
  // conv tap is a plain TMA box.
  const int Wo = W * up, Ho = H * up;
  const int oct_per_pix = CR >> 3;
  const int out_pix = (Wo + 2) * Ho;
  const int total = (p_end - p_begin) * oct_per_pix;
  // (pixel, channel octet) of this thread's item advance incrementally: no integer division in the loop; Ho is a
  // power of two in every reference geometry (shift), otherwise one division per item
  int pl = tid / oct_per_pix, oc = tid - pl * oct_per_pix;
  const int step_p = nthreads / oct_per_pix, step_o = nthreads - step_p * oct_per_pix;
  const int sh_h = (Ho & (Ho - 1)) == 0 ? 31 - __clz(Ho) : -1;
  // Two items per step, software-pipelined one step ahead: the loads of step k+1 are in flight while step k is
  // consumed, and the loads of the FIRST step are issued before the GroupNorm prologue below (its dependent moment
  // loads and the first data loads overlap; small passes are pure latency).
  struct Item {
    size_t o;
    int c;
    bool live, zero;
    float4 v0, v1;
  };
  auto issue = [&](int i, Item (&it)[U]) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      it[u].live = i + u * nthreads < total;
      while (oc >= oct_per_pix) { oc -= oct_per_pix; ++pl; }
      const int po = p_begin + pl;
      const int c = ch_lo + (oc << 3);
      pl += step_p; oc += step_o;                         // advance to this thread's next item
      const int wp = sh_h >= 0 ? po >> sh_h : po / Ho;
      const int ho = po - wp * Ho;
      int wo = wp - 1;
      const bool halo = wo < 0 || wo >= Wo;
      if (wo < 0) wo += Wo;
      if (wo >= Wo) wo -= Wo;
      it[u].o = (static_cast<size_t>(b) * out_pix + po) * C + c;
      it[u].c = c - ch_lo;
      it[u].zero = halo && !a.circular;
      it[u].v0 = it[u].v1 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (it[u].live && !it[u].zero) {
        const int pin = (up == 2) ? (wo >> 1) * H + (ho >> 1) : wo * H + ho;
        const size_t pix = static_cast<size_t>(b) * W * H + pin;
        const float* src = (c < c0) ? a.x0 + pix * c0 + c : a.x1 + pix * c1 + (c - c0);
        it[u].v0 = ld_act4<COH>(reinterpret_cast<const float4*>(src));
        it[u].v1 = ld_act4<COH>(reinterpret_cast<const float4*>(src) + 1);
      }
    }
  };
  Item cur[U];
  issue(tid, cur);
  if (norm) {
    // group moments: either the (sum, sum^2) per (image, group) of rldm_gn_stats, or the per channel-PAIR moments
    // that the producing convolutions accumulated in their epilogues (x0's pairs, then x1's for a skip concat)
    const int cpg = C / G;
    const double inv_n = 1.0 / (static_cast<double>(W) * H * cpg);
    for (int cr = tid; cr < CR; cr += nthreads) {
      const int c = ch_lo + cr;
      const int g = c / cpg;
      double s = 0.0, ss = 0.0;
      if (a.sums) {
        s = ld_actd<COH>(a.sums + (static_cast<size_t>(b) * G + g) * 2);
        ss = ld_actd<COH>(a.sums + (static_cast<size_t>(b) * G + g) * 2 + 1);
      } else {
        for (int cc = g * cpg; cc < (g + 1) * cpg; cc += 2) {
          const double* pr = cc < c0 ? a.pairs0 + (static_cast<size_t>(b) * (c0 / 2) + cc / 2) * 2
                                     : a.pairs1 + (static_cast<size_t>(b) * (c1 / 2) + (cc - c0) / 2) * 2;
          s += ld_actd<COH>(pr);
          ss += ld_actd<COH>(pr + 1);
        }
      }
      const double mean = s * inv_n;
      double var = ss * inv_n - mean * mean;              // the cancellation is why the moments are doubles
      if (var < 0) var = 0;
      const float rstd = rsqrtf(static_cast<float>(var) + a.eps);
      const float sa = rstd * __ldg(a.gamma + c);
      sc[cr] = sa;
      sf[cr] = __ldg(a.beta + c) - static_cast<float>(mean) * sa;
    }
    if (bar_id == 0) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nthreads) : "memory");
  }
  __half* out = a.out; __half* out_lo = a.out_lo; __half* raw = a.raw; __half* raw_lo = a.raw_lo;
  for (int i = tid; i < total; i += U * nthreads) {
    Item nxt[PIPE ? U : 1];
    if (PIPE) {
#pragma unroll
      for (int u = 0; u < U; ++u) nxt[u].live = false;
      if (i + U * nthreads < total) issue(i + U * nthreads, reinterpret_cast<Item(&)[U]>(nxt));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!cur[u].live) continue;
      const size_t o = cur[u].o;
      if (cur[u].zero) {
        *reinterpret_cast<uint4*>(out + o) = make_uint4(0, 0, 0, 0);
        if (out_lo) *reinterpret_cast<uint4*>(out_lo + o) = make_uint4(0, 0, 0, 0);
        if (raw) *reinterpret_cast<uint4*>(raw + o) = make_uint4(0, 0, 0, 0);
        if (raw_lo) *reinterpret_cast<uint4*>(raw_lo + o) = make_uint4(0, 0, 0, 0);
        continue;
      }
      const int c = cur[u].c;
      float v[8] = {cur[u].v0.x, cur[u].v0.y, cur[u].v0.z, cur[u].v0.w, cur[u].v1.x, cur[u].v1.y, cur[u].v1.z, cur[u].v1.w};
      if (raw) store_split8(v, raw, raw_lo, o);       // second output: the un-normalised operand (1x1 shortcut input)
      if (norm) {
        const float4 a0 = *reinterpret_cast<const float4*>(sc + c), a1 = *reinterpret_cast<const float4*>(sc + c + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(sf + c), b1 = *reinterpret_cast<const float4*>(sf + c + 4);
        v[0] = fmaf(v[0], a0.x, b0.x); v[1] = fmaf(v[1], a0.y, b0.y); v[2] = fmaf(v[2], a0.z, b0.z); v[3] = fmaf(v[3], a0.w, b0.w);
        v[4] = fmaf(v[4], a1.x, b1.x); v[5] = fmaf(v[5], a1.y, b1.y); v[6] = fmaf(v[6], a1.z, b1.z); v[7] = fmaf(v[7], a1.w, b1.w);
      }
      if (a.silu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = silu_f(v[j]);
      }
      store_split8(v, out, out_lo, o);
    }
    if (PIPE) {
#pragma unroll
      for (int u = 0; u < U; ++u) cur[u] = nxt[u];
    } else if (i + U * nthreads < total) {
      issue(i + U * nthreads, cur);
    }
  }
}

// virtual grid (ceil(outpix/pix_per_block), B) of 256-thread blocks over the whole tensor (rldm_prep, fused levels)
template <bool COH>
__device__ __forceinline__ void prep_block(const PrepArgs& a, int vbx, int vby, float* shf) {
  const int out_pix = (a.W * a.up + 2) * a.H * a.up;
  const int p_begin = vbx * a.pix_per_block;
  prep_range<COH>(a, vby, p_begin, min(p_begin + a.pix_per_block, out_pix), 0, a.c0 + a.c1, shf, threadIdx.x, blockDim.x, 0);
}

// ------------------------------------------------------------------------------------------------
// attention core on the legacy tensor path (mma.sync m16n8k16, fp16 x fp16 -> fp32), split-fp16 operands: the kernel
// of the SHORT sequences (N < 256 or N not a multiple of 128; the tcgen05 kernel takes the rest).
// head_dim 8 is packed as K = [hi | lo]:  S = [q_hi|q_lo] . [k_hi;k_hi] + [q_hi|q_lo] . [k_lo;k_lo],
// and P.V is issued as P_hi.V_hi + P_lo.V_hi + P_hi.V_lo with the FlashAttention-2 register trick (the C
// fragments of two 8-key score blocks ARE the A fragment of one 16-key P block).  ~22-bit operands throughout.
// virtual grid (N/64, C/8, B), 128 threads = 4 warps x 16 queries; K/V of the (b, head) staged per 256-key chunk.
__device__ __forceinline__ void mma_f16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
// split (x, y) into fp16 hi and lo words
__device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x, y);
  const float2 hf = __half22float2(h);
  hi = pack_h2(h);
  lo = pack_h2(__floats2half2_rn(x - hf.x, y - hf.y));
}

constexpr int kAtKT = 256;            // keys per shared-memory chunk
constexpr int kAtVP = kAtKT + 8;      // padded pitch of the transposed V rows (bank-conflict-free B fragments)
constexpr int kAtSmemHalves = 2 * kAtKT * 8 + 2 * 8 * kAtVP;      // sk_hi, sk_lo, sv_hi, sv_lo (fp16 elements)

// tid: 0..127 inside the virtual block; bar_id: named barrier of its 128 threads (0 = __syncthreads of a 128-thread CTA)
template <bool COH>
__device__ __forceinline__ void attention_tc_block(const float* qkv, __half* out, __half* out_lo, int N, int C, int H,
                                                   int vbx, int vby, int vbz, __half* sm, int tid, int bar_id) {
  __half* sk_hi = sm;
  __half* sk_lo = sk_hi + kAtKT * 8;
  __half* sv_hi = sk_lo + kAtKT * 8;                     // transposed: [d][key]
  __half* sv_lo = sv_hi + 8 * kAtVP;
  auto sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory"); };
  const int b = vbz, hd = vby;
  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const size_t row = 3 * static_cast<size_t>(C);
  const float* base = qkv + static_cast<size_t>(b) * N * row + hd * 8;
  const int q0 = vbx * 64 + warp * 16;
  // Q fragment (constant over the key loop); softmax scale 1/sqrt(8) and log2(e) folded in
  const float qs = 0.35355339059327373f * 1.4426950408889634f;
  uint32_t qa[4];
  {
    const float2* p0 = reinterpret_cast<const float2*>(base + (q0 + g) * row + 2 * t);
    const float2* p1 = reinterpret_cast<const float2*>(base + (q0 + g + 8) * row + 2 * t);
    const float2 x0 = COH ? __ldcg(p0) : __ldg(p0);
    const float2 x1 = COH ? __ldcg(p1) : __ldg(p1);
    split2(x0.x * qs, x0.y * qs, qa[0], qa[2]);     // a0 = hi(row g), a2 = lo(row g)   (K cols 8.. = lo part)
    split2(x1.x * qs, x1.y * qs, qa[1], qa[3]);     // a1 = hi(row g+8), a3 = lo(row g+8)
  }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  float o[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = 0; k0 < N; k0 += kAtKT) {
    const int kt = min(kAtKT, N - k0);             // multiple of 64 (checked on the host)
    sync();
    for (int i = tid; i < kt; i += 128) {
      const float* kp = base + (k0 + i) * row + C;
      const float* vp = base + (k0 + i) * row + 2 * C;
      const float4 ka = ld_act4<COH>(reinterpret_cast<const float4*>(kp)), kb = ld_act4<COH>(reinterpret_cast<const float4*>(kp) + 1);
      const float4 va = ld_act4<COH>(reinterpret_cast<const float4*>(vp)), vb = ld_act4<COH>(reinterpret_cast<const float4*>(vp) + 1);
      uint32_t h[4], l[4];
      split2(ka.x, ka.y, h[0], l[0]); split2(ka.z, ka.w, h[1], l[1]);
      split2(kb.x, kb.y, h[2], l[2]); split2(kb.z, kb.w, h[3], l[3]);
      *reinterpret_cast<uint4*>(sk_hi + i * 8) = make_uint4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<uint4*>(sk_lo + i * 8) = make_uint4(l[0], l[1], l[2], l[3]);
      const float vv[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
#pragma unroll
      for (int d = 0; d < 8; ++d) {
        const __half hh = __float2half_rn(vv[d]);
        sv_hi[d * kAtVP + i] = hh;
        sv_lo[d * kAtVP + i] = __float2half_rn(vv[d] - __half2float(hh));
      }
    }
    sync();
    for (int c0 = 0; c0 < kt; c0 += 64) {
      // ---- S = Q K^T for 64 keys: 8 blocks of 8 keys, 2 MMAs each
      float sfr[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sfr[j][0] = sfr[j][1] = sfr[j][2] = sfr[j][3] = 0.f;
        const int key = c0 + j * 8 + g;
        const uint32_t kh = *reinterpret_cast<const uint32_t*>(sk_hi + key * 8 + 2 * t);
        const uint32_t kl = *reinterpret_cast<const uint32_t*>(sk_lo + key * 8 + 2 * t);
        mma_f16_16816(sfr[j], qa, kh, kh);          // q_hi.k_hi + q_lo.k_hi
        mma_f16_16816(sfr[j], qa, kl, kl);          // q_hi.k_lo (+ q_lo.k_lo)
      }
      // ---- online softmax (rows g and g+8; a row is spread over the 4 threads of a quad)
      float r0 = sfr[0][0], r1 = sfr[0][2];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        r0 = fmaxf(r0, fmaxf(sfr[j][0], sfr[j][1]));
        r1 = fmaxf(r1, fmaxf(sfr[j][2], sfr[j][3]));
      }
      r0 = fmaxf(r0, __shfl_xor_sync(0xffffffffu, r0, 1)); r0 = fmaxf(r0, __shfl_xor_sync(0xffffffffu, r0, 2));
      r1 = fmaxf(r1, __shfl_xor_sync(0xffffffffu, r1, 1)); r1 = fmaxf(r1, __shfl_xor_sync(0xffffffffu, r1, 2));
      const float n0 = fmaxf(m0, r0), n1 = fmaxf(m1, r1);
      const float cr0 = exp2f(m0 - n0), cr1 = exp2f(m1 - n1);
      m0 = n0; m1 = n1;
      l0 *= cr0; l1 *= cr1;
      o[0] *= cr0; o[1] *= cr0; o[2] *= cr1; o[3] *= cr1;
      // ---- P = exp2(S - m), O += P V : 4 blocks of 16 keys, 3 MMAs each
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        uint32_t ph[4], pl[4];
        {
          const float p0 = exp2f(sfr[2 * jj][0] - m0), p1 = exp2f(sfr[2 * jj][1] - m0);
          const float p2 = exp2f(sfr[2 * jj][2] - m1), p3 = exp2f(sfr[2 * jj][3] - m1);
          const float p4 = exp2f(sfr[2 * jj + 1][0] - m0), p5 = exp2f(sfr[2 * jj + 1][1] - m0);
          const float p6 = exp2f(sfr[2 * jj + 1][2] - m1), p7 = exp2f(sfr[2 * jj + 1][3] - m1);
          l0 += (p0 + p1) + (p4 + p5);
          l1 += (p2 + p3) + (p6 + p7);
          split2(p0, p1, ph[0], pl[0]);   // row g,   keys 2t,2t+1
          split2(p2, p3, ph[1], pl[1]);   // row g+8, keys 2t,2t+1
          split2(p4, p5, ph[2], pl[2]);   // row g,   keys 8+2t,..
          split2(p6, p7, ph[3], pl[3]);   // row g+8, keys 8+2t,..
        }
        const int kb16 = c0 + jj * 16;
        const uint32_t vh0 = *reinterpret_cast<const uint32_t*>(sv_hi + g * kAtVP + kb16 + 2 * t);
        const uint32_t vh1 = *reinterpret_cast<const uint32_t*>(sv_hi + g * kAtVP + kb16 + 8 + 2 * t);
        const uint32_t vl0 = *reinterpret_cast<const uint32_t*>(sv_lo + g * kAtVP + kb16 + 2 * t);
        const uint32_t vl1 = *reinterpret_cast<const uint32_t*>(sv_lo + g * kAtVP + kb16 + 8 + 2 * t);
        mma_f16_16816(o, ph, vh0, vh1);
        mma_f16_16816(o, pl, vh0, vh1);
        mma_f16_16816(o, ph, vl0, vl1);
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  // W-padded operand layout (B, W+2, H, C): token n lands at padded pixel H + n
  const size_t ob = (static_cast<size_t>(b) * (N + 2 * H) + H) * C + hd * 8 + 2 * t;
  uint32_t h0, lo0, h1, lo1;
  split2(o[0] * i0, o[1] * i0, h0, lo0);
  split2(o[2] * i1, o[3] * i1, h1, lo1);
  *reinterpret_cast<uint32_t*>(out + ob + static_cast<size_t>(q0 + g) * C) = h0;
  *reinterpret_cast<uint32_t*>(out + ob + static_cast<size_t>(q0 + g + 8) * C) = h1;
  if (out_lo) {
    *reinterpret_cast<uint32_t*>(out_lo + ob + static_cast<size_t>(q0 + g) * C) = lo0;
    *reinterpret_cast<uint32_t*>(out_lo + ob + static_cast<size_t>(q0 + g + 8) * C) = lo1;
  }
}

}  // namespace rldm
