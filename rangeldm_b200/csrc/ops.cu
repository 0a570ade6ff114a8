// ops.cu -- the HBM/L2-bound kernels around the tensor-core convolution: GroupNorm statistics,
// the fused normalise+SiLU+concat+upsample+cast "prep", boundary convolutions, attention core,
// time embedding, the fused scheduler step, and the program runner.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "blocks.cuh"

namespace rldm {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static EnvSwitches g_env;
static bool g_env_loaded = false;
static void load_env() {
  auto tri = [](const char* name) { const char* e = getenv(name); return !e ? 1 : (e[0] == '0' ? 0 : (e[0] == 'n' ? 2 : 1)); };
  const char* e = getenv("RLDM_PDL");
  g_env.pdl = e ? atoi(e) : 2;
  e = getenv("RLDM_PREP_PDL_MAX");
  // every prep pass carries the attribute (round 2, plain-fp16 UNet: 228.4 -> 229.7 images/s against the round-1 cut-off of
  // 2e7 elements, which kept it off the decoder's full-resolution passes)
  g_env.prep_pdl_max = e ? static_cast<size_t>(atoll(e)) : static_cast<size_t>(-1);
  g_env.conv_wt = tri("RLDM_CONV_WT");
  g_env.conv_wt_halo = tri("RLDM_CONV_WT_HALO");
  g_env.conv_persistent = getenv("RLDM_NO_PERSISTENT") == nullptr;
  g_env.conv_mt1 = getenv("RLDM_CONV_MT1") != nullptr;
  g_env.conv_mt2_res = getenv("RLDM_CONV_MT2_RES") != nullptr;
  e = getenv("RLDM_WT_PDL");
  g_env.wt_pdl = e ? atoi(e) != 0 : true;
  g_env.wt_pdl_all = e ? atoi(e) == 2 : false;
  e = getenv("RLDM_EMIT_MAXCL");
  g_env.emit_maxcl = e ? atoi(e) : 8;
  e = getenv("RLDM_EMIT_MAXCLM");
  g_env.emit_maxclm = e ? atoi(e) : 1;
  e = getenv("RLDM_PDL_EXTRA");
  g_env.pdl_extra = e ? atoi(e) : 0;
  g_env.attn_mmasync = getenv("RLDM_ATTN_MMASYNC") != nullptr;
  g_env.attn_cudacore = getenv("RLDM_ATTN_CUDACORE") != nullptr;
  e = getenv("RLDM_SMALL_BN64");
  g_env.small_bn64 = e ? atoi(e) : 128;
  g_env.small_bn64_all = getenv("RLDM_SMALL_BN64_ALL") != nullptr;
  e = getenv("RLDM_ATTN_POLY");
  g_env.attn_poly = e ? atoi(e) : 2;
  g_env.nco_pp1 = getenv("RLDM_NCO_PP1") != nullptr;
  int dev = 0;
  g_env.n_sms = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&g_env.n_sms, cudaDevAttrMultiProcessorCount, dev);
  if (g_env.n_sms <= 0) g_env.n_sms = 148;
  g_env_loaded = true;
}
const EnvSwitches& env() {
  if (!g_env_loaded) load_env();
  return g_env;
}
// Programmatic dependent launch policy (RLDM_PDL = 0 | 1 | 2, default 2).  Measured on B200 inside the trajectory
// graph (scripts/launch_gap.py): programmatic edges save ~1 us per node on the latency-bound launches (small
// convolutions 9.6 -> 8.8 us, prep + conv pairs 11.5 -> 10.1 us), but make the long persistent convolutions, the
// attention kernel and the big elementwise passes SLOWER (decoder +8 %) -- early-launched dependents sit on SM
// resources for the whole primary.  Mode 2 therefore marks only the latency-bound launches; 1 marks all.
bool pdl_enabled() { return env().pdl == 1; }
bool pdl_enabled_small() { return env().pdl == 1 || env().pdl == 2; }
bool pdl_enabled_class(int cls) { return env().pdl == 1 || (env().pdl == 2 && (env().pdl_extra & cls) != 0); }

// ------------------------------------------------------------------------------------------------
// GroupNorm statistics.  grid (chunks, B), block 256.  Thread = one float4 of channels, striding
// over the pixels of its chunk; per-channel partials are folded to per-group doubles in shared
// memory and leave the block as one atomicAdd(double) per (group, moment).
__global__ void __launch_bounds__(256)
gn_stats_kernel(const float* __restrict__ x0, int c0, const float* __restrict__ x1, int c1,
                double* __restrict__ sums, int P, int G, int pix_per_block) {
  pdl_entry();
  extern __shared__ double sh[];  // [2][G]
  const int C = c0 + c1;
  const int q_per_pix = C >> 2;                 // float4 quads per pixel
  const int b = blockIdx.y;
  const int p_begin = blockIdx.x * pix_per_block;
  const int p_end = min(p_begin + pix_per_block, P);
  for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) sh[i] = 0.0;
  __syncthreads();
  const int cpg = C / G;
  // each thread owns quad (threadIdx.x % q_per_pix) when blockDim is a multiple of q_per_pix;
  // otherwise it walks quads generically.
  const int total_q = (p_end - p_begin) * q_per_pix;
  if (blockDim.x % q_per_pix == 0) {
    const int quad = threadIdx.x % q_per_pix;
    const int c = quad << 2;
    float s[4] = {0, 0, 0, 0}, ss[4] = {0, 0, 0, 0};
    for (int p = p_begin + threadIdx.x / q_per_pix; p < p_end; p += blockDim.x / q_per_pix) {
      const size_t pix = static_cast<size_t>(b) * P + p;
      const float4 v = (c < c0) ? __ldg(reinterpret_cast<const float4*>(x0 + pix * c0 + c))
                                : __ldg(reinterpret_cast<const float4*>(x1 + pix * c1 + (c - c0)));
      s[0] += v.x; ss[0] += v.x * v.x;
      s[1] += v.y; ss[1] += v.y * v.y;
      s[2] += v.z; ss[2] += v.z * v.z;
      s[3] += v.w; ss[3] += v.w * v.w;
    }
    if ((cpg & 3) == 0) {   // the whole quad lies in one group
      atomicAdd(&sh[c / cpg], (double)s[0] + (double)s[1] + (double)s[2] + (double)s[3]);
      atomicAdd(&sh[G + c / cpg], (double)ss[0] + (double)ss[1] + (double)ss[2] + (double)ss[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        atomicAdd(&sh[(c + j) / cpg], (double)s[j]);
        atomicAdd(&sh[G + (c + j) / cpg], (double)ss[j]);
      }
    }
  } else {
    for (int i = threadIdx.x; i < total_q; i += blockDim.x) {
      const int p = p_begin + i / q_per_pix;
      const int c = (i % q_per_pix) << 2;
      const size_t pix = static_cast<size_t>(b) * P + p;
      const float4 v = (c < c0) ? __ldg(reinterpret_cast<const float4*>(x0 + pix * c0 + c))
                                : __ldg(reinterpret_cast<const float4*>(x1 + pix * c1 + (c - c0)));
      const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        atomicAdd(&sh[(c + j) / cpg], (double)vv[j]);
        atomicAdd(&sh[G + (c + j) / cpg], (double)vv[j] * vv[j]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) {
    const int g = i % G, mom = i / G;
    atomicAdd(&sums[(static_cast<size_t>(b) * G + g) * 2 + mom], sh[i]);
  }
}

// ------------------------------------------------------------------------------------------------
// prep: y = [silu]([gn](concat(x0,x1))) cast to fp16, optionally nearest-2x upsampled (body: blocks.cuh).
// grid (ceil(outpix/pix_per_block), B), block 256.
__global__ void __launch_bounds__(256, 4)
prep_kernel(const PrepArgs a) {
  pdl_entry();
  extern __shared__ float shf[];  // scale[C], shift[C]
  prep_block<false>(a, blockIdx.x, blockIdx.y, shf);
}

// ------------------------------------------------------------------------------------------------
// conv_ref: one thread per (output pixel, output channel); same contract as conv_tc.
__global__ void conv_ref_kernel(const __half* __restrict__ x, const __half* __restrict__ x_lo,
                                const __half* __restrict__ wgt,
                                const float* __restrict__ bias, const float* __restrict__ temb,
                                int temb_stride, const float* __restrict__ residual,
                                float* __restrict__ out, int B, int W, int H, int Cin, int Cout,
                                int ks, int stride, int pad_lo, int circular) {
  pdl_entry();
  const int Wo = W / stride, Ho = H / stride;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(B) * Wo * Ho * Cout;
  if (idx >= total) return;
  const int n = idx % Cout;
  const size_t m = idx / Cout;
  const int ho = m % Ho;
  const int wo = (m / Ho) % Wo;
  const int b = m / (static_cast<size_t>(Ho) * Wo);
  float acc = 0.f;
  for (int i = 0; i < ks; ++i) {
    const int w = stride * wo + i - pad_lo + 1;     // column in the W-padded operand (halo = wrap or zeros)
    for (int j = 0; j < ks; ++j) {
      const int h = stride * ho + j - pad_lo;
      if (h < 0 || h >= H) continue;
      const size_t xo = ((static_cast<size_t>(b) * (W + 2) + w) * H + h) * Cin;
      const __half* wr = wgt + (static_cast<size_t>(i * ks + j) * Cout + n) * Cin;
      const __half* wl = wr + static_cast<size_t>(ks) * ks * Cout * Cin;   // low-order plane
      for (int c = 0; c < Cin; c += 2) {
        float2 a = __half22float2(*reinterpret_cast<const __half2*>(x + xo + c));
        float2 w2 = __half22float2(*reinterpret_cast<const __half2*>(wr + c));
        if (x_lo) {
          const float2 al = __half22float2(*reinterpret_cast<const __half2*>(x_lo + xo + c));
          const float2 wl2 = __half22float2(*reinterpret_cast<const __half2*>(wl + c));
          a.x += al.x; a.y += al.y; w2.x += wl2.x; w2.y += wl2.y;
        }
        acc = fmaf(a.x, w2.x, acc);
        acc = fmaf(a.y, w2.y, acc);
      }
    }
  }
  if (bias) acc += bias[n];
  if (temb) acc += temb[static_cast<size_t>(b) * temb_stride + n];
  if (residual) acc += residual[idx];
  out[idx] = acc;
}

// ------------------------------------------------------------------------------------------------
// conv_in: (B,Cin,W,H) fp32 ref layout -> (B,W,H,Cout) fp32 cl, 3x3 circular/zero, Cin <= 16.
// Each block owns a CONTIGUOUS range of kCinPix-pixel chunks: the weight matrix [9*Cin][Cout] is staged in shared
// memory once per block (it used to be re-read for every 32 pixels, which was most of the kernel's traffic); per
// chunk the im2col columns (9*Cin floats per pixel, wrap and zero pad resolved once) are staged as [K][32 pixels],
// then thread = (4 pixels, 4 output channels) runs a K-long chain of 16 FMAs per pair of shared float4 loads.
// stats (optional): channel-pair moments of the output, [B][Cout/2][2] doubles, accumulated per block while the
// image index stays the same (needs W*H % kCinPix == 0) -- the GroupNorm of the first ResnetBlock2D reads them.
constexpr int kCinPix = 32;
__global__ void __launch_bounds__(256)
conv_in_kernel(const float* __restrict__ x0, int c0, const float* __restrict__ x1, int c1,
               const float* __restrict__ wgt, const float* __restrict__ bias,
               float* __restrict__ out, int B, int W, int H, int Cout, int circular, double* __restrict__ stats,
               int chunks_per_block) {
  extern __shared__ float sh_ci[];
  __shared__ float red_s[512], red_q[512];
  const int Cin = c0 + c1;
  const int K = 9 * Cin;
  float* w_s = sh_ci;                    // [K][Cout]
  float* in_s = sh_ci + K * Cout;        // [kCinPix][K]
  __shared__ int koff[9 * 16];           // k -> (kernel row i, kernel column j, input channel c), K <= 144
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const int tap = k / Cin, c = k - tap * Cin;
    const int i = tap / 3, j = tap - i * 3;
    koff[k] = (i << 16) | (j << 8) | c;
  }
  pdl_trigger();
  for (int i = threadIdx.x * 4; i < K * Cout; i += blockDim.x * 4)    // weights do not depend on prior kernels
    *reinterpret_cast<float4*>(w_s + i) = __ldg(reinterpret_cast<const float4*>(wgt + i));
  pdl_wait();
  const size_t total_pix = static_cast<size_t>(B) * W * H;
  const size_t pix_per_img = static_cast<size_t>(W) * H;
  const int n_chunks = static_cast<int>((total_pix + kCinPix - 1) / kCinPix);
  const int q_per_pix = Cout >> 2;
  const int lanes_pix = blockDim.x / q_per_pix;
  const int quad = threadIdx.x % q_per_pix;
  const int pl = threadIdx.x / q_per_pix;
  const int co = quad << 2;
  const float4 bv = bias ? __ldg(reinterpret_cast<const float4*>(bias + co)) : make_float4(0, 0, 0, 0);
  float s01 = 0.f, q01 = 0.f, s23 = 0.f, q23 = 0.f;
  int cur_img = -1;
  // fold the per-thread pair moments over the pixel lanes and add them to image `img` (one double atomic per pair)
  auto flush = [&](int img) {
    red_s[pl * (Cout >> 1) + (co >> 1)] = s01; red_q[pl * (Cout >> 1) + (co >> 1)] = q01;
    red_s[pl * (Cout >> 1) + (co >> 1) + 1] = s23; red_q[pl * (Cout >> 1) + (co >> 1) + 1] = q23;
    __syncthreads();
    for (int t = threadIdx.x; t < (Cout >> 1); t += blockDim.x) {
      double ds = 0.0, dq = 0.0;
      for (int e = 0; e < lanes_pix; ++e) { ds += red_s[e * (Cout >> 1) + t]; dq += red_q[e * (Cout >> 1) + t]; }
      double* st = stats + (static_cast<size_t>(img) * (Cout >> 1) + t) * 2;
      atomicAdd(st, ds); atomicAdd(st + 1, dq);
    }
    __syncthreads();
    s01 = q01 = s23 = q23 = 0.f;
  };
  const int c_begin = blockIdx.x * chunks_per_block;
  const int c_end = min(n_chunks, c_begin + chunks_per_block);
  for (int chunk = c_begin; chunk < c_end; ++chunk) {
    const size_t p_begin = static_cast<size_t>(chunk) * kCinPix;
    if (stats) {
      const int img = static_cast<int>(p_begin / pix_per_img);        // uniform over the block
      if (img != cur_img) {
        if (cur_img >= 0) flush(cur_img);
        cur_img = img;
      }
    }
    __syncthreads();                     // previous chunk's in_s fully consumed
    // this thread's pixel is fixed (blockDim.x is a multiple of kCinPix): its coordinates are resolved once per chunk
    // (32-bit arithmetic; the (tap, channel) decomposition of k comes from a table built once per block -- the integer
    // divisions of the first version were a third of the kernel's instructions), then the K loop runs in batches of
    // independent loads (the staging is bound by load latency otherwise)
    {
      const int p = threadIdx.x % kCinPix;
      const int pp = static_cast<int>(p_begin) + p;
      const bool live = pp < static_cast<int>(total_pix);
      const int b = live ? pp / static_cast<int>(pix_per_img) : 0;
      const int pin = live ? pp - b * static_cast<int>(pix_per_img) : 0;
      const int w = pin / H, h = pin - w * H;
      const float* xb0 = x0 + static_cast<size_t>(b) * c0 * W * H;
      const float* xb1 = x1 ? x1 + static_cast<size_t>(b) * c1 * W * H : nullptr;
      constexpr int kBatch = 6;
      const int k_step = blockDim.x / kCinPix;
      for (int k0 = threadIdx.x / kCinPix; k0 < K; k0 += k_step * kBatch) {
        float a[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          const int k = k0 + u * k_step;
          a[u] = 0.f;
          if (live && k < K) {
            const int e = koff[k];                             // (i, j, c) = (e >> 16, (e >> 8) & 255, e & 255)
            const int c = e & 255;
            int wi = w + (e >> 16) - 1;
            const int hj = h + ((e >> 8) & 255) - 1;
            bool ok = hj >= 0 && hj < H;
            if (circular) {
              if (wi < 0) wi += W;
              if (wi >= W) wi -= W;
            } else {
              ok = ok && wi >= 0 && wi < W;
            }
            if (ok)
              a[u] = (c < c0) ? __ldg(xb0 + (c * W + wi) * H + hj) : __ldg(xb1 + ((c - c0) * W + wi) * H + hj);
          }
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          const int k = k0 + u * k_step;
          if (k < K) in_s[k * kCinPix + p] = a[u];          // [K][kCinPix]
        }
      }
    }
    __syncthreads();
    // thread = 4 output channels x 4 pixels: every weight float4 read from shared memory feeds 16 FMAs
    for (int pg = pl; pg < kCinPix / 4; pg += lanes_pix) {
      float4 acc[4] = {bv, bv, bv, bv};
#pragma unroll 3
      for (int k = 0; k < K; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(in_s + k * kCinPix + 4 * pg);
        const float4 wv = *reinterpret_cast<const float4*>(w_s + k * Cout + co);
        acc[0].x = fmaf(a.x, wv.x, acc[0].x); acc[0].y = fmaf(a.x, wv.y, acc[0].y);
        acc[0].z = fmaf(a.x, wv.z, acc[0].z); acc[0].w = fmaf(a.x, wv.w, acc[0].w);
        acc[1].x = fmaf(a.y, wv.x, acc[1].x); acc[1].y = fmaf(a.y, wv.y, acc[1].y);
        acc[1].z = fmaf(a.y, wv.z, acc[1].z); acc[1].w = fmaf(a.y, wv.w, acc[1].w);
        acc[2].x = fmaf(a.z, wv.x, acc[2].x); acc[2].y = fmaf(a.z, wv.y, acc[2].y);
        acc[2].z = fmaf(a.z, wv.z, acc[2].z); acc[2].w = fmaf(a.z, wv.w, acc[2].w);
        acc[3].x = fmaf(a.w, wv.x, acc[3].x); acc[3].y = fmaf(a.w, wv.y, acc[3].y);
        acc[3].z = fmaf(a.w, wv.z, acc[3].z); acc[3].w = fmaf(a.w, wv.w, acc[3].w);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const size_t pp = p_begin + 4 * pg + q;
        if (pp < total_pix) {
          *reinterpret_cast<float4*>(out + pp * Cout + co) = acc[q];
          s01 += acc[q].x + acc[q].y; q01 += acc[q].x * acc[q].x + acc[q].y * acc[q].y;
          s23 += acc[q].z + acc[q].w; q23 += acc[q].z * acc[q].z + acc[q].w * acc[q].w;
        }
      }
    }
  }
  if (stats && cur_img >= 0) flush(cur_img);
}

// conv_out: (B,W+2,H,Cin) fp16 clp (hi [+ lo]) -> (B,Cout,W,H) fp32 ref layout, Cout in {2,4,8}.
// Thread = one output pixel (consecutive threads = consecutive beams h, so the 9 neighbour rows are shared
// through L1 and the ref-layout stores are coalesced); weights [9][Cout][Cin] sit in shared memory and are read
// as warp-wide broadcasts.
// LPP lanes share one pixel (splitting the channel loop, butterfly-reduced) when there are too few pixels to
// fill the machine with one thread each.
template <int COUT, int LPP>
__global__ void __launch_bounds__(256)
conv_out_kernel(const __half* __restrict__ x, const __half* __restrict__ x_lo, const float* __restrict__ wgt,
                const float* __restrict__ bias, float* __restrict__ out, int B, int W, int H,
                int Cin, int circular) {
  extern __shared__ float w_s[];     // [9][COUT][Cin]
  pdl_trigger();
  for (int i = threadIdx.x * 4; i < 9 * COUT * Cin; i += blockDim.x * 4)
    *reinterpret_cast<float4*>(w_s + i) = __ldg(reinterpret_cast<const float4*>(wgt + i));
  pdl_wait();
  __syncthreads();
  const size_t total_pix = static_cast<size_t>(B) * W * H;
  const int sub = threadIdx.x % LPP;
  const int ppp = blockDim.x / LPP;                    // pixels per pass of the block
  // grid-stride over pixel groups: the weights are staged once per block, not once per 16 pixels
  for (size_t g0 = static_cast<size_t>(blockIdx.x) * ppp; g0 < total_pix; g0 += static_cast<size_t>(gridDim.x) * ppp) {
  const size_t praw = g0 + threadIdx.x / LPP;
  const bool live = praw < total_pix;
  const size_t pp = live ? praw : total_pix - 1;
  const int h = pp % H;
  const int w = (pp / H) % W;
  const int b = pp / (static_cast<size_t>(H) * W);
  float acc[COUT];
#pragma unroll
  for (int n = 0; n < COUT; ++n) acc[n] = (bias && sub == 0) ? __ldg(bias + n) : 0.f;
  for (int i = 0; i < 3; ++i) {
    const int wi = w + i;                             // W-padded operand: halo columns hold the wrap (or zeros)
    for (int j = 0; j < 3; ++j) {
      const int hj = h + j - 1;
      if (hj < 0 || hj >= H) continue;
      const size_t xo = ((static_cast<size_t>(b) * (W + 2) + wi) * H + hj) * Cin;
      const float* wt = w_s + (i * 3 + j) * COUT * Cin;
      for (int c = sub * 8; c < Cin; c += 8 * LPP) {
        const uint4 hv = __ldg(reinterpret_cast<const uint4*>(x + xo + c));
        float a[8];
        {
          const __half2* hp = reinterpret_cast<const __half2*>(&hv);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 f = __half22float2(hp[k]);
            a[2 * k] = f.x; a[2 * k + 1] = f.y;
          }
        }
        if (x_lo) {
          const uint4 lv = __ldg(reinterpret_cast<const uint4*>(x_lo + xo + c));
          const __half2* lp = reinterpret_cast<const __half2*>(&lv);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 f = __half22float2(lp[k]);
            a[2 * k] += f.x; a[2 * k + 1] += f.y;
          }
        }
#pragma unroll
        for (int n = 0; n < COUT; ++n) {
          const float4 w0 = *reinterpret_cast<const float4*>(wt + n * Cin + c);
          const float4 w1 = *reinterpret_cast<const float4*>(wt + n * Cin + c + 4);
          acc[n] = fmaf(a[0], w0.x, acc[n]); acc[n] = fmaf(a[1], w0.y, acc[n]);
          acc[n] = fmaf(a[2], w0.z, acc[n]); acc[n] = fmaf(a[3], w0.w, acc[n]);
          acc[n] = fmaf(a[4], w1.x, acc[n]); acc[n] = fmaf(a[5], w1.y, acc[n]);
          acc[n] = fmaf(a[6], w1.z, acc[n]); acc[n] = fmaf(a[7], w1.w, acc[n]);
        }
      }
    }
  }
#pragma unroll
  for (int n = 0; n < COUT; ++n) {
#pragma unroll
    for (int o = LPP / 2; o > 0; o >>= 1) acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], o);
    if (live && sub == 0) out[((static_cast<size_t>(b) * COUT + n) * W + w) * H + h] = acc[n];
  }
  }
}

// ------------------------------------------------------------------------------------------------
// norm_conv_out: conv_norm_out (GroupNorm) + SiLU + conv_out (3x3, Cout <= 8) in ONE kernel.
// x (B,W,H,Cin) fp32 cl -> out (B,Cout,W,H) fp32 ref layout.  grid (ceil(W/TW), B), block 256.
// A block stages TW output columns plus one halo column on each side -- normalised and activated on the way in
// (GroupNorm moments as in prep_kernel: group sums or channel-pair moments) -- as fp32 in shared memory with a row
// pitch of Cin + 4*LPP floats (conflict-free float4 reads), the weights [9][Cout][Cin] next to it, and then
// LPP lanes per pixel run the 9-tap x Cin FMA chain from shared memory.  Replaces a prep launch, the fp16 hi/lo
// operand round trip (8 B per element written and 9x re-read through L1) and the old one-thread-per-pixel conv_out.
template <int COUT>
__global__ void __launch_bounds__(256)
norm_conv_out_kernel(const float* __restrict__ x, const double* __restrict__ sums, const double* __restrict__ pairs,
                     const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int G, int silu,
                     const float* __restrict__ wgt, const float* __restrict__ bias, float* __restrict__ out,
                     int W, int H, int Cin, int circular, int TW, int LPP, int PP) {
  extern __shared__ float sh_no[];
  const int pitch = PP > 1 ? Cin + 4 : Cin + 4 * LPP;
  float* sc = sh_no;                       // [Cin]
  float* sf = sc + Cin;                    // [Cin]
  float* w_s = sf + Cin;                   // [9][COUT][Cin]
  float* tile = w_s + 9 * COUT * Cin;      // [(TW+2)*H][pitch]
  pdl_trigger();
  for (int i = threadIdx.x * 4; i < 9 * COUT * Cin; i += blockDim.x * 4)
    *reinterpret_cast<float4*>(w_s + i) = __ldg(reinterpret_cast<const float4*>(wgt + i));
  pdl_wait();
  const int b = blockIdx.y;
  const int w0 = blockIdx.x * TW;
  const bool norm = sums != nullptr || pairs != nullptr;
  if (norm) {
    const int cpg = Cin / G;
    const double inv_n = 1.0 / (static_cast<double>(W) * H * cpg);
    for (int c = threadIdx.x; c < Cin; c += blockDim.x) {
      const int g = c / cpg;
      double s = 0.0, ss = 0.0;
      if (sums) {
        s = sums[(static_cast<size_t>(b) * G + g) * 2];
        ss = sums[(static_cast<size_t>(b) * G + g) * 2 + 1];
      } else {
        for (int cc = g * cpg; cc < (g + 1) * cpg; cc += 2) {
          const double* pr = pairs + (static_cast<size_t>(b) * (Cin / 2) + cc / 2) * 2;
          s += pr[0];
          ss += pr[1];
        }
      }
      const double mean = s * inv_n;
      double var = ss * inv_n - mean * mean;
      if (var < 0) var = 0;
      const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
      const float a = rstd * gamma[c];
      sc[c] = a;
      sf[c] = beta[c] - static_cast<float>(mean) * a;
    }
  }
  __syncthreads();
  // ---- stage (TW + 2) columns: normalise + SiLU on the way in; wrap (or zero) outside [0, W)
  const int c4n = Cin >> 2;
  const int n_items = (TW + 2) * H * c4n;
  // H and Cin/4 are powers of two in every reference config: shifts instead of integer divisions (sh < 0: generic)
  const int sh_c = (c4n & (c4n - 1)) == 0 ? 31 - __clz(c4n) : -1;
  const int sh_h = (H & (H - 1)) == 0 ? 31 - __clz(H) : -1;
  // eight independent 16 B loads in flight per thread before any of them is consumed (the staging is bound by the
  // global-load latency: 20-24 items per thread = three round trips instead of six with four in flight)
  constexpr int kBatch = 8;
  for (int i0 = threadIdx.x; i0 < n_items; i0 += blockDim.x * kBatch) {
    float4 v[kBatch];
    int rr[kBatch], cc[kBatch];
    bool okk[kBatch];
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int i = i0 + u * blockDim.x;
      const int r = sh_c >= 0 ? i >> sh_c : i / c4n;        // row of the tile = col * H + h
      const int c = (i - r * c4n) << 2;
      const int col = sh_h >= 0 ? r >> sh_h : r / H;
      const int h = r - col * H;
      int wc = w0 - 1 + col;
      bool ok = i < n_items;
      if (circular) {
        if (wc < 0) wc += W;
        if (wc >= W) wc -= W;
      } else {
        ok = ok && wc >= 0;
      }
      ok = ok && wc < W;                   // also: columns past a ragged last block are never consumed
      rr[u] = r; cc[u] = c; okk[u] = ok;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok) v[u] = __ldg(reinterpret_cast<const float4*>(x + ((static_cast<size_t>(b) * W + wc) * H + h) * Cin + c));
    }
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      if (i0 + u * blockDim.x >= n_items) break;
      const int c = cc[u];
      float4 t = v[u];
      if (okk[u]) {
        if (norm) {
          t.x = fmaf(t.x, sc[c], sf[c]); t.y = fmaf(t.y, sc[c + 1], sf[c + 1]);
          t.z = fmaf(t.z, sc[c + 2], sf[c + 2]); t.w = fmaf(t.w, sc[c + 3], sf[c + 3]);
        }
        if (silu) { t.x = silu_f(t.x); t.y = silu_f(t.y); t.z = silu_f(t.z); t.w = silu_f(t.w); }
      }
      *reinterpret_cast<float4*>(tile + rr[u] * pitch + c) = t;
    }
  }
  __syncthreads();
  const int sub = threadIdx.x % LPP;
  if (PP == 4) {
    // ---- 3x3 conv from shared memory, register-blocked: a thread owns FOUR vertically adjacent output pixels (and
    // LPP lanes split the channel loop).  Per (kernel column, channel quad) it loads the 6 input rows once and every
    // weight quad once for all four pixels: 18 LDS.128 per 192 FMA instead of 5 per 16 -- the one-pixel loop below is
    // bound by the shared-memory load pipe, not by the FMAs.
    const int hq = H >> 2;
    const int n_grp = TW * hq;
    for (int pg = threadIdx.x / LPP; pg < n_grp; pg += blockDim.x / LPP) {
      const int col = pg / hq, h0 = (pg - col * hq) << 2;
      const int w = w0 + col;
      float acc[4][COUT];
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int n = 0; n < COUT; ++n) acc[q][n] = 0.f;
      for (int i = 0; i < 3; ++i) {
        const float* xcol = tile + (col + i) * H * pitch;
        const float* wt = w_s + i * 3 * COUT * Cin;
        for (int c = sub * 4; c < Cin; c += 4 * LPP) {
          float4 xr[6];
#pragma unroll
          for (int r = 0; r < 6; ++r) {
            const int hj = h0 - 1 + r;
            xr[r] = (hj >= 0 && hj < H) ? *reinterpret_cast<const float4*>(xcol + hj * pitch + c)
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int j = 0; j < 3; ++j) {
#pragma unroll
            for (int n = 0; n < COUT; ++n) {
              const float4 wv = *reinterpret_cast<const float4*>(wt + (j * COUT + n) * Cin + c);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 a = xr[q + j];
                acc[q][n] = fmaf(a.x, wv.x, acc[q][n]); acc[q][n] = fmaf(a.y, wv.y, acc[q][n]);
                acc[q][n] = fmaf(a.z, wv.z, acc[q][n]); acc[q][n] = fmaf(a.w, wv.w, acc[q][n]);
              }
            }
          }
        }
      }
#pragma unroll
      for (int n = 0; n < COUT; ++n) {
        const float bn = bias ? __ldg(bias + n) : 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float v = acc[q][n];
          for (int o = LPP >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          acc[q][n] = v + bn;
        }
        // out is (B, Cout, W, H): the four pixels are contiguous along H
        if (sub == 0 && w < W)
          *reinterpret_cast<float4*>(out + ((static_cast<size_t>(b) * COUT + n) * W + w) * H + h0) =
              make_float4(acc[0][n], acc[1][n], acc[2][n], acc[3][n]);
      }
    }
    return;
  }
  // ---- 3x3 conv from shared memory: LPP lanes per output pixel split the channel loop
  const int n_pix = TW * H;
  for (int p = threadIdx.x / LPP; p < n_pix; p += blockDim.x / LPP) {
    const int col = sh_h >= 0 ? p >> sh_h : p / H;
    const int h = p - col * H;
    const int w = w0 + col;
    float acc[COUT];
#pragma unroll
    for (int n = 0; n < COUT; ++n) acc[n] = 0.f;
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) {
        const int hj = h + j - 1;
        if (hj < 0 || hj >= H) continue;
        const float* xr = tile + ((col + i) * H + hj) * pitch;
        const float* wt = w_s + (i * 3 + j) * COUT * Cin;
#pragma unroll 4
        for (int c = sub * 4; c < Cin; c += 4 * LPP) {
          const float4 a = *reinterpret_cast<const float4*>(xr + c);
#pragma unroll
          for (int n = 0; n < COUT; ++n) {
            const float4 wv = *reinterpret_cast<const float4*>(wt + n * Cin + c);
            acc[n] = fmaf(a.x, wv.x, acc[n]); acc[n] = fmaf(a.y, wv.y, acc[n]);
            acc[n] = fmaf(a.z, wv.z, acc[n]); acc[n] = fmaf(a.w, wv.w, acc[n]);
          }
        }
      }
    }
#pragma unroll
    for (int n = 0; n < COUT; ++n) {
      for (int o = LPP >> 1; o > 0; o >>= 1) acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], o);
      if (sub == 0 && w < W) out[((static_cast<size_t>(b) * COUT + n) * W + w) * H + h] = acc[n] + (bias ? __ldg(bias + n) : 0.f);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// attention core, head_dim 8.  grid (ceil(N/128), C/8, B), block 128: thread = one query.
// K and V of the (b, head) are staged in shared memory in tiles of KT keys; every thread scans
// them with broadcast reads (two passes per tile are avoided by an online softmax).
constexpr int kAttnKT = 512;
__global__ void __launch_bounds__(128)
attention_kernel(const float* __restrict__ qkv, __half* __restrict__ out, __half* __restrict__ out_lo, int N,
                 int C, int H) {
  pdl_entry();
  __shared__ float4 sk[kAttnKT * 2 + 16];
  __shared__ float4 sv[kAttnKT * 2 + 16];
  const int b = blockIdx.z, hd = blockIdx.y;
  const int qi = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t row = 3 * static_cast<size_t>(C);
  const float* base = qkv + static_cast<size_t>(b) * N * row + hd * 8;
  // fold softmax scale 1/sqrt(8) and log2(e) into q so the inner loop is one FFMA chain + ex2
  const float qs = 0.35355339059327373f * 1.4426950408889634f;
  float q[8];
  {
    const int qq = min(qi, N - 1);
    const float4 a = __ldg(reinterpret_cast<const float4*>(base + qq * row));
    const float4 c = __ldg(reinterpret_cast<const float4*>(base + qq * row) + 1);
    q[0] = a.x * qs; q[1] = a.y * qs; q[2] = a.z * qs; q[3] = a.w * qs;
    q[4] = c.x * qs; q[5] = c.y * qs; q[6] = c.z * qs; q[7] = c.w * qs;
  }
  float mx = -INFINITY, l = 0.f;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int k0 = 0; k0 < N; k0 += kAttnKT) {
    const int kt = min(kAttnKT, N - k0);
    __syncthreads();
    for (int i = threadIdx.x; i < ((kt + 7) & ~7) * 2; i += blockDim.x) {
      const int kk = k0 + (i >> 1);
      const bool in = (i >> 1) < kt;   // pad the tile to a multiple of 8 keys with zeros (masked below)
      sk[i] = in ? __ldg(reinterpret_cast<const float4*>(base + kk * row + C) + (i & 1)) : make_float4(0, 0, 0, 0);
      sv[i] = in ? __ldg(reinterpret_cast<const float4*>(base + kk * row + 2 * C) + (i & 1)) : make_float4(0, 0, 0, 0);
    }
    __syncthreads();
    for (int j0 = 0; j0 < kt; j0 += 8) {
      float s[8];
      float cmx = mx;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 ka = sk[2 * (j0 + j)], kb = sk[2 * (j0 + j) + 1];
        float t = q[0] * ka.x;
        t = fmaf(q[1], ka.y, t); t = fmaf(q[2], ka.z, t); t = fmaf(q[3], ka.w, t);
        t = fmaf(q[4], kb.x, t); t = fmaf(q[5], kb.y, t); t = fmaf(q[6], kb.z, t);
        t = fmaf(q[7], kb.w, t);
        s[j] = (j0 + j < kt) ? t : -INFINITY;
        cmx = fmaxf(cmx, s[j]);
      }
      const float corr = exp2f(mx - cmx);   // mx=-inf on the first block -> 0
      mx = cmx;
      l *= corr;
#pragma unroll
      for (int d = 0; d < 8; ++d) acc[d] *= corr;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float pj = exp2f(s[j] - mx);
        l += pj;
        const float4 va = sv[2 * (j0 + j)], vb = sv[2 * (j0 + j) + 1];
        acc[0] = fmaf(pj, va.x, acc[0]); acc[1] = fmaf(pj, va.y, acc[1]);
        acc[2] = fmaf(pj, va.z, acc[2]); acc[3] = fmaf(pj, va.w, acc[3]);
        acc[4] = fmaf(pj, vb.x, acc[4]); acc[5] = fmaf(pj, vb.y, acc[5]);
        acc[6] = fmaf(pj, vb.z, acc[6]); acc[7] = fmaf(pj, vb.w, acc[7]);
      }
    }
  }
  if (qi < N) {
    const float inv = 1.0f / l;
    __align__(16) __half2 h[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(acc[2 * j] * inv, acc[2 * j + 1] * inv);
    // W-padded operand layout (B, W+2, H, C): token n = w*H + h lands at padded pixel H + n
    const size_t o = (static_cast<size_t>(b) * (N + 2 * H) + H + qi) * C + hd * 8;
    *reinterpret_cast<uint4*>(out + o) = *reinterpret_cast<const uint4*>(h);
    if (out_lo) {
      __align__(16) __half2 l[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 hf = __half22float2(h[j]);
        l[j] = __floats2half2_rn(acc[2 * j] * inv - hf.x, acc[2 * j + 1] * inv - hf.y);
      }
      *reinterpret_cast<uint4*>(out_lo + o) = *reinterpret_cast<const uint4*>(l);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// attention core on the legacy tensor path (mma.sync), short sequences (body: blocks.cuh).  grid (N/64, C/8, B), block 128.
__global__ void __launch_bounds__(128)
attention_tc_kernel(const float* __restrict__ qkv, __half* __restrict__ out, __half* __restrict__ out_lo,
                    int N, int C, int H) {
  pdl_entry();
  __shared__ __align__(16) __half sm[kAtSmemHalves];
  attention_tc_block<false>(qkv, out, out_lo, N, C, H, blockIdx.x, blockIdx.y, blockIdx.z, sm, threadIdx.x, 0);
}

// ------------------------------------------------------------------------------------------------
// time embedding: three launches of ONE small-GEMM kernel over R rows (R = batch for a single forward, R = number of
// sampling steps when a trajectory evaluates the embedding of its whole timestep table at once):
//   (1) sinusoid(t) -> linear_1 -> SiLU        K = D0,  N = D4
//   (2)             -> linear_2 -> SiLU        K = D4,  N = D4      (silu(emb) is what every resnet projects)
//   (3)             -> all time_emb_proj rows  K = D4,  N = T
// grid = ceil(N / 8), block 256: warp = one output column n (its weight row lives in registers, K/32 per lane), the
// R input rows are staged in shared memory in chunks; a dot product is a per-lane FMA chain + one butterfly.  Every
// weight is read exactly once per launch (9.7 MB for the C3 projections) by ~N/8 CTAs in parallel; the old kernels
// walked all rows of a matrix with ONE CTA per batch row (84 us for 8 rows).
constexpr int kTembMaxK = 1024;          // weight row in registers: K/32 <= 32 per lane
constexpr int kTembRowsChunk = 16;       // input rows staged per pass: 16 x 1024 floats = 64 KB max
__global__ void __launch_bounds__(256)
temb_linear_kernel(const float* __restrict__ in, const float* __restrict__ t, const float* __restrict__ w,
                   const float* __restrict__ bias, float* __restrict__ out, int R, int K, int N, int silu_out) {
  extern __shared__ float sh_t[];        // [rows chunk][K]
  pdl_entry();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * (blockDim.x >> 5) + warp;
  const int kpl = K >> 5;                // K % 32 == 0 (host)
  float wr[kTembMaxK / 32];
#pragma unroll
  for (int j = 0; j < kTembMaxK / 32; ++j) wr[j] = (j < kpl && n < N) ? __ldg(w + static_cast<size_t>(n) * K + lane + 32 * j) : 0.f;
  const float bn = (n < N && bias) ? __ldg(bias + n) : 0.f;
  for (int r0 = 0; r0 < R; r0 += kTembRowsChunk) {
    const int rows = min(kTembRowsChunk, R - r0);
    __syncthreads();
    if (t != nullptr) {                  // sinusoidal embedding of the timestep (flip_sin_to_cos: [cos | sin], divisor half)
      const int half = K >> 1;
      for (int i = threadIdx.x; i < rows * half; i += blockDim.x) {
        const int r = i / half, c = i - r * half;
        const float f = expf(-9.210340371976184f * static_cast<float>(c) / static_cast<float>(half));
        const float a = t[r0 + r] * f;
        sh_t[r * K + c] = cosf(a);
        sh_t[r * K + half + c] = sinf(a);
      }
    } else {
      for (int i = threadIdx.x; i < rows * K; i += blockDim.x) sh_t[i] = in[static_cast<size_t>(r0) * K + i];
    }
    __syncthreads();
    if (n < N) {
      for (int r = 0; r < rows; ++r) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < kTembMaxK / 32; ++j)
          if (j < kpl) s = fmaf(wr[j], sh_t[r * K + lane + 32 * j], s);
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) {
          s += bn;
          out[static_cast<size_t>(r0 + r) * N + n] = silu_out ? s / (1.0f + expf(-s)) : s;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// fused scheduler step (float4 vectorised; scalar tail).
__global__ void __launch_bounds__(256)
sched_step_kernel(const float* __restrict__ k, const float* x,
                  const float* __restrict__ eps, const float* x0_prev,
                  const float* __restrict__ noise, float* x_out,
                  float* x0_out, int64_t n) {
  // x / x_out and x0_prev / x0_out may alias (the trajectory program updates the latents and the previous x0 in
  // place: every thread reads its own elements before it writes them), so those four carry no __restrict__
  pdl_entry();
  const float k0 = k[0], k1 = k[1], k2 = k[2], k3 = k[3], k4 = k[4], k5 = k[5], k6 = k[6];
  const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  if (i + 4 <= n) {
    const float4 xv = *reinterpret_cast<const float4*>(x + i);
    const float4 ev = *reinterpret_cast<const float4*>(eps + i);
    const float4 pv = x0_prev ? *reinterpret_cast<const float4*>(x0_prev + i) : make_float4(0, 0, 0, 0);
    const float4 nv = noise ? *reinterpret_cast<const float4*>(noise + i) : make_float4(0, 0, 0, 0);
    float4 x0, xo;
#define RLDM_STEP(f)                                                            \
  x0.f = fmaf(k0, xv.f, k1 * ev.f);                                             \
  xo.f = k2 * xv.f + k3 * x0.f + k4 * pv.f + k5 * ev.f + k6 * nv.f;
    RLDM_STEP(x) RLDM_STEP(y) RLDM_STEP(z) RLDM_STEP(w)
#undef RLDM_STEP
    if (x0_out) *reinterpret_cast<float4*>(x0_out + i) = x0;
    *reinterpret_cast<float4*>(x_out + i) = xo;
  } else {
    for (int64_t j = i; j < n; ++j) {
      const float x0 = fmaf(k0, x[j], k1 * eps[j]);
      const float xo = k2 * x[j] + k3 * x0 + k4 * (x0_prev ? x0_prev[j] : 0.f) + k5 * eps[j] +
                       k6 * (noise ? noise[j] : 0.f);
      if (x0_out) x0_out[j] = x0;
      x_out[j] = xo;
    }
  }
}

// zero fill (replaces cudaMemsetAsync so the launch chain stays kernel -> kernel for PDL); n16 = 16-byte units
__global__ void zero_kernel(uint4* __restrict__ p, size_t n16) {
  pdl_entry();
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n16;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    p[i] = make_uint4(0, 0, 0, 0);
}
int zero_fill(void* p, size_t bytes, cudaStream_t st) {
  if (bytes % 16 != 0 || (reinterpret_cast<uintptr_t>(p) & 15) != 0) {
    cudaError_t e = cudaMemsetAsync(p, 0, bytes, st);
    if (e != cudaSuccess) { set_error("memset: %s", cudaGetErrorString(e)); return 2; }
    return 0;
  }
  const size_t n16 = bytes / 16;
  unsigned grid = static_cast<unsigned>((n16 + 255) / 256);
  if (grid > 1184) grid = 1184;
  if (grid == 0) return 0;
  RLDM_CUDA(launch_pdl_cls(8, zero_kernel, dim3(grid), dim3(256), 0, st, reinterpret_cast<uint4*>(p), n16));
  return 0;
}

__global__ void scale_kernel(const float* __restrict__ x, float a, float* __restrict__ y, int64_t n) {
  pdl_entry();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) y[i] = a * x[i];
}

// layout helpers: ref (B,C,W,H) <-> cl (B,W,H,C); thread per element (tiny boundary tensors only).
__global__ void ref_to_cl_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int P,
                                 size_t total) {
  pdl_entry();
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;  // dst index
  if (i >= total) return;
  const int c = i % C;
  const size_t bp = i / C;
  const int p = bp % P;
  const size_t b = bp / P;
  dst[i] = src[(b * C + c) * P + p];
}
__global__ void cl_to_ref_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int P,
                                 size_t total) {
  pdl_entry();
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;  // dst index
  if (i >= total) return;
  const int p = i % P;
  const size_t bc = i / P;
  const int c = bc % C;
  const size_t b = bc / C;
  dst[i] = src[(b * P + p) * C + c];
}

}  // namespace rldm

using namespace rldm;

extern "C" int rldm_version(void) { return RLDM_VERSION; }
extern "C" void rldm_reload_env(void) { load_env(); }
extern "C" const char* rldm_last_error(void) { return g_err; }

extern "C" int rldm_gn_stats(const float* x0, int c0, const float* x1, int c1, double* sums, int B,
                             int P, int G, void* stream) {
  const int C = c0 + c1;
  RLDM_CHECK(c0 % 4 == 0 && c1 % 4 == 0 && C % G == 0, "gn_stats: bad channels c0=%d c1=%d G=%d", c0, c1, G);
  RLDM_CHECK(x1 != nullptr || c1 == 0, "gn_stats: x1 NULL with c1=%d", c1);
  // ~4 CTAs per SM over the whole tensor, at least 16 pixels each
  int chunks = (592 + B - 1) / B;
  int ppb = (P + chunks - 1) / chunks;
  if (ppb < 16) ppb = 16;
  chunks = (P + ppb - 1) / ppb;
  const int q = C / 4;   // block = whole pixels so each thread keeps one channel quad in registers
  const int threads = q <= 256 ? (256 / q) * q : 256;
  RLDM_CUDA(launch_pdl(gn_stats_kernel, dim3(chunks, B), dim3(threads), 2 * G * sizeof(double), as_stream(stream), x0, c0, x1, c1, sums, P, G, ppb));
  RLDM_LAUNCH_CHECK();
  return 0;
}

extern "C" int rldm_prep(const float* x0, int c0, const float* x1, int c1, const double* sums,
                         const double* pairs0, const double* pairs1,
                         const float* gamma, const float* beta, float eps, int G, int silu, int up,
                         int circular, uint16_t* out, uint16_t* out_lo, uint16_t* raw, uint16_t* raw_lo, int B,
                         int W, int H, void* stream) {
  const int C = c0 + c1;
  RLDM_CHECK(c0 % 8 == 0 && c1 % 8 == 0, "prep: channels must be multiples of 8 (c0=%d c1=%d)", c0, c1);
  RLDM_CHECK(up == 1 || up == 2, "prep: up must be 1 or 2");
  RLDM_CHECK(!(sums || pairs0) || (gamma && beta && G > 0 && C % G == 0), "prep: GroupNorm needs gamma/beta/G");
  RLDM_CHECK(!pairs0 || ((C / G) % 2 == 0 && c0 % 2 == 0 && (c1 == 0 || pairs1)),
             "prep: channel-pair moments need an even group size and moments for both concat sources");
  const int out_pix = (W * up + 2) * H * up;
  int chunks = (592 + B - 1) / B;
  int ppb = (out_pix + chunks - 1) / chunks;
  if (ppb < 8) ppb = 8;
  chunks = (out_pix + ppb - 1) / ppb;
  // preps up to the top-level UNet tensors may start under the tail of the producing kernel (RLDM_PDL=2); the
  // full-resolution decoder passes measured slower with it
  PrepArgs a;
  a.x0 = x0; a.x1 = x1; a.sums = sums; a.pairs0 = pairs0; a.pairs1 = pairs1; a.gamma = gamma; a.beta = beta;
  a.out = reinterpret_cast<__half*>(out); a.out_lo = reinterpret_cast<__half*>(out_lo);
  a.raw = reinterpret_cast<__half*>(raw); a.raw_lo = reinterpret_cast<__half*>(raw_lo);
  a.eps = eps; a.c0 = c0; a.c1 = c1; a.G = G; a.silu = silu; a.up = up; a.circular = circular; a.W = W; a.H = H;
  a.pix_per_block = ppb;
  if (static_cast<size_t>(B) * out_pix * C <= env().prep_pdl_max)
    RLDM_CUDA(launch_pdl_small(prep_kernel, dim3(chunks, B), dim3(256), 2 * C * sizeof(float), as_stream(stream), a));
  else
    RLDM_CUDA(launch_pdl(prep_kernel, dim3(chunks, B), dim3(256), 2 * C * sizeof(float), as_stream(stream), a));
  RLDM_LAUNCH_CHECK();
  return 0;
}

extern "C" int rldm_conv_ref(const uint16_t* x, const uint16_t* x_lo, const uint16_t* wgt, const float* bias,
                             const float* temb, int temb_stride, const float* residual, float* out,
                             int B, int W, int H, int Cin, int Cout, int ks, int stride, int pad_lo,
                             int circular, void* stream) {
  const size_t total = static_cast<size_t>(B) * (W / stride) * (H / stride) * Cout;
  RLDM_CUDA(launch_pdl(conv_ref_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, as_stream(stream), reinterpret_cast<const __half*>(x), reinterpret_cast<const __half*>(x_lo),
      reinterpret_cast<const __half*>(wgt), bias, temb,
      temb_stride, residual, out, B, W, H, Cin, Cout, ks, stride, pad_lo, circular));
  RLDM_LAUNCH_CHECK();
  return 0;
}

static int conv_in_impl(const float* x0, int c0, const float* x1, int c1, const float* wgt, const float* bias, float* out,
                        int B, int W, int H, int Cout, int circular, double* stats, void* stream) {
  RLDM_CHECK(x1 != nullptr || c1 == 0, "conv_in: x1 NULL with c1=%d", c1);
  RLDM_CHECK(Cout % 4 == 0 && Cout <= 1024 && 256 % (Cout / 4) == 0, "conv_in: unsupported Cout=%d", Cout);
  RLDM_CHECK(!stats || (W * H) % kCinPix == 0, "conv_in: fused moments need W*H %% %d == 0 (got %d)", kCinPix, W * H);
  const size_t total_pix = static_cast<size_t>(B) * W * H;
  RLDM_CHECK(c0 + c1 >= 1 && c0 + c1 <= 16, "conv_in: 1..16 input channels (got %d + %d)", c0, c1);
  RLDM_CHECK(total_pix < (1ull << 31) && static_cast<size_t>(c0 + c1) * W * H < (1ull << 31), "conv_in: tensor too large for 32-bit pixel indices");
  const int K = 9 * (c0 + c1);
  const size_t smem = (static_cast<size_t>(K) * Cout + static_cast<size_t>(kCinPix) * K) * sizeof(float);
  RLDM_CHECK(smem <= 200 * 1024, "conv_in: 9*Cin*Cout too large for shared memory (Cin=%d Cout=%d)", c0 + c1, Cout);
  static size_t smem_set = 0;
  if (smem > 44 * 1024 && smem > smem_set) {
    RLDM_CUDA(cudaFuncSetAttribute(conv_in_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    smem_set = 200 * 1024;
  }
  // contiguous chunk ranges over ~4 blocks per SM (fewer when the weights fill the shared memory)
  const int n_chunks = static_cast<int>((total_pix + kCinPix - 1) / kCinPix);
  const int per_sm = smem > 100 * 1024 ? 1 : (smem > 50 * 1024 ? 2 : 4);
  int blocks = 148 * per_sm;
  if (blocks > n_chunks) blocks = n_chunks;
  const int cpb = (n_chunks + blocks - 1) / blocks;
  blocks = (n_chunks + cpb - 1) / cpb;
  RLDM_CUDA(launch_pdl_cls(4, conv_in_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), smem, as_stream(stream), x0, c0, x1, c1,
                       wgt, bias, out, B, W, H, Cout, circular, stats, cpb));
  RLDM_LAUNCH_CHECK();
  return 0;
}

extern "C" int rldm_conv_in(const float* x0, int c0, const float* x1, int c1, const float* wgt,
                            const float* bias, float* out, int B, int W, int H, int Cout,
                            int circular, void* stream) {
  return conv_in_impl(x0, c0, x1, c1, wgt, bias, out, B, W, H, Cout, circular, nullptr, stream);
}

extern "C" int rldm_conv_in_stats(const float* x0, int c0, const float* x1, int c1, const float* wgt,
                                  const float* bias, float* out, int B, int W, int H, int Cout,
                                  int circular, double* stats, void* stream) {
  return conv_in_impl(x0, c0, x1, c1, wgt, bias, out, B, W, H, Cout, circular, stats, stream);
}

extern "C" int rldm_conv_out(const uint16_t* x, const uint16_t* x_lo, const float* wgt, const float* bias, float* out,
                             int B, int W, int H, int Cin, int Cout, int circular, void* stream) {
  RLDM_CHECK(Cin % 8 == 0, "conv_out: Cin must be a multiple of 8");
  const size_t total_pix = static_cast<size_t>(B) * W * H;
  const __half* xh = reinterpret_cast<const __half*>(x);
  const __half* xl = reinterpret_cast<const __half*>(x_lo);
  cudaStream_t st = as_stream(stream);
  const size_t smem = static_cast<size_t>(9) * Cout * Cin * sizeof(float);
  RLDM_CHECK(smem <= 48 * 1024, "conv_out: 9*Cout*Cin weights exceed 48 KB of shared memory");
  const bool wide = Cin % 64 == 0;        // 8 lanes per pixel: each warp load touches 4 full 128 B lines
  const size_t groups = (total_pix * (wide ? 8 : 1) + 255) / 256;
  const unsigned grid = static_cast<unsigned>(groups < 1184 ? groups : 1184);
#define RLDM_CO(N)                                                                                              \
  if (wide) RLDM_CUDA(launch_pdl_cls(4, conv_out_kernel<N, 8>, dim3(grid), dim3(256), smem, st, xh, xl, wgt, bias, out, B, W, H, Cin, circular)); \
  else RLDM_CUDA(launch_pdl_cls(4, conv_out_kernel<N, 1>, dim3(grid), dim3(256), smem, st, xh, xl, wgt, bias, out, B, W, H, Cin, circular));
  switch (Cout) {
    case 2: RLDM_CO(2) break;
    case 4: RLDM_CO(4) break;
    case 8: RLDM_CO(8) break;
    default: RLDM_CHECK(false, "conv_out: Cout must be 2, 4 or 8 (got %d)", Cout);
  }
#undef RLDM_CO
  RLDM_LAUNCH_CHECK();
  return 0;
}

extern "C" int rldm_norm_conv_out(const float* x, const double* sums, const double* pairs, const float* gamma,
                                  const float* beta, float eps, int G, int silu, const float* wgt, const float* bias,
                                  float* out, int B, int W, int H, int Cin, int Cout, int circular, void* stream) {
  RLDM_CHECK(Cin % 4 == 0, "norm_conv_out: Cin must be a multiple of 4");
  RLDM_CHECK(!(sums || pairs) || (gamma && beta && G > 0 && Cin % G == 0), "norm_conv_out: GroupNorm needs gamma/beta/G");
  RLDM_CHECK(!pairs || (Cin / G) % 2 == 0, "norm_conv_out: channel-pair moments need an even group size");
  RLDM_CHECK(H >= 1 && H <= 256, "norm_conv_out: H out of range (%d)", H);
  // lanes per pixel: keep ~256 threads busy on a TW x H pixel tile
  // four pixels per thread (register-blocked inner loop) whenever H allows it; RLDM_NCO_PP1=1: one pixel per thread
  const int PP = (H % 4 == 0 && Cout <= 4 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && !env().nco_pp1) ? 4 : 1;
  int TW = 8, LPP = 1;
  size_t smem = 0;
  for (;; TW >>= 1) {
    LPP = 1;
    while (LPP < 8 && (TW * H / PP) * LPP * 2 <= 256 && Cin % (8 * LPP) == 0) LPP *= 2;
    smem = (2 * static_cast<size_t>(Cin) + 9 * static_cast<size_t>(Cout) * Cin +
            static_cast<size_t>(TW + 2) * H * (PP > 1 ? Cin + 4 : Cin + 4 * LPP)) * sizeof(float);
    if (smem <= 110 * 1024 || TW == 1) break;
  }
  RLDM_CHECK(smem <= 220 * 1024, "norm_conv_out: tile does not fit shared memory (H=%d Cin=%d Cout=%d)", H, Cin, Cout);
  cudaStream_t st = as_stream(stream);
  dim3 grid((W + TW - 1) / TW, B);
#define RLDM_NCO(N)                                                                                                   \
  {                                                                                                                   \
    static size_t smem_set = 0;                                                                                       \
    if (smem > 44 * 1024 && smem > smem_set) {                                                                        \
      RLDM_CUDA(cudaFuncSetAttribute(norm_conv_out_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); \
      smem_set = 220 * 1024;                                                                                          \
    }                                                                                                                 \
    RLDM_CUDA(launch_pdl_cls(4, norm_conv_out_kernel<N>, grid, dim3(256), smem, st, x, sums, pairs, gamma, beta, eps, G, silu, wgt, \
                         bias, out, W, H, Cin, circular, TW, LPP, PP));                                               \
  }
  switch (Cout) {
    case 2: RLDM_NCO(2) break;
    case 4: RLDM_NCO(4) break;
    case 8: RLDM_NCO(8) break;
    default: RLDM_CHECK(false, "norm_conv_out: Cout must be 2, 4 or 8 (got %d)", Cout);
  }
#undef RLDM_NCO
  RLDM_LAUNCH_CHECK();
  return 0;
}

int rldm_attention_umma(const float* qkv, uint16_t* out, uint16_t* out_lo, int B, int N, int C, int H, void* stream);

extern "C" int rldm_attention(const float* qkv, uint16_t* out, uint16_t* out_lo, int B, int N, int C,
                              int H, void* stream) {
  RLDM_CHECK(C % 8 == 0, "attention: C %% 8 != 0");
  if (!env().attn_mmasync && !env().attn_cudacore) {   // tcgen05 kernel: N a multiple of 128, >= 256
    const int rc = rldm_attention_umma(qkv, out, out_lo, B, N, C, H, stream);
    if (rc >= 0) return rc;
  }
  if (N % 64 == 0 && !env().attn_cudacore) {   // tensor-path kernel; the CUDA-core kernel covers ragged N
    RLDM_CUDA(launch_pdl_cls(2, attention_tc_kernel, dim3(N / 64, C / 8, B), dim3(128), 0, as_stream(stream), qkv, reinterpret_cast<__half*>(out), reinterpret_cast<__half*>(out_lo), N, C, H));
    RLDM_LAUNCH_CHECK();
    return 0;
  }
  const int threads = N >= 128 ? 128 : ((N + 31) / 32) * 32;
  RLDM_CUDA(launch_pdl(attention_kernel, dim3((N + threads - 1) / threads, C / 8, B), dim3(threads), 0, as_stream(stream), qkv, reinterpret_cast<__half*>(out), reinterpret_cast<__half*>(out_lo), N, C, H));
  RLDM_LAUNCH_CHECK();
  return 0;
}

static int temb_linear(const float* in, const float* t, const float* w, const float* b, float* out, int R, int K, int N,
                       int silu_out, cudaStream_t st) {
  RLDM_CHECK(K % 32 == 0 && K <= kTembMaxK, "temb: inner dimension %d must be a multiple of 32 and <= %d", K, kTembMaxK);
  const size_t smem = static_cast<size_t>(kTembRowsChunk) * K * sizeof(float);
  static size_t smem_set = 0;
  if (smem > 44 * 1024 && smem > smem_set) {
    RLDM_CUDA(cudaFuncSetAttribute(temb_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    smem_set = 64 * 1024;
  }
  RLDM_CUDA(launch_pdl_cls(8, temb_linear_kernel, dim3((N + 7) / 8), dim3(256), smem, st, in, t, w, b, out, R, K, N, silu_out));
  return 0;
}

extern "C" int rldm_temb(const float* t, const float* w1, const float* b1, const float* w2,
                         const float* b2, const float* wp, const float* bp, float* scratch,
                         float* out, int B, int D0, int D4, int T, void* stream) {
  // scratch: [2][B][D4] floats (hidden layer, then silu(emb))
  cudaStream_t st = as_stream(stream);
  if (B <= 0) return 0;
  float* h1 = scratch;
  float* semb = scratch + static_cast<size_t>(B) * D4;
  if (int rc = temb_linear(nullptr, t, w1, b1, h1, B, D0, D4, 1, st)) return rc;
  if (int rc = temb_linear(h1, nullptr, w2, b2, semb, B, D4, D4, 1, st)) return rc;
  if (T > 0)
    if (int rc = temb_linear(semb, nullptr, wp, bp, out, B, D4, T, 0, st)) return rc;
  RLDM_LAUNCH_CHECK();
  return 0;
}

extern "C" int rldm_sched_step(const float* k, const float* x, const float* eps,
                               const float* x0_prev, const float* noise, float* x_out,
                               float* x0_out, int64_t n, void* stream) {
  const int64_t nthreads = (n + 3) / 4;
  RLDM_CUDA(launch_pdl_cls(8, sched_step_kernel, dim3(static_cast<unsigned>((nthreads + 255) / 256)), dim3(256), 0, as_stream(stream), k, x, eps, x0_prev, noise, x_out, x0_out, n));
  RLDM_LAUNCH_CHECK();
  return 0;
}

extern "C" int rldm_scale(const float* x, float a, float* y, int64_t n, void* stream) {
  RLDM_CUDA(launch_pdl_cls(8, scale_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, as_stream(stream), x, a, y, n));
  RLDM_LAUNCH_CHECK();
  return 0;
}

extern "C" int rldm_ref_to_cl(const float* src, float* dst, int B, int C, int W, int H, void* stream) {
  const size_t total = static_cast<size_t>(B) * C * W * H;
  RLDM_CUDA(launch_pdl(ref_to_cl_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, as_stream(stream), src, dst, C, W * H, total));
  RLDM_LAUNCH_CHECK();
  return 0;
}
extern "C" int rldm_cl_to_ref(const float* src, float* dst, int B, int C, int W, int H, void* stream) {
  const size_t total = static_cast<size_t>(B) * C * W * H;
  RLDM_CUDA(launch_pdl(cl_to_ref_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, as_stream(stream), src, dst, C, W * H, total));
  RLDM_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// range image -> point cloud (`ldm/dataset.py:228-276`, point_cloud_to_range_image.to_pc_torch), the step the
// reference runs on every generated batch right after the decode (`ldm/inference.py:171`).
// img (B,C,W,H) fp32 ref layout -> points (B, W*H, 3|4) fp32: x, y, z [, remission]; optional depth (B, W*H) =
// |xyz| for the `depth < 90 m` mask of the .bin writer (`ldm/inference.py:176-178`).  Thread = one pixel
// (consecutive threads = consecutive beams h: coalesced reads, 16 B stores).  sin/cos of the 64 beam inclinations are
// staged in shared memory once per block; the azimuth sin/cos is computed per thread (precise sincosf).
namespace rldm {
__global__ void __launch_bounds__(256)
range_to_points_kernel(const float* __restrict__ img, int C, int W, int H, const float* __restrict__ incl,
                       const float* __restrict__ height, int mode, float mean, float stdv, float fill,
                       float* __restrict__ points, float* __restrict__ depth, size_t total) {
  extern __shared__ float sh_rp[];      // sin(incl)[H], cos(incl)[H], height[H]
  pdl_entry();
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    float sv, cv;
    sincosf(__ldg(incl + i), &sv, &cv);
    sh_rp[i] = sv; sh_rp[H + i] = cv; sh_rp[2 * H + i] = __ldg(height + i);
  }
  __syncthreads();
  const int P = C > 1 ? 4 : 3;
  for (size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int h = idx % H;
    const int w = (idx / H) % W;
    const size_t b = idx / (static_cast<size_t>(H) * W);
    const float v = __ldg(img + (b * C * W + w) * H + h);
    // every step is rounded separately (__f*_rn: no FMA contraction), like the op-by-op PyTorch reference -- the
    // azimuth `t * 2 pi - pi` cancels near zero, where a fused multiply-add would move points by tens of microns
    float r;
    if (mode == 1) r = __fsub_rn(exp2f(__fmul_rn(v, 6.0f)), 1.0f);      // log encoding      (`:241-242`)
    else if (mode == 2) r = __fdiv_rn(1.0f, fmaxf(v, 0.0001f));         // inverse encoding  (`:243-244`)
    else r = __fadd_rn(__fmul_rn(v, stdv), mean);                       // linear            (`:245-246`)
    if (r < 0.0f) r = fill;                                             // `:256`
    const float z = __fsub_rn(sh_rp[2 * H + h], __fmul_rn(r, sh_rp[h]));             // `:259`
    const float xy = __fmul_rn(r, sh_rp[H + h]);                                     // `:262`
    // azi = (W - 0.5 - w) / W * 2 pi - pi                                              `:266`, same operation order
    const float kPi = 3.14159265358979323846f;
    const float t = __fdiv_rn(__fsub_rn(__fsub_rn(static_cast<float>(W), 0.5f), static_cast<float>(w)), static_cast<float>(W));
    const float azi = __fsub_rn(__fmul_rn(__fmul_rn(t, 2.0f), kPi), kPi);
    float sa, ca;
    sincosf(azi, &sa, &ca);
    const float x = __fmul_rn(xy, ca), y = __fmul_rn(xy, sa);                        // `:269-270`
    if (P == 4) {
      const float rem = __ldg(img + ((b * C + 1) * W + w) * H + h);
      *reinterpret_cast<float4*>(points + idx * 4) = make_float4(x, y, z, rem);
    } else {
      points[idx * 3] = x; points[idx * 3 + 1] = y; points[idx * 3 + 2] = z;
    }
    if (depth) depth[idx] = sqrtf(x * x + y * y + z * z);
  }
}
}  // namespace rldm

// ------------------------------------------------------------------------------------------------
// point cloud -> bird's-eye-view volume (`ldm/dataset.py:278-294` to_voxel, `:13-132` _splat_points_to_volumes).
// splat: thread = one point; its 8 trilinear votes go to the density and feature volumes with float atomics
// (the reference's 16 scatter_add_ passes).  finalize: thread = one voxel, features / clamp(density), log(density + 1).
namespace rldm {
__global__ void __launch_bounds__(256)
voxel_splat_kernel(const float* __restrict__ points, int P, size_t total, int N, float cx, float cy, float cz, float hx,
                   float hy, float hz, int D, int Hh, int Ww, float* __restrict__ dens, float* __restrict__ feat) {
  pdl_entry();
  const size_t n_vox = static_cast<size_t>(D) * Hh * Ww;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t b = i / N;
    const float* pt = points + i * P;
    // local volume coordinates in [-1, 1] (`:282-283`), then continuous voxel indices (`:66-68`)
    const float fx = ((pt[0] - cx) / hx + 1.0f) * 0.5f * static_cast<float>(Ww - 1);
    const float fy = ((pt[1] - cy) / hy + 1.0f) * 0.5f * static_cast<float>(Hh - 1);
    const float fz = ((pt[2] - cz) / hz + 1.0f) * 0.5f * static_cast<float>(D - 1);
    const float f = P > 3 ? pt[3] : 0.0f;
    const float bx = floorf(fx), by = floorf(fy), bz = floorf(fz);
    const float rx = fx - bx, ry = fy - by, rz = fz - bz;
    // indices as 64-bit integers: far-away points (range fill value, inverse encoding) stay out of bounds
    const long long X = static_cast<long long>(bx), Y = static_cast<long long>(by), Z = static_cast<long long>(bz);
    float* db = dens + b * n_vox;
    float* fb = feat + b * n_vox;
#pragma unroll
    for (int xd = 0; xd < 2; ++xd) {
      const float wx = xd ? rx : 1.0f - rx;
      const long long X_ = X + xd;
#pragma unroll
      for (int yd = 0; yd < 2; ++yd) {
        const float wy = yd ? ry : 1.0f - ry;
        const long long Y_ = Y + yd;
#pragma unroll
        for (int zd = 0; zd < 2; ++zd) {
          const float wz = zd ? rz : 1.0f - rz;
          const long long Z_ = Z + zd;
          if (X_ < 0 || X_ >= Ww || Y_ < 0 || Y_ >= Hh || Z_ < 0 || Z_ >= D) continue;
          const float w = wx * wy * wz;
          const size_t idx = (static_cast<size_t>(Z_) * Hh + Y_) * Ww + X_;
          atomicAdd(db + idx, w);
          atomicAdd(fb + idx, w * f);
        }
      }
    }
  }
}
__global__ void __launch_bounds__(256)
voxel_finalize_kernel(const float* __restrict__ dens, const float* __restrict__ feat, float* __restrict__ voxel,
                      size_t n_vox, int B, int normalize, float min_weight) {
  pdl_entry();
  const size_t total = n_vox * B;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t b = i / n_vox, v = i - b * n_vox;
    const float d = dens[i];
    voxel[(2 * b) * n_vox + v] = normalize ? logf(d + 1.0f) : d;         // `:288-289`
    voxel[(2 * b + 1) * n_vox + v] = feat[i] / fmaxf(d, min_weight);       // `:126-128`
  }
}
}  // namespace rldm

extern "C" int rldm_points_to_voxel(const float* points, int B, int N, int P, const float* pc_range6, int D, int Hh, int Ww,
                                    int normalize, float* scratch, float* voxel, void* stream) {
  RLDM_CHECK(P == 3 || P == 4, "points_to_voxel: points must have 3 or 4 columns (got %d)", P);
  RLDM_CHECK(D >= 1 && Hh >= 1 && Ww >= 1, "points_to_voxel: bad grid");
  const size_t n_vox = static_cast<size_t>(D) * Hh * Ww;
  if (B == 0) return 0;
  cudaStream_t st = as_stream(stream);
  const int rc = zero_fill(scratch, 2 * n_vox * B * sizeof(float), st);
  if (rc) return rc;
  const float cx = (pc_range6[3] + pc_range6[0]) / 2, cy = (pc_range6[4] + pc_range6[1]) / 2, cz = (pc_range6[5] + pc_range6[2]) / 2;
  const float hx = (pc_range6[3] - pc_range6[0]) / 2, hy = (pc_range6[4] - pc_range6[1]) / 2, hz = (pc_range6[5] - pc_range6[2]) / 2;
  const size_t total = static_cast<size_t>(B) * N;
  if (total) {
    size_t blocks = (total + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    RLDM_CUDA(launch_pdl(voxel_splat_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, st, points, P, total, N, cx, cy,
                         cz, hx, hy, hz, D, Hh, Ww, scratch, scratch + n_vox * B));
  }
  size_t blocks = (n_vox * B + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  RLDM_CUDA(launch_pdl(voxel_finalize_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, st, scratch,
                       scratch + n_vox * B, voxel, n_vox, B, normalize, 1e-4f));
  RLDM_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// point cloud -> range image (`ldm/dataset.py:159-226`: projection, miss-value fill, normalisation).
// project: thread = one point: beam = argmin |incl_i - atan2(h_i - z, |xy|)| (`kitti360_range_image.py:51-61`), column
//   from the azimuth (round-half-even like np.round), range with the beam's height removed; the NEAREST return of a
//   pixel wins (the reference sorts by descending range and lets the last assignment win) = 64-bit atomicMin on
//   (range bits << 32 | point index).
// resolve: thread = one pixel: decode the winner, fill holes from the right neighbour (`fill_noise`), remaining holes
//   get the fill value, car-window mask from the 2-pixel neighbourhood, range encoding + normalisation, written in
//   the (C, W, H) layout of the dataset sample.
namespace rldm {
__global__ void __launch_bounds__(256)
range_project_kernel(const float* __restrict__ pc, int N, const float* __restrict__ incl, const float* __restrict__ height,
                     int H, int W, float fill_range, unsigned long long* __restrict__ keys) {
  extern __shared__ float sh_pr[];      // incl[H], height[H]
  pdl_entry();
  for (int i = threadIdx.x; i < H; i += blockDim.x) { sh_pr[i] = __ldg(incl + i); sh_pr[H + i] = __ldg(height + i); }
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    const float4 p = __ldg(reinterpret_cast<const float4*>(pc) + i);
    const float xy = sqrtf(__fadd_rn(__fmul_rn(p.x, p.x), __fmul_rn(p.y, p.y)));
    int row = 0;
    float best = INFINITY;
    for (int r = 0; r < H; ++r) {                       // first minimum, like np.argmin
      const float e = fabsf(__fsub_rn(sh_pr[r], atan2f(__fsub_rn(sh_pr[H + r], p.z), xy)));
      if (e < best) { best = e; row = r; }
    }
    const float azi = atan2f(p.y, p.x);
    const float kPi = 3.14159265358979323846f;
    // col = W - 1.0 + 0.5 - (azi + pi) / (2 pi) * W     (`:163`), rounded half to even (`np.round`)
    const float cf = __fsub_rn(__fadd_rn(__fsub_rn(static_cast<float>(W), 1.0f), 0.5f),
                               __fmul_rn(__fdiv_rn(__fadd_rn(azi, kPi), __fmul_rn(2.0f, kPi)), static_cast<float>(W)));
    int col = static_cast<int>(rintf(cf));
    if (col == W) col = W - 1;
    if (col < 0) col = 0;
    const float z = __fsub_rn(p.z, sh_pr[H + row]);                                      // `:168`
    float rng = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(p.x, p.x), __fmul_rn(p.y, p.y)), __fmul_rn(z, z)));
    if (rng > fill_range) rng = fill_range;                                              // `:170`
    const unsigned long long key = (static_cast<unsigned long long>(__float_as_uint(rng)) << 32) | static_cast<unsigned>(i);
    atomicMin(keys + static_cast<size_t>(row) * W + col, key);
  }
}
__device__ __forceinline__ float2 range_pixel(const unsigned long long* keys, const float* pc, int H, int W, int h, int w,
                                              int mode) {
  const unsigned long long k = keys[static_cast<size_t>(h) * W + w];
  if (k == ~0ull) return make_float2(-1.0f, -1.0f);
  const float rng = __uint_as_float(static_cast<unsigned>(k >> 32));
  const float rem = __ldg(pc + static_cast<size_t>(static_cast<unsigned>(k)) * 4 + 3);
  float v = rng;
  if (mode == 1) v = __fdiv_rn(log2f(__fadd_rn(rng, 1.0f)), 6.0f);
  else if (mode == 2) v = __fdiv_rn(1.0f, rng);
  return make_float2(v, rem);
}
// value after `fill_noise` (`:186-190`): a hole takes its right neighbour (circular in w), from the ORIGINAL image
__device__ __forceinline__ float2 range_filled(const unsigned long long* keys, const float* pc, int H, int W, int h, int w,
                                               int mode) {
  const float2 v = range_pixel(keys, pc, H, W, h, w, mode);
  if (v.x != -1.0f) return v;
  return range_pixel(keys, pc, H, W, h, w + 1 == W ? 0 : w + 1, mode);
}
__global__ void __launch_bounds__(256)
range_resolve_kernel(const unsigned long long* __restrict__ keys, const float* __restrict__ pc, int H, int W, int mode,
                     float mean, float stdv, float fill_range, float fill_rem, float* __restrict__ image,
                     unsigned char* __restrict__ mask, unsigned char* __restrict__ car) {
  pdl_entry();
  const int total = H * W;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int h = i % H, w = i / H;                      // output layout (C, W, H): consecutive threads = consecutive h
    const float2 orig = range_pixel(keys, pc, H, W, h, w, mode);
    float2 v = orig;
    bool m = orig.x > 0.0f;                              // `:194` range_image_mask
    if (orig.x == -1.0f) {
      v = range_pixel(keys, pc, H, W, h, w + 1 == W ? 0 : w + 1, mode);
      m = v.x > 0.0f;
    }
    const bool still = v.x == -1.0f;                     // `:201`
    bool cw = false;
    if (still) {                                         // `:203-209`: any filled value two pixels away
      const float d = range_filled(keys, pc, H, W, (h + H - 2) % H, w, mode).x;
      const float t = range_filled(keys, pc, H, W, (h + 2) % H, w, mode).x;
      const float r = range_filled(keys, pc, H, W, h, (w + W - 2) % W, mode).x;
      const float l = range_filled(keys, pc, H, W, h, (w + 2) % W, mode).x;
      cw = d != -1.0f || t != -1.0f || r != -1.0f || l != -1.0f;
      if (mode == 1) v = make_float2(__fdiv_rn(log2f(__fadd_rn(fill_range, 1.0f)), 6.0f), __fdiv_rn(log2f(__fadd_rn(fill_rem, 1.0f)), 6.0f));
      else if (mode == 2) v = make_float2(__fdiv_rn(1.0f, fill_range), fill_rem);
      else v = make_float2(fill_range, fill_rem);
    }
    if (mode == 0) v.x = __fdiv_rn(__fsub_rn(v.x, mean), stdv);                        // `:223-226`
    image[static_cast<size_t>(w) * H + h] = v.x;
    image[static_cast<size_t>(total) + static_cast<size_t>(w) * H + h] = v.y;
    mask[static_cast<size_t>(w) * H + h] = m ? 1 : 0;
    car[static_cast<size_t>(w) * H + h] = cw ? 1 : 0;
  }
}
__global__ void fill_keys_kernel(unsigned long long* keys, int n) {
  pdl_entry();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) keys[i] = ~0ull;
}
}  // namespace rldm

extern "C" int rldm_points_to_range(const float* pc, int N, const float* incl, const float* height, int H, int W, int mode,
                                    float mean, float stdv, float fill_range, float fill_rem, unsigned long long* keys,
                                    float* image, unsigned char* mask, unsigned char* car_window, void* stream) {
  RLDM_CHECK(H >= 3 && H <= 1024 && W >= 3, "points_to_range: bad image size (H=%d W=%d)", H, W);
  RLDM_CHECK(mode >= 0 && mode <= 2, "points_to_range: mode must be 0 (linear), 1 (log) or 2 (inverse)");
  RLDM_CHECK((reinterpret_cast<uintptr_t>(pc) & 15) == 0, "points_to_range: points must be 16 B aligned (N x 4 fp32)");
  cudaStream_t st = as_stream(stream);
  const int total = H * W;
  RLDM_CUDA(launch_pdl(fill_keys_kernel, dim3((total + 255) / 256), dim3(256), 0, st, keys, total));
  if (N > 0) {
    int blocks = (N + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    RLDM_CUDA(launch_pdl(range_project_kernel, dim3(blocks), dim3(256), 2 * H * sizeof(float), st, pc, N, incl, height, H, W,
                         fill_range, keys));
  }
  RLDM_CUDA(launch_pdl(range_resolve_kernel, dim3((total + 255) / 256), dim3(256), 0, st, keys, pc, H, W, mode, mean, stdv,
                       fill_range, fill_rem, image, mask, car_window));
  RLDM_LAUNCH_CHECK();
  return 0;
}

extern "C" int rldm_range_to_points(const float* img, int B, int C, int W, int H, const float* incl, const float* height,
                                    int mode, float mean, float stdv, float fill, float* points, float* depth,
                                    void* stream) {
  RLDM_CHECK(C >= 1 && H >= 1 && H <= 1024, "range_to_points: bad shape (C=%d H=%d)", C, H);
  RLDM_CHECK(mode >= 0 && mode <= 2, "range_to_points: mode must be 0 (linear), 1 (log) or 2 (inverse)");
  const size_t total = static_cast<size_t>(B) * W * H;
  if (total == 0) return 0;
  size_t blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  RLDM_CUDA(launch_pdl(range_to_points_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 3 * H * sizeof(float),
                       as_stream(stream), img, C, W, H, incl, height, mode, mean, stdv, fill, points, depth, total));
  RLDM_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
namespace rldm {
__global__ void stamp_kernel(unsigned long long* slot) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  *slot = t;
}
}  // namespace rldm

static int run_ops(const rldm_op* ops, int n_ops, unsigned long long* stamps, void* stream) {
  if (stamps) {
    stamp_kernel<<<1, 1, 0, as_stream(stream)>>>(stamps);
    RLDM_LAUNCH_CHECK();
  }
  for (int k = 0; k < n_ops; ++k) {
    const rldm_op& o = ops[k];
    int rc = 0;
    switch (o.kind) {
      case RLDM_OP_GN_STATS:
        rc = rldm_gn_stats((const float*)o.p[0], o.i[0], (const float*)o.p[1], o.i[1], (double*)o.p[2],
                           o.i[2], o.i[3], o.i[4], stream);
        break;
      case RLDM_OP_PREP:
        rc = rldm_prep((const float*)o.p[0], o.i[0], (const float*)o.p[1], o.i[1], (const double*)o.p[2],
                       (const double*)o.p[9], (const double*)o.p[10], (const float*)o.p[3], (const float*)o.p[4], o.f[0], o.i[2], o.i[3], o.i[4], o.i[8],
                       (uint16_t*)o.p[5], (uint16_t*)o.p[6], (uint16_t*)o.p[7], (uint16_t*)o.p[8], o.i[5], o.i[6], o.i[7],
                       stream);
        break;
      case RLDM_OP_CONV_TC:
        if (o.p[11] || o.p[17]) {       // the convolution produces its own operand (no rldm_prep in front of it)
          rldm_conv_src ms, ss;
          ms.x0 = (const float*)o.p[11]; ms.x1 = (const float*)o.p[12]; ms.pairs0 = (const double*)o.p[13];
          ms.pairs1 = (const double*)o.p[14]; ms.gamma = (const float*)o.p[15]; ms.beta = (const float*)o.p[16];
          ms.eps = o.f[0]; ms.c0 = o.i[13]; ms.c1 = o.i[14]; ms.G = o.i[15]; ms.silu = o.i[16]; ms.up = o.i[17]; ms.circular = o.i[9];
          ss.x0 = (const float*)o.p[17]; ss.x1 = (const float*)o.p[18]; ss.pairs0 = ss.pairs1 = nullptr; ss.gamma = ss.beta = nullptr;
          ss.eps = 0.f; ss.c0 = o.i[18]; ss.c1 = o.i[19]; ss.G = 0; ss.silu = 0; ss.up = 1; ss.circular = o.i[9];
          rc = rldm_conv_tc_fused(o.p[11] ? &ms : nullptr, o.p[17] ? &ss : nullptr, (const uint16_t*)o.p[0], (const uint16_t*)o.p[6],
                                  (const uint16_t*)o.p[1], (const float*)o.p[2], (const float*)o.p[3], o.i[0], (const float*)o.p[4],
                                  (float*)o.p[5], o.i[1], o.i[2], o.i[3], o.i[4], o.i[5], o.i[6], o.i[7], o.i[8], o.i[9], o.i[10],
                                  (double*)o.p[7], (const uint16_t*)o.p[8], (const uint16_t*)o.p[9], (const uint16_t*)o.p[10],
                                  o.i[11], o.i[12], stream);
          break;
        }
        if (o.p[19]) {                  // the epilogue also emits the next GroupNorm's operand (p[19..21], i[20..22], f[1])
          rldm_conv_emit em;
          em.out = (uint16_t*)o.p[19]; em.gamma = (const float*)o.p[20]; em.beta = (const float*)o.p[21];
          em.eps = o.f[1]; em.G = o.i[20]; em.silu = o.i[21]; em.circular = o.i[22];
          rc = rldm_conv_tc_emit(&em, (const uint16_t*)o.p[0], (const uint16_t*)o.p[6], (const uint16_t*)o.p[1],
                                 (const float*)o.p[2], (const float*)o.p[3], o.i[0], (const float*)o.p[4],
                                 (float*)o.p[5], o.i[1], o.i[2], o.i[3], o.i[4], o.i[5], o.i[6], o.i[7], o.i[8], o.i[9],
                                 o.i[10], (double*)o.p[7], (const uint16_t*)o.p[8], (const uint16_t*)o.p[9],
                                 (const uint16_t*)o.p[10], o.i[11], o.i[12], stream);
          break;
        }
        rc = rldm_conv_tc_ex((const uint16_t*)o.p[0], (const uint16_t*)o.p[6], (const uint16_t*)o.p[1],
                             (const float*)o.p[2], (const float*)o.p[3], o.i[0], (const float*)o.p[4],
                             (float*)o.p[5], o.i[1], o.i[2], o.i[3], o.i[4], o.i[5], o.i[6], o.i[7], o.i[8], o.i[9],
                             o.i[10], (double*)o.p[7], (const uint16_t*)o.p[8], (const uint16_t*)o.p[9],
                             (const uint16_t*)o.p[10], o.i[11], o.i[12], stream);
        break;
      case RLDM_OP_CONV_UP2:
        rc = rldm_conv_tc_up2((const uint16_t*)o.p[0], (const uint16_t*)o.p[1], (const uint16_t*)o.p[2], (const float*)o.p[3],
                              (float*)o.p[4], o.i[0], o.i[1], o.i[2], o.i[3], o.i[4], o.i[5], (double*)o.p[5], o.i[6], stream);
        break;
      case RLDM_OP_CONV_REF:
        rc = rldm_conv_ref((const uint16_t*)o.p[0], (const uint16_t*)o.p[6], (const uint16_t*)o.p[1], (const float*)o.p[2],
                           (const float*)o.p[3], o.i[0], (const float*)o.p[4], (float*)o.p[5], o.i[1],
                           o.i[2], o.i[3], o.i[4], o.i[5], o.i[6], o.i[7], o.i[8], o.i[9], stream);
        break;
      case RLDM_OP_CONV_IN:
        rc = rldm_conv_in_stats((const float*)o.p[0], o.i[0], (const float*)o.p[1], o.i[1], (const float*)o.p[2],
                                (const float*)o.p[3], (float*)o.p[4], o.i[2], o.i[3], o.i[4], o.i[5], o.i[6],
                                (double*)o.p[5], stream);
        break;
      case RLDM_OP_CONV_OUT:
        rc = rldm_conv_out((const uint16_t*)o.p[0], (const uint16_t*)o.p[4], (const float*)o.p[1], (const float*)o.p[2],
                           (float*)o.p[3], o.i[0], o.i[1], o.i[2], o.i[3], o.i[4], o.i[5], stream);
        break;
      case RLDM_OP_NORM_CONV_OUT:
        rc = rldm_norm_conv_out((const float*)o.p[0], (const double*)o.p[1], (const double*)o.p[2], (const float*)o.p[3],
                                (const float*)o.p[4], o.f[0], o.i[0], o.i[1], (const float*)o.p[5], (const float*)o.p[6],
                                (float*)o.p[7], o.i[2], o.i[3], o.i[4], o.i[5], o.i[6], o.i[7], stream);
        break;
      case RLDM_OP_ATTENTION:
        rc = rldm_attention((const float*)o.p[0], (uint16_t*)o.p[1], (uint16_t*)o.p[2], o.i[0], o.i[1], o.i[2], o.i[3], stream);
        break;
      case RLDM_OP_TEMB:
        rc = rldm_temb((const float*)o.p[0], (const float*)o.p[1], (const float*)o.p[2],
                       (const float*)o.p[3], (const float*)o.p[4], (const float*)o.p[5],
                       (const float*)o.p[6], (float*)o.p[7], (float*)o.p[8], o.i[0], o.i[1], o.i[2],
                       o.i[3], stream);
        break;
      case RLDM_OP_SCHED_STEP:
        rc = rldm_sched_step((const float*)o.p[0], (const float*)o.p[1], (const float*)o.p[2],
                             (const float*)o.p[3], (const float*)o.p[4], (float*)o.p[5], (float*)o.p[6],
                             o.n, stream);
        break;
      case RLDM_OP_MEMSET:
        rc = zero_fill(o.p[0], static_cast<size_t>(o.n), as_stream(stream));
        break;
      case RLDM_OP_AXPY:
        rc = rldm_scale((const float*)o.p[0], o.f[0], (float*)o.p[1], o.n, stream);
        break;
      case RLDM_OP_FUSED:
        rc = rldm_fused_run((rldm_fused*)o.p[0], stream);
        break;
      default:
        set_error("rldm_run: unknown op kind %d at index %d", o.kind, k);
        rc = 3;
    }
    if (rc) return rc;
    if (stamps) {     // plain (non-PDL) launch: starts only when op k has completed and flushed
      stamp_kernel<<<1, 1, 0, as_stream(stream)>>>(stamps + k + 1);
      RLDM_LAUNCH_CHECK();
    }
  }
  return 0;
}

extern "C" int rldm_run(const rldm_op* ops, int n_ops, void* stream) { return run_ops(ops, n_ops, nullptr, stream); }

extern "C" int rldm_run_timed(const rldm_op* ops, int n_ops, unsigned long long* stamps, void* stream) {
  RLDM_CHECK(stamps != nullptr, "rldm_run_timed: stamps must hold n_ops + 1 device uint64 slots");
  return run_ops(ops, n_ops, stamps, stream);
}
