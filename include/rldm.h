/* rldm.h -- C ABI of librldm.so, the sm_100a kernel library behind rangeldm_b200.
 *
 * Drop-in boundary.  The reference (WoodwindHu/RangeLDM) has no FFI of its own: its hot path is
 * Python calling third-party `diffusers` modules.  These entry points are what the
 * replacement classes (`rangeldm_b200.UNet2DModel`, `AutoencoderKL`, the schedulers and the
 * pipelines) bind through ctypes; each cites the reference arithmetic it replaces.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer borrowed from a live torch.Tensor for the duration of the
 *    call; nothing is allocated, freed or synchronised inside a call;
 *  - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); all launches are
 *    stream-ordered and CUDA-graph capturable;
 *  - return value 0 = success, non-zero = error; `rldm_last_error()` returns a thread-local,
 *    NUL-terminated description of the last failure;
 *  - "cl" (channels-last) activations are (B, W, H, C) row-major: C contiguous, then H (beams,
 *    zero padded), then W (azimuth, CIRCULAR), then B.  The reference layout "ref" is
 *    (B, C, W, H) (`ldm/dataset.py:230,330`);
 *  - half = IEEE fp16 (uint16_t storage).  Convolutions multiply fp16 operands on tcgen05 tensor
 *    cores and accumulate in fp32 (TMEM); everything else is fp32.
 */
#ifndef RLDM_H_
#define RLDM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RLDM_VERSION 100

int rldm_version(void);
const char* rldm_last_error(void);
/* The RLDM_* environment switches are read once at first use; this re-reads them. */
void rldm_reload_env(void);

/* ---- GroupNorm statistics ------------------------------------------------------------------
 * Replaces the reduction half of F.group_norm in ResnetBlock2D.norm1/norm2, Attention.group_norm,
 * conv_norm_out (diffusers) == `Normalize` (`vae/sgm/.../model.py:59-62`).
 * x0:(B,P,C0) [+ x1:(B,P,C1), a virtual channel concat == torch.cat([h, skip], 1)] fp32 cl.
 * Accumulates (sum, sum of squares) per (b, group) into `sums`[B][G][2] (double), which the caller
 * must have zeroed.  G groups over C0+C1 channels. */
int rldm_gn_stats(const float* x0, int c0, const float* x1, int c1, double* sums,
                  int B, int P, int G, void* stream);

/* ---- prep: (GroupNorm-apply) (+SiLU) (+concat) (+nearest 2x upsample) -> fp16 cl -----------
 * Replaces F.group_norm's normalise half + F.silu (`model.py:343-345,351-352`), torch.cat of the
 * skip connection (UpBlock2D) and F.interpolate(scale_factor=2, "nearest") (`model.py:121-122`).
 * x0:(B,W,H,C0) [+x1:(B,W,H,C1)] fp32 cl.  GroupNorm moments come either as `sums` [B][G][2] (from
 * rldm_gn_stats) or as channel-pair moments `pairs0` [B][C0/2][2] (+ `pairs1` [B][C1/2][2]) accumulated by
 * the producing rldm_conv_tc epilogues; all NULL -> no normalisation (raw cast).
 * out: the tensor-core OPERAND layout "clp": (B, W*up + 2, H*up, C0+C1) fp16, channels-last and
 * W-PADDED -- padded column wp holds image column (wp-1) mod W*up, so the circular halo of
 * `ldm/utils.py:47` (F.pad(..., mode="circular")) is materialised by the producer for free
 * (circular=0: zero halo) and every conv tap is one plain TMA box.  up in {1,2}.
 * out_lo (optional, same shape): the residual y - fp16(y), so that out + out_lo carries ~22
 * significant bits (split-fp16 operand).  raw / raw_lo (optional, same shape): a second output
 * holding the UN-normalised, un-activated input (the operand of a ResnetBlock2D's 1x1
 * conv_shortcut), produced from the same read. */
int rldm_prep(const float* x0, int c0, const float* x1, int c1, const double* sums,
              const double* pairs0, const double* pairs1,
              const float* gamma, const float* beta, float eps, int G, int silu, int up,
              int circular, uint16_t* out, uint16_t* out_lo, uint16_t* raw, uint16_t* raw_lo, int B,
              int W, int H, void* stream);

/* ---- the hot op: circular implicit-GEMM convolution on tcgen05 ------------------------------
 * Replaces `Conv2d._conv_forward` (`ldm/utils.py:40-58`, twin `vae/sgm/.../model.py:93-108`):
 * wrap-pad W, zero-pad H, F.conv2d(pad 0); and the nn.Linear projections of Attention (ks=1).
 *   x   : (B, W+2, H, Cin) fp16 clp (W-padded operand layout written by rldm_prep / rldm_attention),
 *         Cin % 64 == 0
 *   x_lo: NULL -> plain fp16 operands (1 MMA per K step, 11-bit significands);
 *         else the low-order half of a split-fp16 activation (see rldm_prep) and `wgt` must hold
 *         TWO planes [hi|lo]: the kernel issues Ah*Wh + Al*Wh + Ah*Wl into one fp32 accumulator
 *         ("fp16x3", ~fp32-accurate products; the default of the engine)
 *   wgt : [planes][ks*ks][Cout][Cin] fp16, tap = index_along_W * ks + index_along_H
 *         (torch weight (Cout,Cin,kh,kw): kh <-> W, kw <-> H)
 *   out : (B, Wo, Ho, Cout) fp32 cl;  Wo = W/stride, Ho = H/stride
 *   pad_lo: taps read input (stride*wo + i - pad_lo) mod W, stride*ho + j - pad_lo (0 outside H).
 *         ks=3,pad_lo=1: the symmetric circular conv; ks=3,stride=2,pad_lo=0: the VAE-encoder
 *         Downsample2D(padding=0) asymmetric pad (`ldm/utils.py:109-111`).
 *   circular: informational (1 = the producer wrote wrap halos, 0 = zero halos); the kernel reads
 *         whatever the halo columns of `x` hold.
 *   epilogue: out = acc + bias[c] + temb[b*temb_stride + c] + residual[b,wo,ho,c] (NULL = skip)
 *   split_k: 0 = choose automatically; 2, 4 or 8: the K loop (taps x channel chunks) is split over a
 *         thread-block CLUSTER of split_k CTAs per tile; partial tiles are reduced through distributed
 *         shared memory in a fixed order (deterministic; no atomics, no zero fill).
 *   stats: optional fused GroupNorm statistics of the finished output: (sum, sum of squares) per
 *         (image, channel PAIR) are ADDED into stats[B][Cout/2][2] (double, zeroed by the caller).
 *         rldm_prep folds pairs into the groups of the consuming GroupNorm (also across a skip concat),
 *         which replaces the rldm_gn_stats pass.  Needs Wo*Ho >= 64.
 * Requires Ho a power of two <= 128, Wo*Ho a multiple or a divisor of 128, Cout % 64 == 0. */
int rldm_conv_tc(const uint16_t* x, const uint16_t* x_lo, const uint16_t* wgt, const float* bias, const float* temb,
                 int temb_stride, const float* residual, float* out, int B, int W, int H, int Cin,
                 int Cout, int ks, int stride, int pad_lo, int circular, int split_k, double* stats,
                 void* stream);

/* rldm_conv_tc with ResnetBlock2D's 1x1 `conv_shortcut` (`model.py:356-360`; diffusers ResnetBlock2D) folded into
 * the K loop: out = conv(x; wgt) + conv1x1(sc_x; sc_wgt) + bias [+ temb] [+ residual].  sc_x (+ sc_x_lo):
 * (B, W+2, H, sc_cin) fp16 clp on the same grid as the output (stride must be 1), sc_wgt: [planes][Cout][sc_cin]
 * fp16; `bias` must already hold conv.bias + conv_shortcut.bias.  Saves one launch and the fp32 round trip of the
 * shortcut branch per block. */
int rldm_conv_tc_shortcut(const uint16_t* x, const uint16_t* x_lo, const uint16_t* wgt, const float* bias,
                          const float* temb, int temb_stride, const float* residual, float* out, int B, int W, int H,
                          int Cin, int Cout, int ks, int stride, int pad_lo, int circular, int split_k, double* stats,
                          const uint16_t* sc_x, const uint16_t* sc_x_lo, const uint16_t* sc_wgt, int sc_cin,
                          void* stream);

/* rldm_conv_tc_shortcut with an explicit operand precision (`terms`; the engine chooses it per layer):
 *   3: split-fp16 x3 -- x + x_lo and two weight planes [hi|lo]: Xh*Wh + Xl*Wh + Xh*Wl (~22-bit operands);
 *   2: activations single fp16 (x_lo ignored), two weight planes: Xh*Wh + Xh*Wl -- no low-order activation plane is
 *      written or read (half the activation operand bytes), the weights keep ~22 bits;
 *   1: plain fp16, one weight plane;
 *   0: infer from x_lo (3 when given, else 1), like rldm_conv_tc / rldm_conv_tc_shortcut. */
int rldm_conv_tc_ex(const uint16_t* x, const uint16_t* x_lo, const uint16_t* wgt, const float* bias, const float* temb,
                    int temb_stride, const float* residual, float* out, int B, int W, int H, int Cin, int Cout, int ks,
                    int stride, int pad_lo, int circular, int split_k, double* stats, const uint16_t* sc_x,
                    const uint16_t* sc_x_lo, const uint16_t* sc_wgt, int sc_cin, int terms, void* stream);

/* rldm_conv_tc_ex that PRODUCES ITS OWN OPERAND: instead of a rldm_prep launch in front of the convolution, every CTA of
 * the small-layer kernel turns the fp32 source into the part of the fp16 operand it is about to read (its tile's input
 * window, halo columns included, x the channel chunks of its K slice), writes it to `x` (+ `x_lo`), which are then
 * caller-provided SCRATCH buffers of the operand's shape (B, W+2, H, Cin), and loads it back by TMA.  Same arithmetic
 * as rldm_prep followed by rldm_conv_tc_ex (bit-identical operands).
 *   src    : source of the convolution's operand: concat(x0 (B,W/up,H/up,c0), x1 (.., c1)) fp32 cl, optional GroupNorm
 *            (channel-pair moments pairs0/pairs1 as accumulated by the producing epilogues, gamma, beta, eps, G groups),
 *            optional SiLU, nearest upsampling up in {1,2}, circular / zero halo
 *   sc_src : source of the folded 1x1 shortcut's operand (raw cast: no norm), written to sc_x (+ sc_x_lo); or NULL
 * Only layers that run on the small-layer kernel qualify (rldm_conv_tc_fusable() == 1: fewer 128 x 128 tiles than SMs,
 * Cin and sc_cin <= 512); others return an error. */
typedef struct rldm_conv_src {
  const float* x0; const float* x1;
  const double* pairs0; const double* pairs1;
  const float* gamma; const float* beta;
  float eps;
  int c0, c1, G, silu, up, circular;
} rldm_conv_src;
int rldm_conv_tc_fused(const rldm_conv_src* src, const rldm_conv_src* sc_src, const uint16_t* x, const uint16_t* x_lo,
                       const uint16_t* wgt, const float* bias, const float* temb, int temb_stride, const float* residual,
                       float* out, int B, int W, int H, int Cin, int Cout, int ks, int stride, int pad_lo, int circular,
                       int split_k, double* stats, const uint16_t* sc_x, const uint16_t* sc_x_lo, const uint16_t* sc_wgt,
                       int sc_cin, int terms, void* stream);
int rldm_conv_tc_fusable(int B, int W, int H, int Cin, int Cout, int ks, int stride, int pad_lo, int sc_cin, int has_residual);

/* rldm_conv_tc_ex that also EMITS THE NEXT GroupNorm's OPERAND: where the consumer of this convolution's output is
 * F.group_norm (+ F.silu) in front of another convolution (ResnetBlock2D norm2 -> conv2, the GroupNorm of the next
 * ResnetBlock2D / Attention; diffusers), the epilogue normalises its own output and writes the
 * fp16 W-padded operand (B, W/stride + 2, H/stride, Cout) that a rldm_prep launch would have produced -- the
 * (image, group) moments are completed inside a thread-block cluster that covers whole images (DSMEM), so no second
 * launch has to wait for the whole grid.  `out` may be NULL when nothing else reads the fp32 result; `stats` still
 * accumulates the channel-pair moments when given.  One fp16 plane only (consumers running below split-fp16 x3).
 * Only layers for which rldm_conv_tc_emittable() == 1 qualify: small-layer kernel, 64..1024 output pixels per image,
 * group size a power of two >= 4 that divides the output-channel tile. */
typedef struct rldm_conv_emit {
  uint16_t* out;                 /* (B, Wo+2, Ho, Cout) fp16 */
  const float* gamma; const float* beta;
  float eps;
  int G, silu, circular;         /* groups of the consumer's GroupNorm; SiLU after it; halo columns wrap (1) or zero (0) */
} rldm_conv_emit;
int rldm_conv_tc_emit(const rldm_conv_emit* emit, const uint16_t* x, const uint16_t* x_lo, const uint16_t* wgt,
                      const float* bias, const float* temb, int temb_stride, const float* residual, float* out,
                      int B, int W, int H, int Cin, int Cout, int ks, int stride, int pad_lo, int circular,
                      int split_k, double* stats, const uint16_t* sc_x, const uint16_t* sc_x_lo,
                      const uint16_t* sc_wgt, int sc_cin, int terms, void* stream);
int rldm_conv_tc_emittable(int B, int W, int H, int Cin, int Cout, int ks, int stride, int pad_lo, int sc_cin,
                           int has_residual, int G);

/* Upsample2D (`vae/sgm/modules/diffusionmodules/model.py:120-125`; diffusers Upsample2D): F.interpolate(nearest, 2x)
 * followed by the 3x3 convolution, with the upsampling FOLDED into the convolution: every output pixel (2w+a, 2h+b)
 * sees a 2x2 neighbourhood of the low-resolution input, so the layer is four 2x2 "phase" convolutions over the
 * low-resolution operand (4/9 of the multiply-adds, a quarter of the operand bytes), each writing one pixel phase.
 *   x, x_lo : (B, W+2, H, Cin) fp16 operand of the LOW-resolution tensor (rldm_prep with up = 1)
 *   wgt     : [4 phases: 2a+b][planes][4 taps: 2*ti+tj][Cout][Cin] fp16; phase (a, b), tap (ti, tj) holds the sum of the
 *             3x3 taps (kw, kh) that read the same input pixel: a=0: ti=0 <- kw 0, ti=1 <- kw 1+2; a=1: ti=0 <- kw 0+1,
 *             ti=1 <- kw 2 (same for b, tj, kh)
 *   out     : (B, 2W, 2H, Cout) fp32;  stats: optional channel-pair moments of the output (all four launches add to it)
 * Only layers for which rldm_conv_tc_up2_ok() == 1 (role-swapped kernel: Cout % 128 == 0, whole 256-pixel units, more
 * tiles than SMs). */
int rldm_conv_tc_up2(const uint16_t* x, const uint16_t* x_lo, const uint16_t* wgt, const float* bias, float* out,
                     int B, int W, int H, int Cin, int Cout, int circular, double* stats, int terms, void* stream);
int rldm_conv_tc_up2_ok(int B, int W, int H, int Cin, int Cout);

/* CUDA-core restatement of rldm_conv_tc with the identical contract (split_k ignored); used by the
 * GPU tests to isolate tensor-core descriptor bugs from precision, never by the product path. */
int rldm_conv_ref(const uint16_t* x, const uint16_t* x_lo, const uint16_t* wgt, const float* bias, const float* temb,
                  int temb_stride, const float* residual, float* out, int B, int W, int H, int Cin,
                  int Cout, int ks, int stride, int pad_lo, int circular, void* stream);

/* ---- boundary convolutions with tiny channel counts (CUDA cores, fp32) ----------------------
 * conv_in: x0 (B,C0,W,H) [+ x1 (B,C1,W,H)] fp32 REF layout, a virtual channel concat that replaces
 * torch.cat([latents, pos_encoding | condition], 1) (`ldm/pipelines.py:238,358,498`)
 * -> out (B,W,H,Cout) fp32 cl.  wgt [9][C0+C1][Cout] fp32. */
int rldm_conv_in(const float* x0, int c0, const float* x1, int c1, const float* wgt,
                 const float* bias, float* out, int B, int W, int H, int Cout, int circular,
                 void* stream);
/* conv_in that also accumulates the channel-pair moments of its output into stats[B][Cout/2][2] (double, zeroed by
 * the caller; same contract as rldm_conv_tc's `stats`), so the first GroupNorm needs no statistics pass.
 * Needs W*H % 32 == 0.  stats == NULL: identical to rldm_conv_in. */
int rldm_conv_in_stats(const float* x0, int c0, const float* x1, int c1, const float* wgt, const float* bias,
                       float* out, int B, int W, int H, int Cout, int circular, double* stats, void* stream);
/* conv_out: x (+ optional x_lo) (B,W+2,H,Cin) fp16 clp (already GN+SiLU'd by rldm_prep) -> out
 * (B,Cout,W,H) fp32 REF layout, Cout in {2,4,8}.  wgt [9][Cout][Cin] fp32. */
int rldm_conv_out(const uint16_t* x, const uint16_t* x_lo, const float* wgt, const float* bias, float* out, int B, int W,
                  int H, int Cin, int Cout, int circular, void* stream);

/* conv_norm_out + SiLU + conv_out fused (UNet2DModel tail, App. A.1; Decoder tail `model.py:1051-1056`): x (B,W,H,Cin)
 * fp32 cl -> GroupNorm(G, eps, gamma, beta; moments from `sums` [B][G][2] or channel-pair `pairs` [B][Cin/2][2], both
 * NULL = no normalisation) -> SiLU (if silu) -> 3x3 circular/zero conv -> out (B,Cout,W,H) fp32 REF layout,
 * Cout in {2,4,8}.  wgt [9][Cout][Cin] fp32 (tap = kW*3 + kH).  One launch instead of rldm_prep + rldm_conv_out and no
 * fp16 operand round trip. */
int rldm_norm_conv_out(const float* x, const double* sums, const double* pairs, const float* gamma, const float* beta,
                       float eps, int G, int silu, const float* wgt, const float* bias, float* out, int B, int W, int H,
                       int Cin, int Cout, int circular, void* stream);

/* ---- attention core ----------------------------------------------------------------------
 * Replaces F.scaled_dot_product_attention in AttnProcessor2_0: heads of
 * dim 8, softmax(QK^T/sqrt(8))V, no mask.  qkv:(B,N,3C) fp32 (q|k|v along the last dim, head h at
 * channels [8h,8h+8)); out (+ optional out_lo, split-fp16): fp16 clp (B, N/H + 2, H, C) -- the W-padded
 * operand layout of the to_out projection (halo columns are not written; a 1x1 conv never reads them). */
int rldm_attention(const float* qkv, uint16_t* out, uint16_t* out_lo, int B, int N, int C, int H,
                   void* stream);

/* ---- time embedding -----------------------------------------------------------------------
 * Replaces Timesteps + TimestepEmbedding + every ResnetBlock2D.time_emb_proj(silu(emb))
 * (SURVEY.md App. A.1 steps 1 and ResnetBlock2D).  t:(B) fp32 timesteps -- B rows: the batch of one forward, or
 * the whole timestep table of a trajectory (one row per sampling step; the conv epilogues then read row `step`
 * for every image, temb_stride 0).
 * w1:[D4][D0] b1:[D4] w2:[D4][D4] b2:[D4]; wp:[T][D4] bp:[T] = all projections stacked row-wise.
 * scratch:(2,B,D4) fp32; out:(B,T) fp32.  D0, D4 multiples of 32, <= 1024.  Three launches. */
int rldm_temb(const float* t, const float* w1, const float* b1, const float* w2, const float* b2,
              const float* wp, const float* bp, float* scratch, float* out, int B, int D0, int D4,
              int T, void* stream);

/* ---- fused scheduler step -----------------------------------------------------------------
 * Replaces DDIMScheduler.step / DDPMScheduler.step / DPMSolverMultistepScheduler.step
 * (call sites `ldm/pipelines.py:106,244-246,362,502`) with ONE kernel:
 *    x0   = k[0]*x + k[1]*eps
 *    xout = k[2]*x + k[3]*x0 + k[4]*x0_prev + k[5]*eps + k[6]*noise
 * x0_prev / noise / x0_out may be NULL (their coefficient must then be 0).  k: 7 floats on the
 * DEVICE (a row of the per-step coefficient table), so the step is graph-replayable. */
int rldm_sched_step(const float* k, const float* x, const float* eps, const float* x0_prev,
                    const float* noise, float* x_out, float* x0_out, int64_t n, void* stream);

/* ---- layout helpers ----------------------------------------------------------------------- */
int rldm_ref_to_cl(const float* src, float* dst, int B, int C, int W, int H, void* stream);
int rldm_cl_to_ref(const float* src, float* dst, int B, int C, int W, int H, void* stream);

/* ---- whole-forward programs ----------------------------------------------------------------
 * A program is a flat array of rldm_op records built once by the host (all pointers resolved);
 * rldm_run launches them back to back on `stream` with no host work in between.  This is what
 * UNet2DModel.forward / AutoencoderKL.decode / the sampling loop execute. */
enum {
  RLDM_OP_GN_STATS = 1, RLDM_OP_PREP = 2, RLDM_OP_CONV_TC = 3, RLDM_OP_CONV_IN = 4,
  RLDM_OP_CONV_OUT = 5, RLDM_OP_ATTENTION = 6, RLDM_OP_TEMB = 7, RLDM_OP_SCHED_STEP = 8,
  RLDM_OP_MEMSET = 9, RLDM_OP_CONV_REF = 10, RLDM_OP_AXPY = 11, RLDM_OP_NORM_CONV_OUT = 12,
  RLDM_OP_FUSED = 13,    /* p[0] = rldm_fused handle: a compiled run of small ops (below) */
  RLDM_OP_CONV_UP2 = 14  /* rldm_conv_tc_up2: p = x, x_lo, wgt, bias, out, stats; i = B, W, H, Cin, Cout, circular, terms */
};
typedef struct rldm_op {
  int32_t kind;
  int32_t i[23];      /* integer arguments in the order of the matching entry point */
  float f[2];         /* float arguments (eps) */
  void* p[24];        /* pointer arguments in the order of the matching entry point */
  int64_t n;          /* element / byte count where the entry point takes one */
} rldm_op;
int rldm_run(const rldm_op* ops, int n_ops, void* stream);

/* ---- fused runs of small layers -------------------------------------------------------------
 * Levels 1..n of the UNet are chains of tiny dependent ops (GroupNorm-apply passes, 3x3 / 1x1 convolutions of a few
 * K steps, short attention).  A maximal run of consecutive PREP / CONV_TC / ATTENTION ops that rldm_fused_supported()
 * accepts can be compiled into ONE persistent launch: all SMs walk the run phase by phase, separated by grid-wide
 * barriers instead of kernel boundaries; barriers, TMEM and the TMA ring are set up once.  Same operand layouts and
 * results as the stand-alone entry points up to the fp32 summation order of the K split (which is fixed:
 * deterministic).  The launch needs every CTA resident at once: do not run two fused handles concurrently on
 * different streams of one device.
 *   rldm_fused_supported : 1 when the op can be part of a run
 *   rldm_fused_ws_bytes  : bytes of split-K workspace the run needs (shared by all runs of one stream)
 *   rldm_fused_create    : compiles the run (device-side phase table + tensor maps; synchronous, not capturable)
 *   rldm_fused_run       : one launch, stream-ordered, CUDA-graph capturable
 */
typedef struct rldm_fused rldm_fused;
int rldm_fused_supported(const rldm_op* op);
long long rldm_fused_ws_bytes(const rldm_op* ops, int n_ops);
int rldm_fused_create(const rldm_op* ops, int n_ops, float* ws, long long ws_bytes, rldm_fused** out);
int rldm_fused_run(rldm_fused* h, void* stream);
void rldm_fused_destroy(rldm_fused* h);
/* Profiling variant: the same launches, with a one-thread kernel after every op (and one before the first) that
 * writes %globaltimer (ns) into stamps[0..n_ops]; stamps[k+1]-stamps[k] is op k's serialised, cache-warm
 * duration.  Graph-capturable.  Used by bench.py's roofline pass and scripts/, never by the sampling path. */
int rldm_run_timed(const rldm_op* ops, int n_ops, unsigned long long* stamps, void* stream);

/* ---- range image -> point cloud -----------------------------------------
 * Replaces `point_cloud_to_range_image.to_pc_torch` (`ldm/dataset.py:228-276`), called on every generated batch
 * (`ldm/inference.py:171`).  img (B,C,W,H) fp32 reference layout, channel 0 = encoded range, channel 1 = remission.
 * mode 0: r = img*std + mean; 1: r = 2^(6 img) - 1 (`log`); 2: r = 1/max(img, 1e-4) (`inverse`); r < 0 -> fill
 * (range_fill_value[0]).  incl / height: per-beam inclination and sensor height tables (length H, e.g.
 * `ldm/kitti360_range_image.py:19-48`).  points (B, W*H, C > 1 ? 4 : 3) fp32 = x, y, z [, remission], point index
 * w*H + h; depth (optional, B x W*H) = |xyz|, the quantity the .bin writer masks with `< 90` (`inference.py:176-178`). */
int rldm_range_to_points(const float* img, int B, int C, int W, int H, const float* incl, const float* height, int mode,
                         float mean, float stdv, float fill, float* points, float* depth, void* stream);

/* Point cloud -> bird's-eye-view volume: `to_voxel` (`ldm/dataset.py:278-294`) = `_splat_points_to_volumes`
 * (`:13-132`, trilinear votes of every point, the reference's 16 scatter_add_ passes) + features / clamp(density, 1e-4)
 * + log(density + 1) when `normalize`.  points (B,N,P) fp32 with P = 3 or 4 (x, y, z [, remission]) in metres,
 * pc_range6 HOST pointer {xmin,ymin,zmin,xmax,ymax,zmax}; grid (D,Hh,Ww) = the reference's grid_sizes.
 * scratch: 2*B*D*Hh*Ww floats (zeroed here); voxel: (B, 2*D, Hh, Ww) fp32 = [densities, features]
 * (`voxel = torch.cat([volume_densities, volume_features], dim=1)`, `:292`).  Votes are float atomics: sums agree with
 * the reference to rounding order. */
int rldm_points_to_voxel(const float* points, int B, int N, int P, const float* pc_range6, int D, int Hh, int Ww,
                         int normalize, float* scratch, float* voxel, void* stream);

/* ---- point cloud -> range image --------------------------------------------
 * Replaces `point_cloud_to_range_image.__call__` + `process_miss_value` + `normalize` (`ldm/dataset.py:159-226`) with
 * the KITTI beam assignment of `ldm/kitti360_range_image.py:51-61`, i.e. the sample `RangeDataset.__getitem__` builds
 * (`:327-336`).  pc (N,4) fp32 x,y,z,remission (16 B aligned); incl/height per-beam tables (length H).  The nearest
 * return of a pixel wins (equal ranges: the lower point index).  keys: H*W uint64 scratch.
 * image (2, W, H) fp32 = [encoded range (mode as in rldm_range_to_points; linear is (r - mean)/std), remission];
 * mask, car_window (W, H) uint8 = `range_image_mask`, `car_window_mask`. */
int rldm_points_to_range(const float* pc, int N, const float* incl, const float* height, int H, int W, int mode,
                         float mean, float stdv, float fill_range, float fill_rem, unsigned long long* keys,
                         float* image, unsigned char* mask, unsigned char* car_window, void* stream);

/* y = a*x (elementwise, fp32), e.g. latents / scaling_factor (`ldm/pipelines.py:365`). */
int rldm_scale(const float* x, float a, float* y, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RLDM_H_ */
