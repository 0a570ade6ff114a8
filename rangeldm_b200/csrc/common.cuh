// common.cuh -- error plumbing and sm_100a PTX wrappers shared by the librldm kernels.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/rldm.h"

namespace rldm {

void set_error(const char* fmt, ...);

#define RLDM_CHECK(cond, ...)          \
  do {                                 \
    if (!(cond)) {                     \
      rldm::set_error(__VA_ARGS__);    \
      return 1;                        \
    }                                  \
  } while (0)

#define RLDM_CUDA(call)                                                              \
  do {                                                                               \
    cudaError_t e__ = (call);                                                        \
    if (e__ != cudaSuccess) {                                                        \
      rldm::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return 2;                                                                      \
    }                                                                                \
  } while (0)

#define RLDM_LAUNCH_CHECK() RLDM_CUDA(cudaGetLastError())

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Programmatic dependent launch: kernels launched with the programmatic-stream-serialization attribute may start
// while the preceding kernel is still running; they execute `griddepcontrol.wait` before their first access to
// memory a previous kernel may have written (or may still read) -- the wait returns only when the preceding grid has
// completed and flushed, so correctness is that of plain stream order -- and every kernel fires
// `griddepcontrol.launch_dependents` early.  Without the attribute both instructions are no-ops.  Which launches get
// the attribute is a measured policy (ops.cu: pdl_mode): by default only the latency-bound ones.
// Environment switches of librldm.so, read ONCE at first use (rldm_reload_env() re-reads them: the GPU tests flip
// them between cases).  Defaults are the measured best; every other setting exists so a number in DESIGN.md can be
// re-checked.
struct EnvSwitches {
  int pdl;                  // RLDM_PDL = 0 | 1 | 2 (default 2: only latency-bound launches carry the PDL attribute)
  size_t prep_pdl_max;      // RLDM_PREP_PDL_MAX: largest prep pass (elements) that carries it (default 2e7)
  int conv_wt;              // RLDM_CONV_WT:      1 role-swapped conv kernel (default) | 0 off | 2 ("nores")
  int conv_wt_halo;         // RLDM_CONV_WT_HALO: 1 pixel windows inside it (default) | 0 off | 2 ("nores")
  bool conv_persistent;     // RLDM_NO_PERSISTENT unset
  bool conv_mt1;            // RLDM_CONV_MT1:     pixel-M persistent kernel with one tile per unit
  bool conv_mt2_res;        // RLDM_CONV_MT2_RES: two tiles per unit also for 128-wide layers with a residual
  bool wt_pdl;              // RLDM_WT_PDL != 0:  PDL on single-wave role-swapped launches (default on)
  int emit_maxcl;           // RLDM_EMIT_MAXCL: largest cluster (CTAs) of an emitting launch that spans several M tiles
  int emit_maxclm;          // RLDM_EMIT_MAXCLM: emitting convolutions only for images of at most this many 128-pixel tiles
  int pdl_extra;            // RLDM_PDL_EXTRA: bit mask of launch classes that also carry the PDL attribute in mode 2
  bool wt_pdl_all;          // RLDM_WT_PDL = 2: ... on every role-swapped / persistent convolution launch (experiment)
  int small_bn64;           // RLDM_SMALL_BN64: 64-wide tiles for 1x1 conv layers with at most this many 128x128 tiles (default 128)
  bool small_bn64_all;      // RLDM_SMALL_BN64_ALL: ... for the 3x3 layers as well (experiment)
  int attn_poly;            // RLDM_ATTN_POLY: exponential pairs of every 8 on the FMA pipe (tcgen05 attention, fp16 P), 0..4
  bool attn_mmasync;        // RLDM_ATTN_MMASYNC: mma.sync attention kernel for every shape it covers
  bool attn_cudacore;       // RLDM_ATTN_CUDACORE: CUDA-core attention kernel for every shape
  bool nco_pp1;             // RLDM_NCO_PP1:      norm_conv_out with one pixel per thread
  int n_sms;
};
const EnvSwitches& env();

bool pdl_enabled();
// launches marked latency-bound (small convolutions, small prep passes)
bool pdl_enabled_small();
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl_small(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                           cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled_small() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// launches of class `cls` (bit of RLDM_PDL_EXTRA: 1 tcgen05 attention, 2 short attention, 4 conv_in / conv_out, 8 scheduler
// step / scale / fill / time embedding) carry the attribute in mode 2 when their bit is set
bool pdl_enabled_class(int cls);
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl_cls(int cls, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                         cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled_class(cls) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                     cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ------------------------------------------------------------------------------------------------
// device helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
// x * sigmoid(x) with MUFU.EX2 + MUFU.RCP (~2 ulp): the IEEE division `x / (1 + e^-x)` costs ~20 instructions per
// element (FCHK + refinement + slow-path branch) and made the GroupNorm-apply passes issue-bound.
// The .ftz forms drop the denormal guards nvcc wraps around ex2.approx / the division (two FSETP + FMUL pairs per value).
__device__ __forceinline__ float silu_f(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return x * r;
}
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Experiment switches (compile time): where a kernel fires launch_dependents.
#ifdef RLDM_PDL_OPS_NOTRIGGER
__device__ __forceinline__ void pdl_entry() { pdl_wait(); }
#else
__device__ __forceinline__ void pdl_entry() { pdl_trigger(); pdl_wait(); }
#endif
#ifdef RLDM_PDL_CONV_LATE
__device__ __forceinline__ void pdl_trigger_conv_early() {}
__device__ __forceinline__ void pdl_trigger_conv_late() { pdl_trigger(); }
#else
__device__ __forceinline__ void pdl_trigger_conv_early() { pdl_trigger(); }
__device__ __forceinline__ void pdl_trigger_conv_late() {}
#endif

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier ----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      " selp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a lost arrival traps (error surfaces on the host) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) {
      printf("rldm: mbarrier timeout block(%d,%d,%d) thread %d\n", blockIdx.x, blockIdx.y,
             blockIdx.z, threadIdx.x);
      __trap();
    }
  }
}

// ---- TMA ---------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint64_t* bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4}], [%2];" ::"r"(dst),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, uint64_t* bar, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4, %5, %6}], [%2];" ::"r"(dst),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// ---- tcgen05 / TMEM ----------------------------------------------------------------------------
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(dst_smem)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem], fp16 operands, fp32 accumulate; issued by ONE thread.
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
      " tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread i <-> TMEM lane i).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor for a K-major, SWIZZLE_128B operand tile whose rows are 128 B
// (64 fp16) and whose 8-row groups are 1024 B apart (cute::UMMA::SmemDescriptor, version 1).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);   // start address  [0,14)
  d |= static_cast<uint64_t>(1) << 16;                      // LBO (ignored for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;              // SBO = 1024 B   [32,46)
  d |= static_cast<uint64_t>(1) << 46;                      // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;                      // SWIZZLE_128B
  return d;
}
// The same for SWIZZLE_64B: rows of 64 B (32 fp16), 8-row groups 512 B apart.
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);   // start address  [0,14)
  d |= static_cast<uint64_t>(1) << 16;                      // LBO (ignored for swizzled K-major)
  d |= static_cast<uint64_t>(512 >> 4) << 32;               // SBO = 512 B    [32,46)
  d |= static_cast<uint64_t>(1) << 46;                      // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(4) << 61;                      // SWIZZLE_64B
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): fp16 x fp16 -> fp32, K-major A and B.
__host__ __device__ constexpr uint32_t umma_idesc_f16(uint32_t M, uint32_t N) {
  return (1u << 4) /*c=f32*/ | (0u << 7) /*a=f16*/ | (0u << 10) /*b=f16*/ | ((N >> 3) << 17) |
         ((M >> 4) << 24);
}

}  // namespace rldm
