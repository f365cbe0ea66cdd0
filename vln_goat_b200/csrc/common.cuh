// Shared device/host helpers for libgoat_sm100 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/goat_sm100.h"

namespace goat {

// ---------------------------------------------------------------------------------------------
// error plumbing (C-ABI functions return non-zero and leave a message for goat_last_error())
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);

#define GOAT_CHECK(cond, ...)                                  \
  do {                                                         \
    if (!(cond)) {                                             \
      ::goat::set_error(__VA_ARGS__);                          \
      return GOAT_ERR_INVALID;                                 \
    }                                                          \
  } while (0)

#define GOAT_CUDA(expr)                                                              \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      ::goat::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
                        __FILE__, __LINE__);                                         \
      return GOAT_ERR_CUDA;                                                          \
    }                                                                                \
  } while (0)

#define GOAT_LAUNCH_CHECK()                                                          \
  do {                                                                               \
    cudaError_t _e = cudaGetLastError();                                             \
    if (_e != cudaSuccess) {                                                         \
      ::goat::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                        __FILE__, __LINE__);                                         \
      return GOAT_ERR_CUDA;                                                          \
    }                                                                                \
  } while (0)

inline int dtype_size(int dt) { return dt == GOAT_F32 ? 4 : 2; }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---------------------------------------------------------------------------------------------
// scalar conversions
// ---------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// pack two floats into one 32-bit word of 16-bit values (lo = a, hi = b)
template <typename T> __device__ __forceinline__ uint32_t pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <typename T> __device__ __forceinline__ float2 unpack2(uint32_t w);
template <> __device__ __forceinline__ float2 unpack2<__half>(uint32_t w) {
  return __half22float2(*reinterpret_cast<__half2*>(&w));
}
template <> __device__ __forceinline__ float2 unpack2<__nv_bfloat16>(uint32_t w) {
  return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&w));
}

// ---------------------------------------------------------------------------------------------
// math
// ---------------------------------------------------------------------------------------------
// erf-GELU, reference: pretrain_src/model/Bert_backbone.py:41-47
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
// d/dx gelu_erf(x) = Phi(x) + x * phi(x)
__device__ __forceinline__ float dgelu_erf(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Dropout masks are never stored: every kernel regenerates them from (seed, element index), so forward and backward
// (and the tcgen05 and SIMT variants of a kernel) agree by construction.  ONE 32-bit hash serves the element PAIR
// (2k, 2k+1) -- 16 random bits each -- and an element is kept iff its 16 bits >= round(p * 65536) (p is realised to
// within 8e-6).  The hash is two rounds of multiply-and-fold (32x32->64-bit product, high word xor low word): about
// 5 integer instructions per pair once the per-row part is hoisted.
__device__ __forceinline__ uint32_t mulfold(uint32_t a, uint32_t b) {
  const unsigned long long r = (unsigned long long)a * (unsigned long long)b;
  return (uint32_t)(r >> 32) ^ (uint32_t)r;
}
__device__ __forceinline__ uint32_t drop_thr16(float p) { return (uint32_t)__float2uint_rn(p * 65536.0f); }
// 32 random bits for pair `pair_idx` of a linearly indexed tensor (GEMM epilogue, LayerNorm backward, casts)
__device__ __forceinline__ uint32_t drop_hash(unsigned long long seed, unsigned long long pair_idx) {
  const uint32_t x = mulfold((uint32_t)pair_idx ^ (uint32_t)seed ^ 0x9E3779B1u,
                             (uint32_t)(pair_idx >> 32) ^ (uint32_t)(seed >> 32) ^ 0x85EBCA6Bu);
  return mulfold(x ^ 0x7FEB352Du, 0x846CA68Bu);
}
// keep decision of one element (scalar paths)
__device__ __forceinline__ bool drop_keep(unsigned long long seed, unsigned long long idx, uint32_t thr16) {
  const uint32_t r = drop_hash(seed, idx >> 1);
  return ((idx & 1ull) ? (r >> 16) : (r & 0xFFFFu)) >= thr16;
}
// keep decisions of elements idx, idx+1 (idx even)
__device__ __forceinline__ void drop_keep2(unsigned long long seed, unsigned long long idx_even, uint32_t thr16, bool& k0, bool& k1) {
  const uint32_t r = drop_hash(seed, idx_even >> 1);
  k0 = (r & 0xFFFFu) >= thr16;
  k1 = (r >> 16) >= thr16;
}
// 4 consecutive elements starting at idx (idx % 4 == 0): v[j] = keep ? v[j] * scale : 0
__device__ __forceinline__ void drop_apply4(float (&v)[4], unsigned long long seed, unsigned long long idx, uint32_t thr16, float scale) {
  bool k0, k1, k2, k3;
  drop_keep2(seed, idx, thr16, k0, k1);
  drop_keep2(seed, idx + 2, thr16, k2, k3);
  v[0] = k0 ? v[0] * scale : 0.0f;
  v[1] = k1 ? v[1] * scale : 0.0f;
  v[2] = k2 ? v[2] * scale : 0.0f;
  v[3] = k3 ? v[3] * scale : 0.0f;
}
// Attention-probability dropout: a 32-bit key per (seed, batch, head, query row), then one hash per key PAIR (kj, kj+1),
// kj even.  Shared by the SIMT, chunked tcgen05 and pipelined tcgen05 attention kernels.
__device__ __forceinline__ uint32_t attn_drop_rowkey(unsigned long long seed, int b, int heads, int h, int Nq, int qi) {
  return drop_hash(seed, ((unsigned long long)b * heads + h) * (unsigned long long)Nq + qi);
}
__device__ __forceinline__ uint32_t attn_drop_pair(uint32_t rowkey, int kj_even) {
  const uint32_t x = mulfold(rowkey ^ ((uint32_t)(kj_even >> 1) * 0x9E3779B1u), 0x85EBCA6Bu ^ rowkey);
  return mulfold(x ^ 0x7FEB352Du, 0x846CA68Bu);
}
__device__ __forceinline__ bool attn_drop_keep(uint32_t rowkey, int kj, uint32_t thr16) {
  const uint32_t r = attn_drop_pair(rowkey, kj & ~1);
  return ((kj & 1) ? (r >> 16) : (r & 0xFFFFu)) >= thr16;
}

// erf-GELU and its derivative for the 16-bit tensor-core epilogues: Abramowitz-Stegun 7.1.26 (|erf error| < 1.5e-7,
// far below fp16 / bf16 output rounding) with one MUFU.RCP and one MUFU.EX2; exp(-x^2/2) is shared between the
// erf tail and the Gaussian density of the derivative.  The fp32 parity path keeps erff (gelu_erf / dgelu_erf).
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// cdf = Phi(x), e = exp(-x^2 / 2).  h = erfc(|x| / sqrt 2) / 2 = Phi(-|x|) from the A&S polynomial (coefficients pre-halved),
// then Phi(x) = x >= 0 ? 1 - h : h.  ~15 instructions per element including the two MUFU ops.
__device__ __forceinline__ void gelu_parts(float x, float& cdf, float& e) {
  const float t = rcp_approx(fmaf(0.3275911f * 0.70710678118654752440f, fabsf(x), 1.0f));
  float p = fmaf(t, 0.5f * 1.061405429f, 0.5f * -1.453152027f);
  p = fmaf(p, t, 0.5f * 1.421413741f);
  p = fmaf(p, t, 0.5f * -0.284496736f);
  p = fmaf(p, t, 0.5f * 0.254829592f);
  e = ex2_approx(x * x * -0.72134752044448170368f);      // exp(-x^2 / 2)
  const float h = p * t * e;
  cdf = x >= 0.0f ? 1.0f - h : h;
}
__device__ __forceinline__ float gelu_fast(float x) {
  float cdf, e;
  gelu_parts(x, cdf, e);
  return x * cdf;
}
__device__ __forceinline__ float dgelu_fast(float x) {
  float cdf, e;
  gelu_parts(x, cdf, e);
  return cdf + x * e * 0.39894228040143267794f;
}

// ---------------------------------------------------------------------------------------------
// epilogue shared by the tcgen05 GEMM and the SIMT GEMM
// ---------------------------------------------------------------------------------------------
struct EpiParams {
  const float* bias;     // [N] fp32 or null
  const float* res;      // [M, ldres] fp32 residual added after activation/dropout, or null
  const void* aux_in;    // [M, ldaux] (16-bit T or fp32, same type as `in` operands) for DGELU/DRELU
  void* aux_out;         // [M, ldaux] pre-activation copy written by GELU (same type as out16/T), or null
  void* out;             // [M, ldc] fp32 or T
  void* out2;            // optional second copy of the result in T (16-bit) when out is fp32, or null
  int ldc, ldres, ldaux, ldc2;
  int act;               // goat_act_t
  int out_f32;           // 1: out is fp32, 0: out is T
  int accumulate;        // 1: out (fp32) += alpha * acc with atomics, no other epilogue term (split-K wgrad)
  float alpha;           // scales the accumulator before bias
  float drop_p;          // dropout probability applied after activation, before residual (0 = off)
  unsigned long long drop_seed;
  const unsigned long long* drop_seed_ptr;  // optional device word added to drop_seed
  long long drop_row0;   // global row of this launch's row 0 (a GEMM split over two launches keeps ONE dropout mask)
};

__device__ __forceinline__ unsigned long long eff_seed(unsigned long long seed, const unsigned long long* ptr) {
  return ptr ? seed + *ptr : seed;
}

template <typename T>
__device__ __forceinline__ float epi_apply(const EpiParams& ep, int m, int n, float acc) {
  float v = acc * ep.alpha;
  if (ep.bias) v += __ldg(ep.bias + n);
  if (ep.act == GOAT_ACT_GELU) {
    if (ep.aux_out) reinterpret_cast<T*>(ep.aux_out)[(size_t)m * ep.ldaux + n] = from_f<T>(v);
    // round the stored pre-activation first so backward (which sees the rounded z) matches forward
    if (ep.aux_out) v = to_f<T>(from_f<T>(v));
    v = gelu_erf(v);
  } else if (ep.act == GOAT_ACT_RELU) {
    v = fmaxf(v, 0.0f);
  } else if (ep.act == GOAT_ACT_DGELU) {
    v *= dgelu_erf(to_f<T>(reinterpret_cast<const T*>(ep.aux_in)[(size_t)m * ep.ldaux + n]));
  } else if (ep.act == GOAT_ACT_DRELU) {
    v = (to_f<T>(reinterpret_cast<const T*>(ep.aux_in)[(size_t)m * ep.ldaux + n]) > 0.0f) ? v : 0.0f;
  } else if (ep.act == GOAT_ACT_TANH) {
    v = tanhf(v);
  }
  if (ep.drop_p > 0.0f) {
    const bool keep = drop_keep(eff_seed(ep.drop_seed, ep.drop_seed_ptr),
                                (unsigned long long)(m + ep.drop_row0) * (unsigned long long)ep.ldc + n, drop_thr16(ep.drop_p));
    v = keep ? v * (1.0f / (1.0f - ep.drop_p)) : 0.0f;
  }
  if (ep.res) v += ep.res[(size_t)m * ep.ldres + n];
  return v;
}

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch.  Kernels launched through launch_pdl() may become resident while the previous kernel
// of the stream is still draining: their prologue (barrier init, TMEM allocation, descriptor prefetch) overlaps its
// tail, and pdl_wait() -- placed before the FIRST global-memory access -- blocks until the predecessor has completed
// and its writes are visible.  pdl_launch_dependents() lets the successor do the same with this kernel.
// A kernel that is launched with the attribute MUST execute pdl_wait() before touching global memory.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pdl_enabled();   // GOAT_PDL=0 turns the launch attribute off (api.cu)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// ---------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier, TMA, tcgen05  (sm_100a)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 1024-byte aligned start inside a dynamic shared memory buffer.  Pointer arithmetic (not an integer round trip) so the
// compiler keeps the shared state space and emits LDS / STS instead of generic LD / ST for everything derived from it.
__device__ __forceinline__ uint8_t* smem_align1024(uint8_t* p) { return p + ((1024u - (smem_u32(p) & 1023u)) & 1023u); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must trap (the context dies, the GPU survives) instead of hanging.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) {  // ~2 s at 2 GHz
      printf("goat: mbarrier wait timed out (block %d,%d thread %d parity %u)\n", blockIdx.x, blockIdx.y,
             threadIdx.x, parity);
      __trap();
    }
  }
}

__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
// 2-D tiled TMA load global -> shared, completion on an mbarrier (complete_tx::bytes)
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// whole warp; writes the TMEM base address to *smem_slot
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem];  kind::f16 covers fp16 and bf16 inputs with fp32 accumulation
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// TMEM -> registers: this thread's lane (row), 32 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// TMEM -> registers: this thread's lane (row), 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 64-bit shared-memory matrix descriptor for tcgen05.mma (SWIZZLE_128B canonical layouts).
// Field layout follows the PTX ISA "shared memory descriptor" table (same bits CUTLASS's
// cute::UMMA::SmemDescriptor names): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), base_offset [49,52), layout_type [61,64) with SWIZZLE_128B = 2.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// 32-bit instruction descriptor for kind::f16: fp32 accumulator, A/B format (0 = fp16, 1 = bf16),
// A/B major-ness (0 = K-major, 1 = MN-major), N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc_f16(int fmt, int a_mn, int b_mn, int M, int N) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <typename T> struct UmmaFmt;
template <> struct UmmaFmt<__half> { static constexpr int value = 0; };
template <> struct UmmaFmt<__nv_bfloat16> { static constexpr int value = 1; };

}  // namespace goat
