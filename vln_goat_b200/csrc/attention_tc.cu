// Attention core on the 5th-gen tensor cores (fp16 / bf16 operands, fp32 accumulation in TMEM).
//
//   forward : O = softmax(scale Q K^T + kmask + bias) V, lse saved
//   backward: P recomputed from lse;  dV = P~^T dO,  dS = P (dP~ - delta),  dQ = scale dS K,  dK = scale dS^T Q
//
// One CTA (128 threads = 128 TMEM lanes = 128 query rows) per (head, batch); keys are processed in chunks of
// 128.  Q / K / V / dO tiles are fetched by 3-D TMA (inner = head slice of 64 channels, token, batch; out-of-range
// tokens are zero-filled) straight from the token-major [B, N, heads*64] tensors into 128B-swizzled shared memory.
// Every contraction is a tcgen05.mma with M = 128:
//     S  = Q K^T      A = Q  (K-major)   B = K  (K-major)    N = keys
//     O  = P V        A = P  (K-major)   B = V  (MN-major)   N = 64
//     dP = dO V^T     A = dO (K-major)   B = V  (K-major)    N = keys
//     dV = P^T dO     A = P  (MN-major)  B = dO (MN-major)   N = 64   (M = keys)
//     dK = dS^T Q     A = dS (MN-major)  B = Q  (MN-major)   N = 64   (M = keys)
//     dQ = dS K       A = dS (K-major)   B = K  (MN-major)   N = 64
// A [rows][64 x 16-bit] tile with the 128-byte swizzle is simultaneously the canonical K-major layout (rows = M/N,
// 64 contiguous = K) and the canonical MN-major layout (rows = K, 64 contiguous = M/N), so P / dS / Q / K / V / dO
// are each staged once and read both ways.  The softmax runs on the TMEM rows (thread r owns query r: tcgen05.ld,
// exp, row sums in registers), writes P (and dS) as 16-bit operands back to shared memory for the second MMA.
// TMEM budget: 256 columns (S 128 + O 64 forward; S 128 + dP 128, then reused for dV 64 | dK 64 | dQ 64 backward).
#include <cuda.h>
#include <stdlib.h>

#include "attn_common.cuh"

namespace goat {

int make_tmap3(CUtensorMap* tm, int dtype, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t ld_elems,
               uint64_t sb_elems, uint32_t box0, uint32_t box1);

namespace {

using namespace attn;

constexpr int TC_THREADS = 128;
constexpr int KC = 128;               // keys per chunk
constexpr int TMEM_COLS = 256;

struct TcArgs {
  int B, heads, Nq, Nk;
  const float* kmask;
  const float* bias;
  float scale;
  float* lse;
  float drop_p;
  unsigned long long drop_seed;
  const unsigned long long* drop_seed_ptr;
  void* O; int ldo; long long sbo;      // forward output / backward: forward output (for delta)
  const void* dO;                        // backward (same layout as O)
  void* dQ; int ldq; long long sbq;
  void* dK; int ldk; long long sbk;
  void* dV; int ldv; long long sbv;
  float* dbias;
};

// write 64 fp32 values as one 128-byte row of 16-bit elements to global memory
template <typename T>
__device__ __forceinline__ void store_global_row64(T* dst, const float* v, float mul) {
  uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    uint4 w;
    w.x = pack2<T>(v[q * 8 + 0] * mul, v[q * 8 + 1] * mul);
    w.y = pack2<T>(v[q * 8 + 2] * mul, v[q * 8 + 3] * mul);
    w.z = pack2<T>(v[q * 8 + 4] * mul, v[q * 8 + 5] * mul);
    w.w = pack2<T>(v[q * 8 + 6] * mul, v[q * 8 + 7] * mul);
    d[q] = w;
  }
}

struct Smem {
  uint8_t* base;
  __device__ explicit Smem(uint8_t* raw) {
    base = smem_align1024(raw);
  }
};

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
constexpr int FWD_SMEM = 3 * TILE + 2 * TILE + 64 + 512 + 1024;  // Q K V | P(2 blocks) | barriers | key mask | alignment

template <typename T>
__global__ void __launch_bounds__(TC_THREADS)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, TcArgs p) {
  extern __shared__ uint8_t smem_raw[];
  Smem sm(smem_raw);
  uint8_t* sQ = sm.base;
  uint8_t* sK = sQ + TILE;
  uint8_t* sV = sK + TILE;
  uint8_t* sP = sV + TILE;
  uint64_t* bar_q = reinterpret_cast<uint64_t*>(sP + 2 * TILE);
  uint64_t* bar_kv = bar_q + 1;
  uint64_t* bar_mma = bar_q + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_q + 3);
  float* kml = reinterpret_cast<float*>(bar_q + 8);   // [128] additive key mask of the current chunk, log2 domain

  const int tid = threadIdx.x, warp = tid >> 5;
  const int h = blockIdx.x, b = blockIdx.y;
  const int q0 = blockIdx.z * 128;          // this CTA's tile of 128 query rows (long instructions: Nq up to 514)
  if (tid == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(bar_q, 1); mbar_init(bar_kv, 1); mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<TMEM_COLS>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_row = tmem + ((uint32_t)(warp * 32) << 16);
  constexpr int FMT = UmmaFmt<T>::value;

  if (tid == 0) {
    mbar_arrive_expect_tx(bar_q, TILE);
    tma_load_3d(sQ, &tmQ, bar_q, h * 64, q0, b);
  }
  const int r = q0 + tid;                   // global query row; the smem / TMEM row of this thread is tid
  const bool rv = r < p.Nq;
  const bool warp_live = q0 + warp * 32 < p.Nq;
  const float* brow = p.bias ? p.bias + ((long long)b * p.Nq + (rv ? r : 0)) * p.Nk : nullptr;
  const float* krow = p.kmask ? p.kmask + (long long)b * p.Nk : nullptr;
  const unsigned long long seed = p.drop_p > 0.f ? eff_seed(p.drop_seed, p.drop_seed_ptr) : 0ull;
  const int nchunks = (p.Nk + KC - 1) / KC;
  uint32_t ph_kv = 0, ph_mma = 0;
  float m = -INFINITY, l = 0.f;

  // S[128 x nk16] = Q K_c^T into TMEM columns [0, nk16)
  auto issue_qk = [&](int nk16) {
    const uint32_t idesc = make_idesc_f16(FMT, 0, 0, 128, nk16);
    const uint32_t a0 = smem_u32(sQ), b0 = smem_u32(sK);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      umma_f16(tmem, make_smem_desc_sw128(a0 + k * 32, 0, 1024), make_smem_desc_sw128(b0 + k * 32, 0, 1024), idesc,
               k ? 1u : 0u);
  };
  const float sl2 = p.scale * LOG2E;
  const bool drop = p.drop_p > 0.f;
  const uint32_t thr16 = drop_thr16(p.drop_p);
  const float keep_scale = drop ? 1.f / (1.f - p.drop_p) : 1.f;
  const uint32_t rowkey = drop ? attn_drop_rowkey(seed, b, p.heads, h, p.Nq, r) : 0u;
  auto fill_kml = [&](int c) {
    const int key = c * KC + tid;
    kml[tid] = key < p.Nk ? (krow ? __ldg(krow + key) * LOG2E : 0.f) : -INFINITY;
  };
  // Branch-free 16-key blocks (same arithmetic as attention_pipe.cu): masked / out-of-range keys carry -inf in the staged
  // mask; rows past Nq see zero-filled Q rows (S = 0) and are never stored.  One hash per key PAIR for dropout, the per-row
  // key hoisted -- the per-element lambdas this replaces were ~40 dependent instructions per score on ONE warp per scheduler.
  auto block_max = [&](int c, int nk16, float mm) -> float {
#pragma unroll 1
    for (int kb = 0; kb * 16 < nk16; ++kb) {
      uint32_t rr[16];
      float kv[16];
      tmem_ld_32x32b_x16(t_row + kb * 16, rr);
      load_kv16c(kml, brow, kb * 16, c * KC + kb * 16, p.Nk, kv);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) mm = fmaxf(mm, fmaf(__uint_as_float(rr[j]), sl2, kv[j]));
    }
    return mm;
  };

  // ---- pass A (only when the keys do not fit one chunk): exact row maxima
  if (nchunks > 1) {
    for (int c = 0; c < nchunks; ++c) {
      const int nk = min(KC, p.Nk - c * KC), nk16 = (nk + 15) & ~15;
      if (tid == 0) {
        mbar_arrive_expect_tx(bar_kv, TILE);
        tma_load_3d(sK, &tmK, bar_kv, h * 64, c * KC, b);
        if (c == 0) mbar_wait(bar_q, 0);
        mbar_wait(bar_kv, ph_kv);
        tcgen05_fence_after();
        issue_qk(nk16);
        umma_commit(bar_mma);
      }
      ph_kv ^= 1;
      fill_kml(c);
      mbar_wait(bar_mma, ph_mma);
      ph_mma ^= 1;
      tcgen05_fence_after();
      __syncthreads();
      if (warp_live) m = block_max(c, nk16, m);
      tcgen05_fence_before();
      __syncthreads();
    }
  }

  // ---- pass B: P = exp(S - m), O += P V
  for (int c = 0; c < nchunks; ++c) {
    const int nk = min(KC, p.Nk - c * KC), nk16 = (nk + 15) & ~15;
    if (tid == 0) {
      mbar_arrive_expect_tx(bar_kv, 2 * TILE);
      tma_load_3d(sK, &tmK, bar_kv, h * 64, c * KC, b);
      tma_load_3d(sV, &tmV, bar_kv, h * 64, c * KC, b);
      if (c == 0 && nchunks == 1) mbar_wait(bar_q, 0);
      mbar_wait(bar_kv, ph_kv);
      tcgen05_fence_after();
      issue_qk(nk16);
      umma_commit(bar_mma);
    }
    ph_kv ^= 1;
    fill_kml(c);
    mbar_wait(bar_mma, ph_mma);
    ph_mma ^= 1;
    tcgen05_fence_after();
    __syncthreads();
    if (warp_live) {
      if (nchunks == 1) m = block_max(c, nk16, m);
      const float mref = (m == -INFINITY) ? 0.f : m;
#pragma unroll 1
      for (int kb = 0; kb * 16 < nk16; ++kb) {
        uint32_t rr[16];
        float kv[16], pv[16];
        tmem_ld_32x32b_x16(t_row + kb * 16, rr);
        load_kv16c(kml, brow, kb * 16, c * KC + kb * 16, p.Nk, kv);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float e = ex2_approx(fmaf(__uint_as_float(rr[j]), sl2, kv[j]) - mref);
          l += e;
          pv[j] = e;
        }
        if (drop) drop_mul16(pv, rowkey, c * KC + kb * 16, thr16, keep_scale);
        store_row16<T>(sP, tid, kb * 16, pv);
      }
    }
    fence_proxy_async();
    tcgen05_fence_before();
    __syncthreads();
    if (tid == 0) {
      tcgen05_fence_after();
      const uint32_t idesc = make_idesc_f16(FMT, 0, 1, 128, 64);
      const uint32_t a0 = smem_u32(sP), b0 = smem_u32(sV);
      for (int kk = 0; kk * 16 < nk16; ++kk)
        umma_f16(tmem + 128, make_smem_desc_sw128(a0 + (kk >> 2) * TILE + (kk & 3) * 32, 0, 1024),
                 make_smem_desc_sw128(b0 + kk * 2048, 8192, 1024), idesc, (c | kk) ? 1u : 0u);
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, ph_mma);
    ph_mma ^= 1;
    tcgen05_fence_after();
  }

  // ---- epilogue: O / l -> global, lse
  if (warp_live) {
    float o[64];
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      uint32_t rr[32];
      tmem_ld_32x32b_x32(t_row + 128 + g * 32, rr);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) o[g * 32 + j] = __uint_as_float(rr[j]);
    }
    if (rv) {
      const float inv = l > 0.f ? 1.f / l : 0.f;
      T* dst = reinterpret_cast<T*>(p.O) + (long long)b * p.sbo + (long long)r * p.ldo + h * 64;
      store_global_row64<T>(dst, o, inv);
      if (p.lse) p.lse[((long long)b * p.heads + h) * p.Nq + r] = l > 0.f ? (m + __log2f(l)) * LN2 : -INFINITY;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<TMEM_COLS>(tmem);
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
constexpr int BWD_SMEM = 4 * TILE + 4 * TILE + 64 + 512 + 1024;  // Q dO K V | P(2) dS(2) | barriers | key mask | alignment

// KV_OUT = true : Nq <= 128, one CTA per (head, batch) produces dQ, dK and dV.
// KV_OUT = false: Nq > 128, one CTA per (head, batch, QUERY tile) produces only its dQ rows (accumulated over the key
//                 chunks in registers); dK / dV come from attn_bwd_tc_kv_kernel below (one CTA per KEY chunk, accumulated
//                 over the query tiles in TMEM).  S / dP are recomputed by both passes: the tensor work of this
//                 HBM- and latency-bound op is negligible, and no cross-CTA reduction (atomics, workspace) is needed.
template <typename T, bool KV_OUT>
__global__ void __launch_bounds__(TC_THREADS)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO, TcArgs p) {
  extern __shared__ uint8_t smem_raw[];
  Smem sm(smem_raw);
  uint8_t* sQ = sm.base;
  uint8_t* sdO = sQ + TILE;
  uint8_t* sK = sdO + TILE;
  uint8_t* sV = sK + TILE;
  uint8_t* sP = sV + TILE;
  uint8_t* sdS = sP + 2 * TILE;
  uint64_t* bar_q = reinterpret_cast<uint64_t*>(sdS + 2 * TILE);
  uint64_t* bar_kv = bar_q + 1;
  uint64_t* bar_mma = bar_q + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_q + 3);
  float* kml = reinterpret_cast<float*>(bar_q + 8);   // [128] additive key mask of the current chunk, log2 domain

  const int tid = threadIdx.x, warp = tid >> 5;
  const int h = blockIdx.x, b = blockIdx.y;
  const int q0 = blockIdx.z * 128;
  if (tid == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmdO);
    mbar_init(bar_q, 1); mbar_init(bar_kv, 1); mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<TMEM_COLS>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_row = tmem + ((uint32_t)(warp * 32) << 16);
  constexpr int FMT = UmmaFmt<T>::value;

  if (tid == 0) {
    mbar_arrive_expect_tx(bar_q, 2 * TILE);
    tma_load_3d(sQ, &tmQ, bar_q, h * 64, q0, b);
    tma_load_3d(sdO, &tmdO, bar_q, h * 64, q0, b);
  }
  const int r = q0 + tid;                   // global query row
  const bool rv = r < p.Nq;
  const float* brow = p.bias ? p.bias + ((long long)b * p.Nq + (rv ? r : 0)) * p.Nk : nullptr;
  const float* krow = p.kmask ? p.kmask + (long long)b * p.Nk : nullptr;
  float* dbrow = p.dbias ? p.dbias + ((long long)b * p.Nq + (rv ? r : 0)) * p.Nk : nullptr;
  const unsigned long long seed = p.drop_p > 0.f ? eff_seed(p.drop_seed, p.drop_seed_ptr) : 0ull;
  const int nchunks = (p.Nk + KC - 1) / KC;
  const int nq16 = (min(p.Nq - q0, 128) + 15) & ~15;

  // delta_r = sum_d dO[r,d] O[r,d];  lse_r
  float delta = 0.f, lse = 0.f;
  if (rv) {
    const long long off = (long long)b * p.sbo + (long long)r * p.ldo + h * 64;
    const uint4* po = reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(p.O) + off);
    const uint4* pg = reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(p.dO) + off);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const uint4 a = __ldg(po + q), g = __ldg(pg + q);
      float2 x, y;
      x = unpack2<T>(a.x); y = unpack2<T>(g.x); delta += x.x * y.x + x.y * y.y;
      x = unpack2<T>(a.y); y = unpack2<T>(g.y); delta += x.x * y.x + x.y * y.y;
      x = unpack2<T>(a.z); y = unpack2<T>(g.z); delta += x.x * y.x + x.y * y.y;
      x = unpack2<T>(a.w); y = unpack2<T>(g.w); delta += x.x * y.x + x.y * y.y;
    }
    lse = p.lse[((long long)b * p.heads + h) * p.Nq + r] * LOG2E;
  }
  const float sl2 = p.scale * LOG2E;
  const bool drop = p.drop_p > 0.f;
  const uint32_t thr16 = drop_thr16(p.drop_p);
  const float keep_scale = drop ? 1.f / (1.f - p.drop_p) : 1.f;
  const uint32_t rowkey = drop ? attn_drop_rowkey(seed, b, p.heads, h, p.Nq, r) : 0u;
  if (lse == -INFINITY) lse = INFINITY;     // a fully masked row: every probability exp2(s - inf) = 0 instead of NaN
  float dq[64];
#pragma unroll
  for (int j = 0; j < 64; ++j) dq[j] = 0.f;

  uint32_t ph_kv = 0, ph_mma = 0;
  for (int c = 0; c < nchunks; ++c) {
    const int nk = min(KC, p.Nk - c * KC), nk16 = (nk + 15) & ~15;
    if (tid == 0) {
      mbar_arrive_expect_tx(bar_kv, 2 * TILE);
      tma_load_3d(sK, &tmK, bar_kv, h * 64, c * KC, b);
      tma_load_3d(sV, &tmV, bar_kv, h * 64, c * KC, b);
      if (c == 0) mbar_wait(bar_q, 0);
      mbar_wait(bar_kv, ph_kv);
      tcgen05_fence_after();
      const uint32_t idesc = make_idesc_f16(FMT, 0, 0, 128, nk16);
      const uint32_t qs0 = smem_u32(sQ), k0 = smem_u32(sK), g0 = smem_u32(sdO), v0 = smem_u32(sV);
#pragma unroll
      for (int k = 0; k < 4; ++k)   // S = Q K^T -> cols [0,128)
        umma_f16(tmem, make_smem_desc_sw128(qs0 + k * 32, 0, 1024), make_smem_desc_sw128(k0 + k * 32, 0, 1024), idesc,
                 k ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < 4; ++k)   // dP = dO V^T -> cols [128,256)
        umma_f16(tmem + 128, make_smem_desc_sw128(g0 + k * 32, 0, 1024), make_smem_desc_sw128(v0 + k * 32, 0, 1024),
                 idesc, k ? 1u : 0u);
      umma_commit(bar_mma);
    }
    ph_kv ^= 1;
    {
      const int key = c * KC + tid;
      kml[tid] = key < p.Nk ? (krow ? __ldg(krow + key) * LOG2E : 0.f) : -INFINITY;
    }
    mbar_wait(bar_mma, ph_mma);
    ph_mma ^= 1;
    tcgen05_fence_after();
    __syncthreads();

    // one branch-free pass over 16-key blocks: P = exp2(S - lse), P~ = dropout(P), dS = P~ dP - P delta (rows past Nq see
    // zero-filled Q / dO rows: finite P, dP = 0, delta = 0)
#pragma unroll 1
    for (int kb = 0; kb * 16 < nk16; ++kb) {
      uint32_t rs[16], rp[16];
      float kv[16], pt[16], ds[16];
      tmem_ld_32x32b_x16(t_row + kb * 16, rs);
      tmem_ld_32x32b_x16(t_row + 128 + kb * 16, rp);
      load_kv16c(kml, brow, kb * 16, c * KC + kb * 16, p.Nk, kv);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float pr = ex2_approx(fmaf(__uint_as_float(rs[j]), sl2, kv[j]) - lse);
        pt[j] = pr;
        ds[j] = -pr * delta;
      }
      if (drop) drop_mul16(pt, rowkey, c * KC + kb * 16, thr16, keep_scale);
#pragma unroll
      for (int j = 0; j < 16; ++j) ds[j] = fmaf(pt[j], __uint_as_float(rp[j]), ds[j]);
      if (dbrow && rv) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (c * KC + kb * 16 + j < p.Nk) atomicAdd(dbrow + c * KC + kb * 16 + j, ds[j]);
      }
      if (KV_OUT) store_row16<T>(sP, tid, kb * 16, pt);
      store_row16<T>(sdS, tid, kb * 16, ds);
    }
    fence_proxy_async();
    tcgen05_fence_before();
    __syncthreads();
    if (tid == 0) {
      tcgen05_fence_after();
      const uint32_t pa = smem_u32(sP), sa = smem_u32(sdS), qs0 = smem_u32(sQ), k0 = smem_u32(sK), g0 = smem_u32(sdO);
      const uint32_t id_tt = make_idesc_f16(FMT, 1, 1, 128, 64);
      const uint32_t id_nt = make_idesc_f16(FMT, 0, 1, 128, 64);
      if (KV_OUT) {
        for (int kq = 0; kq * 16 < nq16; ++kq)   // dV[key, d] = sum_q P[q,key] dO[q,d]   -> cols [0,64)
          umma_f16(tmem, make_smem_desc_sw128(pa + kq * 2048, TILE, 1024), make_smem_desc_sw128(g0 + kq * 2048, 8192, 1024),
                   id_tt, kq ? 1u : 0u);
        for (int kq = 0; kq * 16 < nq16; ++kq)   // dK[key, d] = sum_q dS[q,key] Q[q,d]  -> cols [64,128)
          umma_f16(tmem + 64, make_smem_desc_sw128(sa + kq * 2048, TILE, 1024),
                   make_smem_desc_sw128(qs0 + kq * 2048, 8192, 1024), id_tt, kq ? 1u : 0u);
      }
      for (int kk = 0; kk * 16 < nk16; ++kk)   // dQ[q, d] = sum_key dS[q,key] K[key,d]  -> cols [128,192)
        umma_f16(tmem + 128, make_smem_desc_sw128(sa + (kk >> 2) * TILE + (kk & 3) * 32, 0, 1024),
                 make_smem_desc_sw128(k0 + kk * 2048, 8192, 1024), id_nt, kk ? 1u : 0u);
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, ph_mma);
    ph_mma ^= 1;
    tcgen05_fence_after();

    {
      if (KV_OUT) {
        const int key = c * KC + tid;   // this thread's TMEM lane is a KEY for dV / dK
        const bool kv = tid < nk;
        float t[64];
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint32_t rr[32];
          tmem_ld_32x32b_x32(t_row + g * 32, rr);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) t[g * 32 + j] = __uint_as_float(rr[j]);
        }
        if (kv) store_global_row64<T>(reinterpret_cast<T*>(p.dV) + (long long)b * p.sbv + (long long)key * p.ldv + h * 64, t, 1.f);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint32_t rr[32];
          tmem_ld_32x32b_x32(t_row + 64 + g * 32, rr);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) t[g * 32 + j] = __uint_as_float(rr[j]);
        }
        if (kv) store_global_row64<T>(reinterpret_cast<T*>(p.dK) + (long long)b * p.sbk + (long long)key * p.ldk + h * 64, t, p.scale);
      }
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        uint32_t rr[32];
        tmem_ld_32x32b_x32(t_row + 128 + g * 32, rr);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) dq[g * 32 + j] += __uint_as_float(rr[j]);
      }
    }
    tcgen05_fence_before();
    __syncthreads();
  }
  if (rv) store_global_row64<T>(reinterpret_cast<T*>(p.dQ) + (long long)b * p.sbq + (long long)r * p.ldq + h * 64, dq, p.scale);
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<TMEM_COLS>(tmem);
}

// dK / dV of one 128-key chunk for Nq > 128: one CTA per (head, batch, key chunk) walks the query tiles; per tile
// S = Q_t K_c^T and dP = dO_t V_c^T are recomputed, P~ / dS go to shared memory, and dV += P~^T dO_t, dK += dS^T Q_t
// ACCUMULATE in TMEM across the tiles (columns [256,320) / [320,384)); one store at the end.  Thread r owns query row
// qt*128 + r while the scores are processed and key row c*128 + r in the epilogue.
constexpr int TMEM_COLS_KV = 512;

template <typename T>
__global__ void __launch_bounds__(TC_THREADS)
attn_bwd_tc_kv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO, TcArgs p) {
  extern __shared__ uint8_t smem_raw[];
  Smem sm(smem_raw);
  uint8_t* sQ = sm.base;
  uint8_t* sdO = sQ + TILE;
  uint8_t* sK = sdO + TILE;
  uint8_t* sV = sK + TILE;
  uint8_t* sP = sV + TILE;
  uint8_t* sdS = sP + 2 * TILE;
  uint64_t* bar_q = reinterpret_cast<uint64_t*>(sdS + 2 * TILE);
  uint64_t* bar_kv = bar_q + 1;
  uint64_t* bar_mma = bar_q + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_q + 3);
  float* kml = reinterpret_cast<float*>(bar_q + 8);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int h = blockIdx.x, b = blockIdx.y, c = blockIdx.z;
  if (tid == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmdO);
    mbar_init(bar_q, 1); mbar_init(bar_kv, 1); mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<TMEM_COLS_KV>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_row = tmem + ((uint32_t)(warp * 32) << 16);
  constexpr int FMT = UmmaFmt<T>::value;
  constexpr uint32_t COL_S = 0, COL_DP = 128, COL_DV = 256, COL_DK = 320;

  const int nk = min(KC, p.Nk - c * KC), nk16 = (nk + 15) & ~15;
  if (tid == 0) {
    mbar_arrive_expect_tx(bar_kv, 2 * TILE);
    tma_load_3d(sK, &tmK, bar_kv, h * 64, c * KC, b);
    tma_load_3d(sV, &tmV, bar_kv, h * 64, c * KC, b);
  }
  const float* krow = p.kmask ? p.kmask + (long long)b * p.Nk : nullptr;
  {
    const int key = c * KC + tid;
    kml[tid] = key < p.Nk ? (krow ? __ldg(krow + key) * LOG2E : 0.f) : -INFINITY;
  }
  const unsigned long long seed = p.drop_p > 0.f ? eff_seed(p.drop_seed, p.drop_seed_ptr) : 0ull;
  const float sl2 = p.scale * LOG2E;
  const bool drop = p.drop_p > 0.f;
  const uint32_t thr16 = drop_thr16(p.drop_p);
  const float keep_scale = drop ? 1.f / (1.f - p.drop_p) : 1.f;
  const int nqt = (p.Nq + 127) / 128;
  uint32_t ph_q = 0, ph_mma = 0;
  for (int qt = 0; qt < nqt; ++qt) {
    const int q0 = qt * 128;
    const int nq16 = (min(p.Nq - q0, 128) + 15) & ~15;
    if (tid == 0) {
      mbar_arrive_expect_tx(bar_q, 2 * TILE);
      tma_load_3d(sQ, &tmQ, bar_q, h * 64, q0, b);
      tma_load_3d(sdO, &tmdO, bar_q, h * 64, q0, b);
    }
    // this thread's query row of the tile: lse and delta = sum_d dO O
    const int r = q0 + tid;
    const bool rv = r < p.Nq;
    const float* brow = p.bias ? p.bias + ((long long)b * p.Nq + (rv ? r : 0)) * p.Nk : nullptr;
    float delta = 0.f, lse = 0.f;
    if (rv) {
      const long long off = (long long)b * p.sbo + (long long)r * p.ldo + h * 64;
      const uint4* po = reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(p.O) + off);
      const uint4* pg = reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(p.dO) + off);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint4 a = __ldg(po + q), g = __ldg(pg + q);
        float2 x, y;
        x = unpack2<T>(a.x); y = unpack2<T>(g.x); delta += x.x * y.x + x.y * y.y;
        x = unpack2<T>(a.y); y = unpack2<T>(g.y); delta += x.x * y.x + x.y * y.y;
        x = unpack2<T>(a.z); y = unpack2<T>(g.z); delta += x.x * y.x + x.y * y.y;
        x = unpack2<T>(a.w); y = unpack2<T>(g.w); delta += x.x * y.x + x.y * y.y;
      }
      lse = p.lse[((long long)b * p.heads + h) * p.Nq + r] * LOG2E;
      if (lse == -INFINITY) lse = INFINITY;
    }
    if (tid == 0) {
      if (qt == 0) mbar_wait(bar_kv, 0);
      mbar_wait(bar_q, ph_q);
      tcgen05_fence_after();
      const uint32_t idesc = make_idesc_f16(FMT, 0, 0, 128, nk16);
      const uint32_t qs0 = smem_u32(sQ), k0 = smem_u32(sK), g0 = smem_u32(sdO), v0 = smem_u32(sV);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_f16(tmem + COL_S, make_smem_desc_sw128(qs0 + k * 32, 0, 1024), make_smem_desc_sw128(k0 + k * 32, 0, 1024),
                 idesc, k ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_f16(tmem + COL_DP, make_smem_desc_sw128(g0 + k * 32, 0, 1024), make_smem_desc_sw128(v0 + k * 32, 0, 1024),
                 idesc, k ? 1u : 0u);
      umma_commit(bar_mma);
    }
    ph_q ^= 1;
    mbar_wait(bar_mma, ph_mma);
    ph_mma ^= 1;
    tcgen05_fence_after();
    __syncthreads();     // kml visible (first tile)
#pragma unroll 1
    for (int kb = 0; kb * 16 < nk16; ++kb) {
      uint32_t rs[16], rp[16];
      float kv[16], pt[16], ds[16];
      tmem_ld_32x32b_x16(t_row + COL_S + kb * 16, rs);
      tmem_ld_32x32b_x16(t_row + COL_DP + kb * 16, rp);
      load_kv16c(kml, brow, kb * 16, c * KC + kb * 16, p.Nk, kv);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float pr = ex2_approx(fmaf(__uint_as_float(rs[j]), sl2, kv[j]) - lse);
        pt[j] = pr;
        ds[j] = -pr * delta;
      }
      if (drop) drop_mul16(pt, attn_drop_rowkey(seed, b, p.heads, h, p.Nq, r), c * KC + kb * 16, thr16, keep_scale);
#pragma unroll
      for (int j = 0; j < 16; ++j) ds[j] = fmaf(pt[j], __uint_as_float(rp[j]), ds[j]);
      store_row16<T>(sP, tid, kb * 16, pt);
      store_row16<T>(sdS, tid, kb * 16, ds);
    }
    fence_proxy_async();
    tcgen05_fence_before();
    __syncthreads();
    if (tid == 0) {
      tcgen05_fence_after();
      const uint32_t pa = smem_u32(sP), sa = smem_u32(sdS), qs0 = smem_u32(sQ), g0 = smem_u32(sdO);
      const uint32_t id_tt = make_idesc_f16(FMT, 1, 1, 128, 64);
      for (int kq = 0; kq * 16 < nq16; ++kq)   // dV[key, d] += sum_q P~[q,key] dO[q,d]
        umma_f16(tmem + COL_DV, make_smem_desc_sw128(pa + kq * 2048, TILE, 1024),
                 make_smem_desc_sw128(g0 + kq * 2048, 8192, 1024), id_tt, (qt | kq) ? 1u : 0u);
      for (int kq = 0; kq * 16 < nq16; ++kq)   // dK[key, d] += sum_q dS[q,key] Q[q,d]
        umma_f16(tmem + COL_DK, make_smem_desc_sw128(sa + kq * 2048, TILE, 1024),
                 make_smem_desc_sw128(qs0 + kq * 2048, 8192, 1024), id_tt, (qt | kq) ? 1u : 0u);
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, ph_mma);    // Q / dO / P / dS tiles may be overwritten by the next query tile
    ph_mma ^= 1;
    tcgen05_fence_after();
    tcgen05_fence_before();
    __syncthreads();
  }
  {
    const int key = c * KC + tid;
    const bool kv = tid < nk;
    float t[64];
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      uint32_t rr[32];
      tmem_ld_32x32b_x32(t_row + COL_DV + g * 32, rr);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) t[g * 32 + j] = __uint_as_float(rr[j]);
    }
    if (kv) store_global_row64<T>(reinterpret_cast<T*>(p.dV) + (long long)b * p.sbv + (long long)key * p.ldv + h * 64, t, 1.f);
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      uint32_t rr[32];
      tmem_ld_32x32b_x32(t_row + COL_DK + g * 32, rr);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) t[g * 32 + j] = __uint_as_float(rr[j]);
    }
    if (kv) store_global_row64<T>(reinterpret_cast<T*>(p.dK) + (long long)b * p.sbk + (long long)key * p.ldk + h * 64, t, p.scale);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<TMEM_COLS_KV>(tmem);
}

int fill(const goat_attn_args* a, TcArgs* t) {
  t->B = a->B; t->heads = a->heads; t->Nq = a->Nq; t->Nk = a->Nk;
  t->kmask = a->kmask; t->bias = a->bias; t->scale = a->scale; t->lse = a->lse;
  t->drop_p = a->drop_p; t->drop_seed = a->drop_seed;
  t->drop_seed_ptr = reinterpret_cast<const unsigned long long*>(a->drop_seed_ptr);
  t->O = a->O; t->ldo = a->ldo; t->sbo = a->sbo;
  t->dO = a->dO;
  t->dQ = a->dQ; t->ldq = a->ldq; t->sbq = a->sbq;
  t->dK = a->dK; t->ldk = a->ldk; t->sbk = a->sbk;
  t->dV = a->dV; t->ldv = a->ldv; t->sbv = a->sbv;
  t->dbias = a->dbias;
  return GOAT_OK;
}

template <typename T>
int fwd_launch(const goat_attn_args* a, cudaStream_t st) {
  static bool cfg = false;
  if (!cfg) {
    GOAT_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
    cfg = true;
  }
  CUtensorMap tq, tk, tv;
  int rc;
  const uint64_t cols = (uint64_t)a->heads * 64;
  if ((rc = make_tmap3(&tq, a->dtype, a->Q, cols, a->Nq, a->B, a->ldq, a->sbq, 64, 128))) return rc;
  if ((rc = make_tmap3(&tk, a->dtype, a->K, cols, a->Nk, a->B, a->ldk, a->sbk, 64, 128))) return rc;
  if ((rc = make_tmap3(&tv, a->dtype, a->V, cols, a->Nk, a->B, a->ldv, a->sbv, 64, 128))) return rc;
  TcArgs t;
  fill(a, &t);
  dim3 grid(a->heads, a->B, (a->Nq + 127) / 128);
  attn_fwd_tc_kernel<T><<<grid, TC_THREADS, FWD_SMEM, st>>>(tq, tk, tv, t);
  GOAT_LAUNCH_CHECK();
  return GOAT_OK;
}

template <typename T>
int bwd_launch(const goat_attn_args* a, cudaStream_t st) {
  static bool cfg = false;
  if (!cfg) {
    GOAT_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    GOAT_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    GOAT_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kv_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    cfg = true;
  }
  CUtensorMap tq, tk, tv, tg;
  int rc;
  const uint64_t cols = (uint64_t)a->heads * 64;
  if ((rc = make_tmap3(&tq, a->dtype, a->Q, cols, a->Nq, a->B, a->ldq, a->sbq, 64, 128))) return rc;
  if ((rc = make_tmap3(&tk, a->dtype, a->K, cols, a->Nk, a->B, a->ldk, a->sbk, 64, 128))) return rc;
  if ((rc = make_tmap3(&tv, a->dtype, a->V, cols, a->Nk, a->B, a->ldv, a->sbv, 64, 128))) return rc;
  if ((rc = make_tmap3(&tg, a->dtype, a->dO, cols, a->Nq, a->B, a->ldo, a->sbo, 64, 128))) return rc;
  TcArgs t;
  fill(a, &t);
  if (a->Nq <= 128) {
    dim3 grid(a->heads, a->B);
    attn_bwd_tc_kernel<T, true><<<grid, TC_THREADS, BWD_SMEM, st>>>(tq, tk, tv, tg, t);
    GOAT_LAUNCH_CHECK();
    return GOAT_OK;
  }
  // long query sequences (RxR / 512-token instructions): dQ per query tile, then dK / dV per key chunk
  dim3 gq(a->heads, a->B, (a->Nq + 127) / 128);
  attn_bwd_tc_kernel<T, false><<<gq, TC_THREADS, BWD_SMEM, st>>>(tq, tk, tv, tg, t);
  GOAT_LAUNCH_CHECK();
  dim3 gk(a->heads, a->B, (a->Nk + KC - 1) / KC);
  attn_bwd_tc_kv_kernel<T><<<gk, TC_THREADS, BWD_SMEM, st>>>(tq, tk, tv, tg, t);
  GOAT_LAUNCH_CHECK();
  return GOAT_OK;
}

bool ok16(const void* p, long long ld, long long sb) {
  return aligned16(p) && (ld % 8) == 0 && (sb % 8) == 0 && ld >= 64 && sb >= 0;
}

}  // namespace

// The tensor-core path takes 16-bit operands and TMA-compatible strides; any number of query rows / keys (128-row
// query tiles on grid.z, 128-key chunks in the kernel).
bool attn_tc_eligible(const goat_attn_args* a, bool bwd) {
  if (a->dtype != GOAT_F16 && a->dtype != GOAT_BF16) return false;
  if (a->Nq < 1 || a->Nk < 1 || a->D != 64) return false;
  if (a->heads > 65535 || a->B > 65535 || a->Nq > 65535 * 128 || a->Nk > 65535 * 128) return false;
  if (a->B > 1 && (a->sbq <= 0 || a->sbk <= 0 || a->sbv <= 0 || a->sbo <= 0)) return false;  // broadcast operands
  if (!ok16(a->Q, a->ldq, a->sbq) || !ok16(a->K, a->ldk, a->sbk) || !ok16(a->V, a->ldv, a->sbv) ||
      !ok16(a->O, a->ldo, a->sbo))
    return false;
  if (bwd && (!ok16(a->dO, a->ldo, a->sbo) || !aligned16(a->dQ) || !aligned16(a->dK) || !aligned16(a->dV))) return false;
  return true;
}

bool attn_pipe_eligible(const goat_attn_args* a);
int attn_fwd_pipe(const goat_attn_args* a, cudaStream_t st);
int attn_bwd_pipe(const goat_attn_args* a, cudaStream_t st);

// GOAT_ATTN_LEGACY=1 keeps the one-CTA-per-(head,batch) kernels of this file for every shape (A/B comparisons)
static bool legacy_only() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GOAT_ATTN_LEGACY");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

int attn_fwd_tc(const goat_attn_args* a, cudaStream_t st) {
  if (!legacy_only() && attn_pipe_eligible(a)) return attn_fwd_pipe(a, st);
  return a->dtype == GOAT_F16 ? fwd_launch<__half>(a, st) : fwd_launch<__nv_bfloat16>(a, st);
}
int attn_bwd_tc(const goat_attn_args* a, cudaStream_t st) {
  if (!legacy_only() && attn_pipe_eligible(a)) return attn_bwd_pipe(a, st);
  return a->dtype == GOAT_F16 ? bwd_launch<__half>(a, st) : bwd_launch<__nv_bfloat16>(a, st);
}

}  // namespace goat
