// Persistent, warp-specialised attention core for the GOAT shapes (Nq <= 128 queries, Nk <= 128 keys, d = 64):
// one CTA per SM walks the (batch, head) work items; the fixed per-item latency chain of the one-CTA-per-item
// kernels in attention_tc.cu (TMA -> MMA -> softmax -> MMA -> store, 26 % of the training step at 6-12 % issue
// utilisation, profiles/r01_ncu_attention_full.txt) is overlapped across items instead:
//
//   warp 0      TMA producer: Q / K / V (/ dO) head slices of item i+1, i+2 into a smem ring while item i computes
//   warp 1      MMA issuer (one lane): S = Q K^T (and dP = dO V^T) of item i+1 is issued right after the second
//               MMA group of item i, so it runs under the epilogue of item i
//   warps 2-9   256 compute threads, TWO per query row (each takes alternate 32-key groups): softmax on the TMEM
//               rows in the log2 domain, P (and dS) written once as 16-bit operands into 128B-swizzled smem,
//               then the epilogue (O / l, or dV, dK, dQ) straight from TMEM to global memory
//
// forward : O = softmax(scale Q K^T + kmask + bias) V, lse saved
// backward: P from lse;  delta_r = sum_d dO O (= sum_k P~ dP; computed up front, under the score MMAs);  dS = P~ dP - P delta;
//           dV = P~^T dO, dK = scale dS^T Q, dQ = scale dS K            (P~ = dropout-masked, rescaled P)
// TMEM: forward  O [0,64) | S [64,192);   backward dV [0,64) | dK [64,128) | dQ [128,192) | S [192,320) | dP [320,448).
// Every contraction is a tcgen05.mma with M = 128; operand layouts / descriptors are those of attention_tc.cu.
#include <cuda.h>
#include <stdlib.h>

#include "attn_common.cuh"

namespace goat {

int make_tmap3(CUtensorMap* tm, int dtype, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t ld_elems,
               uint64_t sb_elems, uint32_t box0, uint32_t box1);

namespace {

using namespace attn;

constexpr int P_THREADS = 320;       // warp 0 TMA, warp 1 MMA, warps 2..9 compute
constexpr int P_COMPUTE = 256;
constexpr int B_STAGES = 2;

struct PArgs {
  int B, heads, Nq, Nk, n_items;
  const float* kmask;
  const float* bias;
  float scale;
  float* lse;
  float drop_p;
  unsigned long long drop_seed;
  const unsigned long long* drop_seed_ptr;
  void* O; int ldo; long long sbo;
  void* dQ; int ldq; long long sbq;
  void* dK; int ldk; long long sbk;
  void* dV; int ldv; long long sbv;
  float* dbias;
};

__device__ __forceinline__ void compute_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ float load_kml(const PArgs& p, int b, int key) {
  if (key >= p.Nk) return -INFINITY;
  return p.kmask ? __ldg(p.kmask + (long long)b * p.Nk + key) * LOG2E : 0.f;
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
// Two configurations: <3 stages, 1 CTA per SM> (item i+1, i+2 prefetched under item i) and <1 stage, 2 CTAs per SM>
// (84 KB each, 256 TMEM columns each): two independent item chains per SM hide each other's TMA -> MMA -> softmax -> MMA
// -> store latency, which is what bounds these 37..80-row items (ncu: one chain in flight, issue slots 25 % busy).
template <int STAGES>
constexpr int f_smem() { return STAGES * 3 * TILE + 2 * TILE + 128 + (2 * 128 + 2 * 128 + 2 * 128) * 4 + 1024; }

template <typename T, int F_STAGES, int MINB>
__global__ void __launch_bounds__(P_THREADS, MINB)
attn_fwd_pipe_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, PArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_align1024(smem_raw);
  uint8_t* sP = base + F_STAGES * 3 * TILE;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sP + 2 * TILE);
  uint64_t* empty_bar = full_bar + F_STAGES;
  uint64_t* s_full = empty_bar + F_STAGES;
  uint64_t* p_full = s_full + 1;
  uint64_t* o_full = p_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);
  float* kml = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + 128);   // [2][128]
  float* xm = kml + 256;                                                                // [2][128] row-max exchange
  float* xl = xm + 256;                                                                 // [2][128] row-sum exchange

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
#pragma unroll
    for (int s = 0; s < F_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(s_full, 1); mbar_init(p_full, 1); mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();                 // setup above overlapped the previous kernel's tail
  pdl_launch_dependents();
  constexpr int FMT = UmmaFmt<T>::value;
  constexpr uint32_t COL_O = 0, COL_S = 64;
  const int n_local = (p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int nk16 = (p.Nk + 15) & ~15;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < n_local; ++i) {
        const int item = blockIdx.x + i * gridDim.x;
        const int b = item / p.heads, h = item % p.heads;
        const int s = i % F_STAGES;
        mbar_wait(&empty_bar[s], (((uint32_t)(i / F_STAGES)) & 1) ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], 3 * TILE);
        uint8_t* st = base + s * 3 * TILE;
        tma_load_3d(st, &tmQ, &full_bar[s], h * 64, 0, b);
        tma_load_3d(st + TILE, &tmK, &full_bar[s], h * 64, 0, b);
        tma_load_3d(st + 2 * TILE, &tmV, &full_bar[s], h * 64, 0, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_f16(FMT, 0, 0, 128, nk16);
      const uint32_t idesc_o = make_idesc_f16(FMT, 0, 1, 128, 64);
      auto mma_s = [&](int i) {     // S = Q K^T of local item i
        const int s = i % F_STAGES;
        mbar_wait(&full_bar[s], ((uint32_t)(i / F_STAGES)) & 1);
        tcgen05_fence_after();
        const uint32_t a0 = smem_u32(base + s * 3 * TILE), b0 = a0 + TILE;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tmem + COL_S, make_smem_desc_sw128(a0 + k * 32, 0, 1024), make_smem_desc_sw128(b0 + k * 32, 0, 1024),
                   idesc_s, k ? 1u : 0u);
        umma_commit(s_full);
      };
      if (n_local > 0) mma_s(0);
      for (int i = 0; i < n_local; ++i) {
        const int s = i % F_STAGES;
        mbar_wait(p_full, (uint32_t)i & 1);     // softmax(i) done: P in smem, S columns free, O(i-1) consumed
        tcgen05_fence_after();
        const uint32_t a0 = smem_u32(sP), b0 = smem_u32(base + s * 3 * TILE + 2 * TILE);
        for (int kk = 0; kk * 16 < nk16; ++kk)
          umma_f16(tmem + COL_O, make_smem_desc_sw128(a0 + (kk >> 2) * TILE + (kk & 3) * 32, 0, 1024),
                   make_smem_desc_sw128(b0 + kk * 2048, 8192, 1024), idesc_o, kk ? 1u : 0u);
        umma_commit(o_full);
        umma_commit(&empty_bar[s]);
        if (i + 1 < n_local) mma_s(i + 1);      // runs under the epilogue of item i
      }
    }
  } else {
    const int cidx = threadIdx.x - 64;        // 0..255
    const int lg = warp & 3;                  // TMEM lane group this warp may access
    const int ch = (warp - 2) >> 2;           // which alternate 32-key groups this thread takes
    const int r = lg * 32 + lane;
    const bool rv = r < p.Nq;
    const bool live = lg * 32 < p.Nq;         // warp-uniform
    const uint32_t t_lane = tmem + ((uint32_t)(lg * 32) << 16);
    const int nblk = nk16 >> 4;               // 16-key blocks the S MMA writes (columns past nk16 hold stale data)
    const float sl2 = p.scale * LOG2E;
    const bool drop = p.drop_p > 0.f;
    const unsigned long long seed = drop ? eff_seed(p.drop_seed, p.drop_seed_ptr) : 0ull;
    const uint32_t thr16 = drop_thr16(p.drop_p);
    const float keep_scale = drop ? 1.f / (1.f - p.drop_p) : 1.f;
    if (n_local > 0 && cidx < 128) kml[cidx] = load_kml(p, (int)blockIdx.x / p.heads, cidx);
    compute_bar();
    for (int i = 0; i < n_local; ++i) {
      const int item = blockIdx.x + i * gridDim.x;
      const int b = item / p.heads, h = item % p.heads;
      const int par = i & 1;
      const float* km = kml + par * 128;
      const bool has_next = i + 1 < n_local;
      float km_next = 0.f;
      if (has_next && cidx < 128) km_next = load_kml(p, (item + (int)gridDim.x) / p.heads, cidx);
      const float* brow = p.bias ? p.bias + ((long long)b * p.Nq + (rv ? r : 0)) * p.Nk : nullptr;
      mbar_wait(s_full, (uint32_t)i & 1);
      tcgen05_fence_after();
      // pass 1: row maximum over this thread's 16-key blocks.  Branch-free per element: masked / out-of-range keys
      // carry -inf in the staged mask, rows past Nq see S = 0 (zero-filled Q rows) and are never stored.
      float m = -INFINITY;
      if (live) {
#pragma unroll 1
        for (int kb = ch; kb < nblk; kb += 2) {
          uint32_t rr[16];
          float kv[16];
          tmem_ld_32x32b_x16(t_lane + COL_S + kb * 16, rr);
          load_kv16(km, brow, kb * 16, p.Nk, kv);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) m = fmaxf(m, fmaf(__uint_as_float(rr[j]), sl2, kv[j]));
        }
      }
      xm[ch * 128 + r] = m;
      compute_bar();
      m = fmaxf(m, xm[(ch ^ 1) * 128 + r]);
      const float mref = (m == -INFINITY) ? 0.f : m;
      // pass 2: P = exp2(S - m) -> smem (16-bit), partial row sums
      float l = 0.f;
      if (live) {
#pragma unroll 1
        for (int kb = ch; kb < nblk; kb += 2) {
          uint32_t rr[16];
          float kv[16], pv[16];
          tmem_ld_32x32b_x16(t_lane + COL_S + kb * 16, rr);
          load_kv16(km, brow, kb * 16, p.Nk, kv);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float e = ex2_approx(fmaf(__uint_as_float(rr[j]), sl2, kv[j]) - mref);
            l += e;
            pv[j] = e;
          }
          if (drop) drop_mul16(pv, attn_drop_rowkey(seed, b, p.heads, h, p.Nq, r), kb * 16, thr16, keep_scale);
          store_row16<T>(sP, r, kb * 16, pv);
        }
      }
      xl[ch * 128 + r] = l;
      if (has_next && cidx < 128) kml[(par ^ 1) * 128 + cidx] = km_next;
      fence_proxy_async();
      tcgen05_fence_before();
      compute_bar();
      if (cidx == 0) mbar_arrive(p_full);
      // epilogue: O / l -> global, lse
      mbar_wait(o_full, (uint32_t)i & 1);
      tcgen05_fence_after();
      if (live) {
        uint32_t rr[32];
        tmem_ld_32x32b_x32(t_lane + COL_O + ch * 32, rr);
        tmem_ld_wait();
        if (rv) {
          const float lt = xl[r] + xl[128 + r];
          const float inv = lt > 0.f ? 1.f / lt : 0.f;
          T* dst = reinterpret_cast<T*>(p.O) + (long long)b * p.sbo + (long long)r * p.ldo + h * 64 + ch * 32;
          store_global_row32<T>(dst, rr, inv);
          if (ch == 0 && p.lse) p.lse[((long long)b * p.heads + h) * p.Nq + r] = lt > 0.f ? (m + __log2f(lt)) * LN2 : -INFINITY;
        }
      }
      tcgen05_fence_before();
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem);
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
constexpr int B_SMEM = B_STAGES * 4 * TILE + 4 * TILE + 128 + (2 * 128 + 2 * 128) * 4 + 1024;

template <typename T>
__global__ void __launch_bounds__(P_THREADS, 1)
attn_bwd_pipe_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO, PArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_align1024(smem_raw);
  uint8_t* sP = base + B_STAGES * 4 * TILE;
  uint8_t* sdS = sP + 2 * TILE;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sdS + 2 * TILE);
  uint64_t* empty_bar = full_bar + B_STAGES;
  uint64_t* s_full = empty_bar + B_STAGES;
  uint64_t* p_full = s_full + 1;
  uint64_t* o_full = p_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);
  float* kml = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + 128);   // [2][128]
  float* xd = kml + 256;                                                                // [2][128] delta exchange

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmdO);
#pragma unroll
    for (int s = 0; s < B_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(s_full, 1); mbar_init(p_full, 1); mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();                 // setup above overlapped the previous kernel's tail
  pdl_launch_dependents();
  constexpr int FMT = UmmaFmt<T>::value;
  constexpr uint32_t COL_DV = 0, COL_DK = 64, COL_DQ = 128, COL_S = 192, COL_DP = 320;
  const int n_local = (p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int nk16 = (p.Nk + 15) & ~15;
  const int nq16 = (p.Nq + 15) & ~15;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < n_local; ++i) {
        const int item = blockIdx.x + i * gridDim.x;
        const int b = item / p.heads, h = item % p.heads;
        const int s = i % B_STAGES;
        mbar_wait(&empty_bar[s], (((uint32_t)(i / B_STAGES)) & 1) ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], 4 * TILE);
        uint8_t* st = base + s * 4 * TILE;
        tma_load_3d(st, &tmQ, &full_bar[s], h * 64, 0, b);
        tma_load_3d(st + TILE, &tmdO, &full_bar[s], h * 64, 0, b);
        tma_load_3d(st + 2 * TILE, &tmK, &full_bar[s], h * 64, 0, b);
        tma_load_3d(st + 3 * TILE, &tmV, &full_bar[s], h * 64, 0, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_f16(FMT, 0, 0, 128, nk16);
      const uint32_t id_tt = make_idesc_f16(FMT, 1, 1, 128, 64);
      const uint32_t id_nt = make_idesc_f16(FMT, 0, 1, 128, 64);
      auto mma_s = [&](int i) {     // S = Q K^T, dP = dO V^T of local item i
        const int s = i % B_STAGES;
        mbar_wait(&full_bar[s], ((uint32_t)(i / B_STAGES)) & 1);
        tcgen05_fence_after();
        const uint32_t q0 = smem_u32(base + s * 4 * TILE), g0 = q0 + TILE, k0 = q0 + 2 * TILE, v0 = q0 + 3 * TILE;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tmem + COL_S, make_smem_desc_sw128(q0 + k * 32, 0, 1024), make_smem_desc_sw128(k0 + k * 32, 0, 1024),
                   idesc_s, k ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tmem + COL_DP, make_smem_desc_sw128(g0 + k * 32, 0, 1024), make_smem_desc_sw128(v0 + k * 32, 0, 1024),
                   idesc_s, k ? 1u : 0u);
        umma_commit(s_full);
      };
      if (n_local > 0) mma_s(0);
      for (int i = 0; i < n_local; ++i) {
        const int s = i % B_STAGES;
        mbar_wait(p_full, (uint32_t)i & 1);     // P / dS in smem, S / dP columns free, outputs of item i-1 consumed
        tcgen05_fence_after();
        const uint32_t q0 = smem_u32(base + s * 4 * TILE), g0 = q0 + TILE, k0 = q0 + 2 * TILE;
        const uint32_t pa = smem_u32(sP), sa = smem_u32(sdS);
        for (int kq = 0; kq * 16 < nq16; ++kq)   // dV[key, d] = sum_q P~[q,key] dO[q,d]
          umma_f16(tmem + COL_DV, make_smem_desc_sw128(pa + kq * 2048, TILE, 1024),
                   make_smem_desc_sw128(g0 + kq * 2048, 8192, 1024), id_tt, kq ? 1u : 0u);
        for (int kq = 0; kq * 16 < nq16; ++kq)   // dK[key, d] = sum_q dS[q,key] Q[q,d]
          umma_f16(tmem + COL_DK, make_smem_desc_sw128(sa + kq * 2048, TILE, 1024),
                   make_smem_desc_sw128(q0 + kq * 2048, 8192, 1024), id_tt, kq ? 1u : 0u);
        for (int kk = 0; kk * 16 < nk16; ++kk)   // dQ[q, d] = sum_key dS[q,key] K[key,d]
          umma_f16(tmem + COL_DQ, make_smem_desc_sw128(sa + (kk >> 2) * TILE + (kk & 3) * 32, 0, 1024),
                   make_smem_desc_sw128(k0 + kk * 2048, 8192, 1024), id_nt, kk ? 1u : 0u);
        umma_commit(o_full);
        umma_commit(&empty_bar[s]);
        if (i + 1 < n_local) mma_s(i + 1);      // runs under the epilogue of item i
      }
    }
  } else {
    const int cidx = threadIdx.x - 64;
    const int lg = warp & 3;
    const int ch = (warp - 2) >> 2;
    const int r = lg * 32 + lane;
    const bool rv = r < p.Nq;
    const bool live_q = lg * 32 < nq16;       // warp-uniform: rows below nq16 must be written (zeros past Nq)
    const bool live_o = lg * 32 < max(p.Nq, p.Nk);
    const uint32_t t_lane = tmem + ((uint32_t)(lg * 32) << 16);
    const int nblk = nk16 >> 4;               // 16-key blocks the S / dP MMAs write
    const float sl2 = p.scale * LOG2E;
    const bool drop = p.drop_p > 0.f;
    const unsigned long long seed = drop ? eff_seed(p.drop_seed, p.drop_seed_ptr) : 0ull;
    const uint32_t thr16 = drop_thr16(p.drop_p);
    const float keep_scale = drop ? 1.f / (1.f - p.drop_p) : 1.f;
    float lse_next = 0.f;
    if (n_local > 0) {
      const int b0 = (int)blockIdx.x / p.heads, h0 = (int)blockIdx.x % p.heads;
      if (cidx < 128) kml[cidx] = load_kml(p, b0, cidx);
      if (rv) lse_next = p.lse[((long long)b0 * p.heads + h0) * p.Nq + r] * LOG2E;
    }
    compute_bar();
    for (int i = 0; i < n_local; ++i) {
      const int item = blockIdx.x + i * gridDim.x;
      const int b = item / p.heads, h = item % p.heads;
      const int par = i & 1;
      const float* km = kml + par * 128;
      const bool has_next = i + 1 < n_local;
      // a fully masked row has lse = -inf: +inf here makes every probability exp2(s - inf) = 0 instead of NaN
      const float lse = (lse_next == -INFINITY) ? INFINITY : lse_next;
      float km_next = 0.f;
      if (has_next) {
        const int nb = (item + (int)gridDim.x) / p.heads, nh = (item + (int)gridDim.x) % p.heads;
        if (cidx < 128) km_next = load_kml(p, nb, cidx);
        if (rv) lse_next = p.lse[((long long)nb * p.heads + nh) * p.Nq + r] * LOG2E;
      }
      const float* brow = p.bias ? p.bias + ((long long)b * p.Nq + (rv ? r : 0)) * p.Nk : nullptr;
      float* dbrow = p.dbias ? p.dbias + ((long long)b * p.Nq + (rv ? r : 0)) * p.Nk : nullptr;
      // delta_r = sum_d dO[r,d] O[r,d] (= sum_k P~ dP): each of the row's two threads takes 32 of the 64 channels -- dO from the
      // staged tile, O from global memory -- BEFORE waiting for the score MMAs, so this overlaps them; with delta in hand
      // one pass over the scores produces P~ and dS (no probability arrays held across a barrier, no spills)
      {
        const int s = i % B_STAGES;
        mbar_wait(&full_bar[s], ((uint32_t)(i / B_STAGES)) & 1);
        float part = 0.f;
        if (rv) {
          uint8_t* sdO_t = base + s * 4 * TILE + TILE;
          const uint4* po = reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(p.O) + (long long)b * p.sbo +
                                                           (long long)r * p.ldo + h * 64 + ch * 32);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 g = *reinterpret_cast<const uint4*>(sw_chunk(sdO_t, r, ch * 4 + q));
            const uint4 o = __ldg(po + q);
            float2 x, y;
            x = unpack2<T>(o.x); y = unpack2<T>(g.x); part += x.x * y.x + x.y * y.y;
            x = unpack2<T>(o.y); y = unpack2<T>(g.y); part += x.x * y.x + x.y * y.y;
            x = unpack2<T>(o.z); y = unpack2<T>(g.z); part += x.x * y.x + x.y * y.y;
            x = unpack2<T>(o.w); y = unpack2<T>(g.w); part += x.x * y.x + x.y * y.y;
          }
        }
        xd[ch * 128 + r] = part;
      }
      compute_bar();
      const float delta = xd[r] + xd[128 + r];
      mbar_wait(s_full, (uint32_t)i & 1);
      tcgen05_fence_after();
      // one pass: P = exp2(S - lse), P~ = dropout(P), dS = P~ dP - P delta.  Branch-free per element; rows past Nq see
      // zero-filled Q / dO rows (finite P, dP = 0, delta = 0) and are never stored.
#pragma unroll 1
      for (int t = 0; t < 4; ++t) {
        const int kb = ch + 2 * t;
        if (live_q && kb < nblk) {
          uint32_t rs[16], rp[16];
          float kv[16], pt[16], ds[16];
          tmem_ld_32x32b_x16(t_lane + COL_S + kb * 16, rs);
          tmem_ld_32x32b_x16(t_lane + COL_DP + kb * 16, rp);
          load_kv16(km, brow, kb * 16, p.Nk, kv);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float pr = ex2_approx(fmaf(__uint_as_float(rs[j]), sl2, kv[j]) - lse);
            pt[j] = pr;
            ds[j] = -pr * delta;
          }
          if (drop) drop_mul16(pt, attn_drop_rowkey(seed, b, p.heads, h, p.Nq, r), kb * 16, thr16, keep_scale);
#pragma unroll
          for (int j = 0; j < 16; ++j) ds[j] = fmaf(pt[j], __uint_as_float(rp[j]), ds[j]);
          if (dbrow && rv) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (kb * 16 + j < p.Nk) atomicAdd(dbrow + kb * 16 + j, ds[j]);
          }
          store_row16<T>(sP, r, kb * 16, pt);
          store_row16<T>(sdS, r, kb * 16, ds);
        }
      }
      if (has_next && cidx < 128) kml[(par ^ 1) * 128 + cidx] = km_next;
      fence_proxy_async();
      tcgen05_fence_before();
      compute_bar();
      if (cidx == 0) mbar_arrive(p_full);
      // epilogue: dV, dK (TMEM lane = key), dQ (TMEM lane = query); this thread takes 32 of the 64 channels
      mbar_wait(o_full, (uint32_t)i & 1);
      tcgen05_fence_after();
      if (live_o) {
        uint32_t rv_[32], rk_[32], rq_[32];
        tmem_ld_32x32b_x32(t_lane + COL_DV + ch * 32, rv_);
        tmem_ld_32x32b_x32(t_lane + COL_DK + ch * 32, rk_);
        tmem_ld_32x32b_x32(t_lane + COL_DQ + ch * 32, rq_);
        tmem_ld_wait();
        const long long co = h * 64 + ch * 32;
        if (r < p.Nk) {
          store_global_row32<T>(reinterpret_cast<T*>(p.dV) + (long long)b * p.sbv + (long long)r * p.ldv + co, rv_, 1.f);
          store_global_row32<T>(reinterpret_cast<T*>(p.dK) + (long long)b * p.sbk + (long long)r * p.ldk + co, rk_, p.scale);
        }
        if (rv) store_global_row32<T>(reinterpret_cast<T*>(p.dQ) + (long long)b * p.sbq + (long long)r * p.ldq + co, rq_, p.scale);
      }
      tcgen05_fence_before();
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

void fill(const goat_attn_args* a, PArgs* t) {
  t->B = a->B; t->heads = a->heads; t->Nq = a->Nq; t->Nk = a->Nk; t->n_items = a->B * a->heads;
  t->kmask = a->kmask; t->bias = a->bias; t->scale = a->scale; t->lse = a->lse;
  t->drop_p = a->drop_p; t->drop_seed = a->drop_seed;
  t->drop_seed_ptr = reinterpret_cast<const unsigned long long*>(a->drop_seed_ptr);
  t->O = a->O; t->ldo = a->ldo; t->sbo = a->sbo;
  t->dQ = a->dQ; t->ldq = a->ldq; t->sbq = a->sbq;
  t->dK = a->dK; t->ldk = a->ldk; t->sbk = a->sbk;
  t->dV = a->dV; t->ldv = a->ldv; t->sbv = a->sbv;
  t->dbias = a->dbias;
}

static int env_flag(const char* name, int dflt) {
  const char* v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}

template <typename T>
int fwd_launch(const goat_attn_args* a, cudaStream_t st) {
  static bool cfg = false;
  static const int two = env_flag("GOAT_ATTN_FWD_2CTA", 1);
  if (!cfg) {
    GOAT_CUDA(cudaFuncSetAttribute(attn_fwd_pipe_kernel<T, 3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, f_smem<3>()));
    GOAT_CUDA(cudaFuncSetAttribute(attn_fwd_pipe_kernel<T, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, f_smem<1>()));
    cfg = true;
  }
  CUtensorMap tq, tk, tv;
  int rc;
  const uint64_t cols = (uint64_t)a->heads * 64;
  if ((rc = make_tmap3(&tq, a->dtype, a->Q, cols, a->Nq, a->B, a->ldq, a->sbq, 64, 128))) return rc;
  if ((rc = make_tmap3(&tk, a->dtype, a->K, cols, a->Nk, a->B, a->ldk, a->sbk, 64, 128))) return rc;
  if ((rc = make_tmap3(&tv, a->dtype, a->V, cols, a->Nk, a->B, a->ldv, a->sbv, 64, 128))) return rc;
  PArgs t;
  fill(a, &t);
  if (two && t.n_items > num_sms()) {
    const int slots = 2 * num_sms();
    const int grid = t.n_items < slots ? t.n_items : slots;
    GOAT_CUDA(launch_pdl(attn_fwd_pipe_kernel<T, 1, 2>, dim3(grid), dim3(P_THREADS), (size_t)f_smem<1>(), st, tq, tk, tv, t));
  } else {
    const int grid = t.n_items < num_sms() ? t.n_items : num_sms();
    GOAT_CUDA(launch_pdl(attn_fwd_pipe_kernel<T, 3, 1>, dim3(grid), dim3(P_THREADS), (size_t)f_smem<3>(), st, tq, tk, tv, t));
  }
  GOAT_LAUNCH_CHECK();
  return GOAT_OK;
}

template <typename T>
int bwd_launch(const goat_attn_args* a, cudaStream_t st) {
  static bool cfg = false;
  if (!cfg) {
    GOAT_CUDA(cudaFuncSetAttribute(attn_bwd_pipe_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, B_SMEM));
    cfg = true;
  }
  CUtensorMap tq, tk, tv, tg;
  int rc;
  const uint64_t cols = (uint64_t)a->heads * 64;
  if ((rc = make_tmap3(&tq, a->dtype, a->Q, cols, a->Nq, a->B, a->ldq, a->sbq, 64, 128))) return rc;
  if ((rc = make_tmap3(&tk, a->dtype, a->K, cols, a->Nk, a->B, a->ldk, a->sbk, 64, 128))) return rc;
  if ((rc = make_tmap3(&tv, a->dtype, a->V, cols, a->Nk, a->B, a->ldv, a->sbv, 64, 128))) return rc;
  if ((rc = make_tmap3(&tg, a->dtype, a->dO, cols, a->Nq, a->B, a->ldo, a->sbo, 64, 128))) return rc;
  PArgs t;
  fill(a, &t);
  const int grid = t.n_items < num_sms() ? t.n_items : num_sms();
  GOAT_CUDA(launch_pdl(attn_bwd_pipe_kernel<T>, dim3(grid), dim3(P_THREADS), (size_t)B_SMEM, st, tq, tk, tv, tg, t));
  GOAT_LAUNCH_CHECK();
  return GOAT_OK;
}

}  // namespace

// single key chunk only; longer key sequences (RxR, 512 tokens) stay on the chunked kernels of attention_tc.cu
bool attn_pipe_eligible(const goat_attn_args* a) { return a->Nk <= 128 && a->Nq <= 128; }

int attn_fwd_pipe(const goat_attn_args* a, cudaStream_t st) {
  return a->dtype == GOAT_F16 ? fwd_launch<__half>(a, st) : fwd_launch<__nv_bfloat16>(a, st);
}
int attn_bwd_pipe(const goat_attn_args* a, cudaStream_t st) {
  return a->dtype == GOAT_F16 ? bwd_launch<__half>(a, st) : bwd_launch<__nv_bfloat16>(a, st);
}

}  // namespace goat
