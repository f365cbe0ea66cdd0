// tcgen05 GEMM for sm_100a: out[M,N] = epilogue(A[M,K] * B[N,K]^T), fp16/bf16 operands, fp32 accumulation.
//
// Persistent, warp-specialised: each CTA (one per SM) walks a static list of 128 x 128 output tiles
// (tile = blockIdx.x + i * gridDim.x; N fastest so concurrently running CTAs share A panels in L2).
//   warp 0    : TMA producer  -- cp.async.bulk.tensor tiles of A and B into a STAGES-deep smem ring
//   warp 1    : TMEM allocator + MMA issuer -- one elected lane issues tcgen05.mma (UMMA 128 x 128 x 16) into one
//               of TWO TMEM accumulators, so the epilogue of tile i overlaps the main loop of tile i+1
//   warps 2-9 : epilogue -- warp w owns TMEM lanes 32*(w%4).. and one column half; tcgen05.ld 32 columns, transpose
//               through a private padded smem patch so that every global access (bias / residual / GELU aux /
//               output) is a coalesced 16-byte access along the row, then bias / GELU / dGELU / dropout / residual
// Split-K (accumulate mode, wgrad): the K range is cut into `splits` tile-sized pieces that are scheduled as
// independent tiles and reduced with vector fp32 atomics (red.global.add.v4.f32) into a zero-initialised output --
// a 768 x 768 x 5120 weight-gradient GEMM has only 36 output tiles for 148 SMs otherwise.
// Operands may be K-major (contraction dim contiguous: activations and weights in forward) or
// MN-major (dY and W in dgrad, dY and X in wgrad); both use the SWIZZLE_128B canonical layouts, the
// major-ness goes into the instruction descriptor and the smem matrix descriptors.
// Ragged M / N / K edges: TMA zero-fills out-of-bounds box elements; the epilogue predicates stores.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace goat {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 x 2 B = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int EPI_WARPS = 8;
constexpr int GEMM_THREADS = 64 + EPI_WARPS * 32;
constexpr int PATCH_LD = 36;                       // floats per staged row (32 + 4 padding: conflict-free both ways)
constexpr int PATCH_BYTES = 32 * PATCH_LD * 4;     // one warp's 32 x 32 transpose patch

template <int BLOCK_N, int STAGES>
struct Cfg {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_BYTES = EPI_WARPS * PATCH_BYTES;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024;  // +1024: manual alignment
};

struct Sched {
  int tiles_m, tiles_n, splits, kb_per_split, num_kb, num_tiles;
  int kb_wrap;   // > 0: k-blocks [kb_wrap, num_kb) re-read A from k-block (kb - kb_wrap) against the second B operand (B_lo)
};

__device__ __forceinline__ void tile_coords(const Sched& sc, int t, int& m0, int& n0, int& kb0, int& kb1) {
  const int tn = t % sc.tiles_n;
  const int r = t / sc.tiles_n;
  const int tm = r % sc.tiles_m;
  const int ks = r / sc.tiles_m;
  m0 = tm * BLOCK_M;
  n0 = tn * 128;
  kb0 = ks * sc.kb_per_split;
  kb1 = min(sc.num_kb, kb0 + sc.kb_per_split);
}

// 4 consecutive outputs (row m, columns n..n+3) through the fused epilogue, vector global accesses
template <typename T>
__device__ __forceinline__ void epi_store4(const EpiParams& ep, int m, int n, float4 acc, unsigned long long seed) {
  float v[4] = {acc.x * ep.alpha, acc.y * ep.alpha, acc.z * ep.alpha, acc.w * ep.alpha};
  if (ep.accumulate) {
    atomicAdd(reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + (size_t)m * ep.ldc + n),
              make_float4(v[0], v[1], v[2], v[3]));
    return;
  }
  if (ep.bias) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + n));
    v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
  }
  if (ep.act == GOAT_ACT_GELU) {
    if (ep.aux_out) {
      uint2 w;
      w.x = pack2<T>(v[0], v[1]);
      w.y = pack2<T>(v[2], v[3]);
      *reinterpret_cast<uint2*>(reinterpret_cast<T*>(ep.aux_out) + (size_t)m * ep.ldaux + n) = w;
      // GELU of the ROUNDED pre-activation: backward only ever sees the stored 16-bit z
      float2 f;
      f = unpack2<T>(w.x); v[0] = f.x; v[1] = f.y;
      f = unpack2<T>(w.y); v[2] = f.x; v[3] = f.y;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = gelu_fast(v[j]);
  } else if (ep.act == GOAT_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.0f);
  } else if (ep.act == GOAT_ACT_TANH) {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = tanhf(v[j]);
  } else if (ep.act == GOAT_ACT_DGELU || ep.act == GOAT_ACT_DRELU) {
    const uint2 w = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const T*>(ep.aux_in) + (size_t)m * ep.ldaux + n));
    const float2 a = unpack2<T>(w.x), b = unpack2<T>(w.y);
    const float z[4] = {a.x, a.y, b.x, b.y};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      v[j] = (ep.act == GOAT_ACT_DGELU) ? v[j] * dgelu_fast(z[j]) : (z[j] > 0.0f ? v[j] : 0.0f);
  }
  if (ep.drop_p > 0.0f)   // vector path: ldc % 4 == 0 and n % 4 == 0, so the element index is a multiple of 4
    drop_apply4(v, seed, (unsigned long long)(m + ep.drop_row0) * (unsigned long long)ep.ldc + n, drop_thr16(ep.drop_p), 1.0f / (1.0f - ep.drop_p));
  if (ep.res) {
    const float4 b = *reinterpret_cast<const float4*>(ep.res + (size_t)m * ep.ldres + n);
    v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
  }
  if (ep.out_f32) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + (size_t)m * ep.ldc + n) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    uint2 w;
    w.x = pack2<T>(v[0], v[1]);
    w.y = pack2<T>(v[2], v[3]);
    *reinterpret_cast<uint2*>(reinterpret_cast<T*>(ep.out) + (size_t)m * ep.ldc + n) = w;
  }
  if (ep.out2) {
    uint2 w;
    w.x = pack2<T>(v[0], v[1]);
    w.y = pack2<T>(v[2], v[3]);
    *reinterpret_cast<uint2*>(reinterpret_cast<T*>(ep.out2) + (size_t)m * ep.ldc2 + n) = w;
  }
}

template <typename T>
__device__ __forceinline__ void epi_store1(const EpiParams& ep, int m, int n, float acc) {
  if (ep.accumulate) {
    atomicAdd(reinterpret_cast<float*>(ep.out) + (size_t)m * ep.ldc + n, acc * ep.alpha);
    return;
  }
  const float val = epi_apply<T>(ep, m, n, acc);
  if (ep.out_f32) reinterpret_cast<float*>(ep.out)[(size_t)m * ep.ldc + n] = val;
  else reinterpret_cast<T*>(ep.out)[(size_t)m * ep.ldc + n] = from_f<T>(val);
  if (ep.out2) reinterpret_cast<T*>(ep.out2)[(size_t)m * ep.ldc2 + n] = from_f<T>(val);
}

template <typename T, int BLOCK_N, int STAGES, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmB2, EpiParams ep,
                 int M, int N, int K, Sched sc) {
  static_assert(BLOCK_N == 128, "tile scheduler and TMEM double buffering assume 128-wide tiles");
  using C = Cfg<BLOCK_N, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  float* patches = reinterpret_cast<float*>(smem + STAGES * C::STAGE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * C::STAGE_BYTES + C::EPI_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<2 * BLOCK_N>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                 // everything above overlapped the previous kernel's tail
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;  // running k-block counter across tiles: stage = it % STAGES
      for (int t = blockIdx.x; t < sc.num_tiles; t += gridDim.x) {
        int m0, n0, kb0, kb1;
        tile_coords(sc, t, m0, n0, kb0, kb1);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_arrive_expect_tx(&full_bar[s], C::STAGE_BYTES);
          uint8_t* sa = smem + s * C::STAGE_BYTES;
          uint8_t* sb = sa + C::A_BYTES;
          const bool lo = sc.kb_wrap > 0 && kb >= sc.kb_wrap;
          const int k0 = (lo ? kb - sc.kb_wrap : kb) * BLOCK_K;
          const CUtensorMap* mapB = lo ? &tmB2 : &tmB;
          if (!A_MN) {
            tma_load_2d(sa, &tmA, &full_bar[s], k0, m0);  // box {64 k, 128 m}
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_M / 64; ++j)       // boxes {64 m, 64 k}
              tma_load_2d(sa + j * 8192, &tmA, &full_bar[s], m0 + 64 * j, k0);
          }
          if (!B_MN) {
            tma_load_2d(sb, mapB, &full_bar[s], k0, n0);  // box {64 k, BLOCK_N n}
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_N / 64; ++j)
              tma_load_2d(sb + j * 8192, mapB, &full_bar[s], n0 + 64 * j, k0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(UmmaFmt<T>::value, A_MN ? 1 : 0, B_MN ? 1 : 0, BLOCK_M, BLOCK_N);
      uint32_t it = 0;
      int local = 0;
      for (int t = blockIdx.x; t < sc.num_tiles; t += gridDim.x, ++local) {
        int m0, n0, kb0, kb1;
        tile_coords(sc, t, m0, n0, kb0, kb1);
        const int buf = local & 1;
        mbar_wait(&tmem_empty_bar[buf], (((uint32_t)local >> 1) & 1) ^ 1);   // epilogue drained this accumulator
        tcgen05_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(buf * BLOCK_N);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + s * C::STAGE_BYTES);
          const uint32_t sb = sa + C::A_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // K-major : 8-row groups 1024 B apart (SBO); advance 16 elements = 32 B inside the swizzle row.
            // MN-major: 64-element MN groups 8192 B apart (LBO), 8-k-row groups 1024 B apart (SBO);
            //           advance 16 k-rows = 2048 B.
            const uint64_t da = A_MN ? make_smem_desc_sw128(sa + k * 2048, 8192, 1024)
                                     : make_smem_desc_sw128(sa + k * 32, 0, 1024);
            const uint64_t db = B_MN ? make_smem_desc_sw128(sb + k * 2048, 8192, 1024)
                                     : make_smem_desc_sw128(sb + k * 32, 0, 1024);
            umma_f16(tacc, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);  // frees this smem stage when the MMAs above have read it
        }
        umma_commit(&tmem_full_bar[buf]);  // accumulator complete
      }
    }
  } else {
    const int ew = warp - 2;
    const int lg = warp & 3;          // TMEM lane group this warp may access
    const int chalf = ew >> 2;        // which 64-column half of the tile
    float* patch = patches + ew * (PATCH_BYTES / 4);
    const unsigned long long seed = ep.drop_p > 0.0f ? eff_seed(ep.drop_seed, ep.drop_seed_ptr) : 0ull;
    const bool vec_ok = ((ep.ldc & 3) == 0) && (!ep.res || (ep.ldres & 3) == 0) &&
                        ((!ep.aux_in && !ep.aux_out) || (ep.ldaux & 3) == 0) && (!ep.out2 || (ep.ldc2 & 3) == 0);
    const int rl = lane >> 3;         // row within a group of 4
    const int c4 = (lane & 7) * 4;    // column offset of this lane's float4
    int local = 0;
    for (int t = blockIdx.x; t < sc.num_tiles; t += gridDim.x, ++local) {
      int m0, n0, kb0, kb1;
      tile_coords(sc, t, m0, n0, kb0, kb1);
      const int buf = local & 1;
      mbar_wait(&tmem_full_bar[buf], ((uint32_t)local >> 1) & 1);
      tcgen05_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t)(buf * BLOCK_N) + ((uint32_t)(lg * 32) << 16);
      const bool has_k = kb1 > kb0;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        const int cbase = chalf * 64 + c * 32;
        uint32_t r[32];
        tmem_ld_32x32b_x32(tacc + (uint32_t)cbase, r);
        tmem_ld_wait();
        if (c == 1) {
          // every TMEM read of this warp for this tile is done: hand the accumulator back to the MMA warp
          tcgen05_fence_before();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
        }
        const int nc = n0 + cbase;
        if (nc >= N || m0 + lg * 32 >= M || !has_k) continue;   // warp-uniform
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(patch + lane * PATCH_LD + j) =
              make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        __syncwarp();
#pragma unroll 2
        for (int i = 0; i < 8; ++i) {
          const int row_l = i * 4 + rl;
          const int m = m0 + lg * 32 + row_l;
          const int n = nc + c4;
          if (m < M && n < N) {
            const float4 acc = *reinterpret_cast<const float4*>(patch + row_l * PATCH_LD + c4);
            if (vec_ok && n + 4 <= N) {
              epi_store4<T>(ep, m, n, acc, seed);
            } else {
              const float a4[4] = {acc.x, acc.y, acc.z, acc.w};
              for (int j = 0; j < 4 && n + j < N; ++j) epi_store1<T>(ep, m, n + j, a4[j]);
            }
          }
        }
        __syncwarp();
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<2 * BLOCK_N>(tmem_base);
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D row-major tensor [outer, inner] (inner contiguous), 16-bit elements, 128B-swizzled boxes
int make_tmap(CUtensorMap* tm, int dtype, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
              uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn enc = get_encode_tiled();
  GOAT_CHECK(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  const cuuint64_t gdim[2] = {inner, outer};
  const cuuint64_t gstr[1] = {ld_elems * 2};
  const cuuint32_t box[2] = {box_inner, box_outer};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = dtype == GOAT_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUresult r = enc(tm, dt, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GOAT_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (inner %llu outer %llu ld %llu)", (int)r,
             (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld_elems);
  return GOAT_OK;
}

// same 2-D map with the 64-byte swizzle: 32-column x 16-bit output tiles of the TMA-store epilogue (gemm_umma2.cu)
int make_tmap_sw64(CUtensorMap* tm, int dtype, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                   uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn enc = get_encode_tiled();
  GOAT_CHECK(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  const cuuint64_t gdim[2] = {inner, outer};
  const cuuint64_t gstr[1] = {ld_elems * 2};
  const cuuint32_t box[2] = {box_inner, box_outer};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = dtype == GOAT_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUresult r = enc(tm, dt, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GOAT_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(store) failed with CUresult %d (inner %llu outer %llu ld %llu)", (int)r,
             (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld_elems);
  return GOAT_OK;
}

// 3-D map over a token-major [d2 = batch][d1 = token][d0 = channel] tensor (channel contiguous, token stride ld,
// batch stride sb, in elements); boxes of box0 channels x box1 tokens x 1 batch, 128B swizzle, zero OOB fill.
int make_tmap3(CUtensorMap* tm, int dtype, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t ld_elems,
               uint64_t sb_elems, uint32_t box0, uint32_t box1) {
  EncodeTiledFn enc = get_encode_tiled();
  GOAT_CHECK(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  const cuuint64_t gdim[3] = {d0, d1, d2};
  const cuuint64_t gstr[2] = {ld_elems * 2, (d2 > 1 ? sb_elems : (uint64_t)d1 * ld_elems) * 2};
  const cuuint32_t box[3] = {box0, box1, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapDataType dt = dtype == GOAT_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUresult r = enc(tm, dt, 3, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GOAT_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(3d) failed with CUresult %d (dims %llu %llu %llu ld %llu sb %llu)",
             (int)r, (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2,
             (unsigned long long)ld_elems, (unsigned long long)sb_elems);
  return GOAT_OK;
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

int gemm_umma2(const goat_gemm_args& a, const EpiParams& ep, int block_n, cudaStream_t stream);

namespace {

template <typename T, int BLOCK_N, int STAGES, bool A_MN, bool B_MN>
int launch(const goat_gemm_args& a, const EpiParams& ep, cudaStream_t stream) {
  using C = Cfg<BLOCK_N, STAGES>;
  auto kern = gemm_umma_kernel<T, BLOCK_N, STAGES, A_MN, B_MN>;
  static bool configured = false;
  if (!configured) {
    GOAT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    configured = true;
  }
  CUtensorMap tmA, tmB;
  int rc;
  if (!A_MN) rc = make_tmap(&tmA, a.dtype, a.A, a.K, a.M, a.lda, BLOCK_K, BLOCK_M);
  else rc = make_tmap(&tmA, a.dtype, a.A, a.M, a.K, a.lda, 64, BLOCK_K);
  if (rc) return rc;
  if (!B_MN) rc = make_tmap(&tmB, a.dtype, a.B, a.K, a.N, a.ldb, BLOCK_K, BLOCK_N);
  else rc = make_tmap(&tmB, a.dtype, a.B, a.N, a.K, a.ldb, 64, BLOCK_K);
  if (rc) return rc;
  CUtensorMap tmB2 = tmB;
  const bool use_lo = a.B_lo != nullptr && !B_MN && !ep.accumulate;
  if (use_lo && (rc = make_tmap(&tmB2, a.dtype, a.B_lo, a.K, a.N, a.ldb, BLOCK_K, BLOCK_N))) return rc;
  Sched sc;
  sc.tiles_m = (a.M + BLOCK_M - 1) / BLOCK_M;
  sc.tiles_n = (a.N + BLOCK_N - 1) / BLOCK_N;
  sc.num_kb = (a.K + BLOCK_K - 1) / BLOCK_K;
  sc.splits = 1;
  const int mn = sc.tiles_m * sc.tiles_n;
  if (ep.accumulate && mn < num_sms()) {
    // split K so that about one wave of tiles exists, keeping at least 4 k-blocks (256 of K) per split
    int want = (num_sms() + mn - 1) / mn;
    int cap = sc.num_kb / 4;
    if (cap < 1) cap = 1;
    sc.splits = want < cap ? want : cap;
  }
  sc.kb_wrap = 0;
  if (use_lo) {            // second pass of the K loop over A against B_lo (never combined with split-K)
    sc.kb_wrap = sc.num_kb;
    sc.num_kb *= 2;
  }
  sc.kb_per_split = (sc.num_kb + sc.splits - 1) / sc.splits;
  sc.splits = (sc.num_kb + sc.kb_per_split - 1) / sc.kb_per_split;   // no empty splits
  sc.num_tiles = mn * sc.splits;
  const int grid = sc.num_tiles < num_sms() ? sc.num_tiles : num_sms();
  GOAT_CUDA(launch_pdl(kern, dim3(grid), dim3(GEMM_THREADS), C::SMEM_BYTES, stream, tmA, tmB, tmB2, ep, a.M, a.N, a.K, sc));
  return GOAT_OK;
}

template <typename T>
int dispatch(const goat_gemm_args& a, const EpiParams& ep, cudaStream_t stream) {
  if (!a.a_mn_major && !a.b_mn_major) return launch<T, 128, 5, false, false>(a, ep, stream);
  if (!a.a_mn_major && a.b_mn_major) return launch<T, 128, 5, false, true>(a, ep, stream);
  if (a.a_mn_major && a.b_mn_major) return launch<T, 128, 5, true, true>(a, ep, stream);
  return launch<T, 128, 5, true, false>(a, ep, stream);
}

}  // namespace

bool gemm_umma_eligible(const goat_gemm_args& a) {
  if (a.dtype != GOAT_F16 && a.dtype != GOAT_BF16) return false;
  // TMA needs 16-byte global strides (leading dimensions multiples of 8 elements); the extents themselves may be ragged
  // (out-of-range elements are zero-filled), so K need not be a multiple of 8 once the rows are padded -- the
  // 50265-wide vocabulary gradient is read that way
  if (a.K < 16 || (a.lda & 7) || (a.ldb & 7)) return false;
  if (!aligned16(a.A) || !aligned16(a.B)) return false;
  return true;
}

// Tile selection.  M <= 128 (heads, poolers): one CTA per 128 x 128 tile.  Otherwise a CTA pair per 256 x BLOCK_N tile
// (gemm_umma2.cu); 256-wide tiles halve the L2 bytes per FLOP and are taken when enough of them exist to occupy most
// of the 74 pairs.  GOAT_GEMM_2CTA=0 / GOAT_GEMM_BN=128|256 override (tuning and A/B measurements).
static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}

int gemm_umma(const goat_gemm_args& a, const EpiParams& ep, cudaStream_t stream) {
  static const int use2 = env_int("GOAT_GEMM_2CTA", 1);
  static const int force_bn = env_int("GOAT_GEMM_BN", 0);
  if (use2 && a.M > 128) {
    int bn = force_bn;
    if (bn != 128 && bn != 256) {
      const int tiles256 = ((a.M + 255) / 256) * ((a.N + 255) / 256);
      bn = (a.N >= 256 && (ep.accumulate || tiles256 >= 40)) ? 256 : 128;
    }
    return gemm_umma2(a, ep, bn, stream);
  }
  if (a.dtype == GOAT_F16) return dispatch<__half>(a, ep, stream);
  return dispatch<__nv_bfloat16>(a, ep, stream);
}

}  // namespace goat
