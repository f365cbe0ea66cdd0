// tcgen05 GEMM for sm_100a: out[M,N] = epilogue(A[M,K] * B[N,K]^T), fp16/bf16 operands, fp32 accumulation.
//
// One CTA computes one 128 x BLOCK_N output tile.  Warp roles (192 threads):
//   warp 0   : TMA producer  -- cp.async.bulk.tensor tiles of A and B into a STAGES-deep smem ring
//   warp 1   : TMEM allocator + MMA issuer -- one elected lane issues tcgen05.mma (UMMA 128 x BLOCK_N x 16),
//              accumulator lives in TMEM; tcgen05.commit releases smem stages / signals the epilogue
//   warps 2-5: epilogue -- tcgen05.ld the accumulator (each warp owns TMEM lanes 32*(warp%4)..+31),
//              bias / GELU / dGELU / dropout / residual, vectorised stores
// Operands may be K-major (contraction dim contiguous: activations and weights in forward) or
// MN-major (dY and W in dgrad, dY and X in wgrad); both use the SWIZZLE_128B canonical layouts, the
// major-ness goes into the instruction descriptor and the smem matrix descriptors.
// Ragged M / N / K edges: TMA zero-fills out-of-bounds box elements; the epilogue predicates stores.
#include <cuda.h>

#include "common.cuh"

namespace goat {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 x 2 B = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 192;

template <int BLOCK_N, int STAGES>
struct Cfg {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // +1024: manual 1 KB alignment
};

template <typename T, int BLOCK_N, int STAGES, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, EpiParams ep,
                 int M, int N, int K) {
  using C = Cfg<BLOCK_N, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BLOCK_N;
  const int m0 = blockIdx.y * BLOCK_M;
  const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<BLOCK_N>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], C::STAGE_BYTES);
        uint8_t* sa = smem + s * C::STAGE_BYTES;
        uint8_t* sb = sa + C::A_BYTES;
        const int k0 = kb * BLOCK_K;
        if (!A_MN) {
          tma_load_2d(sa, &tmA, &full_bar[s], k0, m0);  // box {64 k, 128 m}
        } else {
#pragma unroll
          for (int j = 0; j < BLOCK_M / 64; ++j)       // boxes {64 m, 64 k}
            tma_load_2d(sa + j * 8192, &tmA, &full_bar[s], m0 + 64 * j, k0);
        }
        if (!B_MN) {
          tma_load_2d(sb, &tmB, &full_bar[s], k0, n0);  // box {64 k, BLOCK_N n}
        } else {
#pragma unroll
          for (int j = 0; j < BLOCK_N / 64; ++j)
            tma_load_2d(sb + j * 8192, &tmB, &full_bar[s], n0 + 64 * j, k0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(UmmaFmt<T>::value, A_MN ? 1 : 0, B_MN ? 1 : 0, BLOCK_M, BLOCK_N);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
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
          umma_f16(tmem_base, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // frees this smem stage when the MMAs above have read it
      }
      umma_commit(tmem_full_bar);    // accumulator complete
    }
  } else {
    mbar_wait(tmem_full_bar, 0);
    tcgen05_fence_after();
    const int lg = warp & 3;  // TMEM lane group this warp may access
    const int row = m0 + lg * 32 + lane;
    const bool row_ok = row < M;
    const bool vec_ok = ((ep.ldc & 7) == 0) && (!ep.res || (ep.ldres & 3) == 0) && ((ep.ldaux & 7) == 0) &&
                        (!ep.out2 || (ep.ldc2 & 7) == 0);
#pragma unroll 1
    for (int c = 0; c < BLOCK_N / 32; ++c) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(c * 32), r);
      tmem_ld_wait();
      const int nc = n0 + c * 32;
      if (!row_ok || nc >= N) continue;
      if (vec_ok && nc + 32 <= N) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * ep.alpha;
        if (ep.bias) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + nc + j));
            v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
          }
        }
        if (ep.act == GOAT_ACT_GELU) {
          if (ep.aux_out) {
            uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<T*>(ep.aux_out) + (size_t)row * ep.ldaux + nc);
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 w;
              w.x = pack2<T>(v[j], v[j + 1]); w.y = pack2<T>(v[j + 2], v[j + 3]);
              w.z = pack2<T>(v[j + 4], v[j + 5]); w.w = pack2<T>(v[j + 6], v[j + 7]);
              dst[j >> 3] = w;
              float2 f;
              f = unpack2<T>(w.x); v[j] = f.x; v[j + 1] = f.y;
              f = unpack2<T>(w.y); v[j + 2] = f.x; v[j + 3] = f.y;
              f = unpack2<T>(w.z); v[j + 4] = f.x; v[j + 5] = f.y;
              f = unpack2<T>(w.w); v[j + 6] = f.x; v[j + 7] = f.y;
            }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
        } else if (ep.act == GOAT_ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
        } else if (ep.act == GOAT_ACT_TANH) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = tanhf(v[j]);
        } else if (ep.act == GOAT_ACT_DGELU || ep.act == GOAT_ACT_DRELU) {
          const uint4* src =
              reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(ep.aux_in) + (size_t)row * ep.ldaux + nc);
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            const uint4 w = __ldg(src + (j >> 3));
            float z[8];
            float2 f;
            f = unpack2<T>(w.x); z[0] = f.x; z[1] = f.y;
            f = unpack2<T>(w.y); z[2] = f.x; z[3] = f.y;
            f = unpack2<T>(w.z); z[4] = f.x; z[5] = f.y;
            f = unpack2<T>(w.w); z[6] = f.x; z[7] = f.y;
#pragma unroll
            for (int t = 0; t < 8; ++t)
              v[j + t] = (ep.act == GOAT_ACT_DGELU) ? v[j + t] * dgelu_erf(z[t]) : (z[t] > 0.0f ? v[j + t] : 0.0f);
          }
        }
        if (ep.drop_p > 0.0f) {
          const float keep = 1.0f / (1.0f - ep.drop_p);
          const unsigned long long seed = eff_seed(ep.drop_seed, ep.drop_seed_ptr);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float u = rand_uniform(seed, (unsigned long long)row * (unsigned long long)ep.ldc + nc + j);
            v[j] = (u >= ep.drop_p) ? v[j] * keep : 0.0f;
          }
        }
        if (ep.res) {
          const float4* src = reinterpret_cast<const float4*>(ep.res + (size_t)row * ep.ldres + nc);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b = src[j >> 2];
            v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
          }
        }
        if (ep.out_f32) {
          float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + (size_t)row * ep.ldc + nc);
#pragma unroll
          for (int j = 0; j < 32; j += 4) dst[j >> 2] = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
          uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<T*>(ep.out) + (size_t)row * ep.ldc + nc);
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 w;
            w.x = pack2<T>(v[j], v[j + 1]); w.y = pack2<T>(v[j + 2], v[j + 3]);
            w.z = pack2<T>(v[j + 4], v[j + 5]); w.w = pack2<T>(v[j + 6], v[j + 7]);
            dst[j >> 3] = w;
          }
        }
        if (ep.out2) {
          uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<T*>(ep.out2) + (size_t)row * ep.ldc2 + nc);
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 w;
            w.x = pack2<T>(v[j], v[j + 1]); w.y = pack2<T>(v[j + 2], v[j + 3]);
            w.z = pack2<T>(v[j + 4], v[j + 5]); w.w = pack2<T>(v[j + 6], v[j + 7]);
            dst[j >> 3] = w;
          }
        }
      } else {
        // ragged / unaligned edge: scalar path through the shared epilogue
        EpiParams e1 = ep;
        e1.alpha = ep.alpha;
#pragma unroll 1
        for (int j = 0; j < 32; ++j) {
          const int n = nc + j;
          if (n >= N) break;
          const float val = epi_apply<T>(e1, row, n, __uint_as_float(r[j]));
          if (ep.out_f32) reinterpret_cast<float*>(ep.out)[(size_t)row * ep.ldc + n] = val;
          else reinterpret_cast<T*>(ep.out)[(size_t)row * ep.ldc + n] = from_f<T>(val);
          if (ep.out2) reinterpret_cast<T*>(ep.out2)[(size_t)row * ep.ldc2 + n] = from_f<T>(val);
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<BLOCK_N>(tmem_base);
}

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

}  // namespace

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
  dim3 grid((a.N + BLOCK_N - 1) / BLOCK_N, (a.M + BLOCK_M - 1) / BLOCK_M);
  kern<<<grid, GEMM_THREADS, C::SMEM_BYTES, stream>>>(tmA, tmB, ep, a.M, a.N, a.K);
  GOAT_LAUNCH_CHECK();
  return GOAT_OK;
}

template <typename T>
int dispatch(const goat_gemm_args& a, const EpiParams& ep, cudaStream_t stream) {
  if (!a.a_mn_major && !a.b_mn_major) return launch<T, 128, 6, false, false>(a, ep, stream);
  if (!a.a_mn_major && a.b_mn_major) return launch<T, 128, 6, false, true>(a, ep, stream);
  if (a.a_mn_major && a.b_mn_major) return launch<T, 128, 6, true, true>(a, ep, stream);
  return launch<T, 128, 6, true, false>(a, ep, stream);
}

}  // namespace

bool gemm_umma_eligible(const goat_gemm_args& a) {
  if (a.dtype != GOAT_F16 && a.dtype != GOAT_BF16) return false;
  if (a.K < 16 || (a.K & 7) || (a.lda & 7) || (a.ldb & 7)) return false;
  if (!aligned16(a.A) || !aligned16(a.B)) return false;
  return true;
}

int gemm_umma(const goat_gemm_args& a, const EpiParams& ep, cudaStream_t stream) {
  if (a.dtype == GOAT_F16) return dispatch<__half>(a, ep, stream);
  return dispatch<__nv_bfloat16>(a, ep, stream);
}

}  // namespace goat
