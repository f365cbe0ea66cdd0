// Generic SIMT GEMM (fp32 FFMA accumulate) with the same fused epilogue as the tcgen05 kernel.
// Product path for GOAT_F32 (1e-5 parity mode) and for shapes the TMA/UMMA path cannot take
// (K or leading dims not multiples of 8: the 7/14-wide position features, the 1-wide heads).
// Also the on-device cross-check for the tcgen05 kernel (force_simt).
#include "common.cuh"

namespace goat {

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

template <typename T>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const T* __restrict__ A, const T* __restrict__ B, EpiParams ep, int M, int N, int Kfull, long long sAm,
                 long long sAk, long long sBn, long long sBk, int k_chunk) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, 4 x 4 outputs each
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  // loader mapping: pick the thread->element order that makes global reads contiguous
  const bool a_kfast = (sAk == 1), b_kfast = (sBk == 1);
  // split-K (accumulate mode only): grid.z slices of k_chunk; every slice adds its partial product with atomics
  const int kbeg = blockIdx.z * k_chunk;
  const int K = min(Kfull, kbeg + k_chunk);
  for (int k0 = kbeg; k0 < K; k0 += TK) {
#pragma unroll
    for (int t = 0; t < (TM * TK) / 256; ++t) {
      const int e = tid + t * 256;
      const int kk = a_kfast ? (e % TK) : (e / TM);
      const int mm = a_kfast ? (e / TK) : (e % TM);
      const int m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < M && k < K) ? to_f<T>(A[m * sAm + k * sAk]) : 0.0f;
    }
#pragma unroll
    for (int t = 0; t < (TN * TK) / 256; ++t) {
      const int e = tid + t * 256;
      const int kk = b_kfast ? (e % TK) : (e / TN);
      const int nn = b_kfast ? (e / TK) : (e % TN);
      const int n = n0 + nn, k = k0 + kk;
      Bs[kk][nn] = (n < N && k < K) ? to_f<T>(B[n * sBn + k * sBk]) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      if (ep.accumulate) {
        atomicAdd(reinterpret_cast<float*>(ep.out) + (size_t)m * ep.ldc + n, acc[i][j] * ep.alpha);
        continue;
      }
      const float v = epi_apply<T>(ep, m, n, acc[i][j]);
      if (ep.out_f32) reinterpret_cast<float*>(ep.out)[(size_t)m * ep.ldc + n] = v;
      else reinterpret_cast<T*>(ep.out)[(size_t)m * ep.ldc + n] = from_f<T>(v);
      if (ep.out2) reinterpret_cast<T*>(ep.out2)[(size_t)m * ep.ldc2 + n] = from_f<T>(v);
    }
  }
}

template <typename T>
int launch_simt(const goat_gemm_args& a, const EpiParams& ep, cudaStream_t stream) {
  const long long sAm = a.a_mn_major ? 1 : a.lda, sAk = a.a_mn_major ? a.lda : 1;
  const long long sBn = a.b_mn_major ? 1 : a.ldb, sBk = a.b_mn_major ? a.ldb : 1;
  dim3 grid((a.N + TN - 1) / TN, (a.M + TM - 1) / TM);
  int k_chunk = a.K;
  if (ep.accumulate) {
    // weight gradients of the narrow layers (7 / 14-wide position features, 1-wide heads) are [few tiles] x [K = all tokens]:
    // split K so that about four waves of CTAs exist instead of a dozen CTAs walking thousands of k-steps each
    const long long tiles = (long long)grid.x * grid.y;
    long long want = (4 * 148 + tiles - 1) / tiles;
    const long long cap = (a.K + 4 * TK - 1) / (4 * TK);     // at least 64 of K per slice
    if (want > cap) want = cap;
    if (want > 65535) want = 65535;
    if (want > 1) {
      k_chunk = (int)(((a.K + want - 1) / want + TK - 1) / TK * TK);
      grid.z = (unsigned)((a.K + k_chunk - 1) / k_chunk);
    }
  }
  gemm_simt_kernel<T><<<grid, 256, 0, stream>>>(reinterpret_cast<const T*>(a.A), reinterpret_cast<const T*>(a.B), ep,
                                                 a.M, a.N, a.K, sAm, sAk, sBn, sBk, k_chunk);
  GOAT_LAUNCH_CHECK();
  return GOAT_OK;
}

}  // namespace

int gemm_simt(const goat_gemm_args& a, const EpiParams& ep, cudaStream_t stream) {
  if (a.dtype == GOAT_F32) return launch_simt<float>(a, ep, stream);
  if (a.dtype == GOAT_F16) return launch_simt<__half>(a, ep, stream);
  return launch_simt<__nv_bfloat16>(a, ep, stream);
}

}  // namespace goat
