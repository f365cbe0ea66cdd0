// LayerNorm forward / backward (fp32 statistics), column sums and dtype casts.
// Memory-bound helpers: one warp per row, 16-byte vector accesses, two-stage deterministic column reductions.
#include <stdlib.h>

#include "common.cuh"

namespace goat {

namespace {

constexpr int LN_ROWS_PER_CTA = 4;   // 4 warps, one row each per iteration
constexpr int LN_MAX_PER_LANE = 32;  // H <= 1024

template <typename T>
__device__ __forceinline__ float ldx(const void* p, size_t i) { return to_f<T>(reinterpret_cast<const T*>(p)[i]); }

// ---------------------------------------------------------------------------------------------
template <typename TX, typename TY>
__global__ void __launch_bounds__(128)
ln_fwd_kernel(const void* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
              float* __restrict__ y32, void* __restrict__ y16, float* __restrict__ mean, float* __restrict__ rstd, int M,
              int H) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * LN_ROWS_PER_CTA + warp;
  if (row >= M) return;
  float v[LN_MAX_PER_LANE];
  const int n = (H + 31) / 32;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
    if (i < n) {
      const int c = i * 32 + lane;
      v[i] = c < H ? ldx<TX>(x, (size_t)row * H + c) : 0.f;
      s += v[i];
    }
  }
  const float mu = warp_sum(s) / H;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
    if (i < n) {
      const int c = i * 32 + lane;
      const float d = c < H ? v[i] - mu : 0.f;
      q += d * d;
    }
  }
  const float rs = rsqrtf(warp_sum(q) / H + eps);
  if (lane == 0) {
    if (mean) mean[row] = mu;
    if (rstd) rstd[row] = rs;
  }
#pragma unroll
  for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
    if (i < n) {
      const int c = i * 32 + lane;
      if (c < H) {
        const float y = (v[i] - mu) * rs * gamma[c] + beta[c];
        if (y32) y32[(size_t)row * H + c] = y;
        if (y16) reinterpret_cast<TY*>(y16)[(size_t)row * H + c] = from_f<TY>(y);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward: per row   xhat = (x-mu)*rstd, g = dy*gamma,
//   dx = rstd * (g - mean(g) - xhat * mean(g*xhat))
// column partials (dgamma = sum dy*xhat, dbeta = sum dy, dcolsum = sum dx16) per CTA into workspace
// [gridDim.x][3][H]; a second kernel reduces them in a fixed order (deterministic).
template <typename TX, typename TD>
__global__ void __launch_bounds__(128)
ln_bwd_kernel(const float* __restrict__ dy, const void* __restrict__ x, const float* __restrict__ gamma,
              const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ dres,
              float* __restrict__ dx32, void* __restrict__ dx16, float drop_p, unsigned long long drop_seed_,
              const unsigned long long* __restrict__ drop_seed_ptr, float* __restrict__ partial, int M, int H,
              int rows_per_cta) {
  extern __shared__ float sacc[];  // [4 warps][3][H]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = (H + 31) / 32;
  float ag[LN_MAX_PER_LANE], ab[LN_MAX_PER_LANE], ac[LN_MAX_PER_LANE];
#pragma unroll
  for (int i = 0; i < LN_MAX_PER_LANE; ++i) { ag[i] = 0.f; ab[i] = 0.f; ac[i] = 0.f; }
  const int r0 = blockIdx.x * rows_per_cta;
  const int r1 = min(M, r0 + rows_per_cta);
  const float keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const unsigned long long drop_seed = drop_p > 0.f ? eff_seed(drop_seed_, drop_seed_ptr) : 0ull;
  for (int row = r0 + warp; row < r1; row += 4) {
    const float mu = mean[row], rs = rstd[row];
    float xh[LN_MAX_PER_LANE], g[LN_MAX_PER_LANE];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
      if (i < n) {
        const int c = i * 32 + lane;
        if (c < H) {
          const float d = dy[(size_t)row * H + c];
          xh[i] = (ldx<TX>(x, (size_t)row * H + c) - mu) * rs;
          g[i] = d * gamma[c];
          ag[i] += d * xh[i];
          ab[i] += d;
          s1 += g[i];
          s2 += g[i] * xh[i];
        } else { xh[i] = 0.f; g[i] = 0.f; }
      }
    }
    s1 = warp_sum(s1) / H;
    s2 = warp_sum(s2) / H;
#pragma unroll
    for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
      if (i < n) {
        const int c = i * 32 + lane;
        if (c < H) {
          const float d = rs * (g[i] - s1 - xh[i] * s2);
          if (dx32) dx32[(size_t)row * H + c] = d + (dres ? dres[(size_t)row * H + c] : 0.f);
          float dm = d;
          if (drop_p > 0.f)
            dm = drop_keep(drop_seed, (unsigned long long)row * (unsigned long long)H + c, drop_thr16(drop_p)) ? d * keep : 0.f;
          if (dx16) {
            const TD q = from_f<TD>(dm);
            reinterpret_cast<TD*>(dx16)[(size_t)row * H + c] = q;
            dm = to_f<TD>(q);
          }
          ac[i] += dm;
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
    if (i < n) {
      const int c = i * 32 + lane;
      if (c < H) {
        sacc[(warp * 3 + 0) * H + c] = ag[i];
        sacc[(warp * 3 + 1) * H + c] = ab[i];
        sacc[(warp * 3 + 2) * H + c] = ac[i];
      }
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 3 * H; e += blockDim.x) {
    const int k = e / H, c = e % H;
    partial[((size_t)blockIdx.x * 3 + k) * H + c] =
        sacc[(0 * 3 + k) * H + c] + sacc[(1 * 3 + k) * H + c] + sacc[(2 * 3 + k) * H + c] + sacc[(3 * 3 + k) * H + c];
  }
}

__global__ void ln_bwd_finalize_kernel(const float* __restrict__ partial, int nparts, int H, float* dgamma, float* dbeta,
                                       float* dcolsum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= H) return;
  float a = 0.f, b = 0.f, d = 0.f;
  for (int p = 0; p < nparts; ++p) {
    a += partial[((size_t)p * 3 + 0) * H + c];
    b += partial[((size_t)p * 3 + 1) * H + c];
    d += partial[((size_t)p * 3 + 2) * H + c];
  }
  if (dgamma) dgamma[c] = a;
  if (dbeta) dbeta[c] = b;
  if (dcolsum) dcolsum[c] = d;
}


// ---------------------------------------------------------------------------------------------
// Vectorised variants for H = 128 * NV (768 -> NV = 6): one warp per row, lane owns float4 columns
// c = i*128 + lane*4, i < NV.  All global accesses are 16-byte (8-byte for 16-bit outputs), the row stays in
// registers between the statistics and the normalisation pass.
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float4 ld4(const void* p, size_t i) {
  if constexpr (sizeof(T) == 4) {
    return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i);
  } else {
    const uint2 w = *reinterpret_cast<const uint2*>(reinterpret_cast<const T*>(p) + i);
    const float2 a = unpack2<T>(w.x), b = unpack2<T>(w.y);
    return make_float4(a.x, a.y, b.x, b.y);
  }
}
template <typename T>
__device__ __forceinline__ void st4(void* p, size_t i, float4 v) {
  if constexpr (sizeof(T) == 4) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + i) = v;
  } else {
    uint2 w;
    w.x = pack2<T>(v.x, v.y);
    w.y = pack2<T>(v.z, v.w);
    *reinterpret_cast<uint2*>(reinterpret_cast<T*>(p) + i) = w;
  }
}

template <typename TX, typename TY, int NV>
__global__ void __launch_bounds__(256)
ln_fwd_vec_kernel(const void* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                  float* __restrict__ y32, void* __restrict__ y16, float* __restrict__ mean, float* __restrict__ rstd,
                  int M) {
  constexpr int H = NV * 128;
  pdl_wait();
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= M) return;
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = ld4<TX>(x, (size_t)row * H + i * 128 + lane * 4);
    s += v[i].x + v[i].y + v[i].z + v[i].w;
  }
  const float mu = warp_sum(s) * (1.f / H);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mu, b = v[i].y - mu, c = v[i].z - mu, d = v[i].w - mu;
    q += a * a + b * b + c * c + d * d;
  }
  const float rs = rsqrtf(warp_sum(q) * (1.f / H) + eps);
  if (lane == 0) {
    if (mean) mean[row] = mu;
    if (rstd) rstd[row] = rs;
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = i * 128 + lane * 4;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
    float4 y;
    y.x = (v[i].x - mu) * rs * g.x + b.x;
    y.y = (v[i].y - mu) * rs * g.y + b.y;
    y.z = (v[i].z - mu) * rs * g.z + b.z;
    y.w = (v[i].w - mu) * rs * g.w + b.w;
    if (y32) st4<float>(y32, (size_t)row * H + c, y);
    if (y16) st4<TY>(y16, (size_t)row * H + c, y);
  }
}

// ACC = false: per-CTA column partials [gridDim.x][3][H] for ln_bwd_finalize_vec_kernel (deterministic order).
// ACC = true : the CTA totals are added straight into dgamma / dbeta / dcolsum with vector fp32 atomics (destinations
//              zero-initialised or holding earlier contributions, e.g. flat-gradient views) -- no second kernel.
template <typename TX, typename TD, int NV, bool ACC>
__global__ void __launch_bounds__(256)
ln_bwd_vec_kernel(const float* __restrict__ dy, const void* __restrict__ x, const float* __restrict__ gamma,
                  const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ dres,
                  float* __restrict__ dx32, void* __restrict__ dx16, float drop_p, unsigned long long drop_seed_,
                  const unsigned long long* __restrict__ drop_seed_ptr, float* __restrict__ partial, int M,
                  int rows_per_cta, float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dcolsum) {
  constexpr int H = NV * 128;
  extern __shared__ float sacc[];  // [warps][3][H]
  pdl_wait();
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nw = blockDim.x >> 5;       // 4 (partials + finalize) or 8 (atomic accumulate) warps
  float4 ag[NV], ab[NV], ac[NV], gm[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    ag[i] = ab[i] = ac[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    gm[i] = __ldg(reinterpret_cast<const float4*>(gamma + i * 128 + lane * 4));
  }
  const int r0 = blockIdx.x * rows_per_cta;
  const int r1 = min(M, r0 + rows_per_cta);
  const float keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const unsigned long long drop_seed = drop_p > 0.f ? eff_seed(drop_seed_, drop_seed_ptr) : 0ull;
  for (int row = r0 + warp; row < r1; row += nw) {
    const float mu = mean[row], rs = rstd[row];
    float4 xh[NV], g[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const size_t o = (size_t)row * H + i * 128 + lane * 4;
      const float4 d = *reinterpret_cast<const float4*>(dy + o);
      const float4 xv = ld4<TX>(x, o);
      xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
      g[i] = make_float4(d.x * gm[i].x, d.y * gm[i].y, d.z * gm[i].z, d.w * gm[i].w);
      ag[i].x += d.x * xh[i].x; ag[i].y += d.y * xh[i].y; ag[i].z += d.z * xh[i].z; ag[i].w += d.w * xh[i].w;
      ab[i].x += d.x; ab[i].y += d.y; ab[i].z += d.z; ab[i].w += d.w;
      s1 += g[i].x + g[i].y + g[i].z + g[i].w;
      s2 += g[i].x * xh[i].x + g[i].y * xh[i].y + g[i].z * xh[i].z + g[i].w * xh[i].w;
    }
    s1 = warp_sum(s1) * (1.f / H);
    s2 = warp_sum(s2) * (1.f / H);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = i * 128 + lane * 4;
      const size_t o = (size_t)row * H + c;
      float4 d;
      d.x = rs * (g[i].x - s1 - xh[i].x * s2);
      d.y = rs * (g[i].y - s1 - xh[i].y * s2);
      d.z = rs * (g[i].z - s1 - xh[i].z * s2);
      d.w = rs * (g[i].w - s1 - xh[i].w * s2);
      if (dx32) {
        float4 w = d;
        if (dres) {
          const float4 rr = *reinterpret_cast<const float4*>(dres + o);
          w.x += rr.x; w.y += rr.y; w.z += rr.z; w.w += rr.w;
        }
        *reinterpret_cast<float4*>(dx32 + o) = w;
      }
      float4 dm = d;
      if (drop_p > 0.f) {
        float v4[4] = {d.x, d.y, d.z, d.w};
        drop_apply4(v4, drop_seed, (unsigned long long)o, drop_thr16(drop_p), keep);
        dm = make_float4(v4[0], v4[1], v4[2], v4[3]);
      }
      if (dx16) {
        st4<TD>(dx16, o, dm);
        if constexpr (sizeof(TD) == 2) {   // the column sum must see the rounded values the GEMMs will read
          dm.x = to_f<TD>(from_f<TD>(dm.x)); dm.y = to_f<TD>(from_f<TD>(dm.y));
          dm.z = to_f<TD>(from_f<TD>(dm.z)); dm.w = to_f<TD>(from_f<TD>(dm.w));
        }
      }
      ac[i].x += dm.x; ac[i].y += dm.y; ac[i].z += dm.z; ac[i].w += dm.w;
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = i * 128 + lane * 4;
    *reinterpret_cast<float4*>(&sacc[(warp * 3 + 0) * H + c]) = ag[i];
    *reinterpret_cast<float4*>(&sacc[(warp * 3 + 1) * H + c]) = ab[i];
    *reinterpret_cast<float4*>(&sacc[(warp * 3 + 2) * H + c]) = ac[i];
  }
  __syncthreads();
  if constexpr (ACC) {
    for (int e = threadIdx.x * 4; e < 3 * H; e += blockDim.x * 4) {
      float4 t = *reinterpret_cast<const float4*>(&sacc[e]);
      for (int w = 1; w < nw; ++w) {
        const float4 u = *reinterpret_cast<const float4*>(&sacc[w * 3 * H + e]);
        t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
      }
      const int k = e / H, c = e - k * H;
      float* out = k == 0 ? dgamma : (k == 1 ? dbeta : dcolsum);
      if (out) atomicAdd(reinterpret_cast<float4*>(out + c), t);
    }
  } else {
    for (int e = threadIdx.x; e < 3 * H; e += blockDim.x) {
      float t = 0.f;
      for (int w = 0; w < nw; ++w) t += sacc[w * 3 * H + e];
      partial[(size_t)blockIdx.x * 3 * H + e] = t;
    }
  }
}

// partial [nparts][3][H] -> dgamma / dbeta / dcolsum; block (32 columns, 8 part-groups), grid (H/32, 3)
__global__ void __launch_bounds__(256)
ln_bwd_finalize_vec_kernel(const float* __restrict__ partial, int nparts, int H, float* dgamma, float* dbeta,
                           float* dcolsum) {
  __shared__ float red[8][33];
  pdl_wait();
  pdl_launch_dependents();
  const int k = blockIdx.y;
  float* out = k == 0 ? dgamma : (k == 1 ? dbeta : dcolsum);
  if (!out) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float t = 0.f;
  if (c < H)
    for (int p = ty; p < nparts; p += 8) t += partial[((size_t)p * 3 + k) * H + c];
  red[ty][tx] = t;
  __syncthreads();
  if (ty == 0 && c < H) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) a += red[w][tx];
    out[c] = a;
  }
}

inline int ln_bwd_vec_parts(int M) {
  int parts = (M + 15) / 16;      // >= 16 rows per CTA (2 per warp)
  if (parts > 296) parts = 296;   // 2 CTAs per SM on 148 SMs
  if (parts < 1) parts = 1;
  return parts;
}

inline int ln_bwd_parts(int M) {
  int parts = (M + 31) / 32;   // >= 32 rows per CTA
  if (parts > 592) parts = 592;  // 4 CTAs per SM on 148 SMs
  if (parts < 1) parts = 1;
  return parts;
}

// ---------------------------------------------------------------------------------------------
// column sums: partial[rowchunk][N] then fixed-order reduce
template <typename T>
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const void* __restrict__ x, int M, int N, int ld, int rows_per_cta, float* __restrict__ partial) {
  pdl_wait();
  pdl_launch_dependents();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= N) return;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
  float s = 0.f;
  for (int r = r0; r < r1; ++r) s += ldx<T>(x, (size_t)r * ld + c);
  partial[(size_t)blockIdx.y * N + c] = s;
}
__global__ void colsum_final_kernel(const float* __restrict__ partial, int nparts, int N, float* __restrict__ out) {
  pdl_wait();
  pdl_launch_dependents();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += partial[(size_t)p * N + c];
  out[c] = s;
}
inline int colsum_parts(int M) {
  int parts = (M + 63) / 64;
  if (parts > 128) parts = 128;
  if (parts < 1) parts = 1;
  return parts;
}

// One-kernel column sums accumulated with fp32 atomics into a zero-initialised (or partially filled) destination -- the
// bias gradients land directly in the flat gradient buffer, which the optimizer kernel clears every step.
// Block = 8 warps; a warp reads whole 512-byte row segments (16 bytes per lane), blockIdx.x picks the 32*VEC-column
// group, blockIdx.y the row chunk; per-warp register sums -> smem -> one atomic per column and block.
template <typename T>
__global__ void __launch_bounds__(256)
colsum_atomic_kernel(const T* __restrict__ x, int M, int N, int ld, int rows_per_cta, float* __restrict__ out) {
  constexpr int VEC = 16 / sizeof(T);
  __shared__ float red[8][32 * VEC + 1];
  pdl_wait();
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = blockIdx.x * (32 * VEC) + lane * VEC;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
  float acc[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
  if (c0 < N) {   // N % VEC == 0: a lane's VEC columns are all in range or all out
#pragma unroll 4
    for (int r = r0 + warp; r < r1; r += 8) {
      const uint4 w = *reinterpret_cast<const uint4*>(x + (size_t)r * ld + c0);
      if constexpr (sizeof(T) == 4) {
        acc[0] += __uint_as_float(w.x); acc[1] += __uint_as_float(w.y); acc[2] += __uint_as_float(w.z); acc[3] += __uint_as_float(w.w);
      } else {
        const float2 a = unpack2<T>(w.x), b = unpack2<T>(w.y), c = unpack2<T>(w.z), d = unpack2<T>(w.w);
        acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y; acc[4] += c.x; acc[5] += c.y; acc[6] += d.x; acc[7] += d.y;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < VEC; ++j) red[warp][lane * VEC + j] = acc[j];
  __syncthreads();
  for (int c = threadIdx.x; c < 32 * VEC; c += 256) {
    const int col = blockIdx.x * (32 * VEC) + c;
    if (col < N) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += red[w][c];
      atomicAdd(out + col, t);
    }
  }
}

// ---------------------------------------------------------------------------------------------
template <typename S, typename Dt>
__global__ void cast_kernel(const S* __restrict__ src, Dt* __restrict__ dst, long long n, float drop_p,
                            unsigned long long drop_seed_, const unsigned long long* __restrict__ drop_seed_ptr) {
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const long long e = i + 4 < n ? i + 4 : n;
  if (drop_p > 0.f) {
    const unsigned long long seed = eff_seed(drop_seed_, drop_seed_ptr);
    const float keep = 1.f / (1.f - drop_p);
    for (; i < e; ++i)
      dst[i] = from_f<Dt>(drop_keep(seed, (unsigned long long)i, drop_thr16(drop_p)) ? to_f<S>(src[i]) * keep : 0.f);
  } else {
    for (; i < e; ++i) dst[i] = from_f<Dt>(to_f<S>(src[i]));
  }
}

template <typename S>
int cast_from(const void* src, void* dst, int dd, long long n, float p, unsigned long long seed,
              const unsigned long long* seed_ptr, cudaStream_t st) {
  const int threads = 256;
  const long long blocks = (n + threads * 4 - 1) / (threads * 4);
  if (dd == GOAT_F32)
    cast_kernel<S, float><<<(unsigned)blocks, threads, 0, st>>>((const S*)src, (float*)dst, n, p, seed, seed_ptr);
  else if (dd == GOAT_F16)
    cast_kernel<S, __half><<<(unsigned)blocks, threads, 0, st>>>((const S*)src, (__half*)dst, n, p, seed, seed_ptr);
  else if (dd == GOAT_BF16)
    cast_kernel<S, __nv_bfloat16><<<(unsigned)blocks, threads, 0, st>>>((const S*)src, (__nv_bfloat16*)dst, n, p, seed, seed_ptr);
  else GOAT_CHECK(false, "goat_cast: bad dst dtype");
  GOAT_LAUNCH_CHECK();
  return GOAT_OK;
}

}  // namespace
}  // namespace goat

using namespace goat;

extern "C" int goat_layernorm_fwd(const void* x, int x_dtype, const float* gamma, const float* beta, float eps,
                                  float* y32, void* y16, int y16_dtype, float* mean, float* rstd, int M, int H,
                                  goat_stream_t stream) {
  GOAT_CHECK(x && gamma && beta, "goat_layernorm_fwd: null x/gamma/beta");
  GOAT_CHECK(y32 || y16, "goat_layernorm_fwd: no output requested");
  GOAT_CHECK(H > 0 && H <= 32 * LN_MAX_PER_LANE, "goat_layernorm_fwd: H=%d unsupported (max %d)", H, 32 * LN_MAX_PER_LANE);
  GOAT_CHECK(!y16 || y16_dtype == GOAT_F16 || y16_dtype == GOAT_BF16, "goat_layernorm_fwd: y16 dtype must be F16/BF16");
  if (M <= 0) return GOAT_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (H == 768 && aligned16(x) && (!y32 || aligned16(y32)) && (!y16 || aligned16(y16)) && aligned16(gamma) && aligned16(beta)) {
    dim3 g8((M + 7) / 8);
#define LN_FWDV(TX, TY) GOAT_CUDA(launch_pdl(ln_fwd_vec_kernel<TX, TY, 6>, g8, dim3(256), (size_t)0, st, x, gamma, beta, eps, y32, y16, mean, rstd, M))
    const bool yh8 = (y16_dtype == GOAT_F16);
    if (x_dtype == GOAT_F32) { if (yh8) LN_FWDV(float, __half); else LN_FWDV(float, __nv_bfloat16); }
    else if (x_dtype == GOAT_F16) { if (yh8) LN_FWDV(__half, __half); else LN_FWDV(__half, __nv_bfloat16); }
    else if (x_dtype == GOAT_BF16) { if (yh8) LN_FWDV(__nv_bfloat16, __half); else LN_FWDV(__nv_bfloat16, __nv_bfloat16); }
    else GOAT_CHECK(false, "goat_layernorm_fwd: bad x dtype");
#undef LN_FWDV
    GOAT_LAUNCH_CHECK();
    return GOAT_OK;
  }
  dim3 grid((M + LN_ROWS_PER_CTA - 1) / LN_ROWS_PER_CTA);
#define LN_FWD(TX, TY) ln_fwd_kernel<TX, TY><<<grid, 128, 0, st>>>(x, gamma, beta, eps, y32, y16, mean, rstd, M, H)
  const bool yh = (y16_dtype == GOAT_F16);
  if (x_dtype == GOAT_F32) { if (yh) LN_FWD(float, __half); else LN_FWD(float, __nv_bfloat16); }
  else if (x_dtype == GOAT_F16) { if (yh) LN_FWD(__half, __half); else LN_FWD(__half, __nv_bfloat16); }
  else if (x_dtype == GOAT_BF16) { if (yh) LN_FWD(__nv_bfloat16, __half); else LN_FWD(__nv_bfloat16, __nv_bfloat16); }
  else GOAT_CHECK(false, "goat_layernorm_fwd: bad x dtype");
#undef LN_FWD
  GOAT_LAUNCH_CHECK();
  return GOAT_OK;
}

extern "C" size_t goat_layernorm_bwd_workspace_bytes(int M, int H) {
  const int parts = ln_bwd_parts(M) > ln_bwd_vec_parts(M) ? ln_bwd_parts(M) : ln_bwd_vec_parts(M);
  return (size_t)parts * 3 * (size_t)H * sizeof(float);
}

extern "C" int goat_layernorm_bwd(const float* dy, const void* x, int x_dtype, const float* gamma, const float* mean,
                                  const float* rstd, const float* dres, float* dx32, void* dx16, int dx16_dtype,
                                  float drop_p, uint64_t drop_seed, const uint64_t* drop_seed_ptr, float* dgamma,
                                  float* dbeta, float* dcolsum, void* workspace, int M, int H, goat_stream_t stream) {
  GOAT_CHECK(dy && x && gamma && mean && rstd && workspace, "goat_layernorm_bwd: null argument");
  GOAT_CHECK(H > 0 && H <= 32 * LN_MAX_PER_LANE, "goat_layernorm_bwd: H=%d unsupported", H);
  GOAT_CHECK(!dx16 || dx16_dtype == GOAT_F16 || dx16_dtype == GOAT_BF16 || dx16_dtype == GOAT_F32,
             "goat_layernorm_bwd: bad dx16 dtype");
  GOAT_CHECK(drop_p >= 0.f && drop_p < 1.f, "goat_layernorm_bwd: drop_p out of range");
  if (M <= 0) return GOAT_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* partial = reinterpret_cast<float*>(workspace);
  if (H == 768 && aligned16(dy) && aligned16(x) && aligned16(gamma) && (!dres || aligned16(dres)) &&
      (!dx32 || aligned16(dx32)) && (!dx16 || aligned16(dx16))) {
    const int vparts = ln_bwd_vec_parts(M);
    const int vrows = (M + vparts - 1) / vparts;
    constexpr int VSMEM = 4 * 3 * 768 * 4;
    const int ddv = dx16 ? dx16_dtype : GOAT_F16;
#define LN_BWDV(TX, TD)                                                                                               \
  do {                                                                                                                \
    static bool cfgv = false;                                                                                         \
    if (!cfgv) {                                                                                                      \
      GOAT_CUDA(cudaFuncSetAttribute(ln_bwd_vec_kernel<TX, TD, 6, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, VSMEM)); \
      cfgv = true;                                                                                                    \
    }                                                                                                                 \
    GOAT_CUDA(launch_pdl(ln_bwd_vec_kernel<TX, TD, 6, false>, dim3(vparts), dim3(128), (size_t)VSMEM, st, dy, x, gamma, mean, rstd, \
        dres, dx32, dx16, drop_p, (unsigned long long)drop_seed,                                                      \
        reinterpret_cast<const unsigned long long*>(drop_seed_ptr), partial, M, vrows, (float*)nullptr, (float*)nullptr, \
        (float*)nullptr));                                                                                            \
  } while (0)
#define LN_BWDV_X(TX)                                                     \
  do {                                                                    \
    if (ddv == GOAT_F16) LN_BWDV(TX, __half);                             \
    else if (ddv == GOAT_BF16) LN_BWDV(TX, __nv_bfloat16);                \
    else LN_BWDV(TX, float);                                              \
  } while (0)
    if (x_dtype == GOAT_F32) LN_BWDV_X(float);
    else if (x_dtype == GOAT_F16) LN_BWDV_X(__half);
    else if (x_dtype == GOAT_BF16) LN_BWDV_X(__nv_bfloat16);
    else GOAT_CHECK(false, "goat_layernorm_bwd: bad x dtype");
#undef LN_BWDV_X
#undef LN_BWDV
    GOAT_LAUNCH_CHECK();
    GOAT_CUDA(launch_pdl(ln_bwd_finalize_vec_kernel, dim3(768 / 32, 3), dim3(256), (size_t)0, st, (const float*)partial, vparts, H,
                         dgamma, dbeta, dcolsum));
    GOAT_LAUNCH_CHECK();
    return GOAT_OK;
  }
  const int parts = ln_bwd_parts(M);
  const int rows_per_cta = (M + parts - 1) / parts;
  const size_t smem = (size_t)4 * 3 * H * sizeof(float);
#define LN_BWD(TX, TD)                                                                                                \
  do {                                                                                                                \
    static bool cfg = false;                                                                                          \
    if (!cfg && smem > 48 * 1024) {                                                                                   \
      GOAT_CUDA(cudaFuncSetAttribute(ln_bwd_kernel<TX, TD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 3 * 1024 * 4)); \
      cfg = true;                                                                                                     \
    }                                                                                                                 \
    ln_bwd_kernel<TX, TD><<<parts, 128, smem, st>>>(dy, x, gamma, mean, rstd, dres, dx32, dx16, drop_p, drop_seed,    \
                                                    reinterpret_cast<const unsigned long long*>(drop_seed_ptr),      \
                                                    partial, M, H, rows_per_cta);                                     \
  } while (0)
  const int dd = dx16 ? dx16_dtype : GOAT_F16;
#define LN_BWD_X(TX)                                                      \
  do {                                                                    \
    if (dd == GOAT_F16) LN_BWD(TX, __half);                               \
    else if (dd == GOAT_BF16) LN_BWD(TX, __nv_bfloat16);                  \
    else LN_BWD(TX, float);                                               \
  } while (0)
  if (x_dtype == GOAT_F32) LN_BWD_X(float);
  else if (x_dtype == GOAT_F16) LN_BWD_X(__half);
  else if (x_dtype == GOAT_BF16) LN_BWD_X(__nv_bfloat16);
  else GOAT_CHECK(false, "goat_layernorm_bwd: bad x dtype");
#undef LN_BWD_X
#undef LN_BWD
  GOAT_LAUNCH_CHECK();
  ln_bwd_finalize_kernel<<<(H + 127) / 128, 128, 0, st>>>(partial, parts, H, dgamma, dbeta, dcolsum);
  GOAT_LAUNCH_CHECK();
  return GOAT_OK;
}

extern "C" int goat_layernorm_bwd_acc(const float* dy, const void* x, int x_dtype, const float* gamma, const float* mean,
                                      const float* rstd, const float* dres, float* dx32, void* dx16, int dx16_dtype,
                                      float drop_p, uint64_t drop_seed, const uint64_t* drop_seed_ptr, float* dgamma,
                                      float* dbeta, float* dcolsum, int M, int H, goat_stream_t stream) {
  GOAT_CHECK(dy && x && gamma && mean && rstd, "goat_layernorm_bwd_acc: null argument");
  GOAT_CHECK(!dx16 || dx16_dtype == GOAT_F16 || dx16_dtype == GOAT_BF16 || dx16_dtype == GOAT_F32,
             "goat_layernorm_bwd_acc: bad dx16 dtype");
  GOAT_CHECK(drop_p >= 0.f && drop_p < 1.f, "goat_layernorm_bwd_acc: drop_p out of range");
  if (!(H == 768 && aligned16(dy) && aligned16(x) && aligned16(gamma) && (!dres || aligned16(dres)) &&
        (!dx32 || aligned16(dx32)) && (!dx16 || aligned16(dx16)) && (!dgamma || aligned16(dgamma)) &&
        (!dbeta || aligned16(dbeta)) && (!dcolsum || aligned16(dcolsum)))) {
    set_error("goat_layernorm_bwd_acc: needs H == 768 and 16-byte aligned tensors (use goat_layernorm_bwd)");
    return GOAT_ERR_UNSUPPORTED;
  }
  if (M <= 0) return GOAT_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // 8 warps per CTA, ONE CTA per SM: measured at 5120 x 768 (B200): 74 CTAs 17.2 us, 148 CTAs 10.9 us (5.0 TB/s),
  // 222 CTAs 14.2 us, 296 CTAs 13.5 us, 444 CTAs 15.4 us -- past one CTA per SM the same-address atomics of the column
  // sums cost more than the extra loads in flight buy.  GOAT_LN_PARTS overrides (tuning).
  static const int max_parts = [] { const char* e = getenv("GOAT_LN_PARTS"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 148; }();
  int vparts = (M + 15) / 16;
  if (vparts > max_parts) vparts = max_parts;
  const int vrows = (M + vparts - 1) / vparts;
  constexpr int VSMEM = 8 * 3 * 768 * 4;
  const int ddv = dx16 ? dx16_dtype : GOAT_F16;
#define LN_BWDA(TX, TD)                                                                                               \
  do {                                                                                                                \
    static bool cfga = false;                                                                                         \
    if (!cfga) {                                                                                                      \
      GOAT_CUDA(cudaFuncSetAttribute(ln_bwd_vec_kernel<TX, TD, 6, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, VSMEM)); \
      cfga = true;                                                                                                    \
    }                                                                                                                 \
    GOAT_CUDA(launch_pdl(ln_bwd_vec_kernel<TX, TD, 6, true>, dim3(vparts), dim3(256), (size_t)VSMEM, st, dy, x, gamma, mean, rstd, \
        dres, dx32, dx16, drop_p, (unsigned long long)drop_seed,                                                      \
        reinterpret_cast<const unsigned long long*>(drop_seed_ptr), (float*)nullptr, M, vrows, dgamma, dbeta, dcolsum)); \
  } while (0)
#define LN_BWDA_X(TX)                                                     \
  do {                                                                    \
    if (ddv == GOAT_F16) LN_BWDA(TX, __half);                             \
    else if (ddv == GOAT_BF16) LN_BWDA(TX, __nv_bfloat16);                \
    else LN_BWDA(TX, float);                                              \
  } while (0)
  if (x_dtype == GOAT_F32) LN_BWDA_X(float);
  else if (x_dtype == GOAT_F16) LN_BWDA_X(__half);
  else if (x_dtype == GOAT_BF16) LN_BWDA_X(__nv_bfloat16);
  else GOAT_CHECK(false, "goat_layernorm_bwd_acc: bad x dtype");
#undef LN_BWDA_X
#undef LN_BWDA
  return GOAT_OK;
}

extern "C" size_t goat_colsum_workspace_bytes(int M, int N) { return (size_t)colsum_parts(M) * (size_t)N * sizeof(float); }

extern "C" int goat_colsum(const void* x, int dtype, int M, int N, int ld, float* out, void* workspace,
                           goat_stream_t stream) {
  GOAT_CHECK(x && out && workspace, "goat_colsum: null argument");
  GOAT_CHECK(N > 0 && ld >= N, "goat_colsum: bad N/ld");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (M <= 0) {
    GOAT_CUDA(cudaMemsetAsync(out, 0, (size_t)N * sizeof(float), st));
    return GOAT_OK;
  }
  const int parts = colsum_parts(M);
  const int rows_per_cta = (M + parts - 1) / parts;
  dim3 grid((N + 255) / 256, parts);
  float* partial = reinterpret_cast<float*>(workspace);
  if (dtype == GOAT_F32) GOAT_CUDA(launch_pdl(colsum_partial_kernel<float>, grid, dim3(256), (size_t)0, st, x, M, N, ld, rows_per_cta, partial));
  else if (dtype == GOAT_F16) GOAT_CUDA(launch_pdl(colsum_partial_kernel<__half>, grid, dim3(256), (size_t)0, st, x, M, N, ld, rows_per_cta, partial));
  else if (dtype == GOAT_BF16) GOAT_CUDA(launch_pdl(colsum_partial_kernel<__nv_bfloat16>, grid, dim3(256), (size_t)0, st, x, M, N, ld, rows_per_cta, partial));
  else GOAT_CHECK(false, "goat_colsum: bad dtype");
  GOAT_LAUNCH_CHECK();
  GOAT_CUDA(launch_pdl(colsum_final_kernel, dim3((N + 127) / 128), dim3(128), (size_t)0, st, (const float*)partial, parts, N, out));
  GOAT_LAUNCH_CHECK();
  return GOAT_OK;
}

extern "C" int goat_colsum_acc(const void* x, int dtype, int M, int N, int ld, float* out, goat_stream_t stream) {
  GOAT_CHECK(x && out, "goat_colsum_acc: null argument");
  GOAT_CHECK(dtype == GOAT_F32 || dtype == GOAT_F16 || dtype == GOAT_BF16, "goat_colsum_acc: bad dtype");
  const int vec = 16 / dtype_size(dtype);
  GOAT_CHECK(N > 0 && ld >= N && (N % vec) == 0 && (ld % vec) == 0 && aligned16(x),
             "goat_colsum_acc: N and ld must be multiples of %d and x 16-byte aligned", vec);
  if (M <= 0) return GOAT_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int gx = (N + 32 * vec - 1) / (32 * vec);
  int gy = (4 * 148 + gx - 1) / gx;                 // about four CTAs (32 warps) per SM in total
  const int max_gy = (M + 15) / 16;                 // at least 16 rows (2 per warp) per CTA
  if (gy > max_gy) gy = max_gy;
  if (gy < 1) gy = 1;
  const int rows_per_cta = (M + gy - 1) / gy;
  gy = (M + rows_per_cta - 1) / rows_per_cta;
  dim3 grid(gx, gy);
  if (dtype == GOAT_F32)
    GOAT_CUDA(launch_pdl(colsum_atomic_kernel<float>, grid, dim3(256), (size_t)0, st, (const float*)x, M, N, ld, rows_per_cta, out));
  else if (dtype == GOAT_F16)
    GOAT_CUDA(launch_pdl(colsum_atomic_kernel<__half>, grid, dim3(256), (size_t)0, st, (const __half*)x, M, N, ld, rows_per_cta, out));
  else
    GOAT_CUDA(launch_pdl(colsum_atomic_kernel<__nv_bfloat16>, grid, dim3(256), (size_t)0, st, (const __nv_bfloat16*)x, M, N, ld, rows_per_cta, out));
  return GOAT_OK;
}

extern "C" int goat_dropout_cast(const void* src, int src_dtype, void* dst, int dst_dtype, long long n, float drop_p,
                                 uint64_t drop_seed, const uint64_t* drop_seed_ptr, goat_stream_t stream) {
  GOAT_CHECK(src && dst, "goat_cast: null argument");
  GOAT_CHECK(drop_p >= 0.f && drop_p < 1.f, "goat_dropout_cast: drop_p out of range");
  if (n <= 0) return GOAT_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const unsigned long long* sp = reinterpret_cast<const unsigned long long*>(drop_seed_ptr);
  if (src_dtype == GOAT_F32) return cast_from<float>(src, dst, dst_dtype, n, drop_p, drop_seed, sp, st);
  if (src_dtype == GOAT_F16) return cast_from<__half>(src, dst, dst_dtype, n, drop_p, drop_seed, sp, st);
  if (src_dtype == GOAT_BF16) return cast_from<__nv_bfloat16>(src, dst, dst_dtype, n, drop_p, drop_seed, sp, st);
  GOAT_CHECK(false, "goat_cast: bad src dtype");
}

extern "C" int goat_cast(const void* src, int src_dtype, void* dst, int dst_dtype, long long n, goat_stream_t stream) {
  return goat_dropout_cast(src, src_dtype, dst, dst_dtype, n, 0.f, 0, nullptr, stream);
}
