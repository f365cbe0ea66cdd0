// Small fp32 kernels around the transformer blocks: adaptive panorama fusion / CFP attention pooling, p(z)-weighted
// dictionary sums (BACL), the "door" gate (BACL / FACL), row-wise softmax cross-entropy (SAP / MLM / InfoNCE),
// gather-and-reduce over index lists (global-map aggregation, logit fusion) and the RoBERTa input embeddings.
// All of them are tiny next to the GEMMs (O(tokens x 768) bytes): one CTA per batch row or one warp per token,
// coalesced fp32 accesses along the hidden dimension, fp32 math throughout.  Parameter gradients that reduce over
// the batch (pool vectors, gate vectors, embedding tables) are accumulated with atomics into caller-zeroed buffers
// (the flat gradient buffer of engine.FlatParams is zero at step start).
#include <math_constants.h>

#include "common.cuh"

namespace goat {
namespace {

constexpr int POOL_THREADS = 256;
constexpr int POOL_MAX_N = 1024;
constexpr int POOL_MAX_H = 1024;

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[w];
  return t;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = -CUDART_INF_F;
  for (int w = 0; w < (blockDim.x >> 5); ++w) t = fmaxf(t, red[w]);
  return t;
}

// ---------------------------------------------------------------------------------------------
// attention pooling.  mode 0 (pano fusion):  s_n = tanh(x_n . w + b)      out = sum_n a_n x_n
//                     mode 1 (CFP pooling):  s_n = tanh(x_n) . w          out = tanh(sum_n a_n x_n)
// a = softmax over the N tokens of one batch row (no mask: the reference pools over padding too).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(POOL_THREADS)
attn_pool_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, int mode,
                     int Nfull, int H, float* __restrict__ out, float* __restrict__ a_out, float* __restrict__ s_out,
                     const int* __restrict__ n_valid) {
  __shared__ float sc[POOL_MAX_N];
  __shared__ float red[POOL_THREADS / 32];
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* xb = x + (size_t)b * Nfull * H;
  // only the first N tokens are pooled (N = the batch's own padded length when the buffers are padded further)
  const int N = n_valid ? max(1, min(Nfull, *n_valid)) : Nfull;
  for (int n = N + threadIdx.x; n < Nfull; n += POOL_THREADS) {
    if (s_out) s_out[(size_t)b * Nfull + n] = 0.f;
    a_out[(size_t)b * Nfull + n] = 0.f;
  }
  const float bv = (mode == 0 && bias) ? bias[0] : 0.f;
  for (int n = warp; n < N; n += POOL_THREADS / 32) {
    float d = 0.f;
    for (int h = lane; h < H; h += 32) {
      const float v = xb[(size_t)n * H + h];
      d += (mode == 0 ? v : tanhf(v)) * w[h];
    }
    d = warp_sum(d);
    if (lane == 0) sc[n] = mode == 0 ? tanhf(d + bv) : d;
  }
  __syncthreads();
  float m = -CUDART_INF_F;
  for (int n = threadIdx.x; n < N; n += POOL_THREADS) m = fmaxf(m, sc[n]);
  m = block_max(m, red);
  float l = 0.f;
  for (int n = threadIdx.x; n < N; n += POOL_THREADS) l += __expf(sc[n] - m);
  l = block_sum(l, red);
  const float inv = 1.f / l;
  __syncthreads();
  for (int n = threadIdx.x; n < N; n += POOL_THREADS) {
    const float s = sc[n];
    const float a = __expf(s - m) * inv;
    if (s_out) s_out[(size_t)b * Nfull + n] = s;
    a_out[(size_t)b * Nfull + n] = a;
    sc[n] = a;
  }
  __syncthreads();
  for (int h = threadIdx.x; h < H; h += POOL_THREADS) {
    float o = 0.f;
    for (int n = 0; n < N; ++n) o = fmaf(sc[n], xb[(size_t)n * H + h], o);
    out[(size_t)b * H + h] = mode == 0 ? o : tanhf(o);
  }
}

__global__ void __launch_bounds__(POOL_THREADS)
attn_pool_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ x, const float* __restrict__ w,
                     const float* __restrict__ a, const float* __restrict__ s, const float* __restrict__ out, int mode,
                     int Nfull, int H, float* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db,
                     const int* __restrict__ n_valid) {
  __shared__ float dpre[POOL_MAX_H];
  __shared__ float coef[POOL_MAX_N];   // da_n, then ds_n (mode 1) / du_n (mode 0)
  __shared__ float red[POOL_THREADS / 32];
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* xb = x + (size_t)b * Nfull * H;
  const float* ab = a + (size_t)b * Nfull;
  const int N = n_valid ? max(1, min(Nfull, *n_valid)) : Nfull;
  for (int i = N * H + threadIdx.x; i < Nfull * H; i += POOL_THREADS) dx[(size_t)b * Nfull * H + i] = 0.f;
  for (int h = threadIdx.x; h < H; h += POOL_THREADS) {
    float g = dout[(size_t)b * H + h];
    if (mode == 1) {
      const float o = out[(size_t)b * H + h];
      g *= 1.f - o * o;
    }
    dpre[h] = g;
  }
  __syncthreads();
  for (int n = warp; n < N; n += POOL_THREADS / 32) {
    float d = 0.f;
    for (int h = lane; h < H; h += 32) d = fmaf(dpre[h], xb[(size_t)n * H + h], d);
    d = warp_sum(d);
    if (lane == 0) coef[n] = d;
  }
  __syncthreads();
  float c = 0.f;
  for (int n = threadIdx.x; n < N; n += POOL_THREADS) c += ab[n] * coef[n];
  c = block_sum(c, red);
  float dbl = 0.f;
  __syncthreads();
  for (int n = threadIdx.x; n < N; n += POOL_THREADS) {
    float ds = ab[n] * (coef[n] - c);
    if (mode == 0) {
      const float sv = s[(size_t)b * Nfull + n];
      ds *= 1.f - sv * sv;
      dbl += ds;
    }
    coef[n] = ds;
  }
  if (mode == 0 && db) {
    dbl = block_sum(dbl, red);
    if (threadIdx.x == 0) atomicAdd(db, dbl);
  }
  __syncthreads();
  for (int h = threadIdx.x; h < H; h += POOL_THREADS) {
    const float wh = w[h], g = dpre[h];
    float dwh = 0.f;
    for (int n = 0; n < N; ++n) {
      const float xv = xb[(size_t)n * H + h];
      float d = ab[n] * g;
      if (mode == 0) {
        d = fmaf(coef[n], wh, d);
        dwh = fmaf(coef[n], xv, dwh);
      } else {
        const float t = tanhf(xv);
        d = fmaf(coef[n] * wh, 1.f - t * t, d);
        dwh = fmaf(coef[n], t, dwh);
      }
      dx[((size_t)b * Nfull + n) * H + h] = d;
    }
    if (dw) atomicAdd(dw + h, dwh);
  }
}

// ---------------------------------------------------------------------------------------------
// weighted token sum: out[b,:] = sum_n p[b,n] x[b,n,:]    (BACL: sum_z p(z) z)
// ---------------------------------------------------------------------------------------------
__global__ void wsum_fwd_kernel(const float* __restrict__ x, const float* __restrict__ p, int N, int H,
                                float* __restrict__ out) {
  const int b = blockIdx.y;
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  const float* xb = x + (size_t)b * N * H;
  float o = 0.f;
  for (int n = 0; n < N; ++n) o = fmaf(p[(size_t)b * N + n], xb[(size_t)n * H + h], o);
  out[(size_t)b * H + h] = o;
}
__global__ void wsum_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ p, int N, int H,
                                float* __restrict__ dx) {
  const int b = blockIdx.z, n = blockIdx.y;
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  dx[((size_t)b * N + n) * H + h] = p[(size_t)b * N + n] * dout[(size_t)b * H + h];
}

// ---------------------------------------------------------------------------------------------
// door gate: g = sigmoid(aug . wa + ba + ori . wo + bo);  out = g aug + (1 - g) ori     (one warp per row)
// ---------------------------------------------------------------------------------------------
constexpr int GATE_WARPS = 4;
constexpr int GATE_PER_LANE = 32;   // H <= 1024

__global__ void __launch_bounds__(GATE_WARPS * 32)
door_gate_fwd_kernel(const float* __restrict__ aug, const float* __restrict__ ori, const float* __restrict__ wa,
                     const float* __restrict__ ba, const float* __restrict__ wo, const float* __restrict__ bo, int M, int H,
                     float* __restrict__ out, float* __restrict__ gate) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * GATE_WARPS + warp;
  if (row >= M) return;
  const float* a = aug + (size_t)row * H;
  const float* o = ori + (size_t)row * H;
  float z = 0.f;
  for (int h = lane; h < H; h += 32) z += a[h] * wa[h] + o[h] * wo[h];
  z = warp_sum(z) + (ba ? ba[0] : 0.f) + (bo ? bo[0] : 0.f);
  const float g = 1.f / (1.f + __expf(-z));
  if (lane == 0) gate[row] = g;
  for (int h = lane; h < H; h += 32) out[(size_t)row * H + h] = g * a[h] + (1.f - g) * o[h];
}

__global__ void __launch_bounds__(GATE_WARPS * 32)
door_gate_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ aug, const float* __restrict__ ori,
                     const float* __restrict__ wa, const float* __restrict__ wo, const float* __restrict__ gate, int M, int H,
                     int rows_per_cta, float* __restrict__ daug, float* __restrict__ dori, float* __restrict__ dwa,
                     float* __restrict__ dwo, float* __restrict__ dba, float* __restrict__ dbo) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc_a[GATE_PER_LANE], acc_o[GATE_PER_LANE];
#pragma unroll
  for (int i = 0; i < GATE_PER_LANE; ++i) { acc_a[i] = 0.f; acc_o[i] = 0.f; }
  float acc_z = 0.f;
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
  for (int row = r0 + warp; row < r1; row += GATE_WARPS) {
    const float* a = aug + (size_t)row * H;
    const float* o = ori + (size_t)row * H;
    const float* d = dout + (size_t)row * H;
    float dg = 0.f;
    for (int h = lane; h < H; h += 32) dg = fmaf(d[h], a[h] - o[h], dg);
    dg = warp_sum(dg);
    const float g = gate[row];
    const float dz = dg * g * (1.f - g);
    acc_z += dz;
#pragma unroll
    for (int i = 0; i < GATE_PER_LANE; ++i) {
      const int h = i * 32 + lane;
      if (h < H) {
        daug[(size_t)row * H + h] = fmaf(g, d[h], dz * wa[h]);
        dori[(size_t)row * H + h] = fmaf(1.f - g, d[h], dz * wo[h]);
        acc_a[i] = fmaf(dz, a[h], acc_a[i]);
        acc_o[i] = fmaf(dz, o[h], acc_o[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < GATE_PER_LANE; ++i) {
    const int h = i * 32 + lane;
    if (h < H) {
      if (dwa && acc_a[i] != 0.f) atomicAdd(dwa + h, acc_a[i]);
      if (dwo && acc_o[i] != 0.f) atomicAdd(dwo + h, acc_o[i]);
    }
  }
  if (lane == 0 && acc_z != 0.f) {
    if (dba) atomicAdd(dba, acc_z);
    if (dbo) atomicAdd(dbo, acc_z);
  }
}

// ---------------------------------------------------------------------------------------------
// row-wise softmax cross-entropy over strided logits (element (i, j) at logits[i*sr + j*sc]):
//   lse_i = log sum_j exp(x_ij);  loss_i = lse_i - x_i,label_i;  label == ignore_index -> loss 0, no gradient.
// -inf logits (masked actions) contribute exp(-inf) = 0 exactly as in torch.
// ---------------------------------------------------------------------------------------------
constexpr int XENT_THREADS = 256;

__global__ void __launch_bounds__(XENT_THREADS)
xent_fwd_kernel(const float* __restrict__ logits, long long sr, long long sc, const long long* __restrict__ labels, int N,
                long long ignore_index, float* __restrict__ loss, float* __restrict__ lse) {
  __shared__ float red[XENT_THREADS / 32];
  const int i = blockIdx.x;
  const float* row = logits + (size_t)i * sr;
  float m = -CUDART_INF_F;
  for (int j = threadIdx.x; j < N; j += XENT_THREADS) m = fmaxf(m, row[(size_t)j * sc]);
  m = block_max(m, red);
  float l = 0.f;
  for (int j = threadIdx.x; j < N; j += XENT_THREADS) l += __expf(row[(size_t)j * sc] - m);
  l = block_sum(l, red);
  if (threadIdx.x == 0) {
    const float e = m + __logf(l);
    lse[i] = e;
    const long long lab = labels[i];
    // a label outside [0, N) that is not ignore_index is a caller bug (torch raises): no out-of-bounds read, NaN loss
    loss[i] = (lab == ignore_index) ? 0.f : ((lab < 0 || lab >= N) ? CUDART_NAN_F : e - row[(size_t)lab * sc]);
  }
}

__global__ void __launch_bounds__(XENT_THREADS)
xent_bwd_kernel(const float* __restrict__ dloss, const float* __restrict__ logits, long long sr, long long sc,
                const long long* __restrict__ labels, const float* __restrict__ lse, int N, long long ignore_index,
                float* __restrict__ dlogits, long long dsr, long long dsc, int accumulate) {
  const int i = blockIdx.x;
  const long long lab = labels[i];
  const float g = (lab == ignore_index || lab < 0 || lab >= N) ? 0.f : dloss[i];
  const float e = lse[i];
  const float* row = logits + (size_t)i * sr;
  float* drow = dlogits + (size_t)i * dsr;
  for (int j = threadIdx.x + blockIdx.y * XENT_THREADS; j < N; j += XENT_THREADS * gridDim.y) {
    float d = 0.f;
    if (g != 0.f) d = g * (__expf(row[(size_t)j * sc] - e) - (j == lab ? 1.f : 0.f));
    if (accumulate) drow[(size_t)j * dsc] += d;
    else drow[(size_t)j * dsc] = d;
  }
}

// ---------------------------------------------------------------------------------------------
// chunked cross-entropy over a wide class axis (the 50265-word vocabulary of the MLM head): the logits exist only one
// column chunk [c0, c0 + Nc) at a time.  forward keeps running (max, sum exp, picked logit) per row across the chunks;
// backward turns a recomputed logits chunk into the 16-bit gradient operand of the dgrad / wgrad GEMMs.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(XENT_THREADS)
xent_chunk_fwd_kernel(const float* __restrict__ logits, long long ld, const long long* __restrict__ labels, int Nc,
                      long long c0, int first, float* __restrict__ run_m, float* __restrict__ run_l,
                      float* __restrict__ picked) {
  __shared__ float red[XENT_THREADS / 32];
  const int i = blockIdx.x;
  const float* row = logits + (size_t)i * ld;
  float m = -CUDART_INF_F;
  for (int j = threadIdx.x; j < Nc; j += XENT_THREADS) m = fmaxf(m, row[j]);
  m = block_max(m, red);
  const float m_old = first ? -CUDART_INF_F : run_m[i];
  const float m_new = fmaxf(m, m_old);
  float l = 0.f;
  for (int j = threadIdx.x; j < Nc; j += XENT_THREADS) l += __expf(row[j] - m_new);
  l = block_sum(l, red);
  if (threadIdx.x == 0) {
    const float l_old = first ? 0.f : run_l[i];
    run_l[i] = l + ((m_old == -CUDART_INF_F) ? 0.f : l_old * __expf(m_old - m_new));
    run_m[i] = m_new;
    const long long lab = labels[i] - c0;
    if (first) picked[i] = 0.f;
    if (lab >= 0 && lab < Nc) picked[i] = row[lab];
  }
}

template <typename T>
__global__ void __launch_bounds__(XENT_THREADS)
xent_chunk_bwd_kernel(const float* __restrict__ dloss, const float* __restrict__ logits, long long ld,
                      const long long* __restrict__ labels, const float* __restrict__ lse, int Nc, long long c0,
                      long long ignore_index, long long n_classes, T* __restrict__ out, long long ldo) {
  const int i = blockIdx.x;
  const long long lab = labels[i];
  const float g = (lab == ignore_index || lab < 0 || lab >= n_classes) ? 0.f : dloss[i];
  const float e = lse[i];
  const float* row = logits + (size_t)i * ld;
  T* orow = out + (size_t)i * ldo;
  const long long hit = lab - c0;
  for (int j = threadIdx.x + blockIdx.y * XENT_THREADS; j < Nc; j += XENT_THREADS * gridDim.y) {
    float d = 0.f;
    if (g != 0.f) d = g * (__expf(row[j] - e) - (j == hit ? 1.f : 0.f));
    orow[j] = from_f<T>(d);
  }
}

// ---------------------------------------------------------------------------------------------
// gather-and-reduce: out[r,:] = scale_r * sum_{k < K, idx[r,k] >= 0} src[idx[r,k], :]
//   scale_r = 1 (sum) or 1 / #valid (mean; 0 valid entries -> zeros).
// ---------------------------------------------------------------------------------------------
__global__ void segment_reduce_fwd_kernel(const float* __restrict__ src, const int* __restrict__ idx, int K, int H,
                                          int mean, float* __restrict__ out) {
  const int r = blockIdx.y;
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  float o = 0.f;
  int cnt = 0;
  for (int k = 0; k < K; ++k) {
    const int s = idx[(size_t)r * K + k];
    if (s >= 0) {
      o += src[(size_t)s * H + h];
      ++cnt;
    }
  }
  if (mean && cnt > 1) o *= 1.f / (float)cnt;
  out[(size_t)r * H + h] = o;
}
__global__ void segment_reduce_bwd_kernel(const float* __restrict__ dout, const int* __restrict__ idx, int K, int H,
                                          int mean, float* __restrict__ dsrc) {
  const int r = blockIdx.y;
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  int cnt = 0;
  for (int k = 0; k < K; ++k) cnt += idx[(size_t)r * K + k] >= 0 ? 1 : 0;
  if (cnt == 0) return;
  const float g = dout[(size_t)r * H + h] * ((mean && cnt > 1) ? 1.f / (float)cnt : 1.f);
  for (int k = 0; k < K; ++k) {
    const int s = idx[(size_t)r * K + k];
    if (s >= 0) atomicAdd(dsrc + (size_t)s * H + h, g);
  }
}

// ---------------------------------------------------------------------------------------------
// RoBERTa input embeddings: out[m,:] = word[ids[m]] + pos[m % L] + type[0]     (LayerNorm + dropout follow)
// ---------------------------------------------------------------------------------------------
__global__ void embed_fwd_kernel(const long long* __restrict__ ids, const float* __restrict__ word,
                                 const float* __restrict__ pos, const float* __restrict__ type, int L, int H,
                                 float* __restrict__ out) {
  const int m = blockIdx.y;
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  const long long id = ids[m];
  out[(size_t)m * H + h] = word[(size_t)id * H + h] + pos[(size_t)(m % L) * H + h] + type[h];
}
__global__ void embed_bwd_kernel(const float* __restrict__ dout, const long long* __restrict__ ids, int L, int H,
                                 long long pad, float* __restrict__ dword, float* __restrict__ dpos,
                                 float* __restrict__ dtype) {
  const int m = blockIdx.y;
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  const float g = dout[(size_t)m * H + h];
  // nn.Embedding(padding_idx=pad) never accumulates a gradient into row `pad` (word AND position tables)
  if (dword && ids[m] != pad) atomicAdd(dword + (size_t)ids[m] * H + h, g);
  if (dpos && (long long)(m % L) != pad) atomicAdd(dpos + (size_t)(m % L) * H + h, g);
  if (dtype) atomicAdd(dtype + h, g);
}

// ---------------------------------------------------------------------------------------------
// out = dy * act'(ref) in one pass (heads / poolers); sprel_linear (1 -> 1) forward and its two scalar gradients
// ---------------------------------------------------------------------------------------------
template <typename R, typename Dt>
__global__ void act_grad_kernel(const float* __restrict__ dy, const R* __restrict__ ref, int act, Dt* __restrict__ out,
                                long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float g = dy[i];
    if (act == GOAT_ACT_RELU) g = to_f<R>(ref[i]) > 0.f ? g : 0.f;
    else if (act == GOAT_ACT_TANH) { const float y = to_f<R>(ref[i]); g *= 1.f - y * y; }
    else if (act == GOAT_ACT_GELU) g *= dgelu_erf(to_f<R>(ref[i]));
    out[i] = from_f<Dt>(g);
  }
}

// rows of 16-byte words gathered by index (feature bank: panorama rows of 36 x 768 16-bit features); idx < 0 -> zero row
__global__ void gather_rows_kernel(const uint4* __restrict__ src, const int* __restrict__ idx, long long words_per_row,
                                   uint4* __restrict__ out) {
  const int r = blockIdx.y;
  const long long s = idx[r];
  const uint4* srow = src + s * words_per_row;
  uint4* orow = out + (long long)r * words_per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < words_per_row; i += (long long)gridDim.x * blockDim.x)
    orow[i] = s >= 0 ? __ldg(srow + i) : make_uint4(0u, 0u, 0u, 0u);
}

__global__ void act_fwd_kernel(const float* __restrict__ x, int act, float* __restrict__ out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    out[i] = act == GOAT_ACT_GELU ? gelu_erf(v) : act == GOAT_ACT_RELU ? fmaxf(v, 0.f) : act == GOAT_ACT_TANH ? tanhf(v) : v;
  }
}

template <typename R>
int act_grad_launch(const float* dy, const void* ref, int act, void* out, int od, long long n, cudaStream_t st) {
  long long want = (n + 255) / 256;
  const int grid = (int)(want > 148 * 8 ? 148 * 8 : want);
  if (od == GOAT_F32) act_grad_kernel<R, float><<<grid, 256, 0, st>>>(dy, (const R*)ref, act, (float*)out, n);
  else if (od == GOAT_F16) act_grad_kernel<R, __half><<<grid, 256, 0, st>>>(dy, (const R*)ref, act, (__half*)out, n);
  else act_grad_kernel<R, __nv_bfloat16><<<grid, 256, 0, st>>>(dy, (const R*)ref, act, (__nv_bfloat16*)out, n);
  GOAT_LAUNCH_CHECK();
  return GOAT_OK;
}

__global__ void sprel_fwd_kernel(const float* __restrict__ d, const float* __restrict__ w, const float* __restrict__ b,
                                 float* __restrict__ out, long long n) {
  const float ww = w[0], bb = b[0];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = fmaf(d[i], ww, bb);
}
__global__ void __launch_bounds__(256)
sprel_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ d, float* __restrict__ dw, float* __restrict__ db,
                 long long n) {
  __shared__ float red[8];
  float sw = 0.f, sb = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float g = dout[i];
    sw = fmaf(g, d[i], sw);
    sb += g;
  }
  sw = block_sum(sw, red);
  sb = block_sum(sb, red);
  if (threadIdx.x == 0) {
    atomicAdd(dw, sw);
    atomicAdd(db, sb);
  }
}

}  // namespace
}  // namespace goat

using namespace goat;

extern "C" {

int goat_attn_pool_fwd(const float* x, const float* w, const float* bias, int mode, int B, int N, int H, float* out,
                       float* a, float* s, const int* n_valid, goat_stream_t stream) {
  GOAT_CHECK(x && w && out && a, "goat_attn_pool_fwd: null argument");
  GOAT_CHECK(mode == 0 || mode == 1, "goat_attn_pool_fwd: mode must be 0 (pano fusion) or 1 (CFP pooling)");
  GOAT_CHECK(mode == 1 || s, "goat_attn_pool_fwd: mode 0 needs the score buffer s");
  GOAT_CHECK(N >= 1 && N <= POOL_MAX_N && H >= 1 && H <= POOL_MAX_H, "goat_attn_pool_fwd: N=%d / H=%d out of range", N, H);
  if (B <= 0) return GOAT_OK;
  attn_pool_fwd_kernel<<<B, POOL_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, w, bias, mode, N, H, out, a, s,
                                                                                      n_valid);
  GOAT_LAUNCH_CHECK();
  return GOAT_OK;
}

int goat_attn_pool_bwd(const float* dout, const float* x, const float* w, const float* a, const float* s, const float* out,
                       int mode, int B, int N, int H, float* dx, float* dw, float* db, const int* n_valid,
                       goat_stream_t stream) {
  GOAT_CHECK(dout && x && w && a && dx, "goat_attn_pool_bwd: null argument");
  GOAT_CHECK(mode == 0 || mode == 1, "goat_attn_pool_bwd: bad mode");
  GOAT_CHECK(mode == 0 ? s != nullptr : out != nullptr, "goat_attn_pool_bwd: mode 0 needs s, mode 1 needs out");
  GOAT_CHECK(N >= 1 && N <= POOL_MAX_N && H >= 1 && H <= POOL_MAX_H, "goat_attn_pool_bwd: N=%d / H=%d out of range", N, H);
  if (B <= 0) return GOAT_OK;
  attn_pool_bwd_kernel<<<B, POOL_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dout, x, w, a, s, out, mode, N, H,
                                                                                      dx, dw, db, n_valid);
  GOAT_LAUNCH_CHECK();
  return GOAT_OK;
}

int goat_wsum_fwd(const float* x, const float* p, int B, int N, int H, float* out, goat_stream_t stream) {
  GOAT_CHECK(x && p && out, "goat_wsum_fwd: null argument");
  if (B <= 0 || H <= 0) return GOAT_OK;
  GOAT_CHECK(B <= 65535, "goat_wsum_fwd: B too large");
  wsum_fwd_kernel<<<dim3((H + 127) / 128, B), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, p, N, H, out);
  GOAT_LAUNCH_CHECK();
  return GOAT_OK;
}

int goat_wsum_bwd(const float* dout, const float* p, int B, int N, int H, float* dx, goat_stream_t stream) {
  GOAT_CHECK(dout && p && dx, "goat_wsum_bwd: null argument");
  if (B <= 0 || N <= 0 || H <= 0) return GOAT_OK;
  GOAT_CHECK(B <= 65535 && N <= 65535, "goat_wsum_bwd: B / N too large");
  wsum_bwd_kernel<<<dim3((H + 127) / 128, N, B), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dout, p, N, H, dx);
  GOAT_LAUNCH_CHECK();
  return GOAT_OK;
}

int goat_door_gate_fwd(const float* aug, const float* ori, const float* wa, const float* ba, const float* wo,
                       const float* bo, int M, int H, float* out, float* gate, goat_stream_t stream) {
  GOAT_CHECK(aug && ori && wa && wo && out && gate, "goat_door_gate_fwd: null argument");
  GOAT_CHECK(H >= 1 && H <= 32 * GATE_PER_LANE, "goat_door_gate_fwd: H=%d out of range", H);
  if (M <= 0) return GOAT_OK;
  door_gate_fwd_kernel<<<(M + GATE_WARPS - 1) / GATE_WARPS, GATE_WARPS * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      aug, ori, wa, ba, wo, bo, M, H, out, gate);
  GOAT_LAUNCH_CHECK();
  return GOAT_OK;
}

int goat_door_gate_bwd(const float* dout, const float* aug, const float* ori, const float* wa, const float* wo,
                       const float* gate, int M, int H, float* daug, float* dori, float* dwa, float* dwo, float* dba,
                       float* dbo, goat_stream_t stream) {
  GOAT_CHECK(dout && aug && ori && wa && wo && gate && daug && dori, "goat_door_gate_bwd: null argument");
  GOAT_CHECK(H >= 1 && H <= 32 * GATE_PER_LANE, "goat_door_gate_bwd: H=%d out of range", H);
  if (M <= 0) return GOAT_OK;
  int ctas = (M + 15) / 16;
  if (ctas > 592) ctas = 592;
  const int rows = (M + ctas - 1) / ctas;
  door_gate_bwd_kernel<<<ctas, GATE_WARPS * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      dout, aug, ori, wa, wo, gate, M, H, rows, daug, dori, dwa, dwo, dba, dbo);
  GOAT_LAUNCH_CHECK();
  return GOAT_OK;
}

int goat_xent_fwd(const float* logits, long long stride_row, long long stride_col, const long long* labels, int M, int N,
                  long long ignore_index, float* loss, float* lse, goat_stream_t stream) {
  GOAT_CHECK(logits && labels && loss && lse, "goat_xent_fwd: null argument");
  GOAT_CHECK(N >= 1, "goat_xent_fwd: N must be >= 1");
  if (M <= 0) return GOAT_OK;
  xent_fwd_kernel<<<M, XENT_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(logits, stride_row, stride_col, labels, N,
                                                                                  ignore_index, loss, lse);
  GOAT_LAUNCH_CHECK();
  return GOAT_OK;
}

int goat_xent_bwd(const float* dloss, const float* logits, long long stride_row, long long stride_col,
                  const long long* labels, const float* lse, int M, int N, long long ignore_index, float* dlogits,
                  long long dstride_row, long long dstride_col, int accumulate, goat_stream_t stream) {
  GOAT_CHECK(dloss && logits && labels && lse && dlogits, "goat_xent_bwd: null argument");
  if (M <= 0 || N <= 0) return GOAT_OK;
  int gy = (N + XENT_THREADS * 8 - 1) / (XENT_THREADS * 8);
  if (gy < 1) gy = 1;
  if (gy > 64) gy = 64;
  xent_bwd_kernel<<<dim3(M, gy), XENT_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      dloss, logits, stride_row, stride_col, labels, lse, N, ignore_index, dlogits, dstride_row, dstride_col, accumulate);
  GOAT_LAUNCH_CHECK();
  return GOAT_OK;
}

int goat_segment_reduce_fwd(const float* src, const int* idx, int R, int K, int H, int mean, float* out,
                            goat_stream_t stream) {
  GOAT_CHECK(src && idx && out, "goat_segment_reduce_fwd: null argument");
  if (R <= 0 || H <= 0) return GOAT_OK;
  GOAT_CHECK(R <= 65535, "goat_segment_reduce_fwd: too many rows");
  const int th = H >= 128 ? 128 : 32;
  segment_reduce_fwd_kernel<<<dim3((H + th - 1) / th, R), th, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, idx, K, H,
                                                                                                          mean, out);
  GOAT_LAUNCH_CHECK();
  return GOAT_OK;
}

int goat_segment_reduce_bwd(const float* dout, const int* idx, int R, int K, int H, int mean, float* dsrc,
                            goat_stream_t stream) {
  GOAT_CHECK(dout && idx && dsrc, "goat_segment_reduce_bwd: null argument");
  if (R <= 0 || H <= 0) return GOAT_OK;
  GOAT_CHECK(R <= 65535, "goat_segment_reduce_bwd: too many rows");
  const int th = H >= 128 ? 128 : 32;
  segment_reduce_bwd_kernel<<<dim3((H + th - 1) / th, R), th, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dout, idx, K, H,
                                                                                                          mean, dsrc);
  GOAT_LAUNCH_CHECK();
  return GOAT_OK;
}

int goat_embed_fwd(const long long* ids, const float* word, const float* pos, const float* type, int M, int L, int H,
                   float* out, goat_stream_t stream) {
  GOAT_CHECK(ids && word && pos && type && out, "goat_embed_fwd: null argument");
  GOAT_CHECK(L >= 1, "goat_embed_fwd: L must be >= 1");
  if (M <= 0 || H <= 0) return GOAT_OK;
  GOAT_CHECK(M <= 65535, "goat_embed_fwd: too many tokens per call (max 65535)");
  embed_fwd_kernel<<<dim3((H + 127) / 128, M), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(ids, word, pos, type, L, H,
                                                                                                out);
  GOAT_LAUNCH_CHECK();
  return GOAT_OK;
}

int goat_embed_bwd(const float* dout, const long long* ids, int M, int L, int H, long long padding_idx, float* dword,
                   float* dpos, float* dtype, goat_stream_t stream) {
  GOAT_CHECK(dout && ids, "goat_embed_bwd: null argument");
  if (M <= 0 || H <= 0) return GOAT_OK;
  GOAT_CHECK(M <= 65535, "goat_embed_bwd: too many tokens per call (max 65535)");
  embed_bwd_kernel<<<dim3((H + 127) / 128, M), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dout, ids, L, H,
                                                                                                padding_idx, dword, dpos, dtype);
  GOAT_LAUNCH_CHECK();
  return GOAT_OK;
}

int goat_act_grad(const float* dy, const void* ref, int ref_dtype, int act, void* out, int out_dtype, long long n,
                  goat_stream_t stream) {
  GOAT_CHECK(dy && out, "goat_act_grad: null argument");
  GOAT_CHECK(act == GOAT_ACT_NONE || ref, "goat_act_grad: the activation needs its reference tensor");
  GOAT_CHECK(act == GOAT_ACT_NONE || act == GOAT_ACT_RELU || act == GOAT_ACT_TANH || act == GOAT_ACT_GELU,
             "goat_act_grad: bad act %d", act);
  GOAT_CHECK(out_dtype == GOAT_F32 || out_dtype == GOAT_F16 || out_dtype == GOAT_BF16, "goat_act_grad: bad out dtype");
  if (n <= 0) return GOAT_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (act == GOAT_ACT_NONE || ref_dtype == GOAT_F32) return act_grad_launch<float>(dy, ref ? ref : dy, act, out, out_dtype, n, st);
  if (ref_dtype == GOAT_F16) return act_grad_launch<__half>(dy, ref, act, out, out_dtype, n, st);
  if (ref_dtype == GOAT_BF16) return act_grad_launch<__nv_bfloat16>(dy, ref, act, out, out_dtype, n, st);
  GOAT_CHECK(false, "goat_act_grad: bad ref dtype");
}

int goat_gather_rows(const void* src, const int* idx, int R, long long row_bytes, void* out, goat_stream_t stream) {
  GOAT_CHECK(src && idx && out, "goat_gather_rows: null argument");
  GOAT_CHECK(row_bytes > 0 && (row_bytes & 15) == 0 && aligned16(src) && aligned16(out),
             "goat_gather_rows: rows must be multiples of 16 bytes and 16-byte aligned");
  if (R <= 0) return GOAT_OK;
  GOAT_CHECK(R <= 65535, "goat_gather_rows: too many rows");
  const long long words = row_bytes / 16;
  int gx = (int)((words + 255) / 256);
  if (gx > 16) gx = 16;
  gather_rows_kernel<<<dim3(gx, R), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4*>(src), idx, words, reinterpret_cast<uint4*>(out));
  GOAT_LAUNCH_CHECK();
  return GOAT_OK;
}

int goat_act_fwd(const float* x, int act, float* out, long long n, goat_stream_t stream) {
  GOAT_CHECK(x && out, "goat_act_fwd: null argument");
  GOAT_CHECK(act == GOAT_ACT_NONE || act == GOAT_ACT_RELU || act == GOAT_ACT_TANH || act == GOAT_ACT_GELU,
             "goat_act_fwd: bad act %d", act);
  if (n <= 0) return GOAT_OK;
  long long want = (n + 255) / 256;
  act_fwd_kernel<<<(int)(want > 148 * 8 ? 148 * 8 : want), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, act, out, n);
  GOAT_LAUNCH_CHECK();
  return GOAT_OK;
}

int goat_sprel_fwd(const float* d, const float* w, const float* b, float* out, long long n, goat_stream_t stream) {
  GOAT_CHECK(d && w && b && out, "goat_sprel_fwd: null argument");
  if (n <= 0) return GOAT_OK;
  long long want = (n + 255) / 256;
  sprel_fwd_kernel<<<(int)(want > 148 * 8 ? 148 * 8 : want), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d, w, b, out, n);
  GOAT_LAUNCH_CHECK();
  return GOAT_OK;
}

int goat_sprel_bwd(const float* dout, const float* d, float* dw, float* db, long long n, goat_stream_t stream) {
  GOAT_CHECK(dout && d && dw && db, "goat_sprel_bwd: null argument");
  if (n <= 0) return GOAT_OK;
  long long want = (n + 256 * 8 - 1) / (256 * 8);
  sprel_bwd_kernel<<<(int)(want > 148 ? 148 : want), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dout, d, dw, db, n);
  GOAT_LAUNCH_CHECK();
  return GOAT_OK;
}

int goat_xent_chunk_fwd(const float* logits, long long ld, const long long* labels, int M, int Nc, long long c0, int first,
                        float* run_max, float* run_sum, float* picked, goat_stream_t stream) {
  GOAT_CHECK(logits && labels && run_max && run_sum && picked, "goat_xent_chunk_fwd: null argument");
  GOAT_CHECK(Nc >= 1 && ld >= Nc, "goat_xent_chunk_fwd: bad chunk width / leading dimension");
  if (M <= 0) return GOAT_OK;
  xent_chunk_fwd_kernel<<<M, XENT_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(logits, ld, labels, Nc, c0, first,
                                                                                        run_max, run_sum, picked);
  GOAT_LAUNCH_CHECK();
  return GOAT_OK;
}

int goat_xent_chunk_bwd(const float* dloss, const float* logits, long long ld, const long long* labels, const float* lse,
                        int M, int Nc, long long c0, long long ignore_index, long long n_classes, void* out, int out_dtype,
                        long long ldo, goat_stream_t stream) {
  GOAT_CHECK(dloss && logits && labels && lse && out, "goat_xent_chunk_bwd: null argument");
  GOAT_CHECK(out_dtype == GOAT_F32 || out_dtype == GOAT_F16 || out_dtype == GOAT_BF16, "goat_xent_chunk_bwd: bad out dtype");
  GOAT_CHECK(Nc >= 1 && ld >= Nc && ldo >= Nc, "goat_xent_chunk_bwd: bad chunk width / leading dimension");
  if (M <= 0) return GOAT_OK;
  int gy = (Nc + XENT_THREADS * 8 - 1) / (XENT_THREADS * 8);
  if (gy < 1) gy = 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid(M, gy);
  if (out_dtype == GOAT_F32)
    xent_chunk_bwd_kernel<float><<<grid, XENT_THREADS, 0, st>>>(dloss, logits, ld, labels, lse, Nc, c0, ignore_index, n_classes,
                                                                 (float*)out, ldo);
  else if (out_dtype == GOAT_F16)
    xent_chunk_bwd_kernel<__half><<<grid, XENT_THREADS, 0, st>>>(dloss, logits, ld, labels, lse, Nc, c0, ignore_index, n_classes,
                                                                  (__half*)out, ldo);
  else
    xent_chunk_bwd_kernel<__nv_bfloat16><<<grid, XENT_THREADS, 0, st>>>(dloss, logits, ld, labels, lse, Nc, c0, ignore_index,
                                                                         n_classes, (__nv_bfloat16*)out, ldo);
  GOAT_LAUNCH_CHECK();
  return GOAT_OK;
}

}  // extern "C"
