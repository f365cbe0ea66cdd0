// Fused optimizer step over one flat parameter buffer: global-norm clip + AdamW + 16-bit operand shadow.
//
// Replaces the per-tensor Python loop of the reference (P/optim/adamw.py:64-110: ~8 tiny launches per
// parameter tensor) and torch.nn.utils.clip_grad_norm_ (P/train_r2r_goat.py:349-366) with two launches:
//   goat_sumsq      partial[g] = sum of squares of a grid-strided slice of the flat gradient
//   goat_adamw_step every CTA reduces partial[] in a fixed order (deterministic) -> clip coefficient,
//                   then updates p, m, v in place and writes the bf16/fp16 copy the GEMMs read next step.
// HBM-bound: 16 B read + 18 B written per parameter (p, m, v, the 16-bit shadow and the cleared gradient).  Hyper-parameters live in a small DEVICE array so a
// captured CUDA graph can be replayed with a new learning rate / bias correction each step.
#include "common.cuh"

namespace goat {
namespace {

constexpr int OPT_THREADS = 256;
constexpr int SUMSQ_MAX_PARTS = 1184;  // 148 SMs x 8

__global__ void __launch_bounds__(OPT_THREADS) sumsq_kernel(const float* __restrict__ g, long long n,
                                                            float* __restrict__ partial) {
  __shared__ float red[OPT_THREADS / 32];
  float s = 0.f;
  const long long n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = g4[i];
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0)
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) s += g[i] * g[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < OPT_THREADS / 32; ++w) t += red[w];
    partial[blockIdx.x] = t;
  }
}

// hp (device, fp32): [0] lr  [1] beta1  [2] beta2  [3] eps  [4] weight_decay  [5] 1-beta1^t  [6] 1-beta2^t
//                    [7] max_grad_norm (<= 0: no clipping)  [8] gradient pre-scale (1/world for summed grads)
// scaler (device, fp32, optional): the dynamic loss scale of the fp16 path (torch.cuda.amp.GradScaler semantics,
//   P/train_r2r_goat.py:279,325,351-363): [0] scale  [1] clean steps since the last change  [2] 1 if the last step
//   overflowed  [3] skipped steps so far  [4] optimizer steps actually taken.  With a scaler the gradients are divided
//   by scaler[0] on top of hp[8], a non-finite gradient norm SKIPS the update (p, m, v, shadow untouched; the gradient
//   buffer is still cleared), and the bias corrections are computed here from scaler[4] + 1 instead of hp[5], hp[6]
//   (a skipped step does not advance Adam's step count).  goat_scaler_update() then moves the scale.
template <typename TS>
__global__ void __launch_bounds__(OPT_THREADS)
adamw_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             TS* __restrict__ shadow, long long n, long long n_decay, const float* __restrict__ hp,
             const float* __restrict__ partial, int nparts, float* __restrict__ norm_out, int zero_grad,
             const float* __restrict__ scaler, TS* __restrict__ shadow_lo) {
  __shared__ float s_coef;
  __shared__ int s_skip;
  if (threadIdx.x < 32) {
    float t = 0.f;
    for (int i = threadIdx.x; i < nparts; i += 32) t += partial[i];
    t = warp_sum(t);
    if (threadIdx.x == 0) {
      const float pre = scaler ? hp[8] / scaler[0] : hp[8];
      const float norm = sqrtf(t) * pre;
      float coef = pre;
      if (hp[7] > 0.f) {
        const float c = hp[7] / (norm + 1e-6f);  // torch.nn.utils.clip_grad_norm_
        if (c < 1.f) coef *= c;
      }
      s_coef = coef;
      s_skip = (scaler && !isfinite(t)) ? 1 : 0;
      if (blockIdx.x == 0 && norm_out) *norm_out = norm;
    }
  }
  __syncthreads();
  const float coef = s_coef;
  const long long n4 = n >> 2;
  if (s_skip) {   // overflowed fp16 gradients: drop them, change nothing else
    if (zero_grad) {
      for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
        reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (blockIdx.x == 0)
        for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) g[i] = 0.f;
    }
    return;
  }
  const float lr = hp[0], b1 = hp[1], b2 = hp[2], eps = hp[3], wd = hp[4];
  float bc1 = hp[5], bc2 = hp[6];
  if (scaler && hp[5] != 1.0f) {   // hp[5] == 1 exactly: correct_bias off
    const float t = scaler[4] + 1.0f;
    bc1 = 1.0f - powf(b1, t);
    bc2 = 1.0f - powf(b2, t);
  }
  const float step_size = lr * sqrtf(bc2) / bc1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* pa = &pp.x; const float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gr = ga[j] * coef;
      ma[j] = ma[j] * b1 + gr * (1.f - b1);
      va[j] = va[j] * b2 + gr * gr * (1.f - b2);
      float x = pa[j] - step_size * (ma[j] / (sqrtf(va[j]) + eps));
      if (i * 4 + j < n_decay) x -= x * (lr * wd);
      pa[j] = x;
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (shadow) {
      if constexpr (sizeof(TS) == 2) {
        uint2 w;
        w.x = pack2<TS>(pp.x, pp.y);
        w.y = pack2<TS>(pp.z, pp.w);
        reinterpret_cast<uint2*>(shadow)[i] = w;
        if (shadow_lo) {   // second term of the split weight: what the 16-bit copy lost
          const float2 h0 = unpack2<TS>(w.x), h1 = unpack2<TS>(w.y);
          uint2 l;
          l.x = pack2<TS>(pp.x - h0.x, pp.y - h0.y);
          l.y = pack2<TS>(pp.z - h1.x, pp.w - h1.y);
          reinterpret_cast<uint2*>(shadow_lo)[i] = l;
        }
      }
    }
  }
  if (blockIdx.x == 0) {
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
      const float gr = g[i] * coef;
      const float mi = m[i] * b1 + gr * (1.f - b1);
      const float vi = v[i] * b2 + gr * gr * (1.f - b2);
      float x = p[i] - step_size * (mi / (sqrtf(vi) + eps));
      if (i < n_decay) x -= x * (lr * wd);
      p[i] = x; m[i] = mi; v[i] = vi;
      if (zero_grad) g[i] = 0.f;
      if (shadow) {
        if constexpr (sizeof(TS) == 2) {
          shadow[i] = from_f<TS>(x);
          if (shadow_lo) shadow_lo[i] = from_f<TS>(x - to_f<TS>(shadow[i]));
        }
      }
    }
  }
}

// one thread: GradScaler.update() -- halve the scale after an overflow, double it after `interval` clean steps
__global__ void scaler_update_kernel(float* __restrict__ scaler, const float* __restrict__ partial, int nparts, float growth,
                                     float backoff, float interval) {
  float t = 0.f;
  for (int i = 0; i < nparts; ++i) t += partial[i];
  if (!isfinite(t)) {
    scaler[0] = fmaxf(scaler[0] * backoff, 1.0f);
    scaler[1] = 0.f;
    scaler[2] = 1.f;
    scaler[3] += 1.f;
  } else {
    scaler[2] = 0.f;
    scaler[4] += 1.f;
    scaler[1] += 1.f;
    if (scaler[1] >= interval) {
      scaler[0] = fminf(scaler[0] * growth, 16777216.0f);
      scaler[1] = 0.f;
    }
  }
}

}  // namespace
}  // namespace goat

using namespace goat;

extern "C" size_t goat_sumsq_workspace_bytes(void) { return SUMSQ_MAX_PARTS * sizeof(float); }

extern "C" int goat_sumsq(const float* g, long long n, float* partial, int* nparts_out, goat_stream_t stream) {
  GOAT_CHECK(g && partial && nparts_out, "goat_sumsq: null argument");
  GOAT_CHECK(aligned16(g), "goat_sumsq: gradient buffer must be 16-byte aligned");
  long long want = (n / 4 + OPT_THREADS - 1) / OPT_THREADS;
  int parts = (int)(want < 1 ? 1 : (want > SUMSQ_MAX_PARTS ? SUMSQ_MAX_PARTS : want));
  *nparts_out = parts;
  sumsq_kernel<<<parts, OPT_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(g, n, partial);
  GOAT_LAUNCH_CHECK();
  return GOAT_OK;
}

extern "C" int goat_adamw_step(float* p, float* g, float* m, float* v, void* shadow, int shadow_dtype, long long n,
                               long long n_decay, const float* hp, const float* partial, int nparts, float* norm_out,
                               int zero_grad, const float* scaler, void* shadow_lo, goat_stream_t stream) {
  GOAT_CHECK(p && g && m && v && hp && partial, "goat_adamw_step: null argument");
  GOAT_CHECK(aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v) && (!shadow || aligned16(shadow)),
             "goat_adamw_step: buffers must be 16-byte aligned");
  GOAT_CHECK(!shadow || shadow_dtype == GOAT_F16 || shadow_dtype == GOAT_BF16, "goat_adamw_step: shadow dtype must be F16/BF16");
  GOAT_CHECK(!shadow_lo || (shadow && aligned16(shadow_lo)), "goat_adamw_step: shadow_lo needs shadow and 16-byte alignment");
  GOAT_CHECK(nparts >= 1 && nparts <= SUMSQ_MAX_PARTS, "goat_adamw_step: bad nparts");
  if (n <= 0) return GOAT_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  long long want = (n / 4 + OPT_THREADS - 1) / OPT_THREADS;
  const int grid = (int)(want < 1 ? 1 : (want > 148 * 16 ? 148 * 16 : want));
  if (shadow && shadow_dtype == GOAT_F16)
    adamw_kernel<__half><<<grid, OPT_THREADS, 0, st>>>(p, g, m, v, (__half*)shadow, n, n_decay, hp, partial, nparts, norm_out, zero_grad, scaler, (__half*)shadow_lo);
  else if (shadow)
    adamw_kernel<__nv_bfloat16><<<grid, OPT_THREADS, 0, st>>>(p, g, m, v, (__nv_bfloat16*)shadow, n, n_decay, hp, partial, nparts, norm_out, zero_grad, scaler, (__nv_bfloat16*)shadow_lo);
  else
    adamw_kernel<float><<<grid, OPT_THREADS, 0, st>>>(p, g, m, v, (float*)nullptr, n, n_decay, hp, partial, nparts, norm_out, zero_grad, scaler, (float*)nullptr);
  GOAT_LAUNCH_CHECK();
  return GOAT_OK;
}

extern "C" int goat_scaler_update(float* scaler, const float* partial, int nparts, float growth, float backoff,
                                  int interval, goat_stream_t stream) {
  GOAT_CHECK(scaler && partial, "goat_scaler_update: null argument");
  GOAT_CHECK(nparts >= 1 && nparts <= SUMSQ_MAX_PARTS, "goat_scaler_update: bad nparts");
  GOAT_CHECK(growth >= 1.f && backoff > 0.f && backoff <= 1.f && interval >= 1, "goat_scaler_update: bad schedule");
  scaler_update_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream)>>>(scaler, partial, nparts, growth, backoff,
                                                                            (float)interval);
  GOAT_LAUNCH_CHECK();
  return GOAT_OK;
}
