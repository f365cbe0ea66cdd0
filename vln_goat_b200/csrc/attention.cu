// Attention core, fp32-math SIMT version (all dtypes): O = softmax(scale QK^T + kmask + bias) V.
//
// Token-major operands (element (b,t,h,d) at base[b*sb + t*ld + h*64 + d]) so the fused QKV projection
// output is consumed in place and O is written head-merged.  Flash-style: keys are streamed through
// shared memory in chunks of 64 with an online softmax, the [Nq,Nk] score matrix never touches HBM.
// Forward saves lse[b,h,q]; backward recomputes P from it (two kernels: dQ(+dbias) and dK/dV, no atomics
// except the optional head-summed dbias).
// This kernel is HBM / latency bound (~25 FLOP/B), grid = q-tiles x heads x batch.
#include "common.cuh"

namespace goat {

namespace {

constexpr int D = 64;
constexpr int QT = 32;   // query rows per CTA (8 per warp)
constexpr int KC = 64;   // keys per chunk
constexpr int RW = 8;    // rows per warp
constexpr int PAD = 65;  // padded row stride for lane-indexed rows

struct AttnP {
  goat_attn_args a;
};

template <typename T>
__device__ __forceinline__ float ld(const void* base, long long off) {
  return to_f<T>(reinterpret_cast<const T*>(base)[off]);
}

// cooperative load of `rows` x 64 elements (row r <- token t0 + r, zero past ntok) into smem with `stride`
template <typename T>
__device__ __forceinline__ void load_tile(float* dst, int stride, const void* base, long long boff, int ld_, int h,
                                          int t0, int ntok, int rows) {
  for (int e = threadIdx.x; e < rows * (D / 2); e += blockDim.x) {
    const int r = e / (D / 2), c = (e % (D / 2)) * 2;
    const int t = t0 + r;
    float x = 0.f, y = 0.f;
    if (t < ntok) {
      const long long off = boff + (long long)t * ld_ + h * D + c;
      if constexpr (sizeof(T) == 2) {
        const uint32_t w = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const T*>(base) + off);
        const float2 f = unpack2<T>(w);
        x = f.x; y = f.y;
      } else {
        const float2 f = *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(base) + off);
        x = f.x; y = f.y;
      }
    }
    dst[r * stride + c] = x;
    dst[r * stride + c + 1] = y;
  }
}

__device__ __forceinline__ float score_bias(const goat_attn_args& a, int b, int qi, int kj) {
  float s = 0.f;
  if (a.kmask) s += a.kmask[(long long)b * a.Nk + kj];
  if (a.bias) s += a.bias[((long long)b * a.Nq + qi) * a.Nk + kj];
  return s;
}

__device__ __forceinline__ float drop_scale(const goat_attn_args& a, int b, int h, int qi, int kj) {
  if (a.drop_p <= 0.f) return 1.f;
  const uint32_t rowkey = attn_drop_rowkey(eff_seed(a.drop_seed, reinterpret_cast<const unsigned long long*>(a.drop_seed_ptr)),
                                           b, a.heads, h, a.Nq, qi);
  return attn_drop_keep(rowkey, kj, drop_thr16(a.drop_p)) ? 1.f / (1.f - a.drop_p) : 0.f;
}

// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) attn_fwd_kernel(goat_attn_args a) {
  extern __shared__ float sm[];
  float* Qs = sm;                  // [QT][D]
  float* Ks = Qs + QT * D;         // [KC][PAD]
  float* Vs = Ks + KC * PAD;       // [KC][D]
  float* Ps = Vs + KC * D;         // [4][RW][KC]
  const int q0 = blockIdx.x * QT, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  load_tile<T>(Qs, D, a.Q, (long long)b * a.sbq, a.ldq, h, q0, a.Nq, QT);

  float m[RW], l[RW], o0[RW], o1[RW];
#pragma unroll
  for (int r = 0; r < RW; ++r) { m[r] = -INFINITY; l[r] = 0.f; o0[r] = 0.f; o1[r] = 0.f; }

  for (int kc = 0; kc < a.Nk; kc += KC) {
    __syncthreads();
    load_tile<T>(Ks, PAD, a.K, (long long)b * a.sbk, a.ldk, h, kc, a.Nk, KC);
    load_tile<T>(Vs, D, a.V, (long long)b * a.sbv, a.ldv, h, kc, a.Nk, KC);
    __syncthreads();
    float s0[RW], s1[RW];
#pragma unroll
    for (int r = 0; r < RW; ++r) { s0[r] = 0.f; s1[r] = 0.f; }
#pragma unroll 8
    for (int d = 0; d < D; ++d) {
      const float k0 = Ks[lane * PAD + d], k1 = Ks[(lane + 32) * PAD + d];
#pragma unroll
      for (int r = 0; r < RW; ++r) {
        const float q = Qs[(warp * RW + r) * D + d];
        s0[r] = fmaf(q, k0, s0[r]);
        s1[r] = fmaf(q, k1, s1[r]);
      }
    }
    const int j0 = kc + lane, j1 = kc + lane + 32;
#pragma unroll
    for (int r = 0; r < RW; ++r) {
      const int qi = q0 + warp * RW + r;
      const bool qok = qi < a.Nq;
      float x0 = (qok && j0 < a.Nk) ? s0[r] * a.scale + score_bias(a, b, qi, j0) : -INFINITY;
      float x1 = (qok && j1 < a.Nk) ? s1[r] * a.scale + score_bias(a, b, qi, j1) : -INFINITY;
      const float mx = warp_max(fmaxf(x0, x1));
      const float mnew = fmaxf(m[r], mx);
      const float mref = (mnew == -INFINITY) ? 0.f : mnew;
      const float p0 = expf(x0 - mref), p1 = expf(x1 - mref);
      const float corr = expf(m[r] - mref);
      l[r] = l[r] * corr + warp_sum(p0 + p1);
      o0[r] *= corr; o1[r] *= corr;
      m[r] = mnew;
      float* prow = Ps + (warp * RW + r) * KC;
      prow[lane] = (qok && j0 < a.Nk) ? p0 * drop_scale(a, b, h, qi, j0) : 0.f;
      prow[lane + 32] = (qok && j1 < a.Nk) ? p1 * drop_scale(a, b, h, qi, j1) : 0.f;
    }
    __syncwarp();
#pragma unroll 4
    for (int j = 0; j < KC; ++j) {
      const float v0 = Vs[j * D + lane], v1 = Vs[j * D + lane + 32];
#pragma unroll
      for (int r = 0; r < RW; ++r) {
        const float p = Ps[(warp * RW + r) * KC + j];
        o0[r] = fmaf(p, v0, o0[r]);
        o1[r] = fmaf(p, v1, o1[r]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < RW; ++r) {
    const int qi = q0 + warp * RW + r;
    if (qi >= a.Nq) continue;
    const float inv = l[r] > 0.f ? 1.f / l[r] : 0.f;
    const long long off = (long long)b * a.sbo + (long long)qi * a.ldo + h * D;
    reinterpret_cast<T*>(a.O)[off + lane] = from_f<T>(o0[r] * inv);
    reinterpret_cast<T*>(a.O)[off + lane + 32] = from_f<T>(o1[r] * inv);
    if (lane == 0 && a.lse)
      a.lse[((long long)b * a.heads + h) * a.Nq + qi] = (l[r] > 0.f) ? m[r] + logf(l[r]) : -INFINITY;
  }
}

// ---------------------------------------------------------------------------------------------
// dQ (+ dbias): same tiling as forward.  dS = P * (dP - delta), dQ = scale * dS K
template <typename T>
__global__ void __launch_bounds__(128) attn_bwd_dq_kernel(goat_attn_args a) {
  extern __shared__ float sm[];
  float* Qs = sm;                  // [QT][D]
  float* dOs = Qs + QT * D;        // [QT][D]
  float* Ks = dOs + QT * D;        // [KC][PAD]
  float* Vs = Ks + KC * PAD;       // [KC][PAD]
  float* Ss = Vs + KC * PAD;       // [4][RW][KC]
  const int q0 = blockIdx.x * QT, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  load_tile<T>(Qs, D, a.Q, (long long)b * a.sbq, a.ldq, h, q0, a.Nq, QT);
  load_tile<T>(dOs, D, a.dO, (long long)b * a.sbo, a.ldo, h, q0, a.Nq, QT);
  __syncthreads();

  float delta[RW], lse[RW], dq0[RW], dq1[RW];
#pragma unroll
  for (int r = 0; r < RW; ++r) {
    const int qi = q0 + warp * RW + r;
    float t = 0.f;
    lse[r] = 0.f;
    if (qi < a.Nq) {
      const long long off = (long long)b * a.sbo + (long long)qi * a.ldo + h * D;
      t = dOs[(warp * RW + r) * D + lane] * ld<T>(a.O, off + lane) +
          dOs[(warp * RW + r) * D + lane + 32] * ld<T>(a.O, off + lane + 32);
      lse[r] = a.lse[((long long)b * a.heads + h) * a.Nq + qi];
    }
    delta[r] = warp_sum(t);
    dq0[r] = 0.f; dq1[r] = 0.f;
  }

  for (int kc = 0; kc < a.Nk; kc += KC) {
    __syncthreads();
    load_tile<T>(Ks, PAD, a.K, (long long)b * a.sbk, a.ldk, h, kc, a.Nk, KC);
    load_tile<T>(Vs, PAD, a.V, (long long)b * a.sbv, a.ldv, h, kc, a.Nk, KC);
    __syncthreads();
    float s0[RW], s1[RW], p0[RW], p1[RW];
#pragma unroll
    for (int r = 0; r < RW; ++r) { s0[r] = 0.f; s1[r] = 0.f; p0[r] = 0.f; p1[r] = 0.f; }
#pragma unroll 4
    for (int d = 0; d < D; ++d) {
      const float k0 = Ks[lane * PAD + d], k1 = Ks[(lane + 32) * PAD + d];
      const float v0 = Vs[lane * PAD + d], v1 = Vs[(lane + 32) * PAD + d];
#pragma unroll
      for (int r = 0; r < RW; ++r) {
        const float q = Qs[(warp * RW + r) * D + d];
        const float g = dOs[(warp * RW + r) * D + d];
        s0[r] = fmaf(q, k0, s0[r]); s1[r] = fmaf(q, k1, s1[r]);
        p0[r] = fmaf(g, v0, p0[r]); p1[r] = fmaf(g, v1, p1[r]);   // dP~ = dO V^T
      }
    }
    const int j0 = kc + lane, j1 = kc + lane + 32;
#pragma unroll
    for (int r = 0; r < RW; ++r) {
      const int qi = q0 + warp * RW + r;
      const bool qok = qi < a.Nq;
      float ds0 = 0.f, ds1 = 0.f;
      if (qok && j0 < a.Nk) {
        const float p = expf(s0[r] * a.scale + score_bias(a, b, qi, j0) - lse[r]);
        ds0 = p * (p0[r] * drop_scale(a, b, h, qi, j0) - delta[r]);
        if (a.dbias) atomicAdd(a.dbias + ((long long)b * a.Nq + qi) * a.Nk + j0, ds0);
      }
      if (qok && j1 < a.Nk) {
        const float p = expf(s1[r] * a.scale + score_bias(a, b, qi, j1) - lse[r]);
        ds1 = p * (p1[r] * drop_scale(a, b, h, qi, j1) - delta[r]);
        if (a.dbias) atomicAdd(a.dbias + ((long long)b * a.Nq + qi) * a.Nk + j1, ds1);
      }
      float* srow = Ss + (warp * RW + r) * KC;
      srow[lane] = ds0;
      srow[lane + 32] = ds1;
    }
    __syncwarp();
#pragma unroll 4
    for (int j = 0; j < KC; ++j) {
      const float k0 = Ks[j * PAD + lane], k1 = Ks[j * PAD + lane + 32];
#pragma unroll
      for (int r = 0; r < RW; ++r) {
        const float s = Ss[(warp * RW + r) * KC + j];
        dq0[r] = fmaf(s, k0, dq0[r]);
        dq1[r] = fmaf(s, k1, dq1[r]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < RW; ++r) {
    const int qi = q0 + warp * RW + r;
    if (qi >= a.Nq) continue;
    const long long off = (long long)b * a.sbq + (long long)qi * a.ldq + h * D;
    reinterpret_cast<T*>(a.dQ)[off + lane] = from_f<T>(dq0[r] * a.scale);
    reinterpret_cast<T*>(a.dQ)[off + lane + 32] = from_f<T>(dq1[r] * a.scale);
  }
}

// ---------------------------------------------------------------------------------------------
// dK, dV: each CTA owns 32 keys (8 per warp) and streams the queries in chunks of 64.
template <typename T>
__global__ void __launch_bounds__(128) attn_bwd_dkv_kernel(goat_attn_args a) {
  extern __shared__ float sm[];
  float* Ks = sm;                   // [QT][D]   (QT keys here)
  float* Vs = Ks + QT * D;          // [QT][D]
  float* Qs = Vs + QT * D;          // [KC][PAD] (KC queries per chunk)
  float* dOs = Qs + KC * PAD;       // [KC][PAD]
  float* Ps = dOs + KC * PAD;       // [4][RW][KC]
  float* Ss = Ps + 4 * RW * KC;     // [4][RW][KC]
  float* lses = Ss + 4 * RW * KC;   // [KC]
  float* dels = lses + KC;          // [KC]
  const int k0 = blockIdx.x * QT, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  load_tile<T>(Ks, D, a.K, (long long)b * a.sbk, a.ldk, h, k0, a.Nk, QT);
  load_tile<T>(Vs, D, a.V, (long long)b * a.sbv, a.ldv, h, k0, a.Nk, QT);

  float dk0[RW], dk1[RW], dv0[RW], dv1[RW];
#pragma unroll
  for (int r = 0; r < RW; ++r) { dk0[r] = dk1[r] = dv0[r] = dv1[r] = 0.f; }

  for (int qc = 0; qc < a.Nq; qc += KC) {
    __syncthreads();
    load_tile<T>(Qs, PAD, a.Q, (long long)b * a.sbq, a.ldq, h, qc, a.Nq, KC);
    load_tile<T>(dOs, PAD, a.dO, (long long)b * a.sbo, a.ldo, h, qc, a.Nq, KC);
    __syncthreads();
    // delta / lse for the 64 queries of this chunk: 16 per warp
    for (int i = warp * 16; i < warp * 16 + 16; ++i) {
      const int qi = qc + i;
      float t = 0.f;
      if (qi < a.Nq) {
        const long long off = (long long)b * a.sbo + (long long)qi * a.ldo + h * D;
        t = dOs[i * PAD + lane] * ld<T>(a.O, off + lane) + dOs[i * PAD + lane + 32] * ld<T>(a.O, off + lane + 32);
      }
      t = warp_sum(t);
      if (lane == 0) {
        dels[i] = t;
        lses[i] = (qi < a.Nq) ? a.lse[((long long)b * a.heads + h) * a.Nq + qi] : 0.f;
      }
    }
    __syncthreads();
    float s0[RW], s1[RW], p0[RW], p1[RW];
#pragma unroll
    for (int r = 0; r < RW; ++r) { s0[r] = s1[r] = p0[r] = p1[r] = 0.f; }
#pragma unroll 4
    for (int d = 0; d < D; ++d) {
      const float q0 = Qs[lane * PAD + d], q1 = Qs[(lane + 32) * PAD + d];
      const float g0 = dOs[lane * PAD + d], g1 = dOs[(lane + 32) * PAD + d];
#pragma unroll
      for (int r = 0; r < RW; ++r) {
        const float k = Ks[(warp * RW + r) * D + d];
        const float v = Vs[(warp * RW + r) * D + d];
        s0[r] = fmaf(q0, k, s0[r]); s1[r] = fmaf(q1, k, s1[r]);
        p0[r] = fmaf(g0, v, p0[r]); p1[r] = fmaf(g1, v, p1[r]);
      }
    }
    const int i0 = qc + lane, i1 = qc + lane + 32;
#pragma unroll
    for (int r = 0; r < RW; ++r) {
      const int kj = k0 + warp * RW + r;
      const bool kok = kj < a.Nk;
      float pt0 = 0.f, pt1 = 0.f, ds0 = 0.f, ds1 = 0.f;
      if (kok && i0 < a.Nq) {
        const float p = expf(s0[r] * a.scale + score_bias(a, b, i0, kj) - lses[lane]);
        const float dsc = drop_scale(a, b, h, i0, kj);
        pt0 = p * dsc;
        ds0 = p * (p0[r] * dsc - dels[lane]);
      }
      if (kok && i1 < a.Nq) {
        const float p = expf(s1[r] * a.scale + score_bias(a, b, i1, kj) - lses[lane + 32]);
        const float dsc = drop_scale(a, b, h, i1, kj);
        pt1 = p * dsc;
        ds1 = p * (p1[r] * dsc - dels[lane + 32]);
      }
      Ps[(warp * RW + r) * KC + lane] = pt0;
      Ps[(warp * RW + r) * KC + lane + 32] = pt1;
      Ss[(warp * RW + r) * KC + lane] = ds0;
      Ss[(warp * RW + r) * KC + lane + 32] = ds1;
    }
    __syncwarp();
#pragma unroll 4
    for (int i = 0; i < KC; ++i) {
      const float g0 = dOs[i * PAD + lane], g1 = dOs[i * PAD + lane + 32];
      const float q0 = Qs[i * PAD + lane], q1 = Qs[i * PAD + lane + 32];
#pragma unroll
      for (int r = 0; r < RW; ++r) {
        const float p = Ps[(warp * RW + r) * KC + i];
        const float s = Ss[(warp * RW + r) * KC + i];
        dv0[r] = fmaf(p, g0, dv0[r]); dv1[r] = fmaf(p, g1, dv1[r]);
        dk0[r] = fmaf(s, q0, dk0[r]); dk1[r] = fmaf(s, q1, dk1[r]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < RW; ++r) {
    const int kj = k0 + warp * RW + r;
    if (kj >= a.Nk) continue;
    const long long offk = (long long)b * a.sbk + (long long)kj * a.ldk + h * D;
    const long long offv = (long long)b * a.sbv + (long long)kj * a.ldv + h * D;
    reinterpret_cast<T*>(a.dK)[offk + lane] = from_f<T>(dk0[r] * a.scale);
    reinterpret_cast<T*>(a.dK)[offk + lane + 32] = from_f<T>(dk1[r] * a.scale);
    reinterpret_cast<T*>(a.dV)[offv + lane] = from_f<T>(dv0[r]);
    reinterpret_cast<T*>(a.dV)[offv + lane + 32] = from_f<T>(dv1[r]);
  }
}

constexpr int FWD_SMEM = (QT * D + KC * PAD + KC * D + 4 * RW * KC) * 4;
constexpr int DQ_SMEM = (2 * QT * D + 2 * KC * PAD + 4 * RW * KC) * 4;
constexpr int DKV_SMEM = (2 * QT * D + 2 * KC * PAD + 2 * 4 * RW * KC + 2 * KC) * 4;

template <typename T>
int fwd_t(const goat_attn_args& a, cudaStream_t st) {
  static bool cfg = false;
  if (!cfg) {
    GOAT_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
    cfg = true;
  }
  dim3 grid((a.Nq + QT - 1) / QT, a.heads, a.B);
  attn_fwd_kernel<T><<<grid, 128, FWD_SMEM, st>>>(a);
  GOAT_LAUNCH_CHECK();
  return GOAT_OK;
}

template <typename T>
int bwd_t(const goat_attn_args& a, cudaStream_t st) {
  static bool cfg = false;
  if (!cfg) {
    GOAT_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, DQ_SMEM));
    GOAT_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, DKV_SMEM));
    cfg = true;
  }
  dim3 gq((a.Nq + QT - 1) / QT, a.heads, a.B);
  attn_bwd_dq_kernel<T><<<gq, 128, DQ_SMEM, st>>>(a);
  GOAT_LAUNCH_CHECK();
  dim3 gk((a.Nk + QT - 1) / QT, a.heads, a.B);
  attn_bwd_dkv_kernel<T><<<gk, 128, DKV_SMEM, st>>>(a);
  GOAT_LAUNCH_CHECK();
  return GOAT_OK;
}

int check_args(const goat_attn_args* a, bool bwd) {
  GOAT_CHECK(a != nullptr, "goat_attn: null args");
  GOAT_CHECK(a->D == 64, "goat_attn: head dim must be 64 (got %d)", a->D);
  GOAT_CHECK(a->B >= 0 && a->heads > 0 && a->Nq >= 0 && a->Nk >= 0, "goat_attn: bad sizes");
  GOAT_CHECK(a->dtype == GOAT_F32 || a->dtype == GOAT_F16 || a->dtype == GOAT_BF16, "goat_attn: bad dtype");
  GOAT_CHECK(a->Q && a->K && a->V && a->O, "goat_attn: null Q/K/V/O");
  GOAT_CHECK(a->ldq >= a->heads * 64 && a->ldk >= a->heads * 64 && a->ldv >= a->heads * 64 && a->ldo >= a->heads * 64,
             "goat_attn: row stride smaller than heads*64");
  GOAT_CHECK((a->ldq % 2) == 0 && (a->ldk % 2) == 0 && (a->ldv % 2) == 0 && (a->ldo % 2) == 0 && (a->sbq % 2) == 0 &&
                 (a->sbk % 2) == 0 && (a->sbv % 2) == 0 && (a->sbo % 2) == 0,
             "goat_attn: strides must be even");
  GOAT_CHECK(a->drop_p >= 0.f && a->drop_p < 1.f, "goat_attn: drop_p out of range");
  GOAT_CHECK(a->heads <= 65535 && a->B <= 65535, "goat_attn: heads/batch exceed grid limits");
  if (bwd) GOAT_CHECK(a->dO && a->dQ && a->dK && a->dV && a->lse, "goat_attn_core_bwd: null dO/dQ/dK/dV/lse");
  return GOAT_OK;
}

}  // namespace

bool attn_tc_eligible(const goat_attn_args* a, bool bwd);
int attn_fwd_tc(const goat_attn_args* a, cudaStream_t st);
int attn_bwd_tc(const goat_attn_args* a, cudaStream_t st);

}  // namespace goat

using namespace goat;

extern "C" int goat_attn_core_fwd(const goat_attn_args* a, goat_stream_t stream) {
  int rc = check_args(a, false);
  if (rc) return rc;
  if (a->B == 0 || a->Nq == 0) return GOAT_OK;
  GOAT_CHECK(a->Nk > 0, "goat_attn_core_fwd: Nk must be > 0");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!a->force_simt && attn_tc_eligible(a, false)) return attn_fwd_tc(a, st);
  if (a->dtype == GOAT_F32) return fwd_t<float>(*a, st);
  if (a->dtype == GOAT_F16) return fwd_t<__half>(*a, st);
  return fwd_t<__nv_bfloat16>(*a, st);
}

extern "C" int goat_attn_core_bwd(const goat_attn_args* a, goat_stream_t stream) {
  int rc = check_args(a, true);
  if (rc) return rc;
  if (a->B == 0 || a->Nq == 0 || a->Nk == 0) return GOAT_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!a->force_simt && attn_tc_eligible(a, true)) return attn_bwd_tc(a, st);
  if (a->dtype == GOAT_F32) return bwd_t<float>(*a, st);
  if (a->dtype == GOAT_F16) return bwd_t<__half>(*a, st);
  return bwd_t<__nv_bfloat16>(*a, st);
}
