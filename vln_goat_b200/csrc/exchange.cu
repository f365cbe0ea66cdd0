// The data-parallel gradient exchange fused with the optimizer step, over NVLink peer memory (one process per GPU).
//
// The reference exchanges gradients with DDP's bucketed all-reduce (P/utils/misc.py:52-58, find_unused_parameters=True)
// and then runs AdamW over every parameter on every rank.  The NCCL form of the sharded step here (engine.FlatParams.
// sharded_step) is reduce-scatter -> clip + AdamW on 1/world -> all-gather; measured on 2 B200s (profiles/
// r02o_overlap_probe.log) those collectives reach ~430 GB/s per direction and cost 2.6 ms of a 10.5 ms step, and they
// cannot hide under the backward pass: the persistent tcgen05 GEMM owns every SM, so a concurrent NCCL kernel only
// pushes the GEMMs into a second wave.  So the exchange is made as short as the wires allow instead, by doing the
// arithmetic AT the ends of the transfers:
//
//   goat_peer_reduce_sumsq  rank r reads ITS shard [lo, lo+S) of the flat gradient from every rank's buffer (peer loads
//                           through NVLink, own copy from HBM), sums them in rank order (deterministic), keeps the reduced
//                           shard locally and emits the partial sums of squares for the global norm
//                           = reduce-scatter + sumsq, no staging copy
//   (one scalar all-reduce of the squared norm: also the barrier "every rank has read my gradients")
//   goat_adamw_step_peers   clip + AdamW on the shard (same arithmetic as goat_adamw_step), and the results are STORED
//                           INTO EVERY RANK's buffers: the 16-bit operand shadow (hi, lo) of every element and the fp32
//                           value of the elements the kernels read in fp32 (>= n_fp32_from)
//                           = AdamW + all-gather + tail broadcast, no staging copy
//
// Peer buffers are the caller's own device allocations (torch's caching allocator, cudaMalloc segments): goat_peer_export
// returns the CUDA IPC handle of the allocation a pointer lies in plus the pointer's offset inside it, the other ranks
// map it with goat_peer_open; the handles travel through the process group's object all-gather on the host side.
#include <cuda.h>
#include <string.h>

#include "common.cuh"

namespace goat {
namespace {

constexpr int XCH_THREADS = 256;
constexpr int XCH_MAX_PARTS = 1184;   // == SUMSQ_MAX_PARTS (optim.cu): the partial buffer is shared

struct PeerPtrs {
  void* p[GOAT_MAX_PEERS];
};
// p[r] for a run-time r without indexing the kernel parameter dynamically (which would copy it to local memory)
__device__ __forceinline__ void* peer_at(const PeerPtrs& t, int r) {
  void* x = t.p[0];
#pragma unroll
  for (int q = 1; q < GOAT_MAX_PEERS; ++q)
    if (q == r) x = t.p[q];
  return x;
}

__global__ void __launch_bounds__(XCH_THREADS)
peer_reduce_sumsq_kernel(PeerPtrs g, int world, long long lo, long long n, float* __restrict__ out,
                         float* __restrict__ partial) {
  __shared__ float red[XCH_THREADS / 32];
  float s = 0.f;
  const long long n4 = n >> 2;
  // two independent float4 columns per thread and iteration: 2 x world 16-byte loads in flight (NVLink round trips are
  // ~2 us; the link needs ~2 MB in flight per GPU)
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += 2 * stride) {
    const long long i2 = i + stride;
    const bool two = i2 < n4;
    float4 v[GOAT_MAX_PEERS], w[GOAT_MAX_PEERS];
#pragma unroll
    for (int r = 0; r < GOAT_MAX_PEERS; ++r) {
      if (r < world) {
        const float4* src = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(g.p[r]) + lo);
        v[r] = __ldcs(src + i);
        if (two) w[r] = __ldcs(src + i2);
      }
    }
    float4 a = v[0], b = two ? w[0] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 1; r < GOAT_MAX_PEERS; ++r) {
      if (r < world) {
        a.x += v[r].x; a.y += v[r].y; a.z += v[r].z; a.w += v[r].w;
        if (two) { b.x += w[r].x; b.y += w[r].y; b.z += w[r].z; b.w += w[r].w; }
      }
    }
    reinterpret_cast<float4*>(out)[i] = a;
    s += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
    if (two) {
      reinterpret_cast<float4*>(out)[i2] = b;
      s += b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
    }
  }
  if (blockIdx.x == 0) {
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
      float a = 0.f;
#pragma unroll
      for (int r = 0; r < GOAT_MAX_PEERS; ++r)
        if (r < world) a += reinterpret_cast<const float*>(g.p[r])[lo + i];
      out[i] = a;
      s += a * a;
    }
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < XCH_THREADS / 32; ++w) t += red[w];
    partial[blockIdx.x] = t;
  }
}

// Same update as adamw_kernel (optim.cu) on elements [lo, lo + n) of the flat buffers; `g` is the reduced shard
// (indexed from 0).  hp / scaler / partial as there.  The new values go to every rank: shadow (and shadow_lo) for all
// elements, the fp32 value for elements >= n_fp32_from.  The own rank's fp32 master is always written.
template <typename TS>
__global__ void __launch_bounds__(XCH_THREADS)
adamw_peers_kernel(PeerPtrs pp, PeerPtrs sh, PeerPtrs shlo, int world, int rank, const float* __restrict__ g,
                   float* __restrict__ m, float* __restrict__ v, long long lo, long long n, long long n_decay,
                   long long n_fp32_from, const float* __restrict__ hp, const float* __restrict__ partial, int nparts,
                   float* __restrict__ norm_out, const float* __restrict__ scaler) {
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
        const float c = hp[7] / (norm + 1e-6f);
        if (c < 1.f) coef *= c;
      }
      s_coef = coef;
      s_skip = (scaler && !isfinite(t)) ? 1 : 0;
      if (blockIdx.x == 0 && norm_out) *norm_out = norm;
    }
  }
  __syncthreads();
  if (s_skip) return;   // overflowed fp16 gradients: nothing changes anywhere
  const float coef = s_coef;
  const float lr = hp[0], b1 = hp[1], b2 = hp[2], eps = hp[3], wd = hp[4];
  float bc1 = hp[5], bc2 = hp[6];
  if (scaler && hp[5] != 1.0f) {
    const float t = scaler[4] + 1.0f;
    bc1 = 1.0f - powf(b1, t);
    bc2 = 1.0f - powf(b2, t);
  }
  const float step_size = lr * sqrtf(bc2) / bc1;
  float* p_own = reinterpret_cast<float*>(peer_at(pp, rank));
  // A warp owns 256 consecutive elements and every lane two float4 columns of them (elements 4*lane.. and 128 + 4*lane..):
  // each fp32 load / store instruction covers 512 contiguous bytes, each 16-bit store 256 -- whole 32-byte sectors, which
  // is what peer stores need (16-byte stores at a 32-byte stride half-filled every sector and cost the owners of the fp32
  // region 1.8 ms instead of 1.0 ms at 8 GPUs, profiles/r02p_peer_phases_n8_strided_stores.txt).
  const long long nblk = n >> 8;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  for (long long blk = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; blk < nblk; blk += warps) {
    float x[2][4], mm[2][4], vv[2][4];
    float4 gg[2];
    long long e[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const long long i = blk * 256 + h * 128 + lane * 4;     // index inside the shard
      e[h] = lo + i;
      *reinterpret_cast<float4*>(x[h]) = *reinterpret_cast<const float4*>(p_own + e[h]);
      gg[h] = __ldcs(reinterpret_cast<const float4*>(g + i));
      *reinterpret_cast<float4*>(mm[h]) = *reinterpret_cast<const float4*>(m + e[h]);
      *reinterpret_cast<float4*>(vv[h]) = *reinterpret_cast<const float4*>(v + e[h]);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float* ga = &gg[h].x;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float gr = ga[j] * coef;
        mm[h][j] = mm[h][j] * b1 + gr * (1.f - b1);
        vv[h][j] = vv[h][j] * b2 + gr * gr * (1.f - b2);
        float y = x[h][j] - step_size * (mm[h][j] / (sqrtf(vv[h][j]) + eps));
        if (e[h] + j < n_decay) y -= y * (lr * wd);
        x[h][j] = y;
      }
      *reinterpret_cast<float4*>(m + e[h]) = *reinterpret_cast<const float4*>(mm[h]);
      *reinterpret_cast<float4*>(v + e[h]) = *reinterpret_cast<const float4*>(vv[h]);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float4 x4 = *reinterpret_cast<const float4*>(x[h]);
      const bool fp32_everywhere = e[h] >= n_fp32_from;   // region boundaries are multiples of 8: never inside a float4
      if (!fp32_everywhere) *reinterpret_cast<float4*>(p_own + e[h]) = x4;
      uint2 hi = make_uint2(0u, 0u), l2 = hi;
      if constexpr (sizeof(TS) == 2) {
        hi.x = pack2<TS>(x4.x, x4.y);
        hi.y = pack2<TS>(x4.z, x4.w);
        const float2 h0 = unpack2<TS>(hi.x), h1 = unpack2<TS>(hi.y);
        l2.x = pack2<TS>(x4.x - h0.x, x4.y - h0.y);
        l2.y = pack2<TS>(x4.z - h1.x, x4.w - h1.y);
      }
#pragma unroll
      for (int r = 0; r < GOAT_MAX_PEERS; ++r) {
        if (r < world) {
          // 4 bytes per element and rank either way: the fp32 value (its 16-bit copies are re-derived locally by
          // goat_split_cast) or hi + lo.  Storing all three for the fp32 region made its owners (the last ranks: embedding
          // tables) send twice the bytes of the others -- 2.7 ms instead of 1.0 ms at 8 GPUs, everybody waiting for them.
          if (fp32_everywhere) {
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(pp.p[r]) + e[h]) = x4;
          } else if constexpr (sizeof(TS) == 2) {
            if (sh.p[r]) *reinterpret_cast<uint2*>(reinterpret_cast<TS*>(sh.p[r]) + e[h]) = hi;
            if (shlo.p[r]) *reinterpret_cast<uint2*>(reinterpret_cast<TS*>(shlo.p[r]) + e[h]) = l2;
          }
        }
      }
    }
  }
  const long long n4 = nblk << 6;    // the scalar tail below starts at element n4 * 4 = nblk * 256
  if (blockIdx.x == 0) {
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
      const long long e = lo + i;
      const float gr = g[i] * coef;
      const float mi = m[e] * b1 + gr * (1.f - b1);
      const float vi = v[e] * b2 + gr * gr * (1.f - b2);
      float x = p_own[e] - step_size * (mi / (sqrtf(vi) + eps));
      if (e < n_decay) x -= x * (lr * wd);
      m[e] = mi; v[e] = vi;
      p_own[e] = x;
#pragma unroll
      for (int r = 0; r < GOAT_MAX_PEERS; ++r) {
        if (r >= world) continue;
        if (e >= n_fp32_from) reinterpret_cast<float*>(pp.p[r])[e] = x;
        if constexpr (sizeof(TS) == 2) {
          if (e < n_fp32_from) {
            const TS h = from_f<TS>(x);
            if (sh.p[r]) reinterpret_cast<TS*>(sh.p[r])[e] = h;
            if (shlo.p[r]) reinterpret_cast<TS*>(shlo.p[r])[e] = from_f<TS>(x - to_f<TS>(h));
          }
        }
      }
    }
  }
}

// ---- rank synchronisation through flags in peer memory --------------------------------------------------------
// Every rank owns a small signal block: flags[GOAT_MAX_PEERS] (u32, flags[q] = last epoch rank q announced to me) and
// vals[2][GOAT_MAX_PEERS] (fp32).  One CTA: thread q < world publishes to rank q (value, system fence, flag), then
// waits until rank q's announcement of this epoch has arrived in the own block.  All work enqueued before this kernel on
// the stream (peer stores of the AdamW kernel, the backward pass) is complete, and therefore visible with the flag.
constexpr int SIG_VALS_OFFSET = 64;   // bytes

__device__ __forceinline__ void signal_and_wait(const PeerPtrs& sig, int world, int rank, unsigned epoch) {
  const int q = threadIdx.x;
  if (q < world) {
    __threadfence_system();
    volatile unsigned* theirs = reinterpret_cast<volatile unsigned*>(peer_at(sig, q)) + rank;
    *theirs = epoch;
    volatile unsigned* mine = reinterpret_cast<volatile unsigned*>(peer_at(sig, rank)) + q;
    while ((int)(*mine - epoch) < 0) { }
    __threadfence_system();
  }
  __syncthreads();
}

__global__ void __launch_bounds__(32) peer_barrier_kernel(PeerPtrs sig, int world, int rank, unsigned epoch) {
  signal_and_wait(sig, world, rank, epoch);
}

// out[0] = sum over ranks (rank order) of (sum of this rank's partial[0..nparts)); identical bits on every rank
__global__ void __launch_bounds__(32)
peer_sum_scalar_kernel(PeerPtrs sig, int world, int rank, unsigned epoch, const float* partial, int nparts, float* out) {
  float t = 0.f;
  for (int i = threadIdx.x; i < nparts; i += 32) t += partial[i];
  t = warp_sum(t);
  const int slot = (int)(epoch & 1u) * GOAT_MAX_PEERS;
  if ((int)threadIdx.x < world) {
    volatile float* theirs = reinterpret_cast<volatile float*>(reinterpret_cast<char*>(peer_at(sig, threadIdx.x)) + SIG_VALS_OFFSET);
    theirs[slot + rank] = t;
  }
  signal_and_wait(sig, world, rank, epoch);
  if (threadIdx.x == 0) {
    const volatile float* mine = reinterpret_cast<const volatile float*>(reinterpret_cast<const char*>(peer_at(sig, rank)) + SIG_VALS_OFFSET);
    float tot = 0.f;
    for (int q = 0; q < world; ++q) tot += mine[slot + q];
    out[0] = tot;
  }
}

// hi = round16(x), lo = round16(x - hi): the split operand copies of an fp32 range (local; HBM-bound, 8 B per element)
template <typename TS>
__global__ void __launch_bounds__(XCH_THREADS)
split_cast_kernel(const float* __restrict__ x, TS* __restrict__ hi, TS* __restrict__ lo, long long n) {
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(x)[i];
    uint2 h;
    h.x = pack2<TS>(a.x, a.y);
    h.y = pack2<TS>(a.z, a.w);
    reinterpret_cast<uint2*>(hi)[i] = h;
    if (lo) {
      const float2 h0 = unpack2<TS>(h.x), h1 = unpack2<TS>(h.y);
      uint2 l;
      l.x = pack2<TS>(a.x - h0.x, a.y - h0.y);
      l.y = pack2<TS>(a.z - h1.x, a.w - h1.y);
      reinterpret_cast<uint2*>(lo)[i] = l;
    }
  }
  if (blockIdx.x == 0) {
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
      const TS h = from_f<TS>(x[i]);
      hi[i] = h;
      if (lo) lo[i] = from_f<TS>(x[i] - to_f<TS>(h));
    }
  }
}

int fill_peers(PeerPtrs& dst, void* const* src, int world) {
  for (int r = 0; r < GOAT_MAX_PEERS; ++r) dst.p[r] = (src && r < world) ? src[r] : nullptr;
  return 0;
}

}  // namespace
}  // namespace goat

using namespace goat;

extern "C" int goat_peer_export(const void* ptr, void* handle_out, unsigned long long* offset_out) {
  GOAT_CHECK(ptr && handle_out && offset_out, "goat_peer_export: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == GOAT_PEER_HANDLE_BYTES, "handle size");
  typedef CUresult (*RangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
  static RangeFn range = nullptr;
  if (!range) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      range = reinterpret_cast<RangeFn>(f);
  }
  GOAT_CHECK(range != nullptr, "cuMemGetAddressRange not available from the driver");
  CUdeviceptr base = 0;
  size_t size = 0;
  GOAT_CHECK(range(&base, &size, reinterpret_cast<CUdeviceptr>(ptr)) == CUDA_SUCCESS,
             "goat_peer_export: %p is not inside a device allocation", ptr);
  cudaIpcMemHandle_t h;
  // fails for memory that is not a plain cudaMalloc allocation (e.g. the caching allocator's expandable segments)
  GOAT_CUDA(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base)));
  memcpy(handle_out, &h, sizeof(h));
  *offset_out = (unsigned long long)(reinterpret_cast<CUdeviceptr>(ptr) - base);
  return GOAT_OK;
}

extern "C" int goat_peer_open(const void* handle, void** base_out) {
  GOAT_CHECK(handle && base_out, "goat_peer_open: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  GOAT_CUDA(cudaIpcOpenMemHandle(base_out, h, cudaIpcMemLazyEnablePeerAccess));
  return GOAT_OK;
}

extern "C" int goat_peer_close(void* ptr) {
  if (ptr) GOAT_CUDA(cudaIpcCloseMemHandle(ptr));
  return GOAT_OK;
}

extern "C" int goat_peer_reduce_sumsq(void* const* g_peers, int world, long long lo, long long n, float* g_shard_out,
                                      float* partial, int* nparts_out, goat_stream_t stream) {
  GOAT_CHECK(g_peers && g_shard_out && partial && nparts_out, "goat_peer_reduce_sumsq: null argument");
  GOAT_CHECK(world >= 1 && world <= GOAT_MAX_PEERS, "goat_peer_reduce_sumsq: world %d outside 1..%d", world, GOAT_MAX_PEERS);
  GOAT_CHECK(lo >= 0 && (lo & 3) == 0 && aligned16(g_shard_out), "goat_peer_reduce_sumsq: shard must start 16-byte aligned");
  for (int r = 0; r < world; ++r) GOAT_CHECK(g_peers[r] && aligned16(g_peers[r]), "goat_peer_reduce_sumsq: bad peer pointer %d", r);
  long long want = (n / 4 + XCH_THREADS - 1) / XCH_THREADS;
  const int parts = (int)(want < 1 ? 1 : (want > XCH_MAX_PARTS ? XCH_MAX_PARTS : want));
  *nparts_out = parts;
  if (n <= 0) {
    GOAT_CUDA(cudaMemsetAsync(partial, 0, sizeof(float), reinterpret_cast<cudaStream_t>(stream)));
    return GOAT_OK;
  }
  PeerPtrs g;
  fill_peers(g, g_peers, world);
  peer_reduce_sumsq_kernel<<<parts, XCH_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(g, world, lo, n, g_shard_out, partial);
  GOAT_LAUNCH_CHECK();
  return GOAT_OK;
}

extern "C" int goat_adamw_step_peers(void* const* p_peers, void* const* shadow_peers, void* const* shadow_lo_peers,
                                     int shadow_dtype, int world, int rank, const float* g_shard, float* m, float* v,
                                     long long lo, long long n, long long n_decay, long long n_fp32_from, const float* hp,
                                     const float* partial, int nparts, float* norm_out, const float* scaler,
                                     goat_stream_t stream) {
  GOAT_CHECK(p_peers && g_shard && m && v && hp && partial, "goat_adamw_step_peers: null argument");
  GOAT_CHECK(world >= 1 && world <= GOAT_MAX_PEERS && rank >= 0 && rank < world, "goat_adamw_step_peers: bad world / rank");
  GOAT_CHECK(lo >= 0 && (lo & 7) == 0 && (n_fp32_from & 7) == 0, "goat_adamw_step_peers: lo and n_fp32_from must be multiples of 8");
  GOAT_CHECK(!shadow_lo_peers || shadow_peers, "goat_adamw_step_peers: shadow_lo needs shadow");
  GOAT_CHECK(!shadow_peers || shadow_dtype == GOAT_F16 || shadow_dtype == GOAT_BF16, "goat_adamw_step_peers: shadow dtype must be F16/BF16");
  GOAT_CHECK(nparts >= 1 && nparts <= XCH_MAX_PARTS, "goat_adamw_step_peers: bad nparts");
  GOAT_CHECK(aligned16(g_shard) && aligned16(m) && aligned16(v), "goat_adamw_step_peers: buffers must be 16-byte aligned");
  for (int r = 0; r < world; ++r) {
    GOAT_CHECK(p_peers[r] && aligned16(p_peers[r]), "goat_adamw_step_peers: bad parameter pointer of rank %d", r);
    GOAT_CHECK(!shadow_peers || (shadow_peers[r] && aligned16(shadow_peers[r])), "goat_adamw_step_peers: bad shadow pointer of rank %d", r);
    GOAT_CHECK(!shadow_lo_peers || (shadow_lo_peers[r] && aligned16(shadow_lo_peers[r])), "goat_adamw_step_peers: bad shadow_lo pointer of rank %d", r);
  }
  if (n <= 0) return GOAT_OK;
  PeerPtrs pp, sh, shlo;
  fill_peers(pp, p_peers, world);
  fill_peers(sh, shadow_peers, world);
  fill_peers(shlo, shadow_lo_peers, world);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  long long want = (n / 4 + XCH_THREADS - 1) / XCH_THREADS;
  const int grid = (int)(want < 1 ? 1 : (want > 148 * 16 ? 148 * 16 : want));
  if (shadow_peers && shadow_dtype == GOAT_F16)
    adamw_peers_kernel<__half><<<grid, XCH_THREADS, 0, st>>>(pp, sh, shlo, world, rank, g_shard, m, v, lo, n, n_decay, n_fp32_from, hp, partial, nparts, norm_out, scaler);
  else if (shadow_peers)
    adamw_peers_kernel<__nv_bfloat16><<<grid, XCH_THREADS, 0, st>>>(pp, sh, shlo, world, rank, g_shard, m, v, lo, n, n_decay, n_fp32_from, hp, partial, nparts, norm_out, scaler);
  else
    adamw_peers_kernel<float><<<grid, XCH_THREADS, 0, st>>>(pp, sh, shlo, world, rank, g_shard, m, v, lo, n, n_decay, n_fp32_from, hp, partial, nparts, norm_out, scaler);
  GOAT_LAUNCH_CHECK();
  return GOAT_OK;
}

extern "C" size_t goat_peer_signal_bytes(void) { return 256; }

extern "C" int goat_peer_barrier(void* const* signal_peers, int world, int rank, unsigned int epoch, goat_stream_t stream) {
  GOAT_CHECK(signal_peers, "goat_peer_barrier: null argument");
  GOAT_CHECK(world >= 1 && world <= GOAT_MAX_PEERS && rank >= 0 && rank < world, "goat_peer_barrier: bad world / rank");
  for (int r = 0; r < world; ++r) GOAT_CHECK(signal_peers[r] && aligned16(signal_peers[r]), "goat_peer_barrier: bad signal pointer %d", r);
  PeerPtrs sig;
  fill_peers(sig, signal_peers, world);
  peer_barrier_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(sig, world, rank, epoch);
  GOAT_LAUNCH_CHECK();
  return GOAT_OK;
}

extern "C" int goat_peer_sum_scalar(void* const* signal_peers, int world, int rank, unsigned int epoch, const float* partial,
                                    int nparts, float* out, goat_stream_t stream) {
  GOAT_CHECK(signal_peers && partial && out, "goat_peer_sum_scalar: null argument");
  GOAT_CHECK(world >= 1 && world <= GOAT_MAX_PEERS && rank >= 0 && rank < world, "goat_peer_sum_scalar: bad world / rank");
  GOAT_CHECK(nparts >= 1 && nparts <= XCH_MAX_PARTS, "goat_peer_sum_scalar: bad nparts");
  for (int r = 0; r < world; ++r) GOAT_CHECK(signal_peers[r] && aligned16(signal_peers[r]), "goat_peer_sum_scalar: bad signal pointer %d", r);
  PeerPtrs sig;
  fill_peers(sig, signal_peers, world);
  peer_sum_scalar_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(sig, world, rank, epoch, partial, nparts, out);
  GOAT_LAUNCH_CHECK();
  return GOAT_OK;
}

extern "C" int goat_split_cast(const float* x, void* hi, void* lo, int dtype, long long n, goat_stream_t stream) {
  GOAT_CHECK(x && hi, "goat_split_cast: null argument");
  GOAT_CHECK(dtype == GOAT_F16 || dtype == GOAT_BF16, "goat_split_cast: dtype must be F16/BF16");
  GOAT_CHECK(aligned16(x) && (reinterpret_cast<uintptr_t>(hi) & 7) == 0 && (reinterpret_cast<uintptr_t>(lo) & 7) == 0,
             "goat_split_cast: x must be 16-byte, hi / lo 8-byte aligned");
  if (n <= 0) return GOAT_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  long long want = (n / 4 + XCH_THREADS - 1) / XCH_THREADS;
  const int grid = (int)(want < 1 ? 1 : (want > 148 * 16 ? 148 * 16 : want));
  if (dtype == GOAT_F16) split_cast_kernel<__half><<<grid, XCH_THREADS, 0, st>>>(x, (__half*)hi, (__half*)lo, n);
  else split_cast_kernel<__nv_bfloat16><<<grid, XCH_THREADS, 0, st>>>(x, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, n);
  GOAT_LAUNCH_CHECK();
  return GOAT_OK;
}
