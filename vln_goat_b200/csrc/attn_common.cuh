// Helpers shared by the tcgen05 attention kernels (attention_pipe.cu: persistent pipelined kernels for Nq, Nk <= 128;
// attention_tc.cu: query-tiled / key-chunked kernels for longer sequences).  sm_100a only.
#pragma once
#include "common.cuh"

namespace goat {
namespace attn {

constexpr int TILE = 128 * 128;      // bytes of a [128 rows][64 x 2 B] tile
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// 16-byte chunk `q` (8 elements) of row `r` inside a [rows][64] 128B-swizzled tile
__device__ __forceinline__ uint8_t* sw_chunk(uint8_t* tile, int r, int q) { return tile + r * 128 + ((q ^ (r & 7)) << 4); }

// 32 consecutive fp32 values (columns c0 .. c0+31 of row r) as 16-bit into the [2 tiles][128 rows][64] operand
template <typename T>
__device__ __forceinline__ void store_row32(uint8_t* buf, int r, int c0, const float (&v)[32]) {
  uint8_t* tile = buf + (c0 >> 6) * TILE;
  const int q0 = (c0 & 63) >> 3;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 w;
    w.x = pack2<T>(v[q * 8 + 0], v[q * 8 + 1]);
    w.y = pack2<T>(v[q * 8 + 2], v[q * 8 + 3]);
    w.z = pack2<T>(v[q * 8 + 4], v[q * 8 + 5]);
    w.w = pack2<T>(v[q * 8 + 6], v[q * 8 + 7]);
    *reinterpret_cast<uint4*>(sw_chunk(tile, r, q0 + q)) = w;
  }
}

// 16 consecutive fp32 values (columns c0 .. c0+15 of row r, c0 % 16 == 0) as 16-bit into the same operand
template <typename T>
__device__ __forceinline__ void store_row16(uint8_t* buf, int r, int c0, const float (&v)[16]) {
  uint8_t* tile = buf + (c0 >> 6) * TILE;
  const int q0 = (c0 & 63) >> 3;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    uint4 w;
    w.x = pack2<T>(v[q * 8 + 0], v[q * 8 + 1]);
    w.y = pack2<T>(v[q * 8 + 2], v[q * 8 + 3]);
    w.z = pack2<T>(v[q * 8 + 4], v[q * 8 + 5]);
    w.w = pack2<T>(v[q * 8 + 6], v[q * 8 + 7]);
    *reinterpret_cast<uint4*>(sw_chunk(tile, r, q0 + q)) = w;
  }
}

// additive log2-domain terms of 16 consecutive keys: staged key mask (+ the optional [Nq,Nk] bias row)
__device__ __forceinline__ void load_kv16(const float* km, const float* brow, int k0, int Nk, float (&kv)[16]) {
  const float4* k4 = reinterpret_cast<const float4*>(km + k0);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 t = k4[q];
    kv[q * 4 + 0] = t.x; kv[q * 4 + 1] = t.y; kv[q * 4 + 2] = t.z; kv[q * 4 + 3] = t.w;
  }
  if (brow) {   // uniform over the CTA
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (k0 + j < Nk) kv[j] = fmaf(__ldg(brow + k0 + j), LOG2E, kv[j]);
  }
}

// 32 fp32 values -> 64 bytes of 16-bit elements in global memory
template <typename T>
__device__ __forceinline__ void store_global_row32(T* dst, const uint32_t (&v)[32], float mul) {
  uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 w;
    w.x = pack2<T>(__uint_as_float(v[q * 8 + 0]) * mul, __uint_as_float(v[q * 8 + 1]) * mul);
    w.y = pack2<T>(__uint_as_float(v[q * 8 + 2]) * mul, __uint_as_float(v[q * 8 + 3]) * mul);
    w.z = pack2<T>(__uint_as_float(v[q * 8 + 4]) * mul, __uint_as_float(v[q * 8 + 5]) * mul);
    w.w = pack2<T>(__uint_as_float(v[q * 8 + 6]) * mul, __uint_as_float(v[q * 8 + 7]) * mul);
    d[q] = w;
  }
}

// same (row key, key pair) hashing as the SIMT kernels (attention.cu) and attention_tc.cu: masks agree across all paths.
// 16 consecutive keys starting at k0 (a multiple of 16): 8 pair hashes
__device__ __forceinline__ void drop_mul16(float (&v)[16], uint32_t rowkey, int k0, uint32_t thr16, float scale) {
#pragma unroll
  for (int j = 0; j < 16; j += 2) {
    const uint32_t r = attn_drop_pair(rowkey, k0 + j);
    v[j] = (r & 0xFFFFu) >= thr16 ? v[j] * scale : 0.f;
    v[j + 1] = (r >> 16) >= thr16 ? v[j + 1] * scale : 0.f;
  }
}

// chunked variant: the staged key mask `km` is indexed by the chunk-local key, the optional bias row by the global key
__device__ __forceinline__ void load_kv16c(const float* km, const float* brow, int k0_local, int k0_global, int Nk,
                                           float (&kv)[16]) {
  const float4* k4 = reinterpret_cast<const float4*>(km + k0_local);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 t = k4[q];
    kv[q * 4 + 0] = t.x; kv[q * 4 + 1] = t.y; kv[q * 4 + 2] = t.z; kv[q * 4 + 3] = t.w;
  }
  if (brow) {   // uniform over the CTA
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (k0_global + j < Nk) kv[j] = fmaf(__ldg(brow + k0_global + j), LOG2E, kv[j]);
  }
}

}  // namespace attn
}  // namespace goat
