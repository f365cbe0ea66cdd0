// 2-CTA tcgen05 GEMM for sm_100a: out[M,N] = epilogue(A[M,K] * B[N,K]^T), fp16/bf16 operands, fp32 accumulation.
//
// Why a CTA pair: with one CTA per 128 x 128 tile every k-block of 64 pulls 32 KB through L2 for 256 tensor-pipe
// clocks of work -- 128 B/clk/SM, three times what the L2 can deliver to 148 SMs (~6.3 KB/clk chip-wide), so the
// single-CTA kernel saturates L2 at about a third of the tensor peak.  Here two CTAs of a cluster (one TPC) share one
// 256 x BLOCK_N tile through tcgen05.mma.cta_group::2: each CTA stages ITS 128 rows of A and only HALF of the B tile
// (BLOCK_N/2 rows), the pair-wide MMA reads both halves, and each CTA's TMEM receives its own 128 x BLOCK_N slice of
// the accumulator.  Bytes through L2 per FLOP drop 1.33x (BLOCK_N 128) or 2x (BLOCK_N 256).
//
// Persistent and warp-specialised like the single-CTA kernel (gemm_umma.cu), per CTA:
//   warp 0    : TMA producer -- own A rows + own half of B into a STAGES-deep ring; the transaction bytes of BOTH
//               CTAs complete on the LEADER's (cluster rank 0) full barrier (cp.async.bulk.tensor .cta_group::2)
//   warp 1    : TMEM allocator (cta_group::2 alloc, both CTAs) + on the leader one elected lane issuing the pair MMAs
//               (UMMA 256 x BLOCK_N x 16) into one of TWO TMEM accumulators; tcgen05.commit multicasts the
//               "stage free" / "accumulator ready" arrivals to both CTAs
//   warps 2-9 : epilogue of the CTA's own 128 x BLOCK_N slice -- tcgen05.ld 32 x 32 patches, transpose through a
//               private padded smem patch so that every global access is a coalesced 16-byte access along the row;
//               the residual / GELU-aux operands of the NEXT patch are fetched before the current one is processed
//               (and those of a tile's first patch before the accumulator-ready wait), so their latency overlaps
//               TMEM reads and math.  Both CTAs' epilogue warps release the accumulator on the leader's barrier.
// Split-K (accumulate mode), K-major / MN-major operands and ragged edges as in gemm_umma.cu.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace goat {

int make_tmap(CUtensorMap* tm, int dtype, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
              uint32_t box_inner, uint32_t box_outer);
int make_tmap_sw64(CUtensorMap* tm, int dtype, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                   uint32_t box_inner, uint32_t box_outer);
int num_sms();

#ifdef GOAT_TIMELINE
__device__ long long g_timeline[4096];
#define TL(slot, idx) do { if (blockIdx.x == 0 && (idx) < 64) g_timeline[(slot) * 64 + (idx)] = clock64(); } while (0)
#else
#define TL(slot, idx) do { } while (0)
#endif

namespace {

constexpr int BM = 128;   // rows of A per CTA; the pair tile is 256 rows
constexpr int BK = 64;    // 64 x 2 B = one 128-byte swizzle row
constexpr int UK = 16;
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 64 + EPI_WARPS * 32;
constexpr int PATCH_LD = 36;                 // floats per staged row (32 + 4 padding: conflict-free both ways)
constexpr int PATCH_FLOATS = 1280;           // one warp's staging area: the 32 x 36 fp32 transpose patch (4608 B) or two
                                             // 2 KB TMA-store tiles (32 rows x 32 x 16 bit, 64B swizzle); 5120 B keeps every
                                             // warp's area 512-byte aligned, which the swizzle pattern needs

template <int BN, int STAGES>
struct Cfg2 {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (BN / 2) * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_BYTES = EPI_WARPS * PATCH_FLOATS * 4;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024;  // +1024: manual alignment
  static_assert((2 * STAGES + 4) * 8 + 4 <= BAR_BYTES, "barrier block too small");
  static_assert(SMEM_BYTES <= 232448, "over the 227 KB dynamic shared memory limit");
};

struct Sched2 {
  int tiles_m, tiles_n, splits, kb_per_split, num_kb, num_tiles;
  int kb_wrap;   // > 0: k-blocks [kb_wrap, num_kb) re-read A from k-block (kb - kb_wrap) against the second B operand (B_lo)
};

template <int BN>
__device__ __forceinline__ void tile_coords2(const Sched2& sc, int t, int& m0, int& n0, int& kb0, int& kb1) {
  const int tn = t % sc.tiles_n;
  const int r = t / sc.tiles_n;
  const int tm = r % sc.tiles_m;
  const int ks = r / sc.tiles_m;
  m0 = tm * (2 * BM);
  n0 = tn * BN;
  kb0 = ks * sc.kb_per_split;
  kb1 = min(sc.num_kb, kb0 + sc.kb_per_split);
}

// ---- cluster / cta_group::2 PTX ----------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_smem_addr`'s twin in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load into this CTA's smem whose transaction bytes complete on a barrier that may live in the peer CTA
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const void* tmap, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_slot) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives (once all previously issued MMAs of this thread completed) on the barrier at this smem offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3)
               : "memory");
}

// ---- epilogue ------------------------------------------------------------------------------------
// The epilogue is compiled once per MODE and selected by a uniform switch outside the persistent tile loop, with the
// per-tile patch loop NOT unrolled: what one launch executes is a few hundred SASS instructions that stay in the 32 KB
// L1.5 instruction cache.  (One body with every bias / activation / dropout / residual / output-type branch, unrolled
// over patches, was ~20k instructions = 300 KB and spent most of its issue slots on instruction fetch.)
enum {
  EPI_ACC = 0,     // fp32 += alpha * acc, vector atomics (split-K weight gradients)
  EPI_T16 = 1,     // alpha, [bias] -> 16-bit out
  EPI_GELU = 2,    // [bias] -> 16-bit pre-activation (aux_out) -> GELU, [dropout] -> 16-bit out
  EPI_DGELU = 3,   // * gelu'(aux_in), [dropout] -> 16-bit out
  EPI_RES32 = 4,   // [bias], [dropout], [+ res] -> fp32 out (+ optional 16-bit copy out2)
  EPI_GENERIC = 5, // anything else (ReLU / tanh heads, unaligned leading dimensions): predicated scalar path
  // 16-bit outputs through TMA stores: TMEM -> registers (one accumulator row per lane) -> 2 KB swizzled smem tile ->
  // cp.async.bulk.tensor store.  No fp32 transpose through shared memory (which competed with the MMA's operand reads and
  // slowed the main loop of the next tile by 25-70 %, profiles/r02g_timeline.log), no per-lane global addressing, ragged
  // edges clipped by the hardware.
  EPI_T16_TMA = 6, EPI_GELU_TMA = 7, EPI_DGELU_TMA = 8
};

__device__ __forceinline__ void tma_store_2d(const void* tmap, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// 16-byte chunk q (8 x 16-bit) of row `row` in a [32 rows][64 B] tile with the 64-byte swizzle (address bits 4-5 ^= bits 7-8)
__device__ __forceinline__ uint8_t* sw64_chunk(uint8_t* tile, int row, int q) {
  return tile + row * 64 + ((q ^ ((row >> 1) & 3)) << 4);
}

// operands of one 32 x 32 patch that do not depend on the accumulator: row (i*4 + lane/8), columns (lane%8)*4..+3
template <int MODE>
struct Pre {
  float4 res[MODE == EPI_RES32 ? 8 : 1];
  uint2 aux[MODE == EPI_DGELU ? 8 : 1];
};

template <typename T, int MODE>
__device__ __forceinline__ void prefetch_patch(const EpiParams& ep, int mrow0, int ncol, int rl, Pre<MODE>& p) {
  if constexpr (MODE == EPI_RES32) {
    if (ep.res) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        p.res[i] = *reinterpret_cast<const float4*>(ep.res + (size_t)(mrow0 + i * 4 + rl) * ep.ldres + ncol);
    }
  }
  if constexpr (MODE == EPI_DGELU) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      p.aux[i] = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const T*>(ep.aux_in) + (size_t)(mrow0 + i * 4 + rl) * ep.ldaux + ncol));
  }
}

// 4 consecutive outputs (row m, columns n..n+3) of a full, aligned patch
template <typename T, int MODE>
__device__ __forceinline__ void epi_out4(const EpiParams& ep, int m, int n, float4 acc, float4 b4, const Pre<MODE>& pre, int i,
                                         float keep, uint32_t thr16, unsigned long long seed) {
  float v[4] = {acc.x * ep.alpha, acc.y * ep.alpha, acc.z * ep.alpha, acc.w * ep.alpha};
  if constexpr (MODE == EPI_ACC) {
    atomicAdd(reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + (size_t)m * ep.ldc + n),
              make_float4(v[0], v[1], v[2], v[3]));
    return;
  }
  v[0] += b4.x; v[1] += b4.y; v[2] += b4.z; v[3] += b4.w;   // zeros when there is no bias
  if constexpr (MODE == EPI_GELU) {
    uint2 w;
    w.x = pack2<T>(v[0], v[1]);
    w.y = pack2<T>(v[2], v[3]);
    *reinterpret_cast<uint2*>(reinterpret_cast<T*>(ep.aux_out) + (size_t)m * ep.ldaux + n) = w;
    // GELU of the ROUNDED pre-activation: backward only ever sees the stored 16-bit z
    float2 f;
    f = unpack2<T>(w.x); v[0] = f.x; v[1] = f.y;
    f = unpack2<T>(w.y); v[2] = f.x; v[3] = f.y;
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = gelu_fast(v[j]);
  }
  if constexpr (MODE == EPI_DGELU) {
    const float2 a = unpack2<T>(pre.aux[i].x), b = unpack2<T>(pre.aux[i].y);
    v[0] *= dgelu_fast(a.x); v[1] *= dgelu_fast(a.y); v[2] *= dgelu_fast(b.x); v[3] *= dgelu_fast(b.y);
  }
  if constexpr (MODE != EPI_T16) {
    if (ep.drop_p > 0.0f) drop_apply4(v, seed, (unsigned long long)(m + ep.drop_row0) * (unsigned long long)ep.ldc + n, thr16, keep);
  }
  if constexpr (MODE == EPI_RES32) {
    if (ep.res) { v[0] += pre.res[i].x; v[1] += pre.res[i].y; v[2] += pre.res[i].z; v[3] += pre.res[i].w; }
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + (size_t)m * ep.ldc + n) = make_float4(v[0], v[1], v[2], v[3]);
    if (ep.out2) {
      uint2 w;
      w.x = pack2<T>(v[0], v[1]);
      w.y = pack2<T>(v[2], v[3]);
      *reinterpret_cast<uint2*>(reinterpret_cast<T*>(ep.out2) + (size_t)m * ep.ldc2 + n) = w;
    }
  } else {
    uint2 w;
    w.x = pack2<T>(v[0], v[1]);
    w.y = pack2<T>(v[2], v[3]);
    *reinterpret_cast<uint2*>(reinterpret_cast<T*>(ep.out) + (size_t)m * ep.ldc + n) = w;
  }
}

// predicated scalar path for ragged patches and the GENERIC mode (kept out of line: it is cold).  `ep` BY VALUE: taking
// the address of the kernel parameter would give it a stack copy that the hot loops then re-read (LDL) after every store.
template <typename T>
__device__ __noinline__ void epi_patch_scalar(const EpiParams ep, const float* patch, int mrow0, int nc, int M, int N, int rl,
                                              int c4) {
#pragma unroll 1
  for (int i = 0; i < 8; ++i) {
    const int row_l = i * 4 + rl;
    const int m = mrow0 + row_l;
#pragma unroll 1
    for (int j = 0; j < 4; ++j) {
      const int n = nc + c4 + j;
      if (m < M && n < N) {
        const float acc = patch[row_l * PATCH_LD + c4 + j];
        if (ep.accumulate) {
          atomicAdd(reinterpret_cast<float*>(ep.out) + (size_t)m * ep.ldc + n, acc * ep.alpha);
        } else {
          const float val = epi_apply<T>(ep, m, n, acc);
          if (ep.out_f32) reinterpret_cast<float*>(ep.out)[(size_t)m * ep.ldc + n] = val;
          else reinterpret_cast<T*>(ep.out)[(size_t)m * ep.ldc + n] = from_f<T>(val);
          if (ep.out2) reinterpret_cast<T*>(ep.out2)[(size_t)m * ep.ldc2 + n] = from_f<T>(val);
        }
      }
    }
  }
}

struct EpiCtx {
  uint64_t* tmem_full_bar;
  uint32_t release_bar0, release_bar1;   // the leader's tmem_empty barriers (shared::cluster addresses)
  uint32_t tmem_base;
  float* patch;
  int cluster_id, num_clusters, row_off, col_off, lg, lane, M, N;
};

template <typename T, int BN, int MODE>
__device__ __forceinline__ void epilogue_loop(const EpiParams& ep, const Sched2& sc, const EpiCtx& cx) {
  constexpr int NCH = BN / 64;      // 32-column patches per epilogue warp and tile
  const int M = cx.M, N = cx.N;
  const int rl = cx.lane >> 3;         // row within a group of 4
  const int c4 = (cx.lane & 7) * 4;    // column offset of this lane's float4
  const unsigned long long seed = ep.drop_p > 0.0f ? eff_seed(ep.drop_seed, ep.drop_seed_ptr) : 0ull;
  const float keep = ep.drop_p > 0.0f ? 1.0f / (1.0f - ep.drop_p) : 1.0f;
  const uint32_t thr16 = drop_thr16(ep.drop_p);
  float* patch = cx.patch;

  int local = 0;
  int t = cx.cluster_id;
  int m0 = 0, n0 = 0, kb0 = 0, kb1 = 0;
  Pre<MODE> cur;
  bool cur_full = false;
  if (t < sc.num_tiles) {
    tile_coords2<BN>(sc, t, m0, n0, kb0, kb1);
    cur_full = MODE != EPI_GENERIC && (m0 + cx.row_off + 32 <= M) && (n0 + cx.col_off + 32 <= N);
    if (cur_full) prefetch_patch<T, MODE>(ep, m0 + cx.row_off, n0 + cx.col_off + c4, rl, cur);
  }
  for (; t < sc.num_tiles; t += cx.num_clusters, ++local) {
    const int buf = local & 1;
    const int mrow0 = m0 + cx.row_off;
    // coordinates of this cluster's next tile (for the cross-tile prefetch)
    int nm0 = 0, nn0 = 0, nkb0 = 0, nkb1 = 0;
    const bool has_next = t + cx.num_clusters < sc.num_tiles;
    if (has_next) tile_coords2<BN>(sc, t + cx.num_clusters, nm0, nn0, nkb0, nkb1);
    if (cx.lane == 0 && cx.lg == 0 && cx.col_off == 0) TL(4, local);
    mbar_wait(&cx.tmem_full_bar[buf], ((uint32_t)local >> 1) & 1);
    tcgen05_fence_after();
    if (cx.lane == 0 && cx.lg == 0 && cx.col_off == 0) TL(5, local);
    const uint32_t tacc = cx.tmem_base + (uint32_t)(buf * BN + cx.col_off) + ((uint32_t)(cx.lg * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < NCH; ++c) {
      const int nc = n0 + cx.col_off + c * 32;
      uint32_t r[32];
      tmem_ld_32x32b_x32(tacc + (uint32_t)(c * 32), r);
      // fetch the accumulator-independent operands of the NEXT patch while TMEM is being read
      Pre<MODE> nxt;
      bool nxt_full = false;
      if (c + 1 < NCH) {
        nxt_full = MODE != EPI_GENERIC && (mrow0 + 32 <= M) && (nc + 64 <= N);
        if (nxt_full) prefetch_patch<T, MODE>(ep, mrow0, nc + 32 + c4, rl, nxt);
      } else if (has_next) {
        nxt_full = MODE != EPI_GENERIC && (nm0 + cx.row_off + 32 <= M) && (nn0 + cx.col_off + 32 <= N);
        if (nxt_full) prefetch_patch<T, MODE>(ep, nm0 + cx.row_off, nn0 + cx.col_off + c4, rl, nxt);
      }
      tmem_ld_wait();
      if (cx.lane == 0 && cx.lg == 0 && cx.col_off == 0) TL(8, local * NCH + c);
      if (c == NCH - 1) {
        // every TMEM read of this warp for this tile is done: hand the accumulator back to the leader's MMA warp
        tcgen05_fence_before();
        if (cx.lane == 0) mbar_arrive_cluster(buf ? cx.release_bar1 : cx.release_bar0);
      }
      if (nc < N && mrow0 < M) {   // warp-uniform
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(patch + cx.lane * PATCH_LD + j) =
              make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        __syncwarp();
        if (cx.lane == 0 && cx.lg == 0 && cx.col_off == 0) TL(9, local * NCH + c);
        if (cur_full) {
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (MODE != EPI_ACC && ep.bias) b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + nc + c4));
          // all 8 smem reads first: a generic-pointer global store may alias shared memory as far as the compiler
          // knows, so interleaved LDS / ST would be issued strictly in program order (one row at a time)
          float4 acc[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] = *reinterpret_cast<const float4*>(patch + (i * 4 + rl) * PATCH_LD + c4);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            epi_out4<T, MODE>(ep, mrow0 + i * 4 + rl, nc + c4, acc[i], b4, cur, i, keep, thr16, seed);
        } else {
          epi_patch_scalar<T>(ep, patch, mrow0, nc, M, N, rl, c4);
        }
        __syncwarp();
        if (cx.lane == 0 && cx.lg == 0 && cx.col_off == 0) TL(10, local * NCH + c);
      }
      cur = nxt;
      cur_full = nxt_full;
    }
    m0 = nm0; n0 = nn0; kb0 = nkb0; kb1 = nkb1;
    if (cx.lane == 0 && cx.lg == 0 && cx.col_off == 0) TL(6, local);
  }
}

// MODE: EPI_T16 (alpha, bias), EPI_GELU (bias -> z tile -> GELU [dropout] -> out tile), EPI_DGELU (* gelu'(aux_in) [dropout])
template <typename T, int BN, int MODE>
__device__ __forceinline__ void epilogue_loop_tma(const EpiParams& ep, const Sched2& sc, const EpiCtx& cx,
                                                  const CUtensorMap* tmC, const CUtensorMap* tmZ) {
  constexpr int NCH = BN / 64;
  const int M = cx.M;
  const int lane = cx.lane;
  const unsigned long long seed = ep.drop_p > 0.0f ? eff_seed(ep.drop_seed, ep.drop_seed_ptr) : 0ull;
  const float keep = ep.drop_p > 0.0f ? 1.0f / (1.0f - ep.drop_p) : 1.0f;
  const uint32_t thr16 = drop_thr16(ep.drop_p);
  uint8_t* obuf = reinterpret_cast<uint8_t*>(cx.patch);     // 2 KB out tile
  uint8_t* zbuf = obuf + 2048;                               // 2 KB pre-activation tile (GELU)
  const T* aux = reinterpret_cast<const T*>(ep.aux_in);

  auto load_aux = [&](int m, int nc, uint4 (&a)[4]) {
    if (MODE == EPI_DGELU && m < M) {
      const uint4* src = reinterpret_cast<const uint4*>(aux + (size_t)m * ep.ldaux + nc);
#pragma unroll
      for (int q = 0; q < 4; ++q) a[q] = __ldg(src + q);
    }
  };

  int local = 0;
  int m0 = 0, n0 = 0, kb0 = 0, kb1 = 0;
  uint4 cur_aux[4] = {};
  int t = cx.cluster_id;
  if (t < sc.num_tiles) {
    tile_coords2<BN>(sc, t, m0, n0, kb0, kb1);
    load_aux(m0 + cx.row_off + lane, n0 + cx.col_off, cur_aux);
  }
  for (; t < sc.num_tiles; t += cx.num_clusters, ++local) {
    const int buf = local & 1;
    const int mrow0 = m0 + cx.row_off;
    const int m = mrow0 + lane;
    int nm0 = 0, nn0 = 0, nkb0 = 0, nkb1 = 0;
    const bool has_next = t + cx.num_clusters < sc.num_tiles;
    if (has_next) tile_coords2<BN>(sc, t + cx.num_clusters, nm0, nn0, nkb0, nkb1);
    mbar_wait(&cx.tmem_full_bar[buf], ((uint32_t)local >> 1) & 1);
    tcgen05_fence_after();
    const uint32_t tacc = cx.tmem_base + (uint32_t)(buf * BN + cx.col_off) + ((uint32_t)(cx.lg * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < NCH; ++c) {
      const int nc = n0 + cx.col_off + c * 32;
      uint32_t r[32];
      tmem_ld_32x32b_x32(tacc + (uint32_t)(c * 32), r);
      uint4 nxt_aux[4] = {};
      if (c + 1 < NCH) load_aux(m, nc + 32, nxt_aux);
      else if (has_next) load_aux(nm0 + cx.row_off + lane, nn0 + cx.col_off, nxt_aux);
      tmem_ld_wait();
      if (c == NCH - 1) {
        tcgen05_fence_before();
        if (lane == 0) mbar_arrive_cluster(buf ? cx.release_bar1 : cx.release_bar0);
      }
      // the previous patch's TMA stores must have finished READING the staging tiles before they are rewritten
      if (lane == 0) bulk_wait_read0();
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 4; ++q) {          // 8 columns per 16-byte chunk
        float v[8];
        float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
        if (MODE != EPI_DGELU && ep.bias) {
          b0 = __ldg(reinterpret_cast<const float4*>(ep.bias + nc + q * 8));
          b1 = __ldg(reinterpret_cast<const float4*>(ep.bias + nc + q * 8 + 4));
        }
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaf(__uint_as_float(r[q * 8 + j]), ep.alpha, bb[j]);
        if constexpr (MODE == EPI_GELU) {
          uint4 z;
          z.x = pack2<T>(v[0], v[1]); z.y = pack2<T>(v[2], v[3]); z.z = pack2<T>(v[4], v[5]); z.w = pack2<T>(v[6], v[7]);
          *reinterpret_cast<uint4*>(sw64_chunk(zbuf, lane, q)) = z;
          // GELU of the ROUNDED pre-activation: backward only ever sees the stored 16-bit z
          const uint32_t zw[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = unpack2<T>(zw[j]);
            v[2 * j] = gelu_fast(f.x);
            v[2 * j + 1] = gelu_fast(f.y);
          }
        }
        if constexpr (MODE == EPI_DGELU) {
          const uint32_t aw[4] = {cur_aux[q].x, cur_aux[q].y, cur_aux[q].z, cur_aux[q].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = unpack2<T>(aw[j]);
            v[2 * j] *= dgelu_fast(f.x);
            v[2 * j + 1] *= dgelu_fast(f.y);
          }
        }
        if constexpr (MODE != EPI_T16) {
          if (ep.drop_p > 0.0f) {
            const unsigned long long idx = (unsigned long long)(m + ep.drop_row0) * (unsigned long long)ep.ldc + nc + q * 8;
            float lo4[4] = {v[0], v[1], v[2], v[3]}, hi4[4] = {v[4], v[5], v[6], v[7]};
            drop_apply4(lo4, seed, idx, thr16, keep);
            drop_apply4(hi4, seed, idx + 4, thr16, keep);
#pragma unroll
            for (int j = 0; j < 4; ++j) { v[j] = lo4[j]; v[4 + j] = hi4[j]; }
          }
        }
        uint4 w;
        w.x = pack2<T>(v[0], v[1]); w.y = pack2<T>(v[2], v[3]); w.z = pack2<T>(v[4], v[5]); w.w = pack2<T>(v[6], v[7]);
        *reinterpret_cast<uint4*>(sw64_chunk(obuf, lane, q)) = w;
      }
      fence_proxy_async();        // the generic-proxy writes above become visible to the TMA engine
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(tmC, obuf, nc, mrow0);
        if (MODE == EPI_GELU) tma_store_2d(tmZ, zbuf, nc, mrow0);
        bulk_commit();
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) cur_aux[q] = nxt_aux[q];
    }
    m0 = nm0; n0 = nn0; kb0 = nkb0; kb1 = nkb1;
  }
  if (lane == 0) bulk_wait_all();   // the staging tiles (and this CTA's shared memory) must outlive the last store
  __syncwarp();
}

template <typename T, int BN, int STAGES, bool A_MN, bool B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
gemm_umma2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmC,
                  const __grid_constant__ CUtensorMap tmZ, EpiParams ep, int M, int N, int K, Sched2 sc, int epi_mode) {
  static_assert(BN == 128 || BN == 256, "pair tile is 256 x 128 or 256 x 256");
  using C = Cfg2<BN, STAGES>;
  constexpr int BNH = BN / 2;       // B rows staged by each CTA
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  float* patches = reinterpret_cast<float*>(smem + STAGES * C::STAGE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * C::STAGE_BYTES + C::EPI_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]  (only the leader's copy is used)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) TL(7, 2);
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (sc.kb_wrap > 0) tma_prefetch_desc(&tmB2);
    if (epi_mode >= EPI_T16_TMA) {
      tma_prefetch_desc(&tmC);
      if (epi_mode == EPI_GELU_TMA) tma_prefetch_desc(&tmZ);
    }
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);    // the leader's producer (arrive.expect_tx of both CTAs' bytes)
      mbar_init(&empty_bar[s], 1);   // the multicast tcgen05.commit
    }
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], 2 * EPI_WARPS);   // epilogue warps of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm<2 * BN>(tmem_slot);
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) TL(7, 0);
  pdl_wait();                 // everything above overlapped the previous kernel's tail
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;  // running k-block counter across tiles: stage = it % STAGES
      for (int t = cluster_id; t < sc.num_tiles; t += num_clusters) {
        int m0, n0, kb0, kb1;
        tile_coords2<BN>(sc, t, m0, n0, kb0, kb1);
        const int my_m = m0 + (int)rank * BM;
        const int my_n = n0 + (int)rank * BNH;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * C::STAGE_BYTES);
          const uint32_t fb = mapa_u32(smem_u32(&full_bar[s]), 0);
          uint8_t* sa = smem + s * C::STAGE_BYTES;
          uint8_t* sb = sa + C::A_BYTES;
          const bool lo = sc.kb_wrap > 0 && kb >= sc.kb_wrap;
          const int k0 = (lo ? kb - sc.kb_wrap : kb) * BK;
          const CUtensorMap* mapB = lo ? &tmB2 : &tmB;
          if (!A_MN) {
            tma_load_2d_2sm(sa, &tmA, fb, k0, my_m);  // box {64 k, 128 m}
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)         // boxes {64 m, 64 k}
              tma_load_2d_2sm(sa + j * 8192, &tmA, fb, my_m + 64 * j, k0);
          }
          if (!B_MN) {
            tma_load_2d_2sm(sb, mapB, fb, k0, my_n);  // box {64 k, BNH n}
          } else {
#pragma unroll
            for (int j = 0; j < BNH / 64; ++j)
              tma_load_2d_2sm(sb + j * 8192, mapB, fb, my_n + 64 * j, k0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(UmmaFmt<T>::value, A_MN ? 1 : 0, B_MN ? 1 : 0, 2 * BM, BN);
      uint32_t it = 0;
      int local = 0;
      for (int t = cluster_id; t < sc.num_tiles; t += num_clusters, ++local) {
        int m0, n0, kb0, kb1;
        tile_coords2<BN>(sc, t, m0, n0, kb0, kb1);
        const int buf = local & 1;
        TL(0, local);
        mbar_wait(&tmem_empty_bar[buf], (((uint32_t)local >> 1) & 1) ^ 1);   // both epilogues drained this accumulator
        tcgen05_fence_after();
        TL(1, local);
        const uint32_t tacc = tmem_base + (uint32_t)(buf * BN);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tcgen05_fence_after();
          if (kb == kb0) TL(2, local);
          const uint32_t sa = smem_u32(smem + s * C::STAGE_BYTES);
          const uint32_t sb = sa + C::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UK; ++k) {
            const uint64_t da = A_MN ? make_smem_desc_sw128(sa + k * 2048, 8192, 1024)
                                     : make_smem_desc_sw128(sa + k * 32, 0, 1024);
            const uint64_t db = B_MN ? make_smem_desc_sw128(sb + k * 2048, 8192, 1024)
                                     : make_smem_desc_sw128(sb + k * 32, 0, 1024);
            umma_f16_2sm(tacc, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit_2sm(&empty_bar[s]);  // frees this smem stage in both CTAs
        }
        umma_commit_2sm(&tmem_full_bar[buf]);  // accumulator complete, both CTAs' epilogues may read
        TL(3, local);
      }
    }
  } else {
    const int ew = warp - 2;
    EpiCtx cx;
    cx.tmem_full_bar = tmem_full_bar;
    cx.release_bar0 = mapa_u32(smem_u32(&tmem_empty_bar[0]), 0);
    cx.release_bar1 = mapa_u32(smem_u32(&tmem_empty_bar[1]), 0);
    cx.tmem_base = tmem_base;
    cx.patch = patches + ew * PATCH_FLOATS;
    cx.cluster_id = cluster_id;
    cx.num_clusters = num_clusters;
    cx.lg = warp & 3;                              // TMEM lane group this warp may access
    cx.row_off = (int)rank * BM + cx.lg * 32;
    cx.col_off = (ew >> 2) * (BN / 2);             // which half of the tile's columns
    cx.lane = lane;
    cx.M = M;
    cx.N = N;
    switch (epi_mode) {   // uniform across the grid
      case EPI_ACC: epilogue_loop<T, BN, EPI_ACC>(ep, sc, cx); break;
      case EPI_T16: epilogue_loop<T, BN, EPI_T16>(ep, sc, cx); break;
      case EPI_GELU: epilogue_loop<T, BN, EPI_GELU>(ep, sc, cx); break;
      case EPI_DGELU: epilogue_loop<T, BN, EPI_DGELU>(ep, sc, cx); break;
      case EPI_RES32: epilogue_loop<T, BN, EPI_RES32>(ep, sc, cx); break;
      case EPI_T16_TMA: epilogue_loop_tma<T, BN, EPI_T16>(ep, sc, cx, &tmC, &tmZ); break;
      case EPI_GELU_TMA: epilogue_loop_tma<T, BN, EPI_GELU>(ep, sc, cx, &tmC, &tmZ); break;
      case EPI_DGELU_TMA: epilogue_loop_tma<T, BN, EPI_DGELU>(ep, sc, cx, &tmC, &tmZ); break;
      default: epilogue_loop<T, BN, EPI_GENERIC>(ep, sc, cx); break;
    }
  }

  if (threadIdx.x == 0) TL(7, 1);
  tcgen05_fence_before();
  cluster_sync_all();   // the peer's smem / barriers / TMEM stay valid until both CTAs are done
  if (warp == 1) tmem_dealloc_2sm<2 * BN>(tmem_base);
}

int epi_mode_of(const EpiParams& ep) {
  const bool vec_ok = ((ep.ldc & 3) == 0) && (!ep.res || (ep.ldres & 3) == 0) &&
                      ((!ep.aux_in && !ep.aux_out) || (ep.ldaux & 3) == 0) && (!ep.out2 || (ep.ldc2 & 3) == 0);
  if (!vec_ok) return EPI_GENERIC;
  if (ep.accumulate) return EPI_ACC;
  if (ep.act == GOAT_ACT_NONE && ep.out_f32) return EPI_RES32;
  if (ep.out_f32 || ep.res || ep.out2) return EPI_GENERIC;
  if (ep.act == GOAT_ACT_NONE && ep.drop_p == 0.0f) return EPI_T16;
  if (ep.act == GOAT_ACT_GELU && ep.aux_out) return EPI_GELU;
  if (ep.act == GOAT_ACT_DGELU && ep.aux_in && !ep.bias) return EPI_DGELU;
  return EPI_GENERIC;
}

// 16-bit output modes take the TMA-store epilogue when the output (and the GELU pre-activation copy) can be described
// by a tensor map: 16-byte aligned base, leading dimension a multiple of 8 elements, N a multiple of the 32-column patch
int epi_mode_tma(int mode, const goat_gemm_args& a, const EpiParams& ep) {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("GOAT_GEMM_TMA_STORE");
    on = (e && *e == '0') ? 0 : 1;
  }
  if (!on || (mode != EPI_T16 && mode != EPI_GELU && mode != EPI_DGELU)) return mode;
  if ((a.N & 31) || (ep.ldc & 7) || !aligned16(ep.out)) return mode;
  if (mode == EPI_GELU && ((ep.ldaux & 7) || !aligned16(ep.aux_out))) return mode;
  if (mode == EPI_DGELU && ((ep.ldaux & 7) || !aligned16(ep.aux_in))) return mode;
  return mode == EPI_T16 ? EPI_T16_TMA : mode == EPI_GELU ? EPI_GELU_TMA : EPI_DGELU_TMA;
}

template <typename T, int BN, int STAGES, bool A_MN, bool B_MN>
int launch2(const goat_gemm_args& a, const EpiParams& ep, cudaStream_t stream) {
  using C = Cfg2<BN, STAGES>;
  auto kern = gemm_umma2_kernel<T, BN, STAGES, A_MN, B_MN>;
  static bool configured = false;
  if (!configured) {
    GOAT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    configured = true;
  }
  CUtensorMap tmA, tmB;
  int rc;
  if (!A_MN) rc = make_tmap(&tmA, a.dtype, a.A, a.K, a.M, a.lda, BK, BM);
  else rc = make_tmap(&tmA, a.dtype, a.A, a.M, a.K, a.lda, 64, BK);
  if (rc) return rc;
  if (!B_MN) rc = make_tmap(&tmB, a.dtype, a.B, a.K, a.N, a.ldb, BK, BN / 2);
  else rc = make_tmap(&tmB, a.dtype, a.B, a.N, a.K, a.ldb, 64, BK);
  if (rc) return rc;
  CUtensorMap tmB2 = tmB;
  const bool use_lo = a.B_lo != nullptr && !B_MN && !ep.accumulate;
  if (use_lo && (rc = make_tmap(&tmB2, a.dtype, a.B_lo, a.K, a.N, a.ldb, BK, BN / 2))) return rc;
  const int pairs = num_sms() / 2;
  Sched2 sc;
  sc.tiles_m = (a.M + 2 * BM - 1) / (2 * BM);
  sc.tiles_n = (a.N + BN - 1) / BN;
  sc.num_kb = (a.K + BK - 1) / BK;
  sc.splits = 1;
  const int mn = sc.tiles_m * sc.tiles_n;
  if (ep.accumulate && mn < pairs) {
    // split K (at least 4 k-blocks = 256 of K per split) so that the tile count fills whole waves of the 74 pairs:
    // cost(s) = waves(mn * s) / s in units of the unsplit tile time.  (Rounding the split count UP to "about one wave"
    // put e.g. 9 x 9 = 81 tiles on 74 pairs: two waves for the work of 1.1.)
    int cap = sc.num_kb / 4;
    if (cap < 1) cap = 1;
    if (cap > 32) cap = 32;
    int best = 1;
    double best_cost = 1e30;
    for (int s = 1; s <= cap; ++s) {
      const int waves = (mn * s + pairs - 1) / pairs;
      const double cost = (double)waves / s + 0.02 * s;    // small penalty per split: extra atomics traffic
      if (cost < best_cost - 1e-9) { best_cost = cost; best = s; }
    }
    sc.splits = best;
  }
  sc.kb_wrap = 0;
  if (use_lo) {            // second pass of the K loop over A against B_lo (never combined with split-K)
    sc.kb_wrap = sc.num_kb;
    sc.num_kb *= 2;
  }
  sc.kb_per_split = (sc.num_kb + sc.splits - 1) / sc.splits;
  sc.splits = (sc.num_kb + sc.kb_per_split - 1) / sc.kb_per_split;   // no empty splits
  sc.num_tiles = mn * sc.splits;
  const int clusters = sc.num_tiles < pairs ? sc.num_tiles : pairs;
  const int mode = epi_mode_tma(epi_mode_of(ep), a, ep);
  CUtensorMap tmC = tmA, tmZ = tmA;
  if (mode >= EPI_T16_TMA) {
    if ((rc = make_tmap_sw64(&tmC, a.dtype, ep.out, a.N, a.M, ep.ldc, 32, 32))) return rc;
    if (mode == EPI_GELU_TMA && (rc = make_tmap_sw64(&tmZ, a.dtype, ep.aux_out, a.N, a.M, ep.ldaux, 32, 32))) return rc;
  }
  GOAT_CUDA(launch_pdl(kern, dim3(2 * clusters), dim3(THREADS), C::SMEM_BYTES, stream, tmA, tmB, tmB2, tmC, tmZ, ep, a.M, a.N,
                       a.K, sc, mode));
  return GOAT_OK;
}

template <typename T, int BN, int STAGES>
int dispatch2(const goat_gemm_args& a, const EpiParams& ep, cudaStream_t stream) {
  if (!a.a_mn_major && !a.b_mn_major) return launch2<T, BN, STAGES, false, false>(a, ep, stream);
  if (!a.a_mn_major && a.b_mn_major) return launch2<T, BN, STAGES, false, true>(a, ep, stream);
  if (a.a_mn_major && a.b_mn_major) return launch2<T, BN, STAGES, true, true>(a, ep, stream);
  return launch2<T, BN, STAGES, true, false>(a, ep, stream);
}

}  // namespace

#ifdef GOAT_TIMELINE
extern "C" int goat_debug_timeline(long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, g_timeline, sizeof(long long) * 4096);
}
#endif

// block_n: 128 or 256
int gemm_umma2(const goat_gemm_args& a, const EpiParams& ep, int block_n, cudaStream_t stream) {
  if (block_n == 256) {
    if (a.dtype == GOAT_F16) return dispatch2<__half, 256, 5>(a, ep, stream);
    return dispatch2<__nv_bfloat16, 256, 5>(a, ep, stream);
  }
  if (a.dtype == GOAT_F16) return dispatch2<__half, 128, 6>(a, ep, stream);
  return dispatch2<__nv_bfloat16, 128, 6>(a, ep, stream);
}

}  // namespace goat
