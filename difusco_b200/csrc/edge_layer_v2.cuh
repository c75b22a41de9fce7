// Fused GNN edge layer, round-2 kernel: CTA pairs (tcgen05 cta_group::2), software-pipelined over tiles.
//
//   e_hat = C e + A h[col] + B h[row]                         gnn_encoder.py:104,110
//   agg  += sigmoid(e_hat) * V h[col]   (row-segment sums)     :112,163,177-191
//   e_til = relu(LN_e(e_hat)) + tau                            :131,135,445
//   e     = e + O silu(LN_O(e_til)) + b_O   (in place)         :449, :339-347
//
// What changed against edge_layer_tc.cuh (which stays for the MIS last layer, graphs with more than MAXSEG node segments
// per 32-edge group, the linear mode and as an A/B reference):
//  * Two CTAs of a cluster work as ONE tensor-core unit: M = 256 (one 128-edge tile per CTA), N = 256, the weight
//    operand B is split by output channel between the two CTAs (each CTA streams only HALF of C and O: the
//    L2 -> SM weight traffic and the shared-memory footprint of B halve).  MMAs are issued by the leader CTA only;
//    completion is multicast to both CTAs' mbarriers.
//  * The fp32 edge tile comes in by TMA (32-column boxes, 128B swizzle) into the operand ring and is converted
//    IN PLACE and row-locally to the bf16 K-major operand (one 128B-swizzled tile per box: 32 hi | 32 lo values per row):
//    the round-1 kernel's register-staged global loads (30 % of all warp samples, long-scoreboard) are gone.
//  * The operand rings (NA x 16 KB A, NB x 16 KB B) do not alias the gather / epilogue staging, so the phases of
//    neighbouring tiles overlap: conversion + GEMM1 of tile t+1 run under E4 of tile t ("X phase").  The residual is
//    preloaded into GEMM2's accumulator (acc2 = e_in + b_O, GEMM2 accumulates), so E4 is a plain copy-out: every worker
//    warp stages its 32 rows of a result box in its own gather buffers and issues the TMA store itself.
//  * E1 takes all of its inputs (A h[col], V h[col], B h[row]) from warp-private cp.async staging that is
//    filled two 8-column steps ahead; the segment-reduce patch reuses the staging buffer.
//  * Layer 0 (input rows are one of two table rows) runs in LUT mode: no GEMM1, no input stream.  The last layer of the
//    sparse TSP encoder accumulates the head's GroupNorm partial sums in its E4 (MODE_GN).
//
// Warp roles (640 threads): warp 0 weight TMA, warp 1 MMA issue + TMEM owner, warp 2 edge-index prefetch + input-box
// TMA, warp 3 idle, warps 4..19 row workers (thread == edge row == TMEM lane, the 256 channels of a row split over the
// 4 warps of a lane quarter).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "edge_layer_tc.cuh"

namespace dfb {
namespace v2 {

#ifndef DFB_NA
#define DFB_NA 5
#define DFB_NB 4
#endif
// The kernel launches with 96 registers per thread (640 threads); the service warpgroup (warps 0..3) then shrinks to 32 and each
// of the four worker warpgroups grows to 112 (setmaxnreg, 128 * 32 + 512 * 112 = 640 * 96): no spills in the worker code.
constexpr int REGS_SERVICE = 32, REGS_WORKER = 112;
constexpr int NA = DFB_NA, NB = DFB_NB;   // ring depths; NA + NB = 9 stages of 16 KB (tuning: -DDFB_NA=6 -DDFB_NB=3 measured no faster)
constexpr int STAGE = 16384;                       // one fp32 box [128 rows x 32 cols] == bf16 hi (8 KB) | lo (8 KB)
constexpr int HALF = 8192;
constexpr int NWORKW = 16;                         // worker warps
constexpr int NSERV = 4;                           // service warps
constexpr int THREADS = (NSERV + NWORKW) * 32;     // 640
constexpr int NWORK = NWORKW * 32;                 // 512
constexpr int MAXSEG = 8;                          // node segments per 32-edge group the kernel stages B h[row] slots for
constexpr int GBUF = 2048 + MAXSEG * 32;           // gather buffer: [32 rows][8 A | 8 V] fp32 + B-row slots x 32 B
constexpr int OFF_A = 0;
constexpr int OFF_B = OFF_A + NA * STAGE;
constexpr int OFF_G = OFF_B + NB * STAGE;
constexpr int OFF_PRM = OFF_G + NWORKW * 2 * GBUF;  // ln_e_g, ln_e_b, tau, ln_o_g, ln_o_b, b_O
constexpr int OFF_IDX = OFF_PRM + 6 * H * 4;        // 2 buffers x { row[128], col[128] }
constexpr int IDX_INTS = 2 * TC_TILE + 8;            // row[128] | col[128] | grp_first[4] | grp_pair[4]
constexpr int OFF_GN = OFF_IDX + 2 * IDX_INTS * 4;   // per worker warp: 8 doubles (4 groups x {sum, sum of squares}) x 2 boxes
constexpr int OFF_SEG = OFF_GN + NWORKW * 16 * 8;      // per worker warp: node of each of its (<= MAXSEG) segments
constexpr int OFF_BAR = OFF_SEG + NWORKW * MAXSEG * 4;
constexpr int SMEM_BYTES = OFF_BAR + 52 * 8;
static_assert(NA <= 6 && NB <= 6, "barrier slots");
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
static_assert(GBUF % 128 == 0, "gather buffers stay 128-byte aligned");
static_assert(OFF_G % 1024 == 0 && (2 * GBUF) % 512 == 0 && 2 * GBUF >= 4096 + 512,
              "every worker warp's gather buffers contain a 1024-byte aligned 4 KB window (its E4 staging, 128B swizzle)");

// UMMA instruction descriptor: D=F32, A=B=BF16, K-major, N=256, M=256 (the pair), cute::UMMA::InstrDescriptor
constexpr uint32_t IDESC2 = (1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((256u >> 4) << 24);
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address (cute Sm100MmaPeerBitMask)

struct Params {
  float* e;
  const float* uvab;
  float* partials;
  GraphDev g;
  LayerParams lp;
  const float* tvec;      // [256] time vector added on edges (TSP) or nullptr (MIS)
  const float* zero_row;  // [1024] zeros: gather source of rows past the end of the edge list
  // LUT mode (layer 0: the input rows are one of two table rows): per-edge selector in caller edge order (or
  // null: always row 0, the MIS e0 = 0 case), cl = C * lut (GEMM1 by table lookup), lut = the residual rows
  const float* lut_x;
  const float* cl;        // [2][256]
  const float* lut;       // [2][256]
  float* debug_acc;       // tests: dump the GEMM1 accumulator [E][256] and stop
  double* gn_part;        // last layer (sparse TSP head): per-(CTA, lane quarter) GroupNorm(32) partial sums [blocks][32][2], or null
  int E;                  // number of valid edge rows (GroupNorm statistics skip the padding rows)
  int* error_flag;
  unsigned long long* phase_cycles;
  int agg_mode;
  int w_row_base;         // row of this layer's C_hi block in the bf16 weight arena
  int n_tiles;
  int probe;
  int reverse;            // sweep the tile pairs from the end of the edge stream to its start
};

// ----------------------------------------------------------------------------------------------
// PTX wrappers specific to the pair kernel
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the LEADER CTA's copy of a barrier, from either CTA of the pair (cutlass umma_arrive_2x1SM_sm0).  No
// cluster-scope release: ptxas turns that into MEMBAR.ALL.GPU + ERRBAR (8 % of all warp samples in the first version);
// what is handed over is either TMEM contents (ordered by tcgen05.wait::st + tcgen05.fence::before_thread_sync) or
// shared memory written for the async proxy (ordered by fence.proxy.async), exactly as in CUTLASS' 2-SM kernels.
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_MASK) : "memory");
}
#ifndef DFB_WAIT_HINT_NS
#define DFB_WAIT_HINT_NS 20000u
#endif
#ifndef DFB_EXP
#define DFB_EXP 0   // timing experiments (results wrong): 2 = no result stores, 4 = stores never awaited, 8 = every CTA stores to its own first tile (L2-resident)
#endif
// bounded wait on a barrier that the peer CTA also arrives on
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity, int* error_flag, int code) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok = 0;
#pragma unroll 1
  for (uint32_t spin = 0;; ++spin) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(addr), "r"(parity), "r"(DFB_WAIT_HINT_NS)
        : "memory");
    if (ok) return;
    if (spin > 400000u) {
      if (error_flag) {
        error_flag[1] = (int)blockIdx.x;
        error_flag[2] = (int)parity;
        error_flag[3] = (int)threadIdx.x;
        atomicExch(error_flag, code);
      }
      __threadfence_system();
      __trap();
    }
  }
}
// L2 eviction policies: the edge stream passes through L2 once per layer (evict_first), the node tensor uvab (32 MB at C2)
// and the weights (1 MB) are re-read by every tile (evict_last) - without the hints the 820 MB/launch stream pushes
// part of uvab out of the 126 MB L2 (ncu: +117 MB of DRAM reads per launch).
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// TMA load whose completion is booked on the leader CTA's barrier (cute SM100_TMA_2SM_LOAD_2D)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d_hint(const CUtensorMap* map, uint32_t src, int c0, int c1, uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void cp_async16_hint(uint32_t dst_smem, const void* src, uint64_t pol) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "l"(pol) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
      " tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(IDESC2), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
      " tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(IDESC2), "r"(accumulate)
      : "memory");
}
// completion of all prior MMAs of this thread -> the barrier at the same offset in BOTH CTAs
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"((uint16_t)3)
      : "memory");
}
// K-major, 64-byte swizzle, 8-row atoms 512 bytes apart (cute::UMMA::SmemDescriptor, layout type 4)
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
  const uint64_t hi = 32ull | (1ull << 14) | (4ull << 29);
  return (hi << 32) | (1ull << 16) | (uint64_t)((smem_addr >> 4) & 0x3fffu);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%8], {%0, %1, %2, %3, %4, %5, %6, %7};"
               ::"r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(taddr)
               : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_read_n() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// byte offset of (row r, 16-byte unit j in [0,4)) inside a [rows][32 bf16] K-major 64B-swizzled tile
__device__ __forceinline__ uint32_t sw64_off(int r, int j) {
  return (uint32_t)((r >> 3) * 512 + (r & 7) * 64 + ((j ^ ((r >> 1) & 3)) << 4));
}

// ----------------------------------------------------------------------------------------------
// MAXAGG: --aggregation max (gnn_encoder.py:188-191): invalid rows contribute -inf and every segment goes through the
// general masked reduce; the sum / mean instantiation (the reference default) carries none of that code.
// GNSTATS: the last layer of the sparse TSP encoder also leaves the head's GroupNorm partial sums (own instantiation: the
// extra live values would otherwise cost the other eleven layers registers in the X phase).
enum { MODE_PLAIN = 0, MODE_GN = 1, MODE_LUT = 2 };
#ifdef DFB_PHASE_PROF
// Tuning build only: clock64 time line of cluster 0's leader CTA for tiles 3..6 (slot = (it - 3) * 128 + actor * 16 + event;
// actors 0..3 = worker warp wq 0 of part 0..3, 4 = MMA thread, 5 = box loads, 6 = stores), layer 5 only.  scripts/probe_tc.py prints it.
__device__ long long g_pair_trace[4 * 128];
#define TRACE(actor, ev, it_) do { if ((P.probe & 128) && blockIdx.x == 0 && P.w_row_base == 5 * 12 * H && (it_) >= 3 && (it_) < 7) \
    g_pair_trace[((it_) - 3) * 128 + (actor) * 16 + (ev)] = clock64(); } while (0)
#else
#define TRACE(actor, ev, it_) do { } while (0)
#endif
template <bool MAXAGG, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
k_edge_layer_pair(const __grid_constant__ CUtensorMap wmap, const __grid_constant__ CUtensorMap emap,
                  const __grid_constant__ CUtensorMap smap, const Params P) {
  constexpr bool GNSTATS = MODE == MODE_GN;
  constexpr bool LUT = MODE == MODE_LUT;   // no GEMM1, no input boxes: acc1 and the residual come from two-row tables
  extern __shared__ __align__(1024) unsigned char smem[];
  float* prm = reinterpret_cast<float*>(smem + OFF_PRM);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  // Every barrier has waiters that see EVERY one of its phases in order (a parity wait that runs two phases ahead of the
  // barrier aliases with the phase before): box_full is indexed by the box of the tile, not by the ring stage, because
  // box b is always converted by the same part (b & 3).
  uint64_t* box_full = bars;        // [8]  TMA input box b of the tile landed (local, tx)
  uint64_t* a_full = bars + 8;      // [NA] LEADER: stage converted by both CTAs (4 warps each)
  uint64_t* a_empty = bars + 14;    // [NA] MMA commit (multicast): stage consumed
  uint64_t* b_full = bars + 20;     // [NB] LEADER: both weight halves landed (tx)
  uint64_t* b_empty = bars + 26;    // [NB] MMA commit (multicast)
  uint64_t* acc_rdy = bars + 32;    // [2]  MMA commit (multicast): GEMM1 / GEMM2 accumulator complete
  uint64_t* a2_full = bars + 34;    // [4]  LEADER: GEMM2 A chunk (64 columns of s) written to TMEM by both CTAs
  // bars + 38 .. 45: unused (the result stores are issued and awaited by the worker warps themselves)
  uint64_t* idx_full = bars + 46;   // [2]  edge endpoints of a tile in shared memory (32 lanes of warp 2)
  uint64_t* idx_free = bars + 48;   // [2]  the 16 worker warps are done with them
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 50);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int n_clusters = gridDim.x >> 1, cid = blockIdx.x >> 1;
  const int n_pairs = (P.n_tiles + 1) >> 1;
  const int n_my = (cid < n_pairs) ? (n_pairs - cid + n_clusters - 1) / n_clusters : 0;
  const bool debug = P.debug_acc != nullptr;
  const uint32_t smem_base = smem_u32(smem);
  // tile pairs in ascending order, or descending when P.reverse: consecutive layers sweep the edge stream in opposite
  // directions, so a layer starts on the rows the previous one wrote last (partly still in L2: -8 % DRAM traffic on every
  // second layer, ncu --cache-control none; storing the tail of a sweep without the evict_first hint changed nothing)
  auto tile_of = [&](int it) {
    const int q = cid + it * n_clusters;
    return 2 * (P.reverse ? n_pairs - 1 - q : q) + (int)rank;
  };

  if (threadIdx.x == 0) {
    if (smem_base & 1023u) {   // the operand swizzles assume a 1024-byte aligned base
      if (P.error_flag) atomicExch(P.error_flag, 99);
      __trap();
    }
    for (int i = 0; i < 8; ++i) mbar_init(&box_full[i], 1);
    for (int i = 0; i < NA; ++i) { mbar_init(&a_full[i], 8); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < NB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    mbar_init(&acc_rdy[0], 1); mbar_init(&acc_rdy[1], 1);
    for (int i = 0; i < 4; ++i) mbar_init(&a2_full[i], 2 * NWORKW);
    for (int i = 0; i < 2; ++i) { mbar_init(&idx_full[i], 32); mbar_init(&idx_free[i], NWORKW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < H; i += THREADS) {
    prm[i] = P.lp.ln_e_g[i];
    prm[H + i] = P.lp.ln_e_b[i];
    prm[2 * H + i] = P.tvec ? P.tvec[i] : 0.0f;
    prm[3 * H + i] = P.lp.ln_o_g[i];
    prm[4 * H + i] = P.lp.ln_o_b[i];
    prm[5 * H + i] = P.lp.b_O[i];
  }
  for (int i = threadIdx.x; i < NWORKW * 16; i += THREADS) reinterpret_cast<double*>(smem + OFF_GN)[i] = 0.0;
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // both CTAs' barriers are initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================== weight TMA (both CTAs, each its N-half) =====================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_SERVICE));
    if (lane == 0) {
      const int i_lo = LUT ? 8 : 0, n_use = debug ? 8 : 16;
      const uint64_t pol_keep = l2_policy_evict_last();
      uint32_t u = 0;
      for (int it = 0; it < n_my; ++it) {
        for (int i = i_lo; i < n_use; ++i, ++u) {
          const uint32_t sb = u % NB, k = u / NB;
          mbar_wait(&b_empty[sb], (k & 1) ^ 1, P.error_flag, 1);
          if (leader) mbar_arrive_expect_tx(&b_full[sb], 2 * STAGE);
          const uint32_t dst = smem_base + OFF_B + sb * STAGE;
          // arena rows of a layer: C_hi | C_lo | O_hi | O_lo (256 each); this CTA's output channels are rows rank*128..+127
          const int rb = P.w_row_base + (i < 8 ? 0 : 512) + (int)rank * 128, kc = (i & 7) * 32;
          tma_load_2d_pair(dst, &wmap, &b_full[sb], kc, rb, pol_keep);
          tma_load_2d_pair(dst + HALF, &wmap, &b_full[sb], kc, rb + 256, pol_keep);
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issue (leader CTA only) =====================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_SERVICE));
    if (leader && lane == 0) {
      uint32_t ub = 0, ga = 0;
#ifdef DFB_PHASE_PROF
      long long wa = 0, wb = 0, w2 = 0, tq;
#define MMA_T0() do { tq = clock64(); } while (0)
#define MMA_T1(acc) do { acc += clock64() - tq; } while (0)
#else
#define MMA_T0() do { } while (0)
#define MMA_T1(acc) do { } while (0)
#endif
      for (int it = 0; it < n_my; ++it) {
        // ---- GEMM1: acc1 = e C^T, 8 K-chunks of 32, A and B from shared memory (LUT mode: the workers fill acc1) ----
        for (int kc = 0; kc < (LUT ? 0 : 8); ++kc, ++ub, ++ga) {
          const uint32_t sb = ub % NB, sa = ga % NA;
          MMA_T0();
          mbar_wait(&b_full[sb], (ub / NB) & 1, P.error_flag, 2);
          MMA_T1(wb);
          MMA_T0();
          mbar_wait_cluster(&a_full[sa], (ga / NA) & 1, P.error_flag, 3);
          MMA_T1(wa);
          tc_fence_after();
          TRACE(4, kc, it);
          const uint32_t a_hi = smem_base + OFF_A + sa * STAGE, a_lo = a_hi + 64;   // hi | lo halves of each 128-byte row
          const uint32_t b_hi = smem_base + OFF_B + sb * STAGE, b_lo = b_hi + HALF;
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint64_t dah = umma_desc_sw128(a_hi + ks * 32), dal = umma_desc_sw128(a_lo + ks * 32);
            const uint64_t dbh = umma_desc_sw64(b_hi + ks * 32), dbl = umma_desc_sw64(b_lo + ks * 32);
            umma2_bf16(tmem_base, dah, dbh, (kc | ks) ? 1u : 0u);
            umma2_bf16(tmem_base, dal, dbh, 1u);
            umma2_bf16(tmem_base, dah, dbl, 1u);
          }
          umma2_commit(&a_empty[sa]);
          umma2_commit(&b_empty[sb]);
        }
        if (!LUT) umma2_commit(&acc_rdy[0]);
        if (debug) continue;
        // ---- GEMM2: acc2 += s O^T on top of the preloaded residual e + b_O; A operand (bf16 hi/lo of s) in TMEM:
        //      k-step j at columns 16 j (8 hi + 8 lo) ----
        for (int kc = 0; kc < 8; ++kc, ++ub) {
          const uint32_t sb = ub % NB;
          MMA_T0();
          if ((kc & 1) == 0) mbar_wait_cluster(&a2_full[kc >> 1], it & 1, P.error_flag, 13);
          MMA_T1(w2);
          if ((kc & 1) == 0) TRACE(4, 8 + (kc >> 1), it);
          MMA_T0();
          mbar_wait(&b_full[sb], (ub / NB) & 1, P.error_flag, 12);
          MMA_T1(wb);
          tc_fence_after();
          const uint32_t b_hi = smem_base + OFF_B + sb * STAGE, b_lo = b_hi + HALF;
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint32_t ta_hi = tmem_base + (uint32_t)(kc * 32 + ks * 16), ta_lo = ta_hi + 8;
            const uint64_t dbh = umma_desc_sw64(b_hi + ks * 32), dbl = umma_desc_sw64(b_lo + ks * 32);
            umma2_bf16_ts(tmem_base + 256u, ta_hi, dbh, 1u);
            umma2_bf16_ts(tmem_base + 256u, ta_lo, dbh, 1u);
            umma2_bf16_ts(tmem_base + 256u, ta_hi, dbl, 1u);
          }
          umma2_commit(&b_empty[sb]);
        }
        umma2_commit(&acc_rdy[1]);
        TRACE(4, 12, it);
      }
#ifdef DFB_PHASE_PROF
      if (P.probe & 128) {   // MMA thread's waits: [21] A stages (GEMM1), [22] weights, [23] GEMM2 A chunks; per tile PAIR
        atomicAdd(P.phase_cycles + 21, (unsigned long long)wa);
        atomicAdd(P.phase_cycles + 22, (unsigned long long)wb);
        atomicAdd(P.phase_cycles + 23, (unsigned long long)w2);
      }
#endif
#undef MMA_T0
#undef MMA_T1
    }
  } else if (warp == 2) {
    // ===================================== edge endpoints + input boxes =====================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_SERVICE));
    uint32_t ga = 0;
    const uint64_t pol_stream = l2_policy_evict_first();
    for (int it = 0; it < n_my; ++it) {
      const int tile = tile_of(it), ib = it & 1;
      if (it >= 2) mbar_wait(&idx_free[ib], ((it >> 1) - 1) & 1, P.error_flag, 20);
      int* s_row = reinterpret_cast<int*>(smem + OFF_IDX) + ib * IDX_INTS;
      int* s_col = s_row + TC_TILE;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int rr = j * 32 + lane, s_edge = tile * TC_TILE + rr;
        const bool ok = tile < P.n_tiles && s_edge < P.g.E;
        s_row[rr] = ok ? __ldg(P.g.row + s_edge) : -1;
        int cv = ok ? __ldg(P.g.col + s_edge) : -1;
        if (LUT && ok && P.lut_x)   // table-row selector of the edge rides in bit 30 of its column index
          cv |= (__ldg(P.lut_x + (P.g.perm ? __ldg(P.g.perm + s_edge) : s_edge)) != 0.0f) ? (1 << 30) : 0;
        s_col[rr] = cv;
      }
      if (lane < 8) {   // (32-edge group, node) pair bookkeeping of the tile's four groups
        const int grp = tile * 4 + (lane & 3);
        const int* src = (lane < 4) ? P.g.grp_first : P.g.grp_pair;
        s_row[2 * TC_TILE + lane] = (grp < P.g.n_groups) ? __ldg(src + grp) : 0;
      }
      mbar_arrive(&idx_full[ib]);
      if (lane == 0) {
        if (!LUT && it + 1 < n_my && tile_of(it + 1) < P.n_tiles && !(P.probe & 256)) {
          // the next tile's 128 edge rows are one contiguous 128 KB block: pull it into L2 now
          const float* nxt = P.e + (size_t)tile_of(it + 1) * TC_TILE * H;
          asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(nxt), "r"(TC_TILE * H * 4), "l"(pol_stream)
                       : "memory");
        }
        for (int b = 0; b < (LUT ? 0 : 8); ++b, ++ga) {
          const uint32_t sa = ga % NA;
          mbar_wait(&a_empty[sa], ((ga / NA) & 1) ^ 1, P.error_flag, 4);
          mbar_arrive_expect_tx(&box_full[b], STAGE);
          TRACE(5, b, it);
          tma_load_2d_hint(smem_base + OFF_A + sa * STAGE, &emap, &box_full[b], 32 * b, tile * TC_TILE, pol_stream);
        }
      }
      __syncwarp();
    }
  } else if (warp == 3) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_SERVICE));
    // (no role: the workers issue their own result stores; the warp keeps the worker warps' ids congruent to their TMEM
    //  lane quarter)
  } else {
    // ===================================== row workers =====================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_WORKER));
    const int wq = warp & 3;                 // TMEM lane quarter this warp may access (hardware rule: warp id % 4)
    const int ww = warp - NSERV;             // 0..15 == part * 4 + wq
    const int part = ww >> 2;                // which 64-column slice of the row
    const int r = wq * 32 + lane;            // tile row == TMEM lane
    const int cbase = part * 64;
    unsigned char* gbuf0 = smem + OFF_G + ww * 2 * GBUF;
    // E4 staging of this warp's 32 rows of a result box: [32 rows][32 fp32], 128B swizzle, inside the warp's own gather
    // buffers (the first 1024-byte aligned window: the region starts at a multiple of 512).  The warp stores it itself
    // (lane 0: TMA store + bulk group) and waits for its own bulk-group read before it reuses the window - no hop
    // through another warp, no barrier shared with the other warps of the part.
    unsigned char* stagewin = gbuf0 + ((1024u - (smem_u32(gbuf0) & 1023u)) & 1023u);
    const uint64_t pol_stream = l2_policy_evict_first();
    const uint32_t t_acc1 = tmem_base + ((uint32_t)(wq * 32) << 16);
    const uint32_t t_acc2 = t_acc1 + 256u;
    auto worker_bar = [] { asm volatile("bar.sync 1, %0;" ::"n"(NWORK) : "memory"); };
    auto stat_buf = [&](int p2) { return reinterpret_cast<float*>(smem + OFF_G + (p2 * 4 + wq) * 2 * GBUF); };
#ifdef DFB_PHASE_PROF
    const bool prof = (P.probe & 128) && ww == 0 && lane == 0;
    long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tp = 0, px[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tx = 0;
#define PHASE(i) do { if (prof) { long long _n = clock64(); pc[i] += _n - tp; tp = _n; } } while (0)
#define XSUB(i) do { if (prof) { long long _n = clock64(); px[i] += _n - tx; tx = _n; } } while (0)
#else
#define PHASE(i) do { } while (0)
#define XSUB(i) do { } while (0)
#endif
    uint32_t gbox = 0;   // global input-box counter of this CTA (operand ring position)
#ifdef DFB_PHASE_PROF
#define WTR(ev) do { if (wq == 0 && lane == 0) TRACE(part, ev, it); } while (0)
#else
#define WTR(ev) do { } while (0)
#endif
    const uint64_t pol_keep = l2_policy_evict_last();
    bool prev_valid = false;   // this thread's row of the previous tile is a real edge (GroupNorm statistics)
    for (int it = 0; it <= n_my; ++it) {
      const bool have_tile = it < n_my;      // X phase converts tile `it` and finishes (E4) tile `it - 1`
      const int tile = tile_of(it);
#ifdef DFB_PHASE_PROF
      if (prof) tp = clock64();
#endif
      // No barrier here: the first write into the gather buffers / staging of this iteration (E4 of the previous tile) waits
      // for GEMM2, whose last K-chunk needs every worker warp's arrival after its last read of the LayerNorm exchange

      // ================= X phase: this part's two boxes (part, part + 4): convert tile `it`, finish tile `it - 1` =========
#ifdef DFB_PHASE_PROF
      if (prof) tx = clock64();
#endif
      WTR(0);
      int xsel = 0;   // LUT mode: table row of this thread's edge
      if (LUT && have_tile) {
        mbar_wait(&idx_full[it & 1], (it >> 1) & 1, P.error_flag, 22);
        const int cv = (reinterpret_cast<const int*>(smem + OFF_IDX) + (it & 1) * IDX_INTS + TC_TILE)[r];
        xsel = (cv >= 0) ? ((cv >> 30) & 1) : 0;
      }
      // Per box: convert (GEMM2's tail of the previous tile drains meanwhile), copy the previous tile's result box out (E4),
      // preload the residual (the box's fp32 values stay in registers across the E4).  Tried and slower: E4 of box `part`
      // first (the workers then idle for the whole GEMM2 tail, ~3.5k cycles including the barrier wake-ups).
      auto load_lut = [&](int b, float4 (&xin)[8]) {
#pragma unroll
        for (int q = 0; q < 8; ++q) xin[q] = __ldg(reinterpret_cast<const float4*>(P.lut + xsel * H + 32 * b) + q);
      };
      auto convert_box = [&](int j, float4 (&xin)[8]) {
        const int b = part + 4 * j;
        const uint32_t g = gbox + b, sa = g % NA;
        mbar_wait(&box_full[b], it & 1, P.error_flag, 5);
        XSUB(0);   // wait for the input box
        WTR(1 + 5 * j);
        unsigned char* stage = smem + OFF_A + sa * STAGE;
#pragma unroll
        for (int q = 0; q < 8; ++q) xin[q] = *reinterpret_cast<const float4*>(stage + sw128_off(r, q));
        // in place and row-local: the A operand is ONE 128B-swizzled K-major tile whose 64 bf16 per row are the box's
        // 32 hi values followed by its 32 lo values, so a row's 128 operand bytes replace exactly its own 128 fp32 bytes
        // (same swizzle as the TMA box) - no other thread's row is touched, no barrier (the fp32 values stay live in
        // registers: they are also the residual preloaded below)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint2 h0, l0, h1, l1;
          split4(xin[2 * q], h0, l0);
          split4(xin[2 * q + 1], h1, l1);
          *reinterpret_cast<uint4*>(stage + sw128_off(r, q)) = make_uint4(h0.x, h0.y, h1.x, h1.y);
          *reinterpret_cast<uint4*>(stage + sw128_off(r, 4 + q)) = make_uint4(l0.x, l0.y, l1.x, l1.y);
        }
        fence_proxy_async();
        tc_fence_before();   // orders this thread's earlier TMEM accesses (previous tile) before the MMA overwrites acc1
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&a_full[sa]);
        XSUB(1);   // conversion
        WTR(2 + 5 * j);
      };
      auto e4_box = [&](int j) {
        // E4 of the previous tile for box b: acc2 already holds e_in + b_O + s O^T (the residual was preloaded)
        const int b = part + 4 * j;
        if (j == 0) {
          mbar_wait(&acc_rdy[1], (it - 1) & 1, P.error_flag, 7);
          tc_fence_after();
        } else {
          if (lane == 0 && !(DFB_EXP & 4)) tma_store_wait_read_n<0>();   // this warp's store of box `part` has read the window
          __syncwarp();
        }
        XSUB(2);   // wait for GEMM2 / the staging
        WTR(3 + 5 * j);
        float gs[8];   // last layer: this row's sums / sums of squares of the 4 GroupNorm groups (8 channels each) of the box
#pragma unroll
        for (int g2 = 0; g2 < 4; ++g2) {   // 8 columns == one GroupNorm group at a time: few live registers next to xin
          uint32_t v[8];
          tmem_ld8(t_acc2 + 32 * b + 8 * g2, v);
          tmem_wait_ld();
          *reinterpret_cast<float4*>(stagewin + sw128_off(lane, 2 * g2)) =
              make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3]));
          *reinterpret_cast<float4*>(stagewin + sw128_off(lane, 2 * g2 + 1)) =
              make_float4(__uint_as_float(v[4]), __uint_as_float(v[5]), __uint_as_float(v[6]), __uint_as_float(v[7]));
          if (GNSTATS) {
            float sv = 0.f, qv = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float xv = __uint_as_float(v[i]);
              sv += xv;
              qv = fmaf(xv, xv, qv);
            }
            gs[2 * g2] = prev_valid ? sv : 0.f;
            gs[2 * g2 + 1] = prev_valid ? qv : 0.f;
          }
        }
        if (GNSTATS) {
          // 8 values x 32 rows -> lane k (k < 8) holds the warp total of value k (butterfly transpose-reduce: 9 shuffles),
          // accumulated in fp64 per warp: the statistics span all E edges of the call (gnn_encoder.py:400, batch dim 1)
#pragma unroll
          for (int off = 4; off >= 1; off >>= 1) {
            const bool up = (lane & off) != 0;
#pragma unroll
            for (int i = 0; i < off; ++i) {
              const float send = up ? gs[i] : gs[i + off], keep = up ? gs[i + off] : gs[i];
              gs[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
          }
          float tot = gs[0];
          tot += __shfl_xor_sync(0xffffffffu, tot, 8);
          tot += __shfl_xor_sync(0xffffffffu, tot, 16);
          // lane L holds the warp total of value index L & 7 (= 2 * group + {sum, sum of squares})
          if (lane < 8) reinterpret_cast<double*>(smem + OFF_GN)[ww * 16 + j * 8 + lane] += (double)tot;
        }
        fence_proxy_async();   // generic-proxy writes -> visible to the TMA store
        __syncwarp();
        if (lane == 0 && !(DFB_EXP & 2)) {
          tma_store_2d_hint(&smap, smem_u32(stagewin), 32 * b,
                            ((DFB_EXP & 8) ? (int)blockIdx.x : tile_of(it - 1)) * TC_TILE + wq * 32, pol_stream);
          tma_store_commit();
        }
        XSUB(3);   // E4 copy-out
        WTR(4 + 5 * j);
      };
      auto preload_box = [&](int j, const float4 (&xin)[8]) {
        // preload GEMM2's accumulator with the residual: acc2[:, box b] = e_in + b_O (this thread's own lane; E4 of the
        // previous tile has read these columns)
        const int b = part + 4 * j;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t v[16];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 bo = *reinterpret_cast<const float4*>(prm + 5 * H + 32 * b + 16 * hh + 4 * q);
            const float4 xx = xin[4 * hh + q];
            const float2 o0 = add2(make_float2(xx.x, xx.y), make_float2(bo.x, bo.y));
            const float2 o1 = add2(make_float2(xx.z, xx.w), make_float2(bo.z, bo.w));
            v[4 * q] = __float_as_uint(o0.x); v[4 * q + 1] = __float_as_uint(o0.y);
            v[4 * q + 2] = __float_as_uint(o1.x); v[4 * q + 3] = __float_as_uint(o1.y);
          }
          tmem_st16(t_acc2 + 32 * b + 16 * hh, v);
        }
        XSUB(4);   // residual preload
        WTR(5 + 5 * j);
      };
      const bool finish_prev = it > 0 && !debug;
      for (int j = 0; j < 2; ++j) {
        float4 xin[8];
        if (have_tile) {
          if (LUT) load_lut(part + 4 * j, xin); else convert_box(j, xin);
        }
        if (finish_prev) e4_box(j);
        if (have_tile && !debug) preload_box(j, xin);
      }
      gbox += 8;
      if (!have_tile) break;
      if (LUT) {
        // GEMM1 by table lookup: acc1[row][cbase .. cbase + 63] = (C lut[x])[...]  (GEMM2 of the previous tile, the last
        // reader of these columns, completed before the E4 above)
#pragma unroll
        for (int hh = 0; hh < 4; ++hh) {
          uint32_t v[16];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 c4 = __ldg(reinterpret_cast<const float4*>(P.cl + xsel * H + cbase + 16 * hh) + q);
            v[4 * q] = __float_as_uint(c4.x); v[4 * q + 1] = __float_as_uint(c4.y);
            v[4 * q + 2] = __float_as_uint(c4.z); v[4 * q + 3] = __float_as_uint(c4.w);
          }
          tmem_st16(t_acc1 + cbase + 16 * hh, v);
        }
        tmem_wait_st();
      }

      // ---- this tile's edge endpoints, node segments, gather pointers ----
      const int ib = it & 1;
      mbar_wait(&idx_full[ib], (it >> 1) & 1, P.error_flag, 21);
      const int* s_row = reinterpret_cast<const int*>(smem + OFF_IDX) + ib * IDX_INTS;
      const int* s_col = s_row + TC_TILE;
      const int my_row = s_row[r];
      const bool valid = my_row >= 0;
      uint32_t seg_mask;
      {
        int next_row = __shfl_down_sync(0xffffffffu, my_row, 1);
        bool seg_end = valid && (lane == 31 || next_row != my_row);
        seg_mask = __ballot_sync(0xffffffffu, seg_end);
      }
      const int nseg = __popc(seg_mask);
      const float* gptr[4];
      const float* bptr = nullptr;
      // A h[col] | V h[col]: lane moves 16-byte unit (lane & 3) of rows (j*8 + lane/4); units 0,1 = A, 2,3 = V.  Rows
      // past the end of the edge list gather zeros: their messages are exactly 0 (sum / mean) without a select
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int cv = s_col[wq * 32 + j * 8 + (lane >> 2)], uu = lane & 3;
        const int cj = (cv >= 0) ? (cv & 0x3fffffff) : -1;
        gptr[j] = (cj >= 0) ? P.uvab + (size_t)cj * 4 * H + ((uu < 2) ? 2 * H : H) + cbase + (uu & 1) * 4
                            : P.zero_row + ((uu < 2) ? 2 * H : H) + cbase + (uu & 1) * 4;
      }
      // B h[row]: one 32-byte slot per node segment of this warp (the host routes graphs with more than MAXSEG
      // segments per 32-edge group to the single-CTA kernel)
      {
        int* segrow = reinterpret_cast<int*>(smem + OFF_SEG) + ww * MAXSEG;
        const int my_seg = __popc(seg_mask & ((1u << lane) - 1u));
        if (((seg_mask >> lane) & 1u) && my_seg < MAXSEG) segrow[my_seg] = my_row;   // the last row of a segment publishes its node
        __syncwarp();
        const int sl = (lane >> 1) & (MAXSEG - 1);
        if (lane < 2 * MAXSEG && sl < nseg) bptr = P.uvab + (size_t)segrow[sl] * 4 * H + 3 * H + cbase + (lane & 1) * 4;
      }
      const uint32_t goff = (uint32_t)((lane >> 2) * 64 + (((lane & 3) ^ ((lane >> 3) & 3)) << 4));
      auto gather_issue = [&](int step, unsigned char* buf) {
#ifdef DFB_PHASE_PROF
        if (P.probe & 1) { cp_async_commit(); return; }   // tuning build: E1 without its gather traffic (results wrong)
#endif
        const uint32_t b32 = smem_u32(buf);
#pragma unroll
        for (int j = 0; j < 4; ++j) cp_async16_hint(b32 + j * 512 + goff, gptr[j] + step * 8, pol_keep);
        if (bptr) cp_async16_hint(b32 + 2048 + lane * 16, bptr + step * 8, pol_keep);
        cp_async_commit();
      };
      XSUB(5);   // endpoints, segments, gather pointers
      WTR(11);
      if (!debug) {
        if (it > 0) {   // this warp's store of box part + 4 has read the window: the gathers may overwrite it
          if (lane == 0 && !(DFB_EXP & 4)) tma_store_wait_read_n<0>();
          __syncwarp();
        }
        XSUB(6);   // wait for the staging before the first gathers
        gather_issue(0, gbuf0);
        gather_issue(1, gbuf0 + GBUF);
      }
      PHASE(0);   // X phase
      WTR(12);
      const int s_edge = tile * TC_TILE + r;
      prev_valid = valid;
      if (!LUT) {
        mbar_wait(&acc_rdy[0], it & 1, P.error_flag, 6);
        tc_fence_after();
      }
      PHASE(1);   // wait for GEMM1
      WTR(13);
      if (debug) {
#pragma unroll 1
        for (int c0 = cbase; c0 < cbase + 64; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(t_acc1 + c0, v);
          tmem_wait_ld();
          if (valid) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              __stcg(reinterpret_cast<float4*>(P.debug_acc + (size_t)s_edge * H + c0) + j,
                     make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                 __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])));
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&idx_free[it & 1]);
        continue;
      }

      // ================= E1: e_hat, gate, messages, row statistics (8 steps of 8 columns) =================
      const int first_node = s_row[2 * TC_TILE + wq];
      const size_t pair_base = (size_t)s_row[2 * TC_TILE + 4 + wq];
      const uint32_t gb32 = smem_u32(gbuf0);
      // shared-memory addresses that are fixed for the tile (buffer 1 = + GBUF)
      const uint32_t a_row = gb32 + (uint32_t)lane * 64u + ((uint32_t)((lane >> 1) & 3) << 4);   // unit u of my row: a_row ^ (u << 4)
      // rows past the edge list read the last real segment's slot: every value that enters their (discarded) row stays finite
      const uint32_t b_slot = gb32 + 2048u + (uint32_t)min(__popc(seg_mask & ((1u << lane) - 1u)), max(min(nseg, MAXSEG), 1) - 1) * 32u;
      const uint32_t p_st = gb32 + (uint32_t)lane * 4u;                                          // patch[c][lane], pitch 144 B
      const uint32_t p_ld = gb32 + (uint32_t)(lane & 7) * 144u + (uint32_t)(lane >> 3) * 32u;     // my 8 rows of column lane & 7
      // Segment structure of this warp's 32 rows.  With at most two node segments (every k-NN / ER workload: a 32-edge
      // group crosses at most one node boundary when degrees are >= 32) the reduce lane (column lane & 7, rows r_lo..r_lo+7)
      // needs one number: k0 = how many of its 8 rows belong to the first segment; prefix sums give both segment sums.
      // (A kernel variant with only this path compiled in was slower: more spills under the 96-register cap.)
      const int r_lo = (lane >> 3) * 8;
      const bool fast = !MAXAGG && nseg <= 2 && nseg >= 1;
      int k0 = 8;
      float* part0 = nullptr; float* part1 = nullptr;
      if (fast) {
        const int e0 = __ffs(seg_mask) - 1;                       // last row of the first segment
        const int n0 = __shfl_sync(0xffffffffu, my_row, e0);
        k0 = min(max(e0 + 1 - r_lo, 0), 8);
        part0 = P.partials + (pair_base + (size_t)(n0 - first_node)) * H + cbase + (lane & 7);
        if (nseg == 2) {
          const int n1 = __shfl_sync(0xffffffffu, my_row, 31 - __clz(seg_mask));
          part1 = P.partials + (pair_base + (size_t)(n1 - first_node)) * H + cbase + (lane & 7);
        }
      }
      float2 nK = splat2(0.f), S1p = splat2(0.f), S1q = splat2(0.f), Q1p = splat2(0.f), Q1q = splat2(0.f);
      auto lds128 = [](uint32_t addr) {
        float4 v4;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v4.x), "=f"(v4.y), "=f"(v4.z), "=f"(v4.w) : "r"(addr));
        return v4;
      };
      auto sts32 = [](uint32_t addr, float x) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(x) : "memory"); };
      auto e1_step = [&](const int step, const uint32_t bo) {   // bo = byte offset of this step's buffer (0 or GBUF)
        uint32_t v[8];
        tmem_ld8(t_acc1 + cbase + step * 8, v);
        if (step < 7) cp_async_wait<1>(); else cp_async_wait<0>();
        __syncwarp();                       // this warp's pieces of the step have landed
        const float4 a0 = lds128(a_row + bo), a1 = lds128((a_row ^ 16u) + bo);
        const float4 v0 = lds128((a_row ^ 32u) + bo), v1 = lds128((a_row ^ 48u) + bo);
        const float4 b0 = lds128(b_slot + bo), b1 = lds128(b_slot + bo + 16u);
        tmem_wait_ld();
        float2 x[4], m[4];
        x[0] = add2(add2(make_float2(__uint_as_float(v[0]), __uint_as_float(v[1])), make_float2(a0.x, a0.y)), make_float2(b0.x, b0.y));
        x[1] = add2(add2(make_float2(__uint_as_float(v[2]), __uint_as_float(v[3])), make_float2(a0.z, a0.w)), make_float2(b0.z, b0.w));
        x[2] = add2(add2(make_float2(__uint_as_float(v[4]), __uint_as_float(v[5])), make_float2(a1.x, a1.y)), make_float2(b1.x, b1.y));
        x[3] = add2(add2(make_float2(__uint_as_float(v[6]), __uint_as_float(v[7])), make_float2(a1.z, a1.w)), make_float2(b1.z, b1.w));
        if (step == 0) nK = splat2(-x[0].x);
        {
          const float2 d0 = add2(x[0], nK), d1 = add2(x[1], nK), d2 = add2(x[2], nK), d3 = add2(x[3], nK);
          S1p = add2(S1p, add2(d0, d2));
          S1q = add2(S1q, add2(d1, d3));
          Q1p = fma2(d0, d0, Q1p);
          Q1q = fma2(d1, d1, Q1q);
          Q1p = fma2(d2, d2, Q1p);
          Q1q = fma2(d3, d3, Q1q);
        }
#ifdef DFB_PHASE_PROF
        if (P.probe & 2) {   // tuning build: E1 without the sigmoid (results wrong)
          m[0] = mul2(x[0], make_float2(v0.x, v0.y)); m[1] = mul2(x[1], make_float2(v0.z, v0.w));
          m[2] = mul2(x[2], make_float2(v1.x, v1.y)); m[3] = mul2(x[3], make_float2(v1.z, v1.w));
        } else
#endif
        {
          m[0] = mul2(sigmoid_mufu2(x[0]), make_float2(v0.x, v0.y));
          m[1] = mul2(sigmoid_mufu2(x[1]), make_float2(v0.z, v0.w));
          m[2] = mul2(sigmoid_mufu2(x[2]), make_float2(v1.x, v1.y));
          m[3] = mul2(sigmoid_mufu2(x[3]), make_float2(v1.z, v1.w));
        }
        if (MAXAGG && !valid) m[0] = m[1] = m[2] = m[3] = splat2(-INFINITY);   // sum / mean: the gathered V row is zero
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          v[2 * j] = __float_as_uint(x[j].x);
          v[2 * j + 1] = __float_as_uint(x[j].y);
        }
        tmem_st8(t_acc1 + cbase + step * 8, v);
        __syncwarp();                       // every lane has read its A / V / B values: the buffer becomes the patch
#pragma unroll
        for (int j = 0; j < 4; ++j) {       // transposed [column][row], 36-float pitch: conflict-free
          sts32(p_st + bo + (uint32_t)(2 * j) * 144u, m[j].x);
          sts32(p_st + bo + (uint32_t)(2 * j + 1) * 144u, m[j].y);
        }
        __syncwarp();
        // row-segment reduction: lane = (column lane & 7, row group lane >> 3); row groups are combined with a fixed
        // shuffle tree; the segment structure is warp-uniform; deterministic, no atomics
        const float4 t0 = lds128(p_ld + bo), t1 = lds128(p_ld + bo + 16u);
        if (fast) {
          // prefix sums in row order; first segment = the first k0 rows, second = the rest (rows past the edge list are 0)
          const float p1 = t0.x, p2 = p1 + t0.y, p3 = p2 + t0.z, p4 = p3 + t0.w;
          const float p5 = p4 + t1.x, p6 = p5 + t1.y, p7 = p6 + t1.z, p8 = p7 + t1.w;
          const float lo4 = (k0 & 2) ? ((k0 & 1) ? p3 : p2) : ((k0 & 1) ? p1 : 0.0f);
          const float hi4 = (k0 & 2) ? ((k0 & 1) ? p7 : p6) : ((k0 & 1) ? p5 : p4);
          float s0 = (k0 & 8) ? p8 : ((k0 & 4) ? hi4 : lo4);
          float s1 = p8 - s0;
          s0 += __shfl_xor_sync(0xffffffffu, s0, 8);
          s1 += __shfl_xor_sync(0xffffffffu, s1, 8);
          s0 += __shfl_xor_sync(0xffffffffu, s0, 16);
          s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
          if (lane < 8) {
            part0[step * 8] = s0;
            if (part1) part1[step * 8] = s1;
          }
        } else {
          const float mv[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
          const int r_hi = r_lo + 7;
          uint32_t mask = seg_mask;
          int start = 0;
          while (mask) {                         // one iteration per node segment present in this warp
            const int end = __ffs(mask) - 1;
            mask &= mask - 1;
            const int lo = max(start, r_lo) - r_lo, hi = min(end, r_hi) - r_lo;   // my 8 rows of this segment
            const uint32_t rm = (hi >= lo) ? (((2u << hi) - 1u) & ~((1u << lo) - 1u)) : 0u;
            float run;
            if (MAXAGG) {
              float q0 = -INFINITY, q1 = -INFINITY;
#pragma unroll
              for (int i = 0; i < 8; i += 2) {
                if ((rm >> i) & 1u) q0 = fmaxf(q0, mv[i]);
                if ((rm >> (i + 1)) & 1u) q1 = fmaxf(q1, mv[i + 1]);
              }
              run = fmaxf(q0, q1);
              run = fmaxf(run, __shfl_xor_sync(0xffffffffu, run, 8));
              run = fmaxf(run, __shfl_xor_sync(0xffffffffu, run, 16));
            } else {
              float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
              for (int i = 0; i < 8; i += 4) {
                if ((rm >> i) & 1u) q0 += mv[i];
                if ((rm >> (i + 1)) & 1u) q1 += mv[i + 1];
                if ((rm >> (i + 2)) & 1u) q2 += mv[i + 2];
                if ((rm >> (i + 3)) & 1u) q3 += mv[i + 3];
              }
              run = (q0 + q1) + (q2 + q3);
              run += __shfl_xor_sync(0xffffffffu, run, 8);
              run += __shfl_xor_sync(0xffffffffu, run, 16);
            }
            const int node = __shfl_sync(0xffffffffu, my_row, end);
            if (lane < 8) P.partials[(pair_base + (size_t)(node - first_node)) * H + cbase + step * 8 + lane] = run;
            start = end + 1;
          }
        }
        __syncwarp();                       // patch may be overwritten by the next gather
        if (step + 2 < 8) gather_issue(step + 2, gbuf0 + bo);
      };
#pragma unroll   // E1 / E2 / E3 fully unrolled: the step offsets fold into immediates (-3.8 % kernel time, fewer spills)
      for (int s2 = 0; s2 < 8; s2 += 2) {   // two steps per iteration: the buffer parity is static
        e1_step(s2, 0u);
        e1_step(s2 + 1, (uint32_t)GBUF);
      }
      const float K1 = -nK.x, S1 = (S1p.x + S1p.y) + (S1q.x + S1q.y), Q1 = (Q1p.x + Q1p.y) + (Q1q.x + Q1q.y);
      tmem_wait_st();
      // LayerNorm statistics of the four column parts of a row are exchanged through the (idle) gather buffers
      float mean1, rstd1;
      {
        float* sb = reinterpret_cast<float*>(gbuf0);
        sb[lane] = K1;
        sb[32 + lane] = S1;
        sb[64 + lane] = Q1;
        worker_bar();
        if (lane == 0) mbar_arrive(&idx_free[it & 1]);   // endpoints of this tile are no longer read
        float kk[4], sp[4], qp[4];
        float msum = 0.f;
#pragma unroll
        for (int p2 = 0; p2 < 4; ++p2) {
          const float* pp = stat_buf(p2);
          kk[p2] = pp[lane]; sp[p2] = pp[32 + lane]; qp[p2] = pp[64 + lane];
          msum += kk[p2] * 64.0f + sp[p2];
        }
        mean1 = msum * (1.0f / H);
        float ss = 0.f;   // sum (x - mean)^2 = sum_p [Q_p - 2 (mean - K_p) S_p + n_p (mean - K_p)^2]
#pragma unroll
        for (int p2 = 0; p2 < 4; ++p2) {
          const float dk = mean1 - kk[p2];
          ss += qp[p2] - 2.0f * dk * sp[p2] + 64.0f * dk * dk;
        }
        rstd1 = rsqrtf(fmaxf(ss * (1.0f / H), 0.0f) + LN_EPS);
      }
      PHASE(2);   // E1
      WTR(14);
      // ================= E2: e_til = relu(LN_e(e_hat)) + tau, statistics for LN_O =================
      float S2, Q2;
      {
        const float2 rs = splat2(rstd1), nm = splat2(-mean1 * rstd1);
        float2 S2p = splat2(0.f), Q2p = splat2(0.f);
#pragma unroll
        for (int c0 = cbase; c0 < cbase + 64; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(t_acc1 + c0, v);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 g4 = *reinterpret_cast<const float4*>(prm + c0 + 4 * j);
            const float4 b4 = *reinterpret_cast<const float4*>(prm + H + c0 + 4 * j);
            const float4 t4 = *reinterpret_cast<const float4*>(prm + 2 * H + c0 + 4 * j);
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const float2 xx = make_float2(__uint_as_float(v[4 * j + 2 * hh]), __uint_as_float(v[4 * j + 2 * hh + 1]));
              const float2 gg = hh ? make_float2(g4.z, g4.w) : make_float2(g4.x, g4.y);
              const float2 bb2 = hh ? make_float2(b4.z, b4.w) : make_float2(b4.x, b4.y);
              const float2 tt = hh ? make_float2(t4.z, t4.w) : make_float2(t4.x, t4.y);
              float2 y = fma2(fma2(xx, rs, nm), gg, bb2);          // LN_e affine
              y = add2(make_float2(fmaxf(y.x, 0.0f), fmaxf(y.y, 0.0f)), tt);   // ReLU + time vector
              S2p = add2(S2p, y);
              Q2p = fma2(y, y, Q2p);
              v[4 * j + 2 * hh] = __float_as_uint(y.x);
              v[4 * j + 2 * hh + 1] = __float_as_uint(y.y);
            }
          }
          tmem_st16(t_acc1 + c0, v);
        }
        S2 = S2p.x + S2p.y;
        Q2 = Q2p.x + Q2p.y;
      }
      tmem_wait_st();
      {
        float* sb = reinterpret_cast<float*>(gbuf0);
        sb[128 + lane] = S2;
        sb[160 + lane] = Q2;
        worker_bar();   // also: every part's e_til is in TMEM before E3 reads columns written by other warps
        S2 = 0.f;
        Q2 = 0.f;
#pragma unroll
        for (int p2 = 0; p2 < 4; ++p2) {
          const float* pp = stat_buf(p2);
          S2 += pp[128 + lane];
          Q2 += pp[160 + lane];
        }
      }
      const float mean2 = S2 * (1.0f / H);
      const float var2 = fmaxf(Q2 * (1.0f / H) - mean2 * mean2, 0.0f);
      const float rstd2 = rsqrtf(var2 + LN_EPS);
      const float2 rs2 = splat2(rstd2), nm2 = splat2(-mean2 * rstd2);
      PHASE(3);   // E2
      // ================= E3: s = silu(LN_O(e_til)) -> GEMM2 A operand, in place in TMEM =================
      // every 64-column K-chunk is produced cooperatively (part p converts columns [64 kc + 16 p, +16)), so chunk 0 is
      // complete after a quarter of E3 and GEMM2 runs underneath the rest
#pragma unroll
      for (int kc = 0; kc < 4; ++kc) {
        const int c0 = kc * 64 + part * 16;
        uint32_t v[16], w16[16];
        tmem_ld16(t_acc1 + c0, v);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 2; ++j) {               // 8 elements each
          float z[8];
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int e0 = 8 * j + 4 * hh;
            const float4 g4 = *reinterpret_cast<const float4*>(prm + 3 * H + c0 + e0);
            const float4 b4 = *reinterpret_cast<const float4*>(prm + 4 * H + c0 + e0);
            const float2 t01 = fma2(fma2(make_float2(__uint_as_float(v[e0]), __uint_as_float(v[e0 + 1])), rs2, nm2),
                                    make_float2(g4.x, g4.y), make_float2(b4.x, b4.y));
            const float2 t23 = fma2(fma2(make_float2(__uint_as_float(v[e0 + 2]), __uint_as_float(v[e0 + 3])), rs2, nm2),
                                    make_float2(g4.z, g4.w), make_float2(b4.z, b4.w));
            const float2 s01 = mul2(t01, sigmoid_mufu2(t01)), s23 = mul2(t23, sigmoid_mufu2(t23));   // SiLU
            z[4 * hh] = s01.x; z[4 * hh + 1] = s01.y; z[4 * hh + 2] = s23.x; z[4 * hh + 3] = s23.y;
          }
          uint2 h0, l0, h1, l1;
          split4(make_float4(z[0], z[1], z[2], z[3]), h0, l0);
          split4(make_float4(z[4], z[5], z[6], z[7]), h1, l1);
          w16[4 * j] = h0.x; w16[4 * j + 1] = h0.y; w16[4 * j + 2] = h1.x; w16[4 * j + 3] = h1.y;             // hi: columns 0..7
          w16[8 + 4 * j] = l0.x; w16[8 + 4 * j + 1] = l0.y; w16[8 + 4 * j + 2] = l1.x; w16[8 + 4 * j + 3] = l1.y;   // lo: columns 8..15
        }
        tmem_st16(t_acc1 + c0, w16);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&a2_full[kc]);
      }
      PHASE(4);   // E3
      WTR(15);
    }
    if (lane == 0) tma_store_wait_all();   // this warp's result stores
    if (GNSTATS && lane < 8) {
      // block = (CTA, lane quarter) supplies all 32 groups: warp (part, wq) owns groups 4 (part + 4 j) + g
      const double* acc = reinterpret_cast<const double*>(smem + OFF_GN) + ww * 16;
      double* dst = P.gn_part + ((size_t)(blockIdx.x * 4 + wq) * 32) * 2;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int grp_id = 4 * (part + 4 * j) + (lane >> 1);
        dst[grp_id * 2 + (lane & 1)] = acc[j * 8 + lane];
      }
    }
#ifdef DFB_PHASE_PROF
    if (prof) {
      for (int i = 0; i < 5; ++i) atomicAdd(P.phase_cycles + 16 + i, (unsigned long long)pc[i]);   // [16..20]: pair kernel phases
      for (int i = 0; i < 8; ++i) atomicAdd(P.phase_cycles + 24 + i, (unsigned long long)px[i]);   // [24..31]: X sub-phases
    }
#endif
#undef PHASE
#undef XSUB
#undef WTR
  }

  // teardown: nobody may leave while the peer can still signal this CTA's barriers or read its shared memory
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
struct State {
  bool ready = false;
  CUtensorMap wmap;   // bf16 weight arena, box 32 K x 128 rows, 64B swizzle
  CUtensorMap smap;   // fp32 edge stream, box 32 columns x 32 rows, 128B swizzle: one worker warp's rows of a result box
  const void* smap_ptr = nullptr;
  long long smap_rows = 0;
  int max_clusters = 0;
};

inline int init(State* st, TcState* tc) {
  cudaError_t e = cudaSuccess;
  const void* fns[6] = {(const void*)k_edge_layer_pair<false, MODE_PLAIN>, (const void*)k_edge_layer_pair<false, MODE_GN>,
                        (const void*)k_edge_layer_pair<false, MODE_LUT>, (const void*)k_edge_layer_pair<true, MODE_PLAIN>,
                        (const void*)k_edge_layer_pair<true, MODE_GN>, (const void*)k_edge_layer_pair<true, MODE_LUT>};
  for (int i = 0; i < 6 && e == cudaSuccess; ++i)
    e = cudaFuncSetAttribute(fns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) {
    tc->err = std::string("cudaFuncSetAttribute(pair kernel): ") + cudaGetErrorString(e);
    return -2;
  }
  st->max_clusters = tc->num_sms / 2;
  return 0;
}

inline int bind_weights(State* st, TcState* tc, const void* arena, int L) {
  cuuint64_t gdim[2] = {(cuuint64_t)H, (cuuint64_t)(L * 12 + 4) * H};
  cuuint64_t gstride[1] = {(cuuint64_t)H * sizeof(uint16_t)};
  cuuint32_t box[2] = {32u, 128u};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = ((PFN_encodeTiled)tc->encode_fn)(&st->wmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(arena), gdim,
                                                gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    tc->err = "cuTensorMapEncodeTiled(weights, 64B swizzle) failed with CUresult " + std::to_string((int)r);
    return -2;
  }
  st->ready = true;
  return 0;
}

// One fused layer that writes e.  LUT mode (cl != nullptr): layer 0, whose input rows are one of two table rows (categorical
// TSP: lut_x selects per edge; MIS: lut_x == nullptr, both tables zero) - GEMM1 and the input stream are replaced by lookups.
// gn_part != nullptr (last layer of the sparse TSP encoder): the kernel also leaves the head's GroupNorm partial sums,
// *gn_blocks = number of [32][2] blocks written.
inline int launch(State* st, TcState* tc, int l, float* e, const float* uvab, float* partials, GraphDev g, LayerParams lp,
                  const float* tvec_edge, int agg_mode, cudaStream_t stream, double* gn_part = nullptr,
                  int* gn_blocks = nullptr, const float* lut_x = nullptr, const float* cl = nullptr,
                  const float* lut = nullptr, int reverse = 0) {
  tc->last_launches = 0;
  if (!st->ready) {
    tc->err = "pair kernel: weights not bound";
    return -1;
  }
  int r = tc_ensure_emap(tc, e, g.E);
  if (r) return r;
  {
    const long long e_rows = (long long)((g.E + TC_TILE - 1) / TC_TILE) * TC_TILE;
    if (st->smap_ptr != (const void*)e || st->smap_rows != e_rows) {
      cuuint64_t gdim[2] = {(cuuint64_t)H, (cuuint64_t)e_rows};
      cuuint64_t gstride[1] = {(cuuint64_t)H * sizeof(float)};
      cuuint32_t box[2] = {32u, 32u};
      cuuint32_t estr[2] = {1u, 1u};
      CUresult cr = ((PFN_encodeTiled)tc->encode_fn)(&st->smap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)e, gdim, gstride, box, estr,
                                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (cr != CUDA_SUCCESS) {
        tc->err = "cuTensorMapEncodeTiled(e, 32-row store box) failed with CUresult " + std::to_string((int)cr);
        return -2;
      }
      st->smap_ptr = (const void*)e;
      st->smap_rows = e_rows;
    }
  }
  Params P;
  P.e = e; P.uvab = uvab; P.partials = partials; P.g = g; P.lp = lp; P.tvec = tvec_edge;
  P.zero_row = tc->zero_row4; P.debug_acc = tc->debug_acc; P.error_flag = tc->error_flag; P.phase_cycles = tc->phase_cycles;
  P.agg_mode = agg_mode; P.w_row_base = l * 12 * H;
  P.gn_part = (tc->debug_acc || cl) ? nullptr : gn_part; P.E = g.E;
  P.lut_x = lut_x; P.cl = cl; P.lut = lut;
  P.n_tiles = (g.E + TC_TILE - 1) / TC_TILE;
  P.probe = tc->probe;
  P.reverse = reverse;
  const int n_pairs = (P.n_tiles + 1) / 2;
  const int clusters = n_pairs < st->max_clusters ? n_pairs : st->max_clusters;

  const int mode = cl ? MODE_LUT : (P.gn_part ? MODE_GN : MODE_PLAIN);
  const int grid = 2 * clusters;
  if (agg_mode == AGG_MAX) {
    if (mode == MODE_LUT) k_edge_layer_pair<true, MODE_LUT><<<grid, THREADS, SMEM_BYTES, stream>>>(st->wmap, tc->emap, st->smap, P);
    else if (mode == MODE_GN) k_edge_layer_pair<true, MODE_GN><<<grid, THREADS, SMEM_BYTES, stream>>>(st->wmap, tc->emap, st->smap, P);
    else k_edge_layer_pair<true, MODE_PLAIN><<<grid, THREADS, SMEM_BYTES, stream>>>(st->wmap, tc->emap, st->smap, P);
  } else {
    if (mode == MODE_LUT) k_edge_layer_pair<false, MODE_LUT><<<grid, THREADS, SMEM_BYTES, stream>>>(st->wmap, tc->emap, st->smap, P);
    else if (mode == MODE_GN) k_edge_layer_pair<false, MODE_GN><<<grid, THREADS, SMEM_BYTES, stream>>>(st->wmap, tc->emap, st->smap, P);
    else k_edge_layer_pair<false, MODE_PLAIN><<<grid, THREADS, SMEM_BYTES, stream>>>(st->wmap, tc->emap, st->smap, P);
  }
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    tc->err = std::string("pair kernel launch: ") + cudaGetErrorString(err);
    return -2;
  }
  if (gn_blocks) *gn_blocks = P.gn_part ? 2 * clusters * 4 : 0;
  tc->last_launches = 1;
  return 0;
}

}  // namespace v2
}  // namespace dfb
