// bf16 GEMM for sm_100a: TMA → 128B-swizzled shared memory → tcgen05.mma (accumulators in TMEM)
// → tcgen05.ld epilogue.  Persistent, warp-specialised, one CTA per SM.
//
//   C[M,N] (bf16 | fp32, optional +=)  =  A ⋅ Bᵀ   with fp32 accumulation
//     A: "K-major"  stored [M,K] (K contiguous)   or "MN-major" stored [K,M] (M contiguous)
//     B: "K-major"  stored [N,K] (K contiguous)   or "MN-major" stored [K,N] (N contiguous)
//
// The three linear-layer GEMMs map onto this without any transposes in memory:
//     y  = x  Wᵀ      : A=x  (K-major)   B=W  (K-major)
//     dx = dy W       : A=dy (K-major)   B=W  (MN-major)
//     dW = dyᵀ x      : A=dy (MN-major)  B=x  (MN-major)     (fp32 accumulate into main_grad)
//
// Tile 128 x 256 x 64, 4-stage TMA ring (48 KB / stage), 2 TMEM accumulator stages (2 x 256 cols =
// the whole 512-column TMEM) so the epilogue of tile i overlaps the main loop of tile i+1.
// Warp roles: w0 = TMA producer (1 lane), w1 = MMA issuer (1 lane), w2 = TMEM alloc, w4..7 = epilogue.
//
// PAIR = 1 is the cta_group::2 variant: a 2-CTA cluster (the two SMs of a TPC) computes a 256 x 256 tile. Each CTA stages its
// own 128 A rows and HALF of B (128 of the 256 N rows), so per-SM shared-memory traffic per FLOP halves (and the 32 KB stages
// allow a 6-deep ring); the leader CTA issues M=256 tcgen05.mma.cta_group::2 instructions that read both CTAs' operands and
// write each CTA's TMEM with its own 128 accumulator rows; both CTAs run the epilogue on their half. Completion is
// multicast to both CTAs' barriers; the peer's TMA transactions and epilogue arrivals land on the leader's barriers.
#include <cuda.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "tc_common.cuh"
#include "tmap.h"

namespace {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int kAccStages = 2;
constexpr int kThreads = 256;
constexpr uint32_t kABytes = BM * BK * 2;  // 16 KB
constexpr uint32_t kStagingBytes = 4 /*epilogue warps*/ * 2 /*buffers*/ * 4096;  // 32 rows x 128 B per buffer
constexpr int kGroupM = 16;  // raster: super-rows of 16 M-tiles keep the A slab L2-resident

// per-variant geometry: PAIR=0 one CTA owns 128 x 256; PAIR=1 a CTA pair owns 256 x 256 and each CTA stages half of B
constexpr uint32_t kCopyChunkBytes = 128 * BK * 2;  // IO = 3 weight-gather copier: one {64 col, 128 row} box = 16 KB
constexpr int kCopyBufs = 2;
constexpr int kGatherSub = 4;  // IO = 3: readiness flags per rank block (a block is released to the MMA tiles in quarters)

constexpr uint32_t kAuxSlabBytes = 4096;  // EPI = 3: one {64 col x 32 row} bf16 slab (gate or up) per epilogue warp

template <int PAIR, int IO = 0, int EPI = 0>
struct Geo {
  static constexpr int kTileM = PAIR ? 2 * BM : BM;        // rows of C per scheduled tile
  static constexpr int kBRows = PAIR ? BN / 2 : BN;        // B rows staged by one CTA
  static constexpr uint32_t kBBytes = kBRows * BK * 2;     // 16 KB | 32 KB
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;
  // IO = 3 gives one ring stage (32 KB) to the weight-gather copier's two 16 KB bounce buffers
  // EPI = 3 (SwiGLU backward) gives two stages to the TMA-staged gate/up slabs its epilogue reads: per epilogue warp a
  // double-buffered pair of 4 KB slabs (the loads run two 64-feature groups ahead of the math)
  // (with IO = 3 the copier already took a stage: single-buffered slabs there, four stages left in both combinations)
  static constexpr int kAuxSets = EPI == 3 ? (IO == 3 ? 1 : 2) : 0;
  static constexpr int kStages = PAIR ? 6 - (IO == 3 ? 1 : 0) - (EPI == 3 ? kAuxSets : 0) : 4;
  static constexpr uint32_t kCopyBytes = IO == 3 ? kCopyBufs * kCopyChunkBytes : 0;
  static constexpr uint32_t kAuxBytes = 4 * kAuxSets * 2 * kAuxSlabBytes;
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + kStagingBytes + kCopyBytes + kAuxBytes + 1024 /*align slack*/ + 512 /*barriers*/;
};

using namespace tc;

__host__ __device__ constexpr uint32_t make_idesc(int a_mn, int b_mn, int m) {
  // c=f32 (1<<4), a=bf16 (1<<7), b=bf16 (1<<10), majors (15,16), N>>3 at 17, M>>4 at 24
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void tile_coords(int tile, int tiles_m, int tiles_n, int& tm, int& tn, int group_m = kGroupM) {
  const int per_group = group_m * tiles_n;
  const int g = tile / per_group;
  const int first_m = g * group_m;
  const int gsz = min(group_m, tiles_m - first_m);
  const int r = tile - g * per_group;
  tm = first_m + r % gsz;
  tn = r / gsz;
}

struct GemmParams {
  int M, N, K;
  int ldc;
  int a_mn, b_mn;
  int c_fp32, accumulate;
  int split_k;  // >1 only with c_fp32 && accumulate: every K slice reduce-adds its partial tile (weight-gradient GEMMs)
  void* C;
  // optional RoPE epilogue (QKV projection): output columns [0, rope_cols) are heads of width rope_D whose interleaved pairs are
  // rotated by the position (row % rope_S) while the tile is still in registers; tables fp32 [rope_S.., rope_D/2]. bf16 outputs only.
  const float* rope_cos;
  const float* rope_sin;
  int rope_S, rope_cols, rope_D;
  // optional SwiGLU epilogue (gate/up projection, CTA-pair scheduler only): B = W13 stored [2·FF, K] as [gate rows | up rows];
  // a tile covers 128 features: the leader CTA stages the 128 GATE rows, its partner the matching 128 UP rows, so every CTA's
  // accumulator holds gate in columns [0,128) and up in [128,256) for the same features. The epilogue stores gate and up (bf16,
  // for the backward) into C = gate_up [M, 2·FF] and h = silu(gate)·up into the second output [M, FF].
  int swiglu_ff;
  int group_m;  // raster super-row height in M tiles (0 = kGroupM)
  // optional SwiGLU-BACKWARD epilogue (EPI = 3; the down-projection's input-gradient GEMM dh = dy·W2, N = FF): the epilogue reads the
  // saved bf16 gate/up activations of the same (row, feature) straight from `aux` = gate_up [M, 2·FF] and stores
  // d_gate = dh·up·silu'(gate) and d_up = dh·silu(gate) into C = d_gate_up [M, 2·FF] — dh never exists in memory and the
  // stand-alone swiglu_bwd pass (3 % of the Llama-1B step in round 1) disappears. swiglu_ff = FF.
  const void* aux;
  int ld_aux;
};

// Fused collective GEMMs over the NVLink symmetric heap (IO template parameter of the kernel):
//   IO = 1  all-gather ⊕ GEMM : row block r of A lives on rank r. Peer memory is NOT cached in the local L2, so a remote A tile
//           must cross NVLink exactly once: the tiles of output column 0 of every remote row block ("gather tiles", scheduled
//           first, owners interleaved) TMA-load their A tiles straight from the owner into the MMA ring, and an otherwise idle
//           warp TMA-stores each consumed ring slot into a local gathered copy of A; a per-row-block flag then releases the
//           remaining column tiles of that block, which read the local copy (L2-resident). Local row blocks are scheduled between
//           the two, so the NVLink transfer hides behind them. (Measured on 2 GPUs: loading every tile from the peer re-fetched
//           each remote row 22 times and ran 3.6x slower than NCCL all-gather + GEMM.)
//   IO = 2  GEMM ⊕ reduce-scatter : every rank computes a full partial C; the epilogue TMA-reduce-adds each fp32 tile into the
//           buffer of the rank that owns those rows (cp.reduce.async.bulk.tensor on a peer-mapped address). Row blocks are visited
//           owner-interleaved (rank+1, rank+2, …, self) so the NVLink reduce traffic is spread over the whole kernel and, at any
//           moment, the ranks target distinct owners.
//   IO = 3  parameter all-gather ⊕ GEMM (ZeRO-3 / reshard_after_forward): the WEIGHT (operand B, stored [rows, cols]) is sharded by
//           rows over the ranks of the FSDP group — rank r owns rows [r·rpr, (r+1)·rpr) in its symmetric-heap shard. Nothing is
//           gathered ahead of the kernel: warp 3 of EVERY CTA is a copier that TMA-loads {64 x 128} boxes of the peers' shards
//           over NVLink into a 16 KB bounce buffer and TMA-stores them into a local full-size copy of the weight, walking the rank
//           blocks in consumption order (own block first); a per-quarter-block counter releases the MMA tiles, whose TMA producer
//           waits on the counters of exactly the B rows it is about to load and reads them from the local copy (L2-resident).
//           K-major B (forward): the output-column order is rotated to start at the own block; MN-major B (input gradient, the
//           contraction runs over the sharded rows): the K loop is rotated instead. The shards are read-only during
//           forward/backward, so no inter-rank barrier is needed around the kernel.
struct PeerMaps {
  CUtensorMap m[8];
  int n;              // ranks
  int rows_per_rank;  // IO 1/2: multiple of the tile height; IO 3: rows of B owned by each rank
  int rank;
  uint32_t* flags;    // IO = 1: one counter per 256-row block of the gathered A (zeroed by the host before the launch)
                      // IO = 3: n * kGatherSub counters (chunks landed per quarter block), zeroed before the launch
  int idle_rounds;    // IO = 1: scheduling rounds a CTA pair sits out after a gather tile (it costs ~3 ordinary tiles: NVLink latency)
  // IO = 3
  int order[8];       // rank blocks in the order the copiers fetch them (= the order the tiles consume them)
  int cpr;            // 128-row chunk rows per rank block = ceil(rows_per_rank / 128)
  int col_chunks;     // cols / 64
  int spc;            // chunk rows per readiness counter = ceil(cpr / kGatherSub)
  int start;          // first output-column tile (K-major B) / first k-block (MN-major B) of the rotated order
  int gate;           // 1: B is gathered by THIS kernel and the tiles wait for it; 0: B is already resident in the local scratch
  // gather-AHEAD list: the weight the NEXT GEMM of the layer sequence consumes (forward: wqkv → wo → w13 → w2 → next layer;
  // backward: the reverse dgrad order). After (or instead of) its own weight the copier pulls this one into ITS scratch region, so
  // the next kernel finds its operand resident and runs un-gated: in steady state every gather hides under the previous GEMM and
  // only the first kernel of a pass waits for NVLink (measured without it, 4 x B200: forward +6 %, dgrad +40…90 % — the first wave
  // of dgrad tiles needs every rank block).
  CUtensorMap nm[8];
  int nn;             // ranks holding the next weight (0 = nothing to prefetch)
  int ncpr, ncol_chunks;
  int norder[8];
};

// IO = 1 tile schedule: [gather tiles][local tiles][dependent tiles]; kind 1 = gather (tn = 0 of a remote block), 0 = local rows,
// 2 = dependent (waits for the block's flag, reads the local gathered copy). Lower indices never wait on higher ones.
__device__ __forceinline__ void ag_tile(int idx, const PeerMaps& pm, int tiles_n, int tpr, int& tm, int& tn, int& kind) {
  const int n = pm.n, G = (n - 1) * tpr, L = tpr * tiles_n;
  if (idx < G) {
    tm = ((pm.rank + 1 + idx % (n - 1)) % n) * tpr + idx / (n - 1);
    tn = 0;
    kind = 1;
  } else if (idx < G + L) {
    const int j = idx - G;
    tm = pm.rank * tpr + j % tpr;
    tn = j / tpr;
    kind = 0;
  } else {
    const int j = idx - G - L, g = j % G;
    tm = ((pm.rank + 1 + g % (n - 1)) % n) * tpr + g / (n - 1);
    tn = 1 + j / G;
    kind = 2;
  }
}
// IO = 1 static schedule with load balancing: iteration `it` of CTA pair `sched_id` → tile index, -1 = idle slot, -2 = done.
// Round 0 hands gather tile p to pair p < G; those pairs then sit out `idle` rounds while the others keep taking tiles.
__device__ __forceinline__ int ag_next(int it, int sched_id, int sched_n, int G, int idle, int num_tiles) {
  if (G == 0 || G > sched_n || idle == 0) {
    const int t = sched_id + it * sched_n;
    return t < num_tiles ? t : -2;
  }
  const int R = num_tiles - G;  // non-gather tiles, numbered in order behind the gather tiles
  int r;
  if (it == 0) {
    if (sched_id < G) return sched_id;
    r = sched_id - G;
  } else if (it <= idle) {
    if (sched_id < G) return -1;
    r = (sched_n - G) * it + (sched_id - G);
  } else {
    r = (sched_n - G) * (idle + 1) + (it - idle - 1) * sched_n + sched_id;
  }
  return r < R ? G + r : -2;
}
// IO = 2 row-block order: raster index t → owner-interleaved tile row
__device__ __forceinline__ int rs_tile_row(int t, const PeerMaps& pm, int tpr) { return ((pm.rank + 1 + t % pm.n) % pm.n) * tpr + t / pm.n; }

// rotate the 16 interleaved pairs held in 32 consecutive fp32 accumulator registers (head-dim offset d0, sequence position pos)
__device__ __forceinline__ void rope_regs(uint32_t (&r)[32], const float* __restrict__ cosb, const float* __restrict__ sinb, int pos, int d0,
                                          int half_d) {
  const float4* c4 = reinterpret_cast<const float4*>(cosb + (int64_t)pos * half_d + (d0 >> 1));
  const float4* s4 = reinterpret_cast<const float4*>(sinb + (int64_t)pos * half_d + (d0 >> 1));
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float4 c = __ldg(c4 + k), sn = __ldg(s4 + k);
    const float cc[4] = {c.x, c.y, c.z, c.w}, ss[4] = {sn.x, sn.y, sn.z, sn.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = (k * 4 + e) * 2;
      const float a = __uint_as_float(r[j]), b = __uint_as_float(r[j + 1]);
      r[j] = __float_as_uint(a * cc[e] - b * ss[e]);
      r[j + 1] = __float_as_uint(a * ss[e] + b * cc[e]);
    }
  }
}

// EPI: 0 = plain, 1 = RoPE on the leading output columns, 2 = SwiGLU (PAIR only), 3 = SwiGLU backward. Separate instantiations on purpose: the fused
// epilogues need 170-230 registers per thread, and a plain GEMM compiled with that footprint fills the SM's register file, so
// the gradient-reduction / optimizer kernels of the comm stream can no longer co-reside with it during the backward pass
// (measured: the 2-GPU step gained only a third of what the 1-GPU step gained when the epilogues were runtime branches).
template <int A_MN, int B_MN, int PAIR, int EPI, int IO>
__global__ void __launch_bounds__(kThreads, 1)
    gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                     const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_h,
                     const __grid_constant__ CUtensorMap tmap_g, const __grid_constant__ CUtensorMap tmap_g2, const GemmParams p,
                     const __grid_constant__ PeerMaps pm) {
  using G = Geo<PAIR, IO, EPI>;
  constexpr int kStages = G::kStages;
  constexpr uint32_t kStageBytes = G::kStageBytes;
  constexpr int kBRows = G::kBRows;
  constexpr int kTileM = G::kTileM;
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;  // 0 = leader (issues the MMAs)
  const int sched_id = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int sched_n = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + kStages * kStageBytes;  // 1024-aligned: 4 warps x 2 buffers x 4 KB
  uint8_t* copy_buf = staging + kStagingBytes;  // IO = 3: kCopyBufs x 16 KB bounce buffers of the weight-gather copier
  uint8_t* aux_buf = copy_buf + G::kCopyBytes;  // EPI = 3: per epilogue warp one gate and one up slab (TMA-loaded)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux_buf + G::kAuxBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;
  uint64_t* tempty_bar = tfull_bar + kAccStages;
  uint64_t* copied_bar = tempty_bar + kAccStages;  // IO = 1: the gather copier has finished reading a consumed ring slot
  uint64_t* cp_bar = copied_bar + kStages;         // IO = 3: a peer box has landed in bounce buffer i
  uint64_t* aux_bar = cp_bar + kCopyBufs;          // EPI = 3: [warp q][set]: that warp's gate/up slabs of one group have landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_bar + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr bool swiglu = PAIR && EPI == 2;
  const int tiles_m = (p.M + kTileM - 1) / kTileM, tiles_n = swiglu ? (p.swiglu_ff + 127) / 128 : (p.N + BN - 1) / BN;
  const int num_kb = (p.K + BK - 1) / BK;
  // split-K: work item w = tile * split_k + slice; slice s covers k-blocks [s*kb_per, min(num_kb, (s+1)*kb_per))
  const int split_k = p.split_k;
  const int kb_per = (num_kb + split_k - 1) / split_k;
  const int num_tiles = tiles_m * tiles_n * split_k;  // = number of work items
  const int ag_G = (IO == 1 && pm.n > 1) ? (pm.n - 1) * (pm.rows_per_rank / kTileM) : 0;  // gather tiles (lowest indices)
  // every role walks the same tile sequence: round-robin over the CTAs (pairs); all-gather mode inserts idle slots (ag_next)
  auto next_tile = [&](int& it) -> int {
    if (IO == 1 && pm.n > 1) {
      for (;;) {
        const int t = ag_next(it++, sched_id, sched_n, ag_G, pm.idle_rounds, num_tiles);
        if (t != -1) return t;
      }
    }
    const int t = sched_id + (it++) * sched_n;
    return t < num_tiles ? t : -2;
  };
  const int group_m = p.group_m > 0 ? p.group_m : kGroupM;
  // IO = 3 rotations (see PeerMaps): K-major B → output columns start at the own rank block; MN-major B → the K loop does
  const int tn_rot = (IO == 3 && B_MN == 0 && pm.gate) ? pm.start : 0;
  const int kb_rot = (IO == 3 && B_MN == 1 && pm.gate) ? pm.start : 0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    prefetch_tmap(&tmap_c);
    prefetch_tmap(&tmap_h);
    if (IO == 3) {
      prefetch_tmap(&tmap_g);
      prefetch_tmap(&tmap_g2);
    }
    if (IO != 0)
      for (int i = 0; i < pm.n; ++i) prefetch_tmap(&pm.m[i]);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
      mbar_init(&copied_bar[i], 1);
    }
    for (int i = 0; i < kCopyBufs; ++i) mbar_init(&cp_bar[i], 1);
    for (int i = 0; i < 8; ++i) mbar_init(&aux_bar[i], 1);
    for (int i = 0; i < kAccStages; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], PAIR ? 8 : 4);  // one arrive per epilogue warp (of both CTAs in a pair)
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (PAIR) tmem_alloc_pair(tmem_slot, 512);
    else tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();  // the peer's barriers must exist before any remote arrive / TMA completion targets them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int tpr = IO != 0 ? pm.rows_per_rank / kTileM : 0;  // tile rows per rank
      // gather tiles have the lowest indices: this CTA's are tiles sched_id, sched_id + sched_n, … below G
      const int gather_fills = ag_G > sched_id ? ((ag_G - sched_id + sched_n - 1) / sched_n) * num_kb : 0;
      int fill = 0;
      // IO = 3: wait until the copiers have landed B rows [r0, r1] in the local full copy. One readiness counter per quarter of a
      // rank block; a counter that was seen complete stays complete for the rest of the kernel (bit in `ready`).
      uint32_t ready = 0;
      auto gate_rows = [&](int r0, int r1) {
        if (IO != 3 || !pm.gate) return;
        const int rpr = pm.rows_per_rank;
        for (int x = r0; x <= r1;) {
          const int b = x / rpr, cr = (x - b * rpr) >> 7;
          const int q = min(cr / pm.spc, kGatherSub - 1);
          const int id = b * kGatherSub + q;
          const int cr_end = q == kGatherSub - 1 ? pm.cpr : min(pm.cpr, (q + 1) * pm.spc);  // chunk rows [q*spc, cr_end)
          if (!((ready >> id) & 1u)) {
            const uint32_t expect = (uint32_t)((cr_end - q * pm.spc) * pm.col_chunks);
            SpinGuard guard;
            while (pb::ld_acquire_gpu_u32(pm.flags + id) < expect) guard.tick();
            fence_proxy_async_all();  // generic-proxy acquire → the TMA (async-proxy) reads of those rows
            ready |= 1u << id;
          }
          x = b * rpr + min(rpr, cr_end * 128);  // first row of the next quarter (or of the next rank block)
        }
      };
      for (int it = 0, tile; (tile = next_tile(it)) >= 0;) {
        int tm, tn, kind = 0;
        if (IO == 1 && pm.n > 1) {
          ag_tile(tile, pm, tiles_n, tpr, tm, tn, kind);
        } else {
          tile_coords(tile / split_k, tiles_m, tiles_n, tm, tn, group_m);
          if (IO == 2) tm = rs_tile_row(tm, pm, tpr);
          if (IO == 3 && tn_rot) tn = (tn + tn_rot) % tiles_n;
        }
        const int kb0 = (tile % split_k) * kb_per, kb1 = min(num_kb, kb0 + kb_per);
        if (kb0 >= kb1) continue;  // empty K slice (split does not divide K): every role skips it identically
        const int m0 = tm * kTileM + (int)crank * BM;      // this CTA's A rows
        // this CTA's share of the B rows (SwiGLU mode: leader = gate rows, partner = the matching up rows)
        const int n0 = swiglu ? (int)crank * p.swiglu_ff + tn * 128 : tn * BN + (int)crank * kBRows;
        if (IO == 3 && B_MN == 0) gate_rows(n0, min(n0 + kBRows, pm.rows_per_rank * pm.n) - 1);  // my B rows have landed locally
        if (IO == 1 && kind == 2) {  // the block's gather tile (both CTAs of that pair) has stored these rows locally
          SpinGuard guard;
          while (pb::ld_acquire_gpu_u32(pm.flags + tm) < 2u) guard.tick();
          fence_proxy_async_all();  // generic-proxy acquire → the TMA (async-proxy) reads below
        }
        for (int kbi = kb0; kbi < kb1; ++kbi) {
          const int kb = kb_rot ? (kbi + kb_rot) % num_kb : kbi;
          if (IO == 3 && B_MN == 1) gate_rows(kb * BK, min(kb * BK + BK, pm.rows_per_rank * pm.n) - 1);
          mbar_wait(&empty_bar[stage], phase ^ 1);
          // the slot's previous fill was a gather fill (they are the first fills of every CTA): the copier must have drained it
          if (IO == 1 && fill >= kStages && fill - kStages < gather_fills) mbar_wait(&copied_bar[stage], phase ^ 1);
          ++fill;
          uint8_t* sa = smem + stage * kStageBytes;
          uint8_t* sb = sa + kABytes;
          // PAIR: the leader's barrier collects the bytes of BOTH CTAs' loads (complete_tx may precede expect_tx in a phase)
          if (!PAIR || crank == 0) mbar_expect_tx(&full_bar[stage], (PAIR ? 2u : 1u) * kStageBytes);
          auto load = [&](const CUtensorMap* m, void* dst, int c0, int c1) {
            if (PAIR) tma_load_2d_pair(m, &full_bar[stage], dst, c0, c1);
            else tma_load_2d(m, &full_bar[stage], dst, c0, c1);
          };
          if (IO == 1) {
            // gather / local tiles read the owner's block (NVLink for remote owners); dependent tiles read the local gathered copy
            const int owner = m0 / pm.rows_per_rank;
            if (kind == 2) load(&tmap_a, sa, kb * BK, m0);
            else load(&pm.m[owner], sa, kb * BK, m0 - owner * pm.rows_per_rank);
          } else if (A_MN == 0) {
            load(&tmap_a, sa, kb * BK, m0);  // box {64 k, 128 m}
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)  // box {64 m, 64 k} → 8 KB each
              load(&tmap_a, sa + j * (BK * 128), m0 + j * 64, kb * BK);
          }
          if (B_MN == 0) {
            load(&tmap_b, sb, kb * BK, n0);  // box {64 k, kBRows n}
          } else {
#pragma unroll
            for (int j = 0; j < kBRows / 64; ++j)
              load(&tmap_b, sb + j * (BK * 128), n0 + j * 64, kb * BK);
          }
          if (++stage == kStages) stage = 0, phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0 && crank == 0) {
      constexpr uint32_t idesc = make_idesc(A_MN, B_MN, kTileM);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int it = 0, tile; (tile = next_tile(it)) >= 0;) {
        const int kb0 = (tile % split_k) * kb_per, kb1 = min(num_kb, kb0 + kb_per);
        if (kb0 >= kb1) continue;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);  // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kStageBytes);
          const uint32_t sb = sa + kABytes;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // K-major: +32 B per UMMA_K inside the 128 B swizzle row.  MN-major: +16 k-rows * 128 B.
            const uint64_t adesc = A_MN == 0 ? make_smem_desc(sa + k * 32, 16, 1024)
                                             : make_smem_desc(sa + k * 2048, BK * 128, 1024);
            const uint64_t bdesc = B_MN == 0 ? make_smem_desc(sb + k * 32, 16, 1024)
                                             : make_smem_desc(sb + k * 2048, BK * 128, 1024);
            if (PAIR) umma_bf16_pair(tmem_d, adesc, bdesc, idesc, ((kb - kb0) | k) != 0 ? 1u : 0u);
            else umma_bf16(tmem_d, adesc, bdesc, idesc, ((kb - kb0) | k) != 0 ? 1u : 0u);
          }
          // smem slot reusable once these MMAs retire (PAIR: in both CTAs)
          if (PAIR) umma_commit_pair(&empty_bar[stage]);
          else umma_commit(&empty_bar[stage]);
          if (++stage == kStages) stage = 0, phase ^= 1;
        }
        // accumulator complete → epilogue (PAIR: of both CTAs)
        if (PAIR) umma_commit_pair(&tfull_bar[acc]);
        else umma_commit(&tfull_bar[acc]);
        if (++acc == kAccStages) acc = 0, acc_phase ^= 1;
      }
    }
  } else if (IO == 1 && warp == 3) {
    // ------------------------------------------------------------------ gather copier (all-gather ⊕ GEMM only)
    // follows the ring in consumption order: once the MMAs have retired a slot (empty barrier), a gather tile's A slab is
    // TMA-stored into the local gathered copy before the producer may refill the slot; after the tile, the block's flag is bumped
    if (lane == 0) {
      const int tpr = pm.rows_per_rank / kTileM;
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0, tile; (tile = next_tile(it)) >= 0;) {
        int tm = 0, tn = 0, kind = 0;
        if (pm.n > 1) ag_tile(tile, pm, tiles_n, tpr, tm, tn, kind);
        if (kind != 1) break;  // gather tiles come first; nothing to copy afterwards (the producer stops waiting for us too)
        const int m0 = tm * kTileM + (int)crank * BM;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase);
          tma_store_2d(&tmap_a, smem + stage * kStageBytes, kb * BK, m0);
          bulk_commit();
          bulk_wait_read<0>();
          mbar_arrive(&copied_bar[stage]);
          if (++stage == kStages) stage = 0, phase ^= 1;
        }
        bulk_wait_all();          // the rows are in local memory …
        fence_proxy_async_all();  // … async-proxy writes ordered before the generic-proxy release below
        pb::red_add_release_gpu_u32(pm.flags + tm, 1u);
      }
    }
  } else if (IO == 3 && warp == 3) {
    // ------------------------------------------------------------------ weight-gather copier (parameter all-gather ⊕ GEMM)
    // Every CTA of the grid takes the {64 col x 128 row} boxes g = blockIdx.x, blockIdx.x + gridDim.x, … of the whole weight in
    // consumption order: peer shard --TMA load (NVLink)--> bounce buffer --TMA store--> local full copy, then bumps the
    // readiness counter of the box's quarter block once the store has fully completed. Two boxes are in flight per CTA
    // (148 x 32 KB per ~3 us of NVLink latency is well above what the MMA tiles consume).
    if (lane == 0) {
      uint32_t ph[kCopyBufs] = {0, 0};
      // one list = one weight: (peer maps, local store map, rank order, geometry); `flags` != nullptr → publish per-quarter counters
      auto gather_list = [&](const CUtensorMap* maps, const CUtensorMap* store_map, const int* order, int n, int cpr, int col_chunks,
                             int spc, uint32_t* flags) {
        const int cpb = cpr * col_chunks;  // boxes per rank block
        const int total = cpb * n;
        const int mine = total > (int)blockIdx.x ? (total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
        auto box = [&](int i, int& owner, int& cr, int& cc) {
          const int g = (int)blockIdx.x + i * (int)gridDim.x;
          const int j = g / cpb, c = g - j * cpb;
          owner = order[j];
          cr = c / col_chunks;
          cc = c - cr * col_chunks;
        };
        auto fetch = [&](int i) {  // rows past the end of the owner's shard are zero-filled by the TMA unit
          int owner, cr, cc;
          box(i, owner, cr, cc);
          const int buf = i % kCopyBufs;
          mbar_expect_tx(&cp_bar[buf], kCopyChunkBytes);
          tma_load_2d(&maps[owner], &cp_bar[buf], copy_buf + buf * kCopyChunkBytes, cc * BK, cr * 128);
        };
        for (int i = 0; i < kCopyBufs && i < mine; ++i) fetch(i);  // both NVLink loads in flight from the start
        int prev_id = -1;
        for (int i = 0; i < mine; ++i) {
          int owner, cr, cc;
          box(i, owner, cr, cc);
          const int buf = i % kCopyBufs;
          mbar_wait(&cp_bar[buf], ph[buf]);
          ph[buf] ^= 1;
          // … and clipped by the store: dimension 1 of the 3-D map is the rows of ONE rank block
          tma_store_3d(store_map, copy_buf + buf * kCopyChunkBytes, cc * BK, cr * 128, owner);
          bulk_commit();
          if (flags != nullptr && prev_id >= 0) {  // the previous box's WRITES are complete (not only its smem reads): publish it
            asm volatile("cp.async.bulk.wait_group 1;" ::: "memory");
            fence_proxy_async_all();
            pb::red_add_release_gpu_u32(flags + prev_id, 1u);
          }
          prev_id = owner * kGatherSub + min(cr / spc, kGatherSub - 1);
          if (i + kCopyBufs < mine) {
            bulk_wait_read<0>();  // this buffer has been read out by the store above
            fetch(i + kCopyBufs);
          }
        }
        bulk_wait_all();  // every box of this list is in local memory (the NEXT kernel relies on that for a gather-ahead list)
        if (flags != nullptr && prev_id >= 0) {
          fence_proxy_async_all();
          pb::red_add_release_gpu_u32(flags + prev_id, 1u);
        }
      };
      if (pm.gate) gather_list(pm.m, &tmap_g, pm.order, pm.n, pm.cpr, pm.col_chunks, pm.spc, pm.flags);
      if (pm.nn > 0) gather_list(pm.nm, &tmap_g2, pm.norder, pm.nn, pm.ncpr, pm.ncol_chunks, 1, nullptr);
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    // TMEM → registers → 128B-swizzled smem slab (32 rows x 128 B per warp) → TMA store / reduce-add.
    // Each warp owns its TMEM lane quadrant, its own staging double-buffer and its own bulk groups,
    // so the four epilogue warps never synchronise with each other.
    const int q = warp & 3;
    uint8_t* my_stage = staging + q * 8192;
    const uint32_t row_sw = (uint32_t)(lane & 7);
    int acc = 0;
    uint32_t acc_phase = 0;
    int buf = 0;
    const int tpr = IO != 0 ? pm.rows_per_rank / kTileM : 0;
    // EPI = 3: the saved gate/up activations of this warp's 32 rows reach the epilogue through TMA — one {64 x 32} slab each per
    // 64-feature group, double-buffered and requested TWO groups ahead (a look-ahead cursor walks this CTA's tile sequence), so an
    // HBM-latency load is never on the epilogue's critical path. (v1: every thread fetched its own row with 16-byte LDGs — 32 lines
    // per instruction, no L1 beside 224 KB of shared memory: 2.0x the MMA time. v2: single-buffered slabs one group ahead: 1.7x.)
    uint32_t aux_phase[2] = {0, 0};
    int aux_k = 0;                  // groups consumed so far (set = aux_k & 1)
    int la_it = 0, la_tile = -2, la_tm = 0, la_tn = 0, la_g = BN / 64;  // look-ahead cursor: next group to request
    int la_k = 0;                   // groups requested so far
    auto aux_request_next = [&]() {  // lane 0 only: request the next group of this CTA's sequence into set (la_k & 1)
      if (la_g == BN / 64) {         // move the cursor to the next tile
        if (la_tile == -1) return;   // sequence exhausted
        const int t = next_tile(la_it);
        if (t < 0) {
          la_tile = -1;
          return;
        }
        la_tile = t;
        tile_coords(t / split_k, tiles_m, tiles_n, la_tm, la_tn, group_m);
        if (IO == 3 && tn_rot) la_tn = (la_tn + tn_rot) % tiles_n;
        la_g = 0;
      }
      constexpr int kSets = G::kAuxSets > 0 ? G::kAuxSets : 1;
      const int set = la_k % kSets;
      const int f0 = la_tn * BN + la_g * 64, r0 = la_tm * kTileM + (int)crank * BM + q * 32;
      uint8_t* dst = aux_buf + (q * kSets + set) * 2 * kAuxSlabBytes;
      mbar_expect_tx(&aux_bar[q * 2 + set], 2 * kAuxSlabBytes);
      tma_load_2d(&tmap_h, &aux_bar[q * 2 + set], dst, f0, r0);
      tma_load_2d(&tmap_h, &aux_bar[q * 2 + set], dst + kAuxSlabBytes, p.swiglu_ff + f0, r0);
      ++la_g;
      ++la_k;
    };
    if (EPI == 3 && lane == 0) {
      for (int i = 0; i < G::kAuxSets; ++i) aux_request_next();
    }
    for (int it = 0, tile; (tile = next_tile(it)) >= 0;) {
      int tm, tn, kind = 0;
      if (IO == 1 && pm.n > 1) {
        ag_tile(tile, pm, tiles_n, tpr, tm, tn, kind);
      } else {
        tile_coords(tile / split_k, tiles_m, tiles_n, tm, tn, group_m);
        if (IO == 2) tm = rs_tile_row(tm, pm, tpr);
        if (IO == 3 && tn_rot) tn = (tn + tn_rot) % tiles_n;
      }

      if ((tile % split_k) * kb_per >= num_kb) continue;  // empty K slice
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int row0 = tm * kTileM + (int)crank * BM + q * 32;
      const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16);
      if (p.c_fp32) {
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {  // 32 fp32 columns = one 128 B row
          uint32_t r[32];
          tmem_ld_32x32b_x32(taddr + c * 32, r);
          if (lane == 0) bulk_wait_read<1>();  // the store that last read this buffer has drained
          __syncwarp();
          tmem_ld_wait();
          const uint32_t sbase = smem_u32(my_stage + buf * 4096) + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            st_shared_v4(sbase + ((j ^ row_sw) << 4), r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
          fence_proxy_async();
          __syncwarp();
          const int col0 = tn * BN + c * 32;
          if (lane == 0 && row0 < p.M && col0 < p.N) {
            if (IO == 2) {  // reduce-scatter: add into the owner rank's buffer (peer-mapped tensor map)
              const int owner = row0 / pm.rows_per_rank;
              tma_reduce_add_2d(&pm.m[owner], my_stage + buf * 4096, col0, row0 - owner * pm.rows_per_rank);
            } else if (p.accumulate) tma_reduce_add_2d(&tmap_c, my_stage + buf * 4096, col0, row0);
            else tma_store_2d(&tmap_c, my_stage + buf * 4096, col0, row0);
          }
          if (lane == 0) bulk_commit();
          buf ^= 1;
        }
      } else if constexpr (EPI == 3) {
        // dh tile (fp32, TMEM) ⊗ saved gate/up (bf16, TMA-staged slabs) → d_gate | d_up (bf16): two 64-column slabs per group, one
        // TMA store each
        const int FF = p.swiglu_ff;
        uint8_t* stage_g = my_stage;
        uint8_t* stage_u = my_stage + 4096;
        const uint32_t sg = smem_u32(stage_g) + lane * 128, su = smem_u32(stage_u) + lane * 128;
#pragma unroll 1
        for (int g = 0; g < BN / 64; ++g) {
          const int f0 = tn * BN + g * 64;
          const bool ok = f0 < FF;  // warp-uniform (FF % 64 == 0)
          constexpr int kSets = G::kAuxSets > 0 ? G::kAuxSets : 1;
          const int set = aux_k % kSets;
          const uint32_t in_g = smem_u32(aux_buf + (q * kSets + set) * 2 * kAuxSlabBytes) + lane * 128, in_u = in_g + kAuxSlabBytes;
          mbar_wait(&aux_bar[q * 2 + set], aux_phase[set]);  // this group's gate / up slabs have landed (requested two groups ago)
          aux_phase[set] ^= 1;
          ++aux_k;
          if (lane == 0) bulk_wait_read<0>();  // both output slabs are free again
          __syncwarp();
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(taddr + g * 64 + half * 32, r);
            uint4 gq[4], uq[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              gq[j] = ld_shared_v4(in_g + (((half * 4 + j) ^ row_sw) << 4));
              uq[j] = ld_shared_v4(in_u + (((half * 4 + j) ^ row_sw) << 4));
            }
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t gw[4] = {gq[j].x, gq[j].y, gq[j].z, gq[j].w}, uw[4] = {uq[j].x, uq[j].y, uq[j].z, uq[j].w};
              uint32_t og[4], ou[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 gf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&gw[e]));
                const float2 uf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&uw[e]));
                const float d0 = __uint_as_float(r[j * 8 + 2 * e]), d1 = __uint_as_float(r[j * 8 + 2 * e + 1]);
                // sigmoid through ex2.approx + rcp.approx (4 instructions): with IEEE division the epilogue needed ~70 instructions per
                // element on only four warps per SM and ran LONGER than the tile's MMAs (fused 489 us vs 250 + 156 us un-fused)
                const float s0 = __fdividef(1.f, 1.f + exp2f(-1.4426950408889634f * gf.x));
                const float s1 = __fdividef(1.f, 1.f + exp2f(-1.4426950408889634f * gf.y));
                const float dg0 = d0 * uf.x * s0 * (1.f + gf.x * (1.f - s0)), dg1 = d1 * uf.y * s1 * (1.f + gf.y * (1.f - s1));
                const float du0 = d0 * gf.x * s0, du1 = d1 * gf.y * s1;
                og[e] = pack_bf16x2(__float_as_uint(dg0), __float_as_uint(dg1));
                ou[e] = pack_bf16x2(__float_as_uint(du0), __float_as_uint(du1));
              }
              st_shared_v4(sg + (((half * 4 + j) ^ row_sw) << 4), og[0], og[1], og[2], og[3]);
              st_shared_v4(su + (((half * 4 + j) ^ row_sw) << 4), ou[0], ou[1], ou[2], ou[3]);
            }
          }
          fence_proxy_async();
          __syncwarp();  // every lane has read the input slabs and written the output slabs
          if (lane == 0) {
            aux_request_next();  // refill the input set just consumed (two groups ahead of the math)
            if (row0 < p.M && ok) {
              tma_store_2d(&tmap_c, stage_g, f0, row0);
              tma_store_2d(&tmap_c, stage_u, FF + f0, row0);
            }
            bulk_commit();
          }
        }
      } else if constexpr (swiglu) {
        // pack 64 fp32 accumulator columns to bf16, stage them in the swizzled slab and TMA-store them at (col0, row0)
        auto emit64 = [&](const uint32_t (&w)[32], const CUtensorMap* map, int col0, bool in_range) {
          if (lane == 0) bulk_wait_read<1>();
          __syncwarp();
          const uint32_t sbase = smem_u32(my_stage + buf * 4096) + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) st_shared_v4(sbase + ((j ^ row_sw) << 4), w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0 && row0 < p.M && in_range) tma_store_2d(map, my_stage + buf * 4096, col0, row0);
          if (lane == 0) bulk_commit();
          buf ^= 1;
        };
        auto load64_packed = [&](uint32_t col, uint32_t (&w)[32]) {
          uint32_t r0[32], r1[32];
          tmem_ld_32x32b_x32(taddr + col, r0);
          tmem_ld_32x32b_x32(taddr + col + 32, r1);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            w[j] = pack_bf16x2(r0[2 * j], r0[2 * j + 1]);
            w[16 + j] = pack_bf16x2(r1[2 * j], r1[2 * j + 1]);
          }
        };
#pragma unroll 1
        for (int g = 0; g < 2; ++g) {  // two 64-feature groups of the tile's 128 features
          const int f0 = tn * 128 + g * 64;
          const bool ok = f0 < p.swiglu_ff;
          uint32_t wg[32], wu[32], wh[32];
          load64_packed(g * 64, wg);
          emit64(wg, &tmap_c, f0, ok);
          load64_packed(128 + g * 64, wu);
          emit64(wu, &tmap_c, p.swiglu_ff + f0, ok);
          // h = silu(gate)·up from the bf16-ROUNDED gate/up — bit-compatible with the stand-alone kernel that reads them back
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const __nv_bfloat162 gb = *reinterpret_cast<const __nv_bfloat162*>(&wg[j]), ub = *reinterpret_cast<const __nv_bfloat162*>(&wu[j]);
            const float2 gf = __bfloat1622float2(gb), uf = __bfloat1622float2(ub);
            const float h0 = gf.x / (1.f + __expf(-gf.x)) * uf.x, h1 = gf.y / (1.f + __expf(-gf.y)) * uf.y;
            wh[j] = pack_bf16x2(__float_as_uint(h0), __float_as_uint(h1));
          }
          emit64(wh, &tmap_h, f0, ok);
        }
      } else {
#pragma unroll 1
        for (int g = 0; g < BN / 64; ++g) {  // 64 bf16 columns = one 128 B row
          uint32_t r0[32], r1[32];
          tmem_ld_32x32b_x32(taddr + g * 64, r0);
          tmem_ld_32x32b_x32(taddr + g * 64 + 32, r1);
          if (lane == 0) bulk_wait_read<1>();
          __syncwarp();
          tmem_ld_wait();
          if (EPI == 1 && tn * BN + g * 64 < p.rope_cols) {  // Q/K head columns: rotate in registers
            const int pos = (row0 + lane) % p.rope_S;
            const int d0 = (tn * BN + g * 64) % p.rope_D;
            rope_regs(r0, p.rope_cos, p.rope_sin, pos, d0, p.rope_D >> 1);
            rope_regs(r1, p.rope_cos, p.rope_sin, pos, d0 + 32, p.rope_D >> 1);
          }
          const uint32_t sbase = smem_u32(my_stage + buf * 4096) + lane * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            st_shared_v4(sbase + ((j ^ row_sw) << 4), pack_bf16x2(r0[8 * j], r0[8 * j + 1]),
                         pack_bf16x2(r0[8 * j + 2], r0[8 * j + 3]), pack_bf16x2(r0[8 * j + 4], r0[8 * j + 5]),
                         pack_bf16x2(r0[8 * j + 6], r0[8 * j + 7]));
#pragma unroll
          for (int j = 0; j < 4; ++j)
            st_shared_v4(sbase + (((j + 4) ^ row_sw) << 4), pack_bf16x2(r1[8 * j], r1[8 * j + 1]),
                         pack_bf16x2(r1[8 * j + 2], r1[8 * j + 3]), pack_bf16x2(r1[8 * j + 4], r1[8 * j + 5]),
                         pack_bf16x2(r1[8 * j + 6], r1[8 * j + 7]));
          fence_proxy_async();
          __syncwarp();
          const int col0 = tn * BN + g * 64;
          if (lane == 0 && row0 < p.M && col0 < p.N) {
            if (p.accumulate) tma_reduce_add_2d(&tmap_c, my_stage + buf * 4096, col0, row0);
            else tma_store_2d(&tmap_c, my_stage + buf * 4096, col0, row0);
          }
          if (lane == 0) bulk_commit();
          buf ^= 1;
        }
      }
      // every TMEM read of this accumulator has completed (wait::ld above) → hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_leader(&tempty_bar[acc]);  // the leader's MMA warp waits for both CTAs' epilogues
        else mbar_arrive(&tempty_bar[acc]);
      }
      if (++acc == kAccStages) acc = 0, acc_phase ^= 1;
    }
    if (lane == 0) {
      if (IO == 2) bulk_wait_all();  // peer reductions must have been performed before this kernel retires (a barrier follows)
      else bulk_wait_read<0>();      // staging smem must outlive the last store's reads
    }
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all();  // no CTA may exit (or free TMEM) while its partner can still signal it / be read by the MMAs
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------- host side
int g_num_sms = 0;

// Split-K for reduce-add outputs. Measured on B200 (profiles/gemm_splitk_r1.json): for grids that already fill the machine,
// splitting K to smooth wave quantisation LOSES 3-25 % — the extra fp32 tile reductions cost more than the idle tail — so the
// automatic policy only splits when the output has fewer than half as many tiles as there are schedulable CTAs (pairs), i.e.
// when most SMs would otherwise sit idle for the whole kernel; slices keep >= 16 k-blocks (K >= 1024) each.
int choose_split_k(int tiles, int units, int num_kb) {
  if (tiles * 2 > units) return 1;
  int s = units / tiles;
  if (s > 8) s = 8;
  while (s > 1 && num_kb / s < 16) --s;
  return s < 1 ? 1 : s;
}

int g_split_k_mode = -1;  // -1 auto, 0/1 off, n>1 forced

template <int A_MN, int B_MN, int PAIR, int EPI = 0, int IO = 0>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& th, const GemmParams& p,
           int max_ctas, cudaStream_t stream, const PeerMaps* pmp = nullptr, const CUtensorMap* tgp = nullptr,
           const CUtensorMap* tg2p = nullptr) {
  static const PeerMaps kNoPeers = {};
  const PeerMaps& pm = pmp ? *pmp : kNoPeers;
  const CUtensorMap& tg = tgp ? *tgp : tc;
  const CUtensorMap& tg2 = tg2p ? *tg2p : tg;
  using G = Geo<PAIR, IO, EPI>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_kernel<A_MN, B_MN, PAIR, EPI, IO>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)G::kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  int grid = g_num_sms;
  if (max_ctas > 0 && max_ctas < grid) grid = max_ctas;
  const int tiles_n = (PAIR && EPI == 2) ? (p.swiglu_ff + 127) / 128 : (p.N + BN - 1) / BN;
  const int tiles = ((p.M + G::kTileM - 1) / G::kTileM) * tiles_n * p.split_k;
  if (!PAIR) {
    if (tiles < grid) grid = tiles;
    gemm_bf16_kernel<A_MN, B_MN, 0, EPI, IO><<<grid, kThreads, G::kSmemBytes, stream>>>(ta, tb, tc, th, tg, tg2, p, pm);
  } else {
    grid &= ~1;
    if (2 * tiles < grid) grid = 2 * tiles;
    if (grid < 2) grid = 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = G::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_bf16_kernel<A_MN, B_MN, 1, EPI, IO>, ta, tb, tc, th, tg, tg2, p, pm);
    if (e != cudaSuccess) return (int)e;
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

// 0 = one CTA per 128x256 tile, 1 = CTA pair per 256x256 tile (default). PB_GEMM_PAIR overrides for A/B testing.
int g_pair_mode = -1;
int pair_mode() {
  if (g_pair_mode < 0) {
    const char* e = getenv("PB_GEMM_PAIR");
    g_pair_mode = e ? (atoi(e) != 0) : 1;  // measured: faster than the single-CTA scheduler (and than cuBLAS) on the Llama shapes
  }
  return g_pair_mode;
}

}  // namespace

// Select the tile scheduler: 0 = single-CTA 128x256 tiles, 1 = CTA-pair 256x256 tiles (cta_group::2), -1 = from PB_GEMM_PAIR.
PB_EXPORT int pb_gemm_set_pair_mode(int mode) {
  const int old = pair_mode();
  g_pair_mode = mode;
  return old;
}

// Split-K policy for fp32 reduce-add outputs: -1 = automatic (default), 0 = never, n > 1 = always n slices (testing).
PB_EXPORT int pb_gemm_set_split_k(int mode) {
  const int old = g_split_k_mode;
  g_split_k_mode = mode;
  return old;
}

// lda/ldb: row stride (elements) of the matrix AS STORED (see header comment).
static int gemm_impl(const void* A, const void* B, void* C, int M, int N, int K, int lda, int ldb, int ldc, int a_mn_major,
                     int b_mn_major, int c_fp32, int accumulate, int max_ctas, const float* rope_cos, const float* rope_sin,
                     int rope_S, int rope_cols, int rope_D, cudaStream_t stream, void* H = nullptr, int ldh = 0, int swiglu_ff = 0) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if ((lda % 8) || (ldb % 8) || (ldc % (c_fp32 ? 4 : 8))) return -1;
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(C)) & 15) return -2;
  CUtensorMap ta, tb;
  int rc;
  if (!a_mn_major)
    rc = pbhost::cached_tmap(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BK, BM);
  else
    rc = pbhost::cached_tmap(&ta, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, 64, BK);
  if (rc) return rc;
  const int pair = (pair_mode() && M > BM) || swiglu_ff > 0;  // a pair needs two M blocks to be worth it; SwiGLU mode requires it
  if (!b_mn_major)
    rc = pbhost::cached_tmap(&tb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, BK, pair ? BN / 2 : BN);
  else
    rc = pbhost::cached_tmap(&tb, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, 64, BK);
  if (rc) return rc;
  CUtensorMap tc;  // epilogue store: 32-row x 128-byte boxes
  rc = c_fp32 ? pbhost::cached_tmap(&tc, C, (uint64_t)M, (uint64_t)N, (uint64_t)ldc, 32, 32, 4)
              : pbhost::cached_tmap(&tc, C, (uint64_t)M, (uint64_t)N, (uint64_t)ldc, 64, 32, 2);
  if (rc) return rc;
  CUtensorMap th = tc;  // second output (SwiGLU mode): h [M, FF]
  if (swiglu_ff > 0) {
    if ((rc = pbhost::cached_tmap(&th, H, (uint64_t)M, (uint64_t)swiglu_ff, (uint64_t)ldh, 64, 32, 2))) return rc;
  }
  GemmParams p{M, N, K, ldc, a_mn_major, b_mn_major, c_fp32, accumulate, 1, C, rope_cos, rope_sin, rope_S, rope_cols, rope_D, swiglu_ff};
  if (c_fp32 && accumulate && g_split_k_mode != 0 && g_split_k_mode != 1) {
    if (g_num_sms == 0) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    int units = (max_ctas > 0 && max_ctas < g_num_sms) ? max_ctas : g_num_sms;
    if (pair) units /= 2;
    const int tile_m = pair ? 2 * BM : BM;
    const int tiles = ((M + tile_m - 1) / tile_m) * ((N + BN - 1) / BN);
    const int num_kb = (K + BK - 1) / BK;
    p.split_k = g_split_k_mode > 1 ? (num_kb / g_split_k_mode >= 1 ? g_split_k_mode : 1) : choose_split_k(tiles, units > 0 ? units : 1, num_kb);
  }
  if (swiglu_ff > 0) return launch<0, 0, 1, 2>(ta, tb, tc, th, p, max_ctas, stream);
  if (rope_cos != nullptr) return pair ? launch<0, 0, 1, 1>(ta, tb, tc, th, p, max_ctas, stream) : launch<0, 0, 0, 1>(ta, tb, tc, th, p, max_ctas, stream);
  if (pair) {
    if (!a_mn_major && !b_mn_major) return launch<0, 0, 1>(ta, tb, tc, th, p, max_ctas, stream);
    if (!a_mn_major && b_mn_major) return launch<0, 1, 1>(ta, tb, tc, th, p, max_ctas, stream);
    if (a_mn_major && !b_mn_major) return launch<1, 0, 1>(ta, tb, tc, th, p, max_ctas, stream);
    return launch<1, 1, 1>(ta, tb, tc, th, p, max_ctas, stream);
  }
  if (!a_mn_major && !b_mn_major) return launch<0, 0, 0>(ta, tb, tc, th, p, max_ctas, stream);
  if (!a_mn_major && b_mn_major) return launch<0, 1, 0>(ta, tb, tc, th, p, max_ctas, stream);
  if (a_mn_major && !b_mn_major) return launch<1, 0, 0>(ta, tb, tc, th, p, max_ctas, stream);
  return launch<1, 1, 0>(ta, tb, tc, th, p, max_ctas, stream);
}

PB_EXPORT int pb_gemm_bf16(const void* A, const void* B, void* C, int M, int N, int K, int lda, int ldb, int ldc,
                           int a_mn_major, int b_mn_major, int c_fp32, int accumulate, int max_ctas,
                           cudaStream_t stream) {
  return gemm_impl(A, B, C, M, N, K, lda, ldb, ldc, a_mn_major, b_mn_major, c_fp32, accumulate, max_ctas, nullptr, nullptr, 1, 0, 64,
                   stream);
}

// y = x·Wᵀ with RoPE applied to output columns [0, rope_cols) in the epilogue (the fused QKV projection). Row r of the output is
// token r of a [B, S] batch (position r % S); heads are rope_D wide; cos/sin are fp32 [>=S, rope_D/2]. bf16 output, K-major operands.
PB_EXPORT int pb_gemm_bf16_rope(const void* A, const void* B, void* C, int M, int N, int K, int lda, int ldb, int ldc, int max_ctas,
                                const float* rope_cos, const float* rope_sin, int rope_S, int rope_cols, int rope_D,
                                cudaStream_t stream) {
  if (rope_cos == nullptr || rope_sin == nullptr || rope_S <= 0 || M % rope_S != 0) return -3;
  if ((rope_D != 64 && rope_D != 128) || rope_cols % rope_D != 0 || rope_cols > N) return -4;
  return gemm_impl(A, B, C, M, N, K, lda, ldb, ldc, 0, 0, 0, 0, max_ctas, rope_cos, rope_sin, rope_S, rope_cols, rope_D, stream);
}

// gate_up = x·W13ᵀ (kept for the backward) AND h = silu(gate)·up in ONE kernel: see GemmParams::swiglu_ff.
//   x [M, K], W13 [2·FF, K] = [gate rows | up rows], gate_up [M, 2·FF], h [M, FF]; all bf16, unit inner stride.
PB_EXPORT int pb_gemm_bf16_swiglu(const void* A, const void* W13, void* gate_up, void* H, int M, int FF, int K, int lda, int ldb,
                                  int ld_gu, int ld_h, cudaStream_t stream) {
  if (FF % 64 != 0 || (ld_h % 8) || (reinterpret_cast<uintptr_t>(H) & 15)) return -5;
  return gemm_impl(A, W13, gate_up, M, 2 * FF, K, lda, ldb, ld_gu, 0, 0, 0, 0, 0, nullptr, nullptr, 1, 0, 64, stream, H, ld_h, FF);
}

// d_gate_up = swiglu'(gate_up) ⊙ (dy·W2) in ONE kernel (EPI = 3): dy [M, K=dim], W2 [dim, FF] (MN-major B: the contraction runs over
// its rows), gate_up [M, 2·FF] saved by the forward, d_gate_up [M, 2·FF]. bf16, unit inner strides, CTA-pair tiles (M > 128).
PB_EXPORT int pb_gemm_bf16_swiglu_bwd(const void* dY, const void* W2, const void* gate_up, void* d_gate_up, int M, int FF, int K,
                                      int lda, int ldb, int ld_gu, int ld_dgu, cudaStream_t stream) {
  if (M <= BM || FF % 64 != 0) return -5;
  if ((lda % 8) || (ldb % 8) || (ld_gu % 8) || (ld_dgu % 8)) return -1;
  if ((reinterpret_cast<uintptr_t>(dY) | reinterpret_cast<uintptr_t>(W2) | reinterpret_cast<uintptr_t>(gate_up) |
       reinterpret_cast<uintptr_t>(d_gate_up)) & 15) return -2;
  CUtensorMap ta, tb, tc;
  int rc;
  if ((rc = pbhost::cached_tmap(&ta, dY, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BK, BM))) return rc;
  if ((rc = pbhost::cached_tmap(&tb, W2, (uint64_t)K, (uint64_t)FF, (uint64_t)ldb, 64, BK))) return rc;
  if ((rc = pbhost::cached_tmap(&tc, d_gate_up, (uint64_t)M, (uint64_t)(2 * FF), (uint64_t)ld_dgu, 64, 32, 2))) return rc;
  CUtensorMap taux;  // saved gate_up [M, 2·FF], {64 x 32} boxes: the epilogue's input slabs
  if ((rc = pbhost::cached_tmap(&taux, gate_up, (uint64_t)M, (uint64_t)(2 * FF), (uint64_t)ld_gu, 64, 32, 2))) return rc;
  GemmParams p{M, FF, K, ld_dgu, 0, 1, 0, 0, 1, d_gate_up, nullptr, nullptr, 1, 0, 64, FF, 0, gate_up, ld_gu};
  return launch<0, 1, 1, 3>(ta, tb, tc, taux, p, 0, stream);
}

// ---------------------------------------------------------------------------------------------------------------------------
// all-gather ⊕ GEMM:  C[n·M_local, N] = [A_0; A_1; …; A_{n-1}] · Bᵀ, A_r = rank r's [M_local, K] block in the symmetric heap.
// a_peers[r] is the address of rank r's block AS MAPPED IN THIS PROCESS. The caller brackets the call with heap barriers (every
// rank's block written before, nobody overwrites it until all ranks are done). K-major operands, bf16 output, CTA-pair tiles.
// a_full: local [n·M_local, K] bf16 scratch that receives the remote row blocks (by-product: the gathered A, minus the local block);
// flags: n·M_local/256 uint32 counters in local memory (zeroed here, on the stream).
// b_mn_major: B stored [K, N] (the transposed weight of a backward GEMM) instead of [N, K].
PB_EXPORT int pb_gemm_allgather(const void* const* a_peers, int n, int rank, const void* B, void* C, void* a_full, uint32_t* flags,
                                int M_local, int N, int K, int lda, int ldb, int ldc, int b_mn_major, cudaStream_t stream) {
  if (n < 1 || n > 8 || rank < 0 || rank >= n || M_local % (2 * BM) != 0) return -6;
  if ((lda % 8) || (ldb % 8) || (ldc % 8) || (K % 8)) return -1;
  PeerMaps pm = {};
  pm.n = n;
  pm.rows_per_rank = M_local;
  pm.rank = rank;
  pm.flags = flags;
  {
    static int idle = -1;
    if (idle < 0) {
      const char* e = getenv("PB_AG_IDLE_ROUNDS");
      idle = e ? atoi(e) : 2;
    }
    pm.idle_rounds = idle;
  }
  int rc;
  if (n > 1) {
    cudaError_t e = cudaMemsetAsync(flags, 0, sizeof(uint32_t) * (size_t)(n * (M_local / (2 * BM))), stream);
    if (e != cudaSuccess) return (int)e;
  }
  for (int r = 0; r < n; ++r)
    if ((rc = pbhost::cached_tmap(&pm.m[r], a_peers[r], (uint64_t)M_local, (uint64_t)K, (uint64_t)lda, BK, BM))) return rc;
  const int M = n * M_local;
  CUtensorMap tfull, tb, tc;
  if ((rc = pbhost::cached_tmap(&tfull, a_full, (uint64_t)M, (uint64_t)K, (uint64_t)K, BK, BM))) return rc;
  if (!b_mn_major) rc = pbhost::cached_tmap(&tb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, BK, BN / 2);
  else rc = pbhost::cached_tmap(&tb, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, 64, BK);
  if (rc) return rc;
  if ((rc = pbhost::cached_tmap(&tc, C, (uint64_t)M, (uint64_t)N, (uint64_t)ldc, 64, 32, 2))) return rc;
  GemmParams p{M, N, K, ldc, 0, b_mn_major, 0, 0, 1, C, nullptr, nullptr, 1, 0, 64, 0};
  return b_mn_major ? launch<0, 1, 1, 0, 1>(tfull, tb, tc, tc, p, 0, stream, &pm) : launch<0, 0, 1, 0, 1>(tfull, tb, tc, tc, p, 0, stream, &pm);
}

// GEMM ⊕ reduce-scatter:  every rank holds a K-shard (A [M, K_local], B [N, K_local]); rank r ends up with rows
// [r·M/n, (r+1)·M/n) of Σ_ranks A·Bᵀ in its fp32 buffer c_peers[r] ([M/n, N], zeroed by the caller before the opening barrier).
PB_EXPORT int pb_gemm_reduce_scatter(const void* A, const void* B, float* const* c_peers, int n, int rank, int M, int N, int K, int lda,
                                     int ldb, int ldc, int b_mn_major, cudaStream_t stream) {
  if (n < 1 || n > 8 || rank < 0 || rank >= n || M % n != 0 || (M / n) % (2 * BM) != 0) return -6;
  if ((lda % 8) || (ldb % 8) || (ldc % 4)) return -1;
  PeerMaps pm = {};
  pm.n = n;
  pm.rows_per_rank = M / n;
  pm.rank = rank;
  int rc;
  for (int r = 0; r < n; ++r)
    if ((rc = pbhost::cached_tmap(&pm.m[r], c_peers[r], (uint64_t)pm.rows_per_rank, (uint64_t)N, (uint64_t)ldc, 32, 32, 4))) return rc;
  CUtensorMap ta, tb;
  if ((rc = pbhost::cached_tmap(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BK, BM))) return rc;
  if (!b_mn_major) rc = pbhost::cached_tmap(&tb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, BK, BN / 2);
  else rc = pbhost::cached_tmap(&tb, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, 64, BK);
  if (rc) return rc;
  GemmParams p{M, N, K, ldc, 0, b_mn_major, 1, 1, 1, c_peers[rank], nullptr, nullptr, 1, 0, 64, 0};
  return b_mn_major ? launch<0, 1, 1, 0, 2>(ta, tb, pm.m[rank], pm.m[rank], p, 0, stream, &pm)
                    : launch<0, 0, 1, 0, 2>(ta, tb, pm.m[rank], pm.m[rank], p, 0, stream, &pm);
}


// ---------------------------------------------------------------------------------------------------------------------------
// parameter all-gather ⊕ GEMM (IO = 3; ZeRO-3 / train.reshard_after_forward): the weight W [rowsW, colsW] is sharded by rows
// over n ranks (rank r: rows [r·rowsW/n, (r+1)·rowsW/n) at w_peers[r], the address AS MAPPED IN THIS PROCESS). The kernel gathers
// it into w_full (local, [rowsW, colsW]) tile by tile WHILE it multiplies; no collective, no barrier (shards are read-only here).
//   b_mn_major = 0 (forward)        C[M, rowsW] = A[M, colsW] · Wᵀ          epi: 0 plain | 1 RoPE on the leading columns | 2 SwiGLU
//   b_mn_major = 1 (input gradient) C[M, colsW] = A[M, rowsW] · W           epi 0 only
// SwiGLU: W = W13 [2·FF, colsW] = [gate rows | up rows], C = gate_up [M, 2·FF], H = silu(gate)·up [M, FF].
// flags: n·4 uint32 in local memory (zeroed here, on the stream). bf16 everywhere, CTA-pair tiles (M > 128).
// epi = 3 (with b_mn_major = 1): SwiGLU-backward epilogue, W = W2 [dim, FF]; H = the saved gate_up [M, 2·FF] (read), C = d_gate_up.
// own_resident = 1: W is already complete in w_full (gathered ahead by the previous kernel): no gather, no gating.
// next_peers != nullptr: additionally gather the weight [next_rows, next_cols] (same n ranks) the NEXT GEMM will consume into
// next_full — un-gated, complete when this kernel retires.
PB_EXPORT int pb_gemm_wgather(const void* A, const void* const* w_peers, int n, int rank, void* w_full, uint32_t* flags, void* C,
                              void* H, int M, int rowsW, int colsW, int lda, int ldc, int ldh, int b_mn_major, int epi,
                              const float* rope_cos, const float* rope_sin, int rope_S, int rope_cols, int rope_D,
                              int own_resident, const void* const* next_peers, void* next_full, int next_rows, int next_cols,
                              cudaStream_t stream) {
  if (n < 1 || n > 8 || rank < 0 || rank >= n || rowsW % n != 0 || colsW % 64 != 0) return -6;
  if ((lda % 8) || (ldc % 8)) return -1;
  if ((epi == 1 || epi == 2) && b_mn_major) return -7;
  if (epi == 3 && !b_mn_major) return -7;
  const int N = b_mn_major ? colsW : rowsW, K = b_mn_major ? rowsW : colsW;
  const int FF = epi == 2 ? rowsW / 2 : (epi == 3 ? colsW : 0);
  if ((epi == 2 || epi == 3) && (FF % 64 != 0 || H == nullptr || (ldh % 8))) return -5;
  if (epi == 1) {
    if (rope_cos == nullptr || rope_sin == nullptr || rope_S <= 0 || M % rope_S != 0) return -3;
    if ((rope_D != 64 && rope_D != 128) || rope_cols % rope_D != 0 || rope_cols > N) return -4;
  }
  if (M <= BM) return -8;  // CTA-pair scheduler only (training shapes); tiny M goes through an explicit gather on the host side
  const int rpr = rowsW / n;
  PeerMaps pm = {};
  pm.n = n;
  pm.rows_per_rank = rpr;
  pm.rank = rank;
  pm.flags = flags;
  pm.cpr = (rpr + 127) / 128;
  pm.col_chunks = colsW / 64;
  pm.spc = (pm.cpr + kGatherSub - 1) / kGatherSub;
  pm.gate = own_resident ? 0 : 1;
  CUtensorMap tg2;
  bool have_next = false;
  if (next_peers != nullptr && next_full != nullptr && next_rows > 0) {
    if (next_rows % n != 0 || next_cols % 64 != 0) return -6;
    const int nrpr = next_rows / n;
    pm.nn = n;
    pm.ncpr = (nrpr + 127) / 128;
    pm.ncol_chunks = next_cols / 64;
    for (int t = 0; t < n; ++t) pm.norder[t] = (rank + 1 + t) % n;  // remote blocks first (they take longest), own block last
    int rc2;
    for (int r = 0; r < n; ++r)
      if ((rc2 = pbhost::cached_tmap(&pm.nm[r], next_peers[r], (uint64_t)nrpr, (uint64_t)next_cols, (uint64_t)next_cols, BK, 128))) return rc2;
    if ((rc2 = pbhost::cached_tmap_blocks(&tg2, next_full, (uint64_t)nrpr, (uint64_t)next_cols, (uint64_t)n, BK, 128))) return rc2;
    have_next = true;
  }
  if (epi == 2 && n % 2 == 0) {
    // a tile needs gate rows (blocks < n/2) and the matching up rows (blocks >= n/2): fetch them pairwise, own pair first
    const int h = n / 2, g0 = rank % h;
    for (int t = 0; t < h; ++t) {
      const int g = (g0 + t) % h;
      pm.order[2 * t] = rank >= h ? g + h : g;
      pm.order[2 * t + 1] = rank >= h ? g : g + h;
    }
    pm.start = (g0 * rpr) / 128;  // tiles are 128 features wide
  } else {
    for (int t = 0; t < n; ++t) pm.order[t] = (rank + t) % n;
    if (epi == 2) pm.start = 0;
    else pm.start = b_mn_major ? (rank * rpr) / BK : (rank * rpr) / BN;
  }
  if (!own_resident) {
    cudaError_t e = cudaMemsetAsync(flags, 0, sizeof(uint32_t) * (size_t)(n * kGatherSub), stream);
    if (e != cudaSuccess) return (int)e;
  }
  int rc;
  for (int r = 0; r < n; ++r)
    if ((rc = pbhost::cached_tmap(&pm.m[r], w_peers[r], (uint64_t)rpr, (uint64_t)colsW, (uint64_t)colsW, BK, 128))) return rc;
  CUtensorMap tg, ta, tb, tc, th;
  if ((rc = pbhost::cached_tmap_blocks(&tg, w_full, (uint64_t)rpr, (uint64_t)colsW, (uint64_t)n, BK, 128))) return rc;
  if ((rc = pbhost::cached_tmap(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BK, BM))) return rc;
  if (!b_mn_major) rc = pbhost::cached_tmap(&tb, w_full, (uint64_t)rowsW, (uint64_t)colsW, (uint64_t)colsW, BK, BN / 2);
  else rc = pbhost::cached_tmap(&tb, w_full, (uint64_t)rowsW, (uint64_t)colsW, (uint64_t)colsW, 64, BK);
  if (rc) return rc;
  if ((rc = pbhost::cached_tmap(&tc, C, (uint64_t)M, (uint64_t)(epi >= 2 ? 2 * FF : N), (uint64_t)ldc, 64, 32, 2))) return rc;
  th = tc;
  if (epi == 2 && (rc = pbhost::cached_tmap(&th, H, (uint64_t)M, (uint64_t)FF, (uint64_t)ldh, 64, 32, 2))) return rc;
  static int group_m = -1;
  if (group_m < 0) {
    const char* ev = getenv("PB_WG_GROUP_M");
    group_m = ev ? atoi(ev) : 0;
  }
  GemmParams p{M, epi == 2 ? 2 * FF : N, K, ldc, 0, b_mn_major, 0, 0, 1, C, rope_cos, rope_sin, rope_S, rope_cols, rope_D, FF, group_m,
               epi == 3 ? H : nullptr, epi == 3 ? ldh : 0};
  const CUtensorMap* tg2p = have_next ? &tg2 : nullptr;
  if (epi == 3) {
    CUtensorMap taux;
    if ((rc = pbhost::cached_tmap(&taux, H, (uint64_t)M, (uint64_t)(2 * FF), (uint64_t)ldh, 64, 32, 2))) return rc;
    return launch<0, 1, 1, 3, 3>(ta, tb, tc, taux, p, 0, stream, &pm, &tg, tg2p);
  }
  if (epi == 2) return launch<0, 0, 1, 2, 3>(ta, tb, tc, th, p, 0, stream, &pm, &tg, tg2p);
  if (epi == 1) return launch<0, 0, 1, 1, 3>(ta, tb, tc, th, p, 0, stream, &pm, &tg, tg2p);
  return b_mn_major ? launch<0, 1, 1, 0, 3>(ta, tb, tc, th, p, 0, stream, &pm, &tg, tg2p)
                    : launch<0, 0, 1, 0, 3>(ta, tb, tc, th, p, 0, stream, &pm, &tg, tg2p);
}
