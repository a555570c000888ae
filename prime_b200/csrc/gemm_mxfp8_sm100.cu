// Block-scaled fp8 (MXFP8: e4m3 values, one UE8M0 power-of-two scale per 32 elements along K) GEMM for sm_100a, plus the
// quantisation kernels that produce its operands.
//
//   C[M,N] (bf16) = Σ_k  (A_q[m,k] · 2^(sfa[m,k/32]-127)) · (B_q[n,k] · 2^(sfb[n,k/32]-127))      fp32 accumulation in TMEM
//
// The scaling is done BY THE TENSOR CORE (tcgen05.mma.kind::mxf8f6f4.block_scale): scale factors are staged in TMEM next to
// the accumulators with tcgen05.cp and consumed per 32-element K block — there is no "promotion" pass over the accumulators
// (reading a 128x192 fp32 tile back every 128 K elements would cost 4x the MMA time at 64 B/clk of TMEM read bandwidth).
//
// Scale-factor memory layout (chosen so the kernel needs NO shuffling between global memory and tcgen05.cp): for a [R, K] operand
// the buffer is  [K/128][ceil(R/128)+1][512 bytes]; one 512-byte "atom" covers 128 rows x 4 K-blocks and stores the byte of
// (row r, block kb) at  (r % 32) * 16 + ((r % 128) / 32) * 4 + kb  — exactly the 32-row x 128-bit pattern tcgen05.cp
// .32x128b.warpx4 broadcasts into the four TMEM lane quadrants. The "+1" atom of padding lets a 192-column tile that starts in
// the middle of an atom always fetch two whole atoms.
//
// Tile 128 x 192 x 128(K bytes): 2 x 192 accumulator columns (double-buffered epilogue overlap) + 2 x (4 + 8) scale-factor columns fit
// the 512-column TMEM; odd N tiles start 64 columns into an SFB atom, which is a +2-column offset of the SFB TMEM address.
// Warp roles as in gemm_sm100.cu: w0 TMA producer, w1 MMA issuer (also issues the tcgen05.cp — cp and mma execute in issue
// order; two alternating 12-column scale slots), w2 TMEM alloc, w4..7 epilogue.
#include <cuda.h>
#include <cuda_fp8.h>

#include <algorithm>

#include "common.cuh"
#include "tc_common.cuh"
#include "tmap.h"

namespace {

using namespace tc;

constexpr int BM = 128, BN = 192, BKB = 128;  // BKB: K elements (= bytes) per stage
constexpr int kAccStages = 2;
constexpr int kThreads = 256;
constexpr uint32_t kABytes = BM * BKB;        // 16 KB
constexpr uint32_t kSfaBytes = 512;           // one atom: 128 rows x 4 K-blocks
constexpr uint32_t kSfbBytes = 1024;          // two atoms
constexpr uint32_t kStagingBytes = 4 * 2 * 4096;
constexpr int kGroupM = 16;
constexpr uint32_t kSfaCol = kAccStages * BN;  // 384: first of two scale-factor slots
constexpr uint32_t kSfCols = 12;               // per slot: SFA 4 columns + SFB 8 columns (two atoms)

// PAIR = 0: one CTA per 128 x 192 tile. PAIR = 1 (cta_group::2): a CTA pair per 256 x 192 tile — each CTA stages its own 128 A rows,
// HALF of B (96 rows), its own SFA atom and the whole SFB; the leader issues M = 256 block-scaled MMAs and the tcgen05.cp copies
// (cta_group::2: every CTA copies from its OWN shared memory into its OWN TMEM). At fp8 rates a stage lasts only 384 clk, so the
// single-CTA ring (4 x 40 KB in flight against ~3000 clk of TMA latency) starves the tensor core; the pair needs 28 KB per CTA per
// stage and keeps 6 stages in flight.
template <int PAIR>
struct Geo {
  static constexpr int kTileM = PAIR ? 2 * BM : BM;
  static constexpr int kBRows = PAIR ? BN / 2 : BN;
  static constexpr uint32_t kBBytes = kBRows * BKB;  // 12 KB | 24 KB
  static constexpr uint32_t kStageBytes = kABytes + kBBytes + kSfaBytes + kSfbBytes;
  static constexpr uint32_t kStagePitch = (kStageBytes + 1023) / 1024 * 1024;  // A/B tiles stay 1024-aligned (SWIZZLE_128B)
  static constexpr int kStages = PAIR ? 6 : 4;
  static constexpr uint32_t kSmemBytes = kStages * kStagePitch + kStagingBytes + 1024 + 256;
};
static_assert(Geo<0>::kSmemBytes <= 232448 && Geo<1>::kSmemBytes <= 232448, "shared memory budget");

// kind::mxf8f6f4 block-scaled instruction descriptor: a/b = e4m3 (0), K-major both, N>>3 at 17, UE8M0 scales (bit 23),
// M>>4 at 24; scale-factor byte selectors: b_sf_id at 4, a_sf_id at 29
__host__ __device__ constexpr uint32_t make_idesc_mx(uint32_t sf_id, int m) {
  return (sf_id << 4) | ((uint32_t)(BN >> 3) << 17) | (1u << 23) | ((uint32_t)(m >> 4) << 24) | (sf_id << 29);
}

template <int PAIR>
__device__ __forceinline__ void umma_mxfp8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum, uint32_t tsfa,
                                           uint32_t tsfb) {
  if (PAIR) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum), "r"(tsfa), "r"(tsfb)
        : "memory");
  } else {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum), "r"(tsfa), "r"(tsfb)
        : "memory");
  }
}
// 32 rows x 16 bytes of shared memory → the same 4 TMEM columns of all four lane quadrants
template <int PAIR>
__device__ __forceinline__ void tmem_cp_sf(uint32_t taddr, uint32_t saddr) {
  // no-swizzle K-major descriptor: 8-row x 16-byte core matrices, 128 bytes apart
  const uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
  if (PAIR) asm volatile("tcgen05.cp.cta_group::2.32x128b.warpx4 [%0], %1;" ::"r"(taddr), "l"(d) : "memory");
  else asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(taddr), "l"(d) : "memory");
}

__device__ __forceinline__ void tile_coords(int tile, int tiles_m, int tiles_n, int& tm, int& tn) {
  const int per_group = kGroupM * tiles_n;
  const int g = tile / per_group;
  const int first_m = g * kGroupM;
  const int gsz = min(kGroupM, tiles_m - first_m);
  const int r = tile - g * per_group;
  tm = first_m + r % gsz;
  tn = r / gsz;
}

struct MxParams {
  int M, N, K;
  int atoms_m, atoms_n;  // atoms per K/128 row of the scale buffers: [K/128][atoms][512 B]
};

// tmap_sfa / tmap_sfb: the scale buffers as [K/128 · atoms] rows of 128 uint32 (512 B), no swizzle; boxes of 1 / 2 rows
template <int PAIR>
__global__ void __launch_bounds__(kThreads, 1)
    gemm_mxfp8_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                      const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_sfa,
                      const __grid_constant__ CUtensorMap tmap_sfb, const MxParams p) {
  using G = Geo<PAIR>;
  constexpr int kStages = G::kStages;
  constexpr uint32_t kStagePitch = G::kStagePitch;
  constexpr uint32_t kBBytes = G::kBBytes;
  constexpr int kTileM = G::kTileM;
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
  const int sched_id = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int sched_n = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + kStages * kStagePitch;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + kStagingBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;
  uint64_t* tempty_bar = tfull_bar + kAccStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + kAccStages);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (p.M + kTileM - 1) / kTileM, tiles_n = (p.N + BN - 1) / BN;
  const int num_kb = p.K / BKB;
  const int num_tiles = tiles_m * tiles_n;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    prefetch_tmap(&tmap_c);
    prefetch_tmap(&tmap_sfa);
    prefetch_tmap(&tmap_sfb);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < kAccStages; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], PAIR ? 8 : 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (PAIR) tmem_alloc_pair(tmem_slot, 512);
    else tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ producer: operands and scale atoms by TMA
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = sched_id; tile < num_tiles; tile += sched_n) {
        int tm, tn;
        tile_coords(tile, tiles_m, tiles_n, tm, tn);
        const int m0 = tm * kTileM + (int)crank * BM;           // this CTA's A rows
        const int n0 = tn * BN + (int)crank * G::kBRows;        // this CTA's share of the B rows
        const int atom_m = tm * (kTileM / 128) + (int)crank;    // this CTA's SFA atom
        const int atom_n0 = (tn * BN) >> 7;                     // first of the tile's two SFB atoms (every CTA needs both)
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kStagePitch;
          uint8_t* sb = sa + kABytes;
          uint8_t* ssfa = sb + kBBytes;
          uint8_t* ssfb = ssfa + kSfaBytes;
          if (!PAIR || crank == 0) mbar_expect_tx(&full_bar[stage], (PAIR ? 2u : 1u) * G::kStageBytes);
          auto load = [&](const CUtensorMap* m, void* dst, int c0, int c1) {
            if (PAIR) tma_load_2d_pair(m, &full_bar[stage], dst, c0, c1);
            else tma_load_2d(m, &full_bar[stage], dst, c0, c1);
          };
          load(&tmap_a, sa, kb * BKB, m0);  // box {128 k-bytes, 128 rows}
          load(&tmap_b, sb, kb * BKB, n0);  // box {128 k-bytes, 192 | 96 rows}
          load(&tmap_sfa, ssfa, 0, kb * p.atoms_m + atom_m);   // box {128 words, 1 atom}
          load(&tmap_sfb, ssfb, 0, kb * p.atoms_n + atom_n0);  // box {128 words, 2 atoms}
          if (++stage == kStages) stage = 0, phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA of a pair)
    if (lane == 0 && crank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      uint32_t fillno = 0;
      // The scales of a stage go into one of two 12-column TMEM slots right before its MMAs (tcgen05.cp and tcgen05.mma execute
      // in issue order). Measured alternatives (profiles/mxfp8_gemm_bench_r1.json): copying one stage AHEAD was slower (it costs a
      // ring stage of TMA lookahead), one slot vs two made no difference, and dropping the copies altogether gains ~15 % — the
      // three copies per 384-clk stage occupy the tensor pipe; moving them to tcgen05.st from idle warps is the known next step.
      auto copy_sf = [&](int st, uint32_t slot) {
        const uint32_t ssfa = smem_u32(smem + st * kStagePitch) + kABytes + kBBytes;
        const uint32_t ssfb = ssfa + kSfaBytes;
        const uint32_t t = tmem_base + kSfaCol + slot * kSfCols;
        tmem_cp_sf<PAIR>(t, ssfa);
        tmem_cp_sf<PAIR>(t + 4, ssfb);
        tmem_cp_sf<PAIR>(t + 8, ssfb + 512);
      };
      for (int tile = sched_id; tile < num_tiles; tile += sched_n) {
        int tm, tn;
        tile_coords(tile, tiles_m, tiles_n, tm, tn);
        // odd N tiles begin 64 columns into their first SFB atom = 2 TMEM columns
        const uint32_t sfb_shift = (uint32_t)(((tn * BN) & 127) >> 5);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          copy_sf(stage, fillno & 1);
          const uint32_t sa = smem_u32(smem + stage * kStagePitch);
          const uint32_t sb = sa + kABytes;
          const uint32_t tsfa = tmem_base + kSfaCol + (fillno & 1) * kSfCols;
          const uint32_t tsfb = tsfa + 4 + sfb_shift;
#pragma unroll
          for (int k = 0; k < BKB / 32; ++k) {  // UMMA_K = 32 elements = 32 bytes inside the 128 B swizzle row
            const uint64_t adesc = make_smem_desc(sa + k * 32, 16, 1024);
            const uint64_t bdesc = make_smem_desc(sb + k * 32, 16, 1024);
            umma_mxfp8<PAIR>(tmem_d, adesc, bdesc, make_idesc_mx((uint32_t)k, kTileM), (kb | k) != 0 ? 1u : 0u, tsfa, tsfb);
          }
          if (PAIR) umma_commit_pair(&empty_bar[stage]);
          else umma_commit(&empty_bar[stage]);
          if (++stage == kStages) stage = 0, phase ^= 1;
          ++fillno;
        }
        if (PAIR) umma_commit_pair(&tfull_bar[acc]);
        else umma_commit(&tfull_bar[acc]);
        if (++acc == kAccStages) acc = 0, acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue: TMEM → bf16 → swizzled slab → TMA store
    const int q = warp & 3;
    uint8_t* my_stage = staging + q * 8192;
    const uint32_t row_sw = (uint32_t)(lane & 7);
    int acc = 0;
    uint32_t acc_phase = 0;
    int buf = 0;
    for (int tile = sched_id; tile < num_tiles; tile += sched_n) {
      int tm, tn;
      tile_coords(tile, tiles_m, tiles_n, tm, tn);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int row0 = tm * kTileM + (int)crank * BM + q * 32;
      const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int g = 0; g < BN / 64; ++g) {
        uint32_t r0[32], r1[32];
        tmem_ld_32x32b_x32(taddr + g * 64, r0);
        tmem_ld_32x32b_x32(taddr + g * 64 + 32, r1);
        if (lane == 0) bulk_wait_read<1>();
        __syncwarp();
        tmem_ld_wait();
        const uint32_t sbase = smem_u32(my_stage + buf * 4096) + lane * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          st_shared_v4(sbase + ((j ^ row_sw) << 4), pack_bf16x2(r0[8 * j], r0[8 * j + 1]), pack_bf16x2(r0[8 * j + 2], r0[8 * j + 3]),
                       pack_bf16x2(r0[8 * j + 4], r0[8 * j + 5]), pack_bf16x2(r0[8 * j + 6], r0[8 * j + 7]));
#pragma unroll
        for (int j = 0; j < 4; ++j)
          st_shared_v4(sbase + (((j + 4) ^ row_sw) << 4), pack_bf16x2(r1[8 * j], r1[8 * j + 1]), pack_bf16x2(r1[8 * j + 2], r1[8 * j + 3]),
                       pack_bf16x2(r1[8 * j + 4], r1[8 * j + 5]), pack_bf16x2(r1[8 * j + 6], r1[8 * j + 7]));
        fence_proxy_async();
        __syncwarp();
        const int col0 = tn * BN + g * 64;
        if (lane == 0 && row0 < p.M && col0 < p.N) tma_store_2d(&tmap_c, my_stage + buf * 4096, col0, row0);
        if (lane == 0) bulk_commit();
        buf ^= 1;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_leader(&tempty_bar[acc]);
        else mbar_arrive(&tempty_bar[acc]);
      }
      if (++acc == kAccStages) acc = 0, acc_phase ^= 1;
    }
    if (lane == 0) bulk_wait_read<0>();
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all();
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------- quantisers
__device__ __forceinline__ size_t sf_offset(int r, int kblk, int atoms_r) {
  // kblk = index of the 32-element block along K
  return ((size_t)(kblk >> 2) * atoms_r + (r >> 7)) * 512 + (size_t)((r & 31) * 16 + ((r & 127) >> 5) * 4 + (kblk & 3));
}
// smallest power of two s with amax / s <= 448 (e4m3 max), as a biased UE8M0 exponent; and 1/s
__device__ __forceinline__ uint32_t ue8m0_for(float amax, float& inv_scale) {
  const uint32_t bits = __float_as_uint(amax * (1.0f / 448.0f));
  uint32_t e = (bits >> 23) & 0xFF;
  if (bits & 0x7FFFFF) e += 1;
  e = min(max(e, 1u), 254u);
  inv_scale = __uint_as_float((254u - e) << 23);
  return e;
}
__device__ __forceinline__ uint32_t pack_e4m3x4(float a, float b, float c, float d) {
  const uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E4M3);
  const uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(c, d), __NV_SATFINITE, __NV_E4M3);
  return lo | (hi << 16);
}

// x bf16 [R, C] (row stride ld) → q e4m3 [R, C] (dense) + scales for blocks of 32 along C. One thread per 32-element block.
__global__ void __launch_bounds__(256) quantize_mxfp8_kernel(const __nv_bfloat16* __restrict__ x, int64_t ld, uint8_t* __restrict__ q,
                                                             uint8_t* __restrict__ sf, int R, int C, int atoms_r) {
  const int blocks_per_row = C >> 5;
  const int64_t total = (int64_t)R * blocks_per_row;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / blocks_per_row), kblk = (int)(i % blocks_per_row);
    const pb::bf16x8* src = reinterpret_cast<const pb::bf16x8*>(x + (int64_t)r * ld + kblk * 32);
    float f[32];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const pb::bf16x8 v = pb::ldg_stream(src + j);
      float t[8];
      pb::unpack8(v, t);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[j * 8 + e] = t[e];
    }
    float amax = 0.f;
#pragma unroll
    for (int e = 0; e < 32; ++e) amax = fmaxf(amax, fabsf(f[e]));
    float inv;
    const uint32_t e8 = ue8m0_for(amax, inv);
    uint4 o[2];
    uint32_t* ow = reinterpret_cast<uint32_t*>(o);
#pragma unroll
    for (int j = 0; j < 8; ++j) ow[j] = pack_e4m3x4(f[4 * j] * inv, f[4 * j + 1] * inv, f[4 * j + 2] * inv, f[4 * j + 3] * inv);
    uint4* dst = reinterpret_cast<uint4*>(q + (int64_t)r * C + kblk * 32);
    dst[0] = o[0];
    dst[1] = o[1];
    sf[sf_offset(r, kblk, atoms_r)] = (uint8_t)e8;
  }
}

// x bf16 [R, C] → qᵀ e4m3 [C, R] with scales for blocks of 32 along R (the contraction dim of the transposed use).
// Block: 128 rows x 32 columns through shared memory; thread t owns column (t % 32), row block (t / 32).
__global__ void __launch_bounds__(128) quantize_mxfp8_t_kernel(const __nv_bfloat16* __restrict__ x, int64_t ld, uint8_t* __restrict__ q,
                                                               uint8_t* __restrict__ sf, int R, int C, int atoms_c) {
  __shared__ float tile[128][33];
  const int r0 = blockIdx.y * 128, c0 = blockIdx.x * 32;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int rr = w; rr < 128; rr += 4) {
    const int r = r0 + rr, c = c0 + lane;
    tile[rr][lane] = (r < R && c < C) ? __bfloat162float(x[(int64_t)r * ld + c]) : 0.f;
  }
  __syncthreads();
  const int c = c0 + lane;
  const int rb = r0 + w * 32;
  if (c >= C || rb >= R) return;
  float f[32];
  float amax = 0.f;
#pragma unroll
  for (int e = 0; e < 32; ++e) {
    f[e] = tile[w * 32 + e][lane];
    amax = fmaxf(amax, fabsf(f[e]));
  }
  float inv;
  const uint32_t e8 = ue8m0_for(amax, inv);
  uint4 o[2];
  uint32_t* ow = reinterpret_cast<uint32_t*>(o);
#pragma unroll
  for (int j = 0; j < 8; ++j) ow[j] = pack_e4m3x4(f[4 * j] * inv, f[4 * j + 1] * inv, f[4 * j + 2] * inv, f[4 * j + 3] * inv);
  uint4* dst = reinterpret_cast<uint4*>(q + (int64_t)c * R + rb);
  dst[0] = o[0];
  dst[1] = o[1];
  sf[sf_offset(c, rb >> 5, atoms_c)] = (uint8_t)e8;
}

int g_mx_pair_mode = -1;  // -1: from PB_MXFP8_PAIR (default on)

template <int PAIR>
int launch_mx(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tsa, const CUtensorMap& tsb,
              const MxParams& p, int max_ctas, cudaStream_t stream) {
  using G = Geo<PAIR>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_mxfp8_kernel<PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  int grid = pbhost::num_sms();
  if (max_ctas > 0 && max_ctas < grid) grid = max_ctas;
  const int tiles = ((p.M + G::kTileM - 1) / G::kTileM) * ((p.N + BN - 1) / BN);
  if (!PAIR) {
    if (tiles < grid) grid = tiles;
    gemm_mxfp8_kernel<0><<<grid, kThreads, G::kSmemBytes, stream>>>(ta, tb, tc, tsa, tsb, p);
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
    cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_mxfp8_kernel<1>, ta, tb, tc, tsa, tsb, p);
    if (e != cudaSuccess) return (int)e;
  }
  PB_CHECK_LAUNCH();
  return 0;
}

}  // namespace

// 0 = single-CTA 128x192 tiles, 1 = CTA-pair 256x192 tiles (cta_group::2, default), -1 = re-read PB_MXFP8_PAIR. Returns the old mode.
PB_EXPORT int pb_gemm_mxfp8_set_pair_mode(int mode) {
  const int old = g_mx_pair_mode;
  g_mx_pair_mode = mode;
  return old;
}

// Bytes of the scale-factor buffer for an operand with `rows` rows and K contraction elements (K % 128 == 0).
PB_EXPORT int64_t pb_mxfp8_sf_bytes(int rows, int K) { return (int64_t)(K / 128) * ((rows + 127) / 128 + 1) * 512; }

// transpose = 0: q [R, C] with blocks along C (C % 128 == 0).   transpose = 1: q [C, R] with blocks along R (R % 128 == 0).
PB_EXPORT int pb_quantize_mxfp8(const void* x, int64_t ld, void* q, void* sf, int R, int C, int transpose, cudaStream_t stream) {
  if (R <= 0 || C <= 0) return 0;
  if ((ld % 8) || ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(q)) & 15)) return -1;
  if (!transpose) {
    if (C % 128) return -2;
    const int64_t total = (int64_t)R * (C / 32);
    const int grid = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
    quantize_mxfp8_kernel<<<grid, 256, 0, stream>>>((const __nv_bfloat16*)x, ld, (uint8_t*)q, (uint8_t*)sf, R, C, (R + 127) / 128 + 1);
  } else {
    if (R % 128) return -2;
    dim3 grid((C + 31) / 32, R / 128);
    quantize_mxfp8_t_kernel<<<grid, 128, 0, stream>>>((const __nv_bfloat16*)x, ld, (uint8_t*)q, (uint8_t*)sf, R, C, (C + 127) / 128 + 1);
  }
  PB_CHECK_LAUNCH();
  return 0;
}

// C[M,N] bf16 = A_q[M,K] · B_q[N,K]ᵀ with per-32 UE8M0 scales (buffers laid out as described at the top of this file).
PB_EXPORT int pb_gemm_mxfp8(const void* Aq, const void* sfa, const void* Bq, const void* sfb, void* C, int M, int N, int K, int ldc,
                            int max_ctas, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (K % 128 || ldc % 8) return -1;
  if ((reinterpret_cast<uintptr_t>(Aq) | reinterpret_cast<uintptr_t>(Bq) | reinterpret_cast<uintptr_t>(C) | reinterpret_cast<uintptr_t>(sfa) |
       reinterpret_cast<uintptr_t>(sfb)) & 15)
    return -2;
  if (g_mx_pair_mode < 0) {
    const char* e = getenv("PB_MXFP8_PAIR");
    g_mx_pair_mode = e ? (atoi(e) != 0) : 1;
  }
  const int pair = g_mx_pair_mode && M > BM;
  const int atoms_m = (M + 127) / 128 + 1, atoms_n = (N + 127) / 128 + 1, kblocks = K / 128;
  CUtensorMap ta, tb, tc, tsa, tsb;
  int rc;
  if ((rc = pbhost::cached_tmap(&ta, Aq, (uint64_t)M, (uint64_t)K, (uint64_t)K, BKB, BM, 1))) return rc;
  if ((rc = pbhost::cached_tmap(&tb, Bq, (uint64_t)N, (uint64_t)K, (uint64_t)K, BKB, pair ? BN / 2 : BN, 1))) return rc;
  if ((rc = pbhost::cached_tmap(&tc, C, (uint64_t)M, (uint64_t)N, (uint64_t)ldc, 64, 32, 2))) return rc;
  if ((rc = pbhost::cached_tmap(&tsa, sfa, (uint64_t)kblocks * atoms_m, 128, 128, 128, 1, -4))) return rc;
  if ((rc = pbhost::cached_tmap(&tsb, sfb, (uint64_t)kblocks * atoms_n, 128, 128, 128, 2, -4))) return rc;
  MxParams p{M, N, K, atoms_m, atoms_n};
  return pair ? launch_mx<1>(ta, tb, tc, tsa, tsb, p, max_ctas, stream) : launch_mx<0>(ta, tb, tc, tsa, tsb, p, max_ctas, stream);
}
