// Flash attention forward, second generation (D = 128): TWO query tiles per CTA and P kept in tensor memory.
//
// Why a second kernel: the first one (attention_sm100.cu) streams a 64 KB K/V tile for every 128 query rows and stages P through
// 64 KB of shared memory, which leaves room for only two K/V stages — its softmax warps spend 18 % of their samples waiting for S
// (profiles/ncu_source_hotspots_fwd_r1.txt) and the K/V stream alone would need ≈14 TB/s of L2 bandwidth at full tensor rate.
// Here a CTA owns a PAIR of 128-row query tiles of one (batch, head): every K/V tile that lands in shared memory feeds both, and
//   * softmax group g (4 warps, thread = query row = TMEM lane) owns query tile g outright: no running-max hand-off, no named
//     barriers, every row statistic stays in registers;
//   * P_g (bf16) is written with tcgen05.st over the first 64 columns of S_g — the columns the group has just read — and P·V takes
//     its A operand from TMEM (tcgen05.mma … [d], [a_tmem], b_desc): no st.shared, no proxy fence, no P buffers;
//   * TMEM: S_0 | S_1 | O_0 | O_1 = 4 x 128 columns. Per group the chain QKᵀ → softmax → P·V is serial (P aliases S), the two
//     groups run half a period apart so the tensor pipe always has the other tile's MMAs to chew on;
//   * O is rescaled lazily (only when a row max grows by more than 2^8, decided per warp) by the group itself.
// Warp roles: w0 TMA loader · w1 MMA issuer · w2 TMEM allocator · w3..6 softmax group 0 · w7..10 softmax group 1.
// Persistent over a heavy-first list of (query-tile pair, batch·head) items; the K/V rings run on one global tile counter.
#include <type_traits>

#include "tc_common.cuh"
#include "tmap.h"

using namespace tc;

namespace {

constexpr int BQ = 128, D = 128;
constexpr int kThreads = 352;
constexpr float kRescaleThreshold = 8.f;  // log2 units
constexpr uint32_t kTileBytes = 128 * D * 2;            // one Q tile: 32 KB as [2 column chunks][128 rows x 128 B]
constexpr uint32_t kOffK = 2 * kTileBytes;
constexpr uint32_t kRingBytes = 2 * kTileBytes;         // 64 KB per K ring and per V ring
constexpr uint32_t kOffV = kOffK + kRingBytes;
constexpr uint32_t kOffStage = kOffV + kRingBytes;      // per group: [128 rows x 128 B] output staging for one 64-column chunk
constexpr uint32_t kOffBar = kOffStage + 2 * 16384;
constexpr uint32_t kSmem = kOffBar + 320;
// BKV = key tile. 128: one S buffer per group, QKᵀ(t+1) can only follow P·V(t) (P aliases S) — the softmax warps then wait for S a
// quarter of the time (ncu source view). 64: the group's 128 S columns hold TWO buffers, QKᵀ(t+1) runs while softmax(t) is busy,
// and the same 64 KB rings hold four 16 KB stages each.

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

struct Fwd2Params {
  float* lse2;
  int B, S, H, Hkv;
  float scale_log2;  // softmax scale * log2(e)
  int causal;
};

template <int BKV>
__global__ void __launch_bounds__(kThreads, 1)
    flash_fwd2_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_kv,
                      const __grid_constant__ CUtensorMap tmap_o, const Fwd2Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // SWIZZLE_128B operands need 1024-byte aligned tiles
  uint8_t* sQ = smem;
  uint8_t* sK = smem + kOffK;
  uint8_t* sV = smem + kOffV;
  constexpr int NSB = 128 / BKV;                        // S buffers per group
  constexpr int kStages = (int)(kRingBytes / (BKV * D * 2));  // 2 | 4
  constexpr uint32_t kKVBytes = BKV * D * 2;            // one K or V tile as [2 column chunks][BKV rows x 128 B]
  constexpr uint32_t kChunk = BKV * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* q_full = bars;         // 1  both Q tiles of an item have landed (one phase per item)
  uint64_t* q_empty = bars + 1;    // 1  every QKᵀ of the item has retired
  uint64_t* k_full = bars + 2;     // up to 4: K / V rings on one global kv-tile counter
  uint64_t* k_empty = bars + 6;
  uint64_t* v_full = bars + 10;
  uint64_t* v_empty = bars + 14;
  // The per-tile barriers exist once per (group, S buffer): a softmax group can run a whole tile ahead of the MMA thread's
  // bookkeeping when it has two buffers, and a single barrier would then advance two phases between two looks at it — a parity wait
  // cannot tell that from zero phases. With one barrier per buffer, tile n+2 can only signal after tile n has been consumed.
  uint64_t* s_full = bars + 18;    // [group][buffer]: S of that buffer's next tile is complete
  uint64_t* p_full = bars + 22;    // [group][buffer] (4 warp arrivals): P is in TMEM, O_g rescaled if it had to be
  uint64_t* o_done = bars + 26;    // [group][buffer]: P·V of that tile has retired (O_g stable)
  uint64_t* o_free = bars + 30;    // 2  per group (4 warp arrivals): the epilogue has read O_g out of TMEM (one phase per item)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 32);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int npairs = p.S / (2 * BQ);
  const int BH = p.B * p.H;
  const int n_items = npairs * BH;
  auto item = [&](int w, int& pb, int& bh) {  // heavy (late) causal pairs first
    pb = npairs - 1 - w / BH;
    bh = w % BH;
  };
  auto tiles_of = [&](int pb, int g) { return p.causal ? (2 * pb + g + 1) * (128 / BKV) : p.S / BKV; };

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_qkv);
    prefetch_tmap(&tmap_kv);
    prefetch_tmap(&tmap_o);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_done[i], 1);
    }
    for (int i = 0; i < 2; ++i) mbar_init(&o_free[i], 4);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA loader
    if (lane == 0) {
      int kt = 0, it = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
        int pb, bh;
        item(w, pb, bh);
        const int b = bh / p.H, h = bh % p.H;
        const int hk = h / (p.H / p.Hkv);
        const int row0 = b * p.S + pb * 2 * BQ;
        const int col_q = h * D, col_k = (p.H + hk) * D, col_v = (p.H + p.Hkv + hk) * D;
        mbar_wait(q_empty, (it & 1) ^ 1);
        mbar_expect_tx(q_full, 2 * kTileBytes);
#pragma unroll
        for (int g = 0; g < 2; ++g)
#pragma unroll
          for (int c = 0; c < 2; ++c) tma_load_2d(&tmap_qkv, q_full, sQ + g * kTileBytes + c * (128 * 128), col_q + c * 64, row0 + g * BQ);
        const int n_kv = tiles_of(pb, 1);
        for (int t = 0; t < n_kv; ++t, ++kt) {
          const int st = kt % kStages;
          const uint32_t ph = (kt / kStages) & 1;
          const int krow = b * p.S + t * BKV;
          mbar_wait(&k_empty[st], ph ^ 1);
          mbar_expect_tx(&k_full[st], kKVBytes);
#pragma unroll
          for (int c = 0; c < 2; ++c) tma_load_2d(&tmap_kv, &k_full[st], sK + st * kKVBytes + c * kChunk, col_k + c * 64, krow);
          mbar_wait(&v_empty[st], ph ^ 1);
          mbar_expect_tx(&v_full[st], kKVBytes);
#pragma unroll
          for (int c = 0; c < 2; ++c) tma_load_2d(&tmap_kv, &v_full[st], sV + st * kKVBytes + c * kChunk, col_v + c * 64, krow);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = idesc_bf16(BQ, BKV, 0, 0);  // S = Q Kᵀ : both K-major
      constexpr uint32_t idesc_o = idesc_bf16(BQ, D, 0, 1);    // O += P V : P (TMEM) K-major, V MN-major
      constexpr int LA = NSB - 1;  // how many tiles QKᵀ may run ahead of the group's P·V
      int kt0 = 0, it = 0;
      int ng[2] = {0, 0};  // per-group tile counters (global): phases of p_full / o_done, S buffer = counter % NSB
      // S_g[buffer of tile n] = Q_g · K(kt)ᵀ; the caller guarantees K(kt) has landed; the buffer is reusable by program order
      auto issue_qk = [&](int g, int kt, int n) {
        const int st = kt % kStages;
        const int buf = n % NSB;
        const uint32_t a0 = smem_u32(sQ + g * kTileBytes), b0 = smem_u32(sK + st * kKVBytes);
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + g * 128 + buf * BKV, make_smem_desc(a0 + c * (128 * 128) + k * 32, 16, 1024),
                      make_smem_desc(b0 + c * kChunk + k * 32, 16, 1024), idesc_s, (c | k) != 0 ? 1u : 0u);
        umma_commit(&s_full[g * 2 + buf]);
      };
      auto wait_k = [&](int kt) {
        mbar_wait(&k_full[kt % kStages], (kt / kStages) & 1);
        tc_fence_after();
      };
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
        int pb, bh;
        item(w, pb, bh);
        const int n0 = tiles_of(pb, 0), n1 = tiles_of(pb, 1);  // n0 <= n1
        mbar_wait(q_full, it & 1);
        // K tile j is used by group 0 (if j < n0) and group 1, always in this order; its ring slot is released behind group 1's
        auto qk_pair_release = [&](int g, int j) {
          if (g == 1) {
            umma_commit(&k_empty[(kt0 + j) % kStages]);
            if (j + 1 == n1) umma_commit(q_empty);  // the item's last QKᵀ
          }
        };
        for (int j = 0; j <= LA && j < n1; ++j) {
          wait_k(kt0 + j);
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            if (j >= (g == 0 ? n0 : n1)) continue;
            issue_qk(g, kt0 + j, ng[g] + j);
            qk_pair_release(g, j);
          }
        }
        for (int t = 0; t < n1; ++t) {
          const int kt = kt0 + t;
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const int ntl = g == 0 ? n0 : n1;
            if (t >= ntl) continue;
            // P_g(t) is in TMEM (and O_g was rescaled if a row max jumped)
            const int buf = ng[g] % NSB;
            const uint32_t bph = (ng[g] / NSB) & 1;
            mbar_wait(&p_full[g * 2 + buf], bph);
            if (t == 0 && it > 0) mbar_wait(&o_free[g], (it - 1) & 1);  // the previous item's epilogue has drained O_g
            mbar_wait(&v_full[kt % kStages], (kt / kStages) & 1);
            tc_fence_after();
            const uint32_t b0 = smem_u32(sV + (kt % kStages) * kKVBytes);
            const uint32_t tP = tmem_base + g * 128 + buf * BKV;
#pragma unroll
            for (int kk = 0; kk < BKV / 16; ++kk)  // 16 keys = 8 TMEM columns of packed bf16 pairs per step
              umma_bf16_ts(tmem_base + 256 + g * 128, tP + kk * 8, make_smem_desc(b0 + kk * 2048, kChunk, 1024), idesc_o,
                           (t | kk) != 0 ? 1u : 0u);
            umma_commit(&o_done[g * 2 + buf]);
            if (g == 1) umma_commit(&v_empty[kt % kStages]);  // group 1 uses every V tile and comes last
            const int tn = t + 1 + LA;  // the QKᵀ that reuses the S buffer P_g(t) has just been read from
            if (tn < ntl) {
              wait_k(kt0 + tn);
              issue_qk(g, kt0 + tn, ng[g] + 1 + LA);
              qk_pair_release(g, tn);
            }
            ++ng[g];
          }
        }
        kt0 += n1;
      }
    }
  } else if (warp >= 3) {
    // ------------------------------------------------------------------ softmax + epilogue: group g owns query tile g
    const int q = warp & 3;           // TMEM lane quadrant this warp may touch
    const int g = (warp - 3) >> 2;
    const int r = q * 32 + lane;      // query row inside the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint32_t tS0 = tmem_base + g * 128 + lane_addr, tO = tmem_base + 256 + g * 128 + lane_addr;
    uint8_t* stage = smem + kOffStage + g * 16384;
    const uint32_t row_sw = (uint32_t)(lane & 7);
    int ng = 0, it = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
      int pb, bh;
      item(w, pb, bh);
      const int b = bh / p.H, h = bh % p.H;
      const int qt = 2 * pb + g;                 // query tile index in the sequence
      const int n_kv = tiles_of(pb, g);
      const int row0 = b * p.S + qt * BQ;
      float m_used = -INFINITY, l = 0.f;         // reference max of the exponentials (log2 domain) and the row sum relative to it
      const int qpos = qt * BQ + r;              // this row's position in the sequence
      for (int t = 0; t < n_kv; ++t, ++ng) {
        const int buf = ng % NSB;
        const uint32_t tS = tS0 + buf * BKV;
        mbar_wait(&s_full[g * 2 + buf], (ng / NSB) & 1);
        tc_fence_after();
        uint32_t v[BKV];
#pragma unroll
        for (int c = 0; c < BKV / 32; ++c) tmem_ld_32x32b_x32(tS + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&v[c * 32]));
        tmem_ld_wait();
        const bool diag = p.causal && (t + 1) * BKV - 1 > qt * BQ;  // some key of this tile lies behind some query of the tile
        const int lim = qpos - t * BKV;                             // keys j <= lim of this tile are visible to this row
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        if (diag) {
#pragma unroll
          for (int j = 0; j < BKV; ++j)
            if (j <= lim) mx4[j & 3] = fmaxf(mx4[j & 3], __uint_as_float(v[j]));
        } else {
#pragma unroll
          for (int j = 0; j < BKV; ++j) mx4[j & 3] = fmaxf(mx4[j & 3], __uint_as_float(v[j]));
        }
        const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * p.scale_log2;
        // lazy rescale: keep the old reference unless this row's max outgrew it by more than 2^8 (exp2 stays far below bf16 / fp32
        // overflow, and the final 1/l uses the same reference, so the result is exact)
        const bool grow = mx > m_used + kRescaleThreshold;
        float alpha = 1.f;
        if (grow) {
          alpha = fast_exp2(m_used - mx);  // 0 on the first tile (m_used = -inf)
          l *= alpha;
          m_used = mx;
        }
        // P·V of the previous tile: already retired whenever S of this tile exists (QKᵀ is issued behind it), so this wait never
        // blocks — but every completed phase of the barrier is observed, which keeps the parity protocol trivially alias-free
        // (and compute-sanitizer's synccheck, which flags phases nobody waited for, quiet)
        if (t > 0) mbar_wait(&o_done[g * 2 + (ng - 1) % NSB], ((ng - 1) / NSB) & 1);
        if (t > 0 && __any_sync(0xffffffffu, grow)) {
          // O_g *= alpha (rows that did not grow multiply by 1)
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < D / 32; ++c) {
            uint32_t o[32];
            tmem_ld_32x32b_x32(tO + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * alpha);
            tmem_st_32x32b_x32(tO + c * 32, o);
          }
        }
        // p = exp2(s·scale − m) → bf16 pairs → TMEM, over the S columns that are already in registers
        float l4[4] = {0.f, 0.f, 0.f, 0.f};
        auto exp_pack_store = [&](auto masked) {
#pragma unroll
          for (int c = 0; c < BKV / 32; ++c) {
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              float e0 = fast_exp2(fmaf(__uint_as_float(v[c * 32 + j]), p.scale_log2, -m_used));
              float e1 = fast_exp2(fmaf(__uint_as_float(v[c * 32 + j + 1]), p.scale_log2, -m_used));
              if (decltype(masked)::value) {
                if (c * 32 + j > lim) e0 = 0.f;
                if (c * 32 + j + 1 > lim) e1 = 0.f;
              }
              l4[(j >> 1) & 3] += e0 + e1;
              pk[j >> 1] = pack_bf16x2(__float_as_uint(e0), __float_as_uint(e1));
            }
            tmem_st_32x32b_x16(tS + c * 16, pk);
          }
        };
        if (diag) exp_pack_store(std::true_type{});
        else exp_pack_store(std::false_type{});
        l += (l4[0] + l4[1]) + (l4[2] + l4[3]);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[g * 2 + buf]);
      }
      // ---- epilogue of the item: O_g / l → bf16 → staging → TMA store, 64 columns at a time
      mbar_wait(&o_done[g * 2 + (ng - 1) % NSB], ((ng - 1) / NSB) & 1);
      tc_fence_after();
      const float inv_l = 1.f / l;
#pragma unroll 1
      for (int c = 0; c < D / 64; ++c) {
        uint32_t o0[32], o1[32];
        tmem_ld_32x32b_x32(tO + c * 64, o0);
        tmem_ld_32x32b_x32(tO + c * 64 + 32, o1);
        tmem_ld_wait();
        if (c == D / 64 - 1) {  // O_g has left TMEM → the next item's first P·V may overwrite it
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&o_free[g]);
        }
        const uint32_t sbase = smem_u32(stage) + r * 128;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint32_t wv[4], wu[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            wv[e] = pack_bf16x2(__float_as_uint(__uint_as_float(o0[8 * k + 2 * e]) * inv_l), __float_as_uint(__uint_as_float(o0[8 * k + 2 * e + 1]) * inv_l));
            wu[e] = pack_bf16x2(__float_as_uint(__uint_as_float(o1[8 * k + 2 * e]) * inv_l), __float_as_uint(__uint_as_float(o1[8 * k + 2 * e + 1]) * inv_l));
          }
          st_shared_v4(sbase + (((uint32_t)k ^ row_sw) << 4), wv[0], wv[1], wv[2], wv[3]);
          st_shared_v4(sbase + (((uint32_t)(k + 4) ^ row_sw) << 4), wu[0], wu[1], wu[2], wu[3]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmap_o, stage + q * 32 * 128, h * D + c * 64, row0 + q * 32);
          bulk_commit();
          bulk_wait_read<0>();  // this warp's staging rows are rewritten by the next chunk
        }
        __syncwarp();
      }
      p.lse2[((int64_t)bh) * p.S + qt * BQ + r] = m_used + log2f(l);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

// D = 128, S % 256 == 0. Same contract as pb_flash_attn_fwd (attention_sm100.cu).
PB_EXPORT int pb_flash_attn_fwd2(const void* qkv, void* out, float* lse2, int B, int S, int H, int Hkv, float scale, int causal,
                                 cudaStream_t stream) {
  if (S % (2 * BQ) != 0 || H % Hkv != 0) return -1;
  static bool configured = false;
  static int bkv = 0;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(flash_fwd2_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(flash_fwd2_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
    if (e != cudaSuccess) return (int)e;
    const char* ev = getenv("PB_ATTN_FWD2_BKV");
    bkv = (ev && atoi(ev) == 64) ? 64 : 128;
    configured = true;
  }
  const uint64_t rows = (uint64_t)B * S, wqkv = (uint64_t)(H + 2 * Hkv) * D, wo = (uint64_t)H * D;
  CUtensorMap tq, tkv, to;
  int rc = pbhost::cached_tmap(&tq, qkv, rows, wqkv, wqkv, 64, 128, 2);
  if (rc) return rc;
  rc = pbhost::cached_tmap(&tkv, qkv, rows, wqkv, wqkv, 64, (uint32_t)bkv, 2);
  if (rc) return rc;
  rc = pbhost::cached_tmap(&to, out, rows, wo, wo, 64, 32, 2);
  if (rc) return rc;
  Fwd2Params p{lse2, B, S, H, Hkv, scale * 1.4426950408889634f, causal};
  const int items = (S / (2 * BQ)) * B * H;
  const int grid = items < pbhost::num_sms() ? items : pbhost::num_sms();
  if (bkv == 64) flash_fwd2_kernel<64><<<grid, kThreads, kSmem, stream>>>(tq, tkv, to, p);
  else flash_fwd2_kernel<128><<<grid, kThreads, kSmem, stream>>>(tq, tkv, to, p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}
