// Flash attention for sm_100a on the fused QKV activation — tcgen05 S = QKᵀ and O += P·V with TMEM accumulators.
//
// Input : qkv [B, S, (H + 2·Hkv)·D] bf16 (RoPE already applied in place), heads laid out [Q heads | K heads | V heads]
// Output: out [B, S, H·D] bf16 (the layout the output-projection GEMM consumes — no transposes, no copies)
//         lse2 [B, H, S] fp32 = log2-domain log-sum-exp of (scale·log2e·s), consumed by the backward kernels
//
// Persistent: one CTA per SM walks a static work list of (128-row query block, batch, head) items, heavy (late) causal blocks
// first. All rings (K/V smem, S in TMEM, P in smem) run on one global tile counter, so the TMA loader and the MMA issuer flow
// straight into the next item while the softmax warps are still normalising/storing the previous one: the per-item pipeline
// fill/drain (≈7 µs measured with one CTA per item — more than the math of a short causal item) is paid once per CTA.
// Warp roles: w0 TMA loader · w1 MMA issuer · w2 TMEM allocator · w3..10 softmax in TWO GROUPS that leapfrog over the kv tiles
//   (group = global tile parity = S/P stage). A 128×128 fp32 score tile takes 1024 clk to read out of TMEM (64 B/clk/SM), 1024 clk
//   of MUFU.EX2 and 1024 clk of MMAs; with every softmax warp on the same tile those phases run back to back, with two groups on
//   consecutive tiles one group's TMEM read overlaps the other's exp2/pack/st.shared. Rows are per-thread (one TMEM lane), so the
//   only cross-group state is the running row max: the owner of tile t publishes m(t) through 1 KB of smem and a 64-thread named
//   barrier (warp q of one group ↔ warp q of the other) BEFORE its exp phase, so the hand-off is off the critical path. Each
//   group keeps its own partial row sum relative to the max it last saw; they are merged in the epilogue.
//   S is double-buffered in TMEM (2 × 128 cols) so QKᵀ of tile t+1 overlaps the softmax of tile t;
//   P is written bf16 into 128B-swizzled smem (double-buffered) as the K-major A operand of P·V;
//   O lives in TMEM (D cols); it is only rescaled when a row max grows by more than 2^8 (lazy rescale),
//   the final 1/l normalisation uses the same stale max, so the result is exact.
#include <type_traits>

#include "tc_common.cuh"
#include "tmap.h"

using namespace tc;

namespace {

constexpr int BQ = 128, BKV = 128;
constexpr int kThreads = 352;  // w0 TMA · w1 MMA · w2 TMEM alloc · w3..10 softmax: group (w-3)/4, TMEM lane quadrant w%4
constexpr float kRescaleThreshold = 8.f;  // log2 units

template <int D>
struct FwdCfg {
  static constexpr int kChunks = D / 64;                 // 64-element (128 B) column chunks per row
  static constexpr uint32_t kQBytes = BQ * D * 2;
  static constexpr uint32_t kKVBytes = BKV * D * 2;
  static constexpr uint32_t kPBytes = BQ * BKV * 2;      // 32 KB (two [128 x 64] swizzled blocks)
  static constexpr uint32_t kOffK = kQBytes;
  static constexpr uint32_t kOffV = kOffK + 2 * kKVBytes;
  static constexpr uint32_t kOffP = kOffV + 2 * kKVBytes;
  static constexpr uint32_t kOffBar = kOffP + 2 * kPBytes;
  static constexpr uint32_t kOffXchg = kOffBar + 256;          // mrow[2][128] running max per tile parity + lrow[2][128] partial sums
  static constexpr uint32_t kSmem = kOffXchg + 2048;           // 226.25 KB at D=128: no room for alignment slack, so the
                                                               // dynamic smem base itself is declared 1024-aligned
};

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct FwdParams {
  float* lse2;
  int B, S, H, Hkv;
  float scale_log2;  // softmax scale * log2(e)
  int causal;
};

template <int D>
__global__ void __launch_bounds__(kThreads, 1)
    flash_fwd_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_o,
                     const FwdParams p) {
  using C = FwdCfg<D>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // SWIZZLE_128B operands need 1024-byte aligned tiles
  uint8_t* sQ = smem;
  uint8_t* sK = smem + C::kOffK;
  uint8_t* sV = smem + C::kOffV;
  uint8_t* sP = smem + C::kOffP;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kOffBar);
  uint64_t* q_full = bars;            // 1   producer → MMA, one phase per work item
  uint64_t* q_empty = bars + 1;       // 1   MMA → producer: every QKᵀ of the item has retired, Q smem reusable
  uint64_t* k_full = bars + 2;        // 2   K/V rings, S and P buffers run on ONE global tile counter across work items
  uint64_t* k_empty = bars + 4;       // 2
  uint64_t* v_full = bars + 6;        // 2
  uint64_t* v_empty = bars + 8;       // 2
  uint64_t* s_full = bars + 10;       // 2
  uint64_t* s_empty = bars + 12;      // 2 (4 warp arrivals: stage i ↔ softmax group i)
  uint64_t* p_full = bars + 14;       // 2 (4 warp arrivals)
  uint64_t* pv_done = bars + 16;      // 2 (commit of P·V for a tile → P buffer free, O stable)
  uint64_t* o_free = bars + 18;       // 1 (8 warp arrivals): the epilogue has read O out of TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);
  float* xchg = reinterpret_cast<float*>(smem + C::kOffXchg);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nqb = p.S / BQ;
  const int BH = p.B * p.H;
  const int n_items = nqb * BH;
  // work item w → (query block, batch·head); heavy (late) causal blocks first, round-robin over the persistent CTAs
  auto item = [&](int w, int& qb, int& bh) {
    qb = nqb - 1 - w / BH;
    bh = w % BH;
  };

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_qkv);
    prefetch_tmap(&tmap_o);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    mbar_init(o_free, 8);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 4);  // stage i is always served by softmax group i (4 warps)
      mbar_init(&p_full[i], 4);
      mbar_init(&pv_done[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_O = tmem_base + 256;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA loader (runs ahead across work items)
    if (lane == 0) {
      int g = 0;  // global kv-tile counter
      int it = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
        int qb, bh;
        item(w, qb, bh);
        const int b = bh / p.H, h = bh % p.H, hk = h / (p.H / p.Hkv);
        const int n_kv = p.causal ? qb + 1 : p.S / BKV;
        const int row0 = b * p.S + qb * BQ;
        const int col_q = h * D, col_k = (p.H + hk) * D, col_v = (p.H + p.Hkv + hk) * D;
        mbar_wait(q_empty, (it & 1) ^ 1);
        mbar_expect_tx(q_full, C::kQBytes);
#pragma unroll
        for (int c = 0; c < C::kChunks; ++c) tma_load_2d(&tmap_qkv, q_full, sQ + c * (BQ * 128), col_q + c * 64, row0);
        for (int t = 0; t < n_kv; ++t, ++g) {
          const int st = g & 1;
          const uint32_t ph = (g >> 1) & 1;
          const int krow = b * p.S + t * BKV;
          mbar_wait(&k_empty[st], ph ^ 1);
          mbar_expect_tx(&k_full[st], C::kKVBytes);
#pragma unroll
          for (int c = 0; c < C::kChunks; ++c)
            tma_load_2d(&tmap_qkv, &k_full[st], sK + st * C::kKVBytes + c * (BKV * 128), col_k + c * 64, krow);
          mbar_wait(&v_empty[st], ph ^ 1);
          mbar_expect_tx(&v_full[st], C::kKVBytes);
#pragma unroll
          for (int c = 0; c < C::kChunks; ++c)
            tma_load_2d(&tmap_qkv, &v_full[st], sV + st * C::kKVBytes + c * (BKV * 128), col_v + c * 64, krow);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = idesc_bf16(BQ, BKV, 0, 0);  // S = Q Kᵀ : both K-major
      constexpr uint32_t idesc_o = idesc_bf16(BQ, D, 0, 1);    // O += P V : P K-major, V MN-major
      int g = 0;
      int it = 0;
      auto issue_s = [&](int gg, bool last_of_item) {
        const int st = gg & 1;
        const uint32_t ph = (gg >> 1) & 1;
        mbar_wait(&k_full[st], ph);
        mbar_wait(&s_empty[st], ph ^ 1);
        tc_fence_after();
        const uint32_t a0 = smem_u32(sQ), b0 = smem_u32(sK + st * C::kKVBytes);
#pragma unroll
        for (int c = 0; c < C::kChunks; ++c)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + st * BKV, make_smem_desc(a0 + c * (BQ * 128) + k * 32, 16, 1024),
                      make_smem_desc(b0 + c * (BKV * 128) + k * 32, 16, 1024), idesc_s, (c | k) != 0 ? 1u : 0u);
        umma_commit(&k_empty[st]);
        umma_commit(&s_full[st]);
        if (last_of_item) umma_commit(q_empty);  // the producer may now overwrite Q with the next item's block
      };
      // can QKᵀ of tile gg be issued right now without blocking? (its K tile has landed and its S stage has been drained)
      auto s_ready = [&](int gg) {
        const int st = gg & 1;
        const uint32_t ph = (gg >> 1) & 1;
        return mbar_test_wait(&k_full[st], ph) && mbar_test_wait(&s_empty[st], ph ^ 1);
      };
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
        int qb, bh;
        item(w, qb, bh);
        const int n_kv = p.causal ? qb + 1 : p.S / BKV;
        mbar_wait(q_full, it & 1);
        issue_s(g, n_kv == 1);
        int s_next = 1;  // next tile (item-local) whose QKᵀ has not been issued yet
        for (int t = 0; t < n_kv; ++t, ++g) {
          if (s_next <= t + 1 && s_next < n_kv) {  // one tile ahead is mandatory (overlaps the softmax of tile t)
            issue_s(g + (s_next - t), s_next + 1 == n_kv);
            ++s_next;
          }
          const int st = g & 1;
          const uint32_t ph = (g >> 1) & 1;
          mbar_wait(&v_full[st], ph);
          // while P of tile t is being produced, opportunistically run QKᵀ TWO tiles ahead: S stage g&1 is free as soon as its
          // softmax group has pulled tile g into registers, so when K(g+2) has landed too the group finds its next tile waiting.
          // Never block on it here — a blocking wait in front of P·V(t) measurably slows the whole pipeline down.
          SpinGuard guard;
          while (!mbar_test_wait(&p_full[st], ph)) {  // test_wait: try_wait may park the thread for a time slice, and the opportunistic issue below must not wait for that
            guard.tick();
            if (s_next == t + 2 && s_next < n_kv && s_ready(g + 2)) {
              issue_s(g + 2, s_next + 1 == n_kv);
              ++s_next;
            }
          }
          if (t == 0 && it > 0) mbar_wait(o_free, (it - 1) & 1);  // previous item's epilogue has drained O
          tc_fence_after();
          const uint32_t a0 = smem_u32(sP + st * C::kPBytes), b0 = smem_u32(sV + st * C::kKVBytes);
#pragma unroll
          for (int kk = 0; kk < BKV / 16; ++kk)
            umma_bf16(tmem_O, make_smem_desc(a0 + (kk >> 2) * (BQ * 128) + (kk & 3) * 32, 16, 1024),
                      make_smem_desc(b0 + kk * 2048, BKV * 128, 1024), idesc_o, (t | kk) != 0 ? 1u : 0u);
          umma_commit(&v_empty[st]);
          umma_commit(&pv_done[st]);
        }
      }
    }
  } else if (warp >= 3) {
    // ------------------------------------------------------------------ softmax + epilogue
    const int q = warp & 3;            // TMEM lane quadrant this warp may touch
    const int grp = (warp - 3) >> 2;   // softmax group: owns the tiles of this global parity, and this half of the O columns
    const int r = q * 32 + lane;       // query row inside the block == TMEM lane
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint32_t row_sw = (uint32_t)(r & 7);
    constexpr int kOC = D / 64;        // 32-column O chunks per warp in the epilogue
    float* mrow = xchg;                // [2][128]
    float* lrow = xchg + 256;          // [2][128]
    // named barriers 1..8: id(from) is signalled by group `from`'s warp q and awaited by the other group's warp q; 9..12: pair sync
    auto publish = [&]() { asm volatile("bar.arrive %0, 64;" ::"r"(1 + q + 4 * grp) : "memory"); };
    auto consume = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + q + 4 * (grp ^ 1)) : "memory"); };
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(9 + q) : "memory"); };
    int g = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
      int qb, bh;
      item(w, qb, bh);
      const int b = bh / p.H, h = bh % p.H;
      const int n_kv = p.causal ? qb + 1 : p.S / BKV;
      const int row0 = b * p.S + qb * BQ;
      const int col_q = h * D;
      float m_ref = -INFINITY, l = 0.f;  // this group's partial row sum, relative to the last running max it has seen
      for (int t = 0; t < n_kv; ++t, ++g) {
        if ((g & 1) != grp) continue;
        const int st = g & 1;
        const uint32_t ph = (g >> 1) & 1;
        mbar_wait(&s_full[st], ph);
        tc_fence_after();
        const uint32_t tS = tmem_base + st * BKV + lane_addr;
        const bool diag = p.causal && (t == qb);
        // one TMEM read per tile: the whole 128-column score row lives in registers (4 loads in flight, one wait)
        uint32_t v[BKV];
#pragma unroll
        for (int c = 0; c < BKV / 32; ++c) tmem_ld_32x32b_x32(tS + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&v[c * 32]));
        tmem_ld_wait();
        // S stage consumed → QKᵀ two tiles ahead may overwrite it while we do the math
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[st]);
        float mx = -INFINITY;
        if (diag) {
#pragma unroll
          for (int j = 0; j < BKV; ++j)
            if (j <= r) mx = fmaxf(mx, __uint_as_float(v[j]));
        } else {
          float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // 4 independent chains: FMNMX latency, not throughput
#pragma unroll
          for (int j = 0; j < BKV; j += 4) {
            m4[0] = fmaxf(m4[0], __uint_as_float(v[j]));
            m4[1] = fmaxf(m4[1], __uint_as_float(v[j + 1]));
            m4[2] = fmaxf(m4[2], __uint_as_float(v[j + 2]));
            m4[3] = fmaxf(m4[3], __uint_as_float(v[j + 3]));
          }
          mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        }
        mx *= p.scale_log2;
        // running max after the previous tile (owned by the other group), then publish ours before the long exp phase
        float m_prev = -INFINITY;
        if (t > 0) {
          consume();
          m_prev = mrow[((g - 1) & 1) * 128 + r];
        }
        // lazy rescale (warp-uniform: the tcgen05.ld/st of the O rescale are warp-collective)
        const bool any_grow = __any_sync(0xffffffffu, (mx - m_prev) > kRescaleThreshold);
        const float m_new = any_grow ? fmaxf(m_prev, mx) : m_prev;
        mrow[(g & 1) * 128 + r] = m_new;
        publish();
        const float alpha = (any_grow && m_prev != -INFINITY) ? fast_exp2(m_prev - m_new) : (any_grow ? 0.f : 1.f);
        if (m_ref != m_new) {  // bring this group's partial sum to the new reference (exp2(-inf) = 0 covers the first tile)
          l *= fast_exp2(m_ref - m_new);
          m_ref = m_new;
        }
        // the P buffer of this stage was last read by P·V two tiles ago (possibly of the previous work item)
        if (g >= 2) mbar_wait(&pv_done[st], ph ^ 1);
        // p = exp2(s·scale − m) → bf16 → 128B-swizzled smem (the K-major A operand of P·V), 32 columns at a time so the score
        // registers retire as we go; four partial sums keep the FADD chain short
        float l4[4] = {0.f, 0.f, 0.f, 0.f};
        uint8_t* pbuf = sP + st * C::kPBytes;
        // only the diagonal tile needs the causal mask: keep the per-element compare/select out of the common path
        auto exp_pack_store = [&](auto masked) {
#pragma unroll
          for (int c = 0; c < BKV / 32; ++c) {
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              float e0 = fast_exp2(fmaf(__uint_as_float(v[c * 32 + j]), p.scale_log2, -m_new));
              float e1 = fast_exp2(fmaf(__uint_as_float(v[c * 32 + j + 1]), p.scale_log2, -m_new));
              if (decltype(masked)::value) {
                if (c * 32 + j > r) e0 = 0.f;
                if (c * 32 + j + 1 > r) e1 = 0.f;
              }
              l4[(j >> 1) & 3] += e0 + e1;
              pk[j >> 1] = pack_bf16x2(__float_as_uint(e0), __float_as_uint(e1));
            }
            const uint32_t sbase = smem_u32(pbuf + (c >> 1) * (BQ * 128)) + r * 128;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint32_t chunk = (uint32_t)((c & 1) * 4 + i);
              st_shared_v4(sbase + ((chunk ^ row_sw) << 4), pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
            }
          }
        };
        if (diag) exp_pack_store(std::true_type{});
        else exp_pack_store(std::false_type{});
        l += (l4[0] + l4[1]) + (l4[2] + l4[3]);
        // rescale O if some row's max moved (needs P·V of the previous tile finished; ordered before ours by p_full below)
        if (any_grow && t > 0) {
          mbar_wait(&pv_done[(g - 1) & 1], ((g - 1) >> 1) & 1);
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < D / 32; ++c) {
            const uint32_t a = tmem_O + lane_addr + c * 32;
            uint32_t o[32];
            tmem_ld_32x32b_x32(a, o);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * alpha);
            tmem_st_32x32b_x32(a, o);
          }
          tmem_st_wait();
        }
        fence_proxy_async();  // P (generic-proxy smem writes) → visible to the tensor core's async proxy
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[st]);
      }
      // ---- epilogue of this work item: O / l → bf16 → swizzled staging (P buffer 0) → TMA store; lse2 = m + log2(l).
      // Meanwhile the MMA warp is already computing S of the next item's first tile and the loader is 1-2 tiles ahead.
      const int gl = g - 1;
      float m_fin;
      if ((gl & 1) == grp) {
        m_fin = m_ref;  // we processed the last tile: our reference IS the final max
      } else {
        consume();      // the other group's publish for the last tile
        m_fin = mrow[(gl & 1) * 128 + r];
      }
      l *= fast_exp2(m_ref - m_fin);  // partial sum → final reference (a group without tiles contributes exp2(-inf)·0 = 0)
      lrow[grp * 128 + r] = l;
      pair_sync();
      l += lrow[(grp ^ 1) * 128 + r];
      mbar_wait(&pv_done[gl & 1], (gl >> 1) & 1);
      if (gl >= 1) mbar_wait(&pv_done[(gl - 1) & 1], ((gl - 1) >> 1) & 1);  // staging aliases P buffer 0: both P·V readers retired
      tc_fence_after();
      const float inv_l = 1.f / l;
      uint8_t* stage = sP;
#pragma unroll 1
      for (int i = 0; i < kOC; ++i) {
        const int c = grp * kOC + i;  // 32-column chunk of O
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_O + lane_addr + c * 32, v);
        tmem_ld_wait();
        const uint32_t sbase = smem_u32(stage + (c >> 1) * (BQ * 128)) + r * 128;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t chunk = (uint32_t)((c & 1) * 4 + k);
          uint32_t wv[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            wv[e] = pack_bf16x2(__float_as_uint(__uint_as_float(v[8 * k + 2 * e]) * inv_l),
                                __float_as_uint(__uint_as_float(v[8 * k + 2 * e + 1]) * inv_l));
          st_shared_v4(sbase + ((chunk ^ row_sw) << 4), wv[0], wv[1], wv[2], wv[3]);
        }
      }
      // O has left TMEM → the first P·V of the next item may overwrite it
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_free);
      if (grp == 0) p.lse2[((int64_t)bh) * p.S + qb * BQ + r] = m_fin + log2f(l);
      fence_proxy_async();
      __syncwarp();
      // D=128: each warp filled one whole 64-column chunk and stores it; D=64: the pair shares chunk 0 → meet, then one store
      if (D == 64) pair_sync();
      if (lane == 0 && (D != 64 || grp == 0)) {
        const int c = D == 64 ? 0 : grp;
        tma_store_2d(&tmap_o, stage + c * (BQ * 128) + q * 32 * 128, col_q + c * 64, row0 + q * 32);
        bulk_commit();
        bulk_wait_read<0>();
      }
      __syncwarp();
      // the staging rows of BOTH warps of the pair alias rows that group 0 writes P into at its next tile, and lrow is reused by
      // the next item: nobody moves on until both stores have been read out of smem
      pair_sync();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int D>
int launch_fwd(const void* qkv, void* out, float* lse2, int B, int S, int H, int Hkv, float scale, int causal,
               cudaStream_t stream) {
  using C = FwdCfg<D>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(flash_fwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const uint64_t rows = (uint64_t)B * S, wqkv = (uint64_t)(H + 2 * Hkv) * D, wo = (uint64_t)H * D;
  CUtensorMap tq, to;
  int rc = pbhost::cached_tmap(&tq, qkv, rows, wqkv, wqkv, 64, 128, 2);
  if (rc) return rc;
  rc = pbhost::cached_tmap(&to, out, rows, wo, wo, 64, 32, 2);
  if (rc) return rc;
  FwdParams p{lse2, B, S, H, Hkv, scale * 1.4426950408889634f, causal};
  const int items = (S / BQ) * B * H;
  const int grid = items < pbhost::num_sms() ? items : pbhost::num_sms();  // persistent: one CTA per SM walks the work list
  flash_fwd_kernel<D><<<grid, kThreads, C::kSmem, stream>>>(tq, to, p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace

PB_EXPORT int pb_flash_attn_fwd2(const void* qkv, void* out, float* lse2, int B, int S, int H, int Hkv, float scale, int causal,
                                 cudaStream_t stream);  // attention_fwd2_sm100.cu

static int g_fwd_variant = -1;  // 1 = first-generation kernel, 2 = two query tiles per CTA + P in TMEM (D = 128, S % 256 == 0); -1 = env

// Forward kernel selection: 1 | 2, -1 = re-read PB_ATTN_FWD. Returns the previous value.
PB_EXPORT int pb_flash_attn_fwd_set_variant(int v) {
  const int old = g_fwd_variant;
  g_fwd_variant = v;
  return old;
}

PB_EXPORT int pb_flash_attn_fwd(const void* qkv, void* out, float* lse2, int B, int S, int H, int Hkv, int D, float scale,
                                int causal, cudaStream_t stream) {
  if (S % BQ != 0 || H % Hkv != 0) return -1;
  if (g_fwd_variant < 0) {
    const char* e = getenv("PB_ATTN_FWD");
    g_fwd_variant = (e && atoi(e) == 1) ? 1 : 2;  // measured: 0.141 → 0.115 ms (B16 S1024 H16 D128 causal), 0.434 → 0.332 ms (S 2048)
  }
  if (g_fwd_variant == 2 && D == 128 && S % (2 * BQ) == 0) return pb_flash_attn_fwd2(qkv, out, lse2, B, S, H, Hkv, scale, causal, stream);
  if (D == 128) return launch_fwd<128>(qkv, out, lse2, B, S, H, Hkv, scale, causal, stream);
  if (D == 64) return launch_fwd<64>(qkv, out, lse2, B, S, H, Hkv, scale, causal, stream);
  return -2;
}
