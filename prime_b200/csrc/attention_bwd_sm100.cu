// Flash-attention backward for sm_100a (tcgen05 / TMEM / TMA), fused-QKV layout, deterministic (no atomics).
//
//   delta[b,h,s]   = Σ_d dO·O                                                   (bwd_delta_kernel, memory bound)
//   dK, dV         : one CTA per 128-row KV block; loops over 64-row query tiles (bwd_dkdv_kernel)
//                    computes the TRANSPOSED scores Sᵀ = K Qᵀ, dPᵀ = V dOᵀ so that each softmax thread owns a KV row
//                    and writes Pᵀ / dSᵀ as K-major A operands:  dV += Pᵀ dO ,  dK += dSᵀ Q
//   dQ             : one CTA per 128-row query block; loops over 64-row KV tiles   (bwd_dq_kernel)
//                    S = Q Kᵀ, dP = dO Vᵀ, dS row-wise,  dQ += dS K
//
// A [rows x 64-column] 128B-swizzled box is simultaneously a K-major operand (rows = M/N) and an MN-major operand
// (rows = K), so each Q / dO / K tile is loaded ONCE by TMA and fed to both kinds of GEMM.
// With dS = P ∘ (dP − delta) · scale the results need no further scaling.  P is recomputed from the forward's
// log2-domain LSE: P = exp2(s·scale·log2e − lse2).
//
// TMEM budget (512 columns):  dK/dV kernel: Sᵀ 2×64 | dPᵀ 2×64 | dV D | dK D      dQ kernel: S 2×64 | dP 2×64 | dQ D
#include <type_traits>

#include "tc_common.cuh"
#include "tmap.h"

using namespace tc;

namespace {

constexpr int kThreads = 384;  // w0 TMA · w1 MMA · w2 TMEM alloc · w3 idle · w4..7 math group 0 (even tiles) · w8..11 math group 1 (odd tiles)

__device__ __forceinline__ void ld_shared_f4(uint32_t saddr, float (&v)[4]) {
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(saddr));
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(256) bwd_delta_kernel(const pb::bf16x8* __restrict__ dout, const pb::bf16x8* __restrict__ out,
                                                        float* __restrict__ delta, int64_t tokens, int S, int H, int vec_per_head,
                                                        float scale) {
  // one warp per TOKEN: it walks the token's H·D row in fully coalesced 512-byte steps (32 lanes x 16 B) and reduces inside groups
  // of `vec_per_head` lanes (D = 128 → 16 lanes, two heads per step; D = 64 → 8 lanes, four heads per step). Round 1 gave a warp to
  // every (token, head) pair, which left half of the lanes (D = 128) or three quarters (D = 64) without a vector to load.
  // Writes delta·scale (the only form the two kernels below use).
  const int64_t tok = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (tok >= tokens) return;
  const int vec_per_tok = H * vec_per_head;
  const int64_t base = tok * vec_per_tok;
  const int64_t bidx = tok / S;
  const int s = (int)(tok - bidx * S);
  for (int i0 = 0; i0 < vec_per_tok; i0 += 32) {
    const int i = i0 + lane;
    float acc = 0.f;
    if (i < vec_per_tok) {
      float a[8], b[8];
      pb::unpack8(pb::ldg_stream(dout + base + i), a);
      pb::unpack8(pb::ldg_stream(out + base + i), b);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc += a[j] * b[j];
    }
    for (int o = vec_per_head >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((lane & (vec_per_head - 1)) == 0 && i < vec_per_tok) {
      const int h = i / vec_per_head;
      delta[(bidx * H + h) * S + s] = acc * scale;
    }
  }
}

struct BwdParams {
  const float* lse2;
  const float* delta;  // rowsum(dO∘O) · scale
  int B, S, H, Hkv;
  float scale, scale_log2;
  int causal;
  unsigned long long* trace;  // optional pipeline trace (tools/attn_trace.py): CTA (0,0) stamps clock64() at role events
  // optional fused RoPE backward: fp32 tables [S, D/2]; dQ and dK leave the kernels already rotated back (inverse rotation,
  // interleaved-pair convention) so no separate pass over the 200 MB dQKV tensor is needed. nullptr = plain attention backward.
  const float* rope_cos;
  const float* rope_sin;
  int split_issue;  // S/dP and the accumulating MMAs (dQ | dV,dK) are issued by two different threads (warp 1 / warp 3)
  int tma3d;        // streamed operand tiles as ONE 3-D TMA box per operand instead of one 2-D box per 64-column chunk
  int l2_prefetch;  // issue cp.async.bulk.prefetch for the input tile a ring-depth ahead (PB_ATTN_BWD_L2PF, A/B switch)
};

// inverse RoPE on 32 consecutive head-dim columns [c32*32, c32*32+32) of one row at sequence position `pos`
template <int D>
__device__ __forceinline__ void rope_inverse_chunk(float (&x)[32], const float* __restrict__ cosb, const float* __restrict__ sinb, int pos,
                                                   int c32) {
  const float4* c4 = reinterpret_cast<const float4*>(cosb + (int64_t)pos * (D / 2) + c32 * 16);
  const float4* s4 = reinterpret_cast<const float4*>(sinb + (int64_t)pos * (D / 2) + c32 * 16);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float4 c = __ldg(c4 + k), sn = __ldg(s4 + k);
    const float cc[4] = {c.x, c.y, c.z, c.w}, ss[4] = {sn.x, sn.y, sn.z, sn.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = (k * 4 + e) * 2;
      const float a = x[j], b = x[j + 1];
      x[j] = a * cc[e] + b * ss[e];
      x[j + 1] = b * cc[e] - a * ss[e];
    }
  }
}

// trace record: [role 0..7][slot] = {event code, tile, clock}; roles 0..3 = dK/dV kernel (loader, MMA, math group 0, math group 1),
// 4..7 = the same four roles of the dQ kernel
__device__ __forceinline__ void trace_ev(const BwdParams& p, int role, int& n, int code, int tile) {
  if (p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && n < 256) {
    unsigned long long* e = p.trace + ((size_t)role * 256 + n) * 3;
    e[0] = (unsigned long long)code;
    e[1] = (unsigned long long)tile;
    e[2] = (unsigned long long)clock64();
    ++n;
  }
}

// write one 32-value fp32 chunk of a row as bf16 into a [rows x 64] swizzled block (half: which 32 columns)
__device__ __forceinline__ void store_row_chunk_bf16(uint32_t block_base, int r, int half, const float (&x)[32]) {
  const uint32_t sbase = block_base + r * 128;
  const uint32_t sw = (uint32_t)(r & 7);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t chunk = (uint32_t)(half * 4 + i);
    st_shared_v4(sbase + ((chunk ^ sw) << 4), pack_bf16x2(__float_as_uint(x[8 * i]), __float_as_uint(x[8 * i + 1])),
                 pack_bf16x2(__float_as_uint(x[8 * i + 2]), __float_as_uint(x[8 * i + 3])),
                 pack_bf16x2(__float_as_uint(x[8 * i + 4]), __float_as_uint(x[8 * i + 5])),
                 pack_bf16x2(__float_as_uint(x[8 * i + 6]), __float_as_uint(x[8 * i + 7])));
  }
}

// same, for 32 values already packed as 16 bf16x2 words
__device__ __forceinline__ void store_row_chunk_packed(uint32_t block_base, int r, int half, const uint32_t (&x)[16]) {
  const uint32_t sbase = block_base + r * 128;
  const uint32_t sw = (uint32_t)(r & 7);
#pragma unroll
  for (int i = 0; i < 4; ++i) st_shared_v4(sbase + (((uint32_t)(half * 4 + i) ^ sw) << 4), x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
}

// ----------------------------------------------------------------------------------------------------- dK / dV
template <int D, int PST>
struct DkvCfg {
  static constexpr int kChunks = D / 64;
  static constexpr uint32_t kKVBytes = 128 * D * 2;           // resident K or V (128 rows)
  static constexpr uint32_t kQBytes = 64 * D * 2;             // one Q or dO tile (64 rows)
  static constexpr uint32_t kPBytes = 128 * 128;              // Pᵀ or dSᵀ : [128 kv rows x 64 q]
  static constexpr uint32_t kOffV = kKVBytes;
  // Smem budget (224 KB of operands): K,V 64 KB + Q/dO ring kQStages x 32 KB + Pᵀ/dSᵀ kPStages x 32 KB. Two variants, A/B-tested
  // (PB_ATTN_BWD_PSTAGES): 4 input stages + ONE Pᵀ/dSᵀ buffer (a math group waits for the previous tile's dV/dK MMAs right before
  // storing) — best while the math phase was the bottleneck; or 3 input stages + TWO buffers (a group only waits for its own
  // tile of two iterations ago), which takes the dV/dK MMAs of the other group's tile off the store's critical path.
  //   PST = 0: Pᵀ/dSᵀ never touch shared memory — the math warps write them (bf16, tcgen05.st) over the S/dP columns they have
  //   just read, and dV/dK take their A operand from TMEM (tcgen05.mma with [tmem] A). The 64 KB go to a 5-deep input ring,
  //   the st.shared + proxy fence disappear, but S/dP of tile it+2 can only be issued behind dV/dK of tile it (same columns).
  static constexpr int kPStages = PST;
  static constexpr int kQStages = PST == 0 ? 5 : PST == 2 ? 3 : 4;
  static constexpr uint32_t kOffQ = 2 * kKVBytes;
  static constexpr uint32_t kOffdO = kOffQ + kQStages * kQBytes;
  static constexpr uint32_t kOffP = kOffdO + kQStages * kQBytes;
  static constexpr uint32_t kOffdS = kOffP + kPStages * kPBytes;
  static constexpr uint32_t kOffBar = kOffdS + kPStages * kPBytes;
  // per-query lse / delta·scale of the current and the next tile of each math group: [group][slot][lse 64 | delta 64] fp32
  static constexpr uint32_t kOffAux = kOffBar + 256;
  static constexpr uint32_t kSmem = kOffAux + 2 * 2 * 512;  // the dynamic smem base is declared 1024-aligned (checked at entry)
  static constexpr uint32_t tS = 0, tdP = 128, tdV = 256, tdK = 256 + D;
};

template <int D, int PST>
__global__ void __launch_bounds__(kThreads, 1)
    bwd_dkdv_kernel(const __grid_constant__ CUtensorMap tmap_qkv128, const __grid_constant__ CUtensorMap tmap_qkv64,
                    const __grid_constant__ CUtensorMap tmap_do64, const __grid_constant__ CUtensorMap tmap_dqkv,
                    const __grid_constant__ CUtensorMap tmap_qkv64_3d, const __grid_constant__ CUtensorMap tmap_do64_3d,
                    const BwdParams p) {
  using C = DkvCfg<D, PST>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  if (smem_u32(smem_raw) & 1023) __trap();  // SWIZZLE_128B tiles need it and there is no slack left to realign by hand
  uint8_t* smem = smem_raw;
  uint8_t* sK = smem;
  uint8_t* sV = smem + C::kOffV;
  uint8_t* sQ = smem + C::kOffQ;
  uint8_t* sdO = smem + C::kOffdO;
  uint8_t* sP = smem + C::kOffP;
  uint8_t* sdS = smem + C::kOffdS;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kOffBar);
  uint64_t* kv_full = bars;        // 1
  uint64_t* q_full = bars + 1;     // up to 5
  uint64_t* q_empty = bars + 6;    // up to 5
  uint64_t* s_full = bars + 11;    // 2
  uint64_t* s_empty = bars + 13;   // 2 (4 warps)
  uint64_t* p_full = bars + 15;    // 2 (4 warps)
  uint64_t* acc_done = bars + 17;  // 2
  uint64_t* all_done = bars + 19;  // 1: every dV/dK MMA of this block has retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int jb = blockIdx.x;  // KV block (128 rows)
  const int b = blockIdx.y / p.Hkv, hk = blockIdx.y % p.Hkv;
  const int group = p.H / p.Hkv;
  const int nq64 = p.S / 64;
  const int it0 = p.causal ? (jb * 128) / 64 : 0;  // first query tile that can see this KV block
  const int tiles_per_head = nq64 - it0;
  const int n_it = tiles_per_head * group;
  const int kvrow0 = b * p.S + jb * 128;
  const int col_k = (p.H + hk) * D, col_v = (p.H + p.Hkv + hk) * D;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_qkv128);
    prefetch_tmap(&tmap_qkv64);
    prefetch_tmap(&tmap_do64);
    prefetch_tmap(&tmap_dqkv);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(kv_full, 1);
    mbar_init(all_done, 1);
    for (int i = 0; i < C::kQStages; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 4);  // stage i is always served by math group i (4 warps)
      mbar_init(&p_full[i], 4);
      mbar_init(&acc_done[i], 1);
    }
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
      mbar_expect_tx(kv_full, 2 * C::kKVBytes);
#pragma unroll
      for (int c = 0; c < C::kChunks; ++c) {
        tma_load_2d(&tmap_qkv128, kv_full, sK + c * (128 * 128), col_k + c * 64, kvrow0);
        tma_load_2d(&tmap_qkv128, kv_full, sV + c * (128 * 128), col_v + c * 64, kvrow0);
      }
      int tr_n = 0;
      for (int it = 0; it < n_it; ++it) {
        const int st = it % C::kQStages;
        const uint32_t ph = (it / C::kQStages) & 1;
        const int h = hk * group + it / tiles_per_head;
        const int qrow = b * p.S + (it0 + it % tiles_per_head) * 64;
        {  // the tile that will be loaded kQStages iterations from now: start moving it HBM → L2 (measured: a cold 32 KB
           // Q/dO stage takes ≈3900 clk to land, and only one load can be in flight with every slot held until dV/dK retire)
          const int nx = it + C::kQStages;
          if (p.l2_prefetch && nx < n_it) {
            const int hn = hk * group + nx / tiles_per_head;
            const int qn = b * p.S + (it0 + nx % tiles_per_head) * 64;
#pragma unroll
            for (int c = 0; c < C::kChunks; ++c) {
              tma_prefetch_l2_2d(&tmap_qkv64, hn * D + c * 64, qn);
              tma_prefetch_l2_2d(&tmap_do64, hn * D + c * 64, qn);
            }
          }
        }
        mbar_wait(&q_empty[st], ph ^ 1);
        trace_ev(p, 0, tr_n, 1, it);  // slot free → issue the Q/dO loads of tile `it`
        mbar_expect_tx(&q_full[st], 2 * C::kQBytes);
        if (p.tma3d) {
          tma_load_3d(&tmap_qkv64_3d, &q_full[st], sQ + st * C::kQBytes, 0, qrow, h * D / 64);
          tma_load_3d(&tmap_do64_3d, &q_full[st], sdO + st * C::kQBytes, 0, qrow, h * D / 64);
        } else {
#pragma unroll
          for (int c = 0; c < C::kChunks; ++c) {
            tma_load_2d(&tmap_qkv64, &q_full[st], sQ + st * C::kQBytes + c * (64 * 128), h * D + c * 64, qrow);
            tma_load_2d(&tmap_do64, &q_full[st], sdO + st * C::kQBytes + c * (64 * 128), h * D + c * 64, qrow);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = idesc_bf16(128, 64, 0, 0);  // Sᵀ[128 kv x 64 q] = K Qᵀ   (both K-major)
      constexpr uint32_t idesc_a = idesc_bf16(128, D, 0, 1);   // dV/dK[128 kv x D] += Pᵀ·dO  (A K-major, B MN-major)
      mbar_wait(kv_full, 0);
      int tr_n = 0;
      auto issue_sd = [&](int it) {
        const int st = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        const int qs = it % C::kQStages;
        mbar_wait(&q_full[qs], (it / C::kQStages) & 1);
        trace_ev(p, 1, tr_n, 1, it);  // Q/dO landed
        if (PST != 0) mbar_wait(&s_empty[st], ph ^ 1);  // PST = 0: program order (issued behind dV/dK of tile it-2) is the guarantee
        trace_ev(p, 1, tr_n, 2, it);  // S/dP stage free → issue S, dP
        tc_fence_after();
        const uint32_t k0 = smem_u32(sK), v0 = smem_u32(sV);
        const uint32_t q0 = smem_u32(sQ + qs * C::kQBytes), d0 = smem_u32(sdO + qs * C::kQBytes);
#pragma unroll
        for (int c = 0; c < C::kChunks; ++c)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + C::tS + st * 64, make_smem_desc(k0 + c * (128 * 128) + k * 32, 16, 1024),
                      make_smem_desc(q0 + c * (64 * 128) + k * 32, 16, 1024), idesc_s, (c | k) != 0 ? 1u : 0u);
#pragma unroll
        for (int c = 0; c < C::kChunks; ++c)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + C::tdP + st * 64, make_smem_desc(v0 + c * (128 * 128) + k * 32, 16, 1024),
                      make_smem_desc(d0 + c * (64 * 128) + k * 32, 16, 1024), idesc_s, (c | k) != 0 ? 1u : 0u);
        umma_commit(&s_full[st]);
      };
      // S/dP run up to TWO tiles ahead: TMEM stage it&1 is free as soon as its math group has pulled tile `it` into registers (long
      // before it finishes the math), so tile it+2 can already be waiting in TMEM when that group comes back. The second tile is
      // issued opportunistically from the P-wait polling loop: a blocking wait for its Q/dO load in front of dV/dK of tile `it`
      // stalls the whole pipeline (measured with a 3-slot ring: slower than one-ahead; the trace shows the MMA thread parked there).
      auto sd_ready = [&](int it) {  // can S/dP of tile `it` be issued without blocking?
        return mbar_test_wait(&q_full[it % C::kQStages], (it / C::kQStages) & 1) && mbar_test_wait(&s_empty[it & 1], ((it >> 1) & 1) ^ 1);
      };
      if (PST == 0) {
        // Pᵀ/dSᵀ live in TMEM on top of S/dP of their own stage: S/dP(it+2) must follow dV/dK(it) in the (in-order) tensor pipe
        issue_sd(0);
        if (n_it > 1) issue_sd(1);
        for (int it = 0; it < n_it; ++it) {
          const int st = it & 1;
          mbar_wait(&p_full[st], (it >> 1) & 1);
          trace_ev(p, 1, tr_n, 3, it);
          tc_fence_after();
          const int qs = it % C::kQStages;
          const uint32_t q0 = smem_u32(sQ + qs * C::kQBytes), d0 = smem_u32(sdO + qs * C::kQBytes);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)  // K = 64 query rows = 8 TMEM columns of packed bf16 pairs per step
            umma_bf16_ts(tmem_base + C::tdV, tmem_base + C::tS + st * 64 + kk * 8, make_smem_desc(d0 + kk * 2048, 64 * 128, 1024), idesc_a,
                         (it | kk) != 0 ? 1u : 0u);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16_ts(tmem_base + C::tdK, tmem_base + C::tdP + st * 64 + kk * 8, make_smem_desc(q0 + kk * 2048, 64 * 128, 1024), idesc_a,
                         (it | kk) != 0 ? 1u : 0u);
          umma_commit(&q_empty[qs]);
          if (it + 2 < n_it) issue_sd(it + 2);
        }
        umma_commit(all_done);
      } else if (p.split_issue) {
        for (int it = 0; it < n_it; ++it) issue_sd(it);  // dV/dK are issued by warp 3 (see bwd_dq_kernel); s_empty keeps this ≤ 2 tiles ahead
      } else {
      issue_sd(0);
      int sd_next = 1;  // next tile whose S/dP has not been issued
      for (int it = 0; it < n_it; ++it) {
        if (sd_next <= it + 1 && sd_next < n_it) issue_sd(sd_next++);  // one tile ahead is mandatory
        const int st = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        {  // two tiles ahead whenever it costs nothing; never block in front of dV/dK of tile `it`
          SpinGuard guard;
          while (!mbar_test_wait(&p_full[st], ph)) {  // test_wait: try_wait may park the thread for a time slice, and the opportunistic issue below must not wait for that
            guard.tick();
            if (sd_next == it + 2 && sd_next < n_it && sd_ready(sd_next)) issue_sd(sd_next++);
          }
        }
        trace_ev(p, 1, tr_n, 3, it);  // Pᵀ/dSᵀ ready → issue dV, dK
        tc_fence_after();
        const int qs = it % C::kQStages;
        const uint32_t pbuf = PST == 2 ? (uint32_t)st * C::kPBytes : 0u;  // see DkvCfg
        const uint32_t pa = smem_u32(sP) + pbuf, da = smem_u32(sdS) + pbuf;
        const uint32_t q0 = smem_u32(sQ + qs * C::kQBytes), d0 = smem_u32(sdO + qs * C::kQBytes);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)  // K = 64 query rows
          umma_bf16(tmem_base + C::tdV, make_smem_desc(pa + kk * 32, 16, 1024), make_smem_desc(d0 + kk * 2048, 64 * 128, 1024),
                    idesc_a, (it | kk) != 0 ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_bf16(tmem_base + C::tdK, make_smem_desc(da + kk * 32, 16, 1024), make_smem_desc(q0 + kk * 2048, 64 * 128, 1024),
                    idesc_a, (it | kk) != 0 ? 1u : 0u);
        umma_commit(&q_empty[qs]);
        umma_commit(&acc_done[st]);
      }
      umma_commit(all_done);
      }
    }
  } else if (warp == 3) {
    if (lane == 0 && p.split_issue && PST != 0) {
      // ------------------------------------------------------------------ second MMA issuer: dV += Pᵀ·dO, dK += dSᵀ·Q
      constexpr uint32_t idesc_a = idesc_bf16(128, D, 0, 1);
      mbar_wait(kv_full, 0);
      int tr_n = 128;
      for (int it = 0; it < n_it; ++it) {
        const int st = it & 1;
        const int qs = it % C::kQStages;
        mbar_wait(&p_full[st], (it >> 1) & 1);
        mbar_wait(&q_full[qs], (it / C::kQStages) & 1);  // landed long ago: acquire for this thread
        trace_ev(p, 1, tr_n, 3, it);
        tc_fence_after();
        const uint32_t pbuf = PST == 2 ? (uint32_t)st * C::kPBytes : 0u;
        const uint32_t pa = smem_u32(sP) + pbuf, da = smem_u32(sdS) + pbuf;
        const uint32_t q0 = smem_u32(sQ + qs * C::kQBytes), d0 = smem_u32(sdO + qs * C::kQBytes);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_bf16(tmem_base + C::tdV, make_smem_desc(pa + kk * 32, 16, 1024), make_smem_desc(d0 + kk * 2048, 64 * 128, 1024), idesc_a,
                    (it | kk) != 0 ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_bf16(tmem_base + C::tdK, make_smem_desc(da + kk * 32, 16, 1024), make_smem_desc(q0 + kk * 2048, 64 * 128, 1024), idesc_a,
                    (it | kk) != 0 ? 1u : 0u);
        umma_commit(&q_empty[qs]);
        umma_commit(&acc_done[st]);
      }
      umma_commit(all_done);
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax-backward math + epilogue
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;  // math group: owns the tiles of this parity and, in the epilogue, dV (0) or dK (1)
    const int r = q * 32 + lane;       // KV row in the block == TMEM lane
    const int kv_idx = jb * 128 + r;   // position in the sequence
    const int kv_warp_min = jb * 128 + q * 32;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    // Two math groups leapfrog over the tiles (group g owns the tiles — and therefore the S/dP TMEM stage and the Pᵀ/dSᵀ smem
    // stage — of parity g). TMEM can only be read at 64 B/clk per SM, so one group's 64 KB of S/dP loads overlap the other
    // group's exp2/FMA/pack/st.shared phase instead of every warp queueing on the same resource at the same time.
    int tr_n = 0;
    // lse and delta·scale are per QUERY, i.e. per register column of this kv-row-per-thread layout: 128 warp-uniform values per
    // tile. As 32 LDG.128 in the math phase they were its critical path (ncu source view: the first FFMA/FMUL after each load
    // carried the stall samples; with the smem carve-out at 224 KB the loads mostly miss L1). Now warp 0 of the group copies
    // the NEXT tile's 512 bytes global → smem with one cp.async per lane while the current tile is being processed, and the
    // math phase reads them back as broadcast LDS.128 (29 clk).
    const uint32_t aux_base = smem_u32(smem + C::kOffAux) + (uint32_t)half * 1024u;
    auto stage_aux = [&](int itn, uint32_t slot) {
      const int hn = hk * group + itn / tiles_per_head;
      const int64_t off = ((int64_t)(b * p.H + hn)) * p.S + (it0 + itn % tiles_per_head) * 64 + (lane & 15) * 4;
      const float* src = (lane & 16) ? p.delta + off : p.lse2 + off;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(aux_base + slot * 512u + (uint32_t)lane * 16u), "l"(src) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (q == 0 && half < n_it) stage_aux(half, 0u);
    for (int it = half; it < n_it; it += 2) {
      const int st = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      const int h = hk * group + it / tiles_per_head;
      const int qpos0 = (it0 + it % tiles_per_head) * 64;
      const uint32_t aux = aux_base + (uint32_t)(((it - half) >> 1) & 1) * 512u;
      if (q == 0 && lane == 0) trace_ev(p, 2 + half, tr_n, 1, it);  // start waiting for S/dP
      mbar_wait(&s_full[st], ph);
      if (q == 0 && lane == 0) trace_ev(p, 2 + half, tr_n, 2, it);  // S/dP ready
      tc_fence_after();
      const bool need_mask = p.causal && (qpos0 < jb * 128 + 128);
      // all four TMEM loads in flight, one wait; the per-query lse/delta rows are warp-uniform 128-bit loads
      uint32_t sv[64], dv[64];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        tmem_ld_32x32b_x32(tmem_base + C::tS + st * 64 + lane_addr + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[c * 32]));
        tmem_ld_32x32b_x32(tmem_base + C::tdP + st * 64 + lane_addr + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&dv[c * 32]));
      }
      // group rendezvous (named barrier 1 + group, 128 threads): every warp has finished reading the other slot (tile it-2),
      // and warp 0's copy of THIS tile's values — issued a whole tile ago — is complete and visible
      if (q == 0) asm volatile("cp.async.wait_group 0;" ::: "memory");
      asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");
      if (q == 0 && it + 2 < n_it) stage_aux(it + 2, (uint32_t)((((it - half) >> 1) + 1) & 1));
      tmem_ld_wait();
      // S/dP of this stage now live in registers → the MMA warp may overwrite the stage with tile it+2
      tc_fence_before();
      __syncwarp();
      if (PST != 0 && lane == 0) mbar_arrive(&s_empty[st]);
      if (q == 0 && lane == 0) trace_ev(p, 2 + half, tr_n, 3, it);  // TMEM loads done
      // only the (at most two) diagonal tiles need the causal mask: keep the per-element compare/select out of the common path
      uint32_t pk[32], dk[32];  // Pᵀ and dSᵀ rows of this thread, packed bf16x2
      auto tile_math = [&](auto masked) {
#pragma unroll
        for (int j4 = 0; j4 < 16; ++j4) {
          // diagonal tiles: per 4-column group the warp's 32 kv rows are either all masked (nothing to compute), all visible
          // (common path) or cut by the diagonal (per-element select) — warp-uniform branches
          bool cut = false;
          if (decltype(masked)::value) {
            const int qa = qpos0 + j4 * 4;
            if (kv_warp_min > qa + 3) {
              pk[2 * j4] = pk[2 * j4 + 1] = dk[2 * j4] = dk[2 * j4 + 1] = 0u;
              continue;
            }
            cut = kv_warp_min + 31 > qa;
          }
          float ls[4], dl[4];
          ld_shared_f4(aux + j4 * 16, ls);
          ld_shared_f4(aux + 256 + j4 * 16, dl);
          auto group4 = [&](auto cut_here) {
            float pv[4], dsv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int qc = j4 * 4 + e;
              pv[e] = fast_exp2(fmaf(__uint_as_float(sv[qc]), p.scale_log2, -ls[e]));
              if (decltype(cut_here)::value && kv_idx > qpos0 + qc) pv[e] = 0.f;
              dsv[e] = pv[e] * fmaf(__uint_as_float(dv[qc]), p.scale, -dl[e]);
            }
            pk[2 * j4] = pack_bf16x2(__float_as_uint(pv[0]), __float_as_uint(pv[1]));
            pk[2 * j4 + 1] = pack_bf16x2(__float_as_uint(pv[2]), __float_as_uint(pv[3]));
            dk[2 * j4] = pack_bf16x2(__float_as_uint(dsv[0]), __float_as_uint(dsv[1]));
            dk[2 * j4 + 1] = pack_bf16x2(__float_as_uint(dsv[2]), __float_as_uint(dsv[3]));
          };
          if (decltype(masked)::value && cut) group4(std::true_type{});
          else group4(std::false_type{});
        }
      };
      if (need_mask) tile_math(std::true_type{});
      else tile_math(std::false_type{});
      if (q == 0 && lane == 0) trace_ev(p, 2 + half, tr_n, 6, it);  // math done (registers hold the packed tile)
      if (PST == 0) {
        // overwrite the S / dP columns of this stage (already in registers) with the packed bf16 rows: the A operands of dV / dK
        if (q == 0 && lane == 0) trace_ev(p, 2 + half, tr_n, 4, it);
        tmem_st_32x32b_x32(tmem_base + C::tS + st * 64 + lane_addr, pk);
        tmem_st_32x32b_x32(tmem_base + C::tdP + st * 64 + lane_addr, dk);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[st]);
        if (q == 0 && lane == 0) trace_ev(p, 2 + half, tr_n, 5, it);
        continue;
      }
      // one buffer: it was last read by dV/dK of the previous tile (the other group's); two buffers: by this group's own
      // tile of two iterations ago
      if (PST == 2) {
        if (it >= 2) mbar_wait(&acc_done[st], ph ^ 1);
      } else if (it >= 1) {
        mbar_wait(&acc_done[(it - 1) & 1], ((it - 1) >> 1) & 1);
      }
      const uint32_t pbuf = PST == 2 ? (uint32_t)st * C::kPBytes : 0u;
      if (q == 0 && lane == 0) trace_ev(p, 2 + half, tr_n, 4, it);  // P/dS buffer free
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        store_row_chunk_packed(smem_u32(sP) + pbuf, r, c, *reinterpret_cast<uint32_t(*)[16]>(&pk[c * 16]));
        store_row_chunk_packed(smem_u32(sdS) + pbuf, r, c, *reinterpret_cast<uint32_t(*)[16]>(&dk[c * 16]));
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[st]);
      if (q == 0 && lane == 0) trace_ev(p, 2 + half, tr_n, 5, it);  // math + stores done
    }
    // epilogue: dV, dK → bf16 → staging (the Pᵀ/dSᵀ region, 64 KB) → TMA stores into dqkv
    mbar_wait(all_done, 0);  // committed after the last tile's MMAs (parity waits on the per-stage barriers can alias, see bwd_dq)
    tc_fence_after();
    uint8_t* stage = sQ;  // the Q/dO rings are idle now: [which(dV,dK)][chunk c] blocks of [128 rows x 128 B]
    {
      const int which = half;  // warps 4..7 drain dV, warps 8..11 drain dK
#pragma unroll 1
      for (int c32 = 0; c32 < D / 32; ++c32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + (which ? C::tdK : C::tdV) + lane_addr + c32 * 32, v);
        tmem_ld_wait();
        float x[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
        if (which == 1 && p.rope_cos != nullptr) rope_inverse_chunk<D>(x, p.rope_cos, p.rope_sin, kv_idx, c32);  // dK only
        store_row_chunk_bf16(smem_u32(stage + (which * C::kChunks + (c32 >> 1)) * (128 * 128)), r, c32 & 1, x);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
#pragma unroll
        for (int c = 0; c < C::kChunks; ++c)
          tma_store_2d(&tmap_dqkv, stage + (which * C::kChunks + c) * (128 * 128) + q * 32 * 128, (which ? col_k : col_v) + c * 64,
                       kvrow0 + q * 32);
        bulk_commit();
        bulk_wait_read<0>();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ----------------------------------------------------------------------------------------------------- dQ
template <int D>
struct DqCfg {
  static constexpr int kChunks = D / 64;
  static constexpr uint32_t kQBytes = 128 * D * 2;   // resident Q or dO (128 rows)
  static constexpr uint32_t kKVBytes = 64 * D * 2;   // one K or V tile (64 rows)
  static constexpr uint32_t kdSBytes = 128 * 128;    // dS : [128 q rows x 64 kv]
  static constexpr uint32_t kOffdO = kQBytes;
  static constexpr int kKVStages = 4;  // K/V ring: a slot is held until dQ += dS·K of its tile retires; S/dP run two tiles ahead
  static constexpr uint32_t kOffK = 2 * kQBytes;
  static constexpr uint32_t kOffV = kOffK + kKVStages * kKVBytes;
  static constexpr uint32_t kOffdS = kOffV + kKVStages * kKVBytes;
  static constexpr uint32_t kOffBar = kOffdS + 2 * kdSBytes;
  static constexpr uint32_t kSmem = kOffBar + 256 + 1024;
  static constexpr uint32_t tS = 0, tdP = 128, tdQ = 256;
  // QT = 1 (D = 128): the resident Q and dO tiles are copied ONCE into tensor memory (2 x 64 columns of packed bf16 — exactly the
  // 128 columns S, dP and dQ leave free) and S = Q·Kᵀ, dP = dO·Vᵀ take their A operand from there. A 128x64x16 MMA with both
  // operands in shared memory fetches 6 KB in its 32 clk (192 B/clk against 128 B/clk of smem bandwidth) and the tensor core
  // re-read all 64 KB of Q and dO for every 64 keys; with A in TMEM the same instruction fetches 2 KB.
  static constexpr uint32_t tQ = 384, tdO = 448;
};

template <int D, int QT>
__global__ void __launch_bounds__(kThreads, 1)
    bwd_dq_kernel(const __grid_constant__ CUtensorMap tmap_qkv128, const __grid_constant__ CUtensorMap tmap_qkv64,
                  const __grid_constant__ CUtensorMap tmap_do128, const __grid_constant__ CUtensorMap tmap_dqkv,
                  const __grid_constant__ CUtensorMap tmap_qkv64_3d, const BwdParams p) {
  using C = DqCfg<D>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sdO = smem + C::kOffdO;
  uint8_t* sK = smem + C::kOffK;
  uint8_t* sV = smem + C::kOffV;
  uint8_t* sdS = smem + C::kOffdS;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kOffBar);
  uint64_t* q_full = bars;         // 1
  uint64_t* kv_full = bars + 1;    // 4
  uint64_t* kv_empty = bars + 5;   // 4
  uint64_t* s_full = bars + 9;     // 2
  uint64_t* s_empty = bars + 11;   // 2 (4 warps)
  uint64_t* p_full = bars + 13;    // 2 (4 warps)
  uint64_t* acc_done = bars + 15;  // 2
  uint64_t* all_done = bars + 17;  // 1: every dQ MMA of this block has retired
  uint64_t* a_in_tmem = bars + 18; // 8 warps: Q and dO have been copied into tensor memory (QT = 1)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nqb = p.S / 128;
  const int qb = nqb - 1 - (int)blockIdx.x;  // heavy blocks first
  const int bh = blockIdx.y;
  const int b = bh / p.H, h = bh % p.H;
  const int hk = h / (p.H / p.Hkv);
  const int n_kv = p.causal ? (qb * 128 + 128) / 64 : p.S / 64;
  const int row0 = b * p.S + qb * 128;
  const int col_q = h * D, col_k = (p.H + hk) * D, col_v = (p.H + p.Hkv + hk) * D;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_qkv128);
    prefetch_tmap(&tmap_qkv64);
    prefetch_tmap(&tmap_do128);
    prefetch_tmap(&tmap_dqkv);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    mbar_init(all_done, 1);
    mbar_init(a_in_tmem, 8);
    for (int i = 0; i < C::kKVStages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 4);  // stage i is always served by math group i (4 warps)
      mbar_init(&p_full[i], 4);
      mbar_init(&acc_done[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, 2 * C::kQBytes);
#pragma unroll
      for (int c = 0; c < C::kChunks; ++c) {
        tma_load_2d(&tmap_qkv128, q_full, sQ + c * (128 * 128), col_q + c * 64, row0);
        tma_load_2d(&tmap_do128, q_full, sdO + c * (128 * 128), col_q + c * 64, row0);
      }
      int tr_n = 0;
      for (int t = 0; t < n_kv; ++t) {
        const int st = t % C::kKVStages;
        const uint32_t ph = (t / C::kKVStages) & 1;
        const int krow = b * p.S + t * 64;
        if (p.l2_prefetch && t + C::kKVStages < n_kv) {
          const int kn = b * p.S + (t + C::kKVStages) * 64;
#pragma unroll
          for (int c = 0; c < C::kChunks; ++c) {
            tma_prefetch_l2_2d(&tmap_qkv64, col_k + c * 64, kn);
            tma_prefetch_l2_2d(&tmap_qkv64, col_v + c * 64, kn);
          }
        }
        mbar_wait(&kv_empty[st], ph ^ 1);
        trace_ev(p, 4, tr_n, 1, t);
        mbar_expect_tx(&kv_full[st], 2 * C::kKVBytes);
        if (p.tma3d) {
          tma_load_3d(&tmap_qkv64_3d, &kv_full[st], sK + st * C::kKVBytes, 0, krow, col_k / 64);
          tma_load_3d(&tmap_qkv64_3d, &kv_full[st], sV + st * C::kKVBytes, 0, krow, col_v / 64);
        } else {
#pragma unroll
          for (int c = 0; c < C::kChunks; ++c) {
            tma_load_2d(&tmap_qkv64, &kv_full[st], sK + st * C::kKVBytes + c * (64 * 128), col_k + c * 64, krow);
            tma_load_2d(&tmap_qkv64, &kv_full[st], sV + st * C::kKVBytes + c * (64 * 128), col_v + c * 64, krow);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = idesc_bf16(128, 64, 0, 0);  // S[128 q x 64 kv] = Q Kᵀ
      constexpr uint32_t idesc_a = idesc_bf16(128, D, 0, 1);   // dQ[128 q x D] += dS · K   (K tile as MN-major B)
      mbar_wait(q_full, 0);
      if (QT) {
        mbar_wait(a_in_tmem, 0);
        tc_fence_after();
      }
      int tr_n = 0;
      auto issue_sd = [&](int t) {
        const int st = t & 1;
        const uint32_t ph = (t >> 1) & 1;
        const int ks = t % C::kKVStages;
        mbar_wait(&kv_full[ks], (t / C::kKVStages) & 1);
        trace_ev(p, 5, tr_n, 1, t);
        mbar_wait(&s_empty[st], ph ^ 1);
        trace_ev(p, 5, tr_n, 2, t);
        tc_fence_after();
        const uint32_t q0 = smem_u32(sQ), d0 = smem_u32(sdO);
        const uint32_t k0 = smem_u32(sK + ks * C::kKVBytes), v0 = smem_u32(sV + ks * C::kKVBytes);
        if (QT) {
#pragma unroll
          for (int kk = 0; kk < D / 16; ++kk)  // 16 head-dim elements = 8 TMEM columns of packed bf16 pairs per step
            umma_bf16_ts(tmem_base + C::tS + st * 64, tmem_base + C::tQ + kk * 8,
                         make_smem_desc(k0 + (kk >> 2) * (64 * 128) + (kk & 3) * 32, 16, 1024), idesc_s, kk != 0 ? 1u : 0u);
#pragma unroll
          for (int kk = 0; kk < D / 16; ++kk)
            umma_bf16_ts(tmem_base + C::tdP + st * 64, tmem_base + C::tdO + kk * 8,
                         make_smem_desc(v0 + (kk >> 2) * (64 * 128) + (kk & 3) * 32, 16, 1024), idesc_s, kk != 0 ? 1u : 0u);
        } else {
#pragma unroll
        for (int c = 0; c < C::kChunks; ++c)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + C::tS + st * 64, make_smem_desc(q0 + c * (128 * 128) + k * 32, 16, 1024),
                      make_smem_desc(k0 + c * (64 * 128) + k * 32, 16, 1024), idesc_s, (c | k) != 0 ? 1u : 0u);
#pragma unroll
        for (int c = 0; c < C::kChunks; ++c)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + C::tdP + st * 64, make_smem_desc(d0 + c * (128 * 128) + k * 32, 16, 1024),
                      make_smem_desc(v0 + c * (64 * 128) + k * 32, 16, 1024), idesc_s, (c | k) != 0 ? 1u : 0u);
        }
        umma_commit(&s_full[st]);
      };
      auto sd_ready = [&](int t) {
        return mbar_test_wait(&kv_full[t % C::kKVStages], (t / C::kKVStages) & 1) && mbar_test_wait(&s_empty[t & 1], ((t >> 1) & 1) ^ 1);
      };
      if (p.split_issue) {
        // this thread only feeds S/dP (back-pressured by s_empty: at most two tiles ahead); dQ is issued by warp 3. One thread doing
        // both spends ≈1150 clk inside a batch of 16 S/dP MMAs (the operand-bound tensor pipe back-pressures the issue) and
        // cannot see the other group's p_full meanwhile — dQ(t) then went out ≈500 clk late on every tile (trace)
        for (int t = 0; t < n_kv; ++t) issue_sd(t);
      } else {
      issue_sd(0);  // S/dP: one tile ahead mandatory, two ahead opportunistic (see the dK/dV kernel)
      int sd_next = 1;
      for (int t = 0; t < n_kv; ++t) {
        if (sd_next <= t + 1 && sd_next < n_kv) issue_sd(sd_next++);
        const int st = t & 1;
        const uint32_t ph = (t >> 1) & 1;
        {
          SpinGuard guard;
          while (!mbar_test_wait(&p_full[st], ph)) {  // test_wait: try_wait may park the thread for a time slice, and the opportunistic issue below must not wait for that
            guard.tick();
            if (sd_next == t + 2 && sd_next < n_kv && sd_ready(sd_next)) issue_sd(sd_next++);
          }
        }
        trace_ev(p, 5, tr_n, 3, t);
        tc_fence_after();
        const int ks = t % C::kKVStages;
        const uint32_t da = smem_u32(sdS + st * C::kdSBytes), k0 = smem_u32(sK + ks * C::kKVBytes);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)  // K = 64 kv rows
          umma_bf16(tmem_base + C::tdQ, make_smem_desc(da + kk * 32, 16, 1024), make_smem_desc(k0 + kk * 2048, 64 * 128, 1024),
                    idesc_a, (t | kk) != 0 ? 1u : 0u);
        umma_commit(&kv_empty[ks]);
        umma_commit(&acc_done[st]);
      }
      umma_commit(all_done);
      }
    }
  } else if (warp == 3) {
    if (lane == 0 && p.split_issue) {
      // ------------------------------------------------------------------ second MMA issuer: dQ += dS·K as soon as dS(t) is in smem
      constexpr uint32_t idesc_a = idesc_bf16(128, D, 0, 1);
      int tr_n = 128;  // upper half of the MMA role's trace slots
      for (int t = 0; t < n_kv; ++t) {
        const int st = t & 1;
        const int ks = t % C::kKVStages;
        mbar_wait(&p_full[st], (t >> 1) & 1);
        mbar_wait(&kv_full[ks], (t / C::kKVStages) & 1);  // landed long ago (S of this tile used it): acquire for this thread
        trace_ev(p, 5, tr_n, 3, t);
        tc_fence_after();
        const uint32_t da = smem_u32(sdS + st * C::kdSBytes), k0 = smem_u32(sK + ks * C::kKVBytes);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_bf16(tmem_base + C::tdQ, make_smem_desc(da + kk * 32, 16, 1024), make_smem_desc(k0 + kk * 2048, 64 * 128, 1024), idesc_a,
                    (t | kk) != 0 ? 1u : 0u);
        umma_commit(&kv_empty[ks]);
        umma_commit(&acc_done[st]);
      }
      umma_commit(all_done);
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;  // math group: owns the kv tiles of this parity and half of the dQ columns in the epilogue
    const int r = q * 32 + lane;
    const int q_idx = qb * 128 + r;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const float lse = p.lse2[((int64_t)bh) * p.S + q_idx];
    const float del_s = p.delta[((int64_t)bh) * p.S + q_idx];
    if (QT) {
      // group 0 copies Q, group 1 copies dO: thread = row = TMEM lane; un-swizzle the row's 16-byte pieces out of the
      // [2 chunks][128 rows x 128 B] tile and store the 64 packed words in head-dim order
      mbar_wait(q_full, 0);
      const uint32_t src = smem_u32(half ? sdO : sQ) + (uint32_t)r * 128u;
      const uint32_t dst = tmem_base + (half ? C::tdO : C::tQ) + lane_addr;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t w[32];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(w[4 * j]), "=r"(w[4 * j + 1]), "=r"(w[4 * j + 2]), "=r"(w[4 * j + 3])
                       : "r"(src + (uint32_t)c * (128u * 128u) + (((uint32_t)j ^ (uint32_t)(r & 7)) << 4)));
        tmem_st_32x32b_x32(dst + c * 32, w);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_in_tmem);
    }
    int tr_n = 0;
    const bool tr = q == 0 && lane == 0;
    for (int t = half; t < n_kv; t += 2) {  // the two math groups leapfrog over the kv tiles (see the dK/dV kernel)
      const int st = t & 1;
      const uint32_t ph = (t >> 1) & 1;
      if (tr) trace_ev(p, 6 + half, tr_n, 1, t);
      mbar_wait(&s_full[st], ph);
      if (tr) trace_ev(p, 6 + half, tr_n, 2, t);
      tc_fence_after();
      const int kv0 = t * 64;
      const bool need_mask = p.causal && (kv0 + 64 > qb * 128);
      uint32_t sv[64], dv[64];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        tmem_ld_32x32b_x32(tmem_base + C::tS + st * 64 + lane_addr + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[c * 32]));
        tmem_ld_32x32b_x32(tmem_base + C::tdP + st * 64 + lane_addr + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&dv[c * 32]));
      }
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[st]);
      if (tr) trace_ev(p, 6 + half, tr_n, 3, t);
      if (t >= 2) mbar_wait(&acc_done[st], ph ^ 1);
      if (tr) trace_ev(p, 6 + half, tr_n, 4, t);
      auto tile_math = [&](auto masked) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float ds[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float pv = fast_exp2(fmaf(__uint_as_float(sv[c * 32 + j]), p.scale_log2, -lse));
            if (decltype(masked)::value && (kv0 + c * 32 + j) > q_idx) pv = 0.f;
            ds[j] = pv * fmaf(__uint_as_float(dv[c * 32 + j]), p.scale, -del_s);
          }
          store_row_chunk_bf16(smem_u32(sdS + st * C::kdSBytes), r, c, ds);
        }
      };
      if (need_mask) tile_math(std::true_type{});
      else tile_math(std::false_type{});
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[st]);
      if (tr) trace_ev(p, 6 + half, tr_n, 5, t);
    }
    // NOT acc_done[last tile's stage]: a group only follows its OWN stage's barrier inside the loop, so it can reach this point
    // while the other stage is still two phases behind — a parity wait would then match the phase BEFORE the previous one and
    // fall through with the last tile's dQ += dS·K still outstanding (found by compute-sanitizer's timing perturbation)
    mbar_wait(all_done, 0);
    tc_fence_after();
    uint8_t* stage = sK;  // the K/V rings (4 x kKVBytes = 64 KB at D = 128) are idle now
    constexpr int kC32PerWarp = D / 64;  // each warp of a quadrant drains half of the D accumulator columns
#pragma unroll 1
    for (int i = 0; i < kC32PerWarp; ++i) {
      const int c32 = half * kC32PerWarp + i;
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem_base + C::tdQ + lane_addr + c32 * 32, v);
      tmem_ld_wait();
      float x[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
      if (p.rope_cos != nullptr) rope_inverse_chunk<D>(x, p.rope_cos, p.rope_sin, q_idx, c32);
      store_row_chunk_bf16(smem_u32(stage + (c32 >> 1) * (128 * 128)), r, c32 & 1, x);
    }
    fence_proxy_async();
    __syncwarp();
    // a 64-column chunk is stored once both 32-column halves are in smem: D=128 → each warp owns one whole chunk;
    // D=64 → the two warps share chunk 0, so pair up through a named barrier and let the first warp issue the store
    if (D == 64) asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
    if (lane == 0 && (D != 64 || half == 0)) {
      if (D == 64) {
        tma_store_2d(&tmap_dqkv, stage + q * 32 * 128, col_q, row0 + q * 32);
      } else {
#pragma unroll
        for (int i = 0; i < kC32PerWarp / 2; ++i) {
          const int c = half * (kC32PerWarp / 2) + i;
          tma_store_2d(&tmap_dqkv, stage + c * (128 * 128) + q * 32 * 128, col_q + c * 64, row0 + q * 32);
        }
      }
      bulk_commit();
      bulk_wait_read<0>();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

unsigned long long* g_bwd_trace = nullptr;
int g_bwd_pstages = -1;  // dK/dV kernel variant: 1 (default) | 2 | 0 = Pᵀ/dSᵀ in TMEM (see DkvCfg); -1 = read PB_ATTN_BWD_PSTAGES

template <int D>
int launch_bwd(const void* qkv, const void* out, const void* dout, const float* lse2, float* delta, void* dqkv, int B, int S,
               int H, int Hkv, float scale, int causal, const float* rope_cos, const float* rope_sin, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(bwd_dkdv_kernel<D, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DkvCfg<D, 1>::kSmem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(bwd_dkdv_kernel<D, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DkvCfg<D, 2>::kSmem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(bwd_dkdv_kernel<D, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DkvCfg<D, 0>::kSmem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(bwd_dq_kernel<D, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DqCfg<D>::kSmem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(bwd_dq_kernel<D, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DqCfg<D>::kSmem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const uint64_t rows = (uint64_t)B * S, wqkv = (uint64_t)(H + 2 * Hkv) * D, wo = (uint64_t)H * D;
  {
    const int64_t warps = (int64_t)rows;  // one warp per token (vec_per_head = D / 8 is 8 or 16: a power of two <= 32)
    const int64_t blocks = (warps * 32 + 255) / 256;
    bwd_delta_kernel<<<(unsigned)blocks, 256, 0, stream>>>((const pb::bf16x8*)dout, (const pb::bf16x8*)out, delta, (int64_t)rows, S,
                                                          H, D / 8, scale);
  }
  CUtensorMap tq128, tq64, tdo64, tdo128, tdq;
  int rc;
  if ((rc = pbhost::cached_tmap(&tq128, qkv, rows, wqkv, wqkv, 64, 128, 2))) return rc;
  if ((rc = pbhost::cached_tmap(&tq64, qkv, rows, wqkv, wqkv, 64, 64, 2))) return rc;
  if ((rc = pbhost::cached_tmap(&tdo64, dout, rows, wo, wo, 64, 64, 2))) return rc;
  if ((rc = pbhost::cached_tmap(&tdo128, dout, rows, wo, wo, 64, 128, 2))) return rc;
  if ((rc = pbhost::cached_tmap(&tdq, dqkv, rows, wqkv, wqkv, 64, 32, 2))) return rc;
  CUtensorMap tq64_3d, tdo64_3d;
  if ((rc = pbhost::cached_tmap3(&tq64_3d, qkv, rows, wqkv, wqkv, 64, D / 64))) return rc;
  if ((rc = pbhost::cached_tmap3(&tdo64_3d, dout, rows, wo, wo, 64, D / 64))) return rc;
  static int split_issue = -1;
  if (split_issue < 0) {
    const char* ev = getenv("PB_ATTN_BWD_SPLIT_ISSUE");
    split_issue = ev ? (atoi(ev) != 0) : 1;
  }
  static int tma3d = -1;
  if (tma3d < 0) {
    const char* ev = getenv("PB_ATTN_BWD_TMA3D");
    tma3d = ev ? (atoi(ev) != 0) : 1;
  }
  static int l2pf = -1;
  if (l2pf < 0) {
    const char* ev = getenv("PB_ATTN_BWD_L2PF");
    l2pf = ev ? (atoi(ev) != 0) : 1;
  }
  BwdParams p{lse2, delta, B, S, H, Hkv, scale, scale * 1.4426950408889634f, causal, g_bwd_trace, rope_cos, rope_sin, split_issue, tma3d, l2pf};
  if (g_bwd_pstages < 0) {
    const char* ev = getenv("PB_ATTN_BWD_PSTAGES");
    // with two MMA issuers the 4-deep input ring + one P buffer measures 2 % faster than 3 stages + two buffers (0.471 vs 0.480 ms)
    g_bwd_pstages = ev ? atoi(ev) : 1;
    if (g_bwd_pstages < 0 || g_bwd_pstages > 2) g_bwd_pstages = 1;
  }
  const int pstages = g_bwd_pstages;
  if (pstages == 0) bwd_dkdv_kernel<D, 0><<<dim3(S / 128, B * Hkv), kThreads, DkvCfg<D, 0>::kSmem, stream>>>(tq128, tq64, tdo64, tdq, tq64_3d, tdo64_3d, p);
  else if (pstages == 2) bwd_dkdv_kernel<D, 2><<<dim3(S / 128, B * Hkv), kThreads, DkvCfg<D, 2>::kSmem, stream>>>(tq128, tq64, tdo64, tdq, tq64_3d, tdo64_3d, p);
  else bwd_dkdv_kernel<D, 1><<<dim3(S / 128, B * Hkv), kThreads, DkvCfg<D, 1>::kSmem, stream>>>(tq128, tq64, tdo64, tdq, tq64_3d, tdo64_3d, p);
  static int dq_ts = -1;
  if (dq_ts < 0) {
    const char* ev = getenv("PB_ATTN_BWD_DQ_TS");
    dq_ts = ev ? (atoi(ev) != 0) : 0;  // measured neutral (0.528 vs 0.524 ms): the kernel is not bound by operand fetch
  }
  if (dq_ts && D == 128) bwd_dq_kernel<D, 1><<<dim3(S / 128, B * H), kThreads, DqCfg<D>::kSmem, stream>>>(tq128, tq64, tdo128, tdq, tq64_3d, p);
  else bwd_dq_kernel<D, 0><<<dim3(S / 128, B * H), kThreads, DqCfg<D>::kSmem, stream>>>(tq128, tq64, tdo128, tdq, tq64_3d, p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace

// Device buffer of 4 roles x 256 events x 3 u64 that CTA (0,0) of the dK/dV kernel fills (nullptr = tracing off).
PB_EXPORT void pb_flash_attn_bwd_set_trace(unsigned long long* buf) { g_bwd_trace = buf; }

// dK/dV kernel variant (see DkvCfg): 1 = one smem Pᵀ/dSᵀ buffer + 4-deep input ring (default), 2 = two buffers + 3-deep ring,
// 0 = Pᵀ/dSᵀ in TMEM as the A operand of tcgen05.mma + 5-deep ring, -1 = re-read PB_ATTN_BWD_PSTAGES. Returns the old value.
PB_EXPORT int pb_flash_attn_bwd_set_variant(int v) {
  const int old = g_bwd_pstages;
  g_bwd_pstages = v;
  return old;
}

// delta: [B, H, S] fp32 scratch.  dqkv: [B, S, (H+2Hkv)·D] bf16, fully overwritten.
PB_EXPORT int pb_flash_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse2, float* delta, void* dqkv,
                                int B, int S, int H, int Hkv, int D, float scale, int causal, cudaStream_t stream) {
  if (S % 128 != 0 || H % Hkv != 0) return -1;
  if (D == 128) return launch_bwd<128>(qkv, out, dout, lse2, delta, dqkv, B, S, H, Hkv, scale, causal, nullptr, nullptr, stream);
  if (D == 64) return launch_bwd<64>(qkv, out, dout, lse2, delta, dqkv, B, S, H, Hkv, scale, causal, nullptr, nullptr, stream);
  return -2;
}

// Same, with the RoPE backward fused into the dQ / dK epilogues (cos/sin: fp32 [S, D/2] tables of the forward rotation).
PB_EXPORT int pb_flash_attn_bwd_rope(const void* qkv, const void* out, const void* dout, const float* lse2, float* delta, void* dqkv,
                                     int B, int S, int H, int Hkv, int D, float scale, int causal, const float* rope_cos,
                                     const float* rope_sin, cudaStream_t stream) {
  if (S % 128 != 0 || H % Hkv != 0) return -1;
  if (D == 128) return launch_bwd<128>(qkv, out, dout, lse2, delta, dqkv, B, S, H, Hkv, scale, causal, rope_cos, rope_sin, stream);
  if (D == 64) return launch_bwd<64>(qkv, out, dout, lse2, delta, dqkv, B, S, H, Hkv, scale, causal, rope_cos, rope_sin, stream);
  return -2;
}
