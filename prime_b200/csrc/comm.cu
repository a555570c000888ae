// NVLink/NVSwitch peer-memory layer and the fused collective+compute kernels.
//
//  * symmetric heap: one cudaMalloc per rank, exported with cudaIpc and mapped by every peer, so a
//    kernel can ld/st any peer's buffer at (peer_base + offset) through NVSwitch;
//  * device-side flags (st.release.sys / ld.acquire.sys) for stream-ordered signalling — no host
//    sync, no NCCL launch on the hot path; every spin has a wall-clock bound and reports a timeout
//    through a device error word instead of hanging the GPU;
//  * FSDP inner step:   grad reduce-scatter (peer loads, fixed rank order)  ⊕ scale ⊕ sum-of-squares,
//                       then clip ⊕ partitioned AdamW ⊕ bf16 cast ⊕ parameter all-gather (peer stores);
//  * DiLoCo outer step: pseudo-gradient ⊕ int8 block quantise, then peer int8 all-gather ⊕ dequant-sum
//                       ⊕ Nesterov ⊕ parameter write-back/all-gather — one kernel pair.
#include "common.cuh"

using namespace pb;

constexpr int kMaxPeers = 8;
constexpr unsigned long long kSpinTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;  // 20 s

struct PeerPtrs {
  void* p[kMaxPeers];
  int n;
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Spin until *flag >= expect (monotone epochs). Returns false on timeout.
__device__ __forceinline__ bool spin_wait_ge(const uint32_t* flag, uint32_t expect, uint32_t* err) {
  const unsigned long long t0 = globaltimer_ns();
  while ((int32_t)(ld_acquire_sys_u32(flag) - expect) < 0) {
    __nanosleep(64);
    if (globaltimer_ns() - t0 > kSpinTimeoutNs) {
      if (err) atomicExch(err, 1u);
      return false;
    }
  }
  return true;
}

// ------------------------------------------------------------------------------------------------
// IPC plumbing (host side)
// ------------------------------------------------------------------------------------------------
PB_EXPORT int pb_ipc_alloc(void** ptr, size_t bytes) {
  cudaError_t e = cudaMalloc(ptr, bytes);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaMemset(*ptr, 0, bytes);
}
PB_EXPORT int pb_ipc_free(void* ptr) { return (int)cudaFree(ptr); }
PB_EXPORT int pb_ipc_handle_size() { return (int)sizeof(cudaIpcMemHandle_t); }
PB_EXPORT int pb_ipc_get_handle(void* ptr, void* handle_out) {
  return (int)cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle_out), ptr);
}
PB_EXPORT int pb_ipc_open_handle(const void* handle, void** ptr_out) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  return (int)cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess);
}
PB_EXPORT int pb_ipc_close_handle(void* ptr) { return (int)cudaIpcCloseMemHandle(ptr); }
PB_EXPORT int pb_enable_peer_access(int peer_device) {
  int can = 0, cur = 0;
  cudaGetDevice(&cur);
  if (peer_device == cur) return 0;
  cudaDeviceCanAccessPeer(&can, cur, peer_device);
  if (!can) return -2;
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();
    return 0;
  }
  return (int)e;
}

// ------------------------------------------------------------------------------------------------
// signal / barrier
// ------------------------------------------------------------------------------------------------
// flags live at the same offset in every rank's heap: slot layout [channel][rank].
__global__ void signal_kernel(PeerPtrs flags, int slot, uint32_t value) {
  if ((int)threadIdx.x < flags.n) {
    __threadfence_system();
    st_release_sys_u32(reinterpret_cast<uint32_t*>(flags.p[threadIdx.x]) + slot, value);
  }
}
PB_EXPORT int pb_signal(const PeerPtrs* flags, int slot, uint32_t value, cudaStream_t stream) {
  signal_kernel<<<1, 32, 0, stream>>>(*flags, slot, value);
  PB_CHECK_LAUNCH();
  return 0;
}

// Full barrier across `flags.n` ranks: publish `epoch` into slot[base + my_idx] on every peer, then wait
// until every slot[base + i] in MY heap reached `epoch`.
__global__ void barrier_kernel(PeerPtrs flags, int base_slot, int my_idx, uint32_t epoch, uint32_t* err) {
  const int t = threadIdx.x;
  if (t < flags.n) {
    __threadfence_system();
    st_release_sys_u32(reinterpret_cast<uint32_t*>(flags.p[t]) + base_slot + my_idx, epoch);
    spin_wait_ge(reinterpret_cast<uint32_t*>(flags.p[my_idx]) + base_slot + t, epoch, err);
  }
}
PB_EXPORT int pb_barrier(const PeerPtrs* flags, int base_slot, int my_idx, uint32_t epoch, uint32_t* err,
                         cudaStream_t stream) {
  barrier_kernel<<<1, 32, 0, stream>>>(*flags, base_slot, my_idx, epoch, err);
  PB_CHECK_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// FSDP phase 1: reduce-scatter of one gradient bucket by peer loads.
//   out[i] = scale * sum_p grads[p][off + i]   (p in fixed rank order → bitwise identical across runs)
//   sumsq_partial[blockIdx] += sum_i out[i]^2  (accumulated over buckets; folded by phase 2)
// If wait_flags != nullptr the kernel first waits for flag[slot_base + p] >= expect for every peer p.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) grad_reduce_kernel(PeerPtrs grads, int64_t off, int64_t n, float scale,
                                                          float* __restrict__ out, float* __restrict__ sumsq_partial,
                                                          const uint32_t* wait_flags, int slot_base, uint32_t expect,
                                                          uint32_t* err) {
  __shared__ float red[32];
  if (wait_flags != nullptr) {
    if ((int)threadIdx.x < grads.n) spin_wait_ge(wait_flags + slot_base + threadIdx.x, expect, err);
    __syncthreads();
  }
  const int64_t nvec = n >> 2;
  float ss = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int p = 0; p < kMaxPeers; ++p) {
      if (p < grads.n) {
        const float4 v = ld_relaxed_sys_f4(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(grads.p[p]) + off) + i);
        acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
      }
    }
    acc.x *= scale, acc.y *= scale, acc.z *= scale, acc.w *= scale;
    reinterpret_cast<float4*>(out)[i] = acc;
    ss += acc.x * acc.x + acc.y * acc.y + acc.z * acc.z + acc.w * acc.w;
  }
  ss = block_sum(ss, red);
  if (threadIdx.x == 0) sumsq_partial[blockIdx.x] += ss;
}

PB_EXPORT int pb_grad_reduce_grid() { return 148 * 2; }

PB_EXPORT int pb_grad_reduce(const PeerPtrs* grads, int64_t off, int64_t n, float scale, float* out, float* sumsq_partial,
                             const uint32_t* wait_flags, int slot_base, uint32_t expect, uint32_t* err, int max_ctas,
                             cudaStream_t stream) {
  if (n % 4 != 0 || off % 4 != 0) return -1;
  int grid = pb_grad_reduce_grid();
  if (max_ctas > 0 && max_ctas < grid) grid = max_ctas;
  grad_reduce_kernel<<<grid, 512, 0, stream>>>(*grads, off, n, scale, out, sumsq_partial, wait_flags, slot_base, expect,
                                               err);
  PB_CHECK_LAUNCH();
  return 0;
}

// Fold this rank's sum-of-squares partials, publish to every peer's norm slot, signal.
__global__ void norm_publish_kernel(float* __restrict__ sumsq_partial, int nparts, PeerPtrs norm_slots, PeerPtrs flags,
                                    int my_idx, int flag_slot, uint32_t epoch) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) {
    s += sumsq_partial[i];
    sumsq_partial[i] = 0.f;  // ready for the next step
  }
  s = block_sum(s, red);
  if ((int)threadIdx.x < norm_slots.n) {
    reinterpret_cast<float*>(norm_slots.p[threadIdx.x])[my_idx] = s;
    __threadfence_system();
    st_release_sys_u32(reinterpret_cast<uint32_t*>(flags.p[threadIdx.x]) + flag_slot + my_idx, epoch);
  }
}
PB_EXPORT int pb_norm_publish(float* sumsq_partial, int nparts, const PeerPtrs* norm_slots, const PeerPtrs* flags,
                              int my_idx, int flag_slot, uint32_t epoch, cudaStream_t stream) {
  norm_publish_kernel<<<1, 512, 0, stream>>>(sumsq_partial, nparts, *norm_slots, *flags, my_idx, flag_slot, epoch);
  PB_CHECK_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// FSDP phase 2: clip ⊕ AdamW on this rank's fp32 master shard ⊕ bf16 cast ⊕ all-gather by peer stores.
// ------------------------------------------------------------------------------------------------
struct AdamArgs {
  float lr, beta1, beta2, eps, weight_decay, bc1, bc2, max_norm;
};

__global__ void __launch_bounds__(512) adamw_push_kernel(float* __restrict__ p32, const float* __restrict__ g32,
                                                         float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                         AdamArgs a, const float* norm_slots, int n_norm,
                                                         const uint32_t* wait_flags, int flag_slot, uint32_t epoch,
                                                         PeerPtrs dst, int64_t dst_off, float* gnorm_out, uint32_t* err) {
  __shared__ float s_clip;
  if (threadIdx.x == 0) {
    float clip = 1.f;
    if (norm_slots != nullptr) {
      float tot = 0.f;
      for (int i = 0; i < n_norm; ++i) {
        if (wait_flags) spin_wait_ge(wait_flags + flag_slot + i, epoch, err);
        tot += reinterpret_cast<const volatile float*>(norm_slots)[i];
      }
      const float gn = sqrtf(tot);
      if (gnorm_out && blockIdx.x == 0) *gnorm_out = gn;
      if (a.max_norm > 0.f) clip = fminf(1.f, a.max_norm / (gn + 1e-6f));
    }
    s_clip = clip;
  }
  __syncthreads();
  const float clip = s_clip;
  const float step_size = a.lr / a.bc1;
  const float inv_sqrt_bc2 = rsqrtf(a.bc2);
  const float decay = 1.f - a.lr * a.weight_decay;
  const int64_t nvec = n >> 3;  // 8 elements per thread-iteration → one 16-byte bf16 store per peer
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    float out[8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t j = i * 2 + h;
      float4 p = reinterpret_cast<float4*>(p32)[j];
      const float4 g = reinterpret_cast<const float4*>(g32)[j];
      float4 mm = reinterpret_cast<float4*>(m)[j];
      float4 vv = reinterpret_cast<float4*>(v)[j];
      float* pp = &p.x;
      const float* gp = &g.x;
      float* mp = &mm.x;
      float* vp = &vv.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float gk = gp[k] * clip;
        mp[k] = a.beta1 * mp[k] + (1.f - a.beta1) * gk;
        vp[k] = a.beta2 * vp[k] + (1.f - a.beta2) * gk * gk;
        const float denom = sqrtf(vp[k]) * inv_sqrt_bc2 + a.eps;
        pp[k] = pp[k] * decay - step_size * mp[k] / denom;
        out[h * 4 + k] = pp[k];
      }
      reinterpret_cast<float4*>(p32)[j] = p;
      reinterpret_cast<float4*>(m)[j] = mm;
      reinterpret_cast<float4*>(v)[j] = vv;
    }
    const bf16x8 packed = pack8(out);
#pragma unroll
    for (int q = 0; q < kMaxPeers; ++q)
      if (q < dst.n) stg_v4(reinterpret_cast<bf16x8*>(reinterpret_cast<__nv_bfloat16*>(dst.p[q]) + dst_off) + i, packed);
  }
}

PB_EXPORT int pb_adamw_push(float* p32, const float* g32, float* m, float* v, int64_t n, const AdamArgs* a,
                            const float* norm_slots, int n_norm, const uint32_t* wait_flags, int flag_slot,
                            uint32_t epoch, const PeerPtrs* dst, int64_t dst_off, float* gnorm_out, uint32_t* err,
                            cudaStream_t stream) {
  if (n % 8 != 0 || dst_off % 8 != 0) return -1;
  adamw_push_kernel<<<148 * 2, 512, 0, stream>>>(p32, g32, m, v, n, *a, norm_slots, n_norm, wait_flags, flag_slot, epoch,
                                                 *dst, dst_off, gnorm_out, err);
  PB_CHECK_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// DiLoCo outer step, kernel 1: pseudo-gradient ⊕ symmetric int8 block quantisation.
//   delta = theta0 - theta ;  q = round(delta / (absmax/127)) ; one fp32 scale per `block` elements.
// One CTA (256 thr) per 1024-element block, 4 elements per thread.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pseudograd_quant_kernel(const float* __restrict__ theta0,
                                                               const float* __restrict__ theta, int8_t* __restrict__ q,
                                                               float* __restrict__ scales, int64_t n) {
  __shared__ float red[32];
  for (int64_t blk = blockIdx.x; blk * 1024 < n; blk += gridDim.x) {
    const int64_t base = blk * 1024 + threadIdx.x * 4;
    float d[4] = {0.f, 0.f, 0.f, 0.f};
    if (base + 3 < n) {
      const float4 a = *reinterpret_cast<const float4*>(theta0 + base);
      const float4 b = *reinterpret_cast<const float4*>(theta + base);
      d[0] = a.x - b.x, d[1] = a.y - b.y, d[2] = a.z - b.z, d[3] = a.w - b.w;
    } else {
      for (int k = 0; k < 4; ++k)
        if (base + k < n) d[k] = theta0[base + k] - theta[base + k];
    }
    float am = fmaxf(fmaxf(fabsf(d[0]), fabsf(d[1])), fmaxf(fabsf(d[2]), fabsf(d[3])));
    am = block_max(am, red);
    const float scale = am / 127.f;
    const float inv = scale > 0.f ? 1.f / scale : 0.f;
    if (threadIdx.x == 0) scales[blk] = scale;
    char4 o;
    o.x = (signed char)fmaxf(-127.f, fminf(127.f, rintf(d[0] * inv)));
    o.y = (signed char)fmaxf(-127.f, fminf(127.f, rintf(d[1] * inv)));
    o.z = (signed char)fmaxf(-127.f, fminf(127.f, rintf(d[2] * inv)));
    o.w = (signed char)fmaxf(-127.f, fminf(127.f, rintf(d[3] * inv)));
    if (base + 3 < n) {
      *reinterpret_cast<char4*>(q + base) = o;
    } else {
      const signed char ov[4] = {o.x, o.y, o.z, o.w};
      for (int k = 0; k < 4; ++k)
        if (base + k < n) q[base + k] = ov[k];
    }
  }
}

PB_EXPORT int pb_pseudograd_quant(const float* theta0, const float* theta, int8_t* q, float* scales, int64_t n,
                                  cudaStream_t stream) {
  int64_t blocks = (n + 1023) / 1024;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  pseudograd_quant_kernel<<<(unsigned)blocks, 256, 0, stream>>>(theta0, theta, q, scales, n);
  PB_CHECK_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// DiLoCo outer step, kernel 2: int8 all-gather by peer loads ⊕ dequantise-sum (fixed worker order)
// ⊕ Nesterov SGD on theta0 ⊕ reset of the inner fp32 master ⊕ bf16 write-back to the FSDP group.
// `qs.p[w]` / `ss.p[w]` are worker w's int8 payload / scales for THIS shard (peer-mapped).
// ------------------------------------------------------------------------------------------------
struct OuterArgs {
  float lr, momentum, inv_workers;
  int nesterov;
};

__global__ void __launch_bounds__(256) outer_nesterov_kernel(PeerPtrs qs, PeerPtrs ss, float* __restrict__ theta0,
                                                             float* __restrict__ mom, float* __restrict__ theta, int64_t n,
                                                             OuterArgs a, PeerPtrs dst, int64_t dst_off) {
  for (int64_t blk = blockIdx.x; blk * 1024 < n; blk += gridDim.x) {
    const int64_t base = blk * 1024 + threadIdx.x * 4;
    if (base >= n) continue;
    float g[4] = {0.f, 0.f, 0.f, 0.f};
    const bool full = base + 3 < n;
#pragma unroll
    for (int w = 0; w < kMaxPeers; ++w) {
      if (w < qs.n) {
        const float sc = *reinterpret_cast<const volatile float*>(reinterpret_cast<const float*>(ss.p[w]) + blk);
        const int8_t* qp = reinterpret_cast<const int8_t*>(qs.p[w]) + base;
        if (full) {
          int packed;
          asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(packed) : "l"(qp));
          g[0] += (float)(int8_t)(packed & 0xff) * sc;
          g[1] += (float)(int8_t)((packed >> 8) & 0xff) * sc;
          g[2] += (float)(int8_t)((packed >> 16) & 0xff) * sc;
          g[3] += (float)(int8_t)((packed >> 24) & 0xff) * sc;
        } else {
          for (int k = 0; k < 4; ++k)
            if (base + k < n) g[k] += (float)reinterpret_cast<const volatile int8_t*>(qp)[k] * sc;
        }
      }
    }
    float outv[4];
    for (int k = 0; k < 4; ++k) {
      if (base + k < n) {
        const float gk = g[k] * a.inv_workers;
        const float mk = a.momentum * mom[base + k] + gk;
        mom[base + k] = mk;
        const float upd = a.nesterov ? gk + a.momentum * mk : mk;
        const float t = theta0[base + k] - a.lr * upd;
        theta0[base + k] = t;
        theta[base + k] = t;
        outv[k] = t;
      }
    }
    if (full) {
      const __nv_bfloat162 lo = __floats2bfloat162_rn(outv[0], outv[1]), hi = __floats2bfloat162_rn(outv[2], outv[3]);
      uint2 pk;
      pk.x = *reinterpret_cast<const uint32_t*>(&lo);
      pk.y = *reinterpret_cast<const uint32_t*>(&hi);
#pragma unroll
      for (int q = 0; q < kMaxPeers; ++q)
        if (q < dst.n) *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(dst.p[q]) + dst_off + base) = pk;
    } else {
      for (int k = 0; k < 4; ++k)
        if (base + k < n)
          for (int q = 0; q < dst.n; ++q)
            (reinterpret_cast<__nv_bfloat16*>(dst.p[q]) + dst_off)[base + k] = __float2bfloat16(outv[k]);
    }
  }
}

PB_EXPORT int pb_outer_nesterov(const PeerPtrs* qs, const PeerPtrs* ss, float* theta0, float* mom, float* theta, int64_t n,
                                const OuterArgs* a, const PeerPtrs* dst, int64_t dst_off, cudaStream_t stream) {
  if (dst_off % 4 != 0) return -1;
  int64_t blocks = (n + 1023) / 1024;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  outer_nesterov_kernel<<<(unsigned)blocks, 256, 0, stream>>>(*qs, *ss, theta0, mom, theta, n, *a, *dst, dst_off);
  PB_CHECK_LAUNCH();
  return 0;
}

// fp32 (uncompressed) outer path: all-gather of pseudo-gradients by peer loads, same update.
__global__ void __launch_bounds__(256) outer_nesterov_f32_kernel(PeerPtrs thetas /*peer inner masters*/,
                                                                 float* __restrict__ theta0, float* __restrict__ mom,
                                                                 float* __restrict__ theta_tmp, int64_t n, OuterArgs a) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float t0 = theta0[i];
    float g = 0.f;
    for (int w = 0; w < thetas.n; ++w) g += t0 - *reinterpret_cast<const volatile float*>(reinterpret_cast<const float*>(thetas.p[w]) + i);
    g *= a.inv_workers;
    const float mk = a.momentum * mom[i] + g;
    mom[i] = mk;
    const float upd = a.nesterov ? g + a.momentum * mk : mk;
    const float t = t0 - a.lr * upd;
    theta0[i] = t;
    theta_tmp[i] = t;  // caller copies into the inner master after a barrier (peers still read theta)
  }
}
PB_EXPORT int pb_outer_nesterov_f32(const PeerPtrs* thetas, float* theta0, float* mom, float* theta_tmp, int64_t n,
                                    const OuterArgs* a, cudaStream_t stream) {
  outer_nesterov_f32_kernel<<<148 * 4, 256, 0, stream>>>(*thetas, theta0, mom, theta_tmp, n, *a);
  PB_CHECK_LAUNCH();
  return 0;
}

// fp32 master → bf16 parameter buffers of every FSDP peer (used after checkpoint load / outer f32 path).
__global__ void __launch_bounds__(512) cast_push_kernel(const float* __restrict__ src, int64_t n, PeerPtrs dst,
                                                        int64_t dst_off) {
  const int64_t nvec = n >> 3;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(src)[2 * i], b = reinterpret_cast<const float4*>(src)[2 * i + 1];
    const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    const bf16x8 pk = pack8(f);
#pragma unroll
    for (int q = 0; q < kMaxPeers; ++q)
      if (q < dst.n) stg_v4(reinterpret_cast<bf16x8*>(reinterpret_cast<__nv_bfloat16*>(dst.p[q]) + dst_off) + i, pk);
  }
}
PB_EXPORT int pb_cast_push(const float* src, int64_t n, const PeerPtrs* dst, int64_t dst_off, cudaStream_t stream) {
  if (n % 8 != 0 || dst_off % 8 != 0) return -1;
  cast_push_kernel<<<148 * 2, 512, 0, stream>>>(src, n, *dst, dst_off);
  PB_CHECK_LAUNCH();
  return 0;
}
