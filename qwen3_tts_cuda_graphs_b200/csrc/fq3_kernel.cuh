// fq3_kernel.cuh — the persistent weight-streaming decode kernel (sm_100a), v2.
//
// Replaces the CUDA-graph replays of talker_graph.py:97-107,198-214 and predictor_graph.py:115-167 and
// the eager per-frame glue of generate.py:149-199 (reference paths under /root/reference/faster_qwen3_tts).
//
// Structure of one CTA (one per SM, all co-resident):
//   warp 16      producer: walks the phase program and streams this CTA's weight rows of every GEMV phase
//                through a shared-memory ring with cp.async.bulk (TMA bulk copy).  Weight addresses never
//                depend on activations or sampled ids, so it runs ahead of the consumers across phases.
//   warps 0..15  consumers: per phase, poll-read the input activations (LL words), then each warp consumes
//                the ring tiles independently (row -> warp round-robin, no CTA barrier per tile) and
//                publishes its output elements as LL words.
// There is no grid barrier: phases are chained by the data itself (value + epoch in one 8-byte word).
#pragma once
#include "fq3_common.cuh"

namespace fq3 {

// =================================================================================================
// PTX helpers
// =================================================================================================
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cnt(uint64_t* bar, uint32_t n) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}

__device__ __forceinline__ void cbar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory"); }

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ int ld_volatile_shared_i32(const int* p) {
  int v;
  asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_shared_i32(int* p, int v) {
  asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}

// LL words: relaxed gpu-scope 8-byte accesses served by L2 (the coherence point).
__device__ __forceinline__ LLWord ll_ld(const LLWord* p) {
  LLWord w;
  asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(w.x), "=r"(w.y) : "l"(p) : "memory");
  return w;
}
// two adjacent words with one 16-byte request (each word is validated by its own epoch, so tearing is harmless)
__device__ __forceinline__ void ll_ld2(const LLWord* p, LLWord& a, LLWord& b) {
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(b.x), "=r"(b.y) : "l"(p) : "memory");
}
__device__ __forceinline__ void ll_st(LLWord* p, float v, uint32_t ep) {
  asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(ep) : "memory");
}

__device__ __forceinline__ float bf16r(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// 8 bf16 products into two independent accumulators (halves the dependent FMA chain)
__device__ __forceinline__ void dot8(const uint4& w, const uint4& x, float& a0, float& a1) {
  a0 = fmaf(bf_lo(w.x), bf_lo(x.x), a0);
  a1 = fmaf(bf_hi(w.x), bf_hi(x.x), a1);
  a0 = fmaf(bf_lo(w.y), bf_lo(x.y), a0);
  a1 = fmaf(bf_hi(w.y), bf_hi(x.y), a1);
  a0 = fmaf(bf_lo(w.z), bf_lo(x.z), a0);
  a1 = fmaf(bf_hi(w.z), bf_hi(x.z), a1);
  a0 = fmaf(bf_lo(w.w), bf_lo(x.w), a0);
  a1 = fmaf(bf_hi(w.w), bf_hi(x.w), a1);
}

// =================================================================================================
// Shared-memory layout
// =================================================================================================
struct Smem {
  uint64_t* full;    // [kMaxStages]
  uint64_t* empty;   // [kMaxStages]
  int* ctl;          // [0] producer go flag, [4..] broadcast scratch
  unsigned char* scratch;
  Phase* prog;
  unsigned char* xbuf;
  unsigned char* ring;
};

__device__ __forceinline__ Smem carve_smem(unsigned char* base, const LaunchParams& p) {
  Smem s;
  s.full = reinterpret_cast<uint64_t*>(base);
  s.empty = s.full + kMaxStages;
  s.ctl = reinterpret_cast<int*>(base + kCtlOffset);
  s.scratch = base + kHeaderBytes;
  s.prog = reinterpret_cast<Phase*>(s.scratch + kScratchBytes);
  s.xbuf = reinterpret_cast<unsigned char*>(s.prog) + p.prog_bytes;
  s.ring = s.xbuf + p.xbuf_bytes;
  return s;
}

// =================================================================================================
// Watchdog: a protocol bug must surface as an error, never as a hung GPU.
// =================================================================================================
__device__ __noinline__ void device_fault(const LaunchParams& p, int code, int phase, int detail) {
  volatile int* e = p.err;
  if (e[0] == 0) {
    e[1] = blockIdx.x;
    e[2] = phase;
    e[3] = detail;
    e[0] = code;
  }
  __threadfence_system();
  asm volatile("trap;");
}

struct Spin {
  unsigned long long t0;
  unsigned n;
  __device__ __forceinline__ Spin() : t0(0), n(0) {}
  __device__ __forceinline__ void tick(const LaunchParams& p, int code, int phase, int detail) {
    if ((++n & 0x3ff) == 0) {
      unsigned long long t = globaltimer_ns();
      if (t0 == 0) t0 = t;
      else if (t - t0 > p.watchdog_ns) device_fault(p, code, phase, detail);
    }
  }
};

__device__ __noinline__ void mbar_wait_slow(uint64_t* bar, uint32_t parity, const LaunchParams& p, int code, int phase) {
  Spin s;
  while (!mbar_try_wait(bar, parity)) s.tick(p, code, phase, (int)parity);
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, const LaunchParams& p, int code, int phase) {
  if (mbar_try_wait(bar, parity)) return;
  mbar_wait_slow(bar, parity, p, code, phase);
}

// Slow path of an LL read: poll until the word carries `ep`.
__device__ __noinline__ LLWord ll_spin(const LLWord* ptr, uint32_t ep, const LaunchParams& p, int phase) {
  Spin s;
  LLWord w;
  unsigned ns = 64;
  do {
    // back off: thousands of threads polling L2 at full rate starve the weight stream of L2 request slots
    __nanosleep(ns);
    if (ns < 256) ns += 64;
    s.tick(p, DE_LL_WAIT, phase, (int)ep);
    w = ll_ld(ptr);
  } while (w.y != ep);
  return w;
}
// Validate a word that was loaded earlier (loads are issued in batches so their L2 round trips overlap);
// ep == 0 accepts whatever is there — inputs written before the launch.
__device__ __forceinline__ float ll_check(LLWord w, const LLWord* ptr, uint32_t ep, const LaunchParams& p, int phase) {
  if (ep != 0 && w.y != ep && !(p.debug & 1)) w = ll_spin(ptr, ep, p, phase);
  return __uint_as_float(w.x);
}
__device__ __forceinline__ float ll_wait(const LLWord* ptr, uint32_t ep, const LaunchParams& p, int phase) {
  return ll_check(ll_ld(ptr), ptr, ep, p, phase);
}

__device__ __forceinline__ void prof_mark(const LaunchParams& p, int pidx, int slot) {
  if (p.prof && threadIdx.x == 0 && (int)blockIdx.x == p.prof_cta && pidx >= 0) p.prof[(size_t)pidx * 8 + slot] = clock64();
}

// =================================================================================================
// GEMV phase
// =================================================================================================
struct GemvPlan {
  int r0, r1, rt, ntiles, rpu;
};
__device__ __forceinline__ GemvPlan gemv_plan(const Phase& ph, const LaunchParams& p, int cta, int G) {
  GemvPlan g;
  g.rpu = (ph.flags & F_SWIGLU) ? 2 : 1;
  const int N = (int)ph.N, K = (int)ph.K;
  const int nunits = N / g.rpu;
  const int base = nunits / G, rem = nunits - base * G;  // balanced: the first `rem` CTAs take one more unit
  const int u0 = cta * base + min(cta, rem);
  const int u1 = u0 + base + (cta < rem ? 1 : 0);
  g.r0 = u0 * g.rpu;
  g.r1 = u1 * g.rpu;
  int rt = min(kRowsPerTileMax, p.stage_bytes / (K * 2));
  rt -= rt % g.rpu;
  g.rt = max(rt, g.rpu);
  g.ntiles = (g.r1 - g.r0 + g.rt - 1) / g.rt;
  return g;
}

__device__ __forceinline__ int phase_rows(const Phase& ph, const LaunchParams& p) {
  return (ph.flags & F_ROWS2) ? 2 * p.n_rows : p.n_rows;
}

// Load the activation rows of a GEMV phase into shared memory (bf16), optionally RMS-normalised.
// HF rounding points (Qwen3RMSNorm): fp32 mean-square, x*rsqrt -> bf16, * weight -> bf16.
// All 512 consumer threads take part: one LL round trip, gamma fetched alongside.
__device__ __noinline__ void load_x(const Phase& ph, const LaunchParams& p, unsigned char* smem_base, int M, int row_off, uint32_t ep_in,
                       int pidx) {
  const Smem sm = carve_smem(smem_base, p);
  const int K = (int)ph.K;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const LLWord* in = reinterpret_cast<const LLWord*>(p.bufs[ph.in_buf]);
  const int ld = p.ld[ph.in_buf];
  bf16* xs = reinterpret_cast<bf16*>(sm.xbuf);
  const bool norm = (ph.flags & F_PRENORM) != 0;
  float* red = reinterpret_cast<float*>(sm.scratch);  // [kMaxRows][kConsumerWarps]
  constexpr int kPairs = 7;  // K <= 6720: each thread owns the element pairs 2*(tid + i*480), +1 of every row
  if (!norm) {
    for (int m = 0; m < M; ++m) {
      const LLWord* src = in + (size_t)(m + row_off) * ld;
      LLWord w[kPairs][2];
#pragma unroll
      for (int i = 0; i < kPairs; ++i) {
        const int k = 2 * (threadIdx.x + i * kConsumerThreads);
        if (k < K) ll_ld2(src + k, w[i][0], w[i][1]);
      }
#pragma unroll
      for (int i = 0; i < kPairs; ++i) {
        const int k = 2 * (threadIdx.x + i * kConsumerThreads);
        if (k < K) {
          const float v0 = ll_check(w[i][0], src + k, ep_in, p, pidx), v1 = ll_check(w[i][1], src + k + 1, ep_in, p, pidx);
          *reinterpret_cast<uint32_t*>(xs + (size_t)m * K + k) = pack_bf16x2(v0, v1);
        }
      }
    }
    cbar_sync();
    return;
  }
  const bf16* gamma = (ph.flags & F_ABSPTR) ? reinterpret_cast<const bf16*>(p.lin_gamma)
                                            : reinterpret_cast<const bf16*>(p.arena + (size_t)ph.g_off * 16);
  const float eps = (ph.flags & F_ABSPTR) ? p.lin_eps : p.stacks[ph.stack].eps;
  for (int m = 0; m < M; ++m) {
    const LLWord* src = in + (size_t)(m + row_off) * ld;
    float v[kPairs][2];
    uint32_t g[kPairs];
    LLWord w[kPairs][2];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < kPairs; ++i) {
      const int k = 2 * (threadIdx.x + i * kConsumerThreads);
      if (k < K) {
        ll_ld2(src + k, w[i][0], w[i][1]);
        g[i] = __ldg(reinterpret_cast<const uint32_t*>(gamma + k));
      }
    }
#pragma unroll
    for (int i = 0; i < kPairs; ++i) {
      const int k = 2 * (threadIdx.x + i * kConsumerThreads);
      if (k < K) {
        v[i][0] = ll_check(w[i][0], src + k, ep_in, p, pidx);
        v[i][1] = ll_check(w[i][1], src + k + 1, ep_in, p, pidx);
        ss = fmaf(v[i][0], v[i][0], ss);
        ss = fmaf(v[i][1], v[i][1], ss);
      }
    }
    ss = warp_sum(ss);
    if (lane == 0) red[m * kConsumerWarps + warp] = ss;
    cbar_sync();
    float tot = 0.f;
#pragma unroll
    for (int wi = 0; wi < kConsumerWarps; ++wi) tot += red[m * kConsumerWarps + wi];
    const float rs = rsqrtf(tot / (float)K + eps);
    const bool wr = (ph.flags & F_WRITE_NORMED) && (int)blockIdx.x == (m % (int)gridDim.x);
#pragma unroll
    for (int i = 0; i < kPairs; ++i) {
      const int k = 2 * (threadIdx.x + i * kConsumerThreads);
      if (k < K) {
        const float y0 = bf16r(bf16r(v[i][0] * rs) * bf_lo(g[i])), y1 = bf16r(bf16r(v[i][1] * rs) * bf_hi(g[i]));
        const uint32_t pk = pack_bf16x2(y0, y1);
        *reinterpret_cast<uint32_t*>(xs + (size_t)m * K + k) = pk;
        if (wr) *reinterpret_cast<uint32_t*>(reinterpret_cast<bf16*>(p.bufs[BUF_HID]) + (size_t)m * p.ld[BUF_HID] + k) = pk;
      }
    }
  }
  cbar_sync();
}

__device__ __forceinline__ float silu_f(float x) { return x / (1.f + expf(-x)); }

// One warp computes one unit (1 row, or a gate/up pair) of one tile for MT activation rows and publishes it.
template <int MT, int RPU>
__device__ __noinline__ void gemv_unit(const Phase& ph, const LaunchParams& p, const uint4* __restrict__ wrow,
                                          const uint4* __restrict__ xs, int cpr, int out_col, int M, uint32_t ep) {
  const int lane = threadIdx.x & 31;
  // residual value: issue the L2 read now, consume it in the epilogue (hides the round trip behind the dot product)
  LLWord resw = make_uint2(0u, 0u);
  if ((ph.flags & F_RESID) && lane < M)
    resw = ll_ld(reinterpret_cast<const LLWord*>(p.bufs[ph.res_buf]) + (size_t)lane * p.ld[ph.res_buf] + out_col);
  float a0[RPU][MT], a1[RPU][MT];
#pragma unroll
  for (int r = 0; r < RPU; ++r)
#pragma unroll
    for (int m = 0; m < MT; ++m) a0[r][m] = a1[r][m] = 0.f;
#pragma unroll 2
  for (int c = lane; c < cpr; c += 32) {
    uint4 w[RPU];
#pragma unroll
    for (int r = 0; r < RPU; ++r) w[r] = wrow[(size_t)r * cpr + c];
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      const uint4 x = xs[(size_t)m * cpr + c];
#pragma unroll
      for (int r = 0; r < RPU; ++r) dot8(w[r], x, a0[r][m], a1[r][m]);
    }
  }
  float y[RPU];
#pragma unroll
  for (int r = 0; r < RPU; ++r) y[r] = 0.f;
#pragma unroll
  for (int m = 0; m < MT; ++m) {
#pragma unroll
    for (int r = 0; r < RPU; ++r) {
      const float v = warp_sum(a0[r][m] + a1[r][m]);
      if (lane == m) y[r] = v;
    }
  }
  if (lane < M) {
    // epilogue with PyTorch's bf16 rounding points; lane m owns activation row m
    const int m = lane;
    float out;
    if (RPU == 2) {
      const float g = bf16r(y[0]), u = bf16r(y[1]);
      out = bf16r(bf16r(silu_f(g)) * u);
    } else {
      out = y[0];
      if (ph.flags & F_BIAS) {
        const bf16* bias = (ph.flags & F_ABSPTR) ? reinterpret_cast<const bf16*>(p.lin_bias)
                                                 : reinterpret_cast<const bf16*>(p.arena + (size_t)ph.b_off * 16);
        out += __bfloat162float(bias[out_col]);
      }
      out = bf16r(out);
      if (ph.flags & F_SILU) out = bf16r(silu_f(out));
    }
    if (ph.flags & F_RESID) out = bf16r(__uint_as_float(resw.x) + out);
    ll_st(reinterpret_cast<LLWord*>(p.bufs[ph.out_buf]) + (size_t)m * p.ld[ph.out_buf] + out_col, out, ep);
  }
}

__device__ void gemv_phase_consume(const Phase& ph, const LaunchParams& p, unsigned char* smem_base, unsigned& tile_it, int pidx,
                                   uint32_t ep) {
  const Smem sm = carve_smem(smem_base, p);
  const int G = gridDim.x, cta = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const GemvPlan g = gemv_plan(ph, p, cta, G);
  int M = phase_rows(ph, p), row_off = 0;
  if (ph.flags & F_LAST_ROW) { row_off = M - 1; M = 1; }
  const uint32_t ep_in = (pidx == 0 && ep == p.epoch_base + 1) ? 0u : ep - 1;
  prof_mark(p, pidx, 0);
  if (!(p.debug & 8)) load_x(ph, p, smem_base, M, row_off, ep_in, pidx);
  prof_mark(p, pidx, 1);
  const int cpr = (int)ph.K >> 3;
  const uint4* xs = reinterpret_cast<const uint4*>(sm.xbuf);
  const int upt = g.rt / g.rpu;  // units per tile
  const int nunits = (g.r1 - g.r0) / g.rpu;
  // unit u of the CTA's slice belongs to warp (u % 15).  A warp only touches the tiles that hold its units; the
  // producer pre-arrives on the stage's empty barrier for the warps that have no unit in a tile.
  // mbarrier parity only disambiguates adjacent phases, so no warp may run a full ring ahead of the data:
  // tiles are consumed in rounds of n_stages with a consumer barrier between rounds (one round for the
  // 0.6B shapes, up to three for 12 KB rows).
  const int nrounds = (g.ntiles + p.n_stages - 1) / p.n_stages;
  int u = warp;
#pragma unroll 1
  for (int round = 0; round < nrounds; ++round) {
    if (round > 0) cbar_sync();
    const int t_end = min(g.ntiles, (round + 1) * p.n_stages);
    int t_prev = -1;
    int stage = 0;
#pragma unroll 1
    for (; u < nunits && u / upt < t_end; u += kConsumerWarps) {
      const int t = u / upt;
      const unsigned it = tile_it + (unsigned)t;
      if (t != t_prev) {
        if (t_prev >= 0) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.empty[stage]);
        }
        stage = it % p.n_stages;
        if (lane == 0) mbar_wait(&sm.full[stage], (it / p.n_stages) & 1u, p, DE_FULL_WAIT, pidx);
        __syncwarp();
        t_prev = t;
      }
      if (!(p.debug & 2)) {
        const int j = u - t * upt;
        const uint4* wrow = reinterpret_cast<const uint4*>(sm.ring + (size_t)stage * p.stage_bytes) + (size_t)j * g.rpu * cpr;
        const int oc = g.r0 / g.rpu + u;
        if (g.rpu == 2) {
          if (M == 1) gemv_unit<1, 2>(ph, p, wrow, xs, cpr, oc, M, ep);
          else if (M == 2) gemv_unit<2, 2>(ph, p, wrow, xs, cpr, oc, M, ep);
          else if (M <= 4) gemv_unit<4, 2>(ph, p, wrow, xs, cpr, oc, M, ep);
          else gemv_unit<8, 2>(ph, p, wrow, xs, cpr, oc, M, ep);
        } else {
          if (M == 1) gemv_unit<1, 1>(ph, p, wrow, xs, cpr, oc, M, ep);
          else if (M == 2) gemv_unit<2, 1>(ph, p, wrow, xs, cpr, oc, M, ep);
          else if (M <= 4) gemv_unit<4, 1>(ph, p, wrow, xs, cpr, oc, M, ep);
          else gemv_unit<8, 1>(ph, p, wrow, xs, cpr, oc, M, ep);
        }
      }
    }
    if (t_prev >= 0) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.empty[stage]);
    }
  }
  tile_it += (unsigned)g.ntiles;
  prof_mark(p, pidx, 2);
  cbar_sync();  // xbuf is rewritten by the next phase
  prof_mark(p, pidx, 3);
}

// Producer side of one GEMV phase: stream this CTA's rows through the ring.
__device__ __noinline__ void gemv_phase_produce(const Phase& ph, const LaunchParams& p, unsigned char* smem_base, unsigned& tile_it, int pidx,
                                   uint64_t pol_stream, uint64_t pol_keep) {
  const Smem sm = carve_smem(smem_base, p);
  const GemvPlan g = gemv_plan(ph, p, blockIdx.x, gridDim.x);
  const unsigned char* W = (ph.flags & F_ABSPTR) ? reinterpret_cast<const unsigned char*>(p.lin_W)
                                                 : p.arena + (size_t)ph.w_off * 16;
  const size_t row_bytes = (size_t)ph.K * 2;
  const uint64_t pol = (ph.flags & F_L2_KEEP) ? pol_keep : pol_stream;
  if (p.prof && (int)blockIdx.x == p.prof_cta) p.prof[(size_t)pidx * 8 + 4] = clock64();
  for (int t = 0; t < g.ntiles; ++t, ++tile_it) {
    const int stage = tile_it % p.n_stages;
    const uint32_t parity = ((tile_it / p.n_stages) & 1u) ^ 1u;
    const int row_base = g.r0 + t * g.rt;
    const int rt = min(g.rt, g.r1 - row_base);
    const uint32_t bytes = (uint32_t)(rt * row_bytes);
    mbar_wait(&sm.empty[stage], parity, p, DE_EMPTY_WAIT, pidx);
    if (t == g.ntiles - 1 && p.prof && (int)blockIdx.x == p.prof_cta) p.prof[(size_t)pidx * 8 + 5] = clock64();
    mbar_arrive_expect_tx(&sm.full[stage], bytes);
    bulk_g2s(sm.ring + (size_t)stage * p.stage_bytes, W + (size_t)row_base * row_bytes, bytes, &sm.full[stage], pol);
    const int readers = min(rt / g.rpu, kConsumerWarps);  // units in this tile go to distinct warps (round-robin)
    if (readers < kConsumerWarps) mbar_arrive_cnt(&sm.empty[stage], (uint32_t)(kConsumerWarps - readers));
  }
}

// =================================================================================================
// Attention phase: q/k RMSNorm + RoPE + KV append + GQA decode attention (+ split-KV combine).
// =================================================================================================
struct Group {
  int first_row, nrows, slot, pos0, n_pad, rope_delta;
};
__device__ __forceinline__ int num_groups(const LaunchParams& p) { return p.mode == MODE_PREFILL ? 1 : p.n_rows; }
__device__ __forceinline__ Group get_group(const Phase& ph, const LaunchParams& p, int g, const int* frame_pos) {
  Group r;
  if (p.mode == MODE_PREFILL) {
    r.first_row = 0; r.nrows = p.n_rows; r.slot = p.stream0; r.pos0 = p.pf_pos0; r.n_pad = p.pf_n_pad;
    r.rope_delta = p.pf_rope_delta;
    return r;
  }
  r.nrows = (ph.flags & F_ROWS2) ? 2 : 1;
  r.first_row = g * r.nrows;
  r.slot = p.stream0 + g;
  if (ph.stack == ST_TALKER) {
    const StreamState* st = p.st + r.slot;
    r.pos0 = (p.pos_override >= 0) ? p.pos_override : frame_pos[g];
    r.n_pad = __ldcg(&st->n_pad);
    r.rope_delta = __ldcg(&st->rope_delta);
  } else {
    r.pos0 = (ph.flags & F_ROWS2) ? 0 : (int)ph.aux + 1;
    r.n_pad = 0;
    r.rope_delta = 0;
  }
  return r;
}
// one CTA handles up to 1024 positions of one (sequence, kv head); longer contexts are split
__device__ __forceinline__ int num_splits(int L, int ngroups, int nkv, int G) {
  int cap = G / max(1, ngroups * nkv);
  cap = max(1, min(cap, kMaxSplits));
  int want = (L + 1023) / 1024;
  return max(1, min(want, cap));
}

// RMSNorm over the 128-wide head + rotary embedding; lane owns elements [4*lane, 4*lane+4).
// gamma / cos / sin arrive as 4 packed bf16 (uint2) loaded before the activation wait.
__device__ __forceinline__ void head_norm_rope(float (&x)[4], uint2 g2, float eps, uint2 c2, uint2 s2, int lane) {
  float ss = x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3];
  ss = warp_sum(ss);
  const float rs = rsqrtf(ss * (1.f / kHeadDim) + eps);
  const float g[4] = {bf_lo(g2.x), bf_hi(g2.x), bf_lo(g2.y), bf_hi(g2.y)};
  const float c[4] = {bf_lo(c2.x), bf_hi(c2.x), bf_lo(c2.y), bf_hi(c2.y)};
  const float sn[4] = {bf_lo(s2.x), bf_hi(s2.x), bf_lo(s2.y), bf_hi(s2.y)};
  float y[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) y[i] = bf16r(bf16r(x[i] * rs) * g[i]);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float partner = __shfl_xor_sync(0xffffffffu, y[i], 16);
    const float rot = (lane < 16) ? -partner : partner;
    x[i] = bf16r(bf16r(y[i] * c[i]) + bf16r(rot * sn[i]));
  }
}
__device__ __forceinline__ uint2 ld_bf16x4(const bf16* p) { return __ldg(reinterpret_cast<const uint2*>(p)); }

__device__ __forceinline__ void ll_issue4(const LLWord* ptr, LLWord (&w)[4]) {
  ll_ld2(ptr, w[0], w[1]);
  ll_ld2(ptr + 2, w[2], w[3]);
}
__device__ __forceinline__ void ll_finish4(const LLWord (&w)[4], const LLWord* ptr, uint32_t ep, const LaunchParams& p, int pidx,
                                           float (&x)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = ll_check(w[i], ptr + i, ep, p, pidx);
}

constexpr int kAttnPairsMax = 4;  // (row, q-head) pairs sharing one K/V sweep

// Scratch layout (floats): qs[16][128] | wp[16 warps][kAttnPairsMax][kPartStride]
__device__ __noinline__ void attn_item(const Phase& ph, const LaunchParams& p, unsigned char* smem_base, const Group& gr, int kvh, int sp,
                          int nsplit, uint32_t ep, int pidx) {
  const Smem sm = carve_smem(smem_base, p);
  const StackRt& S = p.stacks[ph.stack];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gq = S.nq / S.nkv;
  const int npairs = gr.nrows * gq;
  const int qb = ph.stack == ST_TALKER ? BUF_TQKV : BUF_PQKV, ab = ph.stack == ST_TALKER ? BUF_TATT : BUF_PATT;
  const LLWord* qkv = reinterpret_cast<const LLWord*>(p.bufs[qb]);
  const int ldq = p.ld[qb];
  LLWord* att = reinterpret_cast<LLWord*>(p.bufs[ab]);
  const int lda = p.ld[ab];
  const bf16* qn = reinterpret_cast<const bf16*>(p.arena + (size_t)ph.g_off * 16);
  const bf16* kn = reinterpret_cast<const bf16*>(p.arena + (size_t)ph.b_off * 16);
  const size_t head_base = (((size_t)ph.layer * S.n_slots + gr.slot) * S.nkv + kvh) * (size_t)S.max_pos * kHeadDim;
  bf16* Kc = S.kcache + head_base;
  bf16* Vc = S.vcache + head_base;
  const int pos_last = gr.pos0 + gr.nrows - 1;
  const int L = pos_last + 1 - gr.n_pad;
  const int per = (L + nsplit - 1) / nsplit;
  const int a = gr.n_pad + sp * per;
  const int b = min(a + per, gr.n_pad + L);
  float* qs = reinterpret_cast<float*>(sm.scratch);
  float* wp = qs + 16 * kHeadDim;
  const float scale = rsqrtf((float)kHeadDim);
  const uint32_t ep_in = ep - 1;
  prof_mark(p, pidx, 0);

  // -- step 1: queries (norm + rope, pre-scaled) to smem; new K/V rows of this chunk to the cache.
  //    Work items 0..npairs-1 are queries, npairs..npairs+nrows-1 are k/v rows; one warp each.
  for (int it = warp; it < npairs + gr.nrows; it += kConsumerWarps) {
    if (it < npairs) {
      const int j = it;
      const int r = j / gq, qh = kvh * gq + (j - r * gq);
      const int rp = min(max(gr.pos0 + r + gr.rope_delta, 0), S.rope_len - 1);
      const LLWord* src = qkv + (size_t)(gr.first_row + r) * ldq + qh * kHeadDim + lane * 4;
      LLWord w[4];
      ll_issue4(src, w);
      const uint2 g2 = ld_bf16x4(qn + lane * 4);
      const uint2 c2 = ld_bf16x4(S.rope_cos + (size_t)rp * kHeadDim + lane * 4);
      const uint2 s2 = ld_bf16x4(S.rope_sin + (size_t)rp * kHeadDim + lane * 4);
      float x[4];
      ll_finish4(w, src, ep_in, p, pidx, x);
      head_norm_rope(x, g2, S.eps, c2, s2, lane);
      *reinterpret_cast<float4*>(qs + j * kHeadDim + lane * 4) =
          make_float4(x[0] * scale, x[1] * scale, x[2] * scale, x[3] * scale);
    } else {
      const int r = it - npairs;
      const int pos = gr.pos0 + r;
      if (pos >= a && pos < b) {
        const int rp = min(max(pos + gr.rope_delta, 0), S.rope_len - 1);
        const LLWord* row = qkv + (size_t)(gr.first_row + r) * ldq;
        const LLWord* ksrc = row + S.nq * kHeadDim + kvh * kHeadDim + lane * 4;
        const LLWord* vsrc = row + (S.nq + S.nkv) * kHeadDim + kvh * kHeadDim + lane * 4;
        LLWord wk[4], wv[4];
        ll_issue4(ksrc, wk);
        ll_issue4(vsrc, wv);
        const uint2 g2 = ld_bf16x4(kn + lane * 4);
        const uint2 c2 = ld_bf16x4(S.rope_cos + (size_t)rp * kHeadDim + lane * 4);
        const uint2 s2 = ld_bf16x4(S.rope_sin + (size_t)rp * kHeadDim + lane * 4);
        float x[4], v[4];
        ll_finish4(wk, ksrc, ep_in, p, pidx, x);
        ll_finish4(wv, vsrc, ep_in, p, pidx, v);
        head_norm_rope(x, g2, S.eps, c2, s2, lane);
        uint2 kk, vv;
        kk.x = pack_bf16x2(x[0], x[1]); kk.y = pack_bf16x2(x[2], x[3]);
        vv.x = pack_bf16x2(v[0], v[1]); vv.y = pack_bf16x2(v[2], v[3]);
        *reinterpret_cast<uint2*>(Kc + (size_t)pos * kHeadDim + lane * 4) = kk;
        *reinterpret_cast<uint2*>(Vc + (size_t)pos * kHeadDim + lane * 4) = vv;
      }
    }
  }
  cbar_sync();
  prof_mark(p, pidx, 1);

  // -- step 2: pairs in batches of <= 4 share one sweep over K/V.  QK^T: one lane per position (no shuffles);
  //            P·V: lane owns 4 output dims, probabilities broadcast by shuffle.  Warp w sweeps positions
  //            a+32w, a+32w+512, ...
  for (int j0 = 0; j0 < npairs; j0 += kAttnPairsMax) {
    const int np = min(kAttnPairsMax, npairs - j0);
    float m_run[kAttnPairsMax], l_run[kAttnPairsMax], acc[kAttnPairsMax][4];
    int pend[kAttnPairsMax];
#pragma unroll
    for (int j = 0; j < kAttnPairsMax; ++j) {
      m_run[j] = -INFINITY; l_run[j] = 0.f;
      acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
      const int r = (j0 + j) / gq;
      pend[j] = (j < np) ? min(b, gr.pos0 + r + 1) : 0;  // causal bound of this pair's row
    }
    for (int pb = a + warp * 32; pb < b; pb += 32 * kConsumerWarps) {
      const int pos = pb + lane;
      const bool valid = pos < b;
      float s[kAttnPairsMax];
#pragma unroll
      for (int j = 0; j < kAttnPairsMax; ++j) s[j] = 0.f;
      if (valid) {
        const uint4* krow = reinterpret_cast<const uint4*>(Kc + (size_t)pos * kHeadDim);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint4 kreg[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) kreg[c] = __ldcg(krow + h * 8 + c);  // 8 independent loads in flight
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 kk = kreg[c];
            const float k0 = bf_lo(kk.x), k1 = bf_hi(kk.x), k2 = bf_lo(kk.y), k3 = bf_hi(kk.y);
            const float k4 = bf_lo(kk.z), k5 = bf_hi(kk.z), k6 = bf_lo(kk.w), k7 = bf_hi(kk.w);
#pragma unroll
            for (int j = 0; j < kAttnPairsMax; ++j) {
              if (j < np) {
                const float* qp = qs + (j0 + j) * kHeadDim + (h * 8 + c) * 8;
                const float4 qa = *reinterpret_cast<const float4*>(qp);
                const float4 qb4 = *reinterpret_cast<const float4*>(qp + 4);
                s[j] += qa.x * k0 + qa.y * k1 + qa.z * k2 + qa.w * k3 + qb4.x * k4 + qb4.y * k5 + qb4.z * k6 + qb4.w * k7;
              }
            }
          }
        }
      }
      float pr[kAttnPairsMax];
#pragma unroll
      for (int j = 0; j < kAttnPairsMax; ++j) {
        const bool ok = valid && (j < np) && (pos < pend[j]);
        const float sv = ok ? s[j] : -INFINITY;
        const float m_new = fmaxf(m_run[j], warp_max(sv));
        const float corr = (m_run[j] == -INFINITY) ? 0.f : expf(m_run[j] - m_new);
        pr[j] = ok ? expf(sv - m_new) : 0.f;
        l_run[j] = l_run[j] * corr + warp_sum(pr[j]);
        acc[j][0] *= corr; acc[j][1] *= corr; acc[j][2] *= corr; acc[j][3] *= corr;
        m_run[j] = m_new;
      }
      const int nvalid = min(32, b - pb);
      for (int i0 = 0; i0 < nvalid; i0 += 8) {
        uint2 vreg[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (i0 + u < nvalid) vreg[u] = __ldcg(reinterpret_cast<const uint2*>(Vc + (size_t)(pb + i0 + u) * kHeadDim + lane * 4));
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (i0 + u < nvalid) {
            const float v0 = bf_lo(vreg[u].x), v1 = bf_hi(vreg[u].x), v2 = bf_lo(vreg[u].y), v3 = bf_hi(vreg[u].y);
#pragma unroll
            for (int j = 0; j < kAttnPairsMax; ++j) {
              if (j < np) {
                const float pj = __shfl_sync(0xffffffffu, pr[j], i0 + u);
                acc[j][0] = fmaf(pj, v0, acc[j][0]); acc[j][1] = fmaf(pj, v1, acc[j][1]);
                acc[j][2] = fmaf(pj, v2, acc[j][2]); acc[j][3] = fmaf(pj, v3, acc[j][3]);
              }
            }
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kAttnPairsMax; ++j) {
      if (j < np) {
        float* w = wp + ((size_t)warp * kAttnPairsMax + j) * kPartStride;
        if (lane == 0) { w[0] = m_run[j]; w[1] = l_run[j]; }
        *reinterpret_cast<float4*>(w + 4 + lane * 4) = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
      }
    }
    prof_mark(p, pidx, 2);
    cbar_sync();
    // merge the warp partials of each pair (fixed order => deterministic)
    for (int o = threadIdx.x; o < np * kHeadDim; o += kConsumerThreads) {
      const int j = o / kHeadDim, dd = o - j * kHeadDim;
      const int r = (j0 + j) / gq, qh = kvh * gq + ((j0 + j) - r * gq);
      const int row = gr.first_row + r;
      float Mx = -INFINITY;
#pragma unroll
      for (int w = 0; w < kConsumerWarps; ++w) Mx = fmaxf(Mx, wp[((size_t)w * kAttnPairsMax + j) * kPartStride]);
      float Lsum = 0.f, O = 0.f;
#pragma unroll
      for (int w = 0; w < kConsumerWarps; ++w) {
        const float* q = wp + ((size_t)w * kAttnPairsMax + j) * kPartStride;
        const float f = (q[0] == -INFINITY) ? 0.f : expf(q[0] - Mx);
        Lsum += q[1] * f;
        O += q[4 + dd] * f;
      }
      if (nsplit == 1) {
        ll_st(att + (size_t)row * lda + qh * kHeadDim + dd, bf16r(Lsum > 0.f ? O / Lsum : 0.f), ep);
      } else {
        float* dst = p.attn_part + (((size_t)row * S.nq + qh) * kMaxSplits + sp) * kPartStride;
        if (dd == 0) { dst[0] = Mx; dst[1] = Lsum; }
        dst[4 + dd] = O;
      }
    }
    cbar_sync();
  }
  prof_mark(p, pidx, 3);
  if (nsplit > 1) {
    // -- step 3: last-arriving split of this (sequence, kv head) combines the partials
    __threadfence();
    cbar_sync();
    int* flag = sm.ctl + 4;
    if (threadIdx.x == 0) {
      unsigned* cnt = p.attn_cnt + (gr.slot - p.stream0) * S.nkv + kvh;
      const unsigned old = atomicAdd(cnt, 1u);
      const int last = (old == (unsigned)(nsplit - 1));
      if (last) { *cnt = 0; __threadfence(); }
      *flag = last;
    }
    cbar_sync();
    if (*flag) {
      for (int o = threadIdx.x; o < npairs * kHeadDim; o += kConsumerThreads) {
        const int j = o / kHeadDim, dd = o - j * kHeadDim;
        const int r = j / gq, qh = kvh * gq + (j - r * gq);
        const int row = gr.first_row + r;
        const float* src = p.attn_part + (((size_t)row * S.nq + qh) * kMaxSplits) * kPartStride;
        float Mx = -INFINITY;
        for (int s = 0; s < nsplit; ++s) Mx = fmaxf(Mx, __ldcg(src + s * kPartStride));
        float Lsum = 0.f, O = 0.f;
        for (int s = 0; s < nsplit; ++s) {
          const float ms = __ldcg(src + s * kPartStride);
          const float f = (ms == -INFINITY) ? 0.f : expf(ms - Mx);
          Lsum += __ldcg(src + s * kPartStride + 1) * f;
          O += __ldcg(src + s * kPartStride + 4 + dd) * f;
        }
        ll_st(att + (size_t)row * lda + qh * kHeadDim + dd, bf16r(Lsum > 0.f ? O / Lsum : 0.f), ep);
      }
    }
    cbar_sync();
  }
}

__device__ void attn_phase(const Phase& ph, const LaunchParams& p, unsigned char* smem_base, uint32_t ep, int pidx,
                           const int* frame_pos) {
  const Smem sm = carve_smem(smem_base, p);
  const StackRt& S = p.stacks[ph.stack];
  const int G = gridDim.x, cta = blockIdx.x;
  const int ng = num_groups(p);
  // item -> CTA: spread items over the whole grid (stride) so concurrent items sit on distant SMs
  int total = 0;
  for (int g = 0; g < ng; ++g) {
    const Group gr = get_group(ph, p, g, frame_pos);
    total += S.nkv * num_splits(gr.pos0 + gr.nrows - gr.n_pad, ng, S.nkv, G);
  }
  const int stride = max(1, G / max(1, total));
  int item = 0;
  bool any = false;
  for (int g = 0; g < ng; ++g) {
    const Group gr = get_group(ph, p, g, frame_pos);
    const int L = gr.pos0 + gr.nrows - gr.n_pad;
    const int nsplit = num_splits(L, ng, S.nkv, G);
    const int nitems = S.nkv * nsplit;
    for (int it = 0; it < nitems; ++it) {
      if (((item + it) * stride) % G == cta) {
        attn_item(ph, p, smem_base, gr, it / nsplit, it % nsplit, nsplit, ep, pidx);
        any = true;
      }
    }
    item += nitems;
  }
  (void)any;  // the KV rows written here are fenced once per step in sample_phase (off the critical path)
}

// =================================================================================================
// Sampling (sampling.py:10-66) — block-wide, 512 consumer threads, one stream per CTA
// =================================================================================================
__device__ __forceinline__ uint32_t f2key(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
  uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);
}
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

struct SampleScratch {
  float* sl;       // [V] working logits
  unsigned* hist;  // [256]
  float* redf;     // [32]
  int* redi;       // [32]
};
__device__ __forceinline__ SampleScratch sample_scratch(unsigned char* scratch) {
  SampleScratch s;
  s.sl = reinterpret_cast<float*>(scratch);
  s.hist = reinterpret_cast<unsigned*>(scratch + 20 * 1024);
  s.redf = reinterpret_cast<float*>(scratch + 21 * 1024);
  s.redi = reinterpret_cast<int*>(scratch + 21 * 1024 + 256);
  return s;
}

__device__ __noinline__ float block_sum(float v, float* red) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  v = warp_sum(v);
  cbar_sync();
  if (lane == 0) red[warp] = v;
  cbar_sync();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < kConsumerWarps; ++w) t += red[w];
  return t;
}
__device__ __noinline__ float block_max(float v, float* red) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  v = warp_max(v);
  cbar_sync();
  if (lane == 0) red[warp] = v;
  cbar_sync();
  float t = -INFINITY;
#pragma unroll
  for (int w = 0; w < kConsumerWarps; ++w) t = fmaxf(t, red[w]);
  return t;
}
// argmax with lowest-index tie-break (torch.argmax on a 1-row tensor)
__device__ __noinline__ int block_argmax(const float* sl, int V, float* redf, int* redi, float* out_val) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < V; i += kConsumerThreads) {
    const float x = sl[i];
    if (x > bv || (x == bv && i < bi)) { bv = x; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  cbar_sync();
  if (lane == 0) { redf[warp] = bv; redi[warp] = bi; }
  cbar_sync();
  bv = redf[0]; bi = redi[0];
#pragma unroll
  for (int w = 1; w < kConsumerWarps; ++w) {
    const float ov = redf[w]; const int oi = redi[w];
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  if (out_val) *out_val = bv;
  return bi == 0x7fffffff ? 0 : bi;
}

// k-th largest value (counting multiplicity) by 4-pass radix select on order-preserving keys.
__device__ __noinline__ float kth_largest(const float* sl, int V, int k, unsigned* hist, int* redi) {
  uint32_t prefix = 0;
  int krem = k;
  const int lane = threadIdx.x & 31;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    cbar_sync();
    if (threadIdx.x < 256) hist[threadIdx.x] = 0;
    cbar_sync();
    for (int i = threadIdx.x; i < V; i += kConsumerThreads) {
      const uint32_t key = f2key(sl[i]);
      const bool match = (pass == 0) || ((key >> (shift + 8)) == (prefix >> (shift + 8)));
      if (match) atomicAdd(&hist[(key >> shift) & 0xffu], 1u);
    }
    cbar_sync();
    if (threadIdx.x < 32) {
      // lane l covers bins [255-8l-7, 255-8l], scanned from the top
      unsigned s = 0;
#pragma unroll
      for (int t = 0; t < 8; ++t) s += hist[255 - 8 * lane - t];
      unsigned incl = s;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
      }
      const unsigned ballot = __ballot_sync(0xffffffffu, incl >= (unsigned)krem);
      const int sel = __ffs(ballot) - 1;
      if (lane == sel) {
        unsigned above = incl - s;
        int digit = 255 - 8 * lane;
        for (int t = 0; t < 8; ++t) {
          const unsigned c = hist[255 - 8 * lane - t];
          if (above + c >= (unsigned)krem) { digit = 255 - 8 * lane - t; break; }
          above += c;
        }
        redi[0] = digit;
        redi[1] = krem - (int)above;
      }
    }
    cbar_sync();
    prefix |= ((uint32_t)redi[0]) << shift;
    krem = redi[1];
  }
  cbar_sync();
  return key2f(prefix);
}

struct SampleArgs {
  int V;
  int do_sample, top_k;
  float top_p, temperature, rep_pen;
  const uint8_t* seen;  // may be null
  int suppress_start;   // ids >= this (except eos) are masked; V => none
  int eos;
  int suppress_eos;
  int round_bf16;       // logits came from a bf16 tensor: round every intermediate like the reference's bf16 ops
  unsigned long long seed, draw;
};

// logits arrive either as LL words (ll != null, epoch ep) or as a plain fp32 array.
__device__ __noinline__ int sample_row(const float* logits_g, const LLWord* ll, uint32_t ep, const LaunchParams* lp, int pidx,
                          const SampleArgs& a, const SampleScratch& sc) {
  float* sl = sc.sl;
  const int V = a.V;
  // 1. load + repetition penalty (sampling.py:22-29, before suppression) + suppression + temperature
  constexpr int kLg = (kMaxVocab + kConsumerThreads - 1) / kConsumerThreads;  // 11
  LLWord lw[kLg];
  if (ll) {
#pragma unroll
    for (int j = 0; j < kLg; ++j) {
      const int i = threadIdx.x + j * kConsumerThreads;
      if (i < V) lw[j] = ll_ld(ll + i);
    }
  }
#pragma unroll
  for (int j = 0; j < kLg; ++j) {
    const int i = threadIdx.x + j * kConsumerThreads;
    if (i >= V) continue;
    float x = ll ? ll_check(lw[j], ll + i, ep, *lp, pidx) : __ldcg(logits_g + i);
    if (a.seen && a.rep_pen != 1.0f && __ldcg(a.seen + i)) {
      x = (x > 0.f) ? x / a.rep_pen : x * a.rep_pen;
      if (a.round_bf16) x = bf16r(x);
    }
    if (i >= a.suppress_start && i != a.eos) x = -INFINITY;
    if (a.suppress_eos && i == a.eos) x = -INFINITY;
    if (a.do_sample) {
      x = x / a.temperature;
      if (a.round_bf16) x = bf16r(x);
    }
    sl[i] = x;
  }
  cbar_sync();
  float vmax;
  const int imax = block_argmax(sl, V, sc.redf, sc.redi, &vmax);
  if (!a.do_sample) return imax;
  // 2. top-k: keep x >= k-th largest (ties survive, sampling.py:54-56)
  if (a.top_k > 0 && a.top_k < V) {
    const float thr = kth_largest(sl, V, a.top_k, sc.hist, sc.redi);
    for (int i = threadIdx.x; i < V; i += kConsumerThreads)
      if (sl[i] < thr) sl[i] = -INFINITY;
    cbar_sync();
  }
  // 3. top-p on the sorted distribution (sampling.py:57-65): rank j is dropped iff cum_j > top_p and j > 0
  if (a.top_p < 1.0f) {
    float loc = 0.f;
    for (int i = threadIdx.x; i < V; i += kConsumerThreads) loc += expf(sl[i] - vmax);
    const float total = block_sum(loc, sc.redf);
    const float P = a.top_p * total;
    uint32_t lo = 0u, hi = 0xffffffffu;  // smallest key t with mass(key >= t) <= P
    for (int it = 0; it < 32; ++it) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      float msum = 0.f;
      for (int i = threadIdx.x; i < V; i += kConsumerThreads) {
        const float x = sl[i];
        if (f2key(x) >= mid) msum += expf(x - vmax);
      }
      msum = block_sum(msum, sc.redf);
      if (msum <= P) hi = mid; else lo = mid + 1;
    }
    const uint32_t tkey = hi;
    // boundary value v* = largest key below tkey; its first j ties (index order) may still fit under P
    uint32_t bk = 0;
    for (int i = threadIdx.x; i < V; i += kConsumerThreads) {
      const uint32_t k = f2key(sl[i]);
      if (k < tkey && k > bk) bk = k;
    }
    const float bval = block_max(key2f(bk), sc.redf);
    float above = 0.f;
    for (int i = threadIdx.x; i < V; i += kConsumerThreads)
      if (f2key(sl[i]) >= tkey) above += expf(sl[i] - vmax);
    above = block_sum(above, sc.redf);
    int keep_ties = 0;
    if (bval > -INFINITY) {
      const float pe = expf(bval - vmax);
      keep_ties = (int)floorf((P - above) / pe);
      if (keep_ties < 0) keep_ties = 0;
    }
    const int seg = (V + kConsumerThreads - 1) / kConsumerThreads;
    const int i0 = threadIdx.x * seg, i1 = min(V, i0 + seg);
    int cnt = 0;
    for (int i = i0; i < i1; ++i) cnt += (sl[i] == bval) ? 1 : 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += n;
    }
    cbar_sync();
    if (lane == 31) sc.redi[warp] = incl;
    cbar_sync();
    int base = 0;
    for (int w = 0; w < warp; ++w) base += sc.redi[w];
    int rank = base + incl - cnt;
    for (int i = i0; i < i1; ++i) {
      const float x = sl[i];
      if (f2key(x) >= tkey) continue;
      if (x == bval && bval > -INFINITY) {
        if (rank < keep_ties) { ++rank; continue; }
        ++rank;
      }
      if (i != imax) sl[i] = -INFINITY;
    }
    cbar_sync();
  }
  // 4. multinomial(softmax) by inverse CDF on one Philox uniform
  const int seg = (V + kConsumerThreads - 1) / kConsumerThreads;
  const int i0 = threadIdx.x * seg, i1 = min(V, i0 + seg);
  float loc = 0.f;
  for (int i = i0; i < i1; ++i) loc += expf(sl[i] - vmax);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float incl = loc;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += n;
  }
  cbar_sync();
  if (lane == 31) sc.redf[warp] = incl;
  if (threadIdx.x == 0) sc.redi[16] = 0x7fffffff;
  cbar_sync();
  float base = 0.f, total = 0.f;
  for (int w = 0; w < kConsumerWarps; ++w) {
    if (w < warp) base += sc.redf[w];
    total += sc.redf[w];
  }
  const uint4 rnd = philox4x32_10(make_uint4((uint32_t)a.draw, (uint32_t)(a.draw >> 32), 0u, 0u),
                                  make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32)));
  const float u = (float)(rnd.x >> 8) * (1.0f / 16777216.0f);
  const float target = u * total;
  const float excl = base + incl - loc;
  if (loc > 0.f && target >= excl && target < excl + loc) {
    float c = excl;
    int pick = -1;
    for (int i = i0; i < i1; ++i) {
      const float e = expf(sl[i] - vmax);
      if (e > 0.f) { pick = i; c += e; if (target < c) break; }
    }
    if (pick >= 0) atomicMin(&sc.redi[16], pick);
  }
  cbar_sync();
  int tok = sc.redi[16];
  if (tok == 0x7fffffff) tok = imax;  // rounding fell off the end of the CDF
  return tok;
}

// Publish one bf16 row (n elements) as LL words with all consumer threads.
__device__ __forceinline__ void publish_row(LLWord* dst, const bf16* src, int n, uint32_t ep) {
  for (int c = threadIdx.x; c < n; c += kConsumerThreads) ll_st(dst + c, __bfloat162float(src[c]), ep);
}

__device__ __noinline__ void sample_phase(const Phase& ph, const LaunchParams& p, unsigned char* smem_base, uint32_t ep, int pidx,
                             const int* frame_done) {
  const Smem sm = carve_smem(smem_base, p);
  const SampleScratch sc = sample_scratch(sm.scratch);
  const int Ht = p.stacks[ST_TALKER].hidden;
  const int ncb = p.n_code_groups - 1;
  const uint32_t ep_in = ep - 1;
  // CTAs without a stream idle through sampling phases: the once-per-step fence that makes this step's KV
  // rows visible to whichever CTA reads them in a later step goes here, off the critical path.
  if ((int)blockIdx.x >= (p.mode == MODE_PREFILL ? 1 : p.n_rows) && threadIdx.x == 0 && ph.kind != SMP_PRED && ph.kind != SMP_PRED_ONLY)
    __threadfence();
  for (int b = blockIdx.x; b < (p.mode == MODE_PREFILL ? 1 : p.n_rows); b += gridDim.x) {
    const int slot = p.stream0 + b;
    StreamState* st = p.st + slot;
    const int done = (p.mode == MODE_FRAMES) ? frame_done[b] : 0;
    LLWord* pin = reinterpret_cast<LLWord*>(p.bufs[p.has_s2m ? BUF_PIN : BUF_PX]);
    const int ldpin = p.ld[p.has_s2m ? BUF_PIN : BUF_PX];
    const LLWord* lgbuf = reinterpret_cast<const LLWord*>(p.bufs[BUF_LOGITS]);
    if (ph.kind == SMP_PRED || ph.kind == SMP_PRED_ONLY) {
      const int i = ph.aux;  // codebook step 0..ncb-1
      const int Vp = p.stacks[ST_PRED].vocab;
      const int lrow = (ph.flags & F_ROWS2) ? 2 * b + 1 : b;
      const LLWord* lg = lgbuf + (size_t)lrow * p.ld[BUF_LOGITS];
      if (p.pred_logits_all) {
        for (int v = threadIdx.x; v < Vp; v += kConsumerThreads)
          p.pred_logits_all[(size_t)i * Vp + v] = ll_wait(lg + v, ep_in, p, pidx);
      }
      SampleArgs a;
      a.V = Vp; a.do_sample = p.sub.do_sample; a.top_k = p.sub.top_k; a.top_p = p.sub.top_p;
      a.temperature = p.sub.temperature; a.rep_pen = 1.0f; a.seen = nullptr; a.suppress_start = Vp; a.eos = -1;
      a.suppress_eos = 0; a.round_bf16 = 1; a.seed = p.pol.seed ^ (0x9E3779B97F4A7C15ull * (unsigned long long)(slot + 1));
      a.draw = __ldcg(&st->draws) + (unsigned long long)(i + 1);
      const int tok = sample_row(nullptr, lg, ep_in, &p, pidx, a, sc);
      if (threadIdx.x == 0) st->cur_codes[i + 1] = tok;
      if (i + 1 < ncb) {
        // next predictor input row: codec_embeds[i](tok)  (predictor_graph.py:144)
        publish_row(pin + (size_t)b * ldpin, p.pred_embeds[i] + (size_t)tok * Ht, Ht, ep);
      } else if (ph.kind == SMP_PRED) {
        // frame complete: append [c0..c15], then build the next talker input (generate.py:159-171)
        const int nfr = __ldcg(&st->n_frames);
        const int c0 = __ldcg(&st->token);
        if (!done) {
          if (threadIdx.x < p.n_code_groups && nfr < p.max_frames) {
            const int code = (threadIdx.x == 0) ? c0 : (threadIdx.x == ncb ? tok : __ldcg(&st->cur_codes[threadIdx.x]));
            st->codes[(size_t)nfr * p.n_code_groups + threadIdx.x] = code;
          }
          if (threadIdx.x == 0) {
            st->n_frames = nfr + 1;
            st->seen[c0] = 1;
            st->draws = __ldcg(&st->draws) + 32ull;
          }
        }
        cbar_sync();
        const int gs = __ldcg(&st->gen_step);
        const int ntr = __ldcg(&st->n_trailing);
        const bf16* text = (gs < ntr) ? st->trailing + (size_t)gs * Ht : st->pad_embed;
        LLWord* tx = reinterpret_cast<LLWord*>(p.bufs[BUF_TX]) + (size_t)b * p.ld[BUF_TX];
        for (int c = threadIdx.x; c < Ht; c += kConsumerThreads) {
          float s = __bfloat162float(p.codec_embed[(size_t)c0 * Ht + c]);
          for (int g = 0; g < ncb; ++g) {
            const int code = (g == ncb - 1) ? tok : __ldcg(&st->cur_codes[g + 1]);
            s += __bfloat162float(p.pred_embeds[g][(size_t)code * Ht + c]);
          }
          const float e = bf16r(s);
          ll_st(tx + c, bf16r(e + __bfloat162float(text[c])), ep);
        }
        // static-cache bound (generate.py:174-177): the frame stays, decoding stops after this step
        if (threadIdx.x == 0 && !done) {
          const int pos = __ldcg(&st->position);
          if (pos >= p.stacks[ST_TALKER].max_pos - 1) st->done = 2;
        }
      }
    } else {
      // SMP_TALKER / SMP_PREFILL: first-codebook sampler (generate.py:124-134, :182-197)
      const int Vt = p.stacks[ST_TALKER].vocab;
      const LLWord* lg = lgbuf + (size_t)b * p.ld[BUF_LOGITS];
      const int nfr = __ldcg(&st->n_frames);
      const int done_now = (ph.kind == SMP_PREFILL) ? 0 : __ldcg(&st->done);  // includes this frame's cache-bound stop
      SampleArgs a;
      a.V = Vt; a.do_sample = p.pol.do_sample; a.top_k = p.pol.top_k; a.top_p = p.pol.top_p;
      a.temperature = p.pol.temperature; a.rep_pen = p.pol.rep_pen;
      a.seen = (ph.kind == SMP_TALKER && nfr > 0) ? st->seen : nullptr;
      a.suppress_start = max(0, Vt - p.pol.suppress_tail); a.eos = p.eos_id;
      a.suppress_eos = (ph.kind == SMP_PREFILL) ? (p.pol.min_new_tokens > 0) : (nfr < p.pol.min_new_tokens);
      a.round_bf16 = 1;
      a.seed = p.pol.seed ^ (0x9E3779B97F4A7C15ull * (unsigned long long)(slot + 1));
      a.draw = __ldcg(&st->draws);
      int tok = sample_row(nullptr, lg, ep_in, &p, pidx, a, sc);
      const bool live = !done_now;
      int new_done = done_now, new_pos = __ldcg(&st->position), new_gs = __ldcg(&st->gen_step);
      if (live) {
        if (ph.kind == SMP_TALKER) { new_pos += 1; new_gs += 1; }
        if (tok == p.eos_id) new_done = 1;  // generate.py:150 — checked before the next frame is built
      } else {
        tok = __ldcg(&st->token);
      }
      cbar_sync();  // everyone has read the old state
      if (threadIdx.x == 0) {
        if (live) { st->token = tok; st->draws = a.draw + 1ull; st->position = new_pos; st->gen_step = new_gs; st->done = new_done; }
        // control record every CTA reads at the top of the next frame
        ll_st(&st->ctl[0], __int_as_float(new_done), ep);
        ll_st(&st->ctl[1], __int_as_float(new_pos), ep);
      }
      // predictor pass-0 input rows [past_hidden ; codec_embed(token)] (generate.py:154-155), addressed by
      // stream slot; always re-published so every reader sees this phase's epoch
      const bf16* hid = reinterpret_cast<const bf16*>(p.bufs[BUF_HID]) + (size_t)b * p.ld[BUF_HID];
      publish_row(pin + (size_t)(2 * slot) * ldpin, hid, Ht, ep);
      publish_row(pin + (size_t)(2 * slot + 1) * ldpin, p.codec_embed + (size_t)tok * Ht, Ht, ep);
      if (threadIdx.x == 0) __threadfence();
    }
    cbar_sync();
  }
}

// =================================================================================================
// Kernel
// =================================================================================================
__global__ void __launch_bounds__(kThreads, 1) fq3_stream_kernel(const __grid_constant__ LaunchParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const Smem sm = carve_smem(smem_raw, p);
  const int tid = threadIdx.x;

  if (tid == 0) {
    for (int s = 0; s < kMaxStages; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], kConsumerWarps);
    }
    sm.ctl[0] = 0;
    fence_barrier_init();
  }
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.prog);
    uint4* dst = reinterpret_cast<uint4*>(sm.prog);
    for (int i = tid; i < p.n_phases * 2; i += kThreads) dst[i] = __ldg(src + i);
  }
  __syncthreads();

  if (tid >= kConsumerThreads) {
    // ------------------------------ producer warp ------------------------------
    if (tid == kConsumerThreads) {
      const uint64_t pol_stream = policy_evict_first(), pol_keep = policy_evict_last();
      unsigned tile_it = 0;
      for (int iter = 0; iter < p.n_iters; ++iter) {
        if (iter > 0) {
          Spin s;
          int go;
          while ((go = ld_volatile_shared_i32(&sm.ctl[0])) >= 0 && go < iter) {
            __nanosleep(32);
            s.tick(p, DE_HANDSHAKE, -1, iter);
          }
          if (go < 0) break;
        }
        for (int i = 0; i < p.n_phases; ++i) {
          const Phase& ph = sm.prog[i];
          if (ph.type == PH_GEMV) gemv_phase_produce(ph, p, smem_raw, tile_it, i, pol_stream, pol_keep);
        }
      }
    }
    return;
  }

  // -------------------------------- consumer warps --------------------------------
  unsigned tile_it = 0;
  int* frame_pos = sm.ctl + 8;    // [4] talker positions of this iteration
  int* frame_done = sm.ctl + 12;  // [4]
  for (int iter = 0; iter < p.n_iters; ++iter) {
    const uint32_t ep0 = p.epoch_base + (uint32_t)iter * (uint32_t)p.n_phases;
    if (p.mode == MODE_FRAMES || p.mode == MODE_TALKER_STEP) {
      // per-stream frame state: from the stream state at launch, afterwards from the sampler's control record
      if (tid < p.n_rows) {
        const StreamState* st = p.st + p.stream0 + tid;
        int done, pos;
        if (iter == 0) {
          done = __ldcg(&st->done);
          pos = __ldcg(&st->position);
        } else {
          done = __float_as_int(ll_wait(&st->ctl[0], ep0, p, -2));
          pos = __float_as_int(ll_wait(&st->ctl[1], ep0, p, -2));
        }
        frame_pos[tid] = pos;
        frame_done[tid] = done;
      }
      cbar_sync();
      if (p.mode == MODE_FRAMES) {
        int all_done = 1;
        for (int b = 0; b < p.n_rows; ++b) all_done &= (frame_done[b] != 0);
        if (iter > 0 && tid == 0) st_volatile_shared_i32(&sm.ctl[0], all_done ? -1 : iter);
        if (all_done && iter > 0) break;
      }
    }
    for (int i = 0; i < p.n_phases; ++i) {
      const Phase& ph = sm.prog[i];
      const uint32_t ep = ep0 + (uint32_t)i + 1u;
      switch (ph.type) {
        case PH_GEMV: gemv_phase_consume(ph, p, smem_raw, tile_it, i, ep); break;
        case PH_ATTN: if (!(p.debug & 4)) attn_phase(ph, p, smem_raw, ep, i, frame_pos); break;
        case PH_SAMPLE: sample_phase(ph, p, smem_raw, ep, i, frame_done); break;
        default: if (tid == 0) device_fault(p, DE_BAD_PHASE, i, ph.type); break;
      }
    }
  }
}

// Small kernels -----------------------------------------------------------------------------------
// Stand-alone sampler behind fq3_sample (sampling.py parity tests and the operator-at-a-time host loop).
__global__ void __launch_bounds__(kConsumerThreads, 1)
fq3_sample_kernel(const float* logits, SampleArgs a, const long long* history, int n_history, uint8_t* seen_scratch,
                  long long* out) {
  __shared__ __align__(16) unsigned char scratch[kScratchBytes];
  const SampleScratch sc = sample_scratch(scratch);
  if (history && n_history > 0 && a.rep_pen != 1.0f) {
    for (int i = threadIdx.x; i < a.V; i += kConsumerThreads) seen_scratch[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n_history; i += kConsumerThreads) {
      const long long t = history[i];
      if (t >= 0 && t < a.V) seen_scratch[t] = 1;
    }
    __syncthreads();
    a.seen = seen_scratch;
  } else {
    a.seen = nullptr;
  }
  const int tok = sample_row(logits, nullptr, 0u, nullptr, 0, a, sc);
  if (threadIdx.x == 0) out[0] = tok;
}

// apply_repetition_penalty (sampling.py:10-29) in place on fp32 logits; history marks a bitmap first.
__global__ void __launch_bounds__(kConsumerThreads, 1)
fq3_rep_penalty_kernel(float* logits, int V, const long long* history, int n_history, float penalty, int round_bf16,
                       uint8_t* seen) {
  for (int i = threadIdx.x; i < V; i += kConsumerThreads) seen[i] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < n_history; i += kConsumerThreads) {
    const long long t = history[i];
    if (t >= 0 && t < V) seen[t] = 1;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < V; i += kConsumerThreads) {
    if (__ldcg(seen + i)) {
      float x = logits[i];
      x = (x > 0.f) ? x / penalty : x * penalty;
      logits[i] = round_bf16 ? bf16r(x) : x;
    }
  }
}

__global__ void fq3_reset_stream_kernel(StreamState* st, int V) {
  for (int i = threadIdx.x; i < V; i += blockDim.x) st->seen[i] = 0;
  if (threadIdx.x == 0) {
    st->token = 0; st->position = 0; st->gen_step = 0; st->n_frames = 0; st->done = 0; st->n_pad = 0;
    st->rope_delta = 0; st->draws = 0ull;  // text conditioning (n_trailing) is set separately and survives a reset
    for (int i = 0; i < 32; ++i) st->cur_codes[i] = 0;
  }
}

__global__ void fq3_set_state_kernel(StreamState* st, int token, int position, int gen_step, int set_mask, int n_pad,
                                     int rope_delta, int n_trailing) {
  if (threadIdx.x == 0) {
    if (set_mask & 1) st->token = token;
    if (set_mask & 2) { st->n_pad = n_pad; st->rope_delta = rope_delta; }
    if (set_mask & 4) st->n_trailing = n_trailing;
    if (set_mask & 8) { st->position = position; st->gen_step = gen_step; }
    if (set_mask & 16) st->done = 0;
  }
}

// K/V import: src [nkv, T, 128] -> cache [nkv, max_pos, 128] of one (layer, slot)
__global__ void fq3_import_kv_kernel(bf16* kdst, bf16* vdst, const bf16* k, const bf16* v, int nkv, int T, int max_pos) {
  const size_t n = (size_t)nkv * T * (kHeadDim / 8);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % (kHeadDim / 8));
    const size_t t = (i / (kHeadDim / 8)) % T;
    const size_t h = i / ((size_t)(kHeadDim / 8) * T);
    const size_t d = (h * max_pos + t) * (kHeadDim / 8) + c;
    reinterpret_cast<uint4*>(kdst)[d] = reinterpret_cast<const uint4*>(k)[i];
    reinterpret_cast<uint4*>(vdst)[d] = reinterpret_cast<const uint4*>(v)[i];
  }
}

__global__ void fq3_codes_to_i64_kernel(const int* cur_codes, int n, long long* out) {
  if (threadIdx.x < n) out[threadIdx.x] = cur_codes[threadIdx.x + 1];
}

// Plain bf16 rows -> LL words (epoch 0: "written before launch") and back.
__global__ void fq3_pack_ll_kernel(LLWord* dst, int ld_dst, const bf16* src, int ld_src, int rows, int cols) {
  const size_t n = (size_t)rows * cols;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / cols, c = i - r * cols;
    dst[r * ld_dst + c] = make_uint2(__float_as_uint(__bfloat162float(src[r * ld_src + c])), 0u);
  }
}
__global__ void fq3_unpack_ll_bf16_kernel(bf16* dst, int ld_dst, const LLWord* src, int ld_src, int rows, int cols) {
  const size_t n = (size_t)rows * cols;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / cols, c = i - r * cols;
    dst[r * ld_dst + c] = __float2bfloat16_rn(__uint_as_float(src[r * ld_src + c].x));
  }
}
__global__ void fq3_unpack_ll_f32_kernel(float* dst, int ld_dst, const LLWord* src, int ld_src, int rows, int cols) {
  const size_t n = (size_t)rows * cols;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / cols, c = i - r * cols;
    dst[r * ld_dst + c] = __uint_as_float(src[r * ld_src + c].x);
  }
}

}  // namespace fq3
