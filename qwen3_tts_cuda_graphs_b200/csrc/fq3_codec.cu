// fq3_codec.cu — kernels of the 12 Hz speech-tokenizer decoder (include/fq3_codec.h).
//
// Activations are channels-last bf16 [time, channels]; every convolution is an implicit GEMM over (tap, c_in)
// whose A rows are time-shifted views of the same buffer (no im2col), so causal padding is a bounds check.
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched with cudaGetDriverEntryPoint, nothing links against libcuda)
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/fq3_codec.h"

typedef __nv_bfloat16 bf16;

namespace {

thread_local std::string g_err;
// launches of the whole process (several scheduler threads may run op lists at once: one per GPU) and of the calling thread alone
std::atomic<int64_t> g_launches{0};
thread_local int64_t t_launches = 0;
inline void count_launches(int64_t n) { g_launches += n; t_launches += n; }

__device__ __forceinline__ float bf16r(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int BM = 64, BN = 64, BK = 32, PAD = 8, LDS = BK + PAD;

__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float silu_f(float x) { return x / (1.f + expf(-x)); }

// Epilogue for one (row, even col) pair of accumulators; order documented in include/fq3_codec.h / DESIGN.md §5.
__device__ __forceinline__ void epilogue_pair(const fq3c_op& o, int row, int col, float v0, float v1) {
  if (row >= o.M || col >= o.N) return;
  const bool has1 = (col + 1 < o.N);
  const int cm0 = col % o.col_mod, cm1 = (col + 1) % o.col_mod;
  if (o.flags & FQ3C_BIAS) {
    const float* b = reinterpret_cast<const float*>(o.bias);
    v0 += b[cm0];
    if (has1) v1 += b[cm1];
  }
  v0 = bf16r(v0);
  v1 = bf16r(v1);
  if (o.flags & FQ3C_SWIGLU) {  // interleaved (gate, up) columns -> one output column
    const float y = bf16r(bf16r(silu_f(v0)) * v1);
    reinterpret_cast<bf16*>(o.C)[(size_t)row * o.ldc + (col >> 1)] = __float2bfloat16_rn(y);
    return;
  }
  if (o.flags & FQ3C_GELU) { v0 = bf16r(gelu_f(v0)); v1 = bf16r(gelu_f(v1)); }
  if (o.flags & FQ3C_SILU) { v0 = bf16r(silu_f(v0)); v1 = bf16r(silu_f(v1)); }
  if (o.flags & FQ3C_SCALE) {
    const float* s = reinterpret_cast<const float*>(o.scale);
    v0 = bf16r(v0 * s[cm0]);
    if (has1) v1 = bf16r(v1 * s[cm1]);
  }
  if (o.flags & FQ3C_RESID) {
    const bf16* r = reinterpret_cast<const bf16*>(o.res) + (size_t)row * o.ldr + col;
    v0 = bf16r(v0 + __bfloat162float(r[0]));
    if (has1) v1 = bf16r(v1 + __bfloat162float(r[1]));
  }
  if (o.flags & FQ3C_CLAMP) { v0 = fminf(fmaxf(v0, -1.f), 1.f); v1 = fminf(fmaxf(v1, -1.f), 1.f); }
  if (o.flags & FQ3C_OUT_F32) {
    float* c = reinterpret_cast<float*>(o.C) + (size_t)row * o.ldc + col;
    c[0] = v0;
    if (has1) c[1] = v1;
  } else {
    bf16* c = reinterpret_cast<bf16*>(o.C) + (size_t)row * o.ldc + col;
    c[0] = __float2bfloat16_rn(v0);
    if (has1) c[1] = __float2bfloat16_rn(v1);
  }
  if (o.flags & FQ3C_SNAKE2) {
    const float* ea = reinterpret_cast<const float*>(o.p0);
    const float* ib = reinterpret_cast<const float*>(o.p1);
    bf16* c2 = reinterpret_cast<bf16*>(o.C2) + (size_t)row * o.ldc + col;
    float s0 = sinf(v0 * ea[cm0]);
    c2[0] = __float2bfloat16_rn(v0 + ib[cm0] * s0 * s0);
    if (has1) {
      float s1 = sinf(v1 * ea[cm1]);
      c2[1] = __float2bfloat16_rn(v1 + ib[cm1] * s1 * s1);
    }
  }
}

// Epilogue for eight consecutive columns of one row (the tcgen05 path: a thread owns a row of the accumulator, so it
// must move whole 16-byte pieces or every store is a partial sector).  Same arithmetic and order as epilogue_pair.
__device__ __forceinline__ void epilogue_row8(const fq3c_op& o, int row, int col, const float (&acc)[8]) {
  const bool vec = (col + 8 <= o.N) && !(o.flags & (FQ3C_SWIGLU | FQ3C_OUT_F32)) && ((o.ldc & 7) == 0) &&
                   (!(o.flags & FQ3C_RESID) || (o.ldr & 7) == 0);
  if (!vec) {
#pragma unroll
    for (int j = 0; j < 4; ++j) epilogue_pair(o, row, col + 2 * j, acc[2 * j], acc[2 * j + 1]);
    return;
  }
  if (row >= o.M) return;
  float v[8];
  int cm[8];
  {
    int c0 = col % o.col_mod;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      cm[j] = c0;
      if (++c0 == o.col_mod) c0 = 0;
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = acc[j];
  if (o.flags & FQ3C_BIAS) {
    const float* b = reinterpret_cast<const float*>(o.bias);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += b[cm[j]];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = bf16r(v[j]);
  if (o.flags & FQ3C_GELU) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = bf16r(gelu_f(v[j]));
  }
  if (o.flags & FQ3C_SILU) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = bf16r(silu_f(v[j]));
  }
  if (o.flags & FQ3C_SCALE) {
    const float* sc = reinterpret_cast<const float*>(o.scale);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = bf16r(v[j] * sc[cm[j]]);
  }
  if (o.flags & FQ3C_RESID) {
    const uint4 rr = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(o.res) + (size_t)row * o.ldr + col);
    const uint32_t rw[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[2 * j] = bf16r(v[2 * j] + __uint_as_float(rw[j] << 16));
      v[2 * j + 1] = bf16r(v[2 * j + 1] + __uint_as_float(rw[j] & 0xffff0000u));
    }
  }
  if (o.flags & FQ3C_CLAMP) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fminf(fmaxf(v[j], -1.f), 1.f);
  }
  {
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
      w[j] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(o.C) + (size_t)row * o.ldc + col) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  if (o.flags & FQ3C_SNAKE2) {
    const float* ea = reinterpret_cast<const float*>(o.p0);
    const float* ib = reinterpret_cast<const float*>(o.p1);
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float s0 = sinf(v[2 * j] * ea[cm[2 * j]]), s1 = sinf(v[2 * j + 1] * ea[cm[2 * j + 1]]);
      const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j] + ib[cm[2 * j]] * s0 * s0, v[2 * j + 1] + ib[cm[2 * j + 1]] * s1 * s1);
      w[j] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(o.C2) + (size_t)row * o.ldc + col) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// Implicit-GEMM convolution / linear layer on the bf16 tensor cores (mma.sync m16n8k16, fp32 accumulate).
__global__ void __launch_bounds__(128) fq3c_gemm_kernel(const fq3c_op o) {
  __shared__ __align__(16) bf16 As[2][BM][LDS];
  __shared__ __align__(16) bf16 Bs[2][BN][LDS];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp >> 1, wn = warp & 1;
  const int m0 = o.m_begin + blockIdx.y * BM, n0 = blockIdx.x * BN;
  const bf16* A = reinterpret_cast<const bf16*>(o.A);
  const bf16* B = reinterpret_cast<const bf16*>(o.B);
  const int KT = (o.K + BK - 1) / BK;

  auto load_stage = [&](int st, int kt) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int ch = tid + i * 128;  // 0..255
      const int r = ch >> 2, c = ch & 3;
      const int k0 = kt * BK + c * 8;
      // A: time-shifted row of the activation buffer
      {
        const int m = m0 + r;
        const int tap = k0 / o.cin, ci = k0 - tap * o.cin;
        const int srow = m + (tap < 8 ? o.tap_off[tap] : 0);
        const bool ok = (m < o.M) && (k0 < o.K) && (srow >= 0) && (srow < o.a_rows);
        const bf16* src = ok ? (A + (size_t)srow * o.lda + ci) : A;
        cp_async16(&As[st][r][c * 8], src, ok);
      }
      {
        const int n = n0 + r;
        const bool ok = (n < o.N) && (k0 < o.K);
        const bf16* src = ok ? (B + (size_t)n * o.K + k0) : B;
        cp_async16(&Bs[st][r][c * 8], src, ok);
      }
    }
    cp_async_commit();
  };

  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;

  load_stage(0, 0);
  for (int kt = 0; kt < KT; ++kt) {
    const int st = kt & 1;
    if (kt + 1 < KT) { load_stage(st ^ 1, kt + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < BK; ks += 16) {
      uint32_t af[2][4], bfr[4][2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int row = wm * 32 + i * 16 + (lane & 15);
        const int col = ks + (lane >> 4) * 8;
        ldmatrix_x4(af[i], &As[st][row][col]);
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {  // each x4 covers two n8 tiles
        uint32_t r[4];
        const int row = wn * 32 + j * 16 + (lane & 7) + ((lane >> 4) << 3);
        const int col = ks + ((lane >> 3) & 1) * 8;
        ldmatrix_x4(r, &Bs[st][row][col]);
        bfr[2 * j][0] = r[0]; bfr[2 * j][1] = r[1]; bfr[2 * j + 1][0] = r[2]; bfr[2 * j + 1][1] = r[3];
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) mma16816(acc[i][j], af[i], bfr[j][0], bfr[j][1]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int row = m0 + wm * 32 + i * 16 + (lane >> 2);
      const int col = n0 + wn * 32 + j * 8 + (lane & 3) * 2;
      epilogue_pair(o, row, col, acc[i][j][0], acc[i][j][1]);
      epilogue_pair(o, row + 8, col, acc[i][j][2], acc[i][j][3]);
    }
}

// =================================================================================================
// tcgen05 path of the implicit GEMM (sm_100a): 128 x BN x 64 tiles, accumulator in tensor memory.
//
// A tile is K-major with the 128-byte swizzle the tensor core expects (the layout a TMA box with SWIZZLE_128B writes):
// row r of the tile is one 128-byte line (64 bf16), its 16-byte chunk c lives at r * 128 + ((c ^ (r & 7)) << 4).  The
// tiles are filled with cp.async so that the time-shifted (tap) rows of the implicit GEMM and the zero padding stay a
// per-chunk address computation; eight lanes cover one 128-byte line of a row, so global reads are coalesced and the
// swizzle makes the shared-memory writes conflict-free.  One thread issues tcgen05.mma (4 x k16 per stage) and commits to
// the stage's mbarrier; the epilogue reads the accumulator with tcgen05.ld (one TMEM lane = one output row per thread).
// =================================================================================================
constexpr int TM = 128, TK = 64, TSTAGES = 6;
constexpr int TLOADERS = 256;                 // warps 0-3 fill A, warps 4-7 fill B; warp 8 issues the MMAs
constexpr int TTHREADS = TLOADERS + 32;
constexpr int TMAXN = 128;
constexpr int T_A_BYTES = TM * TK * 2;        // 16 KB
constexpr int T_B_BYTES = TMAXN * TK * 2;     // 16 KB
constexpr int t_smem(int stages) { return stages * (T_A_BYTES + T_B_BYTES) + 1024 /*alignment*/ + 256 /*barriers + tmem pointer*/; }
constexpr int T_SMEM = t_smem(TSTAGES);

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(20);
}
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  // K-major, SWIZZLE_128B: start address >> 4 | LBO (unused for swizzled K-major) = 1 | SBO = 1024 B (8 rows) | version 1 | layout 2
  return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}

// Warp-specialised: the loader warps run up to TSTAGES k-blocks ahead of the tensor core (full / empty mbarriers per stage,
// cp.async completion counted straight into the full barrier), one thread issues the MMAs, all eight loader warps drain
// the accumulator.  No block-wide barrier inside the k loop.
// Split-K (splits > 1, short-and-wide problems that would leave most SMs idle): blockIdx.z takes a contiguous range of
// k-blocks and stores its raw fp32 tile into the workspace [split][M][Nw]; fq3c_splitk_reduce_kernel adds the splits in
// order and applies the fused epilogue.
// TMA = true: the operand tiles arrive by cp.async.bulk.tensor (SASS UTMALDG) — one elected thread asks for the A box (128 rows x 64
// columns at row m0 + tap offset: rows outside the tensor, i.e. the causal padding, come back as zeros) and the B box (BN x 64) of a
// k-block and both count their bytes on the stage's `full` mbarrier; SWIZZLE_128B writes exactly the layout described above.  Used
// whenever a k-block cannot straddle two taps (one tap, or c_in a multiple of 64); the cp.async loaders stay for the rest.
template <int MINB, bool TMA>  // MINB: CTAs per SM the register budget is planned for: 2 with three stages, 1 with six
__global__ void __launch_bounds__(TTHREADS, MINB) fq3c_gemm_tc5_kernel(const fq3c_op o, const int BN, const int splits, const int Nw, const int NST,
                                                                        const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB) {
  extern __shared__ unsigned char tsmem_raw[];
  const uint32_t raw = smem_u32(tsmem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // swizzle atoms are 1024-byte aligned
  unsigned char* gbase = tsmem_raw + (base - raw);
  const uint32_t a_smem = base, b_smem = base + NST * T_A_BYTES;
  const uint32_t bars = base + NST * (T_A_BYTES + T_B_BYTES);  // full[NST] | empty[NST] | done | tmem base address
  const uint32_t full = bars, empty = bars + 8u * NST, done = bars + 16u * NST;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + NST * (T_A_BYTES + T_B_BYTES) + 16 * NST + 16);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = o.m_begin + blockIdx.y * TM, n0 = blockIdx.x * BN;
  const bf16* A = reinterpret_cast<const bf16*>(o.A);
  const bf16* B = reinterpret_cast<const bf16*>(o.B);
  const int KT_all = (o.K + TK - 1) / TK;
  const int kt_begin = (int)(((long)blockIdx.z * KT_all) / splits), kt_end = (int)(((long)(blockIdx.z + 1) * KT_all) / splits);
  const int KT = kt_end - kt_begin;  // k-blocks of this CTA; `it` below counts them from 0

  if (tid == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(full + 8u * s, TMA ? 1 : TLOADERS);
      mbar_init(empty + 8u * s, 1);
    }
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMAXN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp == 8) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      // instruction descriptor: D = F32, A = B = BF16, both K-major, N = BN, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
      for (int kt = 0; kt < KT; ++kt) {
        const int st = kt % NST;
        mbar_wait(full + 8u * st, (uint32_t)((kt / NST) & 1));
        if (!TMA) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // cp.async (generic proxy) writes -> tensor core (async proxy) reads
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a0 = a_smem + (uint32_t)st * T_A_BYTES, b0 = b_smem + (uint32_t)st * T_B_BYTES;
#pragma unroll
        for (int kk = 0; kk < TK / 16; ++kk) {
          const uint64_t da = umma_desc_sw128(a0 + kk * 32), db = umma_desc_sw128(b0 + kk * 32);
          const uint32_t acc = (kt > 0 || kk > 0) ? 1u : 0u;
          asm volatile(
              "{\n\t.reg .pred p;\n\t"
              "setp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
              ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc)
              : "memory");
        }
        // frees the stage once its MMAs have read it (implies tcgen05.fence::before_thread_sync)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(empty + 8u * st) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(done) : "memory");
    }
  } else {
    // ------------------------------ loaders ------------------------------
    // warps 0-3: A rows r0 + 16 i (i < 8), chunk c; warps 4-7: B rows likewise.  Eight lanes cover one 128-byte line of a row.
    const int lt = tid & 127, c = lt & 7, r0 = lt >> 3;
    const bool is_a = warp < 4;
    if (TMA) {
      if (tid == 0) {
        const uint32_t tx = (uint32_t)(T_A_BYTES + BN * TK * 2);
        for (int kt = 0; kt < KT; ++kt) {
          const int st = kt % NST;
          if (kt >= NST) mbar_wait(empty + 8u * st, (uint32_t)(((kt / NST) - 1) & 1));
          const int k0 = (kt_begin + kt) * TK;
          const int tap = k0 / o.cin, ci = k0 - tap * o.cin;
          const int toff = (tap < 8) ? o.tap_off[tap] : 0;
          const uint32_t fb = full + 8u * st;
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"(tx) : "memory");
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                           a_smem + (uint32_t)st * T_A_BYTES),
                       "l"(&tmA), "r"(ci), "r"(m0 + toff), "r"(fb)
                       : "memory");
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                           b_smem + (uint32_t)st * T_B_BYTES),
                       "l"(&tmB), "r"(k0), "r"(n0), "r"(fb)
                       : "memory");
        }
      }
    } else
    for (int kt = 0; kt < KT; ++kt) {
      const int st = kt % NST;
      if (kt >= NST) {  // one lane per warp polls: 256 spinning threads would swamp the shared-memory pipe the copies need
        if (lane == 0) mbar_wait(empty + 8u * st, (uint32_t)(((kt / NST) - 1) & 1));
        __syncwarp();
      }
      const int k0 = (kt_begin + kt) * TK + c * 8;
      const bool kok = k0 < o.K;
      if (is_a) {
        const int tap = k0 / o.cin, ci = k0 - tap * o.cin;
        const int toff = (tap < 8) ? o.tap_off[tap] : 0;
        const uint32_t sbase = a_smem + (uint32_t)st * T_A_BYTES;
#pragma unroll
        for (int i = 0; i < TM / 16; ++i) {
          const int r = r0 + i * 16;
          const int m = m0 + r;
          const int srow = m + toff;
          const bool ok = kok && (m < o.M) && (srow >= 0) && (srow < o.a_rows);
          const bf16* src = ok ? (A + (size_t)srow * o.lda + ci) : A;
          const uint32_t dst = sbase + (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4);
          const int sz = ok ? 16 : 0;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
        }
      } else {
        const uint32_t sbase = b_smem + (uint32_t)st * T_B_BYTES;
        for (int r = r0; r < BN; r += 16) {
          const int n = n0 + r;
          const bool ok = kok && (n < o.N);
          const bf16* src = ok ? (B + (size_t)n * o.K + k0) : B;
          const uint32_t dst = sbase + (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4);
          const int sz = ok ? 16 : 0;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
        }
      }
      // counted as one arrival on the stage's full barrier when this thread's copies have landed
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(full + 8u * st) : "memory");
    }
    // ------------------------------ epilogue ------------------------------
    if (lane == 0) mbar_wait(done, 0u);
    __syncwarp();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // warp w reads TMEM lanes 32 * (w & 3) .. +31 (= output rows); warps 0-3 take the first half of the columns, 4-7 the second
    const int row = m0 + (warp & 3) * 32 + lane;
    const int chunks = (BN + 31) / 32;  // 32-column chunks in the tile
    const int per = (chunks + 1) / 2;
    const int c_begin = (warp >> 2) * per, c_end = min(chunks, c_begin + per);
    for (int cc = c_begin; cc < c_end; ++cc) {
      uint32_t v[32];
      const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(cc * 32);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
          "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
            "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
            "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
            "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (cc * 32 + 8 * j < BN) {
          float a8[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) a8[q] = __uint_as_float(v[8 * j + q]);
          const int col = n0 + cc * 32 + 8 * j;
          if (splits > 1) {
            if (row < o.M && col < o.N) {
              float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(o.ws) + ((size_t)blockIdx.z * o.M + row) * Nw + col);
              dst[0] = make_float4(a8[0], a8[1], a8[2], a8[3]);
              dst[1] = make_float4(a8[4], a8[5], a8[6], a8[7]);
            }
          } else {
            epilogue_row8(o, row, col, a8);
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMAXN) : "memory");
}

// RVQ dequantiser front half: gather + sum codebook rows. C[m, 0:dim] = sum of the first i0 groups, C[m, dim:2dim] = rest.
// second half of a split-K GEMM: a thread owns eight consecutive columns of a row, adds the splits in order (deterministic)
// and runs the same fused epilogue the single-pass kernel would have run
__global__ void __launch_bounds__(256) fq3c_splitk_reduce_kernel(const fq3c_op o, const int splits, const int Nw) {
  const int groups = Nw >> 3;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)(o.M - o.m_begin) * groups) return;
  const int r_rel = (int)(idx / groups);
  const int row = o.m_begin + r_rel, col = (int)(idx - (long)r_rel * groups) * 8;
  float a8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const float* ws = reinterpret_cast<const float*>(o.ws);
  for (int z = 0; z < splits; ++z) {
    const float4* src = reinterpret_cast<const float4*>(ws + ((size_t)z * o.M + row) * Nw + col);
    const float4 v0 = src[0], v1 = src[1];
    a8[0] += v0.x; a8[1] += v0.y; a8[2] += v0.z; a8[3] += v0.w;
    a8[4] += v1.x; a8[5] += v1.y; a8[6] += v1.z; a8[7] += v1.w;
  }
  epilogue_row8(o, row, col, a8);
}

__global__ void fq3c_rvq_kernel(const fq3c_op o) {
  const long long* codes = reinterpret_cast<const long long*>(o.A);
  const bf16* cb = reinterpret_cast<const bf16*>(o.B);
  const int Q = o.i1, nsem = o.i0, dim = o.i2, cbsize = o.K;
  bf16* C = reinterpret_cast<bf16*>(o.C);
  const int m = blockIdx.x;
  for (int d = threadIdx.x; d < dim; d += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int g = 0; g < Q; ++g) {
      long long c = codes[(size_t)m * Q + g];
      c = c < 0 ? 0 : (c >= cbsize ? cbsize - 1 : c);
      const float v = __bfloat162float(cb[((size_t)g * cbsize + c) * dim + d]);
      if (g < nsem) a += v; else b += v;
    }
    C[(size_t)m * o.ldc + d] = __float2bfloat16_rn(a);
    C[(size_t)m * o.ldc + dim + d] = __float2bfloat16_rn(b);
  }
}

__device__ __forceinline__ float block_sum(float v, float* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < nw; ++i) t += red[i];
  return t;
}

__global__ void fq3c_rmsnorm_kernel(const fq3c_op o) {
  __shared__ float red[32];
  const bf16* x = reinterpret_cast<const bf16*>(o.A) + (size_t)blockIdx.x * o.lda;
  bf16* y = reinterpret_cast<bf16*>(o.C) + (size_t)blockIdx.x * o.ldc;
  const float* w = reinterpret_cast<const float*>(o.scale);
  float ss = 0.f;
  for (int c = threadIdx.x; c < o.N; c += blockDim.x) { const float v = __bfloat162float(x[c]); ss += v * v; }
  ss = block_sum(ss, red);
  const float rs = rsqrtf(ss / o.N + o.f0);
  for (int c = threadIdx.x; c < o.N; c += blockDim.x)
    y[c] = __float2bfloat16_rn(bf16r(bf16r(__bfloat162float(x[c]) * rs) * w[c]));
}

// RMSNorm of the rows a GEMM just finished (fq3c_op.norm_out), with the arithmetic of fq3c_rmsnorm_kernel: same column-strided
// partial sums, same block tree, same rounding points — the fused form changes the launch count, not a bit of the result.
__device__ __forceinline__ void row_rmsnorm(const fq3c_op& o, int row, float* red) {
  const bf16* x = reinterpret_cast<const bf16*>(o.C) + (size_t)row * o.ldc;
  bf16* y = reinterpret_cast<bf16*>(o.norm_out) + (size_t)row * o.norm_ld;
  const float* w = reinterpret_cast<const float*>(o.norm_w);
  float ss = 0.f;
  for (int c = threadIdx.x; c < o.N; c += blockDim.x) { const float v = __bfloat162float(x[c]); ss += v * v; }
  ss = block_sum(ss, red);
  const float rs = rsqrtf(ss / o.N + o.norm_eps);
  for (int c = threadIdx.x; c < o.N; c += blockDim.x)
    y[c] = __float2bfloat16_rn(bf16r(bf16r(__bfloat162float(x[c]) * rs) * w[c]));
}
__global__ void __launch_bounds__(256) fq3c_rownorm_kernel(const fq3c_op o) {
  __shared__ float red[32];
  row_rmsnorm(o, o.m_begin + blockIdx.x, red);
}
// split-K reduction + epilogue + RMSNorm, one CTA per output row: the row's eight-column groups are reduced and written by
// their threads, then the whole CTA reads the finished row back (it is in L1 / L2) for the norm
__global__ void __launch_bounds__(256) fq3c_splitk_reduce_norm_kernel(const fq3c_op o, const int splits, const int Nw) {
  __shared__ float red[32];
  const int row = o.m_begin + blockIdx.x;
  const float* ws = reinterpret_cast<const float*>(o.ws);
  for (int col = threadIdx.x * 8; col < Nw; col += blockDim.x * 8) {
    float a8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int z = 0; z < splits; ++z) {
      const float4* src = reinterpret_cast<const float4*>(ws + ((size_t)z * o.M + row) * Nw + col);
      const float4 v0 = src[0], v1 = src[1];
      a8[0] += v0.x; a8[1] += v0.y; a8[2] += v0.z; a8[3] += v0.w;
      a8[4] += v1.x; a8[5] += v1.y; a8[6] += v1.z; a8[7] += v1.w;
    }
    epilogue_row8(o, row, col, a8);
  }
  __syncthreads();  // the row's global writes are visible to the whole block behind the barrier
  row_rmsnorm(o, row, red);
}

__global__ void fq3c_layernorm_kernel(const fq3c_op o) {
  __shared__ float red[32];
  const bf16* x = reinterpret_cast<const bf16*>(o.A) + (size_t)blockIdx.x * o.lda;
  bf16* y = reinterpret_cast<bf16*>(o.C) + (size_t)blockIdx.x * o.ldc;
  const float* w = reinterpret_cast<const float*>(o.scale);
  const float* b = reinterpret_cast<const float*>(o.bias);
  float s = 0.f;
  for (int c = threadIdx.x; c < o.N; c += blockDim.x) s += __bfloat162float(x[c]);
  const float mean = block_sum(s, red) / o.N;
  float v = 0.f;
  for (int c = threadIdx.x; c < o.N; c += blockDim.x) { const float d = __bfloat162float(x[c]) - mean; v += d * d; }
  const float rs = rsqrtf(block_sum(v, red) / o.N + o.f0);
  for (int c = threadIdx.x; c < o.N; c += blockDim.x)
    y[c] = __float2bfloat16_rn((__bfloat162float(x[c]) - mean) * rs * w[c] + b[c]);
}

// RoPE in place on q and k inside the fused qkv buffer (HF rotate_half convention, bf16 rounding per product).
__global__ void fq3c_rope_kernel(const fq3c_op o) {
  const int t = blockIdx.x, nheads = o.i0, d = o.i1, col0 = o.i2, half = d / 2;
  const int pos = t + (o.p0 ? *reinterpret_cast<const int*>(o.p0) : 0);
  bf16* row = reinterpret_cast<bf16*>(const_cast<void*>(o.A)) + (size_t)t * o.lda + col0;
  for (int idx = threadIdx.x; idx < nheads * half; idx += blockDim.x) {
    const int h = idx / half, i = idx - h * half;
    const float inv = powf(o.f0, -2.0f * i / d);
    const float ang = (float)pos * inv;
    const float c = bf16r(cosf(ang)), s = bf16r(sinf(ang));
    bf16* p = row + h * d;
    const float a = __bfloat162float(p[i]), b = __bfloat162float(p[i + half]);
    p[i] = __float2bfloat16_rn(bf16r(a * c) + bf16r(-b * s));
    p[i + half] = __float2bfloat16_rn(bf16r(b * c) + bf16r(a * s));
  }
}

// Sliding-window causal attention, one warp per (query, head); head_dim in {32, 64, 128}: a lane owns d/32 contiguous
// dims (one 2/4/8-byte load per K and V row), four keys are scored per step so their shuffle reductions overlap; the
// online-softmax updates are applied key by key in order.
template <int PER>
__device__ __forceinline__ void attn_load(const bf16* p, float (&v)[4]) {
  if (PER == 4) {
    const uint2 r = *reinterpret_cast<const uint2*>(p);
    v[0] = __uint_as_float(r.x << 16); v[1] = __uint_as_float(r.x & 0xffff0000u);
    v[2] = __uint_as_float(r.y << 16); v[3] = __uint_as_float(r.y & 0xffff0000u);
  } else if (PER == 2) {
    const uint32_t r = *reinterpret_cast<const uint32_t*>(p);
    v[0] = __uint_as_float(r << 16); v[1] = __uint_as_float(r & 0xffff0000u); v[2] = 0.f; v[3] = 0.f;
  } else {
    v[0] = __bfloat162float(p[0]); v[1] = 0.f; v[2] = 0.f; v[3] = 0.f;
  }
}
template <int PER>
__device__ __forceinline__ void attn_warp(const fq3c_op& o, int t_out, int h, int lane) {
  const int nh = o.i0, nkv = o.i1, d = o.i2, win = o.K;
  const int kvh = h / (nh / nkv);
  const bf16* base = reinterpret_cast<const bf16*>(o.A);
  const int qoff = h * d + lane * PER, koff = nh * d + kvh * d + lane * PER, voff = nh * d + nkv * d + kvh * d + lane * PER;
  float q[4], acc[4] = {0.f, 0.f, 0.f, 0.f};
  const int hist = o.taps, t = t_out + hist;  // stateful decode: `hist` rows of earlier chunks sit in front of the query rows
  const int first = o.p0 ? hist - min(*reinterpret_cast<const int*>(o.p0), hist) : 0;
  attn_load<PER>(base + (size_t)t * o.lda + qoff, q);
  const float scale = rsqrtf((float)d);
  float m = -INFINITY, l = 0.f;
  const int j0 = max(first, t - win + 1);
  for (int jb = j0; jb <= t; jb += 4) {
    float s[4], v[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = min(jb + u, t);
      float k[4];
      attn_load<PER>(base + (size_t)j * o.lda + koff, k);
      attn_load<PER>(base + (size_t)j * o.lda + voff, v[u]);
      s[u] = q[0] * k[0] + q[1] * k[1] + q[2] * k[2] + q[3] * k[3];
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
      for (int u = 0; u < 4; ++u) s[u] += __shfl_xor_sync(0xffffffffu, s[u], off);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (jb + u <= t) {
        const float su = bf16r(bf16r(s[u]) * scale);
        const float mn = fmaxf(m, su), corr = expf(m - mn), p = expf(su - mn);
        l = l * corr + p;
#pragma unroll
        for (int i = 0; i < PER; ++i) acc[i] = acc[i] * corr + p * v[u][i];
        m = mn;
      }
    }
  }
  bf16* out = reinterpret_cast<bf16*>(o.C) + (size_t)t_out * o.ldc + h * d + lane * PER;
#pragma unroll
  for (int i = 0; i < PER; ++i) out[i] = __float2bfloat16_rn(acc[i] / l);
}
// any head_dim <= 128 (lane owns dims lane, lane + 32, ...): the small test configurations
__device__ __forceinline__ void attn_warp_generic(const fq3c_op& o, int t_out, int h, int lane) {
  const int nh = o.i0, nkv = o.i1, d = o.i2, win = o.K;
  const int kvh = h / (nh / nkv);
  const bf16* base = reinterpret_cast<const bf16*>(o.A);
  const int qoff = h * d, koff = nh * d + kvh * d, voff = nh * d + nkv * d + kvh * d;
  const int hist = o.taps, t = t_out + hist;
  const int first = o.p0 ? hist - min(*reinterpret_cast<const int*>(o.p0), hist) : 0;
  const int per = (d + 31) / 32;  // <= 4
  float q[4], acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = 0; i < per; ++i) {
    const int c = lane + i * 32;
    q[i] = c < d ? __bfloat162float(base[(size_t)t * o.lda + qoff + c]) : 0.f;
  }
  const float scale = rsqrtf((float)d);
  float m = -INFINITY, l = 0.f;
  const int j0 = max(first, t - win + 1);
  for (int j = j0; j <= t; ++j) {
    float s = 0.f;
    for (int i = 0; i < per; ++i) {
      const int c = lane + i * 32;
      if (c < d) s += q[i] * __bfloat162float(base[(size_t)j * o.lda + koff + c]);
    }
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    s = bf16r(bf16r(s) * scale);
    const float mn = fmaxf(m, s), corr = expf(m - mn), p = expf(s - mn);
    l = l * corr + p;
    for (int i = 0; i < per; ++i) {
      const int c = lane + i * 32;
      const float v = c < d ? __bfloat162float(base[(size_t)j * o.lda + voff + c]) : 0.f;
      acc[i] = acc[i] * corr + p * v;
    }
    m = mn;
  }
  bf16* out = reinterpret_cast<bf16*>(o.C) + (size_t)t_out * o.ldc + h * d;
  for (int i = 0; i < per; ++i) {
    const int c = lane + i * 32;
    if (c < d) out[c] = __float2bfloat16_rn(acc[i] / l);
  }
}
__global__ void fq3c_attn_kernel(const fq3c_op o) {
  const int nh = o.i0, T = o.M;
  const int wpb = blockDim.x >> 5;
  const int gw = blockIdx.x * wpb + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (gw >= T * nh) return;
  const int t = gw / nh, h = gw - t * nh;
  const bool aligned = ((o.lda * 2) % 8 == 0) && ((o.ldc * 2) % 8 == 0);
  if (o.i2 == 128 && aligned) attn_warp<4>(o, t, h, lane);
  else if (o.i2 == 64 && aligned) attn_warp<2>(o, t, h, lane);
  else attn_warp_generic(o, t, h, lane);
}

// Causal attention of the dense prompt prefill (FQ3C_ATTN with flags bit 0: head_dim 128, no history rows, window >= M, M <= 2048):
// one warp per (query, head), two phases with different lane roles instead of the key-by-key online softmax above —
//   1. four lanes per KEY, eight keys per step: a lane multiplies 32 dims of its key (q from shared memory, four 16-byte loads of the
//      key row), two shuffles finish the dot product (against five per key in the kernel above); the score, rounded to bf16 exactly
//      where the kernel above rounds it, goes to shared memory, and after the warp's max is known the same lane replaces it by
//      p = exp(score - max) — once per key, not once per key and lane;
//   2. lanes take DIMS (four each): every lane walks the keys in order, p straight from shared memory, and accumulates l and its
//      four output dims — no rescaling, no shuffles.
// Same rounding points as attn_warp (score -> bf16 -> x scale -> bf16; fp32 softmax and accumulation; bf16 output); the fp32 sums run in
// another order, which the prefill tests bound against the oracle and the chunked prefill.  M = 240 (ICL prompt): 78 -> ~12 us per layer.
constexpr int APF_WARPS = 8;
__global__ void __launch_bounds__(APF_WARPS * 32) fq3c_attn_prefill_kernel(const fq3c_op o, const int tpad) {
  extern __shared__ unsigned char apf_smem[];
  const int nh = o.i0, nkv = o.i1, T = o.M, win = o.K;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // heavy queries (many keys) first: the grid's tail then consists of the short ones.  (Warps of a CTA = the heads of one query;
  // consecutive queries of one head per CTA was measured: slower, 99 vs 66 us at M = 240.)
  const int gw = blockIdx.x * APF_WARPS + warp;
  if (gw >= T * nh) return;
  const int t = T - 1 - gw / nh, h = gw - (gw / nh) * nh;
  const int kvh = h / (nh / nkv);
  float* qs = reinterpret_cast<float*>(apf_smem) + warp * 128;
  float* sc = reinterpret_cast<float*>(apf_smem + APF_WARPS * 128 * 4) + (size_t)warp * tpad;
  const bf16* base = reinterpret_cast<const bf16*>(o.A);
  const size_t lda = (size_t)o.lda;
  const int koff = nh * 128 + kvh * 128, voff = (nh + nkv) * 128 + kvh * 128;
  {
    const uint2 r = *reinterpret_cast<const uint2*>(base + (size_t)t * lda + h * 128 + lane * 4);
    *reinterpret_cast<float4*>(qs + lane * 4) = make_float4(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xffff0000u),
                                                            __uint_as_float(r.y << 16), __uint_as_float(r.y & 0xffff0000u));
  }
  __syncwarp();
  const float scale = rsqrtf(128.f);
  const int j0 = max(0, t - win + 1);
  float mx = -INFINITY;
  // four lanes per key, eight keys per step: the four lanes of a key read 64 contiguous bytes per load instruction (a warp-wide
  // load touches 8 lines; one lane per key would touch 32 and the L1 tag stage, one line per cycle, becomes the bound: measured)
  const int sub = lane & 3, kq = lane >> 2;
  for (int jb = j0; jb <= t; jb += 16) {  // two groups of eight keys per step: sixteen rows in flight per L2 round trip
    uint4 kk[2][4];
    bool ok[2];
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const int j = jb + g * 8 + kq;
      ok[g] = j <= t;
      const uint4* kr = reinterpret_cast<const uint4*>(base + (size_t)(ok[g] ? j : t) * lda + koff) + sub;
#pragma unroll
      for (int c = 0; c < 4; ++c) kk[g][c] = kr[c * 4];
    }
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      float d0 = 0.f, d1 = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float* qp = qs + (c * 4 + sub) * 8;
        const float4 qa = *reinterpret_cast<const float4*>(qp), qb = *reinterpret_cast<const float4*>(qp + 4);
        d0 += qa.x * __uint_as_float(kk[g][c].x << 16) + qa.y * __uint_as_float(kk[g][c].x & 0xffff0000u) +
              qa.z * __uint_as_float(kk[g][c].y << 16) + qa.w * __uint_as_float(kk[g][c].y & 0xffff0000u);
        d1 += qb.x * __uint_as_float(kk[g][c].z << 16) + qb.y * __uint_as_float(kk[g][c].z & 0xffff0000u) +
              qb.z * __uint_as_float(kk[g][c].w << 16) + qb.w * __uint_as_float(kk[g][c].w & 0xffff0000u);
      }
      float dsum = d0 + d1;
      dsum += __shfl_xor_sync(0xffffffffu, dsum, 1);
      dsum += __shfl_xor_sync(0xffffffffu, dsum, 2);
      const float su = bf16r(bf16r(dsum) * scale);
      if (ok[g]) {
        if (sub == 0) sc[jb + g * 8 + kq - j0] = su;
        mx = fmaxf(mx, su);
      }
    }
  }
  for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  // the exponentials once per key (the lane that scored it), not once per key and lane: expf is ~20 instructions
  for (int j = j0 + lane; j <= t; j += 32) sc[j - j0] = expf(sc[j - j0] - mx);
  __syncwarp();
  float l = 0.f, acc[4] = {0.f, 0.f, 0.f, 0.f};
  const bf16* vp = base + voff + lane * 4;
  int j = j0;
  // sixteen keys per step: the step is one L2 round trip (the V rows of a prompt are not in L1), so all sixteen loads go out before
  // the first one is used; the accumulation itself stays in key order
  for (; j + 15 <= t; j += 16) {
    uint2 r[16];
    float pr[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) r[u] = *reinterpret_cast<const uint2*>(vp + (size_t)(j + u) * lda);
#pragma unroll
    for (int u = 0; u < 16; ++u) pr[u] = sc[j + u - j0];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      l += pr[u];
      acc[0] += pr[u] * __uint_as_float(r[u].x << 16); acc[1] += pr[u] * __uint_as_float(r[u].x & 0xffff0000u);
      acc[2] += pr[u] * __uint_as_float(r[u].y << 16); acc[3] += pr[u] * __uint_as_float(r[u].y & 0xffff0000u);
    }
  }
  for (; j <= t; ++j) {
    const uint2 r0 = *reinterpret_cast<const uint2*>(vp + (size_t)j * lda);
    const float p0 = sc[j - j0];
    l += p0;
    acc[0] += p0 * __uint_as_float(r0.x << 16); acc[1] += p0 * __uint_as_float(r0.x & 0xffff0000u);
    acc[2] += p0 * __uint_as_float(r0.y << 16); acc[3] += p0 * __uint_as_float(r0.y & 0xffff0000u);
  }
  bf16* out = reinterpret_cast<bf16*>(o.C) + (size_t)t * o.ldc + h * 128 + lane * 4;
  const float inv = 1.f / l;
  uint2 w;
  w.x = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(acc[0] * inv)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(acc[1] * inv)) << 16);
  w.y = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(acc[2] * inv)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(acc[3] * inv)) << 16);
  *reinterpret_cast<uint2*>(out) = w;
}

// Dense prefill helper (talker): one warp per (row, head) of the fused qkv rows A [M, lda] = (q heads | k heads | v heads).
//   q / k heads: per-head RMSNorm with bf16 gamma (q: p0, k: p1), HF rounding points, then rotary embedding from the bf16
//                tables B (cos) / bias (sin) at position row + i2 (rotate-half), in place;
//   k / v heads: the finished row is also written to the static KV cache C (K) / C2 (V), laid out [kv_head][ldc][128].
// Same arithmetic as head_norm_rope / the KV append of the decode kernel (fq3_kernel.cuh), so decode steps can attend to it.
__global__ void fq3c_qknorm_rope_kv_kernel(const fq3c_op o) {
  const int nq = o.i0, nkv = o.i1, pos0 = o.i2, nheads = nq + 2 * nkv;
  const int wpb = blockDim.x >> 5;
  const int gw = blockIdx.x * wpb + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (gw >= o.M * nheads) return;
  const int t = gw / nheads, h = gw - t * nheads;
  bf16* row = reinterpret_cast<bf16*>(const_cast<void*>(o.A)) + (size_t)t * o.lda + (size_t)h * 128 + lane * 4;
  uint2 raw = *reinterpret_cast<const uint2*>(row);
  const int pos = pos0 + t;
  if (h < nq + nkv) {
    const bf16* gamma = reinterpret_cast<const bf16*>(h < nq ? o.p0 : o.p1) + lane * 4;
    const bf16* cosp = reinterpret_cast<const bf16*>(o.B) + (size_t)pos * 128 + lane * 4;
    const bf16* sinp = reinterpret_cast<const bf16*>(o.bias) + (size_t)pos * 128 + lane * 4;
    float x[4] = {__uint_as_float(raw.x << 16), __uint_as_float(raw.x & 0xffff0000u), __uint_as_float(raw.y << 16),
                  __uint_as_float(raw.y & 0xffff0000u)};
    float ss = x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3];
    for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
    const float rs = rsqrtf(ss * (1.f / 128.f) + o.f0);
    float y[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) y[i] = bf16r(bf16r(x[i] * rs) * __bfloat162float(gamma[i]));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float partner = __shfl_xor_sync(0xffffffffu, y[i], 16);
      const float rot = (lane < 16) ? -partner : partner;
      x[i] = bf16r(bf16r(y[i] * __bfloat162float(cosp[i])) + bf16r(rot * __bfloat162float(sinp[i])));
    }
    const __nv_bfloat162 lo = __floats2bfloat162_rn(x[0], x[1]), hi = __floats2bfloat162_rn(x[2], x[3]);
    raw = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
    *reinterpret_cast<uint2*>(row) = raw;
  }
  if (h >= nq) {
    const bool is_v = h >= nq + nkv;
    const int kvh = is_v ? h - nq - nkv : h - nq;
    bf16* dst = reinterpret_cast<bf16*>(is_v ? o.C2 : o.C) + ((size_t)kvh * o.ldc + pos) * 128 + lane * 4;
    *reinterpret_cast<uint2*>(dst) = raw;
  }
}

__global__ void fq3c_dwconv_kernel(const fq3c_op o) {
  const size_t n = (size_t)o.M * o.N;
  const bf16* x = reinterpret_cast<const bf16*>(o.A);
  const float* w = reinterpret_cast<const float*>(o.B);
  const float* b = reinterpret_cast<const float*>(o.bias);
  bf16* y = reinterpret_cast<bf16*>(o.C);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int t = (int)(i / o.N), c = (int)(i - (size_t)t * o.N);
    float a = b[c];
    for (int j = 0; j < o.taps; ++j) {
      const int r = t - (o.taps - 1) + j;  // rows [-i0, 0) are the history rows of a stateful decode
      if (r >= -o.i0) a += w[c * o.taps + j] * __bfloat162float(x[(long)r * o.lda + c]);
    }
    y[(size_t)t * o.ldc + c] = __float2bfloat16_rn(a);
  }
}

__global__ void fq3c_copy_kernel(const fq3c_op o) {
  const int nv = o.N >> 3;  // 16-byte pieces per row (N, lda, ldc are multiples of 8)
  const size_t n = (size_t)o.M * nv;
  const bf16* x = reinterpret_cast<const bf16*>(o.A);
  bf16* y = reinterpret_cast<bf16*>(o.C);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t t = i / nv, c = (i - t * nv) * 8;
    *reinterpret_cast<uint4*>(y + t * o.ldc + c) = *reinterpret_cast<const uint4*>(x + t * o.lda + c);
  }
}
__global__ void fq3c_advance_kernel(int* counter, int by) { *counter += by; }
// FQ3C_ROLL: blockIdx.x = history buffer, blockIdx.y = column slice (slices are independent).  dst row r <- src row r + new; when the
// ranges overlap (new < hist) the rows go in batches of `new`: a batch's sources are the next batch's destinations, so a block
// barrier separates them; within a batch sources and destinations are disjoint.
constexpr int ROLL_SLICES = 8;
__global__ void __launch_bounds__(256) fq3c_roll_kernel(const long long* desc, int* counter, int by) {
  const long long* d = desc + (size_t)blockIdx.x * 5;
  bf16* base = reinterpret_cast<bf16*>(d[0]);
  const int hist = (int)d[1], n_new = (int)d[2], cols = (int)d[3], ld = (int)d[4];
  const int nv = cols >> 3;
  const int p0 = (int)(((long)blockIdx.y * nv) / ROLL_SLICES), p1 = (int)(((long)(blockIdx.y + 1) * nv) / ROLL_SLICES);
  const int np = p1 - p0;
  const int batch = n_new >= hist ? hist : n_new;
  for (int r0 = 0; r0 < hist && np > 0; r0 += batch) {
    const int rows = min(batch, hist - r0);
    for (int i = threadIdx.x; i < rows * np; i += blockDim.x) {
      const int r = r0 + i / np, p = p0 + i % np;
      *reinterpret_cast<uint4*>(base + (size_t)r * ld + p * 8) = *reinterpret_cast<const uint4*>(base + (size_t)(r + n_new) * ld + p * 8);
    }
    __syncthreads();
  }
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && counter) *counter += by;
}

__global__ void fq3c_snake_kernel(const fq3c_op o) {
  const size_t n = (size_t)o.M * o.N;
  const bf16* x = reinterpret_cast<const bf16*>(o.A);
  const float* ea = reinterpret_cast<const float*>(o.p0);
  const float* ib = reinterpret_cast<const float*>(o.p1);
  bf16* y = reinterpret_cast<bf16*>(o.C);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int t = (int)(i / o.N), c = (int)(i - (size_t)t * o.N);
    const float v = __bfloat162float(x[(size_t)t * o.lda + c]);
    const float s = sinf(v * ea[c]);
    y[(size_t)t * o.ldc + c] = __float2bfloat16_rn(v + ib[c] * s * s);
  }
}

int fail(const std::string& m) { g_err = m; return -1; }

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    cudaGetLastError();
  }
  return fn;
}
// bf16 matrix [rows, inner] with a row pitch of ld elements, boxes of box_rows x 64 columns, 128-byte swizzle, zeros outside
bool make_map(CUtensorMap* m, const void* base, uint64_t inner, uint64_t rows, uint64_t ld, uint32_t box_rows) {
  EncodeTiledFn fn = encode_tiled();
  if (!fn || !base || inner == 0 || rows == 0) return false;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld * 2) % 16) return false;
  const cuuint64_t dims[2] = {inner, rows};
  const cuuint64_t strides[1] = {ld * 2};
  const cuuint32_t box[2] = {(cuuint32_t)TK, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

extern "C" {

int fq3c_abi_version(void) { return 5; }
const char* fq3c_last_error(void) { return g_err.c_str(); }
int64_t fq3c_launch_count(void) { return g_launches.load(); }

// Opt-in shared-memory sizes are a per-DEVICE function attribute: a process that serves several GPUs (server.py --gpus N: one scheduler
// thread per device) must set them on each device it launches on, not once per process.
static bool device_attrs_ready() {
  static bool done[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
  if (done[dev]) return true;
  if (cudaFuncSetAttribute(fq3c_gemm_tc5_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, T_SMEM) != cudaSuccess ||
      cudaFuncSetAttribute(fq3c_gemm_tc5_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, t_smem(3)) != cudaSuccess ||
      cudaFuncSetAttribute(fq3c_gemm_tc5_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, T_SMEM) != cudaSuccess ||
      cudaFuncSetAttribute(fq3c_gemm_tc5_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, t_smem(3)) != cudaSuccess ||
      cudaFuncSetAttribute(fq3c_attn_prefill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, APF_WARPS * (128 + 2048) * 4) != cudaSuccess)
    return false;
  done[dev] = true;
  return true;
}

int fq3c_run(const fq3c_op* ops, int n_ops, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  for (int i = 0; i < n_ops; ++i) {
    const fq3c_op& o = ops[i];
    if (o.M <= 0) continue;
    switch (o.kind) {
      case FQ3C_GEMM: {
        if (o.K % 8 || o.cin % 8 || o.taps > 8 || o.col_mod <= 0) return fail("gemm: K and cin must be multiples of 8, taps <= 8");
        if (o.m_begin < 0 || o.m_begin >= o.M) return fail("gemm: m_begin outside [0, M)");
        if (o.norm_out && ((o.flags & (FQ3C_SWIGLU | FQ3C_OUT_F32)) || !o.norm_w || o.norm_ld < o.N))
          return fail("gemm: a fused RMSNorm needs a bf16 [M, N] output, weights and norm_ld >= N");
        static int use_tc5 = -1;
        if (use_tc5 < 0) {
          const char* e = getenv("FQ3C_TCGEN05");
          use_tc5 = (e == nullptr || atoi(e) != 0) ? 1 : 0;
        }
        if (use_tc5 && !device_attrs_ready()) return fail("cannot reserve shared memory for the tcgen05 GEMM / prefill attention");
        static int tc5_min_k = -1;
        if (tc5_min_k < 0) { const char* e3 = getenv("FQ3C_TC5_MINK"); tc5_min_k = e3 ? atoi(e3) : TK; }  // one k-block is enough: the vectorised epilogue beats the mma.sync kernel even at K = 96
        if (use_tc5 && o.N >= 16 && o.K >= tc5_min_k) {
          // Tile width: a multiple of 16 columns (a legal UMMA N at M = 128), chosen so that the grid covers the SMs.
          // Tall operands (M >= 128) keep BN >= 64 (every column tile re-reads the A rows); short ones (the transformer at
          // 8-33 frames) are weight-streaming problems: the width that minimises waves x tile cost, down to 16 columns.
          const int mt = (o.M - o.m_begin + TM - 1) / TM;
          int bn = TMAXN;
          if (o.M >= TM) {
            if (((o.N + bn - 1) / bn) * mt < 100) bn = 64;
          } else {
            long best = -1;
            for (int cand = TMAXN; cand >= 16; cand >>= 1) {
              const long tiles = (long)((o.N + cand - 1) / cand) * mt;
              const long cost = ((tiles + 147) / 148) * (cand + 48);  // waves x (columns + fixed per-tile work)
              if (best < 0 || cost < best) { best = cost; bn = cand; }
            }
          }
          // Split-K when the tiles would cover less than half of the SMs and the reduction is long (the transformer's
          // o / down projections at 8-240 rows: 16-32 tiles each walking 32-96 k-blocks alone): wider tiles, every tile
          // split over up to eight CTAs, partial tiles through the caller's workspace.
          const int KT = (o.K + TK - 1) / TK;
          const int Nw = (o.N + 7) & ~7;
          int splits = 1;
          if (o.ws && o.M <= 4 * TM && KT >= 8) {
            const int bn_s = o.N >= 64 ? 64 : 32;
            // the split count must not depend on m_begin: a tail-only decode has to add the same partial sums in the same order
            const int tiles_s = ((o.N + bn_s - 1) / bn_s) * ((o.M + TM - 1) / TM);
            int sp = std::min(std::min(8, KT / 4), 148 / std::max(1, tiles_s));
            while (sp > 1 && (int64_t)sp * o.M * Nw * 4 > o.ws_bytes) --sp;
            if (sp > 1) { splits = sp; bn = bn_s; }
          }
          const int nt = (o.N + bn - 1) / bn;
          bn = std::min(TMAXN, (((o.N + nt - 1) / nt) + 15) / 16 * 16);
          dim3 grid((o.N + bn - 1) / bn, mt, splits);
          // Many tiles (the vocoder convolutions: hundreds of 128-row tiles with 2-11 k-blocks and a long SnakeBeta epilogue):
          // three stages instead of six so that two CTAs share an SM and one's epilogue overlaps the other's main loop.
          static int two_ctas = -1;
          if (two_ctas < 0) { const char* e2 = getenv("FQ3C_TWO_CTAS"); two_ctas = (e2 == nullptr || atoi(e2) != 0) ? 1 : 0; }
          const long n_ctas = (long)grid.x * grid.y * grid.z;
          // operand tiles by TMA tensor loads where a 64-column k-block stays inside one tap (FQ3C_TMA=0: cp.async loaders everywhere)
          static int use_tma = -1;
          if (use_tma < 0) { const char* e4 = getenv("FQ3C_TMA"); use_tma = (e4 == nullptr || atoi(e4) != 0) ? 1 : 0; }
          CUtensorMap tmA, tmB;
          memset(&tmA, 0, sizeof tmA); memset(&tmB, 0, sizeof tmB);
          const bool tma = use_tma && (o.taps == 1 || o.cin % TK == 0) &&
                           make_map(&tmA, o.A, (uint64_t)o.cin, (uint64_t)o.a_rows, (uint64_t)o.lda, TM) &&
                           make_map(&tmB, o.B, (uint64_t)o.K, (uint64_t)o.N, (uint64_t)o.K, (uint32_t)bn);
          const bool two = two_ctas && n_ctas >= 2 * 148 && KT <= 24;
          if (tma) {
            if (two) fq3c_gemm_tc5_kernel<2, true><<<grid, TTHREADS, t_smem(3), s>>>(o, bn, splits, Nw, 3, tmA, tmB);
            else fq3c_gemm_tc5_kernel<1, true><<<grid, TTHREADS, T_SMEM, s>>>(o, bn, splits, Nw, TSTAGES, tmA, tmB);
          } else {
            if (two) fq3c_gemm_tc5_kernel<2, false><<<grid, TTHREADS, t_smem(3), s>>>(o, bn, splits, Nw, 3, tmA, tmB);
            else fq3c_gemm_tc5_kernel<1, false><<<grid, TTHREADS, T_SMEM, s>>>(o, bn, splits, Nw, TSTAGES, tmA, tmB);
          }
          if (splits > 1) {
            if (o.norm_out) {
              fq3c_splitk_reduce_norm_kernel<<<(unsigned)(o.M - o.m_begin), 256, 0, s>>>(o, splits, Nw);
            } else {
              const long n = (long)(o.M - o.m_begin) * (Nw >> 3);
              fq3c_splitk_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(o, splits, Nw);
            }
            count_launches(1);
          } else if (o.norm_out) {
            fq3c_rownorm_kernel<<<(unsigned)(o.M - o.m_begin), 256, 0, s>>>(o);
            count_launches(1);
          }
          break;
        }
        dim3 grid((o.N + BN - 1) / BN, (o.M - o.m_begin + BM - 1) / BM);
        fq3c_gemm_kernel<<<grid, 128, 0, s>>>(o);
        if (o.norm_out) {
          fq3c_rownorm_kernel<<<(unsigned)(o.M - o.m_begin), 256, 0, s>>>(o);
          count_launches(1);
        }
        break;
      }
      case FQ3C_RVQ: fq3c_rvq_kernel<<<o.M, 128, 0, s>>>(o); break;
      case FQ3C_RMSNORM: fq3c_rmsnorm_kernel<<<o.M, 256, 0, s>>>(o); break;
      case FQ3C_LAYERNORM: fq3c_layernorm_kernel<<<o.M, 256, 0, s>>>(o); break;
      case FQ3C_ROPE: fq3c_rope_kernel<<<o.M, 256, 0, s>>>(o); break;
      case FQ3C_ATTN: {
        if (o.i2 > 128) return fail("attn: head_dim > 128");
        static int use_apf = -1;
        if (use_apf < 0) { const char* e5 = getenv("FQ3C_ATTN_PREFILL"); use_apf = (e5 == nullptr || atoi(e5) != 0) ? 1 : 0; }
        if (use_apf && (o.flags & 1) && o.i2 == 128 && o.taps == 0 && !o.p0 && o.M <= 2048 && (o.lda % 8) == 0 && (o.ldc % 4) == 0 &&
            (o.i0 % o.i1) == 0) {
          const int tpad = (o.M + 7) & ~7;
          const size_t smem = (size_t)APF_WARPS * 128 * 4 + (size_t)APF_WARPS * tpad * 4;
          if (!device_attrs_ready()) return fail("cannot reserve shared memory for the prefill attention");
          const int warps = o.M * o.i0;
          fq3c_attn_prefill_kernel<<<(warps + APF_WARPS - 1) / APF_WARPS, APF_WARPS * 32, smem, s>>>(o, tpad);
          break;
        }
        const int warps = o.M * o.i0;
        fq3c_attn_kernel<<<(warps + 7) / 8, 256, 0, s>>>(o);
        break;
      }
      case FQ3C_DWCONV: {
        const size_t n = (size_t)o.M * o.N;
        fq3c_dwconv_kernel<<<(unsigned)std::min<size_t>(2048, (n + 255) / 256), 256, 0, s>>>(o);
        break;
      }
      case FQ3C_QKNORM_ROPE_KV: {
        if (o.lda % 4 || !o.C || !o.C2 || !o.p0 || !o.p1 || !o.B || !o.bias) return fail("qknorm_rope_kv: missing operand");
        const int warps = o.M * (o.i0 + 2 * o.i1);
        fq3c_qknorm_rope_kv_kernel<<<(warps + 7) / 8, 256, 0, s>>>(o);
        break;
      }
      case FQ3C_SNAKE: {
        const size_t n = (size_t)o.M * o.N;
        fq3c_snake_kernel<<<(unsigned)std::min<size_t>(4096, (n + 255) / 256), 256, 0, s>>>(o);
        break;
      }
      case FQ3C_COPY: {
        if ((o.N % 8) || (o.lda % 8) || (o.ldc % 8) || !o.A || !o.C) return fail("copy: N, lda, ldc must be multiples of 8");
        const size_t n = (size_t)o.M * (o.N >> 3);
        fq3c_copy_kernel<<<(unsigned)std::min<size_t>(1024, (n + 255) / 256), 256, 0, s>>>(o);
        break;
      }
      case FQ3C_ROLL: {
        if (!o.A || o.M <= 0) return fail("roll: no descriptors");
        fq3c_roll_kernel<<<dim3((unsigned)o.M, ROLL_SLICES), 256, 0, s>>>(reinterpret_cast<const long long*>(o.A), reinterpret_cast<int*>(o.C), o.i0);
        break;
      }
      case FQ3C_ADVANCE: {
        if (!o.C) return fail("advance: null counter");
        fq3c_advance_kernel<<<1, 1, 0, s>>>(reinterpret_cast<int*>(o.C), o.i0);
        break;
      }
      default: return fail("unknown op kind");
    }
    count_launches(1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(std::string("launch failed at op ") + std::to_string(i) + ": " + cudaGetErrorString(e));
  }
  return 0;
}


/* A plan is static (same ops, same buffers for a given frame count): capture its launches once and replay them as one
 * CUDA graph (about a hundred short kernels per decode; the replay removes the per-launch CPU and scheduling gaps). */
struct fq3c_graph {
  cudaGraphExec_t exec;
  int n_kernels;
};
int fq3c_graph_create(const fq3c_op* ops, int n_ops, void** out) {
  if (!ops || !out || n_ops <= 0) return fail("graph: bad arguments");
  cudaStream_t cs;
  if (cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking) != cudaSuccess) return fail("graph: stream create failed");
  const int64_t before = t_launches;
  cudaGraph_t g = nullptr;
  cudaError_t e = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
  int rc = 0;
  if (e == cudaSuccess) {
    rc = fq3c_run(ops, n_ops, cs);
    e = cudaStreamEndCapture(cs, &g);
  }
  const int n_kernels = (int)(t_launches - before);
  count_launches(-(int64_t)n_kernels);  // nothing ran yet
  cudaStreamDestroy(cs);
  if (rc != 0 || e != cudaSuccess || !g) {
    if (g) cudaGraphDestroy(g);
    return rc != 0 ? rc : fail(std::string("graph: capture failed: ") + cudaGetErrorString(e));
  }
  cudaGraphExec_t ex = nullptr;
  e = cudaGraphInstantiate(&ex, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) return fail(std::string("graph: instantiate failed: ") + cudaGetErrorString(e));
  *out = new fq3c_graph{ex, n_kernels};
  return 0;
}
int fq3c_graph_launch(void* graph, void* stream) {
  fq3c_graph* g = reinterpret_cast<fq3c_graph*>(graph);
  if (!g) return fail("graph: null handle");
  const cudaError_t e = cudaGraphLaunch(g->exec, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail(std::string("graph: launch failed: ") + cudaGetErrorString(e));
  count_launches(g->n_kernels);
  return 0;
}
int fq3c_graph_destroy(void* graph) {
  fq3c_graph* g = reinterpret_cast<fq3c_graph*>(graph);
  if (g) {
    cudaGraphExecDestroy(g->exec);
    delete g;
  }
  return 0;
}

}  // extern "C"
