// fq3_codec.cu — kernels of the 12 Hz speech-tokenizer decoder (include/fq3_codec.h).
//
// Activations are channels-last bf16 [time, channels]; every convolution is an implicit GEMM over (tap, c_in)
// whose A rows are time-shifted views of the same buffer (no im2col), so causal padding is a bounds check.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <string>

#include "../../include/fq3_codec.h"

typedef __nv_bfloat16 bf16;

namespace {

thread_local std::string g_err;
int64_t g_launches = 0;

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

// Implicit-GEMM convolution / linear layer on the bf16 tensor cores (mma.sync m16n8k16, fp32 accumulate).
__global__ void __launch_bounds__(128) fq3c_gemm_kernel(const fq3c_op o) {
  __shared__ __align__(16) bf16 As[2][BM][LDS];
  __shared__ __align__(16) bf16 Bs[2][BN][LDS];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp >> 1, wn = warp & 1;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
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

// RVQ dequantiser front half: gather + sum codebook rows. C[m, 0:dim] = sum of the first i0 groups, C[m, dim:2dim] = rest.
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
  bf16* row = reinterpret_cast<bf16*>(const_cast<void*>(o.A)) + (size_t)t * o.lda + col0;
  for (int idx = threadIdx.x; idx < nheads * half; idx += blockDim.x) {
    const int h = idx / half, i = idx - h * half;
    const float inv = powf(o.f0, -2.0f * i / d);
    const float ang = (float)t * inv;
    const float c = bf16r(cosf(ang)), s = bf16r(sinf(ang));
    bf16* p = row + h * d;
    const float a = __bfloat162float(p[i]), b = __bfloat162float(p[i + half]);
    p[i] = __float2bfloat16_rn(bf16r(a * c) + bf16r(-b * s));
    p[i + half] = __float2bfloat16_rn(bf16r(b * c) + bf16r(a * s));
  }
}

// Sliding-window causal attention, one warp per (query, head); head_dim <= 128.
__global__ void fq3c_attn_kernel(const fq3c_op o) {
  const int nh = o.i0, nkv = o.i1, d = o.i2, win = o.K, T = o.M;
  const int wpb = blockDim.x >> 5;
  const int gw = blockIdx.x * wpb + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (gw >= T * nh) return;
  const int t = gw / nh, h = gw - t * nh, kvh = h / (nh / nkv);
  const bf16* base = reinterpret_cast<const bf16*>(o.A);
  const int qoff = h * d, koff = nh * d + kvh * d, voff = nh * d + nkv * d + kvh * d;
  const int per = (d + 31) / 32;  // <= 4
  float q[4], acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = 0; i < per; ++i) {
    const int c = lane + i * 32;
    q[i] = c < d ? __bfloat162float(base[(size_t)t * o.lda + qoff + c]) : 0.f;
  }
  const float scale = rsqrtf((float)d);
  float m = -INFINITY, l = 0.f;
  const int j0 = max(0, t - win + 1);
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
  bf16* out = reinterpret_cast<bf16*>(o.C) + (size_t)t * o.ldc + h * d;
  for (int i = 0; i < per; ++i) {
    const int c = lane + i * 32;
    if (c < d) out[c] = __float2bfloat16_rn(acc[i] / l);
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
      const int r = t - (o.taps - 1) + j;
      if (r >= 0) a += w[c * o.taps + j] * __bfloat162float(x[(size_t)r * o.lda + c]);
    }
    y[(size_t)t * o.ldc + c] = __float2bfloat16_rn(a);
  }
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

}  // namespace

extern "C" {

int fq3c_abi_version(void) { return 1; }
const char* fq3c_last_error(void) { return g_err.c_str(); }
int64_t fq3c_launch_count(void) { return g_launches; }

int fq3c_run(const fq3c_op* ops, int n_ops, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  for (int i = 0; i < n_ops; ++i) {
    const fq3c_op& o = ops[i];
    if (o.M <= 0) continue;
    switch (o.kind) {
      case FQ3C_GEMM: {
        if (o.K % 8 || o.cin % 8 || o.taps > 8 || o.col_mod <= 0) return fail("gemm: K and cin must be multiples of 8, taps <= 8");
        dim3 grid((o.N + BN - 1) / BN, (o.M + BM - 1) / BM);
        fq3c_gemm_kernel<<<grid, 128, 0, s>>>(o);
        break;
      }
      case FQ3C_RVQ: fq3c_rvq_kernel<<<o.M, 128, 0, s>>>(o); break;
      case FQ3C_RMSNORM: fq3c_rmsnorm_kernel<<<o.M, 256, 0, s>>>(o); break;
      case FQ3C_LAYERNORM: fq3c_layernorm_kernel<<<o.M, 256, 0, s>>>(o); break;
      case FQ3C_ROPE: fq3c_rope_kernel<<<o.M, 256, 0, s>>>(o); break;
      case FQ3C_ATTN: {
        if (o.i2 > 128) return fail("attn: head_dim > 128");
        const int warps = o.M * o.i0;
        fq3c_attn_kernel<<<(warps + 7) / 8, 256, 0, s>>>(o);
        break;
      }
      case FQ3C_DWCONV: {
        const size_t n = (size_t)o.M * o.N;
        fq3c_dwconv_kernel<<<(unsigned)std::min<size_t>(2048, (n + 255) / 256), 256, 0, s>>>(o);
        break;
      }
      case FQ3C_SNAKE: {
        const size_t n = (size_t)o.M * o.N;
        fq3c_snake_kernel<<<(unsigned)std::min<size_t>(4096, (n + 255) / 256), 256, 0, s>>>(o);
        break;
      }
      default: return fail("unknown op kind");
    }
    g_launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(std::string("launch failed at op ") + std::to_string(i) + ": " + cudaGetErrorString(e));
  }
  return 0;
}

}  // extern "C"
