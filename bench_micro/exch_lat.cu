// Microbenchmark: latency floor of one "phase" of the persistent decode kernel on B200.
//   A phase = every CTA publishes its slice of an N-element activation vector as LL words {payload, epoch} (8 bytes,
//   single store => atomic), then every CTA poll-reads the WHOLE vector into shared memory, block-barriers, and
//   (optionally) does a GEMV-sized amount of shared-memory math before publishing the next vector.
// Variants: values per word (1 fp32 | 2 bf16), 8- or 16-byte polling loads, back-off sleep, a background TMA weight
// stream per CTA (models the producer warp), and the amount of per-phase math.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exch_lat exch_lat.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mb_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ bool mb_try(uint64_t* b, uint32_t ph) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(s32(b)), "r"(ph) : "memory");
  return ok;
}
__device__ __forceinline__ void bulk(void* d, const void* s, uint32_t n, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(d)), "l"(s), "r"(n), "r"(s32(b)) : "memory");
}
__device__ __forceinline__ uint2 ll_ld(const uint2* p) {
  uint2 w;
  asm volatile("ld.relaxed.gpu.global.v2.u32 {%0,%1}, [%2];" : "=r"(w.x), "=r"(w.y) : "l"(p) : "memory");
  return w;
}
__device__ __forceinline__ uint4 ll_ld2(const uint2* p) {
  uint4 w;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w) : "l"(p) : "memory");
  return w;
}
__device__ __forceinline__ void ll_st(uint2* p, uint32_t v, uint32_t ep) {
  asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v), "r"(ep) : "memory");
}
__device__ __forceinline__ float warp_sum(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct Params {
  uint2* buf[4];      // rotating LL buffers (publish-before-read lets a CTA run two phases ahead of the slowest reader)
  int nwords;         // words per vector
  int iters;
  int sleep_ns;       // 0 = tight spin
  int wide;           // 1: 16-byte polling loads (two words)
  int work_rows;      // rows per warp of fake GEMV math per phase (0 = none)
  int work_k;         // K of the fake GEMV (elements, multiple of 256)
  int stream;         // 1: background TMA stream per CTA
  const uint8_t* wbase;
  size_t wbytes_per_cta;
  int stage_bytes, n_stages;
  long long* cyc;     // per-CTA total cycles
  float* sink;
};

constexpr int CT = 480;  // consumer threads

__global__ void __launch_bounds__(512, 1) k_exch(const Params p) {
  extern __shared__ __align__(128) uint8_t sm[];
  uint64_t* full = (uint64_t*)sm;            // [16]
  float* xs = (float*)(sm + 256);            // up to 8192 floats
  uint8_t* wsm = sm + 256 + 32768;           // fake weights (work) 16 KB
  uint8_t* ring = wsm + 16384;
  const int tid = threadIdx.x, G = gridDim.x, cta = blockIdx.x;
  if (tid == 0) { *(volatile int*)(sm + 128) = 0; for (int i = 0; i < 16; ++i) mb_init(&full[i], 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
  for (int i = tid; i < 4096; i += 512) ((float*)wsm)[i] = 0.001f * (i & 31);
  __syncthreads();
  if (tid >= CT) {
    // background weight stream: one thread keeps n_stages bulk copies in flight and recycles them itself
    if (tid == CT && p.stream) {
      const uint8_t* src = p.wbase + (size_t)cta * p.wbytes_per_cta;
      const size_t ntiles = p.wbytes_per_cta / p.stage_bytes;
      volatile int* stop = (volatile int*)(sm + 128);
      size_t t = 0;
      for (; t < (size_t)p.n_stages && t < ntiles; ++t) { mb_expect(&full[t], p.stage_bytes); bulk(ring + t * p.stage_bytes, src + t * p.stage_bytes, p.stage_bytes, &full[t]); }
      size_t done = 0;
      while (!*stop) {
        const int s = done % p.n_stages; const uint32_t ph = (done / p.n_stages) & 1;
        while (!mb_try(&full[s], ph)) {}
        ++done;
        const size_t nt = (t % ntiles);
        mb_expect(&full[s], p.stage_bytes);
        bulk(ring + (size_t)s * p.stage_bytes, src + nt * p.stage_bytes, p.stage_bytes, &full[s]);
        ++t;
      }
      // drain
      for (size_t d = done; d < t; ++d) { const int s = d % p.n_stages; const uint32_t ph = (d / p.n_stages) & 1; while (!mb_try(&full[s], ph)) {} }
      p.cyc[G + cta] = (long long)t;
    }
    return;
  }
  const int warp = tid >> 5, lane = tid & 31;
  const int N = p.nwords;
  // this CTA's slice of the vector
  const int base = N / G, rem = N - base * G;
  const int w0 = cta * base + min(cta, rem), w1 = w0 + base + (cta < rem ? 1 : 0);
  float carry = 1.0f;
  long long t0 = clock64();
  for (int it = 1; it <= p.iters; ++it) {
    uint2* out = p.buf[it & 3];
    const uint2* in = p.buf[(it + 3) & 3];
    // ---- publish (one warp per word, lane 0 stores: like the GEMV epilogue)
    for (int w = w0 + warp; w < w1; w += 15) if (lane == 0) ll_st(out + w, __float_as_uint(carry), (uint32_t)it);
    // ---- poll-read the previous vector (epoch it-1; iteration 1 reads zeros => skip check)
    const uint32_t want = (uint32_t)(it - 1);
    if (p.wide) {
      for (int k = 2 * tid; k < N; k += 2 * CT) {
        uint4 v = ll_ld2(in + k);
        if (want) { unsigned n = 0; while (v.y != want || v.w != want) { if (++n > (1u << 22)) { p.sink[G] = 1.f; break; } if (p.sleep_ns) __nanosleep(p.sleep_ns); v = ll_ld2(in + k); } }
        xs[k] = __uint_as_float(v.x); xs[k + 1] = __uint_as_float(v.z);
      }
    } else {
      for (int k = tid; k < N; k += CT) {
        uint2 v = ll_ld(in + k);
        if (want) { unsigned n = 0; while (v.y != want) { if (++n > (1u << 22)) { p.sink[G] = 1.f; break; } if (p.sleep_ns) __nanosleep(p.sleep_ns); v = ll_ld(in + k); } }
        xs[k] = __uint_as_float(v.x);
      }
    }
    asm volatile("bar.sync 1, 480;" ::: "memory");
    // ---- fake GEMV: work_rows rows per warp over work_k elements from smem (weights smem-resident)
    float acc = 0.f;
    for (int r = 0; r < p.work_rows; ++r) {
      float a0 = 0.f, a1 = 0.f;
      for (int c = lane; c < p.work_k / 8; c += 32) {
        const float4 wv = ((const float4*)wsm)[(c + r * 7) & 1023];
        const float4 x0 = ((const float4*)xs)[(2 * c) & 2047], x1 = ((const float4*)xs)[(2 * c + 1) & 2047];
        a0 += wv.x * x0.x + wv.y * x0.y + wv.z * x0.z + wv.w * x0.w;
        a1 += wv.x * x1.x + wv.y * x1.y + wv.z * x1.z + wv.w * x1.w;
      }
      acc += warp_sum(a0 + a1);
    }
    carry = acc * 1e-9f + 1.0f;
    asm volatile("bar.sync 1, 480;" ::: "memory");  // xs reuse
  }
  long long t1 = clock64();
  if (tid == 0) { p.cyc[cta] = t1 - t0; *(volatile int*)(sm + 128) = 1; p.sink[cta] = carry; }
}

// L2 hit latency by pointer chase through a small buffer with gpu-scope relaxed loads
__global__ void k_chase(const uint32_t* buf, int n, long long* out) {
  uint32_t i = 0;
  long long t0 = clock64();
  for (int k = 0; k < n; ++k) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(i) : "l"(buf + i) : "memory");
  long long t1 = clock64();
  out[0] = (t1 - t0) / n; out[1] = i;
}

int main(int argc, char** argv) {
  setvbuf(stdout, nullptr, _IONBF, 0);
  int dev = 0; CK(cudaSetDevice(dev));
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, dev));
  const int G = pr.multiProcessorCount;
  int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev));
  printf("%s, %d SMs, max clock %d MHz\n", pr.name, G, clk_khz / 1000);
  // chase
  {
    const int n = 1 << 16; uint32_t* h = (uint32_t*)malloc(n * 4);
    for (int i = 0; i < n; ++i) h[i] = (uint32_t)((i * 9973u + 4099u) % n);
    uint32_t* d; long long* o; CK(cudaMalloc(&d, n * 4)); CK(cudaMalloc(&o, 16));
    CK(cudaMemcpy(d, h, n * 4, cudaMemcpyHostToDevice));
    k_chase<<<1, 1>>>(d, 2000, o); k_chase<<<1, 1>>>(d, 20000, o); CK(cudaDeviceSynchronize());
    long long r[2]; CK(cudaMemcpy(r, o, 16, cudaMemcpyDeviceToHost));
    printf("L2 chase latency (ld.relaxed.gpu): %lld cycles\n", r[0]);
  }
  const size_t wbytes = (size_t)2 << 30;  // 2 GiB stream source
  uint8_t* wb; CK(cudaMalloc(&wb, wbytes)); CK(cudaMemset(wb, 1, wbytes));
  uint2 *b0, *b1, *b2, *b3; CK(cudaMalloc(&b0, 65536 * 8)); CK(cudaMalloc(&b1, 65536 * 8)); CK(cudaMalloc(&b2, 65536 * 8)); CK(cudaMalloc(&b3, 65536 * 8));
  long long* cyc; CK(cudaMalloc(&cyc, 8 * 2 * G)); float* sink; CK(cudaMalloc(&sink, 4 * (G + 1)));
  const int smem = 256 + 32768 + 16384 + 8 * 16384;
  CK(cudaFuncSetAttribute(k_exch, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  struct Cfg { int nwords, sleep, wide, rows, k, stream; };
  const Cfg cfgs[] = {
      {1024, 0, 0, 0, 0, 0},   {1024, 0, 1, 0, 0, 0},   {1024, 64, 1, 0, 0, 0},  {1024, 20, 1, 0, 0, 0},
      {512, 0, 1, 0, 0, 0},    {2048, 0, 1, 0, 0, 0},   {3072, 0, 1, 0, 0, 0},   {4096, 0, 1, 0, 0, 0},
      {1024, 0, 1, 2, 1024, 0}, {1024, 0, 1, 3, 1024, 0}, {3072, 0, 1, 1, 3072, 0},
      {1024, 0, 1, 0, 0, 1},   {1024, 64, 1, 0, 0, 1},  {3072, 0, 1, 0, 0, 1},   {1024, 0, 1, 2, 1024, 1}, {1024, 0, 0, 2, 1024, 1},
  };
  for (const Cfg& c : cfgs) {
    Params p{};
    p.buf[0] = b0; p.buf[1] = b1; p.buf[2] = b2; p.buf[3] = b3; p.nwords = c.nwords; p.iters = 2000; p.sleep_ns = c.sleep; p.wide = c.wide;
    p.work_rows = c.rows; p.work_k = c.k; p.stream = c.stream; p.wbase = wb; p.wbytes_per_cta = wbytes / G / 16384 * 16384;
    p.stage_bytes = 16384; p.n_stages = 8; p.cyc = cyc; p.sink = sink;
    CK(cudaMemset(b0, 0, 65536 * 8)); CK(cudaMemset(b1, 0, 65536 * 8)); CK(cudaMemset(b2, 0, 65536 * 8)); CK(cudaMemset(b3, 0, 65536 * 8)); CK(cudaMemset(cyc, 0, 8 * 2 * G)); CK(cudaMemset(sink, 0, 4 * (G + 1)));
    void* args[] = {&p};
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    CK(cudaLaunchCooperativeKernel((void*)k_exch, dim3(G), dim3(512), args, smem, 0));
    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    long long* h = (long long*)malloc(8 * 2 * G); CK(cudaMemcpy(h, cyc, 8 * 2 * G, cudaMemcpyDeviceToHost));
    long long mx = 0, tiles = 0; for (int i = 0; i < G; ++i) { if (h[i] > mx) mx = h[i]; tiles += h[G + i]; }
    printf("N=%5d sleep=%3d wide=%d work=%dx%-5d stream=%d : %7.3f us/phase (%lld cycles/phase)  stream %.0f GB/s\n", c.nwords, c.sleep, c.wide,
           c.rows, c.k, c.stream, ms * 1000.0 / p.iters, mx / p.iters, c.stream ? tiles * 16384.0 / (ms * 1e-3) / 1e9 : 0.0);
    float bad = 0; CK(cudaMemcpy(&bad, sink + G, 4, cudaMemcpyDeviceToHost)); if (bad != 0.f) printf("   ^^ SPIN LIMIT HIT (protocol bug)\n");
    free(h);
  }
  return 0;
}
