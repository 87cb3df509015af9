// Microbenchmark: how fast can one persistent CTA per SM pull bytes out of HBM on B200?
//   mode 0: cp.async.bulk (TMA bulk) ring, P producer threads each with its own sub-ring, consumers only wait/release
//   mode 1: LDG.128 streaming, W warps, unroll U (xor-reduced so loads are not dead)
//   mode 2: TMA ring + consumer warps that actually read the tile from smem (LDS.128 + xor)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_bw stream_bw.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mb_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mb_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mb_wait(uint64_t* b, uint32_t ph) {
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(s32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void bulk(void* d, const void* s, uint32_t n, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(d)), "l"(s), "r"(n), "r"(s32(b)) : "memory");
}

__device__ __forceinline__ void bulk_hint(void* d, const void* s, uint32_t n, uint64_t* b, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(s32(d)), "l"(s), "r"(n), "r"(s32(b)), "l"(pol) : "memory");
}
// kernel-like pattern: the buffer is a sequence of "matrices" of G*chunk bytes; CTA c reads bytes [c*chunk, (c+1)*chunk) of
// each matrix in turn, in tiles of <= S bytes (last tile of a chunk partial).  hint: 0 none, 1 evict_first, 2 evict_last.
// CW consumer warps each arrive on every tile (like the decode kernel).
__global__ void __launch_bounds__(512, 1) k_chunk(const uint8_t* __restrict__ base, int nmat, int chunk, int S, int NS, int CW, int hint, unsigned* sink) {
  extern __shared__ __align__(128) uint8_t sm[];
  uint64_t* full = (uint64_t*)sm;
  uint64_t* empty = full + 64;
  uint8_t* ring = sm + 1024;
  const int tid = threadIdx.x;
  if (tid == 0) { for (int i = 0; i < NS; ++i) { mb_init(&full[i], 1); mb_init(&empty[i], CW); } asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncthreads();
  const size_t mat_bytes = (size_t)gridDim.x * chunk;
  const int tpc = (chunk + S - 1) / S;
  if (tid == CW * 32) {
    uint64_t pol = 0;
    if (hint == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    if (hint == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    int t = 0;
    for (int m = 0; m < nmat; ++m) {
      const uint8_t* src = base + (size_t)m * mat_bytes + (size_t)blockIdx.x * chunk;
      for (int j = 0; j < tpc; ++j, ++t) {
        const int s = t % NS; const uint32_t ph = ((t / NS) & 1) ^ 1;
        const uint32_t n = min(S, chunk - j * S);
        mb_wait(&empty[s], ph);
        mb_expect(&full[s], n);
        if (hint) bulk_hint(ring + (size_t)s * S, src + (size_t)j * S, n, &full[s], pol);
        else bulk(ring + (size_t)s * S, src + (size_t)j * S, n, &full[s]);
      }
    }
  } else if (tid < CW * 32) {
    const int l = tid & 31;
    const int total = nmat * tpc;
    for (int t = 0; t < total; ++t) {
      const int s = t % NS; const uint32_t ph = (t / NS) & 1;
      if (l == 0) mb_wait(&full[s], ph);
      __syncwarp();
      if (l == 0) mb_arrive(&empty[s]);
    }
  }
}

// each CTA streams `per_cta` bytes starting at base + cta*per_cta (contiguous slice), in tiles of S bytes
template <int MODE>
__global__ void __launch_bounds__(512, 1) k_tma(const uint8_t* __restrict__ base, size_t per_cta, int S, int NS, int P, int CW, unsigned* sink) {
  extern __shared__ __align__(128) uint8_t sm[];
  uint64_t* full = (uint64_t*)sm;
  uint64_t* empty = full + 64;
  uint8_t* ring = sm + 1024;
  const int tid = threadIdx.x;
  const int cons_threads = CW * 32;
  if (tid == 0) { for (int i = 0; i < NS; ++i) { mb_init(&full[i], 1); mb_init(&empty[i], CW); } asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncthreads();
  const uint8_t* src = base + (size_t)blockIdx.x * per_cta;
  const int ntiles = (int)(per_cta / S);
  if (tid >= cons_threads) {
    const int p = tid - cons_threads;   // producer p handles stages s with s % P == p
    if (p < P) {
      for (int t = p; t < ntiles; t += P) {
        const int s = t % NS; const uint32_t ph = ((t / NS) & 1) ^ 1;
        mb_wait(&empty[s], ph);
        mb_expect(&full[s], S);
        bulk(ring + (size_t)s * S, src + (size_t)t * S, S, &full[s]);
      }
    }
  } else {
    unsigned acc = 0;
    const int w = tid >> 5, l = tid & 31;
    for (int t = 0; t < ntiles; ++t) {
      const int s = t % NS; const uint32_t ph = (t / NS) & 1;
      mb_wait(&full[s], ph);
      if (MODE == 2) {
        const uint4* tile = (const uint4*)(ring + (size_t)s * S);
        for (int c = w * 32 + l; c < S / 16; c += cons_threads) { uint4 v = tile[c]; acc ^= v.x ^ v.y ^ v.z ^ v.w; }
      }
      __syncwarp();
      if (l == 0) mb_arrive(&empty[s]);
    }
    if (acc == 0x12345) sink[0] = acc;
  }
}

__global__ void __launch_bounds__(1024, 1) k_ldg(const uint4* __restrict__ base, size_t per_cta16, int U, unsigned* sink) {
  const uint4* src = base + (size_t)blockIdx.x * per_cta16;
  unsigned acc = 0;
  const size_t n = per_cta16;
  size_t i = threadIdx.x;
  const size_t stride = blockDim.x;
  for (; i + (size_t)(U - 1) * stride < n; i += (size_t)U * stride) {
    uint4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) if (u < U) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(src + i + u * stride));
#pragma unroll
    for (int u = 0; u < 8; ++u) if (u < U) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  if (acc == 0x12345) sink[0] = acc;
}

int main(int argc, char** argv) {
  int dev = 0; CK(cudaSetDevice(dev));
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t total = (size_t)2 << 30;  // 2 GiB buffer (>> L2)
  uint8_t* buf; CK(cudaMalloc(&buf, total)); CK(cudaMemset(buf, 1, total));
  unsigned* sink; CK(cudaMalloc(&sink, 4));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  CK(cudaFuncSetAttribute(k_tma<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  CK(cudaFuncSetAttribute(k_tma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  auto run = [&](const char* name, auto launch, size_t bytes) {
    for (int i = 0; i < 2; ++i) launch();
    CK(cudaDeviceSynchronize());
    float best = 1e9;
    for (int r = 0; r < 5; ++r) { cudaEventRecord(a); launch(); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    printf("%-44s %8.1f GB/s  (%.3f ms)\n", name, bytes / best / 1e6, best); fflush(stdout);
  };
  char nm[128];
  int Ss[] = {4096, 8192, 16384, 32768};
  for (int mode : {0, 2}) for (int S : Ss) for (int P : {1, 2, 4}) {
    const int NS = std::min(48, (200 * 1024) / S);
    const int CW = mode == 2 ? 8 : 1;
    size_t per = (total / sms) / S * S;
    snprintf(nm, sizeof nm, "tma mode=%d S=%dK NS=%d P=%d CW=%d", mode, S / 1024, NS, P, CW);
    const size_t smem = 1024 + (size_t)NS * S;
    if (mode == 0) run(nm, [&] { k_tma<0><<<sms, CW * 32 + 32 * ((P + 31) / 32), smem>>>(buf, per, S, NS, P, CW, sink); }, per * sms);
    else run(nm, [&] { k_tma<2><<<sms, CW * 32 + 32, smem>>>(buf, per, S, NS, P, CW, sink); }, per * sms);
  }
  CK(cudaFuncSetAttribute(k_chunk, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  for (int hint : {0, 1, 2}) for (int chunk : {28672, 57344, 86016, 1 << 20}) for (int CW : {1, 15}) {
    const int S = 16384, NS = 9;
    const int nmat = (int)(total / ((size_t)sms * chunk));
    snprintf(nm, sizeof nm, "chunk pattern hint=%d chunk=%dK CW=%d", hint, chunk / 1024, CW);
    run(nm, [&] { k_chunk<<<sms, 512, 1024 + (size_t)NS * S>>>(buf, nmat, chunk, S, NS, CW, hint, sink); }, (size_t)nmat * sms * chunk);
  }
  for (int NSx : {2, 4, 6, 12}) {   // in-flight sweep at S=16K
    const int S = 16384; size_t per = (total / sms) / S * S;
    snprintf(nm, sizeof nm, "tma mode=0 S=16K NS=%d P=1 (in-flight sweep)", NSx);
    run(nm, [&] { k_tma<0><<<sms, 64, 1024 + (size_t)NSx * S>>>(buf, per, S, NSx, 1, 1, sink); }, per * sms);
  }
  for (int T : {256, 512, 1024}) for (int U : {2, 4, 8}) {
    size_t per16 = (total / sms / 16) / ((size_t)T * U) * ((size_t)T * U);
    snprintf(nm, sizeof nm, "ldg.128 threads=%d unroll=%d", T, U);
    run(nm, [&] { k_ldg<<<sms, T>>>((const uint4*)buf, per16, U, sink); }, per16 * 16 * sms);
  }
  for (int blocks : {2, 4}) {  // more CTAs per SM for LDG
    int T = 512, U = 4; size_t per16 = (total / (sms * blocks) / 16) / ((size_t)T * U) * ((size_t)T * U);
    snprintf(nm, sizeof nm, "ldg.128 threads=%d unroll=%d ctas/sm=%d", T, U, blocks);
    run(nm, [&] { k_ldg<<<sms * blocks, T>>>((const uint4*)buf, per16, U, sink); }, per16 * 16 * sms * blocks);
  }
  return 0;
}
