// Per-SM load latency to a set of addresses: is an 8 KB exchange buffer "nearer" to some SMs than to others (two dies, two L2
// partitions)?  One thread per CTA (one CTA per SM) runs a dependent chain of ld.relaxed.gpu on one address and reports the
// average round trip; addresses are spaced 256 B .. 2 MB apart.  Output: for each address the latency seen from every SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_home l2_home.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
__global__ void probe(unsigned long long* buf, const size_t* offs, int n_addr, int iters, unsigned* out, unsigned* smid) {
  if (threadIdx.x != 0) return;
  unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
  smid[blockIdx.x] = sm;
  for (int a = 0; a < n_addr; ++a) {
    unsigned long long* p = buf + offs[a] / 8;
    unsigned long long v = 0;
    // warm
    for (int i = 0; i < 4; ++i) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p + v) : "memory");
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p + v) : "memory");
    long long t1 = clock64();
    out[(size_t)a * gridDim.x + blockIdx.x] = (unsigned)((t1 - t0) / iters) + (unsigned)v;
  }
}
int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t bytes = 64ull << 20;
  unsigned long long* buf; cudaMalloc(&buf, bytes); cudaMemset(buf, 0, bytes);
  std::vector<size_t> offs;
  for (int i = 0; i < 16; ++i) offs.push_back((size_t)i * 256);            // within 4 KB
  for (int i = 1; i < 16; ++i) offs.push_back((size_t)i * 4096);           // within 64 KB
  for (int i = 1; i < 16; ++i) offs.push_back((size_t)i * (64 << 10));     // within 1 MB
  for (int i = 1; i < 16; ++i) offs.push_back((size_t)i * (2 << 20));      // 2 MB pages
  const int n = (int)offs.size();
  size_t* d_offs; cudaMalloc(&d_offs, n * sizeof(size_t)); cudaMemcpy(d_offs, offs.data(), n * sizeof(size_t), cudaMemcpyHostToDevice);
  unsigned *d_out, *d_sm; cudaMalloc(&d_out, (size_t)n * sms * 4); cudaMalloc(&d_sm, sms * 4);
  probe<<<sms, 32>>>(buf, d_offs, n, 64, d_out, d_sm);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed\n"); return 1; }
  std::vector<unsigned> out((size_t)n * sms), sm(sms);
  cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(sm.data(), d_sm, sms * 4, cudaMemcpyDeviceToHost);
  printf("%d SMs; latency in cycles of a dependent ld.relaxed.gpu chain (all CTAs probing the same address at the same time)\n", sms);
  for (int a = 0; a < n; ++a) {
    std::vector<unsigned> v(out.begin() + (size_t)a * sms, out.begin() + (size_t)(a + 1) * sms);
    std::vector<unsigned> s = v; std::sort(s.begin(), s.end());
    int near = 0; for (unsigned x : v) near += x < (s[0] + s[sms - 1]) / 2;
    printf("off %9zu: min %4u  p25 %4u  median %4u  p75 %4u  max %4u   SMs below the midpoint: %3d   first 16 CTAs:", offs[a], s[0], s[sms / 4], s[sms / 2],
           s[3 * sms / 4], s[sms - 1], near);
    for (int c = 0; c < 16; ++c) printf(" %u", v[c]);
    printf("\n");
  }
  // which CTAs are "near" for offset 0 vs for the 2 MB-apart offsets: a signature string per address
  for (int a : {0, 1, 16, 31, 46, 47, 48}) {
    std::vector<unsigned> v(out.begin() + (size_t)a * sms, out.begin() + (size_t)(a + 1) * sms);
    std::vector<unsigned> s = v; std::sort(s.begin(), s.end());
    const unsigned mid = (s[0] + s[sms - 1]) / 2;
    printf("off %9zu near-map by CTA: ", offs[a]);
    for (int c = 0; c < sms; ++c) putchar(v[c] < mid ? 'n' : 'F');
    printf("\n");
  }
  return 0;
}
