// Is a freshly cleared + built 16 MB table L2-resident for the following CAS / gather kernels?
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ uint64_t mix64(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }
__global__ void fill(uint64_t* p, size_t n, uint64_t seed) { size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x; for (; i < n; i += st) p[i] = mix64(seed + i); }
__global__ void clear_k(uint4* t, size_t n16) { size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x; for (; i < n16; i += st) t[i] = make_uint4(0, 0, 0, 0); }
__global__ void cas_k(const uint64_t* keys, size_t n, uint32_t* tab, uint64_t nb) {  // one row per thread, CAS on word 2 of a 32 B bucket
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; if (i >= n) return;
  uint64_t k = keys[i]; uint64_t b = __umul64hi(mix64(k), nb);
  for (;;) { uint32_t* w = tab + b * 8; if (atomicCAS(w + 2, 0u, (uint32_t)i + 1) == 0u) { *(uint64_t*)w = k; break; } if (atomicCAS(w + 6, 0u, (uint32_t)i + 1) == 0u) { *(uint64_t*)(w + 4) = k; break; } w[3] = 1; if (++b == nb) b = 0; }
}
template <int MODE>  // 0 plain ld.global.nc 32B ; 1 evict_last 256-bit; 2: two 16B __ldg
__global__ void gather_k(const uint64_t* keys, size_t n, const uint32_t* tab, uint64_t nb, unsigned long long* out) {
  size_t i = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * 4; if (i >= n) return;
  unsigned long long acc = 0;
  uint64_t a[4], b[4], c[4], d[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    uint64_t bk = __umul64hi(mix64(keys[i + j]), nb); const uint32_t* p = tab + bk * 8;
    if (MODE == 1) asm volatile("ld.global.L1::no_allocate.L2::evict_last.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a[j]), "=l"(b[j]), "=l"(c[j]), "=l"(d[j]) : "l"(p));
    else if (MODE == 0) asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a[j]), "=l"(b[j]), "=l"(c[j]), "=l"(d[j]) : "l"(p));
    else { ulonglong2 x = __ldg((const ulonglong2*)p), y = __ldg((const ulonglong2*)p + 1); a[j] = x.x; b[j] = x.y; c[j] = y.x; d[j] = y.y; }
  }
#pragma unroll
  for (int j = 0; j < 4; j++) acc += a[j] + b[j] + c[j] + d[j];
  if (acc == 0x12345) out[0] = acc;
}
__global__ void stream_k(const ulonglong2* p, size_t n2, unsigned long long* out) { size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x; unsigned long long a = 0; for (; i < n2; i += st) { ulonglong2 v = p[i]; a += v.x ^ v.y; } if (a == 0x1234) out[0] = a; }
int main() {
  const size_t NB = 512 * 1024, NP = 2 * 1024 * 1024;
  uint64_t *bk, *pk; unsigned long long* out; uint32_t* tab; ulonglong2* big;
  CK(cudaMalloc(&bk, NB * 8)); CK(cudaMalloc(&pk, NP * 8)); CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&tab, 64 << 20)); CK(cudaMalloc(&big, 1ull << 30));
  fill<<<1024, 256>>>(bk, NB, 1); fill<<<1024, 256>>>(pk, NP, 2); CK(cudaMemset(big, 1, 1ull << 30)); CK(cudaDeviceSynchronize());
  cudaEvent_t ev[8]; for (auto& e : ev) cudaEventCreate(&e);
  auto run = [&](const char* name, int clear_mode, int gmode, bool flush_between, uint64_t nb) {
    for (int rep = 0; rep < 3; rep++) {
      stream_k<<<148 * 8, 256>>>(big, (1ull << 30) / 16, out);  // flush L2 with 1 GB of reads
      cudaEventRecord(ev[0]);
      if (clear_mode == 0) cudaMemsetAsync(tab, 0, nb * 32); else if (clear_mode == 1) clear_k<<<148 * 8, 256>>>((uint4*)tab, nb * 2);
      cudaEventRecord(ev[1]);
      cas_k<<<(NB + 255) / 256, 256>>>(bk, NB, tab, nb);
      cudaEventRecord(ev[2]);
      if (flush_between) stream_k<<<148 * 8, 256>>>(big, (64ull << 20) / 16, out);
      cudaEventRecord(ev[3]);
      if (gmode == 0) gather_k<0><<<(NP / 4 + 255) / 256, 256>>>(pk, NP, tab, nb, out); else if (gmode == 1) gather_k<1><<<(NP / 4 + 255) / 256, 256>>>(pk, NP, tab, nb, out); else gather_k<2><<<(NP / 4 + 255) / 256, 256>>>(pk, NP, tab, nb, out);
      cudaEventRecord(ev[4]);
      if (gmode == 0) gather_k<0><<<(NP / 4 + 255) / 256, 256>>>(pk, NP, tab, nb, out); else if (gmode == 1) gather_k<1><<<(NP / 4 + 255) / 256, 256>>>(pk, NP, tab, nb, out); else gather_k<2><<<(NP / 4 + 255) / 256, 256>>>(pk, NP, tab, nb, out);
      cudaEventRecord(ev[5]);
      CK(cudaDeviceSynchronize());
      float t[5]; for (int i = 0; i < 5; i++) cudaEventElapsedTime(&t[i], ev[i], ev[i + 1]);
      if (rep == 2) { printf("%-46s clear %6.1f us  cas %6.1f us  (stream %5.1f)  gather#1 %6.1f us  gather#2 %6.1f us\n", name, t[0] * 1e3, t[1] * 1e3, t[2] * 1e3, t[3] * 1e3, t[4] * 1e3); fflush(stdout); }
    }
  };
  uint64_t nb = NB;
  run("memsetAsync, gather nc", 0, 0, false, nb);
  run("clear kernel, gather nc", 1, 0, false, nb);
  run("clear kernel, gather evict_last", 1, 1, false, nb);
  run("clear kernel, gather 2x ldg128", 1, 2, false, nb);
  run("clear kernel, 64MB stream between, nc", 1, 0, true, nb);
  run("clear kernel, 64MB stream between, evict_last", 1, 1, true, nb);
  run("clear kernel, table 2x (load .25), nc", 1, 0, false, nb * 2);
  printf("done\n");
  return 0;
}
