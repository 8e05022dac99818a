// Hardware fact-finding microbenchmarks that drive the kernel design (not product code).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scratch/micro scratch/micro.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x;
}
__device__ __forceinline__ uint64_t splitmix(uint64_t seed, uint64_t i) {
  uint64_t z = seed + (i + 1) * 0x9e3779b97f4a7c15ULL;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL; z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL; return z ^ (z >> 31);
}

__global__ void fill_rand(uint64_t* p, size_t n, uint64_t seed, uint64_t mask) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) p[i] = splitmix(seed, i) & mask;
}

__global__ void k_read(const ulonglong2* __restrict__ p, size_t n2, uint64_t* out) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  uint64_t acc = 0;
  for (; i + 3 * st < n2; i += 4 * st) {
    ulonglong2 a = p[i], b = p[i + st], c = p[i + 2 * st], d = p[i + 3 * st];
    acc += a.x ^ a.y ^ b.x ^ b.y ^ c.x ^ c.y ^ d.x ^ d.y;
  }
  for (; i < n2; i += st) { ulonglong2 a = p[i]; acc += a.x ^ a.y; }
  if (acc == 0x1234567) out[0] = acc;
}
__global__ void k_copy(const ulonglong2* __restrict__ p, ulonglong2* __restrict__ q, size_t n2) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i + 3 * st < n2; i += 4 * st) {
    ulonglong2 a = p[i], b = p[i + st], c = p[i + 2 * st], d = p[i + 3 * st];
    q[i] = a; q[i + st] = b; q[i + 2 * st] = c; q[i + 3 * st] = d;
  }
  for (; i < n2; i += st) q[i] = p[i];
}

// smem histogram with atomics, bins = 1<<bits
template <int MODE>  // 0: plain smem atomicAdd, 1: match_any aggregated, 2: no atomics (hash only, baseline)
__global__ void k_hist(const ulonglong2* __restrict__ keys, size_t n2, int bits, unsigned* gh) {
  extern __shared__ unsigned sh[];
  int nb = 1 << bits;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  unsigned acc = 0;
  for (; i < n2; i += st) {
    ulonglong2 k = keys[i];
    unsigned b0 = (unsigned)(mix64(k.x) >> (64 - bits)), b1 = (unsigned)(mix64(k.y) >> (64 - bits));
    if (MODE == 0) { atomicAdd(&sh[b0], 1u); atomicAdd(&sh[b1], 1u); }
    else if (MODE == 1) {
      unsigned m0 = __match_any_sync(0xffffffffu, b0); if ((__ffs(m0) - 1) == (threadIdx.x & 31)) atomicAdd(&sh[b0], __popc(m0));
      unsigned m1 = __match_any_sync(0xffffffffu, b1); if ((__ffs(m1) - 1) == (threadIdx.x & 31)) atomicAdd(&sh[b1], __popc(m1));
    } else acc += b0 ^ b1;
  }
  __syncthreads();
  if (MODE == 2) { if (acc == 0x12345) gh[0] = acc; return; }
  for (int j = threadIdx.x; j < nb; j += blockDim.x) if (sh[j]) atomicAdd(&gh[j], sh[j]);
}

// random gather of 16B slots from a table of tslots (power of two) slots
__global__ void k_gather(const ulonglong2* __restrict__ keys, size_t n2, const ulonglong2* __restrict__ tab, uint64_t tmask, uint64_t* out) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  uint64_t acc = 0;
  for (; i < n2; i += st) {
    ulonglong2 k = keys[i];
    ulonglong2 a = tab[mix64(k.x) & tmask], b = tab[mix64(k.y) & tmask];
    acc += a.x + a.y + b.x + b.y;
  }
  if (acc == 0x1234567) out[0] = acc;
}
// random gather of 4B
__global__ void k_gather4(const ulonglong2* __restrict__ keys, size_t n2, const unsigned* __restrict__ tab, uint64_t tmask, uint64_t* out) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  uint64_t acc = 0;
  for (; i < n2; i += st) {
    ulonglong2 k = keys[i];
    acc += tab[mix64(k.x) & tmask] + tab[mix64(k.y) & tmask];
  }
  if (acc == 0x1234567) out[0] = acc;
}

// random global atomics: MODE 0 add u64 (RED), 1 min s64 (RED), 2 add with return, 3 CAS u32
template <int MODE>
__global__ void k_gatomic(const ulonglong2* __restrict__ keys, size_t n2, unsigned long long* tab, uint64_t tmask, uint64_t* out) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  uint64_t acc = 0;
  for (; i < n2; i += st) {
    ulonglong2 k = keys[i];
    uint64_t s0 = mix64(k.x) & tmask, s1 = mix64(k.y) & tmask;
    if (MODE == 0) { atomicAdd(&tab[s0], 1ULL); atomicAdd(&tab[s1], 1ULL); }
    else if (MODE == 1) { atomicMin((long long*)&tab[s0], (long long)k.x); atomicMin((long long*)&tab[s1], (long long)k.y); }
    else if (MODE == 2) { acc += atomicAdd(&tab[s0], 1ULL); acc += atomicAdd(&tab[s1], 1ULL); }
    else { acc += atomicCAS((unsigned*)&tab[s0], 0u, 1u); acc += atomicCAS((unsigned*)&tab[s1], 0u, 1u); }
  }
  if (acc == 0x1234567) out[0] = acc;
}

// smem 64-bit atomics into per-CTA table of G groups (direct index = key % G): sum,count(u32),min,max
template <int MODE>  // 0: 4 atomics/row; 1: sum only; 2: sum+count only
__global__ void k_satomic(const ulonglong2* __restrict__ keys, size_t n2, unsigned G, uint64_t* out) {
  extern __shared__ unsigned long long s64[];
  unsigned long long* ssum = s64; long long* smin = (long long*)(s64 + G); long long* smax = (long long*)(s64 + 2 * G);
  unsigned* scnt = (unsigned*)(s64 + 3 * G);
  for (unsigned i = threadIdx.x; i < G; i += blockDim.x) { ssum[i] = 0; smin[i] = INT64_MAX; smax[i] = INT64_MIN; scnt[i] = 0; }
  __syncthreads();
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n2; i += st) {
    ulonglong2 k = keys[i];
    unsigned g0 = (unsigned)(k.x % G), g1 = (unsigned)(k.y % G);
    long long v0 = (long long)(k.x >> 20), v1 = (long long)(k.y >> 20);
    atomicAdd(&ssum[g0], (unsigned long long)v0); atomicAdd(&ssum[g1], (unsigned long long)v1);
    if (MODE != 1) { atomicAdd(&scnt[g0], 1u); atomicAdd(&scnt[g1], 1u); }
    if (MODE == 0) {
      if (v0 < smin[g0]) atomicMin(&smin[g0], v0); if (v1 < smin[g1]) atomicMin(&smin[g1], v1);
      if (v0 > smax[g0]) atomicMax(&smax[g0], v0); if (v1 > smax[g1]) atomicMax(&smax[g1], v1);
    }
  }
  __syncthreads();
  uint64_t acc = 0;
  for (unsigned j = threadIdx.x; j < G; j += blockDim.x) acc += ssum[j] + scnt[j] + smin[j] + smax[j];
  if (acc == 0x1234567) out[0] = acc;
}

// naive scatter: global atomic cursor per partition + 8B key + 4B id write
__global__ void k_scatter(const uint64_t* __restrict__ keys, size_t n, int bits, unsigned* cur, uint64_t* ok, unsigned* oid, size_t cap) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) {
    uint64_t k = keys[i]; unsigned p = (unsigned)(mix64(k) >> (64 - bits));
    unsigned pos = atomicAdd(&cur[p], 1u);
    size_t o = (size_t)p * cap + pos;
    ok[o] = k; oid[o] = (unsigned)i;
  }
}

template <class F> float timeit(F f, int reps = 5) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) { cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
  CK(cudaGetLastError());
  return best;
}

int main() {
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
  printf("device %s SMs %d L2 %d MB smem/blk optin %zu\n", pr.name, pr.multiProcessorCount, pr.l2CacheSize >> 20, pr.sharedMemPerBlockOptin);
  const size_t N = 1ull << 28;  // 2 GiB of keys
  uint64_t *keys, *out, *buf2; CK(cudaMalloc(&keys, N * 8)); CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&buf2, N * 8));
  fill_rand<<<148 * 8, 256>>>(keys, N, 42, ~0ull);
  CK(cudaDeviceSynchronize());
  int SM = pr.multiProcessorCount;
  for (int bpsm : {2, 4, 8}) for (int th : {256, 512}) {
    float ms = timeit([&] { k_read<<<SM * bpsm, th>>>((ulonglong2*)keys, N / 2, out); });
    printf("read  grid=%dxSM th=%d: %.3f ms  %.1f GB/s\n", bpsm, th, ms, N * 8 / ms / 1e6);
  }
  for (int bpsm : {2, 4, 8}) { float ms = timeit([&] { k_copy<<<SM * bpsm, 256>>>((ulonglong2*)keys, (ulonglong2*)buf2, N / 2); });
    printf("copy  grid=%dxSM: %.3f ms  %.1f GB/s (r+w)\n", bpsm, ms, 2.0 * N * 8 / ms / 1e6); }

  unsigned* gh; CK(cudaMalloc(&gh, 4 << 20)); CK(cudaMemset(gh, 0, 4 << 20));
  CK(cudaFuncSetAttribute(k_hist<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(k_hist<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(k_hist<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int bits : {4, 8, 11, 15}) {
    size_t sm = (size_t)4 << bits; int bpsm = sm > 64 * 1024 ? 1 : 2; int th = 1024 / 1;
    float m0 = timeit([&] { k_hist<0><<<SM * bpsm, th, sm>>>((ulonglong2*)keys, N / 2, bits, gh); });
    float m1 = timeit([&] { k_hist<1><<<SM * bpsm, th, sm>>>((ulonglong2*)keys, N / 2, bits, gh); });
    float m2 = timeit([&] { k_hist<2><<<SM * bpsm, th, sm>>>((ulonglong2*)keys, N / 2, bits, gh); });
    printf("hist bits=%2d: atomics %.3f ms (%.1f Grow/s) match_any %.3f ms (%.1f Grow/s) hash-only %.3f ms (%.1f Grow/s)\n", bits, m0, N / m0 / 1e6, m1, N / m1 / 1e6, m2, N / m2 / 1e6);
  }

  const size_t NP = 1ull << 27;  // probes
  for (size_t tb : {(size_t)1 << 20, (size_t)4 << 20, (size_t)32 << 20, (size_t)96 << 20, (size_t)256 << 20, (size_t)1 << 30, (size_t)4 << 30}) {
    size_t tp = 1; while (tp * 2 <= tb) tp *= 2;  // pow2 bytes
    ulonglong2* tab; CK(cudaMalloc(&tab, tp)); CK(cudaMemset(tab, 1, tp));
    float ms = timeit([&] { k_gather<<<SM * 8, 256>>>((ulonglong2*)keys, NP / 2, tab, tp / 16 - 1, out); });
    float ms4 = timeit([&] { k_gather4<<<SM * 8, 256>>>((ulonglong2*)keys, NP / 2, (unsigned*)tab, tp / 4 - 1, out); });
    printf("gather table=%5zu MB: 16B %.3f ms %.1f Gprobe/s | 4B %.3f ms %.1f Gprobe/s\n", tp >> 20, ms, NP / ms / 1e6, ms4, NP / ms4 / 1e6);
    float a0 = timeit([&] { k_gatomic<0><<<SM * 8, 256>>>((ulonglong2*)keys, NP / 2, (unsigned long long*)tab, tp / 8 - 1, out); });
    float a1 = timeit([&] { k_gatomic<1><<<SM * 8, 256>>>((ulonglong2*)keys, NP / 2, (unsigned long long*)tab, tp / 8 - 1, out); });
    float a2 = timeit([&] { k_gatomic<2><<<SM * 8, 256>>>((ulonglong2*)keys, NP / 2, (unsigned long long*)tab, tp / 8 - 1, out); });
    float a3 = timeit([&] { k_gatomic<3><<<SM * 8, 256>>>((ulonglong2*)keys, NP / 2, (unsigned long long*)tab, tp / 8 - 1, out); });
    printf("gatomic table=%5zu MB: red.add64 %.1f Gop/s  red.min64 %.1f  atom.add64 %.1f  cas32 %.1f\n", tp >> 20, NP / a0 / 1e6, NP / a1 / 1e6, NP / a2 / 1e6, NP / a3 / 1e6);
    CK(cudaFree(tab));
  }
  {  // tiny tables: contention
    unsigned long long* tab; CK(cudaMalloc(&tab, 1 << 20)); CK(cudaMemset(tab, 0, 1 << 20));
    for (uint64_t g : {1ull, 16ull, 128ull, 1024ull, 16384ull}) {
      float a0 = timeit([&] { k_gatomic<0><<<SM * 8, 256>>>((ulonglong2*)keys, NP / 2, tab, g - 1, out); });
      float a2 = timeit([&] { k_gatomic<2><<<SM * 8, 256>>>((ulonglong2*)keys, NP / 2, tab, g - 1, out); });
      printf("gatomic slots=%6llu: red.add64 %.2f Gop/s  atom.add64 %.2f Gop/s\n", (unsigned long long)g, NP / a0 / 1e6, NP / a2 / 1e6);
    }
    CK(cudaFree(tab));
  }
  CK(cudaFuncSetAttribute(k_satomic<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(k_satomic<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(k_satomic<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (unsigned G : {4u, 100u, 1000u, 4096u}) {
    size_t sm = (size_t)G * 28 + 64; int bpsm = 2;
    float m0 = timeit([&] { k_satomic<0><<<SM * bpsm, 1024, sm>>>((ulonglong2*)keys, N / 2, G, out); });
    float m1 = timeit([&] { k_satomic<1><<<SM * bpsm, 1024, sm>>>((ulonglong2*)keys, N / 2, G, out); });
    float m2 = timeit([&] { k_satomic<2><<<SM * bpsm, 1024, sm>>>((ulonglong2*)keys, N / 2, G, out); });
    printf("smem-agg G=%5u: sum/cnt/min/max %.1f Grow/s | sum only %.1f | sum+cnt %.1f\n", G, N / m0 / 1e6, N / m1 / 1e6, N / m2 / 1e6);
  }
  {
    unsigned* cur; CK(cudaMalloc(&cur, 4 << 20));
    uint64_t* ok; unsigned* oid; size_t n = 1ull << 27;
    for (int bits : {8, 11, 15}) {
      size_t cap = (n >> bits) * 5 / 4 + 1024; CK(cudaMalloc(&ok, (cap << bits) * 8)); CK(cudaMalloc(&oid, (cap << bits) * 4));
      float ms = timeit([&] { cudaMemsetAsync(cur, 0, 4 << 20); k_scatter<<<SM * 8, 256>>>(keys, n, bits, cur, ok, oid, cap); }, 3);
      printf("naive scatter bits=%2d: %.3f ms %.1f Grow/s\n", bits, ms, n / ms / 1e6);
      CK(cudaFree(ok)); CK(cudaFree(oid));
    }
  }
  printf("done\n");
  return 0;
}
