// Micro-benchmark behind csrc/djoin.cu's partition pass: how fan-out, tile size, the per-(tile, partition) global atomic and
// the short-run write-out each contribute.  Standalone: nvcc -O3 -gencode arch=compute_100a,code=sm_100a part_bench.cu -o part_bench
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint64_t ld_stream_u64(const void* p) {
  uint64_t v; asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p)); return v;
}
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }

struct Args {
  const uint64_t* in; uint64_t n;          // RAW keys (PACKED = false) or packed rows in coarse regions
  const uint32_t* seg_cur; uint32_t seg_cap, n_seg; int seg_shift;   // PACKED: segment s = in[s * seg_cap ..), rows = min(seg_cur[s], seg_cap); partitions of segment s start at s << seg_shift
  int shift; uint32_t P, cap; uint32_t* cursor; uint64_t* out;
};

// MODE bit 0: skip global atomics (fake positions), bit 1: skip the write-out
template <int THREADS, int R, int KPT, bool PACKED, int MODE>
__global__ void __launch_bounds__(THREADS, (THREADS * R * 8 + THREADS * KPT * 8 <= 110 * 1024 && THREADS <= 512) ? 2 : 1) part_kernel(const __grid_constant__ Args a) {
  constexpr int T = THREADS * R, PP = THREADS * KPT, NW = THREADS / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* stage = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* cnt = reinterpret_cast<uint32_t*>(stage + T);
  uint32_t* gbase = cnt + PP;
  __shared__ uint32_t s_warp[NW];
  __shared__ uint32_t s_total;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < PP; i += THREADS) cnt[i] = 0;
  __syncthreads();
  const uint32_t tiles_per_seg = PACKED ? (a.seg_cap + T - 1) / T : 1;
  const uint32_t tiles = PACKED ? tiles_per_seg * a.n_seg : (uint32_t)((a.n + T - 1) / T);
  for (uint32_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    const uint64_t* src = a.in; uint64_t n = a.n, row0 = (uint64_t)t * T; uint32_t pbase = 0;
    if (PACKED) {
      const uint32_t s = t / tiles_per_seg;
      src = a.in + (uint64_t)s * a.seg_cap; n = min(a.seg_cur[s], a.seg_cap); row0 = (uint64_t)(t % tiles_per_seg) * T; pbase = s << a.seg_shift;
      if (row0 >= n) continue;
    }
    uint32_t d[R], rk[R];
    uint64_t key[R];
#pragma unroll
    for (int k = 0; k < R; k++) { const uint64_t row = row0 + (uint64_t)k * THREADS + tid; key[k] = row < n ? ld_stream_u64(src + row) : ~0ULL; }
#pragma unroll
    for (int k = 0; k < R; k++) {
      const uint64_t row = row0 + (uint64_t)k * THREADS + tid;
      d[k] = PACKED ? (uint32_t)(key[k] >> 32) : (uint32_t)key[k];
      rk[k] = row < n ? atomicAdd(&cnt[(d[k] >> a.shift) - pbase], 1u) : 0xFFFFFFFFu;
    }
    __syncthreads();
    uint32_t c[KPT], got[KPT], sum = 0;
#pragma unroll
    for (int j = 0; j < KPT; j++) c[j] = cnt[tid * KPT + j];
#pragma unroll
    for (int j = 0; j < KPT; j++) {
      if (MODE & 1) got[j] = (t * 16u + (uint32_t)j) % (a.cap - 64);
      else got[j] = c[j] ? atomicAdd(&a.cursor[pbase + tid * KPT + j], c[j]) : 0u;
      sum += c[j];
    }
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = lane < NW ? s_warp[lane] : 0, winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += v; }
      if (lane < NW) s_warp[lane] = winc - w;
      if (lane == 31) s_total = winc;
    }
    __syncthreads();
    {
      uint32_t off = s_warp[warp] + inc - sum;
#pragma unroll
      for (int j = 0; j < KPT; j++) {
        cnt[tid * KPT + j] = off;
        gbase[tid * KPT + j] = (pbase + (uint32_t)(tid * KPT + j)) * a.cap + got[j] - off;
        off += c[j];
      }
    }
    __syncthreads();
    const uint32_t total = s_total;
#pragma unroll
    for (int k = 0; k < R; k++) {
      if (rk[k] == 0xFFFFFFFFu) continue;
      const uint32_t id = PACKED ? (uint32_t)key[k] : (uint32_t)(row0 + (uint64_t)k * THREADS + tid);
      stage[cnt[(d[k] >> a.shift) - pbase] + rk[k]] = ((uint64_t)d[k] << 32) | id;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < KPT; j++) cnt[tid * KPT + j] = 0;
    if (!(MODE & 2)) {
      for (uint32_t i = tid; i < total; i += THREADS) {
        const uint64_t v = stage[i];
        const uint32_t part = ((uint32_t)(v >> 32) >> a.shift) - pbase, dst = gbase[part] + i;
        if (dst - (pbase + part) * a.cap < a.cap) a.out[dst] = v;
      }
    }
    __syncthreads();
  }
}

__global__ void gen_kernel(uint64_t* out, uint64_t n, uint64_t dom, uint64_t seed) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) out[i] = mix64(i * 0x9e3779b97f4a7c15ULL + seed) % dom;
}

template <int THREADS, int R, int KPT, bool PACKED, int MODE>
float run(const char* name, Args a, int sms, uint64_t rows, uint32_t cursor_words) {
  constexpr int T = THREADS * R, PP = THREADS * KPT;
  const size_t smem = (size_t)T * 8 + (size_t)PP * 8;
  CK((cudaFuncSetAttribute(part_kernel<THREADS, R, KPT, PACKED, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
  int occ = 0;
  CK((cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, part_kernel<THREADS, R, KPT, PACKED, MODE>, THREADS, smem)));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int it = 0; it < 4; it++) {
    CK(cudaMemset(a.cursor, 0, cursor_words * 4));
    cudaEventRecord(e0);
    part_kernel<THREADS, R, KPT, PACKED, MODE><<<sms * occ, THREADS, smem>>>(a);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (it) best = ms < best ? ms : best;
  }
  printf("%-58s T=%5d thr=%4d occ=%d  %7.3f ms  %6.1f G rows/s  %5.2f TB/s (16 B/row)\n", name, T, THREADS, occ, best, rows / best / 1e6, rows * 16.0 / best / 1e9);
  return best;
}

int main(int argc, char** argv) {
  const int lg = argc > 1 ? atoi(argv[1]) : 29;
  const uint64_t n = 1ULL << lg, dom = 1ULL << 27;
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  uint64_t *in, *out, *out2; uint32_t* cursor;
  CK(cudaMalloc(&in, n * 8));
  const uint32_t P = 4096, cap = (uint32_t)(n / P * 5 / 4 + 2048) & ~3u;
  const uint32_t cap1 = (uint32_t)(n / 64 * 9 / 8 + 8192) & ~3u, cap512 = (uint32_t)(n / 512 * 9 / 8 + 4096) & ~3u;
  CK(cudaMalloc(&out, (uint64_t)P * cap * 8));
  CK(cudaMalloc(&out2, (uint64_t)P * cap * 8));
  CK(cudaMalloc(&cursor, 3 * 8192 * 4));
  gen_kernel<<<sms * 8, 256>>>(in, n, dom, 12345);
  CK(cudaDeviceSynchronize());
  Args a{}; a.in = in; a.n = n; a.cursor = cursor; a.out = out;
  printf("rows 2^%d, key domain 2^27, %d SMs\n", lg, sms);
  a.shift = 15; a.P = 4096; a.cap = cap;
  run<512, 16, 8, false, 0>("A  one pass 4096-way", a, sms, n, 8192);
  run<512, 16, 8, false, 1>("A1 one pass 4096-way, no global atomics", a, sms, n, 8192);
  run<512, 16, 8, false, 2>("A2 one pass 4096-way, no write-out", a, sms, n, 8192);
  run<512, 16, 8, false, 3>("A3 one pass 4096-way, no atomics, no write-out", a, sms, n, 8192);
  run<1024, 16, 4, false, 0>("D  one pass 4096-way, 16 K-row tiles, 1 CTA/SM", a, sms, n, 8192);
  run<1024, 20, 4, false, 0>("D2 one pass 4096-way, 20 K-row tiles, 1 CTA/SM", a, sms, n, 8192);
  a.shift = 18; a.P = 512; a.cap = cap512;
  run<512, 16, 1, false, 0>("B  one pass 512-way", a, sms, n, 8192);
  run<256, 16, 2, false, 0>("B2 one pass 512-way, 256 threads, 4 K tiles", a, sms, n, 8192);
  a.shift = 21; a.P = 64; a.cap = cap1;
  run<512, 16, 1, false, 0>("C  coarse pass 64-way (8 K tiles)", a, sms, n, 8192);
  run<256, 16, 1, false, 0>("C2 coarse pass 64-way (4 K tiles, 256 threads)", a, sms, n, 8192);
  run<256, 8, 1, false, 0>("C3 coarse pass 64-way (2 K tiles, 256 threads)", a, sms, n, 8192);
  // second level: the 64 coarse regions, each split 64-way into the final 4096 ranges
  Args b{}; b.in = out; b.seg_cur = cursor; b.seg_cap = cap1; b.n_seg = 64; b.seg_shift = 6; b.shift = 15; b.P = 4096; b.cap = cap; b.cursor = cursor + 8192; b.out = out2;
  run<512, 16, 1, false, 0>("C  (again, to fill the coarse regions)", a, sms, n, 8192);
  // keep the coarse cursors: run() clears only the cursor block it is given
  run<512, 16, 1, true, 0>("F  fine pass 64 x 64-way (8 K tiles)", b, sms, n, 0);
  run<256, 16, 1, true, 0>("F2 fine pass 64 x 64-way (4 K tiles, 256 threads)", b, sms, n, 0);
  run<256, 8, 1, true, 0>("F3 fine pass 64 x 64-way (2 K tiles, 256 threads)", b, sms, n, 0);
  // 512-way first, then 8-way
  a.shift = 18; a.P = 512; a.cap = cap512;
  run<512, 16, 1, false, 0>("B  (again, to fill the 512 regions)", a, sms, n, 8192);
  Args g{}; g.in = out; g.seg_cur = cursor; g.seg_cap = cap512; g.n_seg = 512; g.seg_shift = 3; g.shift = 15; g.P = 4096; g.cap = cap; g.cursor = cursor + 8192; g.out = out2;
  run<256, 8, 1, true, 0>("G  fine pass 512 x 8-way (2 K tiles, 256 threads)", g, sms, n, 0);
  return 0;
}
