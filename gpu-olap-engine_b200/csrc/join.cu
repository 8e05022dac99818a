// Radix hash join on int64 keys: shared-memory-staged, write-combining radix partition of both sides -> ONE persistent
// pipeline kernel that clears, builds and probes the L2-resident bucket table of every partition.  Replaces
// radix_partition_kernel, build_hash_table_kernel and probe_hash_table_kernel
// (gpu-kernel-compiler/kernels/join_kernels.cuh:45-166) and the stub JoinKernel::join
// (gpu-kernel-compiler/src/join_kernel.rs:39-68).  Also: the multi-GPU shuffle (same partition kernel, destination rank
// instead of radix partition, optionally storing straight into the peers' receive buffers) and the partition pass of the
// partitioned group-by.
//
// Why this shape (measured on B200, scratch/micro.cu -> profiles/r01_microbench_hw_facts.txt):
//   * random 32-byte gathers run at ~260 G/s while the table is L2-resident (<= 64 MB) but 37 G/s once it spills
//     to HBM (4 GB), and global atomics 110-165 G/s vs 20 G/s -> partition until one partition's table fits L2;
//   * a per-row global atomic + scatter (the reference's partition kernel) does 18 G rows/s; shared-memory atomics
//     are NOT a bottleneck (820 G rows/s histogram) -> rank rows with smem atomics, stage the tile in shared memory,
//     write each partition's run with coalesced stores; output positions come from one global atomic per (tile,
//     partition) into fixed-capacity regions, or from a histogram pass when an exact contiguous layout is needed;
//   * one global atomic per match (the reference's probe) serialises on one address (1.5 G/s) -> matches are staged
//     per CTA tile in shared memory and reserved with one atomic per tile.
#include "common.cuh"
#include "djoin.cuh"

#include <algorithm>
#include <cstdlib>

namespace olap {

// ------------------------------------------------------------------------------------------- partitioning
constexpr int kPartThreads = 512;
constexpr int kPartTile = 4096;          // rows staged per CTA iteration (8 per thread); 256-thread CTAs / 2048-row tiles measured 50 % slower (shorter runs)
constexpr int kPartCtasPerSm = 2;
constexpr int kMaxPartBits = 8;          // <= 256 partitions per pass
constexpr int kMaxPay = 4;
constexpr int kMaxPeers = 8;
constexpr uint64_t kShuffleSalt = 0x5bd1e9955bd1e995ULL;  // multi-GPU destination hash = mix64(key ^ salt)               // 8-byte payload columns carried along (group-by)

struct PartArgs {
  const uint64_t* keys;
  const uint32_t* ids_in;                // NULL: implicit row index
  const uint64_t* pay_in[kMaxPay];
  uint64_t n, chunk;                     // rows, rows per CTA (multiple of kPartTile)
  int bits;
  uint64_t salt;                         // xor-ed into the key before hashing: the multi-GPU shuffle uses bits independent of the local join's
  uint32_t* hist;                        // [P][gridDim.x]: counts, then exclusive offsets after the scan
  uint32_t* cursor;                      // ALLOC mode: rows handed out so far per partition (global atomics), [P]
  uint32_t cap;                          // ALLOC mode: partition p owns output rows [p * cap, (p + 1) * cap)
  // PEER mode (multi-GPU shuffle fused with the exchange, <= kMaxPeers partitions): partition p is written straight into
  // rank p's receive buffers (peer-mapped pointers), starting at row peer_off[p] there; exact per-CTA offsets from the histogram
  uint64_t* peer_keys[8];
  uint32_t* peer_ids[8];
  uint32_t peer_off[8];
  const uint32_t* part_begin;            // PEER mode: scanned partition starts (device, [P + 1])
  // RANGE mode (direct-address join on a dense key domain): partition = (key - rmin) >> rshift; keys outside
  // [rmin, rmin + rrange) cannot match and are dropped
  int64_t rmin; uint64_t rrange; int rshift;
  uint64_t* out_keys;
  uint32_t* out_ids;
  uint64_t* pay_out[kMaxPay];
  TermPack terms;                        // rows failing the conjunction are dropped (fused filter for group-by)
  bool vec;                              // keys are 32-byte aligned: 256-bit loads
  bool terms_vec;                        // so are the fused filter's columns
};

// Partition hash: one multiply (Fibonacci hashing of the folded key) for the local radix pass -- the full mix64 cost ~20 of
// the ~85 instructions per row; the salted multi-GPU shuffle keeps mix64 so its bits stay independent of these.
__host__ __device__ __forceinline__ uint64_t part_hash(uint64_t key, uint64_t salt) {
  if (salt) return mix64(key ^ salt);
  return (key ^ (key >> 29)) * 0x9E3779B97F4A7C15ULL;
}

__device__ __forceinline__ void load_keys4(const PartArgs& a, uint64_t i, uint64_t end, uint64_t k[4], bool ok[4]) {
  if (a.vec && i + 3 < end) {
    u64x4 v = ld_stream_u64x4(a.keys + i);
    k[0] = v.a; k[1] = v.b; k[2] = v.c; k[3] = v.d;
    ok[0] = ok[1] = ok[2] = ok[3] = true;
  } else {
#pragma unroll
    for (int e = 0; e < 4; e++) { ok[e] = i + e < end; k[e] = ok[e] ? ld_stream_u64(a.keys + i + e) : 0; }
  }
}
// the fused filter of the partitioned group-by, applied after ALL of a tile's key loads have been issued (a branch between
// the loads kept them from overlapping: 10.4 instead of 9.3 ms for the 2^30-row partition pass)
__device__ __forceinline__ void filter_keys4(const PartArgs& a, uint64_t i, uint64_t end, bool ok[4]) {
  bool f[4];
  if (a.terms_vec) eval_quad<true>(a.terms, i, end, f); else eval_quad<false>(a.terms, i, end, f);
#pragma unroll
  for (int e = 0; e < 4; e++) ok[e] = ok[e] && f[e];
}

template <bool TERMS>
__global__ void __launch_bounds__(kPartThreads) part_hist_kernel(PartArgs a) {
  __shared__ uint32_t s_h[1 << kMaxPartBits];
  const int P = 1 << a.bits;
  for (int i = threadIdx.x; i < P; i += kPartThreads) s_h[i] = 0;
  __syncthreads();
  const uint64_t begin = blockIdx.x * a.chunk, end = min(a.n, begin + a.chunk);
  const int sh = 64 - a.bits;
  for (uint64_t i = begin + (uint64_t)threadIdx.x * 4; i < end; i += (uint64_t)kPartThreads * 4) {
    uint64_t k[4]; bool ok[4];
    load_keys4(a, i, end, k, ok);
    if (TERMS) filter_keys4(a, i, end, ok);
#pragma unroll
    for (int e = 0; e < 4; e++) if (ok[e]) atomicAdd(&s_h[part_hash(k[e], a.salt) >> sh], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < P; i += kPartThreads) a.hist[(uint64_t)i * gridDim.x + blockIdx.x] = s_h[i];
}

// exclusive scan of hist[P * G] in place (partition-major, so element (p, c) becomes the first output position of
// CTA c's rows of partition p); part_begin[p] = offset of (p, 0); part_begin[P] = total.  One CTA.
__global__ void __launch_bounds__(1024) part_scan_kernel(uint32_t* hist, uint32_t m, uint32_t G, uint32_t P, uint32_t* part_begin,
                                                         uint32_t* part_count, uint32_t align_rows) {
  __shared__ uint32_t s_w[32];
  __shared__ uint32_t s_total;
  __shared__ uint32_t s_shift[1 << kMaxPartBits];
  const uint32_t per = (m + 1023) / 1024, lo = threadIdx.x * per, hi = min(m, lo + per);
  uint32_t sum = 0;
  for (uint32_t i = lo; i < hi; i++) sum += hist[i];
  uint32_t inc = sum;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) s_w[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = s_w[lane], winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += t; }
    s_w[lane] = winc - w;
    if (lane == 31) s_total = winc;
  }
  __syncthreads();
  uint32_t run = s_w[warp] + inc - sum;
  for (uint32_t i = lo; i < hi; i++) { uint32_t c = hist[i]; hist[i] = run; run += c; }
  __syncthreads();
  if (align_rows > 1) {
    // multi-GPU send regions: every partition starts on a multiple of align_rows rows (NCCL's peer copies run at 425 GB/s
    // on 16-byte aligned segments and ~170 GB/s otherwise); the regions keep their sizes, gaps appear between them
    if (threadIdx.x == 0) {
      uint32_t end = 0;
      for (uint32_t p = 0; p < P; p++) {
        const uint32_t start = hist[(uint64_t)p * G], size = (p + 1 < P ? hist[(uint64_t)(p + 1) * G] : s_total) - start;
        const uint32_t aligned = (end + align_rows - 1) / align_rows * align_rows;
        s_shift[p] = aligned - start;
        end = aligned + size;
        if (part_count) part_count[p] = size;
      }
      s_total = end;
    }
    __syncthreads();
    for (uint32_t i = lo; i < hi; i++) hist[i] += s_shift[i / G];
    __syncthreads();
  } else if (part_count) {
    for (uint32_t p = threadIdx.x; p < P; p += 1024) part_count[p] = (p + 1 < P ? hist[(uint64_t)(p + 1) * G] : s_total) - hist[(uint64_t)p * G];
  }
  if (part_begin) {
    for (uint32_t p = threadIdx.x; p < P; p += 1024) part_begin[p] = hist[(uint64_t)p * G];
    if (threadIdx.x == 0) part_begin[P] = s_total;
  }
}

// signed min / max of a key column -> out[0], out[1] (initialised to INT64_MAX / INT64_MIN by the caller)
__global__ void __launch_bounds__(256) key_minmax_kernel(const uint64_t* __restrict__ keys, uint64_t n, long long* __restrict__ out) {
  long long lo = INT64_MAX, hi = INT64_MIN;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    const long long k = (long long)ld_stream_u64(keys + i);
    lo = min(lo, k); hi = max(hi, k);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
  if ((threadIdx.x & 31) == 0) { atomicMin(out, lo); atomicMax(out + 1, hi); }
}

// ALLOC = false: exact layout; every CTA's output offsets come from the histogram pass (contiguous partitions, no atomics).
// ALLOC = true : no histogram pass; partition p owns a fixed-capacity region and every tile reserves its run in it with
//                one global atomic per (tile, partition).  Saves a full read of the keys; a region that overflows (skewed
//                keys) is reported through cursor[p] > cap and the caller falls back to the exact layout.
// PEER  = true : exact offsets like ALLOC = false, but partition p (= destination rank p) has its own output arrays:
//                the stores of the write-out phase go over NVLink into the peers' receive buffers -- the exchange of the
//                multi-GPU shuffle happens inside this kernel, run by run, instead of in a collective after it.
// TERMS = true : a fused filter conjunction drops rows before they are ranked (partitioned group-by with a predicate); a
//                template parameter because even dead filter code cost the join's instantiations registers and 8 % of speed.
template <int NPAY, bool IDS, bool ALLOC, bool PEER = false, bool TERMS = false, bool RANGE = false>
__global__ void __launch_bounds__(kPartThreads, NPAY > 1 ? 1 : kPartCtasPerSm) partition_kernel(PartArgs a) {
  constexpr bool BYPART = ALLOC || PEER;  // the staged row remembers its partition, the destination is formed at write-out
  __shared__ uint64_t* s_kptr[kMaxPeers];
  __shared__ uint32_t* s_iptr[kMaxPeers];
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* s_key = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* s_pay = s_key + kPartTile;                               // [NPAY][kPartTile]
  uint32_t* s_id = reinterpret_cast<uint32_t*>(s_pay + (size_t)NPAY * kPartTile);
  uint32_t* s_dst = s_id + (IDS ? kPartTile : 0);                      // final position of every staged row
  __shared__ uint32_t s_cnt[1 << kMaxPartBits], s_off[1 << kMaxPartBits], s_gbase[1 << kMaxPartBits], s_cursor[1 << kMaxPartBits];
  __shared__ uint32_t s_wtot[8], s_total;

  const int P = 1 << a.bits, sh = 64 - a.bits;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < P; i += kPartThreads) {
    s_cnt[i] = 0;
    s_cursor[i] = ALLOC ? 0u : a.hist[(uint64_t)i * gridDim.x + blockIdx.x];
    if (PEER) {
      s_cursor[i] = s_cursor[i] - a.part_begin[i] + a.peer_off[i];
      s_kptr[i] = a.peer_keys[i]; s_iptr[i] = a.peer_ids[i];
    }
  }
  static_assert(kPartTile <= 65536, "rank is packed into 16 bits");
  __syncthreads();
  const uint64_t begin = blockIdx.x * a.chunk, end = min(a.n, begin + a.chunk);

  // software pipeline: the next tile's keys are already in flight (registers) while this tile is ranked, staged and
  // written, so the HBM read latency hides behind the shared-memory phases
  uint64_t kn[2][4]; bool okn[2][4];
#pragma unroll
  for (int j = 0; j < 2; j++) load_keys4(a, begin + (uint64_t)(j * kPartThreads + tid) * 4, end, kn[j], okn[j]);
  if (TERMS) {
#pragma unroll
    for (int j = 0; j < 2; j++) filter_keys4(a, begin + (uint64_t)(j * kPartThreads + tid) * 4, end, okn[j]);
  }
  for (uint64_t tile = begin; tile < end; tile += kPartTile) {
    // phase 1: pick the partition from the high hash bits, rank inside the tile with one smem atomic per row
    uint64_t k[2][4]; uint32_t pr[2][4];  // pr = partition << 16 | rank in the tile; ~0 = dead row
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
      for (int e = 0; e < 4; e++) { k[j][e] = kn[j][e]; pr[j][e] = okn[j][e] ? 0u : 0xFFFFFFFFu; }
#pragma unroll
    for (int j = 0; j < 2; j++) load_keys4(a, tile + kPartTile + (uint64_t)(j * kPartThreads + tid) * 4, end, kn[j], okn[j]);
    if (TERMS) {
#pragma unroll
      for (int j = 0; j < 2; j++) filter_keys4(a, tile + kPartTile + (uint64_t)(j * kPartThreads + tid) * 4, end, okn[j]);
    }
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
      for (int e = 0; e < 4; e++)
        if (pr[j][e] == 0) {
          uint32_t part;
          if (RANGE) {
            const uint64_t d = k[j][e] - (uint64_t)a.rmin;
            if (d >= a.rrange) { pr[j][e] = 0xFFFFFFFFu; continue; }  // outside the build side's key range: no match possible
            part = (uint32_t)(d >> a.rshift);
          } else part = (uint32_t)(part_hash(k[j][e], a.salt) >> sh);
          pr[j][e] = (part << 16) | atomicAdd(&s_cnt[part], 1u);
        }
    __syncthreads();
    // phase 2: exclusive scan of the tile histogram; advance this CTA's global cursors
    uint32_t c = 0, inc = 0, off = 0, got = 0;
    if (tid < 256) {
      c = tid < P ? s_cnt[tid] : 0;
      inc = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
      if (lane == 31) s_wtot[warp] = inc;
      // ALLOC: reserve this tile's run in partition tid's region; the round trip overlaps the scan and the whole of phase 3
      if (ALLOC && c) got = atomicAdd(&a.cursor[tid], c);
    }
    __syncthreads();
    if (tid < 256) {
      uint32_t wbase = 0;
      for (int w = 0; w < warp; w++) wbase += s_wtot[w];
      if (tid < P) {
        off = wbase + inc - c;
        s_off[tid] = off;
        s_cnt[tid] = 0;
        if (!ALLOC) {
          const uint32_t cur = s_cursor[tid];
          s_gbase[tid] = cur - off;  // destination of staged element i of this partition = s_gbase + i (mod 2^32)
          s_cursor[tid] = cur + c;
        }
        if (tid == P - 1) s_total = off + c;
      }
    }
    __syncthreads();
    // phase 3: scatter the tile into shared memory in partition order
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
      for (int e = 0; e < 4; e++)
        if (pr[j][e] != 0xFFFFFFFFu) {
          const uint64_t row = tile + (uint64_t)(j * kPartThreads + tid) * 4 + e;
          const uint32_t part = pr[j][e] >> 16, pos = s_off[part] + (pr[j][e] & 0xFFFFu);
          s_key[pos] = k[j][e];
          s_dst[pos] = BYPART ? part : s_gbase[part] + pos;
          if (IDS) s_id[pos] = a.ids_in ? ld_stream_u32(a.ids_in + row) : (uint32_t)row;
#pragma unroll
          for (int q = 0; q < NPAY; q++) s_pay[(size_t)q * kPartTile + pos] = ld_stream_u64(a.pay_in[q] + row);
        }
    if (ALLOC && tid < P) {
      s_gbase[tid] = (uint32_t)tid * a.cap + got - off;
      s_cursor[tid] = got + c <= a.cap;  // the run fits (all or none: an overflowing region is redone by the caller)
    }
    __syncthreads();
    // phase 4: each partition's run leaves with coalesced stores (write-combined per CTA)
    const uint32_t total = s_total;
    for (uint32_t i = tid; i < total; i += kPartThreads) {
      uint32_t dst = s_dst[i];
      if (PEER) {
        const uint32_t part = dst;
        dst = s_gbase[part] + i;
        s_kptr[part][dst] = s_key[i];
        if (IDS) s_iptr[part][dst] = s_id[i];
        continue;
      }
      if (ALLOC) {
        if (!s_cursor[dst]) continue;
        dst = s_gbase[dst] + i;
      }
      a.out_keys[dst] = s_key[i];
      if (IDS) a.out_ids[dst] = s_id[i];
#pragma unroll
      for (int q = 0; q < NPAY; q++) a.pay_out[q][dst] = s_pay[(size_t)q * kPartTile + i];
    }
    __syncthreads();
  }
}

static inline bool aligned32(const void* p) { return (((uintptr_t)p) & 31) == 0; }

struct PartPlan { unsigned grid; uint64_t chunk; };
static PartPlan plan_partition(olap_ctx* ctx, uint64_t n) {
  unsigned grid = (unsigned)ctx->sm_count * kPartCtasPerSm;
  uint64_t tiles = (n + kPartTile - 1) / kPartTile;
  if (tiles < grid) grid = (unsigned)(tiles ? tiles : 1);
  uint64_t chunk = ((tiles + grid - 1) / grid) * kPartTile;
  return {grid, chunk};
}

template <int NPAY, bool IDS, bool ALLOC = false, bool PEER = false, bool TERMS = false, bool RANGE = false>
static olap_status launch_partition(olap_ctx* ctx, PartArgs& a, unsigned grid, cudaStream_t s) {
  const size_t smem = (size_t)kPartTile * (8 + 8 * NPAY + (IDS ? 4 : 0) + 4);
  OLAP_CUDA((cudaFuncSetAttribute(partition_kernel<NPAY, IDS, ALLOC, PEER, TERMS, RANGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
  partition_kernel<NPAY, IDS, ALLOC, PEER, TERMS, RANGE><<<grid, kPartThreads, smem, s>>>(a);
  count_launch(ctx);
  return OLAP_OK;
}

// (key, payload columns) without row ids: the group-by's partition pass, with or without a fused filter
template <bool ALLOC>
static olap_status launch_payload_partition(olap_ctx* ctx, PartArgs& a, int npay, unsigned grid, cudaStream_t s) {
  const bool t = a.terms.n > 0;
  switch (npay) {
    case 0: return t ? launch_partition<0, false, ALLOC, false, true>(ctx, a, grid, s) : launch_partition<0, false, ALLOC>(ctx, a, grid, s);
    case 1: return t ? launch_partition<1, false, ALLOC, false, true>(ctx, a, grid, s) : launch_partition<1, false, ALLOC>(ctx, a, grid, s);
    case 2: return t ? launch_partition<2, false, ALLOC, false, true>(ctx, a, grid, s) : launch_partition<2, false, ALLOC>(ctx, a, grid, s);
    case 3: return t ? launch_partition<3, false, ALLOC, false, true>(ctx, a, grid, s) : launch_partition<3, false, ALLOC>(ctx, a, grid, s);
    default: return t ? launch_partition<4, false, ALLOC, false, true>(ctx, a, grid, s) : launch_partition<4, false, ALLOC>(ctx, a, grid, s);
  }
}

// partition bounds as two device arrays: rows of partition p are [bounds[p], bounds[P + p])
__global__ void part_bounds_kernel(const uint32_t* __restrict__ part_begin, const uint32_t* __restrict__ cursor, uint32_t cap, uint32_t P,
                                   uint32_t* __restrict__ bounds) {
  for (uint32_t p = threadIdx.x; p < P; p += blockDim.x) {
    if (cursor) { bounds[p] = p * cap; bounds[P + p] = p * cap + min(cursor[p], cap); }
    else { bounds[p] = part_begin[p]; bounds[P + p] = part_begin[p + 1]; }
  }
}

// Capacity of one partition's region in ALLOC mode: 25 % + two tiles above the mean (hash partitions of a relation without
// heavy hitters stay within a fraction of a percent of the mean), a multiple of the tile so regions stay vector-aligned.
static uint64_t alloc_capacity(uint64_t n, uint32_t P) {
  const uint64_t mean = (n + P - 1) / P;
  return (mean + mean / 4 + 2 * kPartTile + kPartTile - 1) / kPartTile * kPartTile;
}

// one pass: (key, row id) of one relation into fixed-capacity regions; cursor[P] must be zeroed by the caller's stream order
static olap_status run_partition_alloc(olap_ctx* ctx, PartArgs& a, cudaStream_t s, const char* part_mark) {
  PartPlan pl = plan_partition(ctx, a.n);
  a.chunk = pl.chunk;
  a.vec = aligned32(a.keys);
  OLAP_CUDA(cudaMemsetAsync(a.cursor, 0, (size_t)(1u << a.bits) * 4, s));
  if (a.rrange) OLAP_TRY((launch_partition<0, true, true, false, false, true>(ctx, a, pl.grid, s)));
  else OLAP_TRY((launch_partition<0, true, true>(ctx, a, pl.grid, s)));
  count_launch(ctx);
  OLAP_CUDA(cudaGetLastError());
  if (part_mark) phase_mark(ctx, s, part_mark);
  return OLAP_OK;
}

// histogram + scan + scatter of one relation.  hist must hold P * grid u32.
static olap_status run_partition(olap_ctx* ctx, PartArgs& a, int npay, bool ids, uint32_t* part_begin, cudaStream_t s,
                                 const char* hist_mark = nullptr, const char* part_mark = nullptr, uint32_t* part_count = nullptr,
                                 uint32_t align_rows = 0) {
  PartPlan pl = plan_partition(ctx, a.n);
  a.chunk = pl.chunk;
  a.vec = aligned32(a.keys);
  const uint32_t P = 1u << a.bits;
  if (a.terms.n) part_hist_kernel<true><<<pl.grid, kPartThreads, 0, s>>>(a);
  else part_hist_kernel<false><<<pl.grid, kPartThreads, 0, s>>>(a);
  part_scan_kernel<<<1, 1024, 0, s>>>(a.hist, P * pl.grid, pl.grid, P, part_begin, part_count, align_rows);
  count_launch(ctx, 2);
  if (hist_mark) phase_mark(ctx, s, hist_mark);
  if (ids) {
    if (npay == 0) OLAP_TRY((launch_partition<0, true>(ctx, a, pl.grid, s)));
    else { set_error("ids + payload partitioning is not instantiated"); return OLAP_ERR_UNSUPPORTED; }
  } else {
    OLAP_TRY((launch_payload_partition<false>(ctx, a, npay, pl.grid, s)));
  }
  OLAP_CUDA(cudaGetLastError());
  if (part_mark) phase_mark(ctx, s, part_mark);
  return OLAP_OK;
}

// ------------------------------------------------------------------------------------------- build / probe
// The table is an array of 32-byte buckets = one L2 sector: {key0, key1, row0, row1, count, maxdisp}.
//   build : ONE atomicAdd on the bucket's counter per visited bucket: 0/1 -> the slot is ours (plain stores follow),
//           >= 2 -> full, try the next bucket.  A key that lands d > 0 buckets past its home bucket h records
//           atomicMax(table[h].maxdisp, d).
//   probe : reads the home bucket (one sector) and must look further only if maxdisp[h] > 0 -- i.e. only when a key
//           with the SAME home was displaced (about 8 % of probes at load 0.5), no matter how long the linear-probing
//           cluster around it is.
// Measured on the way here (scratch/micro2.cu, profiles/r01_*): a CAS-per-slot build of 512 Ki rows took 29 us even
// with the table in L2 (every failed CAS is a dependent ~1 us round trip; the kernel waits for its longest chain);
// recycling slots through a generation tag instead of clearing made every insert a read-modify-write of a DRAM-resident
// line (20 G/s).  The table is cleared right before its build: a pure write allocates the lines in L2 without a DRAM read.
//
// The first version of these kernels (per-lane loops, a shared-memory continuation queue, two ballots per row) executed
// 363 warp instructions per probe row and 148 per build row (ncu, profiles/r01_ncu_join_v2_*): it was issue-bound, not
// memory-bound.  This version keeps every loop warp-uniform, ranks the hits of four rows with one warp scan, walks the
// rare continuations with all of a thread's chains in flight together, and reuses the partition hash (one multiply).
constexpr int kJoinThreads = 256;
constexpr int kJoinRows = 4;                             // rows per thread per sub-round (one 256-bit key load)
constexpr int kJoinSub = 2;                              // sub-rounds per tile: the sync / flush chain is paid once per tile
constexpr int kJoinTile = kJoinThreads * kJoinRows * kJoinSub;   // 2048 rows per work item
constexpr int kProbeStage = 3072;                        // staged pairs per tile (24 KB)
constexpr int kClearTile = 8192;                         // buckets zeroed per work item (256 KB)
constexpr int kMaxParts = 1 << kMaxPartBits;
constexpr uint32_t kPipeRing = 4;                        // table slots the pipeline cycles through (clear / build / probe / draining)

struct Bucket { uint64_t k0, k1; uint32_t r0, r1, cnt, maxdisp; };  // 32 bytes

// table reads: 256-bit, L2 only (.cg): tables are rebuilt in place by other CTAs of the same kernel, so the non-coherent
// path is out
__device__ __forceinline__ Bucket ld_bucket(const Bucket* p) {
  uint64_t a, b, c, d;
  asm volatile("ld.global.cg.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p) : "memory");
  Bucket r;
  r.k0 = a; r.k1 = b; r.r0 = (uint32_t)c; r.r1 = (uint32_t)(c >> 32); r.cnt = (uint32_t)d; r.maxdisp = (uint32_t)(d >> 32);
  return r;
}

// Home bucket: the 32 hash bits below the ones that chose the partition, scaled to the table (any size: exactly one
// bucket per build row, no power-of-two rounding).  Same multiply as the partition pass.
__device__ __forceinline__ uint32_t home_bucket(uint64_t key, int bits, uint32_t n_buckets) {
  return __umulhi((uint32_t)((part_hash(key, 0) << bits) >> 32), n_buckets);
}

// four consecutive rows of a partitioned (or plain) relation; rows outside [begin, end) are dead
template <bool VEC>
__device__ __forceinline__ void load_rows4(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ ids, uint64_t begin, uint64_t end,
                                           uint64_t row0, uint64_t key[4], uint32_t id[4], bool live[4]) {
  if (VEC && row0 >= begin && row0 + 3 < end) {
    u64x4 v = ld_stream_u64x4(keys + row0);
    key[0] = v.a; key[1] = v.b; key[2] = v.c; key[3] = v.d;
    if (ids) { uint4 w = ld_stream_u32x4(ids + row0); id[0] = w.x; id[1] = w.y; id[2] = w.z; id[3] = w.w; }
    else { id[0] = (uint32_t)row0; id[1] = (uint32_t)row0 + 1; id[2] = (uint32_t)row0 + 2; id[3] = (uint32_t)row0 + 3; }
    live[0] = live[1] = live[2] = live[3] = true;
  } else {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint64_t row = row0 + j;
      live[j] = row >= begin && row < end;
      key[j] = live[j] ? ld_stream_u64(keys + row) : 0;
      id[j] = live[j] ? (ids ? ld_stream_u32(ids + row) : (uint32_t)row) : 0;
    }
  }
}

// Insert one tile of kJoinTile rows starting at `tile` (a multiple of 4 rows).  Called by the whole CTA; no barriers.
template <bool VEC>
__device__ __forceinline__ void build_tile(Bucket* __restrict__ table, uint32_t nb, int bits, const uint64_t* __restrict__ keys,
                                           const uint32_t* __restrict__ ids, uint64_t begin, uint64_t end, uint64_t tile) {
  const int tid = threadIdx.x;
#pragma unroll 1
  for (int sub = 0; sub < kJoinSub; sub++) {
    const uint64_t row0 = tile + (uint64_t)(sub * kJoinThreads + tid) * kJoinRows;
    uint64_t key[kJoinRows];
    uint32_t id[kJoinRows];
    bool pend[kJoinRows];
    load_rows4<VEC>(keys, ids, begin, end, row0, key, id, pend);
    uint32_t home[kJoinRows], b[kJoinRows], d[kJoinRows];
#pragma unroll
    for (int j = 0; j < kJoinRows; j++) { home[j] = b[j] = home_bucket(key[j], bits, nb); d[j] = 0; }
    for (;;) {  // warp-uniform: every lane stays until no lane has an insert pending
      uint32_t slot[kJoinRows];
#pragma unroll
      for (int j = 0; j < kJoinRows; j++) slot[j] = pend[j] ? atomicAdd(&table[b[j]].cnt, 1u) : 0u;  // four L2 atomics in flight
      bool more = false;
#pragma unroll
      for (int j = 0; j < kJoinRows; j++)
        if (pend[j]) {
          if (slot[j] < 2) {
            Bucket* bk = table + b[j];
            (&bk->k0)[slot[j]] = key[j];
            (&bk->r0)[slot[j]] = id[j];
            pend[j] = false;
          } else {
            if (++b[j] == nb) b[j] = 0;
            d[j]++;
            more = true;
          }
        }
      if (!__any_sync(0xffffffffu, more)) break;
    }
#pragma unroll
    for (int j = 0; j < kJoinRows; j++) if (d[j]) atomicMax(&table[home[j]].maxdisp, d[j]);
  }
}

struct ProbeShared {
  uint2 stage[kProbeStage];
  uint32_t n;
  unsigned long long base;
};

__device__ __forceinline__ void stage_or_spill(uint2* stage, uint32_t idx, uint32_t brow, uint32_t pid, olap_pair* out, uint64_t max_matches,
                                               unsigned long long* count) {
  if (idx < kProbeStage) stage[idx] = make_uint2(brow, pid);
  else {  // more matches than the stage holds (duplicate-heavy build side): reserve directly
    const unsigned long long pos = atomicAdd(count, 1ULL);
    if (out && pos < max_matches) out[pos] = olap_pair{brow, pid};
  }
}

// Probe one tile of kJoinTile rows starting at `tile` (a multiple of 4 rows) against one table.  Called by the whole CTA.
// PIPE: called from the join pipeline, whose own barriers already cover the reset of sh.n before and the reuse of sh after.
template <bool VEC, bool PIPE = false>
__device__ __forceinline__ void probe_tile(ProbeShared& sh, const Bucket* __restrict__ table, uint32_t nb, int bits, uint32_t pid_base,
                                           const uint64_t* __restrict__ keys, const uint32_t* __restrict__ ids, uint64_t begin, uint64_t end,
                                           uint64_t tile, olap_pair* __restrict__ out, uint64_t max_matches,
                                           unsigned long long* __restrict__ count) {
  const int tid = threadIdx.x, lane = tid & 31;
  if (!PIPE) {
    if (tid == 0) sh.n = 0;
    __syncthreads();
  }
#pragma unroll 1
  for (int sub = 0; sub < kJoinSub; sub++) {
    const uint64_t row0 = tile + (uint64_t)(sub * kJoinThreads + tid) * kJoinRows;
    uint64_t key[kJoinRows];
    uint32_t pid[kJoinRows];
    bool live[kJoinRows];
    load_rows4<VEC>(keys, ids, begin, end, row0, key, pid, live);
    uint32_t b[kJoinRows], rem[kJoinRows];
    Bucket e[kJoinRows];
#pragma unroll
    for (int j = 0; j < kJoinRows; j++) {  // all home-bucket loads are in flight before any is consumed
      pid[j] += pid_base;
      b[j] = home_bucket(key[j], bits, nb);
      e[j].cnt = 0; e[j].maxdisp = 0;
      if (live[j]) e[j] = ld_bucket(table + b[j]);
    }
    // hits of the four rows are ranked with one warp scan and appended with ONE shared-memory atomic per warp
    bool h0[kJoinRows], h1[kJoinRows];
    uint32_t nh = 0;
#pragma unroll
    for (int j = 0; j < kJoinRows; j++) {
      h0[j] = e[j].cnt >= 1 && e[j].k0 == key[j];
      h1[j] = e[j].cnt >= 2 && e[j].k1 == key[j];
      nh += (uint32_t)h0[j] + (uint32_t)h1[j];
      rem[j] = e[j].maxdisp;
    }
    uint32_t inc = nh;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
    if (total) {
      uint32_t base = 0;
      if (lane == 31) base = atomicAdd(&sh.n, total);
      base = __shfl_sync(0xffffffffu, base, 31);
      uint32_t pos = base + inc - nh;
      if (base + total <= (uint32_t)kProbeStage) {  // warp-uniform: the usual case, plain predicated stores
#pragma unroll
        for (int j = 0; j < kJoinRows; j++) {
          if (h0[j]) sh.stage[pos] = make_uint2(e[j].r0, pid[j]);
          pos += h0[j];
          if (h1[j]) sh.stage[pos] = make_uint2(e[j].r1, pid[j]);
          pos += h1[j];
        }
      } else {
#pragma unroll
        for (int j = 0; j < kJoinRows; j++) {
          if (h0[j]) stage_or_spill(sh.stage, pos++, e[j].r0, pid[j], out, max_matches, count);
          if (h1[j]) stage_or_spill(sh.stage, pos++, e[j].r1, pid[j], out, max_matches, count);
        }
      }
    }
    // continuations: a key with this home bucket was displaced by up to rem buckets (8 % of rows).  Warp-uniform loop; each
    // pass moves the first pending chain of every lane one bucket on (one bucket in registers, not four).
    while (__any_sync(0xffffffffu, (rem[0] | rem[1] | rem[2] | rem[3]) != 0)) {
      const int js = rem[0] ? 0 : rem[1] ? 1 : rem[2] ? 2 : 3;
      const bool act = (rem[0] | rem[1] | rem[2] | rem[3]) != 0;
      uint64_t ks = key[3]; uint32_t bs = b[3], ps = pid[3];
      if (js == 0) { ks = key[0]; bs = b[0]; ps = pid[0]; }
      if (js == 1) { ks = key[1]; bs = b[1]; ps = pid[1]; }
      if (js == 2) { ks = key[2]; bs = b[2]; ps = pid[2]; }
      if (act) {
        if (++bs == nb) bs = 0;
        const Bucket x = ld_bucket(table + bs);
        if (x.cnt >= 1 && x.k0 == ks) stage_or_spill(sh.stage, atomicAdd(&sh.n, 1u), x.r0, ps, out, max_matches, count);
        if (x.cnt >= 2 && x.k1 == ks) stage_or_spill(sh.stage, atomicAdd(&sh.n, 1u), x.r1, ps, out, max_matches, count);
#pragma unroll
        for (int j = 0; j < kJoinRows; j++) if (js == j) { b[j] = bs; rem[j]--; }
      }
    }
  }
  __syncthreads();
  const uint32_t n = min(sh.n, (uint32_t)kProbeStage);
  if (tid == 0 && n) sh.base = atomicAdd(count, (unsigned long long)n);  // one global atomic per tile
  __syncthreads();
  if (out && n) {
    const unsigned long long base = sh.base;
    if (base + n <= max_matches) {
      // 16-byte stores: pairs (lead + 2k, lead + 2k + 1) start on a 16-byte aligned output address
      const uint32_t lead = (uint32_t)((reinterpret_cast<uintptr_t>(out + base) >> 3) & 1);
      if (tid == 0 && lead) st_stream_u64(out + base, (uint64_t)sh.stage[0].x | ((uint64_t)sh.stage[0].y << 32));
      uint32_t i = lead + 2 * tid;
      for (; i + 1 < n; i += 2 * kJoinThreads) {
        const uint2 a = sh.stage[i], b = sh.stage[i + 1];
        st_stream_u64x2(out + base + i, make_ulonglong2((uint64_t)a.x | ((uint64_t)a.y << 32), (uint64_t)b.x | ((uint64_t)b.y << 32)));
      }
      if (i < n) st_stream_u64(out + base + i, (uint64_t)sh.stage[i].x | ((uint64_t)sh.stage[i].y << 32));
    } else {
      for (uint32_t i = tid; i < n; i += kJoinThreads) {
        const unsigned long long pos = base + i;
        if (pos < max_matches) st_stream_u64(out + pos, (uint64_t)sh.stage[i].x | ((uint64_t)sh.stage[i].y << 32));
      }
    }
  }
  if (!PIPE) __syncthreads();
}

// ------------------------------------------------------------------------------------------- direct-address tables
// When the build keys are dense (max - min + 1 <= 5 x rows) and unique -- surrogate keys of a dimension table, the
// benchmark's permutation of 0..n-1 -- a partition's table is just row_id[key - min]: 4 bytes per key of the partition's key
// range (2 MB instead of 16 MB at C2), ONE atomicExch per build row (it also detects a duplicate key, which sends the whole
// join back to the bucket tables), no hash, no key compare, no continuation on the probe side.  Partitions are key RANGES
// ((key - min) >> shift) so that every table covers a contiguous slice of the domain.
constexpr uint32_t kDirectEmpty = 0xFFFFFFFFu;

template <bool VEC>
__device__ __forceinline__ void build_tile_direct(uint32_t* __restrict__ table, uint64_t rmin, uint32_t mask, const uint64_t* __restrict__ keys,
                                                  const uint32_t* __restrict__ ids, uint64_t begin, uint64_t end, uint64_t tile,
                                                  uint32_t* __restrict__ dup_flag) {
  const int tid = threadIdx.x;
#pragma unroll 1
  for (int sub = 0; sub < kJoinSub; sub++) {
    const uint64_t row0 = tile + (uint64_t)(sub * kJoinThreads + tid) * kJoinRows;
    uint64_t key[kJoinRows];
    uint32_t id[kJoinRows];
    bool live[kJoinRows];
    load_rows4<VEC>(keys, ids, begin, end, row0, key, id, live);
    uint32_t old[kJoinRows];
#pragma unroll
    for (int j = 0; j < kJoinRows; j++) old[j] = live[j] ? atomicExch(&table[(uint32_t)(key[j] - rmin) & mask], id[j]) : kDirectEmpty;
    bool dup = false;
#pragma unroll
    for (int j = 0; j < kJoinRows; j++) dup = dup || old[j] != kDirectEmpty;
    if (dup) *dup_flag = 1;  // a second row with the same key: this representation cannot hold it
  }
}

__device__ __forceinline__ uint32_t ld_cg_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <bool VEC>
__device__ __forceinline__ void probe_tile_direct(ProbeShared& sh, const uint32_t* __restrict__ table, uint64_t rmin, uint32_t mask,
                                                  uint32_t pid_base, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ ids,
                                                  uint64_t begin, uint64_t end, uint64_t tile, olap_pair* __restrict__ out,
                                                  uint64_t max_matches, unsigned long long* __restrict__ count) {
  const int tid = threadIdx.x, lane = tid & 31;
#pragma unroll 1
  for (int sub = 0; sub < kJoinSub; sub++) {
    const uint64_t row0 = tile + (uint64_t)(sub * kJoinThreads + tid) * kJoinRows;
    uint64_t key[kJoinRows];
    uint32_t pid[kJoinRows], row[kJoinRows];
    bool live[kJoinRows];
    load_rows4<VEC>(keys, ids, begin, end, row0, key, pid, live);
#pragma unroll
    for (int j = 0; j < kJoinRows; j++) row[j] = live[j] ? ld_cg_u32(table + ((uint32_t)(key[j] - rmin) & mask)) : kDirectEmpty;  // four gathers in flight
    uint32_t nh = 0;
#pragma unroll
    for (int j = 0; j < kJoinRows; j++) nh += row[j] != kDirectEmpty;
    uint32_t inc = nh;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
    if (total) {
      uint32_t base = 0;
      if (lane == 31) base = atomicAdd(&sh.n, total);
      base = __shfl_sync(0xffffffffu, base, 31);
      uint32_t pos = base + inc - nh;  // a tile never yields more pairs than rows here: the stage (>= kJoinTile pairs) cannot overflow
#pragma unroll
      for (int j = 0; j < kJoinRows; j++) {
        if (row[j] != kDirectEmpty) sh.stage[pos] = make_uint2(row[j], pid[j] + pid_base);
        pos += row[j] != kDirectEmpty;
      }
    }
  }
  __syncthreads();
  const uint32_t n = sh.n;
  if (tid == 0 && n) sh.base = atomicAdd(count, (unsigned long long)n);  // one global atomic per tile
  __syncthreads();
  if (out && n) {
    const unsigned long long base = sh.base;
    if (base + n <= max_matches) {
      const uint32_t lead = (uint32_t)((reinterpret_cast<uintptr_t>(out + base) >> 3) & 1);
      if (tid == 0 && lead) st_stream_u64(out + base, (uint64_t)sh.stage[0].x | ((uint64_t)sh.stage[0].y << 32));
      uint32_t i = lead + 2 * tid;
      for (; i + 1 < n; i += 2 * kJoinThreads) {
        const uint2 a = sh.stage[i], b = sh.stage[i + 1];
        st_stream_u64x2(out + base + i, make_ulonglong2((uint64_t)a.x | ((uint64_t)a.y << 32), (uint64_t)b.x | ((uint64_t)b.y << 32)));
      }
      if (i < n) st_stream_u64(out + base + i, (uint64_t)sh.stage[i].x | ((uint64_t)sh.stage[i].y << 32));
    } else {
      for (uint32_t i = tid; i < n; i += kJoinThreads) {
        const unsigned long long pos = base + i;
        if (pos < max_matches) st_stream_u64(out + pos, (uint64_t)sh.stage[i].x | ((uint64_t)sh.stage[i].y << 32));
      }
    }
  }
}

// one table, one relation: the single-table (broadcast) strategy and the per-partition fallback
template <bool VEC>
__global__ void __launch_bounds__(kJoinThreads) build_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ ids,
                                                             uint64_t begin, uint64_t end, Bucket* __restrict__ table,
                                                             uint32_t n_buckets, int bits) {
  const uint64_t first = begin & ~3ULL;  // tiles start 4-row aligned so vector loads stay aligned; rows < begin are masked
  for (uint64_t tile = first + blockIdx.x * (uint64_t)kJoinTile; tile < end; tile += (uint64_t)gridDim.x * kJoinTile)
    build_tile<VEC>(table, n_buckets, bits, keys, ids, begin, end, tile);
}

template <bool VEC>
__global__ void __launch_bounds__(kJoinThreads, 4) probe_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ ids,
                                                                uint64_t begin, uint64_t end, const Bucket* __restrict__ table,
                                                                uint32_t n_buckets, int bits, uint32_t pid_base,
                                                                olap_pair* __restrict__ out, uint64_t max_matches,
                                                                unsigned long long* __restrict__ count) {
  __shared__ ProbeShared sh;
  const uint64_t first = begin & ~3ULL;
  for (uint64_t tile = first + blockIdx.x * (uint64_t)kJoinTile; tile < end; tile += (uint64_t)gridDim.x * kJoinTile)
    probe_tile<VEC>(sh, table, n_buckets, bits, pid_base, keys, ids, begin, end, tile, out, max_matches, count);
}

// ------------------------------------------------------------------------------------------- the join pipeline
// All partitions in ONE persistent kernel.  Work items (a 256 KB clear, a 2048-row build tile, a 2048-row probe tile) are
// handed out by a global ticket in this order:
//     step s = { clear(table of partition s), build(partition s-1), probe(partition s-2) }
// so three partitions are in flight, ~1300 items apart.  An item waits (thread 0 spins on an acquire load, the CTA on a
// barrier) until the items it depends on -- all earlier in ticket order, hence already claimed by running CTAs -- have
// signalled their per-partition counter: build(p) needs every clear(p), probe(p) every build(p), and clear(p) every
// probe(p - R) because table p reuses slot p mod R of a ring of R L2-resident tables.  No grid-wide barrier, no host
// round trip per partition, no ramp / tail per partition: the per-partition launches this replaces (768 at C2) were
// ~1.5 waves of CTAs each.
struct PipeArgs {
  const uint64_t* bkeys; const uint32_t* bids; const uint64_t* pkeys; const uint32_t* pids;
  const uint32_t* bb; const uint32_t* pb;          // device, [2P]: rows of partition p are [x[p], x[P + p])
  uint32_t P; int bits; uint32_t R; uint32_t pid_base; uint32_t load_pct;
  uint32_t direct_entries;                         // DIRECT: keys per partition (2^rshift); tables hold 4-byte row ids
  uint64_t rmin; uint32_t* dup;                    // DIRECT: smallest build key; set to 1 when a build key repeats
  Bucket* arena; uint64_t stride;                  // ring slot r = arena + r * stride buckets
  uint32_t* sync;                                  // [0] ticket, then cleared[P], built[P], probed[P]; zeroed before the launch
  uint32_t grid_dbg;
  unsigned long long* waited;                      // [6]: cycles waited by clear / build / probe items, then the number of waits
  olap_pair* out; uint64_t max_matches; unsigned long long* count;
  uint32_t step_start[kMaxParts + 3];              // first ticket of every step; [P + 2] = number of items
};

// tiles of `tile` rows covering [begin & ~3, end)
__host__ __device__ __forceinline__ uint32_t tiles_of(uint32_t begin, uint32_t end, uint32_t tile) {
  if (end <= begin) return 0;
  return (uint32_t)(((uint64_t)end - (begin & ~3u) + tile - 1) / tile);
}
struct PartWork { uint32_t n_clear, n_build, n_probe, nbk; };
// buckets of a table: load_pct / 100 two-slot buckets per build row (100 = load factor 0.5)
__host__ __device__ __forceinline__ uint32_t table_buckets(uint32_t rows, uint32_t load_pct) {
  const uint64_t b = ((uint64_t)rows * load_pct + 99) / 100;
  return b > 64 ? (uint32_t)b : 64u;
}
__host__ __device__ __forceinline__ PartWork part_work(uint32_t b0, uint32_t b1, uint32_t p0, uint32_t p1, uint32_t load_pct, uint32_t direct_entries = 0) {
  PartWork w;
  w.nbk = direct_entries ? direct_entries / 8 : table_buckets(b1 - b0, load_pct);  // direct tables: size in 32-byte units (for the clear items)
  const bool active = b1 > b0 && p1 > p0;          // inner join: an empty side yields nothing
  w.n_clear = active ? (w.nbk + kClearTile - 1) / kClearTile : 0;
  w.n_build = active ? tiles_of(b0, b1, kJoinTile) : 0;
  w.n_probe = active ? tiles_of(p0, p1, kJoinTile) : 0;
  return w;
}

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void pipe_wait(const uint32_t* counter, uint32_t need, unsigned long long* waited) {
  if (threadIdx.x == 0 && need && ld_acquire_u32(counter) < need) {
    const long long t0 = clock64();
    while (ld_acquire_u32(counter) < need) __nanosleep(128);
    atomicAdd(waited, (unsigned long long)(clock64() - t0));  // cycles spent waiting, per phase (OLAP_JOIN_DEBUG=8 prints them)
    atomicAdd(waited + 3, 1ULL);
  }
  __syncthreads();
}
__device__ __forceinline__ void pipe_signal(uint32_t* counter) {
  __syncthreads();
  if (threadIdx.x == 0) { __threadfence(); atomicAdd(counter, 1u); }
}

template <bool DIRECT>
__global__ void __launch_bounds__(kJoinThreads, DIRECT ? 6 : 4) join_pipeline_kernel(const __grid_constant__ PipeArgs a) {
  __shared__ ProbeShared sh;
  __shared__ uint32_t s_ticket;
  const int tid = threadIdx.x;
  const uint32_t P = a.P;
  uint32_t *ticket = a.sync, *cleared = a.sync + 1, *built = cleared + P, *probed = built + P;
  const uint32_t total = a.step_start[P + 2];
  if (tid == 0) s_ticket = atomicAdd(ticket, 1u);
  __syncthreads();
  for (;;) {
    const uint32_t t = s_ticket;  // written before the barrier that ended the previous item; rewritten only after this item's barriers
    if (t >= total) break;
    uint32_t next = 0;
    if (tid == 0) next = atomicAdd(ticket, 1u);  // in flight while this item runs
    uint32_t lo = 0, hi = P + 2;                 // step s with step_start[s] <= t < step_start[s + 1]
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (a.step_start[mid] <= t) lo = mid; else hi = mid; }
    const uint32_t s = lo;
    uint32_t idx = t - a.step_start[s];
    int phase = 0;
    uint32_t p = s;
    if (s < P) {
      const PartWork w = part_work(a.bb[s], a.bb[P + s], a.pb[s], a.pb[P + s], a.load_pct, a.direct_entries);
      if (idx >= w.n_clear) { idx -= w.n_clear; phase = 1; }
    } else phase = 1;
    if (phase == 1) {
      p = s - 1;
      uint32_t nbd = 0;
      if (s >= 1 && s <= P) nbd = part_work(a.bb[p], a.bb[P + p], a.pb[p], a.pb[P + p], a.load_pct, a.direct_entries).n_build;
      if (idx >= nbd) { idx -= nbd; phase = 2; p = s - 2; }
    }
    const uint32_t b0 = a.bb[p], b1 = a.bb[P + p], p0 = a.pb[p], p1 = a.pb[P + p];
    const PartWork w = part_work(b0, b1, p0, p1, a.load_pct, a.direct_entries);
    Bucket* table = a.arena + (uint64_t)(p % a.R) * a.stride;
    if (phase == 0) {
      // the slot's previous user: the most recent ACTIVE partition q = p - R, p - 2R, ... (an inactive one has no items and
      // signals nothing; the chain clear -> build -> probe of q already orders everything before q)
      uint32_t need = 0, q = p;
      while (q >= a.R && !need) { q -= a.R; need = part_work(a.bb[q], a.bb[P + q], a.pb[q], a.pb[P + q], a.load_pct, a.direct_entries).n_probe; }
      pipe_wait(probed + q, need, a.waited + 0);
      const uint32_t c0 = idx * kClearTile, c1 = min(w.nbk, c0 + kClearTile);
      const uint64_t fill = DIRECT ? ~0ULL : 0ULL;  // direct tables: every entry = kDirectEmpty
      const u64x4 z = {fill, fill, fill, fill};
      for (uint32_t i = c0 + tid; i < c1; i += kJoinThreads) st_stream_u64x4(table + i, z);
      pipe_signal(cleared + p);
    } else if (phase == 1) {
      pipe_wait(cleared + p, w.n_clear, a.waited + 1);
      if (DIRECT) build_tile_direct<true>(reinterpret_cast<uint32_t*>(table), a.rmin, a.direct_entries - 1, a.bkeys, a.bids, b0, b1,
                                          (b0 & ~3u) + (uint64_t)idx * kJoinTile, a.dup);
      else build_tile<true>(table, w.nbk, a.bits, a.bkeys, a.bids, b0, b1, (b0 & ~3u) + (uint64_t)idx * kJoinTile);
      pipe_signal(built + p);
    } else {
      if (tid == 0) sh.n = 0;  // ordered before the tile by pipe_wait's barrier
      pipe_wait(built + p, w.n_build, a.waited + 2);
      if (DIRECT) probe_tile_direct<true>(sh, reinterpret_cast<const uint32_t*>(table), a.rmin, a.direct_entries - 1, a.pid_base, a.pkeys, a.pids, p0,
                                          p1, (p0 & ~3u) + (uint64_t)idx * kJoinTile, a.out, a.max_matches, a.count);
      else probe_tile<true, true>(sh, table, w.nbk, a.bits, a.pid_base, a.pkeys, a.pids, p0, p1, (p0 & ~3u) + (uint64_t)idx * kJoinTile, a.out,
                                  a.max_matches, a.count);
      pipe_signal(probed + p);
    }
    if (tid == 0) s_ticket = next;
    __syncthreads();
  }
}

// table arena: grown on demand, kept across calls; every table is cleared right before its build
// The arena is ONE buffer per context while calls may come in on different streams: every user first makes its stream wait for
// the previous user's event (arena_acquire) and records its own behind its last kernel (arena_release), so two joins on different
// streams of one context run their pipelines one after the other instead of on top of each other.
static olap_status ensure_arena(olap_ctx* ctx, uint64_t bytes) {
  if (ctx->table_arena_bytes >= bytes) return OLAP_OK;
  if (ctx->table_arena) {
    if (ctx->arena_used) OLAP_CUDA(cudaEventSynchronize(ctx->arena_ev));  // the last user is done with the old buffer
    OLAP_CUDA(cudaFree(ctx->table_arena)); ctx->table_arena = nullptr; ctx->table_arena_bytes = 0;
  }
  OLAP_CUDA(cudaMalloc(&ctx->table_arena, bytes));
  ctx->table_arena_bytes = bytes;
  return OLAP_OK;
}
static olap_status arena_acquire(olap_ctx* ctx, cudaStream_t s) {
  if (ctx->arena_used) OLAP_CUDA(cudaStreamWaitEvent(s, ctx->arena_ev, 0));
  return OLAP_OK;
}
static olap_status arena_release(olap_ctx* ctx, cudaStream_t s) {
  OLAP_CUDA(cudaEventRecord(ctx->arena_ev, s));
  ctx->arena_used = true;
  return OLAP_OK;
}

static unsigned grid_rows(olap_ctx* ctx, uint64_t rows, uint64_t rows_per_cta, int per_sm) {
  uint64_t g = (rows + rows_per_cta - 1) / rows_per_cta;
  const uint64_t cap = (uint64_t)ctx->sm_count * per_sm;
  if (g > cap) g = cap;
  return (unsigned)(g ? g : 1);
}

static uint64_t env_u64(const char* name, uint64_t dflt) {
  const char* v = getenv(name);
  if (!v || !*v) return dflt;
  return strtoull(v, nullptr, 10);
}

static void launch_build(olap_ctx* ctx, const uint64_t* keys, const uint32_t* ids, uint64_t begin, uint64_t end, Bucket* table, uint64_t nbk,
                         int bits, cudaStream_t s) {
  const unsigned grid = grid_rows(ctx, end - (begin & ~3ULL), kJoinTile, 8);
  if (aligned32(keys) && (!ids || aligned32(ids)))
    build_kernel<true><<<grid, kJoinThreads, 0, s>>>(keys, ids, begin, end, table, (uint32_t)nbk, bits);
  else
    build_kernel<false><<<grid, kJoinThreads, 0, s>>>(keys, ids, begin, end, table, (uint32_t)nbk, bits);
  count_launch(ctx);
}

static void launch_probe(olap_ctx* ctx, const uint64_t* keys, const uint32_t* ids, uint64_t begin, uint64_t end, const Bucket* table,
                         uint64_t nbk, int bits, uint32_t pid_base, olap_pair* out, uint64_t max_matches, unsigned long long* count,
                         cudaStream_t s) {
  const unsigned grid = grid_rows(ctx, end - (begin & ~3ULL), kJoinTile, 8);
  if (aligned32(keys) && (!ids || aligned32(ids)))
    probe_kernel<true><<<grid, kJoinThreads, 0, s>>>(keys, ids, begin, end, table, (uint32_t)nbk, bits, pid_base, out, max_matches, count);
  else
    probe_kernel<false><<<grid, kJoinThreads, 0, s>>>(keys, ids, begin, end, table, (uint32_t)nbk, bits, pid_base, out, max_matches, count);
  count_launch(ctx);
}

// Radix-partition (key, row id) of one side of the join.  out_keys / out_ids must hold side_rows_cap(n, bits) rows.  First choice
// is the one-pass ALLOC layout; if a region overflows (heavy hitters) the exact histogram layout is used instead.
// d_bounds / h_bounds: [2P], rows of partition p are [x[p], x[P + p]).
static uint64_t side_rows_cap(uint64_t n, int bits, uint32_t used_parts = 0) {
  const uint32_t P = 1u << bits;
  const uint64_t cap = alloc_capacity(n, used_parts ? used_parts : P);  // key ranges: only `used_parts` of the P regions receive rows
  return std::max<uint64_t>(n, (uint64_t)P * cap);
}

struct KeyRange { int64_t rmin = 0; uint64_t rrange = 0; int rshift = 0; uint32_t used = 0; };  // rrange > 0: partition by key range (direct-address join); used = ranges that can hold rows (<= P, P is rounded up to a power of two)

static olap_status partition_side(olap_ctx* ctx, const uint64_t* keys, uint64_t n, int bits, uint64_t* out_keys, uint32_t* out_ids,
                                  uint32_t* d_bounds, std::vector<uint32_t>& h_bounds, Scratch& sc, cudaStream_t s, const char* hist_mark,
                                  const char* part_mark, const KeyRange* range = nullptr, bool* range_overflow = nullptr) {
  const uint32_t P = 1u << bits;
  h_bounds.resize(2 * (size_t)P);
  PartArgs a;
  memset(&a, 0, sizeof(a));
  a.bits = bits; a.keys = keys; a.n = n; a.out_keys = out_keys; a.out_ids = out_ids;
  if (range) { a.rmin = range->rmin; a.rrange = range->rrange; a.rshift = range->rshift; }
  const uint64_t cap = alloc_capacity(n, range ? range->used : P);
  if ((range || env_u64("OLAP_JOIN_ALLOC", 1)) && (uint64_t)P * cap < 0xFFFFFFFFULL) {
    uint32_t* cursor = nullptr;
    OLAP_TRY(sc.get<uint32_t>(P, &cursor));
    a.cursor = cursor; a.cap = (uint32_t)cap;
    OLAP_TRY(run_partition_alloc(ctx, a, s, part_mark));
    part_bounds_kernel<<<1, 256, 0, s>>>(nullptr, cursor, a.cap, P, d_bounds);
    count_launch(ctx);
    std::vector<uint32_t> used(P);
    OLAP_CUDA(cudaMemcpyAsync(used.data(), cursor, (size_t)P * 4, cudaMemcpyDeviceToHost, s));
    OLAP_CUDA(cudaMemcpyAsync(h_bounds.data(), d_bounds, 2 * (size_t)P * 4, cudaMemcpyDeviceToHost, s));
    OLAP_CUDA(cudaStreamSynchronize(s));
    bool overflow = false;
    for (uint32_t p = 0; p < P; p++) overflow |= used[p] > a.cap;
    if (!overflow) return OLAP_OK;
  }
  if (range) {  // key ranges of very different sizes (skewed probe side): the caller goes back to hash partitions
    if (range_overflow) *range_overflow = true;
    return OLAP_OK;
  }
  uint32_t *hist = nullptr, *beg = nullptr;
  OLAP_TRY(sc.get<uint32_t>((uint64_t)P * ctx->sm_count * kPartCtasPerSm, &hist));
  OLAP_TRY(sc.get<uint32_t>(P + 1, &beg));
  a.cursor = nullptr; a.cap = 0; a.hist = hist;
  OLAP_TRY(run_partition(ctx, a, 0, true, beg, s, hist_mark, part_mark));
  part_bounds_kernel<<<1, 256, 0, s>>>(beg, nullptr, 0, P, d_bounds);
  count_launch(ctx);
  OLAP_CUDA(cudaMemcpyAsync(h_bounds.data(), d_bounds, 2 * (size_t)P * 4, cudaMemcpyDeviceToHost, s));
  OLAP_CUDA(cudaStreamSynchronize(s));
  return OLAP_OK;
}

constexpr olap_status OLAP_ERR_INTERNAL_RETRY = (olap_status)1000;  // never leaves this file

// min / max of the build keys; r->rrange > 0 when they are dense enough for direct-address tables: the key range is at most
// 5 x the rows (a direct table then takes <= 20 bytes per build row against 32 for the bucket table; 5 rather than 4 so that a
// rank holding a quarter of a dense global domain qualifies whatever its exact row count) and fits 32-bit offsets.
// The range of a (pointer, rows) column is remembered in the context: a join against the same dimension table does not pay
// the extra read again.  A remembered range is only a hint -- the build-side partition pass flags any key outside it and the
// caller then calls this again with refresh = true.
static olap_status detect_dense_keys(olap_ctx* ctx, const uint64_t* keys, uint64_t n, cudaStream_t s, KeyRange* r, bool refresh = false) {
  long long mm[2];
  bool fresh = false;
  auto it = ctx->key_stats.find(keys);
  if (!refresh && env_u64("OLAP_KEY_STATS", 1) && it != ctx->key_stats.end() && it->second.n == n && it->second.have_range) {
    mm[0] = it->second.min; mm[1] = it->second.max;
  } else {
    Scratch sc(ctx, s);
    long long* d = nullptr;
    OLAP_TRY(sc.get<long long>(2, &d));
    const long long init[2] = {INT64_MAX, INT64_MIN};
    OLAP_CUDA(cudaMemcpyAsync(d, init, 16, cudaMemcpyHostToDevice, s));
    key_minmax_kernel<<<(unsigned)ctx->sm_count * 8, 256, 0, s>>>(keys, n, d);
    count_launch(ctx);
    OLAP_CUDA(cudaGetLastError());
    OLAP_CUDA(cudaMemcpyAsync(mm, d, 16, cudaMemcpyDeviceToHost, s));
    OLAP_CUDA(cudaStreamSynchronize(s));
    phase_mark(ctx, s, "minmax");
    fresh = true;
  }
  const uint64_t span = (uint64_t)mm[1] - (uint64_t)mm[0];  // max - min, exact in unsigned arithmetic
  *r = KeyRange();
  if (mm[1] >= mm[0] && span < 0xFFFFFFF0ULL && span + 1 <= 5 * n) { r->rmin = mm[0]; r->rrange = span + 1; }
  if (fresh) {
    // only a DENSE range is remembered: it is verified by the pass that uses it.  A remembered "sparse" verdict would never be
    // checked again and would keep new, dense contents at the same address off the fast path for good.
    if (r->rrange) {
      if (ctx->key_stats.size() > 4096) ctx->key_stats.clear();
      auto& st = ctx->key_stats[keys];
      st.n = n; st.min = mm[0]; st.max = mm[1]; st.have_range = true;
    } else if (it != ctx->key_stats.end()) ctx->key_stats.erase(it);
  }
  return OLAP_OK;
}

// min / max of a rank's shard of a key column for the multi-GPU join (shard.cu), remembered per (pointer, rows) in a map of its
// own: a shard's range is judged against the GLOBAL row count, not the local one, and is verified by the shuffle pass.
olap_status key_min_max(olap_ctx* ctx, const uint64_t* keys, uint64_t n, cudaStream_t s, int64_t* mn, int64_t* mx) {
  auto it = ctx->shard_key_stats.find(keys);
  if (env_u64("OLAP_KEY_STATS", 1) && it != ctx->shard_key_stats.end() && it->second.n == n) { *mn = it->second.min; *mx = it->second.max; return OLAP_OK; }
  Scratch sc(ctx, s);
  long long* d = nullptr;
  OLAP_TRY(sc.get<long long>(2, &d));
  const long long init[2] = {INT64_MAX, INT64_MIN};
  long long mm[2];
  OLAP_CUDA(cudaMemcpyAsync(d, init, 16, cudaMemcpyHostToDevice, s));
  key_minmax_kernel<<<(unsigned)ctx->sm_count * 8, 256, 0, s>>>(keys, n, d);
  count_launch(ctx);
  OLAP_CUDA(cudaGetLastError());
  OLAP_CUDA(cudaMemcpyAsync(mm, d, 16, cudaMemcpyDeviceToHost, s));
  OLAP_CUDA(cudaStreamSynchronize(s));
  if (ctx->shard_key_stats.size() > 4096) ctx->shard_key_stats.clear();
  auto& st = ctx->shard_key_stats[keys];
  st.n = n; st.min = mm[0]; st.max = mm[1]; st.have_range = true;
  *mn = mm[0]; *mx = mm[1];
  return OLAP_OK;
}

// The direct join with shared-memory tables (djoin.cu), run to completion: *retry = 0 when pairs / count stand, else the OR of
// (1 << kFlag*) telling the caller why another path is needed.
static olap_status smem_join(olap_ctx* ctx, const uint64_t* bk, uint64_t nb, const uint64_t* pk, uint64_t np, const KeyRange& r, uint32_t pid_base,
                             olap_pair* out, uint64_t max_matches, unsigned long long* d_count, int hot, cudaStream_t s, uint32_t* retry) {
  DJBuild b;
  RInput in;
  in.keys = bk; in.n = nb;
  olap_status st = dj_build_prepare(ctx, in, r.rmin, r.rrange, 0, s, &b);
  Scratch sc(ctx, s);
  uint32_t* d_flags = nullptr;
  if (st == OLAP_OK) st = sc.get<uint32_t>(kFlagWords, &d_flags);
  if (st == OLAP_OK) {
    in.keys = pk; in.n = np; in.id_base = pid_base;
    st = dj_probe(b, in, np, out, max_matches, d_count, hot, 0, s, d_flags);
  }
  uint32_t hf[kFlagWords] = {0, 0, 0, 0};
  if (st == OLAP_OK) {
    cudaError_t e = cudaMemcpyAsync(hf, d_flags, sizeof(hf), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) st = cuda_fail(e, "reading the join flags", __FILE__, __LINE__);
  }
  dj_build_free(b, s);
  *retry = 0;
  for (int i = 0; i < kFlagWords; i++) if (hf[i]) *retry |= 1u << i;
  return st;
}

// The build side of a join, prepared once and probed by one or many batches of the probe side (the reference executor hands
// the right side over as a Vec<RecordBatch>; gpu-olap-core/src/executor.rs:66-69).
struct BuildSide {
  olap_ctx* ctx = nullptr;
  int strategy = OLAP_JOIN_RADIX;
  uint64_t nb = 0;
  // broadcast: one table for the whole build side
  Bucket* table = nullptr; uint64_t nbk = 0;
  // radix: the partitioned build side
  uint64_t* pkb = nullptr; uint32_t* pib = nullptr; int bits = 0; uint32_t P = 0;
  KeyRange range;                // range.rrange > 0: direct-address tables over key ranges instead of bucket tables over hash partitions
  uint32_t* d_bounds = nullptr;  // partition bounds on the device, [2P]
  std::vector<uint32_t> hb;      // ... and on the host
  uint64_t max_part = 0;
  // radix, dense unique keys: partitioned by key range for the shared-memory direct join (djoin.cu).  The general layout above
  // is then prepared only if a probe reports that this representation cannot serve (repeated key, skew)
  DJBuild dj; bool use_dj = false; bool general_ready = true;
  const uint64_t* bkeys = nullptr;  // the caller keeps the build keys alive while the handle lives
  cudaEvent_t last_use = nullptr;   // recorded behind every probe: the handle's memory is freed in order behind it
};

static olap_status build_side_free(BuildSide& b, cudaStream_t s) {
  if (b.use_dj) { dj_build_free(b.dj, s); b.use_dj = false; }
  if (b.table) { pool_free(b.ctx, b.table, s); b.table = nullptr; }
  if (b.pkb) { pool_free(b.ctx, b.pkb, s); b.pkb = nullptr; }
  if (b.pib) { pool_free(b.ctx, b.pib, s); b.pib = nullptr; }
  if (b.d_bounds) { pool_free(b.ctx, b.d_bounds, s); b.d_bounds = nullptr; }
  return OLAP_OK;
}

static olap_status build_side_prepare(olap_ctx* ctx, const uint64_t* bk, uint64_t nb, int strategy, cudaStream_t s, BuildSide* b) {
  b->ctx = ctx; b->strategy = strategy; b->nb = nb;
  if (nb == 0) return OLAP_OK;
  if (strategy == OLAP_JOIN_BROADCAST) {
    b->nbk = std::max<uint64_t>(nb, 64);  // one 2-slot bucket per build row: load factor 0.5
    OLAP_TRY(pool_alloc(ctx, b->nbk * 32, s, (void**)&b->table));
    OLAP_CUDA(cudaMemsetAsync(b->table, 0, b->nbk * 32, s));
    count_launch(ctx);
    launch_build(ctx, bk, nullptr, 0, nb, b->table, b->nbk, 0, s);
    phase_mark(ctx, s, "build");
    OLAP_CUDA(cudaGetLastError());
    return OLAP_OK;
  }
  // partitions of ~512 Ki build rows: a 16 MB table each, so the few tables the pipeline has in flight stay L2-resident
  const uint64_t target = env_u64("OLAP_JOIN_PART_ROWS", 512 * 1024);
  int bits = 1;
  while (bits < kMaxPartBits && ((nb + target - 1) / target) > (1ULL << bits)) bits++;
  if (b->range.rrange) {  // direct-address join: <= 256 key ranges of 2^rshift keys (>= 64 Ki keys = 256 KB tables)
    int sh = 16;
    while (((b->range.rrange + (1ULL << sh) - 1) >> sh) > (1ULL << kMaxPartBits)) sh++;
    b->range.rshift = sh;
    const uint64_t used = (b->range.rrange + (1ULL << sh) - 1) >> sh;
    b->range.used = (uint32_t)used;
    bits = 1;
    while ((1ULL << bits) < used) bits++;
  }
  b->bits = bits; b->P = 1u << bits;
  Scratch sc(ctx, s);
  const uint64_t rows = side_rows_cap(nb, bits, b->range.used);
  OLAP_TRY(pool_alloc(ctx, rows * 8 + 32, s, (void**)&b->pkb));
  OLAP_TRY(pool_alloc(ctx, rows * 4 + 32, s, (void**)&b->pib));
  OLAP_TRY(pool_alloc(ctx, 2 * (size_t)b->P * 4, s, (void**)&b->d_bounds));
  bool range_overflow = false;
  OLAP_TRY(partition_side(ctx, bk, nb, bits, b->pkb, b->pib, b->d_bounds, b->hb, sc, s, "hist_build", "partition_build",
                          b->range.rrange ? &b->range : nullptr, &range_overflow));
  if (range_overflow) return OLAP_ERR_INTERNAL_RETRY;
  b->max_part = 0;
  for (uint32_t p = 0; p < b->P; p++) b->max_part = std::max<uint64_t>(b->max_part, b->hb[b->P + p] - b->hb[p]);
  return OLAP_OK;
}

static olap_status build_side_probe(BuildSide& b, const uint64_t* pk, uint64_t np, uint32_t pid_base, olap_pair* out, uint64_t max_matches,
                                    unsigned long long* d_count, cudaStream_t s) {
  olap_ctx* ctx = b.ctx;
  if (b.nb == 0 || np == 0) return OLAP_OK;  // executor.rs:231-233: an empty side yields no rows
  if (b.strategy == OLAP_JOIN_BROADCAST) {
    launch_probe(ctx, pk, nullptr, 0, np, b.table, b.nbk, 0, pid_base, out, max_matches, d_count, s);
    phase_mark(ctx, s, "probe");
    OLAP_CUDA(cudaGetLastError());
    return OLAP_OK;
  }
  const uint32_t P = b.P;
  const int bits = b.bits;
  Scratch sc(ctx, s);
  uint64_t* pkp = nullptr;
  uint32_t *pip = nullptr, *d_pb = nullptr;
  const uint64_t rows = side_rows_cap(np, bits, b.range.used);
  OLAP_TRY(sc.get<uint64_t>(rows + 4, &pkp));
  OLAP_TRY(sc.get<uint32_t>(rows + 8, &pip));
  OLAP_TRY(sc.get<uint32_t>(2 * (uint64_t)P, &d_pb));
  std::vector<uint32_t> hp;
  const bool direct = b.range.rrange != 0;
  bool range_overflow = false;
  OLAP_TRY(partition_side(ctx, pk, np, bits, pkp, pip, d_pb, hp, sc, s, "hist_probe", "partition_probe", direct ? &b.range : nullptr, &range_overflow));
  if (range_overflow) return OLAP_ERR_INTERNAL_RETRY;
  const uint32_t* hb = b.hb.data();
  // buckets per 100 build rows.  150: measured in round 1 on sparse random keys (13.9 ms vs 15.5 ms at 100; dense keys preferred 100,
  // 11.4 vs 12.3 ms) -- and dense unique keys no longer come here, they take the shared-memory direct path (djoin.cu)
  const uint32_t load_pct = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(env_u64("OLAP_JOIN_LOAD", 150), 50), 400);
  const uint32_t direct_entries = direct ? (1u << b.range.rshift) : 0;  // direct tables: one 4-byte row id per key of the range
  const uint64_t nbk = direct ? direct_entries / 8 : table_buckets((uint32_t)b.max_part, load_pct);  // 32-byte units per table slot
  auto work = [&](uint32_t p) { return part_work(hb[p], hb[P + p], hp[p], hp[P + p], load_pct, direct_entries); };

  if (direct || env_u64("OLAP_JOIN_PIPE", 1)) {
    // the join pipeline: one persistent kernel over all partitions (see join_pipeline_kernel)
    PipeArgs pa;
    memset(&pa, 0, sizeof(pa));
    // Ring of table slots.  clear(p) waits for probe(p - R), and with the step order {clear(s), build(s-1), probe(s-2)} that
    // probe is EARLIER in ticket order only if R >= 3 (R == P: no slot is ever reused).  Fewer slots would let a CTA spin
    // on items it -- or CTAs queued behind it -- still has to run: a deadlock, not a slowdown.  Large (skewed) partitions
    // therefore get at least 3 slots even when those no longer fit the L2.
    uint64_t R = std::max<uint64_t>(env_u64("OLAP_JOIN_RING", kPipeRing), 3);
    while (R > 3 && R * nbk * 32 > (96ULL << 20)) R--;
    R = std::min<uint64_t>(R, P);
    OLAP_TRY(ensure_arena(ctx, R * nbk * 32));
    OLAP_TRY(arena_acquire(ctx, s));
    uint32_t* sync = nullptr;
    const uint64_t sync_words = 1 + 3 * (uint64_t)P + 1 + 12;  // counters, padding to 8 bytes, 6 x u64 wait statistics
    OLAP_TRY(sc.get<uint32_t>(sync_words, &sync));
    OLAP_CUDA(cudaMemsetAsync(sync, 0, sync_words * 4, s));
    pa.waited = reinterpret_cast<unsigned long long*>(sync + ((1 + 3 * (uint64_t)P + 1) & ~1ULL));
    pa.bkeys = b.pkb; pa.bids = b.pib; pa.pkeys = pkp; pa.pids = pip; pa.bb = b.d_bounds; pa.pb = d_pb;
    pa.P = P; pa.bits = bits; pa.R = (uint32_t)R; pa.pid_base = pid_base; pa.load_pct = load_pct;
    uint32_t* d_dup = nullptr;
    if (direct) {
      OLAP_TRY(sc.get<uint32_t>(1, &d_dup));
      OLAP_CUDA(cudaMemsetAsync(d_dup, 0, 4, s));
      pa.direct_entries = direct_entries; pa.rmin = (uint64_t)b.range.rmin; pa.dup = d_dup;
    }
    pa.arena = (Bucket*)ctx->table_arena; pa.stride = nbk; pa.sync = sync;
    pa.out = out; pa.max_matches = max_matches; pa.count = d_count;
    uint64_t items = 0;
    for (uint32_t st = 0; st < P + 2; st++) {
      pa.step_start[st] = (uint32_t)items;
      if (st < P) items += work(st).n_clear;
      if (st >= 1 && st <= P) items += work(st - 1).n_build;
      if (st >= 2) items += work(st - 2).n_probe;
    }
    pa.step_start[P + 2] = (uint32_t)items;
    if (items) {
      int occ = 0;
      if (direct) OLAP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, join_pipeline_kernel<true>, kJoinThreads, 0));
      else OLAP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, join_pipeline_kernel<false>, kJoinThreads, 0));
      if (occ < 1) { set_error("join_pipeline_kernel does not fit on an SM"); return OLAP_ERR_CUDA; }
      const uint64_t grid = std::min<uint64_t>((uint64_t)occ * ctx->sm_count, items);
      pa.grid_dbg = (uint32_t)grid;
      if (direct) join_pipeline_kernel<true><<<(unsigned)grid, kJoinThreads, 0, s>>>(pa);
      else join_pipeline_kernel<false><<<(unsigned)grid, kJoinThreads, 0, s>>>(pa);
      count_launch(ctx, 2);
    }
    phase_mark(ctx, s, direct ? "build_probe_direct" : "build_probe");
    OLAP_CUDA(cudaGetLastError());
    OLAP_TRY(arena_release(ctx, s));
    if (direct) {  // a repeated build key cannot live in a direct table: the caller redoes the join with bucket tables
      uint32_t dup = 0;
      OLAP_CUDA(cudaMemcpyAsync(&dup, d_dup, 4, cudaMemcpyDeviceToHost, s));
      OLAP_CUDA(cudaStreamSynchronize(s));
      if (dup) return OLAP_ERR_INTERNAL_RETRY;
    }
    if (env_u64("OLAP_JOIN_DEBUG", 0) & 8) {
      unsigned long long w[6];
      OLAP_CUDA(cudaMemcpyAsync(w, pa.waited, sizeof(w), cudaMemcpyDeviceToHost, s));
      OLAP_CUDA(cudaStreamSynchronize(s));
      fprintf(stderr, "[join pipeline] items %llu grid %u ring %llu | waits clear %llu (%.1f Mcyc) build %llu (%.1f Mcyc) probe %llu (%.1f Mcyc)\n",
              (unsigned long long)items, pa.grid_dbg, (unsigned long long)R, w[3], w[0] / 1e6, w[4], w[1] / 1e6, w[5], w[2] / 1e6);
    }
    return OLAP_OK;
  }

  // OLAP_JOIN_PIPE=0: three launches per partition on S side streams (kept for A/B timing; same time as the pipeline at C2)
  const uint64_t dbg = env_u64("OLAP_JOIN_DEBUG", 0);  // timing decomposition only: 1 skips probes, 2 skips builds, 4 marks every kernel
  int S = (int)std::min<uint64_t>(ctx->streams.size(), env_u64("OLAP_JOIN_STREAMS", 3));
  if (dbg & 4) S = 1;
  if (S < 1) S = 1;
  if (S > 8) S = 8;
  OLAP_TRY(ensure_arena(ctx, (uint64_t)S * nbk * 32));
  OLAP_TRY(arena_acquire(ctx, s));
  OLAP_CUDA(cudaEventRecord(ctx->fork_ev, s));
  for (int k = 0; k < S; k++) OLAP_CUDA(cudaStreamWaitEvent(ctx->streams[k], ctx->fork_ev, 0));
  for (uint32_t p = 0; p < P; p++) {
    const uint64_t b0 = hb[p], b1 = hb[P + p], p0 = hp[p], p1 = hp[P + p];
    if (b1 == b0 || p1 == p0) continue;  // inner join: an empty side yields nothing
    const int k = p % S;
    cudaStream_t st = (dbg & 4) ? s : ctx->streams[k];
    Bucket* table = (Bucket*)ctx->table_arena + (uint64_t)k * nbk;
    const uint64_t tb = table_buckets((uint32_t)(b1 - b0), load_pct);  // this partition's table
    if (!(dbg & 2)) {
      OLAP_CUDA(cudaMemsetAsync(table, 0, tb * 32, st));
      count_launch(ctx);
      if (dbg & 4) phase_mark(ctx, st, "p_clear");
      launch_build(ctx, b.pkb, b.pib, b0, b1, table, tb, bits, st);
      if (dbg & 4) phase_mark(ctx, st, "p_build");
    }
    if (!(dbg & 1)) launch_probe(ctx, pkp, pip, p0, p1, table, tb, bits, pid_base, out, max_matches, d_count, st);
    if (dbg & 4) phase_mark(ctx, st, "p_probe");
  }
  for (int k = 0; k < S; k++) {
    OLAP_CUDA(cudaEventRecord(ctx->join_ev[k], ctx->streams[k]));
    OLAP_CUDA(cudaStreamWaitEvent(s, ctx->join_ev[k], 0));
  }
  phase_mark(ctx, s, "build_probe");
  OLAP_CUDA(cudaGetLastError());
  OLAP_TRY(arena_release(ctx, s));
  return OLAP_OK;
}

// Partition pass of the partitioned group-by (groupby.cu): (key, up to 4 value columns) with the filter fused in.
// Same two layouts as the join's sides: one pass into fixed-capacity regions, or -- when a region overflows (heavy hitters)
// -- the exact histogram layout.  The outputs must hold group_by_partition_capacity(n, bits) rows; bounds: device [2P],
// rows of partition p are [bounds[p], bounds[P + p]); h_bounds receives the same on the host.
uint64_t group_by_partition_capacity(uint64_t n, int bits) { return side_rows_cap(n, bits); }

olap_status group_by_partition_rows(olap_ctx* ctx, const uint64_t* keys, uint64_t n, const TermPack& terms, int n_vcols,
                                    const uint64_t* const* vcols, int bits, uint64_t* out_keys, uint64_t* const* out_vcols,
                                    uint32_t* bounds, uint32_t* h_bounds, cudaStream_t s) {
  if (n_vcols > kMaxPay) { set_error("partitioned group-by carries at most %d value columns", kMaxPay); return OLAP_ERR_UNSUPPORTED; }
  const uint32_t P = 1u << bits;
  Scratch sc(ctx, s);
  PartArgs a;
  memset(&a, 0, sizeof(a));
  a.bits = bits; a.keys = keys; a.n = n; a.out_keys = out_keys; a.terms = terms;
  a.terms_vec = true;
  for (int k = 0; k < terms.n; k++) a.terms_vec = a.terms_vec && aligned32(terms.lhs[k]) && (!terms.rhs[k] || aligned32(terms.rhs[k]));
  for (int q = 0; q < n_vcols; q++) { a.pay_in[q] = vcols[q]; a.pay_out[q] = out_vcols[q]; }
  const uint64_t cap = alloc_capacity(n, P);
  if (env_u64("OLAP_GB_ALLOC", 1) && (uint64_t)P * cap < 0xFFFFFFFFULL) {
    uint32_t* cursor = nullptr;
    OLAP_TRY(sc.get<uint32_t>(P, &cursor));
    a.cursor = cursor; a.cap = (uint32_t)cap;
    PartPlan pl = plan_partition(ctx, n);
    a.chunk = pl.chunk; a.vec = aligned32(keys);
    OLAP_CUDA(cudaMemsetAsync(cursor, 0, (size_t)P * 4, s));
    OLAP_TRY((launch_payload_partition<true>(ctx, a, n_vcols, pl.grid, s)));
    part_bounds_kernel<<<1, 256, 0, s>>>(nullptr, cursor, a.cap, P, bounds);
    count_launch(ctx, 2);
    std::vector<uint32_t> used(P);
    OLAP_CUDA(cudaMemcpyAsync(used.data(), cursor, (size_t)P * 4, cudaMemcpyDeviceToHost, s));
    OLAP_CUDA(cudaMemcpyAsync(h_bounds, bounds, 2 * (size_t)P * 4, cudaMemcpyDeviceToHost, s));
    OLAP_CUDA(cudaStreamSynchronize(s));
    bool overflow = false;
    for (uint32_t p = 0; p < P; p++) overflow |= used[p] > a.cap;
    if (!overflow) return OLAP_OK;
  }
  uint32_t *hist = nullptr, *beg = nullptr;
  OLAP_TRY(sc.get<uint32_t>((uint64_t)P * ctx->sm_count * kPartCtasPerSm, &hist));
  OLAP_TRY(sc.get<uint32_t>(P + 1, &beg));
  a.cursor = nullptr; a.cap = 0; a.hist = hist;
  OLAP_TRY(run_partition(ctx, a, n_vcols, false, beg, s));
  part_bounds_kernel<<<1, 256, 0, s>>>(beg, nullptr, 0, P, bounds);
  count_launch(ctx);
  OLAP_CUDA(cudaMemcpyAsync(h_bounds, bounds, 2 * (size_t)P * 4, cudaMemcpyDeviceToHost, s));
  OLAP_CUDA(cudaStreamSynchronize(s));
  return OLAP_OK;
}

}  // namespace olap

using namespace olap;

extern "C" {

static olap_status finish_count(olap_ctx* ctx, unsigned long long* d_count, uint64_t* host_count, olap_pair* out_pairs, uint64_t max_matches,
                                cudaStream_t s) {
  if (!host_count) return OLAP_OK;
  OLAP_CUDA(cudaMemcpyAsync(host_count, d_count, 8, cudaMemcpyDeviceToHost, s));
  OLAP_CUDA(cudaStreamSynchronize(s));
  if (out_pairs && *host_count > max_matches) {
    set_error("join produced %llu pairs but max_matches is %llu (count is exact, pairs truncated)", (unsigned long long)*host_count,
              (unsigned long long)max_matches);
    return OLAP_ERR_OVERFLOW;
  }
  return OLAP_OK;
}

olap_status olap_hash_join_i64(olap_ctx* ctx, const int64_t* build_keys, uint64_t n_build, const int64_t* probe_keys, uint64_t n_probe,
                               olap_pair* out_pairs, uint64_t max_matches, uint64_t* out_count_dptr, uint64_t* host_count,
                               int32_t strategy, void* stream) {
  OLAP_REQUIRE(ctx, "ctx is NULL");
  OLAP_REQUIRE(n_build == 0 || build_keys, "build_keys is NULL");
  OLAP_REQUIRE(n_probe == 0 || probe_keys, "probe_keys is NULL");
  OLAP_REQUIRE(n_build < 0xFFFFFFFFULL && n_probe < 0xFFFFFFFFULL, "row ids are u32: each side must have < 2^32 - 1 rows");
  OLAP_REQUIRE(strategy >= OLAP_JOIN_AUTO && strategy <= OLAP_JOIN_BROADCAST, "unknown join strategy");
  if (strategy == OLAP_JOIN_SORT_MERGE) {
    set_error("SortMergeJoin is out of scope: the planner never emits it (physical_plan.rs:144)");
    return OLAP_ERR_UNSUPPORTED;
  }
  std::lock_guard<std::mutex> g(ctx->mu);
  OLAP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream_of(stream);
  if (strategy == OLAP_JOIN_AUTO) strategy = olap_choose_join_strategy(n_build);
  if (!out_pairs) max_matches = 0;

  Scratch sc(ctx, s);
  unsigned long long* d_count = (unsigned long long*)out_count_dptr;
  if (!d_count) OLAP_TRY(sc.get<unsigned long long>(1, &d_count));
  OLAP_CUDA(cudaMemsetAsync(d_count, 0, 8, s));
  phase_begin(ctx, s);
  BuildSide b;
  olap_status st = OLAP_OK;
  if (n_build && n_probe) {  // executor.rs:231-233: an empty side yields no rows
    bool done = false;
    if (strategy == OLAP_JOIN_RADIX && n_build >= (1u << 20) && env_u64("OLAP_JOIN_DIRECT", 1)) {
      // dense build keys?  (min / max: remembered per column, else a 1 GB read at C2: 0.2 ms)  Then direct-address tables over key
      // ranges; a repeated build key or badly skewed probe keys send the join back to the general path below.
      KeyRange r;
      st = detect_dense_keys(ctx, (const uint64_t*)build_keys, n_build, s, &r);
      bool try_l2_tables = st == OLAP_OK && r.rrange != 0;
      if (try_l2_tables && djoin_eligible(r.rrange) && env_u64("OLAP_JOIN_SMEM", 1)) {
        auto ps = ctx->key_stats.find(probe_keys);
        int hot = (ps != ctx->key_stats.end() && ps->second.n == n_probe && ps->second.skewed) || env_u64("OLAP_JOIN_HOT", 0);
        for (int attempt = 0; attempt < 4 && st == OLAP_OK && !done; attempt++) {
          uint32_t retry = 0;
          st = smem_join(ctx, (const uint64_t*)build_keys, n_build, (const uint64_t*)probe_keys, n_probe, r, 0, out_pairs, max_matches, d_count, hot, s, &retry);
          if (st != OLAP_OK) break;
          if (!retry) { done = true; break; }
          OLAP_CUDA(cudaMemsetAsync(d_count, 0, 8, s));
          if (retry & (1u << kFlagOutOfRange)) {  // the remembered key range is stale: recompute it
            st = detect_dense_keys(ctx, (const uint64_t*)build_keys, n_build, s, &r, true);
            if (st != OLAP_OK || !r.rrange || !djoin_eligible(r.rrange)) break;
          } else if ((retry & (1u << kFlagOverflow)) && !(retry & (1u << kFlagDup)) && !hot) {  // skewed probe keys: pull the heavy hitters out first
            hot = 1;
            auto& st2 = ctx->key_stats[probe_keys];
            st2.n = n_probe; st2.skewed = true;
          } else break;
        }
        try_l2_tables = false;  // whatever defeated the shared-memory tables (a repeated key, skew) defeats the L2 direct tables too
      }
      if (!done && try_l2_tables && st == OLAP_OK) {
        b.range = r;
        st = build_side_prepare(ctx, (const uint64_t*)build_keys, n_build, strategy, s, &b);
        if (st == OLAP_OK) st = build_side_probe(b, (const uint64_t*)probe_keys, n_probe, 0, out_pairs, max_matches, d_count, s);
        build_side_free(b, s);
        if (st == OLAP_ERR_INTERNAL_RETRY) { st = OLAP_OK; b = BuildSide(); OLAP_CUDA(cudaMemsetAsync(d_count, 0, 8, s)); }
        else done = true;
      }
    }
    if (!done && st == OLAP_OK) {
      st = build_side_prepare(ctx, (const uint64_t*)build_keys, n_build, strategy, s, &b);
      if (st == OLAP_OK) st = build_side_probe(b, (const uint64_t*)probe_keys, n_probe, 0, out_pairs, max_matches, d_count, s);
      build_side_free(b, s);
    }
  }
  phase_end(ctx, s);
  OLAP_TRY(st);
  return finish_count(ctx, d_count, host_count, out_pairs, max_matches, s);
}

}  // extern "C"

struct olap_join_build { BuildSide b; };

extern "C" {

static olap_status dj_side_prepare(olap_ctx* ctx, BuildSide& b, cudaStream_t s, bool refresh) {
  KeyRange r;
  OLAP_TRY(detect_dense_keys(ctx, b.bkeys, b.nb, s, &r, refresh));
  if (b.use_dj) { dj_build_free(b.dj, s); b.use_dj = false; }
  if (!r.rrange || !djoin_eligible(r.rrange)) return OLAP_OK;
  RInput in;
  in.keys = b.bkeys; in.n = b.nb;
  OLAP_TRY(dj_build_prepare(ctx, in, r.rmin, r.rrange, 0, s, &b.dj));
  b.use_dj = true;
  return OLAP_OK;
}

olap_status olap_join_build_create(olap_ctx* ctx, const int64_t* build_keys, uint64_t n_build, int32_t strategy, void* stream,
                                   olap_join_build** out) {
  OLAP_REQUIRE(ctx && out, "NULL argument");
  OLAP_REQUIRE(n_build == 0 || build_keys, "build_keys is NULL");
  OLAP_REQUIRE(n_build < 0xFFFFFFFFULL, "row ids are u32: the build side must have < 2^32 - 1 rows");
  OLAP_REQUIRE(strategy == OLAP_JOIN_AUTO || strategy == OLAP_JOIN_RADIX || strategy == OLAP_JOIN_BROADCAST, "unknown join strategy");
  std::lock_guard<std::mutex> g(ctx->mu);
  OLAP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream_of(stream);
  if (strategy == OLAP_JOIN_AUTO) strategy = olap_choose_join_strategy(n_build);
  auto* h = new olap_join_build();
  BuildSide& b = h->b;
  olap_status st = OLAP_OK;
  if (strategy == OLAP_JOIN_RADIX && n_build >= (1u << 20) && env_u64("OLAP_JOIN_DIRECT", 1) && env_u64("OLAP_JOIN_SMEM", 1)) {
    b.ctx = ctx; b.strategy = strategy; b.nb = n_build; b.bkeys = (const uint64_t*)build_keys;
    st = dj_side_prepare(ctx, b, s, false);
    if (st == OLAP_OK && b.use_dj) b.general_ready = false;
  }
  if (st == OLAP_OK && !b.use_dj) st = build_side_prepare(ctx, (const uint64_t*)build_keys, n_build, strategy, s, &b);
  if (st != OLAP_OK) { build_side_free(b, s); delete h; return st; }
  *out = h;
  return OLAP_OK;
}

olap_status olap_join_probe(olap_join_build* build, const int64_t* probe_keys, uint64_t n_probe, uint32_t probe_row_base, olap_pair* out_pairs,
                            uint64_t max_matches, uint64_t* out_count_dptr, uint64_t* host_count, void* stream) {
  OLAP_REQUIRE(build, "build handle is NULL");
  OLAP_REQUIRE(n_probe == 0 || probe_keys, "probe_keys is NULL");
  OLAP_REQUIRE(n_probe < 0xFFFFFFFFULL && (uint64_t)probe_row_base + n_probe < 0xFFFFFFFFULL, "row ids are u32");
  BuildSide& b = build->b;
  olap_ctx* ctx = b.ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  OLAP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream_of(stream);
  if (!out_pairs) max_matches = 0;
  Scratch sc(ctx, s);
  unsigned long long* d_count = (unsigned long long*)out_count_dptr;
  if (!d_count) OLAP_TRY(sc.get<unsigned long long>(1, &d_count));
  OLAP_CUDA(cudaMemsetAsync(d_count, 0, 8, s));
  bool done = b.nb == 0 || n_probe == 0;
  if (!done && b.use_dj) {
    uint32_t* d_flags = nullptr;
    OLAP_TRY(sc.get<uint32_t>(kFlagWords, &d_flags));
    int hot = 0;
    for (int attempt = 0; attempt < 4 && b.use_dj && !done; attempt++) {
      RInput in;
      in.keys = (const uint64_t*)probe_keys; in.n = n_probe; in.id_base = probe_row_base;
      OLAP_TRY(dj_probe(b.dj, in, n_probe, out_pairs, max_matches, d_count, hot, 0, s, d_flags));
      uint32_t hf[kFlagWords];
      OLAP_CUDA(cudaMemcpyAsync(hf, d_flags, sizeof(hf), cudaMemcpyDeviceToHost, s));
      OLAP_CUDA(cudaStreamSynchronize(s));
      if (!hf[0] && !hf[1] && !hf[2] && !hf[3]) { done = true; break; }
      OLAP_CUDA(cudaMemsetAsync(d_count, 0, 8, s));
      if (hf[kFlagOutOfRange]) OLAP_TRY(dj_side_prepare(ctx, b, s, true));          // stale remembered key range
      else if (hf[kFlagOverflow] && !hf[kFlagDup] && !hot) hot = 1;                  // skewed probe keys
      else { dj_build_free(b.dj, s); b.use_dj = false; }                             // repeated build key: bucket tables from now on
    }
    if (!done && b.use_dj) { dj_build_free(b.dj, s); b.use_dj = false; }
  }
  if (!done) {
    if (!b.general_ready) {
      OLAP_TRY(build_side_prepare(ctx, b.bkeys, b.nb, b.strategy, s, &b));
      b.general_ready = true;
    }
    OLAP_TRY(build_side_probe(b, (const uint64_t*)probe_keys, n_probe, probe_row_base, out_pairs, max_matches, d_count, s));
  }
  if (!b.last_use) OLAP_CUDA(cudaEventCreateWithFlags(&b.last_use, cudaEventDisableTiming));
  OLAP_CUDA(cudaEventRecord(b.last_use, s));
  return finish_count(ctx, d_count, host_count, out_pairs, max_matches, s);
}

olap_status olap_join_build_destroy(olap_join_build* build) {
  if (!build) return OLAP_OK;
  olap_ctx* ctx = build->b.ctx;
  {
    std::lock_guard<std::mutex> g(ctx->mu);
    cudaSetDevice(ctx->device);
    // probes may still be running on the caller's stream (asynchronous when no host count was asked for): free behind them
    if (build->b.last_use) { cudaStreamWaitEvent(ctx->streams[0], build->b.last_use, 0); }
    build_side_free(build->b, ctx->streams[0]);
    if (build->b.last_use) { cudaEventDestroy(build->b.last_use); build->b.last_use = nullptr; }
  }
  delete build;
  return OLAP_OK;
}

static olap_status partition_api(olap_ctx* ctx, const int64_t* keys, const uint32_t* row_ids, uint64_t n_rows, int32_t bits, uint64_t salt,
                                 int64_t* out_keys, uint32_t* out_row_ids, uint32_t* out_part_begin, uint32_t* out_part_count, void* stream) {
  OLAP_REQUIRE(ctx, "ctx is NULL");
  OLAP_REQUIRE(bits >= 1 && bits <= kMaxPartBits, "bits must be in [1, 8]");
  OLAP_REQUIRE(n_rows < 0xFFFFFFFFULL, "row ids are u32");
  OLAP_REQUIRE(out_part_begin && (n_rows == 0 || (keys && out_keys && out_row_ids)), "NULL pointer");
  std::lock_guard<std::mutex> g(ctx->mu);
  OLAP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream_of(stream);
  const uint32_t P = 1u << bits;
  if (n_rows == 0) {
    OLAP_CUDA(cudaMemsetAsync(out_part_begin, 0, (P + 1) * 4, s));
    if (out_part_count) OLAP_CUDA(cudaMemsetAsync(out_part_count, 0, P * 4, s));
    return OLAP_OK;
  }
  Scratch sc(ctx, s);
  uint32_t* hist = nullptr;
  OLAP_TRY(sc.get<uint32_t>((uint64_t)P * ctx->sm_count * kPartCtasPerSm, &hist));
  PartArgs a;
  memset(&a, 0, sizeof(a));
  a.bits = bits; a.salt = salt; a.keys = (const uint64_t*)keys; a.ids_in = row_ids; a.n = n_rows; a.hist = hist;
  a.out_keys = (uint64_t*)out_keys; a.out_ids = out_row_ids;
  return run_partition(ctx, a, 0, true, out_part_begin, s, nullptr, nullptr, out_part_count, out_part_count ? 16 : 0);
}

olap_status olap_radix_partition_i64(olap_ctx* ctx, const int64_t* keys, const uint32_t* row_ids, uint64_t n_rows, int32_t bits,
                                     int64_t* out_keys, uint32_t* out_row_ids, uint32_t* out_part_begin, void* stream) {
  return partition_api(ctx, keys, row_ids, n_rows, bits, 0, out_keys, out_row_ids, out_part_begin, nullptr, stream);
}

olap_status olap_shuffle_partition_i64(olap_ctx* ctx, const int64_t* keys, const uint32_t* row_ids, uint64_t n_rows, int32_t bits,
                                       int64_t* out_keys, uint32_t* out_row_ids, uint32_t* out_part_begin, uint32_t* out_part_count,
                                       void* stream) {
  return partition_api(ctx, keys, row_ids, n_rows, bits, kShuffleSalt, out_keys, out_row_ids, out_part_begin, out_part_count, stream);
}

// Shuffle fused with the exchange.  scratch: olap_shuffle_scratch_words(bits) u32 = [P][grid] histogram, then [P + 1] starts.
uint64_t olap_shuffle_scratch_words(olap_ctx* ctx, int32_t bits) {
  if (!ctx || bits < 0 || bits > 3) return 0;
  const uint64_t P = 1ULL << bits;
  return P * (uint64_t)ctx->sm_count * kPartCtasPerSm + P + 1;
}

static olap_status shuffle_args(olap_ctx* ctx, const int64_t* keys, uint64_t n_rows, int32_t bits, uint32_t* scratch, PartArgs* a, PartPlan* pl) {
  OLAP_REQUIRE(ctx && keys && scratch, "NULL argument");
  OLAP_REQUIRE(bits >= 1 && bits <= 3, "the fused shuffle serves 2, 4 or 8 ranks");
  OLAP_REQUIRE(n_rows > 0 && n_rows < 0xFFFFFFFFULL, "row ids are u32 and the fused shuffle needs at least one row");
  memset(a, 0, sizeof(*a));
  *pl = plan_partition(ctx, n_rows);
  a->bits = bits; a->salt = kShuffleSalt; a->keys = (const uint64_t*)keys; a->n = n_rows; a->hist = scratch;
  a->chunk = pl->chunk; a->vec = aligned32(keys);
  a->part_begin = scratch + (1ULL << bits) * ctx->sm_count * kPartCtasPerSm;
  return OLAP_OK;
}

olap_status olap_shuffle_count_i64(olap_ctx* ctx, const int64_t* keys, uint64_t n_rows, int32_t bits, uint32_t* scratch,
                                   uint32_t* out_part_count, void* stream) {
  PartArgs a; PartPlan pl;
  OLAP_TRY(shuffle_args(ctx, keys, n_rows, bits, scratch, &a, &pl));
  OLAP_REQUIRE(out_part_count, "out_part_count is NULL");
  std::lock_guard<std::mutex> g(ctx->mu);
  OLAP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream_of(stream);
  const uint32_t P = 1u << bits;
  // the kernels index the histogram by the launched grid: clear the slots of a full-size grid first
  OLAP_CUDA(cudaMemsetAsync(scratch, 0, olap_shuffle_scratch_words(ctx, bits) * 4, s));
  if (a.terms.n) part_hist_kernel<true><<<pl.grid, kPartThreads, 0, s>>>(a);
  else part_hist_kernel<false><<<pl.grid, kPartThreads, 0, s>>>(a);
  part_scan_kernel<<<1, 1024, 0, s>>>(a.hist, P * pl.grid, pl.grid, P, const_cast<uint32_t*>(a.part_begin), out_part_count, 0);
  count_launch(ctx, 3);
  OLAP_CUDA(cudaGetLastError());
  return OLAP_OK;
}

olap_status olap_shuffle_scatter_i64(olap_ctx* ctx, const int64_t* keys, uint64_t n_rows, int32_t bits, uint32_t* scratch,
                                     int64_t* const* dest_keys, uint32_t* const* dest_row_ids, const uint64_t* dest_row_offset, void* stream) {
  PartArgs a; PartPlan pl;
  OLAP_TRY(shuffle_args(ctx, keys, n_rows, bits, scratch, &a, &pl));
  OLAP_REQUIRE(dest_keys && dest_row_ids && dest_row_offset, "NULL destination table");
  const uint32_t P = 1u << bits;
  for (uint32_t p = 0; p < P; p++) {
    OLAP_REQUIRE(dest_keys[p] && dest_row_ids[p], "NULL destination pointer");
    OLAP_REQUIRE(dest_row_offset[p] + n_rows < 0xFFFFFFFFULL, "destination offsets are u32 rows");
    a.peer_keys[p] = (uint64_t*)dest_keys[p]; a.peer_ids[p] = dest_row_ids[p]; a.peer_off[p] = (uint32_t)dest_row_offset[p];
  }
  std::lock_guard<std::mutex> g(ctx->mu);
  OLAP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream_of(stream);
  OLAP_TRY((launch_partition<0, true, false, true>(ctx, a, pl.grid, s)));
  OLAP_CUDA(cudaGetLastError());
  return OLAP_OK;
}

uint32_t olap_shuffle_dest_of(int64_t key, int32_t bits) {
  if (bits <= 0) return 0;
  return (uint32_t)(mix64((uint64_t)key ^ kShuffleSalt) >> (64 - bits));
}

}  // extern "C"
