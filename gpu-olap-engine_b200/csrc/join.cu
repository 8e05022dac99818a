// Radix hash join on int64 keys: histogram -> scan -> shared-memory-staged, write-combining radix partition ->
// per-partition open-addressing build/probe in an L2-resident table.  Replaces radix_partition_kernel,
// build_hash_table_kernel and probe_hash_table_kernel (gpu-kernel-compiler/kernels/join_kernels.cuh:45-166) and the
// stub JoinKernel::join (gpu-kernel-compiler/src/join_kernel.rs:39-68).
//
// Why this shape (measured on B200, scratch/micro.cu -> profiles/r01_microbench.txt):
//   * random 16-byte gathers run at ~260 G/s while the table is L2-resident (<= 64 MB) but 37 G/s once it spills
//     to HBM (4 GB), and global atomics 110-165 G/s vs 20 G/s -> partition until one partition's table fits L2;
//   * a per-row global atomic + scatter (the reference's partition kernel) does 18 G rows/s; shared-memory atomics
//     are NOT a bottleneck (820 G rows/s histogram) -> rank rows with smem atomics, stage the tile in shared memory,
//     write each partition's run with coalesced stores, and take the per-CTA output offsets from a histogram pass so
//     that the partition kernel issues no global atomics at all;
//   * one global atomic per match (the reference's probe) serialises on one address (1.5 G/s) -> matches are staged
//     per CTA tile in shared memory and reserved with one atomic per tile.
// Table slots carry a generation tag, so the table arena is never cleared between partitions or joins.
#include "common.cuh"

#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>

namespace olap {

// ------------------------------------------------------------------------------------------- partitioning
constexpr int kPartThreads = 512;
constexpr int kPartTile = 4096;          // rows staged per CTA iteration (8 per thread); 256-thread CTAs / 2048-row tiles measured 50 % slower (shorter runs)
constexpr int kPartCtasPerSm = 2;
constexpr int kMaxPartBits = 8;          // <= 256 partitions per pass
constexpr int kMaxPay = 4;
constexpr uint64_t kShuffleSalt = 0x5bd1e9955bd1e995ULL;  // multi-GPU destination hash = mix64(key ^ salt)               // 8-byte payload columns carried along (group-by)

struct PartArgs {
  const uint64_t* keys;
  const uint32_t* ids_in;                // NULL: implicit row index
  const uint64_t* pay_in[kMaxPay];
  uint64_t n, chunk;                     // rows, rows per CTA (multiple of kPartTile)
  int bits;
  uint64_t salt;                         // xor-ed into the key before hashing: the multi-GPU shuffle uses bits independent of the local join's
  uint32_t* hist;                        // [P][gridDim.x]: counts, then exclusive offsets after the scan
  uint64_t* out_keys;
  uint32_t* out_ids;
  uint64_t* pay_out[kMaxPay];
  TermPack terms;                        // rows failing the conjunction are dropped (fused filter for group-by)
  bool vec;                              // keys are 32-byte aligned: 256-bit loads
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
  if (a.terms.n) {
#pragma unroll
    for (int e = 0; e < 4; e++) if (ok[e]) ok[e] = eval_terms_row(a.terms, i + e);
  }
}

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

template <int NPAY, bool IDS>
__global__ void __launch_bounds__(kPartThreads, NPAY > 1 ? 1 : kPartCtasPerSm) partition_kernel(PartArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* s_key = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* s_pay = s_key + kPartTile;                               // [NPAY][kPartTile]
  uint32_t* s_id = reinterpret_cast<uint32_t*>(s_pay + (size_t)NPAY * kPartTile);
  uint32_t* s_dst = s_id + (IDS ? kPartTile : 0);                      // final position of every staged row
  __shared__ uint32_t s_cnt[1 << kMaxPartBits], s_off[1 << kMaxPartBits], s_gbase[1 << kMaxPartBits], s_cursor[1 << kMaxPartBits];
  __shared__ uint32_t s_wtot[8], s_total;

  const int P = 1 << a.bits, sh = 64 - a.bits;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < P; i += kPartThreads) { s_cnt[i] = 0; s_cursor[i] = a.hist[(uint64_t)i * gridDim.x + blockIdx.x]; }
  __syncthreads();
  const uint64_t begin = blockIdx.x * a.chunk, end = min(a.n, begin + a.chunk);

  // software pipeline: the next tile's keys are already in flight (registers) while this tile is ranked, staged and
  // written, so the HBM read latency hides behind the shared-memory phases
  uint64_t kn[2][4]; bool okn[2][4];
#pragma unroll
  for (int j = 0; j < 2; j++) load_keys4(a, begin + (uint64_t)(j * kPartThreads + tid) * 4, end, kn[j], okn[j]);
  for (uint64_t tile = begin; tile < end; tile += kPartTile) {
    // phase 1: pick the partition from the high hash bits, rank inside the tile with one smem atomic per row
    uint64_t k[2][4]; bool ok[2][4]; uint32_t part[2][4], rank[2][4];
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
      for (int e = 0; e < 4; e++) { k[j][e] = kn[j][e]; ok[j][e] = okn[j][e]; }
#pragma unroll
    for (int j = 0; j < 2; j++) load_keys4(a, tile + kPartTile + (uint64_t)(j * kPartThreads + tid) * 4, end, kn[j], okn[j]);
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
      for (int e = 0; e < 4; e++)
        if (ok[j][e]) { part[j][e] = (uint32_t)(part_hash(k[j][e], a.salt) >> sh); rank[j][e] = atomicAdd(&s_cnt[part[j][e]], 1u); }
    __syncthreads();
    // phase 2: exclusive scan of the tile histogram; advance this CTA's global cursors
    uint32_t c = 0, inc = 0;
    if (tid < 256) {
      c = tid < P ? s_cnt[tid] : 0;
      inc = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
      if (lane == 31) s_wtot[warp] = inc;
    }
    __syncthreads();
    if (tid < 256) {
      uint32_t wbase = 0;
      for (int w = 0; w < warp; w++) wbase += s_wtot[w];
      if (tid < P) {
        const uint32_t off = wbase + inc - c, cur = s_cursor[tid];
        s_off[tid] = off;
        s_gbase[tid] = cur - off;  // destination of staged element i of this partition = s_gbase + i (mod 2^32)
        s_cursor[tid] = cur + c;
        s_cnt[tid] = 0;
        if (tid == P - 1) s_total = off + c;
      }
    }
    __syncthreads();
    // phase 3: scatter the tile into shared memory in partition order
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
      for (int e = 0; e < 4; e++)
        if (ok[j][e]) {
          const uint64_t row = tile + (uint64_t)(j * kPartThreads + tid) * 4 + e;
          const uint32_t pos = s_off[part[j][e]] + rank[j][e];
          s_key[pos] = k[j][e];
          s_dst[pos] = s_gbase[part[j][e]] + pos;
          if (IDS) s_id[pos] = a.ids_in ? ld_stream_u32(a.ids_in + row) : (uint32_t)row;
#pragma unroll
          for (int q = 0; q < NPAY; q++) s_pay[(size_t)q * kPartTile + pos] = ld_stream_u64(a.pay_in[q] + row);
        }
    __syncthreads();
    // phase 4: each partition's run leaves with coalesced stores (write-combined per CTA, no global atomics)
    const uint32_t total = s_total;
    for (uint32_t i = tid; i < total; i += kPartThreads) {
      const uint32_t dst = s_dst[i];
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

template <int NPAY, bool IDS>
static olap_status launch_partition(olap_ctx* ctx, PartArgs& a, unsigned grid, cudaStream_t s) {
  const size_t smem = (size_t)kPartTile * (8 + 8 * NPAY + (IDS ? 4 : 0) + 4);
  OLAP_CUDA(cudaFuncSetAttribute(partition_kernel<NPAY, IDS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  partition_kernel<NPAY, IDS><<<grid, kPartThreads, smem, s>>>(a);
  count_launch(ctx);
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
  part_hist_kernel<<<pl.grid, kPartThreads, 0, s>>>(a);
  part_scan_kernel<<<1, 1024, 0, s>>>(a.hist, P * pl.grid, pl.grid, P, part_begin, part_count, align_rows);
  count_launch(ctx, 2);
  if (hist_mark) phase_mark(ctx, s, hist_mark);
  if (ids) {
    if (npay == 0) OLAP_TRY((launch_partition<0, true>(ctx, a, pl.grid, s)));
    else { set_error("ids + payload partitioning is not instantiated"); return OLAP_ERR_UNSUPPORTED; }
  } else {
    switch (npay) {
      case 0: OLAP_TRY((launch_partition<0, false>(ctx, a, pl.grid, s))); break;
      case 1: OLAP_TRY((launch_partition<1, false>(ctx, a, pl.grid, s))); break;
      case 2: OLAP_TRY((launch_partition<2, false>(ctx, a, pl.grid, s))); break;
      case 3: OLAP_TRY((launch_partition<3, false>(ctx, a, pl.grid, s))); break;
      default: OLAP_TRY((launch_partition<4, false>(ctx, a, pl.grid, s))); break;
    }
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
//           cluster around it is.  Those continuations go to a per-CTA queue and are processed in parallel phases
//           (one dependent L2 round trip per phase for the whole tile).
// Measured on the way here (scratch/micro2.cu, profiles/r01_*): a CAS-per-slot build of 512 Ki rows took 29 us even
// with the table in L2 (every failed CAS is a dependent ~1 us round trip; the kernel waits for its longest chain);
// recycling slots through a generation tag instead of clearing made every insert a read-modify-write of a DRAM-resident
// line (20 G/s); "walk while the bucket is full / was passed" made each warp loop ~5 times per row group with 4 lanes
// active (28 us per 2048-row tile of pure dependent latency).  The table is cleared right before its build: a pure
// write allocates the lines in L2 without a DRAM read.
// Home bucket = mulhi(hash << partition_bits, n_buckets): any table size (exactly 32 B per build row at load 0.5, no
// power-of-two rounding), independent of the bits that chose the partition.
constexpr int kBuildThreads = 256;
constexpr int kProbeThreads = 256;
constexpr int kProbeRows = 4;                            // rows per thread per sub-round (one 256-bit key load)
constexpr int kProbeSub = 2;                             // sub-rounds per tile: the sync/atomic/flush chain is paid once per tile
constexpr int kProbeTile = kProbeThreads * kProbeRows * kProbeSub;   // 2048 probe rows per tile
constexpr int kProbeStage = 3072;                        // staged pairs per tile (24 KB)
constexpr int kProbeQueue = 512;                         // continuation queue entries per phase buffer

struct Bucket { uint64_t k0, k1; uint32_t r0, r1, cnt, maxdisp; };  // 32 bytes

// table reads: 256-bit, L2 only (.cg): the rounds kernel rebuilds tables in place, so the non-coherent path is out
__device__ __forceinline__ Bucket ld_bucket(const Bucket* p) {
  uint64_t a, b, c, d;
  asm volatile("ld.global.cg.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p) : "memory");
  Bucket r;
  r.k0 = a; r.k1 = b; r.r0 = (uint32_t)c; r.r1 = (uint32_t)(c >> 32); r.cnt = (uint32_t)d; r.maxdisp = (uint32_t)(d >> 32);
  return r;
}

__device__ __forceinline__ uint64_t home_bucket(uint64_t key, int bits, uint64_t n_buckets) {
  return __umul64hi(mix64(key) << bits, n_buckets);
}

// The tables one kernel works on: table r lives at arena + r * stride and has nb[r] buckets; build rows
// [bb[r], bb[r+1]) and probe rows [pb[r], pb[r+1]) belong to it.  nr == 1 for the per-partition kernels.
constexpr int kMaxRound = 8;
struct Tables {
  Bucket* arena; uint64_t stride; int nr; int bits;
  uint32_t pid_base;  // added to every probe row id (probe side arriving in batches)
  uint32_t nb[kMaxRound], bb[kMaxRound + 1], pb[kMaxRound + 1];
};
__device__ __forceinline__ int table_of(const uint32_t* begin, int nr, uint64_t row) {
  int r = 0;
  for (int q = 1; q < nr; q++) r += row >= begin[q];
  return r;
}

__device__ __forceinline__ void build_row(const Tables& t, int r, uint64_t key, uint32_t id) {
  Bucket* table = t.arena + (uint64_t)r * t.stride;
  const uint64_t nb = t.nb[r];
  const uint64_t home = home_bucket(key, t.bits, nb);
  uint64_t b = home;
  uint32_t d = 0;
  for (;;) {
    Bucket* bk = table + b;
    const uint32_t slot = atomicAdd(&bk->cnt, 1u);
    if (slot == 0) { bk->k0 = key; bk->r0 = id; break; }
    if (slot == 1) { bk->k1 = key; bk->r1 = id; break; }
    if (++b == nb) b = 0;
    d++;
  }
  if (d) atomicMax(&table[home].maxdisp, d);
}

// one row per thread: every insert's key load, atomic and stores are in flight together
__global__ void __launch_bounds__(kBuildThreads) build_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ ids,
                                                              uint64_t begin, uint64_t end, Bucket* __restrict__ table,
                                                              uint64_t n_buckets, int bits) {
  Tables t;
  t.arena = table; t.stride = 0; t.nr = 1; t.bits = bits; t.pid_base = 0; t.nb[0] = (uint32_t)n_buckets;
  const uint64_t stride = (uint64_t)gridDim.x * kBuildThreads;
  for (uint64_t row = begin + blockIdx.x * (uint64_t)kBuildThreads + threadIdx.x; row < end; row += stride)
    build_row(t, 0, ld_stream_u64(keys + row), ids ? ld_stream_u32(ids + row) : (uint32_t)row);
}

struct ProbeShared {
  uint2 stage[kProbeStage];
  uint64_t q_key[2][kProbeQueue];
  uint32_t q_pid[2][kProbeQueue], q_bkt[2][kProbeQueue], q_rem[2][kProbeQueue];
  uint8_t q_tab[2][kProbeQueue];
  uint32_t q_n[2];
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

// one continuation step: look at bucket bk for key, emit hits (per-thread smem atomic: continuations are rare)
__device__ __forceinline__ void probe_step(ProbeShared& sh, const Bucket* table, uint32_t bk, uint64_t key, uint32_t pid, olap_pair* out,
                                           uint64_t max_matches, unsigned long long* count) {
  const Bucket x = ld_bucket(table + bk);
  if (x.cnt >= 1 && x.k0 == key) stage_or_spill(sh.stage, atomicAdd(&sh.n, 1u), x.r0, pid, out, max_matches, count);
  if (x.cnt >= 2 && x.k1 == key) stage_or_spill(sh.stage, atomicAdd(&sh.n, 1u), x.r1, pid, out, max_matches, count);
}

// Probe one tile of kProbeTile rows starting at `tile` (a multiple of 4 rows) against the tables in t; rows outside
// [begin, end) are masked.  Called by the whole CTA.  VEC: keys / ids may be read with 256 / 128-bit loads.
template <bool VEC>
__device__ __forceinline__ void probe_tile(ProbeShared& sh, const Tables& t, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ ids,
                                           uint64_t begin, uint64_t end, uint64_t tile, olap_pair* __restrict__ out, uint64_t max_matches,
                                           unsigned long long* __restrict__ count) {
  const int tid = threadIdx.x, lane = tid & 31;
  const unsigned lt = (1u << lane) - 1;
  if (tid == 0) { sh.n = 0; sh.q_n[0] = 0; sh.q_n[1] = 0; }
  __syncthreads();
#pragma unroll 1
  for (int sub = 0; sub < kProbeSub; sub++) {
    const uint64_t row0 = tile + (uint64_t)(sub * kProbeThreads + tid) * kProbeRows;
    uint64_t key[kProbeRows];
    uint32_t pid[kProbeRows];
    bool live[kProbeRows];
    if (VEC && row0 >= begin && row0 + 3 < end) {
      u64x4 v = ld_stream_u64x4(keys + row0);
      key[0] = v.a; key[1] = v.b; key[2] = v.c; key[3] = v.d;
      if (ids) { uint4 w = ld_stream_u32x4(ids + row0); pid[0] = w.x; pid[1] = w.y; pid[2] = w.z; pid[3] = w.w; }
      else { pid[0] = (uint32_t)row0; pid[1] = (uint32_t)row0 + 1; pid[2] = (uint32_t)row0 + 2; pid[3] = (uint32_t)row0 + 3; }
      live[0] = live[1] = live[2] = live[3] = true;
    } else {
#pragma unroll
      for (int j = 0; j < kProbeRows; j++) {
        const uint64_t row = row0 + j;
        live[j] = row >= begin && row < end;
        key[j] = live[j] ? ld_stream_u64(keys + row) : 0;
        pid[j] = live[j] ? (ids ? ld_stream_u32(ids + row) : (uint32_t)row) : 0;
      }
    }
#pragma unroll
    for (int j = 0; j < kProbeRows; j++) pid[j] += t.pid_base;
    uint32_t b[kProbeRows];
    int tab[kProbeRows];
    Bucket e[kProbeRows];
#pragma unroll
    for (int j = 0; j < kProbeRows; j++) {  // all home-bucket loads are in flight before any is consumed
      tab[j] = table_of(t.pb, t.nr, row0 + j);
      b[j] = (uint32_t)home_bucket(key[j], t.bits, t.nb[tab[j]]);
      e[j].cnt = 0; e[j].maxdisp = 0;
      if (live[j]) e[j] = ld_bucket(t.arena + (uint64_t)tab[j] * t.stride + b[j]);
    }
#pragma unroll
    for (int j = 0; j < kProbeRows; j++) {
      const Bucket x = e[j];
      // every lane is here (convergent): hits are ranked with two ballots and appended with ONE shared-memory atomic
      // per warp -- a same-address ATOMS costs ~5 cycles, one per hit would dominate the tile
      const bool h0 = x.cnt >= 1 && x.k0 == key[j], h1 = x.cnt >= 2 && x.k1 == key[j];
      const unsigned m0 = __ballot_sync(0xffffffffu, h0), m1 = __ballot_sync(0xffffffffu, h1);
      if (m0 | m1) {
        const uint32_t n0 = __popc(m0);
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&sh.n, n0 + __popc(m1));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (h0) stage_or_spill(sh.stage, base + __popc(m0 & lt), x.r0, pid[j], out, max_matches, count);
        if (h1) stage_or_spill(sh.stage, base + n0 + __popc(m1 & lt), x.r1, pid[j], out, max_matches, count);
      }
      // a key with this home bucket was displaced: queue the row for the continuation phases
      const bool cont = x.maxdisp != 0;
      const unsigned mc = __ballot_sync(0xffffffffu, cont);
      if (mc) {
        uint32_t qb = 0;
        if (lane == 0) qb = atomicAdd(&sh.q_n[0], __popc(mc));
        qb = __shfl_sync(0xffffffffu, qb, 0);
        if (cont) {
          const uint32_t qi = qb + __popc(mc & lt);
          const uint32_t nbk = t.nb[tab[j]];
          uint32_t nb1 = b[j] + 1;
          if (nb1 == nbk) nb1 = 0;
          if (qi < kProbeQueue) {
            sh.q_key[0][qi] = key[j]; sh.q_pid[0][qi] = pid[j]; sh.q_bkt[0][qi] = nb1; sh.q_rem[0][qi] = x.maxdisp; sh.q_tab[0][qi] = (uint8_t)tab[j];
          } else {  // queue full (pathological skew): walk the chain right here
            const Bucket* table = t.arena + (uint64_t)tab[j] * t.stride;
            for (uint32_t r = x.maxdisp; r; r--) { probe_step(sh, table, nb1, key[j], pid[j], out, max_matches, count); if (++nb1 == nbk) nb1 = 0; }
          }
        }
      }
    }
  }
  __syncthreads();
  // continuation phases: every queued row looks at its next bucket; rows with displacement left re-queue
  for (int cur = 0;; cur ^= 1) {
    const uint32_t qn = min(sh.q_n[cur], (uint32_t)kProbeQueue);
    if (qn == 0) break;
    for (uint32_t i = tid; i < qn; i += kProbeThreads) {
      const uint64_t k = sh.q_key[cur][i];
      const uint32_t p = sh.q_pid[cur][i], rem = sh.q_rem[cur][i] - 1;
      const int r = sh.q_tab[cur][i];
      const uint32_t nbk = t.nb[r];
      const Bucket* table = t.arena + (uint64_t)r * t.stride;
      uint32_t bk = sh.q_bkt[cur][i];
      probe_step(sh, table, bk, k, p, out, max_matches, count);
      if (rem) {
        if (++bk == nbk) bk = 0;
        const uint32_t qi = atomicAdd(&sh.q_n[cur ^ 1], 1u);
        if (qi < kProbeQueue) { sh.q_key[cur ^ 1][qi] = k; sh.q_pid[cur ^ 1][qi] = p; sh.q_bkt[cur ^ 1][qi] = bk; sh.q_rem[cur ^ 1][qi] = rem; sh.q_tab[cur ^ 1][qi] = (uint8_t)r; }
        else for (uint32_t q = rem; q; q--) { probe_step(sh, table, bk, k, p, out, max_matches, count); if (++bk == nbk) bk = 0; }
      }
    }
    __syncthreads();
    if (tid == 0) sh.q_n[cur] = 0;
    __syncthreads();
  }
  __syncthreads();
  const uint32_t n = min(sh.n, (uint32_t)kProbeStage);
  if (tid == 0 && n) sh.base = atomicAdd(count, (unsigned long long)n);  // one global atomic per tile
  __syncthreads();
  if (out && n) {
    const unsigned long long base = sh.base;
    for (uint32_t i = tid; i < n; i += kProbeThreads) {
      const unsigned long long pos = base + i;
      if (pos < max_matches) st_stream_u64(out + pos, (uint64_t)sh.stage[i].x | ((uint64_t)sh.stage[i].y << 32));
    }
  }
  __syncthreads();
}

template <bool VEC>
__global__ void __launch_bounds__(kProbeThreads, 4) probe_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ ids,
                                                                 uint64_t begin, uint64_t end, const Bucket* __restrict__ table,
                                                                 uint64_t n_buckets, int bits, uint32_t pid_base,
                                                                 olap_pair* __restrict__ out, uint64_t max_matches,
                                                                 unsigned long long* __restrict__ count) {
  __shared__ ProbeShared sh;
  __shared__ Tables t;
  if (threadIdx.x == 0) { t.arena = const_cast<Bucket*>(table); t.stride = 0; t.nr = 1; t.bits = bits; t.pid_base = pid_base; t.nb[0] = (uint32_t)n_buckets; t.pb[0] = 0; t.pb[1] = 0; }
  __syncthreads();
  const uint64_t first = begin & ~3ULL;  // tiles start 4-row aligned so vector loads stay aligned; rows < begin are masked
  for (uint64_t tile = first + blockIdx.x * (uint64_t)kProbeTile; tile < end; tile += (uint64_t)gridDim.x * kProbeTile)
    probe_tile<VEC>(sh, t, keys, ids, begin, end, tile, out, max_matches, count);
}

// All partitions in ONE cooperative persistent kernel, R partitions (tables) per round:
//   clear the R tables -> grid.sync -> build them -> grid.sync -> probe them -> grid.sync
// Every phase runs across the whole GPU over thousands of rows/tiles (instead of 3 small launches per partition whose
// ~1.5 waves of CTAs were latency-chain-bound), while the R tables (<= ~64 MB) stay L2-resident between build and probe.
struct RoundsArgs {
  const uint64_t* bkeys; const uint32_t* bids; const uint64_t* pkeys; const uint32_t* pids;
  const uint32_t* bbegin; const uint32_t* pbegin;  // device, [P + 1]
  uint32_t P; int bits; int R;
  Bucket* arena; uint64_t stride;
  olap_pair* out; uint64_t max_matches; unsigned long long* count;
};

__global__ void __launch_bounds__(kProbeThreads, 4) join_rounds_kernel(RoundsArgs a) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  __shared__ ProbeShared sh;
  __shared__ Tables t;
  const int tid = threadIdx.x;
  const uint64_t gtid = blockIdx.x * (uint64_t)kProbeThreads + tid, gthreads = (uint64_t)gridDim.x * kProbeThreads;
  for (uint32_t p0 = 0; p0 < a.P; p0 += a.R) {
    const int nr = (int)min((uint32_t)a.R, a.P - p0);
    if (tid == 0) { t.arena = a.arena; t.stride = a.stride; t.nr = nr; t.bits = a.bits; t.pid_base = 0; }
    if (tid <= nr) { t.bb[tid] = a.bbegin[p0 + tid]; t.pb[tid] = a.pbegin[p0 + tid]; }
    if (tid < nr) t.nb[tid] = max(a.bbegin[p0 + tid + 1] - a.bbegin[p0 + tid], 64u);
    __syncthreads();
    // 1. clear (a pure write: the lines are allocated in L2 without a DRAM read)
    for (int r = 0; r < nr; r++) {
      uint4* tb = reinterpret_cast<uint4*>(a.arena + (uint64_t)r * a.stride);
      const uint64_t n16 = (uint64_t)t.nb[r] * 2;
      for (uint64_t i = gtid; i < n16; i += gthreads) tb[i] = make_uint4(0, 0, 0, 0);
    }
    __threadfence();
    grid.sync();
    // 2. build
    for (uint64_t row = t.bb[0] + gtid; row < t.bb[nr]; row += gthreads)
      build_row(t, table_of(t.bb, nr, row), ld_stream_u64(a.bkeys + row), ld_stream_u32(a.bids + row));
    __threadfence();
    grid.sync();
    // 3. probe
    const uint64_t begin = t.pb[0], end = t.pb[nr];
    for (uint64_t tile = (begin & ~3ULL) + blockIdx.x * (uint64_t)kProbeTile; tile < end; tile += (uint64_t)gridDim.x * kProbeTile)
      probe_tile<true>(sh, t, a.pkeys, a.pids, begin, end, tile, a.out, a.max_matches, a.count);
    grid.sync();  // the tables are reused by the next round
  }
}

static uint64_t pow2_at_least(uint64_t x) { uint64_t p = 1; while (p < x) p <<= 1; return p; }

// table arena: grown on demand, kept across calls; every table is cleared on its stream right before its build
static olap_status ensure_arena(olap_ctx* ctx, uint64_t bytes) {
  if (ctx->table_arena_bytes >= bytes) return OLAP_OK;
  if (ctx->table_arena) { OLAP_CUDA(cudaDeviceSynchronize()); OLAP_CUDA(cudaFree(ctx->table_arena)); ctx->table_arena = nullptr; ctx->table_arena_bytes = 0; }
  OLAP_CUDA(cudaMalloc(&ctx->table_arena, bytes));
  ctx->table_arena_bytes = bytes;
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
  std::vector<uint32_t> hb;  // partition boundaries on the host
  uint64_t max_part = 0;
};

static olap_status build_side_free(BuildSide& b, cudaStream_t s) {
  if (b.table) { pool_free(b.ctx, b.table, s); b.table = nullptr; }
  if (b.pkb) { pool_free(b.ctx, b.pkb, s); b.pkb = nullptr; }
  if (b.pib) { pool_free(b.ctx, b.pib, s); b.pib = nullptr; }
  return OLAP_OK;
}

static olap_status build_side_prepare(olap_ctx* ctx, const uint64_t* bk, uint64_t nb, int strategy, cudaStream_t s, BuildSide* b) {
  b->ctx = ctx; b->strategy = strategy; b->nb = nb;
  if (nb == 0) return OLAP_OK;
  if (strategy == OLAP_JOIN_BROADCAST) {
    b->nbk = std::max<uint64_t>(nb, 64);  // one 2-slot bucket per build row: load factor 0.5
    OLAP_TRY(pool_alloc(ctx, b->nbk * 32, s, (void**)&b->table));
    OLAP_CUDA(cudaMemsetAsync(b->table, 0, b->nbk * 32, s));
    build_kernel<<<grid_rows(ctx, nb, kBuildThreads, 16), kBuildThreads, 0, s>>>(bk, nullptr, 0, nb, b->table, b->nbk, 0);
    count_launch(ctx, 2);
    phase_mark(ctx, s, "build");
    OLAP_CUDA(cudaGetLastError());
    return OLAP_OK;
  }
  // partitions of ~512 Ki build rows: a 16 MB table each, so a few tables in flight stay L2-resident
  const uint64_t target = env_u64("OLAP_JOIN_PART_ROWS", 512 * 1024);
  int bits = 1;
  while (bits < kMaxPartBits && ((nb + target - 1) / target) > (1ULL << bits)) bits++;
  b->bits = bits; b->P = 1u << bits;
  Scratch sc(ctx, s);
  uint32_t *hist = nullptr, *beg = nullptr;
  OLAP_TRY(pool_alloc(ctx, nb * 8 + 16, s, (void**)&b->pkb));
  OLAP_TRY(pool_alloc(ctx, nb * 4 + 16, s, (void**)&b->pib));
  OLAP_TRY(sc.get<uint32_t>((uint64_t)b->P * ctx->sm_count * kPartCtasPerSm, &hist));
  OLAP_TRY(sc.get<uint32_t>(b->P + 1, &beg));
  PartArgs a;
  memset(&a, 0, sizeof(a));
  a.bits = bits; a.keys = bk; a.n = nb; a.hist = hist; a.out_keys = b->pkb; a.out_ids = b->pib;
  OLAP_TRY(run_partition(ctx, a, 0, true, beg, s, "hist_build", "partition_build"));
  b->hb.resize(b->P + 1);
  OLAP_CUDA(cudaMemcpyAsync(b->hb.data(), beg, (b->P + 1) * 4, cudaMemcpyDeviceToHost, s));
  OLAP_CUDA(cudaStreamSynchronize(s));
  b->max_part = 0;
  for (uint32_t p = 0; p < b->P; p++) b->max_part = std::max<uint64_t>(b->max_part, b->hb[p + 1] - b->hb[p]);
  return OLAP_OK;
}

static olap_status build_side_probe(BuildSide& b, const uint64_t* pk, uint64_t np, uint32_t pid_base, olap_pair* out, uint64_t max_matches,
                                    unsigned long long* d_count, cudaStream_t s) {
  olap_ctx* ctx = b.ctx;
  if (b.nb == 0 || np == 0) return OLAP_OK;  // executor.rs:231-233: an empty side yields no rows
  if (b.strategy == OLAP_JOIN_BROADCAST) {
    if (aligned32(pk))
      probe_kernel<true><<<grid_rows(ctx, np, kProbeTile, 8), kProbeThreads, 0, s>>>(pk, nullptr, 0, np, b.table, b.nbk, 0, pid_base, out, max_matches, d_count);
    else
      probe_kernel<false><<<grid_rows(ctx, np, kProbeTile, 8), kProbeThreads, 0, s>>>(pk, nullptr, 0, np, b.table, b.nbk, 0, pid_base, out, max_matches, d_count);
    count_launch(ctx);
    phase_mark(ctx, s, "probe");
    OLAP_CUDA(cudaGetLastError());
    return OLAP_OK;
  }
  const uint32_t P = b.P;
  const int bits = b.bits;
  Scratch sc(ctx, s);
  uint64_t* pkp = nullptr;
  uint32_t *pip = nullptr, *hist = nullptr, *beg = nullptr;
  OLAP_TRY(sc.get<uint64_t>(np, &pkp));
  OLAP_TRY(sc.get<uint32_t>(np, &pip));
  OLAP_TRY(sc.get<uint32_t>((uint64_t)P * ctx->sm_count * kPartCtasPerSm, &hist));
  OLAP_TRY(sc.get<uint32_t>(P + 1, &beg));
  PartArgs a;
  memset(&a, 0, sizeof(a));
  a.bits = bits; a.keys = pk; a.n = np; a.hist = hist; a.out_keys = pkp; a.out_ids = pip;
  OLAP_TRY(run_partition(ctx, a, 0, true, beg, s, "hist_probe", "partition_probe"));
  std::vector<uint32_t> hp(P + 1);
  OLAP_CUDA(cudaMemcpyAsync(hp.data(), beg, (P + 1) * 4, cudaMemcpyDeviceToHost, s));
  OLAP_CUDA(cudaStreamSynchronize(s));
  const uint32_t* hb = b.hb.data();
  const uint64_t nbk = std::max<uint64_t>(b.max_part, 64);  // buckets per table: exactly 32 B per build row of the largest partition
  const uint64_t dbg = env_u64("OLAP_JOIN_DEBUG", 0);  // timing decomposition only: 1 skips probes, 2 skips builds, 4 marks every kernel
  if (env_u64("OLAP_JOIN_ROUND", 0) && !dbg) {
    // opt-in (OLAP_JOIN_ROUND=R): one cooperative persistent kernel, R tables per round.  Measured 23 ms vs 13.8 ms for the
    // stream-overlapped per-partition kernels at C2: the same dependent-latency chains, now exposed between grid syncs.
    int R = (int)env_u64("OLAP_JOIN_ROUND", 4);
    if (R > kMaxRound) R = kMaxRound;
    if (R < 1) R = 1;
    OLAP_TRY(ensure_arena(ctx, (uint64_t)R * nbk * 32));
    uint32_t* dbeg = nullptr;
    OLAP_TRY(sc.get<uint32_t>(P + 1, &dbeg));
    OLAP_CUDA(cudaMemcpyAsync(dbeg, hb, (P + 1) * 4, cudaMemcpyHostToDevice, s));
    RoundsArgs ra;
    ra.bkeys = b.pkb; ra.bids = b.pib; ra.pkeys = pkp; ra.pids = pip; ra.bbegin = dbeg; ra.pbegin = beg;
    ra.P = P; ra.bits = bits; ra.R = R; ra.arena = (Bucket*)ctx->table_arena; ra.stride = nbk;
    ra.out = out; ra.max_matches = max_matches; ra.count = d_count;
    int occ = 0;
    OLAP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, join_rounds_kernel, kProbeThreads, 0));
    if (occ < 1) { set_error("join_rounds_kernel does not fit on an SM"); return OLAP_ERR_CUDA; }
    void* kargs[] = {&ra};
    OLAP_CUDA(cudaLaunchCooperativeKernel((void*)join_rounds_kernel, dim3((unsigned)(occ * ctx->sm_count)), dim3(kProbeThreads), kargs, 0, s));
    count_launch(ctx);
    phase_mark(ctx, s, "build_probe");
    OLAP_CUDA(cudaGetLastError());
    OLAP_CUDA(cudaStreamSynchronize(s));  // dbeg / hb stay valid until the kernel is done
    return OLAP_OK;
  }
  int S = (int)std::min<uint64_t>(ctx->streams.size(), env_u64("OLAP_JOIN_STREAMS", 3));
  if (dbg & 4) S = 1;
  if (S < 1) S = 1;
  if (S > 8) S = 8;
  OLAP_TRY(ensure_arena(ctx, (uint64_t)S * nbk * 32));

  // fork: tables of different partitions are built and probed on S side streams so their ramps and tails overlap
  OLAP_CUDA(cudaEventRecord(ctx->fork_ev, s));
  for (int k = 0; k < S; k++) OLAP_CUDA(cudaStreamWaitEvent(ctx->streams[k], ctx->fork_ev, 0));
  uint64_t launches = 0;
  for (uint32_t p = 0; p < P; p++) {
    const uint64_t b0 = hb[p], b1 = hb[p + 1], p0 = hp[p], p1 = hp[p + 1];
    if (b1 == b0 || p1 == p0) continue;  // inner join: an empty side yields nothing
    const int k = p % S;
    cudaStream_t st = (dbg & 4) ? s : ctx->streams[k];
    Bucket* table = (Bucket*)ctx->table_arena + (uint64_t)k * nbk;
    const uint64_t tb = std::max<uint64_t>(b1 - b0, 64);  // this partition's table: one bucket per build row
    if (!(dbg & 2)) {
      OLAP_CUDA(cudaMemsetAsync(table, 0, tb * 32, st));
      if (dbg & 4) phase_mark(ctx, st, "p_clear");
      build_kernel<<<grid_rows(ctx, b1 - b0, kBuildThreads, 16), kBuildThreads, 0, st>>>(b.pkb, b.pib, b0, b1, table, tb, bits);
      if (dbg & 4) phase_mark(ctx, st, "p_build");
    }
    if (!(dbg & 1))
      probe_kernel<true><<<grid_rows(ctx, p1 - p0, kProbeTile, 8), kProbeThreads, 0, st>>>(pkp, pip, p0, p1, table, tb, bits, pid_base, out,
                                                                                        max_matches, d_count);
    if (dbg & 4) phase_mark(ctx, st, "p_probe");
    launches += 3;
  }
  count_launch(ctx, launches);
  for (int k = 0; k < S; k++) {
    OLAP_CUDA(cudaEventRecord(ctx->join_ev[k], ctx->streams[k]));
    OLAP_CUDA(cudaStreamWaitEvent(s, ctx->join_ev[k], 0));
  }
  phase_mark(ctx, s, "build_probe");
  OLAP_CUDA(cudaGetLastError());
  return OLAP_OK;
}

// Partition pass of the partitioned group-by (groupby.cu): (key, up to 4 value columns) with the filter fused in.
olap_status group_by_partition_rows(olap_ctx* ctx, const uint64_t* keys, uint64_t n, const TermPack& terms, int n_vcols,
                                    const uint64_t* const* vcols, int bits, uint64_t* out_keys, uint64_t* const* out_vcols,
                                    uint32_t* part_begin, cudaStream_t s) {
  if (n_vcols > kMaxPay) { set_error("partitioned group-by carries at most %d value columns", kMaxPay); return OLAP_ERR_UNSUPPORTED; }
  Scratch sc(ctx, s);
  uint32_t* hist = nullptr;
  OLAP_TRY(sc.get<uint32_t>((uint64_t)(1u << bits) * ctx->sm_count * kPartCtasPerSm, &hist));
  PartArgs a;
  memset(&a, 0, sizeof(a));
  a.bits = bits; a.keys = keys; a.n = n; a.hist = hist; a.out_keys = out_keys; a.terms = terms;
  for (int q = 0; q < n_vcols; q++) { a.pay_in[q] = vcols[q]; a.pay_out[q] = out_vcols[q]; }
  return run_partition(ctx, a, n_vcols, false, part_begin, s);
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
    st = build_side_prepare(ctx, (const uint64_t*)build_keys, n_build, strategy, s, &b);
    if (st == OLAP_OK) st = build_side_probe(b, (const uint64_t*)probe_keys, n_probe, 0, out_pairs, max_matches, d_count, s);
    build_side_free(b, s);
  }
  phase_end(ctx, s);
  OLAP_TRY(st);
  return finish_count(ctx, d_count, host_count, out_pairs, max_matches, s);
}

}  // extern "C"

struct olap_join_build { BuildSide b; };

extern "C" {

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
  olap_status st = build_side_prepare(ctx, (const uint64_t*)build_keys, n_build, strategy, s, &h->b);
  if (st != OLAP_OK) { build_side_free(h->b, s); delete h; return st; }
  *out = h;
  return OLAP_OK;
}

olap_status olap_join_probe(olap_join_build* build, const int64_t* probe_keys, uint64_t n_probe, uint32_t probe_row_base, olap_pair* out_pairs,
                            uint64_t max_matches, uint64_t* out_count_dptr, uint64_t* host_count, void* stream) {
  OLAP_REQUIRE(build, "build handle is NULL");
  OLAP_REQUIRE(n_probe == 0 || probe_keys, "probe_keys is NULL");
  OLAP_REQUIRE(n_probe < 0xFFFFFFFFULL && (uint64_t)probe_row_base + n_probe < 0xFFFFFFFFULL, "row ids are u32");
  olap_ctx* ctx = build->b.ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  OLAP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream_of(stream);
  if (!out_pairs) max_matches = 0;
  Scratch sc(ctx, s);
  unsigned long long* d_count = (unsigned long long*)out_count_dptr;
  if (!d_count) OLAP_TRY(sc.get<unsigned long long>(1, &d_count));
  OLAP_CUDA(cudaMemsetAsync(d_count, 0, 8, s));
  OLAP_TRY(build_side_probe(build->b, (const uint64_t*)probe_keys, n_probe, probe_row_base, out_pairs, max_matches, d_count, s));
  return finish_count(ctx, d_count, host_count, out_pairs, max_matches, s);
}

olap_status olap_join_build_destroy(olap_join_build* build) {
  if (!build) return OLAP_OK;
  olap_ctx* ctx = build->b.ctx;
  {
    std::lock_guard<std::mutex> g(ctx->mu);
    cudaSetDevice(ctx->device);
    build_side_free(build->b, ctx->streams[0]);
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

uint32_t olap_shuffle_dest_of(int64_t key, int32_t bits) {
  if (bits <= 0) return 0;
  return (uint32_t)(mix64((uint64_t)key ^ kShuffleSalt) >> (64 - bits));
}

}  // extern "C"
