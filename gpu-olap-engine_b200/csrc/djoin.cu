// Direct-address join over key RANGES with shared-memory tables: the fast path of the radix hash join for dense, unique build
// keys (surrogate keys of a dimension table; the benchmark's permutation of 0..n-1).  Replaces radix_partition_kernel,
// build_hash_table_kernel and probe_hash_table_kernel (gpu-kernel-compiler/kernels/join_kernels.cuh:45-166) on that path.
//
// Round-1 finding this design answers (profiles/r01_ncu_join_v6_details.txt): with L2-resident tables every probe row costs
// one 32-byte sector gather for 4 useful bytes -- 537 M sectors, L2 63 % / L1TEX 66 % busy, 41 % of HBM peak -- and both
// passes move 12 bytes per row (8-byte key + 4-byte id).  Here
//   * rows travel as ONE packed 8-byte word {key - min : u32, row id : u32} (the range check already proved the offset fits),
//     so each pass reads / writes 8 bytes per row instead of 12;
//   * the partition pass fans out to ranges of 2^15 keys (up to 8192 partitions) in ONE pass: an 8192-row tile is ranked with
//     shared-memory atomics, staged in partition order and written as short runs; a run is reserved in its partition's
//     fixed-capacity region with one global atomic per (tile, partition), so runs of different CTAs land next to each other
//     and the L2 (126 MB, a few open lines per partition) merges them into full lines before they reach HBM;
//   * a partition's table is then 2^15 x 4-byte row ids = 128 KB of SHARED memory: the probe's random accesses are LDS
//     bank lookups, the L2 only streams.  One CTA per SM, partitions handed out by a ticket.
#include "common.cuh"
#include "djoin.cuh"

#include <algorithm>

namespace olap {

constexpr int kRShift = 15;                       // log2 keys per range partition = entries of one shared-memory table
constexpr uint32_t kRKeys = 1u << kRShift;
constexpr uint32_t kREmpty = 0xFFFFFFFFu;
constexpr int kRThreads = 256;                    // one partition counter per thread: a pass fans out to <= 256 partitions
constexpr int kRRows = 16;                        // rows per thread per tile
constexpr int kRTile = kRThreads * kRRows;        // 4096 rows ranked / staged per CTA iteration
constexpr int kRFineBits = 6;                     // second level: every coarse region is split into 64 ranges of 2^15 keys
constexpr uint32_t kRMaxParts = 256u << kRFineBits;
constexpr int kHotSlots = 2048;                   // open-addressing slots of the heavy-hitter table (<= 1024 keys)

// Why two levels (scratch/part_bench.cu on a B200, 2^29 rows, 16 B/row of traffic): ONE pass to 4096 partitions is bound by its
// global atomics -- one per (tile, partition), i.e. one per 2 rows with 8 K-row tiles: 6.1 ms (3.7 ms with 16 K-row tiles; 1.4 ms
// with the atomics and the write-out removed) -- while a 64-way pass runs at 1.53 ms (5.6 TB/s, 85 % of the measured copy peak)
// and the 64 x 64-way second level at 1.32 ms (6.5 TB/s: each tile's runs land in one 64-region neighbourhood): 2.85 ms together.
enum { kInRaw = 0, kInRegionsCarry = 1, kInRegionsPos = 2 };

struct RPartArgs {
  // kInRaw: int64 keys; offset = key - rmin, rows with offset >= rrange are dropped (cannot match); row id = id_base + row
  const uint64_t* in; uint64_t n; uint32_t id_base;
  uint64_t rmin, rrange;
  // kInRegions*: the fixed-capacity regions a previous pass (or the multi-GPU shuffle) filled with packed words
  // {offset : u32 hi, id : u32 lo}; region g holds min(reg_cur[g], reg_cap) rows, all of partitions
  // [(g & reg_mask) << reg_bits, ((g & reg_mask) + 1) << reg_bits).  Carry: the id is kept; Pos: the id becomes the row's
  // position g * reg_cap + row (multi-GPU: pairs then index the receive buffer, whose low words hold the source row ids)
  const uint32_t* reg_cur; uint32_t reg_cap, n_reg, reg_mask; int reg_bits;
  int shift;               // partition = (offset >> shift) - first partition of the region
  uint32_t cap;            // partition p owns out[p * cap, (p + 1) * cap)
  uint32_t* cursor;        // rows handed out per partition (may exceed cap: overflow)
  uint64_t* out;
  uint32_t* flags;
  int oob_is_error;        // build side: a key outside the range means the remembered key range is stale
  // heavy hitters (Zipf-skewed probe side, kInRaw only): a probe row whose key is in the hot table never enters a partition --
  // its pair (hot_ids[slot], row) is emitted right here.  The table holds only keys whose build row was found (hot_resolve_kernel).
  const uint32_t* hot_keys;   // [kHotSlots] offset + 1, 0 = empty slot; NULL = no hot keys
  const uint32_t* hot_ids;    // [kHotSlots] build row of the slot's key
  olap_pair* pairs; uint64_t max_matches; unsigned long long* count;
};

__device__ __forceinline__ uint32_t hot_slot_of(uint32_t d) { return (d * 0x9E3779B1u) >> (32 - 11); }
static_assert(kHotSlots == 2048, "hot_slot_of shifts for 2^11 slots");

// CTAs per SM: 4 for raw keys (64 registers); the region passes keep the carried id next to offset and rank -- under a 64-register
// cap ptxas serialises their 16 loads (2.18 ms instead of 1.35 ms for 2^29 rows) -- so they run 3 per SM with 80 registers
template <int INPUT, bool HOT>
__global__ void __launch_bounds__(kRThreads, INPUT == kInRaw ? 4 : 3) rpart_kernel(const __grid_constant__ RPartArgs a) {
  constexpr int R = kRRows, NW = kRThreads / 32;
  __shared__ __align__(16) uint64_t stage[kRTile];                    // rows in partition order (32 KB)
  __shared__ uint32_t cnt[kRThreads];                                 // tile histogram, then exclusive offsets
  __shared__ uint32_t gbase[kRThreads];                               // global position of staged row i = gbase[part] + i
  __shared__ uint32_t hotk[HOT ? kHotSlots : 1];                      // hot keys (8 KB); their build rows are read through L1
  __shared__ uint32_t s_warp[NW];
  __shared__ uint32_t s_total, s_over[2], s_hot[2];  // indexed by tile parity: reset one tile ahead, between barriers
  __shared__ unsigned long long s_hotbase;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  cnt[tid] = 0;
  if (HOT) {
    for (int i = tid; i < kHotSlots; i += kRThreads) hotk[i] = a.hot_keys[i];
  }
  if (tid == 0) { s_over[0] = s_over[1] = 0; s_hot[0] = s_hot[1] = 0; }
  __syncthreads();

  const uint32_t tiles_per_reg = INPUT == kInRaw ? 1u : (a.reg_cap + kRTile - 1) / kRTile;
  const uint32_t tiles = INPUT == kInRaw ? (uint32_t)((a.n + kRTile - 1) / kRTile) : tiles_per_reg * a.n_reg;
  int par = 1;
  for (uint32_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    const uint64_t* src = a.in;
    uint64_t n = a.n, row0 = (uint64_t)t * kRTile;
    uint32_t pbase = 0, idb = a.id_base;
    if (INPUT != kInRaw) {
      const uint32_t g = t / tiles_per_reg;
      src = a.in + (uint64_t)g * a.reg_cap;
      n = min(a.reg_cur[g], a.reg_cap);
      row0 = (uint64_t)(t - g * tiles_per_reg) * kRTile;
      pbase = (g & a.reg_mask) << a.reg_bits;
      idb = g * a.reg_cap;
      if (row0 >= n) continue;  // uniform over the CTA
    }
    par ^= 1;

    // phase 1: offset and partition of every row, rank inside the tile with one shared-memory atomic per row
    uint32_t d[R], rk[R];
    uint64_t key[R];
#pragma unroll
    for (int k = 0; k < R; k++) {
      const uint64_t row = row0 + (uint64_t)k * kRThreads + tid;
      key[k] = row < n ? ld_stream_u64(src + row) : 0;
    }
    bool oob = false;
#pragma unroll
    for (int k = 0; k < R; k++) {
      const uint64_t row = row0 + (uint64_t)k * kRThreads + tid;
      bool ok = row < n;
      if (INPUT != kInRaw) d[k] = (uint32_t)(key[k] >> 32);
      else {
        const uint64_t off = key[k] - a.rmin;
        if (off >= a.rrange) { oob = oob || ok; ok = false; }
        d[k] = (uint32_t)off;
      }
      rk[k] = kREmpty;
      if (HOT && ok) {  // heavy hitter?  (linear probing; the table is at most half full)
        uint32_t s = hot_slot_of(d[k]);
        for (;;) {
          const uint32_t hk = hotk[s];
          if (hk == 0) break;
          if (hk == d[k] + 1) { ok = false; rk[k] = 0x80000000u | s; atomicAdd(&s_hot[par], 1u); break; }
          s = (s + 1) & (kHotSlots - 1);
        }
      }
      if (ok) rk[k] = atomicAdd(&cnt[(d[k] >> a.shift) - pbase], 1u);
    }
    if (INPUT == kInRaw && oob && a.oob_is_error) a.flags[kFlagOutOfRange] = 1;
    __syncthreads();

    // phase 2: exclusive scan of the tile histogram (one partition per thread); every non-empty (tile, partition) run is
    // reserved in the partition's region with one global atomic, whose round trip overlaps the scan
    const uint32_t c = cnt[tid];
    const uint32_t got = c ? atomicAdd(&a.cursor[pbase + tid], c) : 0u;
    if (tid == 0) { s_over[par ^ 1] = 0; s_hot[par ^ 1] = 0; }  // the previous tile is done with them, the next one starts after a barrier
    uint32_t inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    {
      uint32_t wbase = 0;
#pragma unroll
      for (int w = 0; w < NW; w++) wbase += w < warp ? s_warp[w] : 0u;
      const uint32_t off = wbase + inc - c;
      cnt[tid] = off;
      gbase[tid] = (pbase + (uint32_t)tid) * a.cap + got - off;
      if (got + c > a.cap) { s_over[par] = 1; a.flags[kFlagOverflow] = 1; }
      if (tid == kRThreads - 1) s_total = off + c;
    }
    __syncthreads();

    // phase 3: scatter the tile into shared memory in partition order; hot probe rows go to the top of the stage, downwards
    const uint32_t total = s_total;
#pragma unroll
    for (int k = 0; k < R; k++) {
      if (rk[k] == kREmpty) continue;
      const uint32_t id = INPUT == kInRegionsCarry ? (uint32_t)key[k] : idb + (uint32_t)(row0 + (uint64_t)k * kRThreads + tid);
      if (HOT && (rk[k] & 0x80000000u)) {
        const uint32_t pos = kRTile - 1 - (atomicAdd(&s_hot[par], 0x10000u) >> 16);
        stage[pos] = ((uint64_t)id << 32) | __ldg(a.hot_ids + (rk[k] & (kHotSlots - 1)));
      } else {
        stage[cnt[(d[k] >> a.shift) - pbase] + rk[k]] = ((uint64_t)d[k] << 32) | id;
      }
    }
    __syncthreads();
    cnt[tid] = 0;  // the offsets are consumed: clear the histogram for the next tile

    // phase 4: every partition's run leaves with consecutive lanes on consecutive addresses
    if (!s_over[par]) {
      for (uint32_t i = tid; i < total; i += kRThreads) {
        const uint64_t v = stage[i];
        a.out[gbase[((uint32_t)(v >> 32) >> a.shift) - pbase] + i] = v;
      }
    } else {  // some region is full: rows past its capacity are dropped; the caller sees the flag and takes another path
      for (uint32_t i = tid; i < total; i += kRThreads) {
        const uint64_t v = stage[i];
        const uint32_t part = ((uint32_t)(v >> 32) >> a.shift) - pbase, dst = gbase[part] + i;
        if (dst - (pbase + part) * a.cap < a.cap) a.out[dst] = v;
      }
    }
    if (HOT) {
      const uint32_t nh = s_hot[par] & 0xFFFFu;
      if (nh) {
        if (tid == 0) s_hotbase = atomicAdd(a.count, (unsigned long long)nh);
        __syncthreads();
        const unsigned long long base = s_hotbase;
        for (uint32_t i = tid; i < nh; i += kRThreads) {
          const uint64_t v = stage[kRTile - 1 - i];  // {probe row : hi, build row : lo}
          if (a.pairs && base + i < a.max_matches) a.pairs[base + i] = olap_pair{(uint32_t)v, (uint32_t)(v >> 32)};
        }
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------- build + probe
struct DJoinArgs {
  const uint64_t* brows; const uint32_t* bcur; uint32_t bcap;
  const uint64_t* prows; const uint32_t* pcur; uint32_t pcap;
  uint32_t P;
  uint32_t* ticket;
  uint32_t* flags;
  olap_pair* out; uint64_t max_matches; unsigned long long* count;
};

constexpr int kDJThreads = 1024;
constexpr int kDJGroups = 4;                       // independent probe groups inside the CTA (named barriers)
constexpr int kDJGroupThreads = kDJThreads / kDJGroups;
constexpr int kDJRows = 8;                         // probe rows per thread per group tile
constexpr int kDJGroupTile = kDJGroupThreads * kDJRows;

__device__ __forceinline__ void group_bar(int g) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(kDJGroupThreads) : "memory"); }

__global__ void __launch_bounds__(kDJThreads, 1) djoin_kernel(const __grid_constant__ DJoinArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* table = reinterpret_cast<uint32_t*>(smem_raw);   // [kRKeys] build row id per key of this range
  __shared__ uint32_t s_part;
  __shared__ uint32_t s_n[kDJGroups];
  __shared__ unsigned long long s_base[kDJGroups];
  const int tid = threadIdx.x, lane = tid & 31;
  const int g = tid / kDJGroupThreads, gt = tid % kDJGroupThreads;
  if (tid < kDJGroups) s_n[tid] = 0;
  for (;;) {
    __syncthreads();  // the previous partition's probes are done with the table; s_part may be rewritten
    if (tid == 0) s_part = atomicAdd(a.ticket, 1u);
    __syncthreads();
    const uint32_t p = s_part;
    if (p >= a.P) break;
    const uint32_t nb = min(a.bcur[p], a.bcap), np = min(a.pcur[p], a.pcap);
    if (nb == 0 || np == 0) continue;  // inner join: an empty side yields nothing
    {
      const uint4 e = make_uint4(kREmpty, kREmpty, kREmpty, kREmpty);
      uint4* t4 = reinterpret_cast<uint4*>(table);
#pragma unroll
      for (int i = 0; i < (int)(kRKeys / 4 / kDJThreads); i++) t4[i * kDJThreads + tid] = e;
    }
    __syncthreads();
    {
      const uint64_t* rows = a.brows + (uint64_t)p * a.bcap;
      bool dup = false;
      for (uint32_t i0 = 0; i0 < nb; i0 += kDJThreads * 8) {
        uint64_t v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) { const uint32_t i = i0 + k * kDJThreads + tid; v[k] = i < nb ? ld_stream_u64(rows + i) : ~0ULL; }
#pragma unroll
        for (int k = 0; k < 8; k++)
          if (i0 + k * kDJThreads + tid < nb) dup = dup || atomicExch(&table[(uint32_t)(v[k] >> 32) & (kRKeys - 1)], (uint32_t)v[k]) != kREmpty;
      }
      if (dup) a.flags[kFlagDup] = 1;  // a repeated build key: this representation cannot hold it (the caller redoes the join with bucket tables)
    }
    __syncthreads();
    const uint64_t* rows = a.prows + (uint64_t)p * a.pcap;
    for (uint32_t t0 = (uint32_t)g * kDJGroupTile; t0 < np; t0 += kDJGroups * kDJGroupTile) {
      uint64_t v[kDJRows];
#pragma unroll
      for (int k = 0; k < kDJRows; k++) { const uint32_t i = t0 + k * kDJGroupThreads + gt; v[k] = i < np ? ld_stream_u64(rows + i) : ~0ULL; }
      uint32_t brow[kDJRows];
      uint32_t wtotal = 0;
#pragma unroll
      for (int k = 0; k < kDJRows; k++) {
        const uint32_t i = t0 + k * kDJGroupThreads + gt;
        brow[k] = i < np ? table[(uint32_t)(v[k] >> 32) & (kRKeys - 1)] : kREmpty;
        wtotal += __popc(__ballot_sync(0xffffffffu, brow[k] != kREmpty));
      }
      uint32_t wbase = 0;
      if (lane == 0 && wtotal) wbase = atomicAdd(&s_n[g], wtotal);
      wbase = __shfl_sync(0xffffffffu, wbase, 0);
      group_bar(g);
      if (gt == 0) {
        const uint32_t n = s_n[g];
        s_n[g] = 0;
        s_base[g] = n ? atomicAdd(a.count, (unsigned long long)n) : 0ULL;  // one global atomic per group tile
      }
      group_bar(g);
      if (a.out) {
        unsigned long long pos = s_base[g] + wbase;
#pragma unroll
        for (int k = 0; k < kDJRows; k++) {
          const uint32_t m = __ballot_sync(0xffffffffu, brow[k] != kREmpty);
          if (brow[k] != kREmpty) {
            const unsigned long long q = pos + __popc(m & ((1u << lane) - 1));
            if (q < a.max_matches) st_stream_u64(a.out + q, (uint64_t)brow[k] | ((uint64_t)(uint32_t)v[k] << 32));
          }
          pos += __popc(m);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------- heavy hitters
// Sample the probe keys, count the samples per key in a small global table, keep the keys seen at least `min_count` times (at
// most kHotSlots / 2 of them).  All on the device: no host round trip.
struct HotArgs {
  const uint64_t* keys; uint64_t n; uint64_t stride; uint32_t n_samples;
  uint64_t rmin, rrange;
  uint32_t* tab_key; uint32_t* tab_cnt; uint32_t tab_slots;   // pow2, zeroed
  uint32_t* hot_keys; uint32_t* hot_ids; uint32_t* hot_n;     // outputs
  uint32_t min_count;
};

__global__ void hot_sample_kernel(HotArgs h) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= h.n_samples) return;
  const uint64_t row = (uint64_t)i * h.stride + (mix64(i) % h.stride);
  if (row >= h.n) return;
  const uint64_t off = ld_stream_u64(h.keys + row) - h.rmin;
  if (off >= h.rrange) return;
  const uint32_t d = (uint32_t)off;
  uint32_t s = (d * 0x9E3779B1u) & (h.tab_slots - 1);
  for (;;) {
    const uint32_t old = atomicCAS(&h.tab_key[s], 0u, d + 1);
    if (old == 0 || old == d + 1) { atomicAdd(&h.tab_cnt[s], 1u); return; }
    s = (s + 1) & (h.tab_slots - 1);
  }
}

// one CTA: raise the threshold until at most kHotSlots / 2 keys qualify, then build the open-addressing hot table
__global__ void __launch_bounds__(1024) hot_select_kernel(HotArgs h) {
  __shared__ uint32_t s_cnt;
  __shared__ uint32_t s_thr;
  uint32_t thr = h.min_count;
  for (;;) {
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    uint32_t mine = 0;
    for (uint32_t i = threadIdx.x; i < h.tab_slots; i += blockDim.x) mine += h.tab_key[i] && h.tab_cnt[i] >= thr;
    if (mine) atomicAdd(&s_cnt, mine);
    __syncthreads();
    if (threadIdx.x == 0) s_thr = s_cnt <= kHotSlots / 2 ? 0 : thr + (thr + 1) / 2;
    __syncthreads();
    if (s_thr == 0) break;
    thr = s_thr;
    __syncthreads();
  }
  for (uint32_t i = threadIdx.x; i < h.tab_slots; i += blockDim.x) {
    const uint32_t k = h.tab_key[i];
    if (k && h.tab_cnt[i] >= thr) {
      uint32_t s = hot_slot_of(k - 1);
      while (atomicCAS(&h.hot_keys[s], 0u, k) != 0) s = (s + 1) & (kHotSlots - 1);
      atomicAdd(h.hot_n, 1u);
    }
  }
}

// The build row of every hot key: one CTA per slot scans the key's range partition of the partitioned build side (<= 2^15
// rows).  A hot key without a build row cannot match; its slot becomes a tombstone (offset + 1 = 2^32 - 1, which no row of a
// range this path accepts can have) so that probe rows carrying it are neither emitted nor break a probing chain.
__global__ void __launch_bounds__(256) hot_resolve_kernel(uint32_t* hot_keys, uint32_t* hot_ids, const uint64_t* brows, const uint32_t* bcur,
                                                          uint32_t bcap) {
  __shared__ uint32_t s_found;
  const uint32_t slot = blockIdx.x, k = hot_keys[slot];
  if (k == 0) return;
  const uint32_t d = k - 1, p = d >> kRShift, n = min(bcur[p], bcap);
  if (threadIdx.x == 0) s_found = kREmpty;
  __syncthreads();
  const uint64_t* rows = brows + (uint64_t)p * bcap;
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    const uint64_t v = rows[i];
    if ((uint32_t)(v >> 32) == d) s_found = (uint32_t)v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    hot_ids[slot] = s_found;
    if (s_found == kREmpty) hot_keys[slot] = 0xFFFFFFFFu;
  }
}

template <int INPUT, bool HOT>
static olap_status launch_rpart_t(olap_ctx* ctx, const RPartArgs& a, unsigned grid, cudaStream_t s) {
  rpart_kernel<INPUT, HOT><<<grid, kRThreads, 0, s>>>(a);
  count_launch(ctx);
  return OLAP_OK;
}

// grid_limit: CTAs to use at most (0 = fill the GPU); the multi-GPU path runs this next to a shuffle kernel
static olap_status launch_rpart(olap_ctx* ctx, const RPartArgs& a, int input, unsigned grid_limit, cudaStream_t s) {
  const uint64_t tiles = input == kInRaw ? (a.n + kRTile - 1) / kRTile : (uint64_t)((a.reg_cap + kRTile - 1) / kRTile) * a.n_reg;
  if (!tiles) return OLAP_OK;
  uint64_t grid = (uint64_t)ctx->sm_count * (input == kInRaw ? 4 : 3);
  if (grid_limit && grid > grid_limit) grid = grid_limit;
  if (grid > tiles) grid = tiles;
  const bool hot = a.hot_keys != nullptr;
  if (input == kInRaw) return hot ? launch_rpart_t<kInRaw, true>(ctx, a, (unsigned)grid, s) : launch_rpart_t<kInRaw, false>(ctx, a, (unsigned)grid, s);
  if (input == kInRegionsCarry) return launch_rpart_t<kInRegionsCarry, false>(ctx, a, (unsigned)grid, s);
  return launch_rpart_t<kInRegionsPos, false>(ctx, a, (unsigned)grid, s);
}

uint32_t djoin_partitions(uint64_t rrange) { return (uint32_t)((rrange + kRKeys - 1) >> kRShift); }
bool djoin_eligible(uint64_t rrange) { return rrange > 0 && rrange < 0xFFFFFFF0ULL && djoin_partitions(rrange) <= kRMaxParts; }
uint32_t djoin_region_keys_log2() { return kRShift; }

// capacity of one region: slack_pct % above the mean plus half a tile, a multiple of 4 rows (32-byte aligned regions)
static uint32_t region_cap(uint64_t n, uint32_t P, uint32_t slack_pct) {
  const uint64_t mean = (n + P - 1) / P;
  const uint64_t cap = mean + mean * slack_pct / 100 + 2048;
  return (uint32_t)((cap + 3) & ~3ULL);
}

static uint32_t padded_parts(uint32_t P) { return (P + 255) / 256 * 256 + 256; }  // the scan reads one cursor per thread past the last partition

// Rows -> P = ceil(rrange / 2^15) fixed-capacity regions of `cap` rows (cursor[p] rows in region p; > cap: overflow, flagged).
// Raw keys with P > 256 take two levels: 64-key-range-wide coarse regions first (scratch), then 64 ranges out of each.
// unique = build side: the coarse capacity is exact (a coarse region spans 2^21 keys), no slack.
static olap_status partition_rows(olap_ctx* ctx, const RInput& in, int64_t rmin, uint64_t rrange, uint32_t P, uint32_t cap, uint32_t* cursor,
                                  uint64_t* out, uint32_t* flags, bool unique, uint32_t slack_pct, const RPartArgs* hot, unsigned grid_limit,
                                  cudaStream_t s, Scratch& sc, const char* mark_coarse, const char* mark_fine) {
  RPartArgs a;
  memset(&a, 0, sizeof(a));
  a.rmin = (uint64_t)rmin; a.rrange = rrange; a.flags = flags; a.oob_is_error = unique ? 1 : 0;
  if (in.regions) {  // multi-GPU: the shuffle already did the coarse level
    if ((1u << in.reg_bits) > (uint32_t)kRThreads) { set_error("shuffle regions span more than 256 key ranges"); return OLAP_ERR_BAD_ARG; }
    a.in = in.regions; a.reg_cur = in.reg_cur; a.reg_cap = in.reg_cap; a.n_reg = in.n_reg; a.reg_mask = in.reg_mask; a.reg_bits = in.reg_bits;
    a.shift = kRShift; a.cap = cap; a.cursor = cursor; a.out = out;
    OLAP_TRY(launch_rpart(ctx, a, kInRegionsPos, grid_limit, s));
    if (mark_fine) phase_mark(ctx, s, mark_fine);
    return OLAP_OK;
  }
  a.in = in.keys; a.n = in.n; a.id_base = in.id_base;
  if (hot) { a.hot_keys = hot->hot_keys; a.hot_ids = hot->hot_ids; a.pairs = hot->pairs; a.max_matches = hot->max_matches; a.count = hot->count; }
  if (P <= (uint32_t)kRThreads) {
    a.shift = kRShift; a.cap = cap; a.cursor = cursor; a.out = out;
    OLAP_TRY(launch_rpart(ctx, a, kInRaw, grid_limit, s));
    if (mark_fine) phase_mark(ctx, s, mark_fine);
    return OLAP_OK;
  }
  const uint32_t C = (P + (1u << kRFineBits) - 1) >> kRFineBits;
  const uint32_t cap1 = unique ? (kRKeys << kRFineBits) : region_cap(in.n, C, slack_pct);
  if ((uint64_t)C * cap1 >= 0xFFFFFFFFULL) { set_error("partition regions exceed 2^32 rows"); return OLAP_ERR_UNSUPPORTED; }
  uint64_t* coarse = nullptr;
  uint32_t* cur1 = nullptr;
  OLAP_TRY(sc.get<uint64_t>((uint64_t)C * cap1, &coarse));
  OLAP_TRY(sc.get<uint32_t>(padded_parts(C), &cur1));
  OLAP_CUDA(cudaMemsetAsync(cur1, 0, (size_t)padded_parts(C) * 4, s));
  a.shift = kRShift + kRFineBits; a.cap = cap1; a.cursor = cur1; a.out = coarse;
  OLAP_TRY(launch_rpart(ctx, a, kInRaw, grid_limit, s));
  if (mark_coarse) phase_mark(ctx, s, mark_coarse);
  RPartArgs f;
  memset(&f, 0, sizeof(f));
  f.in = coarse; f.reg_cur = cur1; f.reg_cap = cap1; f.n_reg = C; f.reg_mask = 0xFFFFFFFFu; f.reg_bits = kRFineBits;
  f.shift = kRShift; f.cap = cap; f.cursor = cursor; f.out = out; f.flags = flags;
  OLAP_TRY(launch_rpart(ctx, f, kInRegionsCarry, grid_limit, s));
  count_launch(ctx);
  if (mark_fine) phase_mark(ctx, s, mark_fine);
  return OLAP_OK;
}

void dj_build_free(DJBuild& b, cudaStream_t s) {
  if (b.brows) { pool_free(b.ctx, b.brows, s); b.brows = nullptr; }
  if (b.bctl) { pool_free(b.ctx, b.bctl, s); b.bctl = nullptr; }
}

// unique keys: a range of 2^15 keys holds at most 2^15 build rows, so the regions never overflow unless a key repeats
olap_status dj_build_prepare(olap_ctx* ctx, const RInput& in, int64_t rmin, uint64_t rrange, unsigned grid_limit, cudaStream_t s, DJBuild* b) {
  b->ctx = ctx; b->rmin = rmin; b->rrange = rrange;
  b->P = djoin_partitions(rrange); b->PP = padded_parts(b->P);
  OLAP_TRY(pool_alloc(ctx, (uint64_t)b->P * kRKeys * 8 + 32, s, (void**)&b->brows));
  OLAP_TRY(pool_alloc(ctx, ((uint64_t)b->PP + kFlagWords) * 4, s, (void**)&b->bctl));
  OLAP_CUDA(cudaMemsetAsync(b->bctl, 0, ((uint64_t)b->PP + kFlagWords) * 4, s));
  count_launch(ctx);
  Scratch sc(ctx, s);
  OLAP_TRY(partition_rows(ctx, in, rmin, rrange, b->P, kRKeys, b->bctl, b->brows, b->bctl + b->PP, true, 0, nullptr, grid_limit, s, sc,
                          "partition_build_coarse", "partition_build"));
  OLAP_CUDA(cudaGetLastError());
  return OLAP_OK;
}

// Probe one batch: partition it by key range, then build + probe every range in shared memory.  Asynchronous on s.
// d_flags (device, kFlagWords u32, written here) receives the probe-side flags OR-ed with the build side's: a set word means the
// pairs / count of this call must not be used and the caller takes another path (overflow: skewed probe keys, retry with hot = 1
// or bucket tables; out of range: stale remembered key range; dup: repeated build key -> bucket tables).
// np_cap: upper bound of the batch's rows (sizes the regions).
olap_status dj_probe(const DJBuild& b, const RInput& in, uint64_t np_cap, olap_pair* out, uint64_t max_matches, unsigned long long* d_count, int hot,
                     unsigned grid_limit, cudaStream_t s, uint32_t* d_flags) {
  olap_ctx* ctx = b.ctx;
  const uint32_t P = b.P, PP = b.PP;
  const uint32_t slack = hot ? 60 : 25;
  const uint32_t pcap = region_cap(np_cap, P, slack);
  OLAP_CUDA(cudaMemcpyAsync(d_flags, b.bctl + PP, kFlagWords * 4, cudaMemcpyDeviceToDevice, s));
  if ((uint64_t)P * pcap >= 0xFFFFFFFFULL) {  // positions are u32
    OLAP_CUDA(cudaMemsetAsync(d_flags + kFlagOverflow, 1, 4, s));
    return OLAP_OK;
  }
  Scratch sc(ctx, s);
  uint64_t* prows = nullptr;
  uint32_t* ctl = nullptr;  // [PP] probe cursors | ticket (4) | hot keys | hot ids | hot_n (4)
  const uint64_t ctl_words = (uint64_t)PP + 4 + 2 * kHotSlots + 4;
  OLAP_TRY(sc.get<uint64_t>((uint64_t)P * pcap, &prows));
  OLAP_TRY(sc.get<uint32_t>(ctl_words, &ctl));
  OLAP_CUDA(cudaMemsetAsync(ctl, 0, ctl_words * 4, s));
  count_launch(ctx);
  uint32_t *pcur = ctl, *ticket = ctl + PP, *hot_keys = ticket + 4, *hot_ids = hot_keys + kHotSlots, *hot_n = hot_ids + kHotSlots;

  RPartArgs ha;
  memset(&ha, 0, sizeof(ha));
  if (hot && !in.regions) {
    HotArgs h;
    memset(&h, 0, sizeof(h));
    h.n_samples = 1u << 16;
    h.keys = in.keys; h.n = in.n; h.stride = std::max<uint64_t>(h.n / h.n_samples, 1);
    h.rmin = (uint64_t)b.rmin; h.rrange = b.rrange;
    h.tab_slots = 1u << 18;
    uint32_t* tab = nullptr;
    OLAP_TRY(sc.get<uint32_t>(2ULL * h.tab_slots, &tab));
    OLAP_CUDA(cudaMemsetAsync(tab, 0, 2ULL * h.tab_slots * 4, s));
    h.tab_key = tab; h.tab_cnt = tab + h.tab_slots;
    h.hot_keys = hot_keys; h.hot_ids = hot_ids; h.hot_n = hot_n;
    h.min_count = 3;
    hot_sample_kernel<<<(h.n_samples + 255) / 256, 256, 0, s>>>(h);
    hot_select_kernel<<<1, 1024, 0, s>>>(h);
    hot_resolve_kernel<<<kHotSlots, 256, 0, s>>>(hot_keys, hot_ids, b.brows, b.bctl, kRKeys);
    count_launch(ctx, 4);
    phase_mark(ctx, s, "hot_keys");
    ha.hot_keys = hot_keys; ha.hot_ids = hot_ids; ha.pairs = out; ha.max_matches = max_matches; ha.count = d_count;
  } else hot = 0;

  OLAP_TRY(partition_rows(ctx, in, b.rmin, b.rrange, P, pcap, pcur, prows, d_flags, false, slack, hot ? &ha : nullptr, grid_limit, s, sc,
                          "partition_probe_coarse", "partition_probe"));

  DJoinArgs j;
  memset(&j, 0, sizeof(j));
  j.brows = b.brows; j.bcur = b.bctl; j.bcap = kRKeys; j.prows = prows; j.pcur = pcur; j.pcap = pcap; j.P = P;
  j.ticket = ticket; j.flags = d_flags; j.out = out; j.max_matches = max_matches; j.count = d_count;
  const size_t smem = (size_t)kRKeys * 4;
  OLAP_CUDA(cudaFuncSetAttribute(djoin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  unsigned grid = (unsigned)std::min<uint32_t>((uint32_t)ctx->sm_count, P);
  if (grid_limit && grid > grid_limit) grid = grid_limit;
  djoin_kernel<<<grid, kDJThreads, smem, s>>>(j);
  count_launch(ctx);
  phase_mark(ctx, s, "build_probe_smem");
  OLAP_CUDA(cudaGetLastError());
  return OLAP_OK;
}

}  // namespace olap
