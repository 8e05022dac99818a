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
constexpr int kHotSlots = 8192;                   // slots of the heavy-hitter table (32 KB of shared memory)
// The table is 2048 BUCKETS of 4 slots; a key lives in bucket b1(key), or -- only when that one was full -- in bucket b2(key).  A
// lookup is one 16-byte shared-memory load and four compares, plus the same again for the rare row whose first bucket is full: no
// probing chain, no loop, no divergence.  At most a quarter of the slots hold a key (2048 keys: a bucket holds Poisson(1) of them, 2 %
// are full).  What this replaces -- linear probing over 8192 slots, half full -- measured on rpart_kernel<raw, hot, peer> at Zipf 1.0
// (profiles/r02_ncu_zipf_rpart_before.txt): 356 thread instructions per row against 60 without heavy hitters, a third of them in the
// one-lane-at-a-time tail of the probing loop (8.7 iterations per warp and row slot, 3.7 lanes active), 3.8 ms per 2^27-row chunk
// against 0.5 ms.  The keys given up (ranks 2049..4096 of a Zipf(1) distribution over 2^31 rows carry 3 % of the probe rows) travel
// the normal way.
constexpr int kHotBucketBits = 11;
constexpr uint32_t kHotMaxKeys = kHotSlots / 4;
static_assert(kHotSlots == 4 << kHotBucketBits, "2048 buckets of 4 slots");
constexpr uint32_t kHotSampleSlots = 1u << 19;    // counting table of the sampling pass

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
  // [(g % reg_q) * reg_parts, (g % reg_q + 1) * reg_parts).  Carry: the id is kept; Pos: the id becomes the row's position
  // id_base + g * reg_cap + row (multi-GPU: pairs then index the receive buffer, whose low words hold the source row ids)
  const uint32_t* reg_cur; uint32_t reg_cap, n_reg, reg_q, reg_parts;
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
  // PEER output (multi-GPU shuffle, kInRaw only): partition j = offset / peer_subw (sub-ranges need not be a power of two wide, so
  // that the ranks' slices tile the key domain evenly) belongs to rank j / peer_q; its rows go straight into region
  // peer_rank * peer_q + j % peer_q of THAT rank's receive slab (peer-mapped pointer), as {offset - first offset of the rank's
  // slice : u32 hi, source row id : u32 lo}.  Only this rank writes those regions, so the cursors stay in local memory; the
  // counts are published afterwards (shard_publish_kernel).
  uint64_t* peer_out[8]; uint32_t peer_q, peer_subw, peer_rank;
  uint64_t peer_mf; uint32_t peer_mq;   // exact division by multiplication: offset / peer_subw and partition / peer_q (see peer_part / peer_dest)
};

// peer_subw = F << 15 with F <= 256, offsets < 2^32: (offset >> 15) < 2^17 and (x * ceil(2^32 / F)) >> 32 == x / F exactly (the error
// term x / 2^32 < 2^-15 < 1 / F); partitions < 256: (p * ceil(2^24 / Q)) >> 24 == p / Q exactly
static_assert(kRShift == 15, "route_part (djoin.cuh) shifts by 15");
__device__ __forceinline__ uint32_t peer_part(const RPartArgs& a, uint32_t d) { return route_part(d, a.peer_mf); }
__device__ __forceinline__ uint32_t peer_dest(const RPartArgs& a, uint32_t part) { return route_dest(part, a.peer_mq); }

__host__ __device__ __forceinline__ uint32_t hot_b1(uint32_t d) { return (d * 0x9E3779B1u) >> (32 - kHotBucketBits); }
__host__ __device__ __forceinline__ uint32_t hot_b2(uint32_t d) { return ((d ^ 0x5BD1E995u) * 0x85EBCA6Bu) >> (32 - kHotBucketBits); }

// slot of `want` (= offset + 1) in a bucket's four keys, or ~0
__device__ __forceinline__ uint32_t hot_match(const uint4& b, uint32_t want, uint32_t slot0) {
  return b.x == want ? slot0 : b.y == want ? slot0 + 1 : b.z == want ? slot0 + 2 : b.w == want ? slot0 + 3 : 0xFFFFFFFFu;
}

// Heavy-hitter lookup of a thread's R rows.  state[k]: 0 = dead row, 1 = live row to be looked up; on return a live row is 2 (not
// a heavy hitter) or 0x80000000 | slot (heavy hitter).  hotk: the table in shared memory, 16-byte aligned.  Four rows at a time:
// their bucket loads are in flight together, and four buckets are all the registers hold.
template <int R>
__device__ __forceinline__ uint32_t hot_lookup(const uint32_t* hotk, const uint32_t (&d)[R], uint32_t (&state)[R]) {
  static_assert(R % 4 == 0, "rows are looked up four at a time");
  const uint4* hk4 = reinterpret_cast<const uint4*>(hotk);
  uint32_t nhot = 0;
#pragma unroll
  for (int k0 = 0; k0 < R; k0 += 4) {
    uint4 b[4];
#pragma unroll
    for (int j = 0; j < 4; j++) b[j] = hk4[hot_b1(d[k0 + j])];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int k = k0 + j;
      if (state[k] != 1u) continue;
      const uint32_t want = d[k] + 1u;
      uint32_t hit = hot_match(b[j], want, hot_b1(d[k]) << 2);
      if (hit == 0xFFFFFFFFu && b[j].w != 0u) {  // the first bucket is full (slots fill in order): the key may live in its second one
        const uint32_t b2 = hot_b2(d[k]);
        hit = hot_match(hk4[b2], want, b2 << 2);
      }
      if (hit != 0xFFFFFFFFu) { state[k] = 0x80000000u | hit; nhot++; }
      else state[k] = 2u;
    }
  }
  return nhot;
}

// Stage positions of a thread's heavy-hitter rows: the warp scans its per-thread counts and reserves the warp's slots with ONE
// shared-memory atomic per tile (it was a ballot + atomic per row slot: 75 instructions per row in the scatter phase; a per-row
// atomic on the single counter serialised the whole CTA before that -- 38 % of the rows are hot at Zipf 1.0).  The counter ends up
// holding the tile's number of hot rows.  All 32 lanes call it; returns the thread's first slot.
__device__ __forceinline__ uint32_t hot_stage_base(uint32_t* counter, uint32_t nhot) {
  const int lane = threadIdx.x & 31;
  uint32_t incl = nhot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  uint32_t wbase = 0;
  if (lane == 31 && incl) wbase = atomicAdd(counter, incl);
  wbase = __shfl_sync(0xffffffffu, wbase, 31);
  return wbase + incl - nhot;
}

// ROWS: rows per thread and tile.  16 (4096-row tiles) is the shape every variant has; 8 halves the registers held across the tile's
// barriers -- more CTAs per SM to fill the gaps between the phases -- and is instantiated for the plain variants (see launch_rpart).
template <int INPUT, bool HOT, bool PEER = false, int ROWS = kRRows>
__global__ void __launch_bounds__(kRThreads, ROWS == 8 ? (INPUT == kInRaw ? 6 : 5) : ((INPUT == kInRaw && !HOT) ? 4 : 3)) rpart_kernel(const __grid_constant__ RPartArgs a) {
  constexpr int kTile = kRThreads * ROWS;
  static_assert(!PEER || INPUT == kInRaw, "the shuffle reads raw keys");
  __shared__ uint64_t* s_optr[PEER ? kRThreads : 1];                   // PEER: where staged row i of partition j goes = s_optr[j] + i
  constexpr int R = ROWS, NW = kRThreads / 32;
  __shared__ __align__(16) uint64_t stage[kTile];                    // rows in partition order (32 KB)
  __shared__ uint32_t cnt[kRThreads];                                 // tile histogram, then exclusive offsets
  __shared__ uint32_t gbase[kRThreads];                               // global position of staged row i = gbase[part] + i
  extern __shared__ __align__(16) uint32_t hotk[];                    // HOT: the hot keys (32 KB, dynamic); their build rows are read through L1
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

  const uint32_t tiles_per_reg = INPUT == kInRaw ? 1u : (a.reg_cap + kTile - 1) / kTile;
  const uint32_t tiles = INPUT == kInRaw ? (uint32_t)((a.n + kTile - 1) / kTile) : tiles_per_reg * a.n_reg;
  int par = 1;
  for (uint32_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    const uint64_t* src = a.in;
    uint64_t n = a.n, row0 = (uint64_t)t * kTile;
    uint32_t pbase = 0, idb = a.id_base;
    if (INPUT != kInRaw) {
      const uint32_t g = t / tiles_per_reg;
      src = a.in + (uint64_t)g * a.reg_cap;
      n = min(a.reg_cur[g], a.reg_cap);
      row0 = (uint64_t)(t - g * tiles_per_reg) * kTile;
      pbase = (g % a.reg_q) * a.reg_parts;
      idb = a.id_base + g * a.reg_cap;
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
    uint32_t nhot = 0;
    // rk: kREmpty = dead row, 0x80000000 | slot = heavy hitter, else the rank inside the tile
#pragma unroll
    for (int k = 0; k < R; k++) {
      const uint64_t row = row0 + (uint64_t)k * kRThreads + tid;
      bool ok = row < n;
      if (INPUT == kInRegionsPos) ok = ok && key[k] != ~0ULL;  // padding word of the shuffle's 16-byte aligned runs
      if (INPUT != kInRaw) d[k] = (uint32_t)(key[k] >> 32);
      else {
        const uint64_t off = key[k] - a.rmin;
        if (off >= a.rrange) { oob = oob || ok; ok = false; }
        d[k] = (uint32_t)off;
      }
      if (HOT) rk[k] = ok ? 1u : 0u;
      else rk[k] = ok ? atomicAdd(&cnt[(PEER ? peer_part(a, d[k]) : d[k] >> a.shift) - pbase], 1u) : kREmpty;   // one loop: fewest live registers
    }
    if (HOT) {
      nhot = hot_lookup<R>(hotk, d, rk);
#pragma unroll
      for (int k = 0; k < R; k++) {
        if (rk[k] == 1u || rk[k] == 2u) rk[k] = atomicAdd(&cnt[(PEER ? peer_part(a, d[k]) : d[k] >> a.shift) - pbase], 1u);   // live, not a heavy hitter
        else if (rk[k] == 0u) rk[k] = kREmpty;
      }
    }
    if (INPUT == kInRaw && oob && a.oob_is_error) a.flags[kFlagOutOfRange] = 1;
    uint32_t hbase = 0;   // HOT: this thread's first slot among the tile's heavy-hitter rows
    if (HOT) hbase = hot_stage_base(&s_hot[par], nhot);
    __syncthreads();

    // phase 2: exclusive scan of the tile histogram (one partition per thread); every non-empty (tile, partition) run is
    // reserved in the partition's region with one global atomic, whose round trip overlaps the scan
    const uint32_t c = cnt[tid];
    const uint32_t got = c ? atomicAdd(&a.cursor[pbase + tid], c) : 0u;
    if (tid == 0) {
      s_over[par ^ 1] = 0; s_hot[par ^ 1] = 0;  // the previous tile is done with them, the next one starts after a barrier
      if (HOT) {  // the tile's heavy-hitter pairs are reserved in the output now; the round trip is over by the time they are written
        const uint32_t nh = s_hot[par];
        if (nh) s_hotbase = atomicAdd(a.count, (unsigned long long)nh);
      }
    }
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
      if (PEER) {
        const uint32_t dest = min((uint32_t)tid / a.peer_q, 7u);
        gbase[tid] = got - off;  // position inside the region
        s_optr[tid] = a.peer_out[dest] + (uint64_t)(a.peer_rank * a.peer_q + (uint32_t)tid % a.peer_q) * a.cap + (int64_t)(int32_t)(got - off);
      } else gbase[tid] = (pbase + (uint32_t)tid) * a.cap + got - off;
      if (got + c > a.cap) { s_over[par] = 1; a.flags[kFlagOverflow] = 1; }
      if (tid == kRThreads - 1) s_total = off + c;
    }
    __syncthreads();

    // phase 3: scatter the tile into shared memory in partition order; hot probe rows go to the top of the stage, downwards
    const uint32_t total = s_total;
#pragma unroll
    for (int k = 0; k < R; k++) {
      const bool is_hot = HOT && rk[k] != kREmpty && (rk[k] & 0x80000000u);
      if (rk[k] == kREmpty) continue;
      const uint32_t id = INPUT == kInRegionsCarry ? (uint32_t)key[k] : idb + (uint32_t)(row0 + (uint64_t)k * kRThreads + tid);
      if (is_hot) {
        stage[kTile - 1 - hbase++] = ((uint64_t)id << 32) | __ldg(a.hot_ids + (rk[k] & (kHotSlots - 1)));
      } else {
        stage[cnt[(PEER ? peer_part(a, d[k]) : d[k] >> a.shift) - pbase] + rk[k]] = ((uint64_t)d[k] << 32) | id;
      }
    }
    __syncthreads();
    cnt[tid] = 0;  // the offsets are consumed: clear the histogram for the next tile

    // phase 4: every partition's run leaves with consecutive lanes on consecutive addresses
    if (PEER) {
      const bool over = s_over[par] != 0;
      for (uint32_t i = tid; i < total; i += kRThreads) {
        const uint64_t v = stage[i];
        const uint32_t d0 = (uint32_t)(v >> 32), part = peer_part(a, d0);
        if (over && gbase[part] + i >= a.cap) continue;  // the region is full: dropped, flagged
        const uint32_t dl = d0 - peer_dest(a, part) * (a.peer_q * a.peer_subw);
        s_optr[part][i] = ((uint64_t)dl << 32) | (uint32_t)v;
      }
    } else if (!s_over[par]) {
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
      const uint32_t nh = s_hot[par];
      if (nh) {
        const unsigned long long base = s_hotbase;
        for (uint32_t i = tid; i < nh; i += kRThreads) {
          const uint64_t v = stage[kTile - 1 - i];  // {probe row : hi, build row : lo}
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

// 1024 threads = 4 probe groups when the join owns the SM; 768 threads = 3 groups (36 K registers) when it shares the SM with a
// shuffle CTA of the next chunk (multi-GPU: 24 K registers, 70 KB of shared memory)
constexpr int kDJGroupThreads = 256;
constexpr int kDJRows = 8;                         // probe rows per thread per group tile
constexpr int kDJGroupTile = kDJGroupThreads * kDJRows;

__device__ __forceinline__ void group_bar(int g) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(kDJGroupThreads) : "memory"); }

template <int THREADS>
__device__ __forceinline__ void djoin_body(const DJoinArgs& a) {
  constexpr int GROUPS = THREADS / kDJGroupThreads;   // independent probe groups inside the CTA (named barriers)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* table = reinterpret_cast<uint32_t*>(smem_raw);   // [kRKeys] build row id per key of this range
  __shared__ uint32_t s_part;
  __shared__ uint32_t s_n[GROUPS];
  __shared__ unsigned long long s_base[GROUPS];
  const int tid = threadIdx.x, lane = tid & 31;
  const int g = tid / kDJGroupThreads, gt = tid % kDJGroupThreads;
  if (tid < GROUPS) s_n[tid] = 0;
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
      for (uint32_t i = tid; i < kRKeys / 4; i += THREADS) t4[i] = e;
    }
    __syncthreads();
    {
      const uint64_t* rows = a.brows + (uint64_t)p * a.bcap;
      bool dup = false;
      for (uint32_t i0 = 0; i0 < nb; i0 += THREADS * 8) {
        uint64_t v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) { const uint32_t i = i0 + k * THREADS + tid; v[k] = i < nb ? ld_stream_u64(rows + i) : ~0ULL; }
#pragma unroll
        for (int k = 0; k < 8; k++)
          if (i0 + k * THREADS + tid < nb) dup = dup || atomicExch(&table[(uint32_t)(v[k] >> 32) & (kRKeys - 1)], (uint32_t)v[k]) != kREmpty;
      }
      if (dup) a.flags[kFlagDup] = 1;  // a repeated build key: this representation cannot hold it (the caller redoes the join with bucket tables)
    }
    __syncthreads();
    const uint64_t* rows = a.prows + (uint64_t)p * a.pcap;
    for (uint32_t t0 = (uint32_t)g * kDJGroupTile; t0 < np; t0 += GROUPS * kDJGroupTile) {
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

__global__ void __launch_bounds__(1024, 1) djoin_kernel(const __grid_constant__ DJoinArgs a) { djoin_body<1024>(a); }
// 48 registers x 768 threads = 36 K: fits an SM together with one shuffle CTA (96 x 256 = 24 K)
__global__ void __maxnreg__(48) djoin_shared_kernel(const __grid_constant__ DJoinArgs a) { djoin_body<768>(a); }

// ------------------------------------------------------------------------------------------- heavy hitters
// Sample the probe keys, count the samples per key in a small global table, keep the keys seen at least `min_count` times (at
// most kHotMaxKeys of them).  All on the device: no host round trip.
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

// one CTA: raise the threshold until at most kHotMaxKeys keys qualify, then build the open-addressing hot table
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
    if (threadIdx.x == 0) s_thr = s_cnt <= kHotMaxKeys ? 0 : thr + (thr + 1) / 2;
    __syncthreads();
    if (s_thr == 0) break;
    thr = s_thr;
    __syncthreads();
  }
  for (uint32_t i = threadIdx.x; i < h.tab_slots; i += blockDim.x) {
    const uint32_t k = h.tab_key[i];
    if (k && h.tab_cnt[i] >= thr) {
      // first free slot of bucket b1, else of bucket b2 (slots are taken in order and never freed, so a bucket whose last slot is
      // taken is full); a key that finds both full is simply not a heavy hitter
      bool placed = false;
      for (int t = 0; t < 2 && !placed; t++) {
        const uint32_t b = t ? hot_b2(k - 1) : hot_b1(k - 1);
        for (int j = 0; j < 4 && !placed; j++) placed = atomicCAS(&h.hot_keys[4 * b + j], 0u, k) == 0u;
      }
      if (placed) atomicAdd(h.hot_n, 1u);
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

// Kernels of different launches share an SM only if they agree on its shared-memory / L1 split.  In the multi-GPU path (shared =
// true) every kernel of this file asks for the maximum shared-memory carveout, so that a shuffle CTA, partition CTAs and a join
// CTA can be resident together (chunk k + 1's shuffle next to chunk k's join).  On its own a partition pass keeps the default
// split: with the L1 cut down to the minimum it ran 20-27 % slower (1.65 -> 2.09 ms for the 2^29-row coarse pass).
template <class K>
static cudaError_t set_carveout(K kernel, bool shared) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, shared ? (int)cudaSharedmemCarveoutMaxShared : (int)cudaSharedmemCarveoutDefault);
}

template <int INPUT, bool HOT, bool PEER = false, int ROWS = kRRows>
static olap_status launch_rpart_t(olap_ctx* ctx, const RPartArgs& a, unsigned grid, bool shared, cudaStream_t s) {
  OLAP_CUDA((set_carveout(rpart_kernel<INPUT, HOT, PEER, ROWS>, shared)));
  const size_t dyn = HOT ? (size_t)kHotSlots * 4 : 0;
  if (HOT) OLAP_CUDA((cudaFuncSetAttribute(rpart_kernel<INPUT, HOT, PEER, ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn)));
  rpart_kernel<INPUT, HOT, PEER, ROWS><<<grid, kRThreads, dyn, s>>>(a);
  count_launch(ctx);
  return OLAP_OK;
}

// grid_limit: CTAs to use at most (0 = fill the GPU); the multi-GPU path runs this next to a shuffle kernel
static olap_status launch_rpart(olap_ctx* ctx, const RPartArgs& a, int input, unsigned grid_limit, cudaStream_t s) {
  const bool hot = a.hot_keys != nullptr, peer = a.peer_out[0] != nullptr, shared = grid_limit != 0;
  // tile shape: 8 rows per thread for the plain variants when OLAP_RPART_ROWS=8.  A/B knob kept for the record: unlike the group-by's
  // two-column partition pass (groupby.cu, where 5 CTAs per SM gain a fifth), these one-column passes do not profit -- C2 on a B200:
  // coarse probe pass 1.65 ms at 16 rows (4 CTAs per SM, 64 registers) vs 1.80 ms at 8 rows (6 CTAs, 40 registers), second level
  // 1.78 vs 1.79 ms: twice the run reservations and 8-row runs cost what the extra occupancy buys.  16 is the default.
  static const int rows_env = [] { const char* v = getenv("OLAP_RPART_ROWS"); return v && *v ? atoi(v) : 16; }();
  const bool r8 = rows_env == 8 && !hot && !peer && !shared;
  const uint64_t tile = (uint64_t)kRThreads * (r8 ? 8 : kRRows);
  const uint64_t tiles = input == kInRaw ? (a.n + tile - 1) / tile : (uint64_t)((a.reg_cap + tile - 1) / tile) * a.n_reg;
  if (!tiles) return OLAP_OK;
  // one wave of resident CTAs (the tile loop strides by the grid): 4 per SM for raw keys, 3 for the region passes and for the
  // heavy-hitter variant (__launch_bounds__ above) -- launched with 4 per SM, a quarter of its CTAs ran as a second wave at a third
  // of the occupancy
  uint64_t grid = (uint64_t)ctx->sm_count * (r8 ? (input == kInRaw ? 6 : 5) : ((input == kInRaw && !hot) ? 4 : 3));
  if (grid_limit && grid > grid_limit) grid = grid_limit;
  if (grid > tiles) grid = tiles;
  if (r8) {
    if (input == kInRaw) return launch_rpart_t<kInRaw, false, false, 8>(ctx, a, (unsigned)grid, shared, s);
    if (input == kInRegionsCarry) return launch_rpart_t<kInRegionsCarry, false, false, 8>(ctx, a, (unsigned)grid, shared, s);
    return launch_rpart_t<kInRegionsPos, false, false, 8>(ctx, a, (unsigned)grid, shared, s);
  }
  if (input == kInRaw && peer) return hot ? launch_rpart_t<kInRaw, true, true>(ctx, a, (unsigned)grid, shared, s) : launch_rpart_t<kInRaw, false, true>(ctx, a, (unsigned)grid, shared, s);
  if (input == kInRaw) return hot ? launch_rpart_t<kInRaw, true>(ctx, a, (unsigned)grid, shared, s) : launch_rpart_t<kInRaw, false>(ctx, a, (unsigned)grid, shared, s);
  if (input == kInRegionsCarry) return launch_rpart_t<kInRegionsCarry, false>(ctx, a, (unsigned)grid, shared, s);
  return launch_rpart_t<kInRegionsPos, false>(ctx, a, (unsigned)grid, shared, s);
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
    if (in.reg_parts > (uint32_t)kRThreads || !in.reg_q) { set_error("shuffle regions span more than 256 key ranges"); return OLAP_ERR_BAD_ARG; }
    a.in = in.regions; a.reg_cur = in.reg_cur; a.reg_cap = in.reg_cap; a.n_reg = in.n_reg; a.reg_q = in.reg_q; a.reg_parts = in.reg_parts;
    a.id_base = in.id_base;
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
  f.in = coarse; f.reg_cur = cur1; f.reg_cap = cap1; f.n_reg = C; f.reg_q = C; f.reg_parts = 1u << kRFineBits;
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

// ---- probing in three steps, so that rows can arrive batch by batch and still be joined with ONE build of every table:
//   dj_probe_begin  regions for up to np_cap rows (pool memory, cursors zeroed)
//   dj_probe_add    partition one batch of rows into the regions (appends; any number of times)
//   dj_probe_join   build + probe every key range in shared memory over what has been added, then reset the regions
// The multi-GPU path adds a chunk as soon as its shuffle has landed and joins once per group of chunks: a join CTA rebuilds its
// 128 KB table per call, which costs as much as probing 2^15 rows -- chunk-wise joins at 2^14 probe rows per range spent more time
// building than probing (14.3 ms of local work per 2^28 x 2^30 join in 8 chunks against 8.3 ms in one piece).
olap_status dj_probe_begin(const DJBuild& b, uint64_t np_cap, int hot, cudaStream_t s, DJProbe* p) {
  olap_ctx* ctx = b.ctx;
  p->b = &b; p->hot = hot == 1; p->np_cap = np_cap;   // hot = 2: room for skew, but the heavy hitters were handled upstream (multi-GPU shuffle)
  p->slack = hot == 2 ? 200 : hot ? 60 : 25;
  p->pcap = region_cap(np_cap, b.P, p->slack);
  p->prows = nullptr; p->ctl = nullptr;
  if ((uint64_t)b.P * p->pcap >= 0xFFFFFFFFULL) { p->pcap = 0; return OLAP_OK; }  // positions are u32: every add raises the overflow flag
  const uint64_t ctl_words = (uint64_t)b.PP + 4 + 2 * kHotSlots + 4;
  OLAP_TRY(pool_alloc(ctx, (uint64_t)b.P * p->pcap * 8 + 32, s, (void**)&p->prows));
  OLAP_TRY(pool_alloc(ctx, ctl_words * 4, s, (void**)&p->ctl));
  OLAP_CUDA(cudaMemsetAsync(p->ctl, 0, ctl_words * 4, s));
  count_launch(ctx);
  return OLAP_OK;
}

void dj_probe_end(DJProbe& p, cudaStream_t s) {
  if (!p.b) return;
  if (p.prows) { pool_free(p.b->ctx, p.prows, s); p.prows = nullptr; }
  if (p.ctl) { pool_free(p.b->ctx, p.ctl, s); p.ctl = nullptr; }
  p.b = nullptr;
}

// d_flags (device, kFlagWords u32): words are only ever SET here (overflow: skewed probe keys, retry with hot = 1 or bucket
// tables).  out / max_matches / d_count: where heavy-hitter pairs go when p.hot (they are joined in this very pass).
olap_status dj_probe_add(DJProbe& p, const RInput& in, olap_pair* out, uint64_t max_matches, unsigned long long* d_count, unsigned grid_limit,
                         cudaStream_t s, uint32_t* d_flags) {
  const DJBuild& b = *p.b;
  olap_ctx* ctx = b.ctx;
  if (!p.pcap) { OLAP_CUDA(cudaMemsetAsync(d_flags + kFlagOverflow, 1, 4, s)); return OLAP_OK; }
  uint32_t *pcur = p.ctl, *ticket = p.ctl + b.PP, *hot_keys = ticket + 4, *hot_ids = hot_keys + kHotSlots, *hot_n = hot_ids + kHotSlots;
  Scratch sc(ctx, s);
  RPartArgs ha;
  memset(&ha, 0, sizeof(ha));
  bool hot = false;
  if (p.hot && !in.regions) {
    HotArgs h;
    memset(&h, 0, sizeof(h));
    h.n_samples = 1u << 18;
    h.keys = in.keys; h.n = in.n; h.stride = std::max<uint64_t>(h.n / h.n_samples, 1);
    h.rmin = (uint64_t)b.rmin; h.rrange = b.rrange;
    h.tab_slots = kHotSampleSlots;
    uint32_t* tab = nullptr;
    OLAP_TRY(sc.get<uint32_t>(2ULL * h.tab_slots, &tab));
    OLAP_CUDA(cudaMemsetAsync(tab, 0, 2ULL * h.tab_slots * 4, s));
    OLAP_CUDA(cudaMemsetAsync(hot_keys, 0, (2ULL * kHotSlots + 4) * 4, s));
    h.tab_key = tab; h.tab_cnt = tab + h.tab_slots;
    h.hot_keys = hot_keys; h.hot_ids = hot_ids; h.hot_n = hot_n;
    h.min_count = 3;
    hot_sample_kernel<<<(h.n_samples + 255) / 256, 256, 0, s>>>(h);
    hot_select_kernel<<<1, 1024, 0, s>>>(h);
    hot_resolve_kernel<<<kHotSlots, 256, 0, s>>>(hot_keys, hot_ids, b.brows, b.bctl, kRKeys);
    count_launch(ctx, 4);
    phase_mark(ctx, s, "hot_keys");
    ha.hot_keys = hot_keys; ha.hot_ids = hot_ids; ha.pairs = out; ha.max_matches = max_matches; ha.count = d_count;
    hot = true;
  }
  return partition_rows(ctx, in, b.rmin, b.rrange, b.P, p.pcap, pcur, p.prows, d_flags, false, p.slack, hot ? &ha : nullptr, grid_limit, s, sc,
                        "partition_probe_coarse", "partition_probe");
}

// d_flags: dup (a repeated build key -> bucket tables) is set here.  Afterwards the regions are empty again.
olap_status dj_probe_join(DJProbe& p, olap_pair* out, uint64_t max_matches, unsigned long long* d_count, unsigned grid_limit, cudaStream_t s,
                          uint32_t* d_flags) {
  const DJBuild& b = *p.b;
  olap_ctx* ctx = b.ctx;
  if (!p.pcap) return OLAP_OK;
  DJoinArgs j;
  memset(&j, 0, sizeof(j));
  j.brows = b.brows; j.bcur = b.bctl; j.bcap = kRKeys; j.prows = p.prows; j.pcur = p.ctl; j.pcap = p.pcap; j.P = b.P;
  j.ticket = p.ctl + b.PP; j.flags = d_flags; j.out = out; j.max_matches = max_matches; j.count = d_count;
  const size_t smem = (size_t)kRKeys * 4;
  unsigned grid = (unsigned)std::min<uint32_t>((uint32_t)ctx->sm_count, b.P);
  if (grid_limit) {  // shared SMs (multi-GPU: a shuffle CTA of the next chunk lives next to every join CTA)
    OLAP_CUDA(cudaFuncSetAttribute(djoin_shared_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    OLAP_CUDA(set_carveout(djoin_shared_kernel, true));
    djoin_shared_kernel<<<grid, 768, smem, s>>>(j);
  } else {
    OLAP_CUDA(cudaFuncSetAttribute(djoin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    OLAP_CUDA(set_carveout(djoin_kernel, false));
    djoin_kernel<<<grid, 1024, smem, s>>>(j);
  }
  count_launch(ctx);
  phase_mark(ctx, s, "build_probe_smem");
  OLAP_CUDA(cudaGetLastError());
  OLAP_CUDA(cudaMemsetAsync(p.ctl, 0, ((uint64_t)b.PP + 4) * 4, s));  // cursors + ticket: ready for the next group
  return OLAP_OK;
}

// Probe one batch in one go.  Asynchronous on s.  d_flags (device, kFlagWords u32, written here) receives the probe-side flags
// OR-ed with the build side's: a set word means the pairs / count of this call must not be used and the caller takes another path
// (overflow: skewed probe keys, retry with hot = 1 or bucket tables; out of range: stale remembered key range; dup: repeated build
// key -> bucket tables).  np_cap: upper bound of the batch's rows (sizes the regions).
olap_status dj_probe(const DJBuild& b, const RInput& in, uint64_t np_cap, olap_pair* out, uint64_t max_matches, unsigned long long* d_count, int hot,
                     unsigned grid_limit, cudaStream_t s, uint32_t* d_flags) {
  OLAP_CUDA(cudaMemcpyAsync(d_flags, b.bctl + b.PP, kFlagWords * 4, cudaMemcpyDeviceToDevice, s));
  DJProbe p;
  olap_status st = dj_probe_begin(b, np_cap, hot && !in.regions, s, &p);
  if (st == OLAP_OK) st = dj_probe_add(p, in, out, max_matches, d_count, grid_limit, s, d_flags);
  if (st == OLAP_OK) st = dj_probe_join(p, out, max_matches, d_count, grid_limit, s, d_flags);
  dj_probe_end(p, s);
  return st;
}

// ------------------------------------------------------------------------------------------- the shuffle kernel (TMA write-out)
// The multi-GPU shuffle as its own kernel, built to saturate NVLink from ONE CTA per SM so that the local partition + join
// kernels of the previous chunk keep the rest of the SM (registers, shared memory):
//   * every (tile, region) run is padded to an even number of rows (the pad is the all-ones word, which the receiver's partition
//     pass skips) so that runs start 16-byte aligned both in the stage and in the region;
//   * the write-out is one cp.async.bulk (shared -> global, the TMA engine; UBLKCP in SASS) per non-empty run, issued by the thread
//     that owns the partition -- no LSU instructions, no registers, full-size NVLink packets;
//   * the stage is double-buffered: the bulk copies of tile t drain over NVLink while tile t + 1 is loaded, ranked and staged;
//     a thread waits for its own copy of tile t - 1 (cp.async.bulk.wait_group.read 1) before that buffer is overwritten.
// Measured before this kernel (rpart_kernel<raw, peer>, 8-byte stores from 2 CTAs per SM): 436 GB/s per direction at 2 GPUs, and
// the shuffle's register footprint kept the join kernel off the SMs (20.4 ms per step = shuffle 12.3 + local 8, no overlap).
constexpr int kShufStage = kRTile + kRThreads;   // rows per stage buffer: a tile plus one pad word per partition

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 96 registers x 256 threads = 24 K: one shuffle CTA + two partition CTAs (2 x 20 K) or one join CTA (36 K) fill an SM's 64 K
template <bool HOT>
__global__ void __maxnreg__(96) shuffle_kernel(const __grid_constant__ RPartArgs a) {
  constexpr int R = kRRows, NW = kRThreads / 32;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* stage0 = reinterpret_cast<uint64_t*>(smem_raw);           // [2][kShufStage]
  __shared__ uint32_t cnt[kRThreads];
  __shared__ __align__(16) uint32_t hotk[HOT ? kHotSlots : 4];
  __shared__ uint32_t s_warp[NW];
  __shared__ uint32_t s_hot[2];
  __shared__ unsigned long long s_hotbase;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  cnt[tid] = 0;
  if (HOT) {
    for (int i = tid; i < kHotSlots; i += kRThreads) hotk[i] = a.hot_keys[i];
  }
  if (tid == 0) { s_hot[0] = s_hot[1] = 0; }
  __syncthreads();
  // this thread's partition: destination rank, region there
  const uint32_t my_dest = min(peer_dest(a, (uint32_t)tid), 7u);
  uint64_t* const my_region = a.peer_out[my_dest] + (uint64_t)(a.peer_rank * a.peer_q + (uint32_t)tid % a.peer_q) * a.cap;

  const uint32_t tiles = (uint32_t)((a.n + kRTile - 1) / kRTile);
  int par = 1;
  uint64_t key[R];
  {
    const uint64_t row0 = (uint64_t)blockIdx.x * kRTile;
#pragma unroll
    for (int k = 0; k < R; k++) { const uint64_t row = row0 + (uint64_t)k * kRThreads + tid; key[k] = row < a.n ? ld_stream_u64(a.in + row) : 0; }
  }
  for (uint32_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    par ^= 1;
    uint64_t* stage = stage0 + (size_t)par * kShufStage;
    const uint64_t row0 = (uint64_t)t * kRTile;
    // phase 1: offset, partition (= sub-range of the global domain), rank inside the tile.  All of a thread's 16 shared-memory
    // atomics are issued before any result is used: with one CTA of 8 warps per SM nothing else hides their latency.
    uint32_t d[R], pr[R];   // pr = partition << 16 | rank; kREmpty: dead row; bit 31: heavy hitter (slot in the low bits)
    const bool full = row0 + kRTile <= a.n;   // uniform: interior tiles skip the per-row bounds checks
    bool oob = false;
    uint32_t nhot = 0;
#pragma unroll
    for (int k = 0; k < R; k++) {
      bool ok = full || row0 + (uint64_t)k * kRThreads + tid < a.n;
      const uint64_t off = key[k] - a.rmin;
      if (off >= a.rrange) { oob = oob || ok; ok = false; }
      d[k] = (uint32_t)off;
      pr[k] = ok ? 1u : 0u;
    }
    if (HOT) nhot = hot_lookup<R>(hotk, d, pr);
#pragma unroll
    for (int k = 0; k < R; k++) pr[k] = (pr[k] == 1u || pr[k] == 2u) ? 0u : (pr[k] == 0u ? kREmpty : pr[k]);   // 0 = live, to be ranked below
    uint32_t hbase = 0;
    if (HOT) hbase = hot_stage_base(&s_hot[par], nhot);
    uint32_t rank[R];
#pragma unroll
    for (int k = 0; k < R; k++) rank[k] = pr[k] == 0u ? atomicAdd(&cnt[peer_part(a, d[k])], 1u) : 0u;
    if (oob && a.oob_is_error) a.flags[kFlagOutOfRange] = 1;
    // the next tile's keys: in flight during the scan, the staging and the write-out of this one
    {
      const uint64_t nrow0 = row0 + (uint64_t)gridDim.x * kRTile;
      if (nrow0 + kRTile <= a.n) {
#pragma unroll
        for (int k = 0; k < R; k++) key[k] = ld_stream_u64(a.in + nrow0 + (uint64_t)k * kRThreads + tid);
      } else {
#pragma unroll
        for (int k = 0; k < R; k++) { const uint64_t row = nrow0 + (uint64_t)k * kRThreads + tid; key[k] = row < a.n ? ld_stream_u64(a.in + row) : 0; }
      }
    }
#pragma unroll
    for (int k = 0; k < R; k++) if (pr[k] == 0u) pr[k] = (peer_part(a, d[k]) << 16) | rank[k];
    __syncthreads();
    // phase 2: runs padded to an even length; one global atomic (local memory: this rank is the only writer of its regions)
    const uint32_t c = cnt[tid], cp = (c + 1) & ~1u;
    const uint32_t got = c ? atomicAdd(&a.cursor[tid], cp) : 0u;
    if (tid == 0) {
      s_hot[par ^ 1] = 0;
      if (HOT) {
        const uint32_t nh = s_hot[par];
        if (nh) s_hotbase = atomicAdd(a.count, (unsigned long long)nh);
      }
    }
    uint32_t inc = cp;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
    if (lane == 31) s_warp[warp] = inc;
    // the bulk copy this thread issued two tiles ago read from the buffer that is about to be overwritten
    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncthreads();
    uint32_t wbase = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) wbase += w < warp ? s_warp[w] : 0u;
    const uint32_t off = wbase + inc - cp;
    cnt[tid] = off;
    if (c & 1) stage[off + c] = ~0ULL;  // the pad word
    __syncthreads();
    // phase 3: scatter into the stage in partition order, already in the receiver's form {offset in its slice, source row id}
#pragma unroll
    for (int k = 0; k < R; k++) {
      const bool is_hot = HOT && pr[k] != kREmpty && (pr[k] & 0x80000000u);
      if (pr[k] == kREmpty) continue;
      const uint32_t id = a.id_base + (uint32_t)(row0 + (uint64_t)k * kRThreads + tid);
      if (is_hot) {
        stage[kShufStage - 1 - hbase++] = ((uint64_t)id << 32) | __ldg(a.hot_ids + (pr[k] & (kHotSlots - 1)));
      } else {
        const uint32_t part = pr[k] >> 16;
        const uint32_t dl = d[k] - peer_dest(a, part) * (a.peer_q * a.peer_subw);
        stage[cnt[part] + (pr[k] & 0xFFFFu)] = ((uint64_t)dl << 32) | id;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the staged rows become visible to the TMA engine
    __syncthreads();
    // phase 4: one bulk copy per run, shared -> (peer) global.  The reservation's round trip (an atomic in local memory) is first
    // needed here, two barriers after it was issued.
    const bool over = got + cp > a.cap;
    if (over) a.flags[kFlagOverflow] = 1;
    if (c && !over) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(my_region + got), "r"(smem_u32(stage + off)), "r"(cp * 8u)
                   : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    cnt[tid] = 0;
    if (HOT) {
      const uint32_t nh = s_hot[par];
      if (nh) {
        const unsigned long long base = s_hotbase;
        for (uint32_t i = tid; i < nh; i += kRThreads) {
          const uint64_t v = stage[kShufStage - 1 - i];  // {probe row : hi, replicated build row : lo}
          if (a.pairs && base + i < a.max_matches) a.pairs[base + i] = olap_pair{(uint32_t)v, (uint32_t)(v >> 32)};
        }
      }
    }
    __syncthreads();
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all copies have landed before the CTA's shared memory goes away
}

// ------------------------------------------------------------------------------------------- multi-GPU shuffle by key range
// Rows of `keys` -> the peers' receive slabs (see RPartArgs PEER).  cursor: [256] u32 in local memory, zeroed by the caller on the
// same stream; flags: local kFlagWords.  hot: optional heavy-hitter table -- probe rows carrying a hot key stay on this rank and
// are joined here against the replicated build rows of those keys.
olap_status range_shuffle(olap_ctx* ctx, const uint64_t* keys, uint64_t n, uint32_t id_base, int64_t gmin, uint64_t grange, uint32_t q, uint32_t subw,
                          int rank, int world, uint64_t* const* peer_regions, uint32_t cap, uint32_t* cursor, uint32_t* flags, bool build_side,
                          const HotRef* hot, unsigned grid_limit, bool tma, cudaStream_t s) {
  if (world > 8 || (uint64_t)world * q > (uint64_t)kRThreads) { set_error("the range shuffle serves <= 8 ranks and <= 256 regions per pass"); return OLAP_ERR_BAD_ARG; }
  RPartArgs a;
  memset(&a, 0, sizeof(a));
  a.in = keys; a.n = n; a.id_base = id_base; a.rmin = (uint64_t)gmin; a.rrange = grange;
  a.cap = cap; a.cursor = cursor; a.flags = flags; a.oob_is_error = build_side ? 1 : 0;
  for (int i = 0; i < world; i++) a.peer_out[i] = peer_regions[i];
  a.peer_q = q; a.peer_subw = subw; a.peer_rank = (uint32_t)rank;
  {
    const uint64_t F = subw >> kRShift;
    if (F == 0 || F > 256 || (F << kRShift) != subw || q == 0 || q > 256) { set_error("shuffle sub-ranges must be F x 2^15 keys wide, F <= 256"); return OLAP_ERR_BAD_ARG; }
    a.peer_mf = route_mul_f(F);
    a.peer_mq = route_mul_q(q);
  }
  if (hot && hot->hot_keys) { a.hot_keys = hot->hot_keys; a.hot_ids = hot->hot_ids; a.pairs = hot->pairs; a.max_matches = hot->max_matches; a.count = hot->count; }
  if (!tma) {  // plain stores from the partition kernel at full occupancy: the rows go to local memory (outbox / own slab)
    OLAP_TRY(launch_rpart(ctx, a, kInRaw, grid_limit, s));
  } else {
    const uint64_t tiles = (n + kRTile - 1) / kRTile;
    uint64_t grid = grid_limit ? grid_limit : (uint64_t)ctx->sm_count;
    if (grid > tiles) grid = tiles;
    const size_t smem = (size_t)2 * kShufStage * 8;
    if (a.hot_keys) {
      OLAP_CUDA((cudaFuncSetAttribute(shuffle_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
      OLAP_CUDA(set_carveout(shuffle_kernel<true>, true));
      shuffle_kernel<true><<<(unsigned)grid, kRThreads, smem, s>>>(a);
    } else {
      OLAP_CUDA((cudaFuncSetAttribute(shuffle_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
      OLAP_CUDA(set_carveout(shuffle_kernel<false>, true));
      shuffle_kernel<false><<<(unsigned)grid, kRThreads, smem, s>>>(a);
    }
    count_launch(ctx);
  }
  OLAP_CUDA(cudaGetLastError());
  return OLAP_OK;
}

// after the shuffle kernel: this rank's row count of every region it filled, written into the owners' count arrays
__global__ void shard_publish_kernel(const uint32_t* __restrict__ cursor, uint32_t cap, uint32_t q, uint32_t rank, uint32_t world, uint32_t* const* peer_counts8,
                                     uint32_t* p0, uint32_t* p1, uint32_t* p2, uint32_t* p3, uint32_t* p4, uint32_t* p5, uint32_t* p6, uint32_t* p7) {
  uint32_t* pc[8] = {p0, p1, p2, p3, p4, p5, p6, p7};
  const uint32_t j = threadIdx.x;
  if (j < world * q) pc[j / q][rank * q + j % q] = min(cursor[j], cap);
}

olap_status range_shuffle_publish(olap_ctx* ctx, const uint32_t* cursor, uint32_t cap, uint32_t q, int rank, int world, uint32_t* const* peer_counts,
                                  cudaStream_t s) {
  uint32_t* pc[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  for (int i = 0; i < world; i++) pc[i] = peer_counts[i];
  shard_publish_kernel<<<1, 256, 0, s>>>(cursor, cap, q, (uint32_t)rank, (uint32_t)world, nullptr, pc[0], pc[1], pc[2], pc[3], pc[4], pc[5], pc[6], pc[7]);
  count_launch(ctx);
  OLAP_CUDA(cudaGetLastError());
  return OLAP_OK;
}

// ------------------------------------------------------------------------------------------- heavy hitters across ranks
// Host restatement of hot_lookup's rule, for the CPU suite: the slot of `want` (= offset + 1) in a table built by
// hot_table_build_host / hot_select_kernel, or ~0.  Same hashes (hot_b1 / hot_b2), same "second bucket only when the first is full".
uint32_t hot_lookup_host(const uint32_t* tab, uint32_t want) {
  const uint32_t d = want - 1u, b1 = hot_b1(d);
  for (int j = 0; j < 4; j++) if (tab[4 * b1 + j] == want) return 4 * b1 + j;
  if (tab[4 * b1 + 3] == 0u) return 0xFFFFFFFFu;
  const uint32_t b2 = hot_b2(d);
  for (int j = 0; j < 4; j++) if (tab[4 * b2 + j] == want) return 4 * b2 + j;
  return 0xFFFFFFFFu;
}
uint32_t hot_table_slots() { return kHotSlots; }
uint32_t hot_max_keys() { return kHotMaxKeys; }
// Host-side build of the same table, sequentially over the caller's (sorted) keys so that every rank ends up with the same slot for
// the same key.  keys_plus1: offset + 1; tab: hot_table_slots() zeroed words.  Returns the keys placed.
uint32_t hot_table_build_host(const uint32_t* keys_plus1, uint32_t n, uint32_t* tab) {
  uint32_t placed = 0;
  for (uint32_t i = 0; i < n; i++) {
    const uint32_t k = keys_plus1[i];
    if (k == 0 || k == 0xFFFFFFFFu) continue;
    bool done = false;
    for (int t = 0; t < 2 && !done; t++) {
      const uint32_t b = t ? hot_b2(k - 1) : hot_b1(k - 1);
      for (int j = 0; j < 4 && !done; j++)
        if (tab[4 * b + j] == 0) { tab[4 * b + j] = k; done = true; }
    }
    placed += done;
  }
  return placed;
}
uint32_t hot_sample_slots() { return kHotSampleSlots; }

// (offset + 1, sample count) of every sampled key seen at least min_count times -> list (at most cap entries)
__global__ void hot_extract_kernel(const uint32_t* __restrict__ tab_key, const uint32_t* __restrict__ tab_cnt, uint32_t slots, uint32_t min_count,
                                   uint32_t* __restrict__ out_key, uint32_t* __restrict__ out_cnt, uint32_t cap, uint32_t* __restrict__ out_n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < slots; i += gridDim.x * blockDim.x) {
    const uint32_t k = tab_key[i], c = tab_cnt[i];
    if (k && c >= min_count) {
      const uint32_t pos = atomicAdd(out_n, 1u);
      if (pos < cap) { out_key[pos] = k; out_cnt[pos] = c; }
    }
  }
}

// Sample one rank's probe shard (2^18 rows) and list the keys seen >= 3 times.  scratch: 2 x hot_sample_slots() + 2 x cap + 4 u32.
olap_status hot_candidates(olap_ctx* ctx, const uint64_t* keys, uint64_t n, int64_t gmin, uint64_t grange, uint32_t* scratch, uint32_t cap,
                           uint32_t** out_key, uint32_t** out_cnt, uint32_t** out_n, cudaStream_t s) {
  HotArgs h;
  memset(&h, 0, sizeof(h));
  h.n_samples = 1u << 18;   // per rank; the ranks' counts are summed before the cut
  h.keys = keys; h.n = n; h.stride = std::max<uint64_t>(n / h.n_samples, 1);
  h.rmin = (uint64_t)gmin; h.rrange = grange;
  h.tab_slots = kHotSampleSlots;
  h.tab_key = scratch; h.tab_cnt = scratch + h.tab_slots;
  *out_key = scratch + 2 * h.tab_slots; *out_cnt = *out_key + cap; *out_n = *out_cnt + cap;
  OLAP_CUDA(cudaMemsetAsync(scratch, 0, (2ULL * h.tab_slots + 2ULL * cap + 4) * 4, s));
  if (n) {
    hot_sample_kernel<<<(h.n_samples + 255) / 256, 256, 0, s>>>(h);
    hot_extract_kernel<<<64, 256, 0, s>>>(h.tab_key, h.tab_cnt, h.tab_slots, 3, *out_key, *out_cnt, cap, *out_n);
    count_launch(ctx, 2);
  }
  OLAP_CUDA(cudaGetLastError());
  return OLAP_OK;
}

// The build row of every hot key this rank owns, in (source rank, source row) form: one CTA per slot scans the key's range
// partition of the partitioned build side.  out[slot] = ((src_rank << 32) | src_row) + 1, or 0 when the key is not in this rank's
// slice / has no build row (the ranks then take the maximum).
__global__ void __launch_bounds__(256) hot_resolve_global_kernel(const uint32_t* __restrict__ hot_keys, uint64_t* __restrict__ out, const uint64_t* brows,
                                                                 const uint32_t* bcur, uint32_t bcap_fine, const uint64_t* slab_b, uint32_t slab_bcap,
                                                                 uint32_t q, uint64_t slice, uint32_t rank) {
  __shared__ uint32_t s_found;
  const uint32_t slot = blockIdx.x, k = hot_keys[slot];
  if (threadIdx.x == 0) out[slot] = 0;
  if (k == 0 || k == 0xFFFFFFFFu) return;
  const uint64_t d = k - 1;
  if (d / slice != rank) return;
  const uint32_t dl = (uint32_t)(d - (uint64_t)rank * slice), p = dl >> kRShift, n = min(bcur[p], bcap_fine);
  if (threadIdx.x == 0) s_found = kREmpty;
  __syncthreads();
  const uint64_t* rows = brows + (uint64_t)p * bcap_fine;
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    const uint64_t v = rows[i];
    if ((uint32_t)(v >> 32) == dl) s_found = (uint32_t)v;
  }
  __syncthreads();
  if (threadIdx.x == 0 && s_found != kREmpty) {
    const uint32_t pos = s_found;   // position in this rank's build slab
    out[slot] = (((uint64_t)((pos / slab_bcap) / q) << 32) | (uint32_t)slab_b[pos]) + 1;
  }
}

olap_status hot_resolve_global(olap_ctx* ctx, const uint32_t* hot_keys, uint64_t* out, const DJBuild& b, const uint64_t* slab_b, uint32_t slab_bcap,
                               uint32_t q, uint64_t slice, int rank, cudaStream_t s) {
  hot_resolve_global_kernel<<<kHotSlots, 256, 0, s>>>(hot_keys, out, b.brows, b.bctl, kRKeys, slab_b, slab_bcap, q, slice, (uint32_t)rank);
  count_launch(ctx);
  OLAP_CUDA(cudaGetLastError());
  return OLAP_OK;
}

// merged[slot] == 0: no rank holds a build row for the key -> tombstone (its probe rows travel the normal way and match nothing);
// hot_ids[slot] = slot: the pair emitted for a hot probe row names the slot, whose build row is merged[slot] - 1
__global__ void hot_finalize_kernel(uint32_t* hot_keys, uint32_t* hot_ids, const uint64_t* merged) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kHotSlots) return;
  hot_ids[i] = i;
  if (hot_keys[i] && merged[i] == 0) hot_keys[i] = 0xFFFFFFFFu;
}

olap_status hot_finalize(olap_ctx* ctx, uint32_t* hot_keys, uint32_t* hot_ids, const uint64_t* merged, cudaStream_t s) {
  hot_finalize_kernel<<<kHotSlots / 256, 256, 0, s>>>(hot_keys, hot_ids, merged);
  count_launch(ctx);
  OLAP_CUDA(cudaGetLastError());
  return OLAP_OK;
}

}  // namespace olap
