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

#include <algorithm>
#include <cstdlib>

namespace olap {

// ------------------------------------------------------------------------------------------- partitioning
constexpr int kPartThreads = 512;
constexpr int kPartTile = 4096;          // rows staged per CTA iteration (8 per thread)
constexpr int kMaxPartBits = 8;          // <= 256 partitions per pass
constexpr int kMaxPay = 4;               // 8-byte payload columns carried along (group-by)

struct PartArgs {
  const uint64_t* keys;
  const uint32_t* ids_in;                // NULL: implicit row index
  const uint64_t* pay_in[kMaxPay];
  uint64_t n, chunk;                     // rows, rows per CTA (multiple of kPartTile)
  int bits;
  uint32_t* hist;                        // [P][gridDim.x]: counts, then exclusive offsets after the scan
  uint64_t* out_keys;
  uint32_t* out_ids;
  uint64_t* pay_out[kMaxPay];
  TermPack terms;                        // rows failing the conjunction are dropped (fused filter for group-by)
  bool vec;                              // keys are 32-byte aligned: 256-bit loads
};

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
    for (int e = 0; e < 4; e++) if (ok[e]) atomicAdd(&s_h[mix64(k[e]) >> sh], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < P; i += kPartThreads) a.hist[(uint64_t)i * gridDim.x + blockIdx.x] = s_h[i];
}

// exclusive scan of hist[P * G] in place (partition-major, so element (p, c) becomes the first output position of
// CTA c's rows of partition p); part_begin[p] = offset of (p, 0); part_begin[P] = total.  One CTA.
__global__ void __launch_bounds__(1024) part_scan_kernel(uint32_t* hist, uint32_t m, uint32_t G, uint32_t P, uint32_t* part_begin) {
  __shared__ uint32_t s_w[32];
  __shared__ uint32_t s_total;
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
  if (part_begin) {
    for (uint32_t p = threadIdx.x; p < P; p += 1024) part_begin[p] = hist[(uint64_t)p * G];
    if (threadIdx.x == 0) part_begin[P] = s_total;
  }
}

template <int NPAY, bool IDS>
__global__ void __launch_bounds__(kPartThreads, NPAY > 1 ? 1 : 2) partition_kernel(PartArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* s_key = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* s_pay = s_key + kPartTile;                               // [NPAY][kPartTile]
  uint32_t* s_id = reinterpret_cast<uint32_t*>(s_pay + (size_t)NPAY * kPartTile);
  uint8_t* s_part = reinterpret_cast<uint8_t*>(s_id + (IDS ? kPartTile : 0));
  __shared__ uint32_t s_cnt[1 << kMaxPartBits], s_off[1 << kMaxPartBits], s_gbase[1 << kMaxPartBits], s_cursor[1 << kMaxPartBits];
  __shared__ uint32_t s_wtot[8], s_total;

  const int P = 1 << a.bits, sh = 64 - a.bits;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < P; i += kPartThreads) { s_cnt[i] = 0; s_cursor[i] = a.hist[(uint64_t)i * gridDim.x + blockIdx.x]; }
  __syncthreads();
  const uint64_t begin = blockIdx.x * a.chunk, end = min(a.n, begin + a.chunk);

  for (uint64_t tile = begin; tile < end; tile += kPartTile) {
    // phase 1: load 8 rows, pick the partition from the high hash bits, rank inside the tile with one smem atomic
    uint64_t k[2][4]; bool ok[2][4]; uint32_t part[2][4], rank[2][4];
#pragma unroll
    for (int j = 0; j < 2; j++) load_keys4(a, tile + (uint64_t)(j * kPartThreads + tid) * 4, end, k[j], ok[j]);
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
      for (int e = 0; e < 4; e++)
        if (ok[j][e]) { part[j][e] = (uint32_t)(mix64(k[j][e]) >> sh); rank[j][e] = atomicAdd(&s_cnt[part[j][e]], 1u); }
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
          s_part[pos] = (uint8_t)part[j][e];
          if (IDS) s_id[pos] = a.ids_in ? ld_stream_u32(a.ids_in + row) : (uint32_t)row;
#pragma unroll
          for (int q = 0; q < NPAY; q++) s_pay[(size_t)q * kPartTile + pos] = ld_stream_u64(a.pay_in[q] + row);
        }
    __syncthreads();
    // phase 4: each partition's run leaves with coalesced stores (write-combined per CTA, no global atomics)
    const uint32_t total = s_total;
    for (uint32_t i = tid; i < total; i += kPartThreads) {
      const uint32_t dst = s_gbase[s_part[i]] + i;
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
  unsigned grid = (unsigned)ctx->sm_count * 2;
  uint64_t tiles = (n + kPartTile - 1) / kPartTile;
  if (tiles < grid) grid = (unsigned)(tiles ? tiles : 1);
  uint64_t chunk = ((tiles + grid - 1) / grid) * kPartTile;
  return {grid, chunk};
}

template <int NPAY, bool IDS>
static olap_status launch_partition(olap_ctx* ctx, PartArgs& a, unsigned grid, cudaStream_t s) {
  const size_t smem = (size_t)kPartTile * (8 + 8 * NPAY + (IDS ? 4 : 0) + 1);
  OLAP_CUDA(cudaFuncSetAttribute(partition_kernel<NPAY, IDS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  partition_kernel<NPAY, IDS><<<grid, kPartThreads, smem, s>>>(a);
  count_launch(ctx);
  return OLAP_OK;
}

// histogram + scan + scatter of one relation.  hist must hold P * grid u32.
static olap_status run_partition(olap_ctx* ctx, PartArgs& a, int npay, bool ids, uint32_t* part_begin, cudaStream_t s) {
  PartPlan pl = plan_partition(ctx, a.n);
  a.chunk = pl.chunk;
  a.vec = aligned32(a.keys);
  const uint32_t P = 1u << a.bits;
  part_hist_kernel<<<pl.grid, kPartThreads, 0, s>>>(a);
  part_scan_kernel<<<1, 1024, 0, s>>>(a.hist, P * pl.grid, pl.grid, P, part_begin);
  count_launch(ctx, 2);
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
  return OLAP_OK;
}

// ------------------------------------------------------------------------------------------- build / probe
// 16-byte slot {key, build row, generation}; a slot is occupied iff generation == the current one.
constexpr int kBuildThreads = 256;
constexpr int kProbeThreads = 256;
constexpr int kProbeRows = 8;
constexpr int kProbeTile = kProbeThreads * kProbeRows;   // 2048 probe rows per tile
constexpr int kProbeStage = 4096;                        // staged pairs per tile (32 KB)

__global__ void __launch_bounds__(kBuildThreads) build_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ ids,
                                                              uint64_t begin, uint64_t end, uint4* __restrict__ table, uint64_t mask,
                                                              uint32_t gen) {
  const uint64_t stride = (uint64_t)gridDim.x * kBuildThreads;
  for (uint64_t row = begin + blockIdx.x * (uint64_t)kBuildThreads + threadIdx.x; row < end; row += stride) {
    const uint64_t key = ld_stream_u64(keys + row);
    const uint32_t id = ids ? ld_stream_u32(ids + row) : (uint32_t)row;
    uint64_t s = mix64(key) & mask;
    for (;;) {
      // generations only grow, so atomicMax claims a stale slot and leaves a live one untouched in one L2 atomic
      const uint32_t old = atomicMax(reinterpret_cast<uint32_t*>(table + s) + 3, gen);
      if (old != gen) {
        table[s] = make_uint4((uint32_t)key, (uint32_t)(key >> 32), id, gen);
        break;
      }
      s = (s + 1) & mask;
    }
  }
}

__global__ void __launch_bounds__(kProbeThreads) probe_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ ids,
                                                              uint64_t begin, uint64_t end, const uint4* __restrict__ table, uint64_t mask,
                                                              uint32_t gen, olap_pair* __restrict__ out, uint64_t max_matches,
                                                              unsigned long long* __restrict__ count) {
  __shared__ uint2 s_buf[kProbeStage];
  __shared__ uint32_t s_n;
  __shared__ unsigned long long s_base;
  const int tid = threadIdx.x;
  for (uint64_t tile = begin + blockIdx.x * (uint64_t)kProbeTile; tile < end; tile += (uint64_t)gridDim.x * kProbeTile) {
    if (tid == 0) s_n = 0;
    __syncthreads();
    uint64_t key[kProbeRows];
    uint64_t slot[kProbeRows];
    uint4 ent[kProbeRows];
#pragma unroll
    for (int j = 0; j < kProbeRows; j++) {  // all first-slot loads of the tile are in flight before any is consumed
      const uint64_t row = tile + (uint64_t)j * kProbeThreads + tid;
      key[j] = row < end ? ld_stream_u64(keys + row) : 0;
      slot[j] = mix64(key[j]) & mask;
      ent[j] = row < end ? __ldg(table + slot[j]) : make_uint4(0, 0, 0, gen - 1);
    }
#pragma unroll
    for (int j = 0; j < kProbeRows; j++) {
      const uint64_t row = tile + (uint64_t)j * kProbeThreads + tid;
      uint4 e = ent[j];
      uint64_t s = slot[j];
      uint32_t pid = 0xFFFFFFFFu;
      while (e.w == gen) {
        if (e.x == (uint32_t)key[j] && e.y == (uint32_t)(key[j] >> 32)) {
          if (pid == 0xFFFFFFFFu) pid = ids ? ld_stream_u32(ids + row) : (uint32_t)row;
          const uint32_t idx = atomicAdd(&s_n, 1u);
          if (idx < kProbeStage) s_buf[idx] = make_uint2(e.z, pid);
          else {  // more matches than the stage holds (duplicate-heavy build side): reserve directly
            const unsigned long long pos = atomicAdd(count, 1ULL);
            if (out && pos < max_matches) out[pos] = olap_pair{e.z, pid};
          }
        }
        s = (s + 1) & mask;
        e = __ldg(table + s);
      }
    }
    __syncthreads();
    const uint32_t n = min(s_n, (uint32_t)kProbeStage);
    if (tid == 0 && n) s_base = atomicAdd(count, (unsigned long long)n);  // one global atomic per tile
    __syncthreads();
    if (out && n) {
      const unsigned long long base = s_base;
      for (uint32_t i = tid; i < n; i += kProbeThreads) {
        const unsigned long long pos = base + i;
        if (pos < max_matches) st_stream_u64(out + pos, (uint64_t)s_buf[i].x | ((uint64_t)s_buf[i].y << 32));
      }
    }
    __syncthreads();
  }
}

static uint64_t pow2_at_least(uint64_t x) { uint64_t p = 1; while (p < x) p <<= 1; return p; }

// table arena: grown on demand, zeroed once; slots are recycled through the generation tag
static olap_status ensure_arena(olap_ctx* ctx, uint64_t bytes) {
  if (ctx->table_arena_bytes >= bytes) return OLAP_OK;
  if (ctx->table_arena) { OLAP_CUDA(cudaDeviceSynchronize()); OLAP_CUDA(cudaFree(ctx->table_arena)); ctx->table_arena = nullptr; ctx->table_arena_bytes = 0; }
  OLAP_CUDA(cudaMalloc(&ctx->table_arena, bytes));
  OLAP_CUDA(cudaMemset(ctx->table_arena, 0, bytes));
  ctx->table_arena_bytes = bytes;
  ctx->table_generation = 0;
  return OLAP_OK;
}
static olap_status next_generations(olap_ctx* ctx, uint32_t n, uint32_t* first) {
  if ((uint64_t)ctx->table_generation + n >= 0xFFFFFFF0ULL) {
    OLAP_CUDA(cudaDeviceSynchronize());
    OLAP_CUDA(cudaMemset(ctx->table_arena, 0, ctx->table_arena_bytes));
    ctx->table_generation = 0;
  }
  *first = ctx->table_generation + 1;
  ctx->table_generation += n;
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

static olap_status join_broadcast(olap_ctx* ctx, const uint64_t* bk, uint64_t nb, const uint64_t* pk, uint64_t np, olap_pair* out,
                                  uint64_t max_matches, unsigned long long* d_count, cudaStream_t s) {
  const uint64_t cap = pow2_at_least(std::max<uint64_t>(2 * nb, 1024));
  OLAP_TRY(ensure_arena(ctx, cap * 16));
  uint32_t gen;
  OLAP_TRY(next_generations(ctx, 1, &gen));
  uint4* table = (uint4*)ctx->table_arena;
  build_kernel<<<grid_rows(ctx, nb, kBuildThreads * 4, 8), kBuildThreads, 0, s>>>(bk, nullptr, 0, nb, table, cap - 1, gen);
  phase_mark(ctx, s, "build");
  probe_kernel<<<grid_rows(ctx, np, kProbeTile, 4), kProbeThreads, 0, s>>>(pk, nullptr, 0, np, table, cap - 1, gen, out, max_matches, d_count);
  phase_mark(ctx, s, "probe");
  count_launch(ctx, 2);
  OLAP_CUDA(cudaGetLastError());
  return OLAP_OK;
}

static olap_status join_radix(olap_ctx* ctx, const uint64_t* bk, uint64_t nb, const uint64_t* pk, uint64_t np, olap_pair* out,
                              uint64_t max_matches, unsigned long long* d_count, cudaStream_t s) {
  // partitions of ~512 Ki build rows: a 16 MB table each, so a few tables in flight stay L2-resident
  const uint64_t target = env_u64("OLAP_JOIN_PART_ROWS", 512 * 1024);
  int bits = 1;
  while (bits < kMaxPartBits && ((nb + target - 1) / target) > (1ULL << bits)) bits++;
  const uint32_t P = 1u << bits;
  const unsigned gmax = (unsigned)ctx->sm_count * 2;

  Scratch sc(ctx, s);
  uint64_t *pkb = nullptr, *pkp = nullptr;
  uint32_t *pib = nullptr, *pip = nullptr, *hist_b = nullptr, *hist_p = nullptr, *beg = nullptr;
  OLAP_TRY(sc.get<uint64_t>(nb, &pkb));
  OLAP_TRY(sc.get<uint32_t>(nb, &pib));
  OLAP_TRY(sc.get<uint64_t>(np, &pkp));
  OLAP_TRY(sc.get<uint32_t>(np, &pip));
  OLAP_TRY(sc.get<uint32_t>((uint64_t)P * gmax, &hist_b));
  OLAP_TRY(sc.get<uint32_t>((uint64_t)P * gmax, &hist_p));
  OLAP_TRY(sc.get<uint32_t>(2 * (P + 1), &beg));

  PartArgs a;
  memset(&a, 0, sizeof(a));
  a.bits = bits;
  a.keys = bk; a.n = nb; a.hist = hist_b; a.out_keys = pkb; a.out_ids = pib;
  OLAP_TRY(run_partition(ctx, a, 0, true, beg, s));
  phase_mark(ctx, s, "partition_build");
  a.keys = pk; a.n = np; a.hist = hist_p; a.out_keys = pkp; a.out_ids = pip;
  OLAP_TRY(run_partition(ctx, a, 0, true, beg + P + 1, s));
  phase_mark(ctx, s, "partition_probe");

  std::vector<uint32_t> h(2 * (P + 1));
  OLAP_CUDA(cudaMemcpyAsync(h.data(), beg, h.size() * 4, cudaMemcpyDeviceToHost, s));
  OLAP_CUDA(cudaStreamSynchronize(s));
  const uint32_t* hb = h.data();
  const uint32_t* hp = h.data() + P + 1;
  uint64_t max_part = 0;
  for (uint32_t p = 0; p < P; p++) max_part = std::max<uint64_t>(max_part, hb[p + 1] - hb[p]);
  const uint64_t cap = pow2_at_least(std::max<uint64_t>(2 * max_part, 1024));
  int S = (int)std::min<uint64_t>(ctx->streams.size(), env_u64("OLAP_JOIN_STREAMS", 3));
  if (S < 1) S = 1;
  if (S > 8) S = 8;
  OLAP_TRY(ensure_arena(ctx, (uint64_t)S * cap * 16));
  uint32_t gen0;
  OLAP_TRY(next_generations(ctx, P, &gen0));

  // fork: tables of different partitions are built and probed on S side streams so their ramps and tails overlap
  OLAP_CUDA(cudaEventRecord(ctx->fork_ev, s));
  for (int k = 0; k < S; k++) OLAP_CUDA(cudaStreamWaitEvent(ctx->streams[k], ctx->fork_ev, 0));
  uint64_t launches = 0;
  for (uint32_t p = 0; p < P; p++) {
    const uint64_t b0 = hb[p], b1 = hb[p + 1], p0 = hp[p], p1 = hp[p + 1];
    if (b1 == b0 || p1 == p0) continue;  // inner join: an empty side yields nothing
    const int k = p % S;
    cudaStream_t st = ctx->streams[k];
    uint4* table = (uint4*)ctx->table_arena + (uint64_t)k * cap;
    build_kernel<<<grid_rows(ctx, b1 - b0, kBuildThreads * 4, 8), kBuildThreads, 0, st>>>(pkb, pib, b0, b1, table, cap - 1, gen0 + p);
    probe_kernel<<<grid_rows(ctx, p1 - p0, kProbeTile, 4), kProbeThreads, 0, st>>>(pkp, pip, p0, p1, table, cap - 1, gen0 + p, out, max_matches,
                                                                                d_count);
    launches += 2;
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

// Partitioned group-by hook (groupby.cu); filled in by a later milestone.
struct GbParams;
struct OutSpec;
olap_status group_by_partitioned(olap_ctx*, GbParams&, const OutSpec&, uint64_t, int64_t*, uint64_t*, uint64_t, unsigned long long*,
                                 cudaStream_t, bool* handled) {
  *handled = false;
  return OLAP_OK;
}

}  // namespace olap

using namespace olap;

extern "C" {

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
  if (n_build && n_probe) {  // executor.rs:231-233: an empty side yields no rows
    if (strategy == OLAP_JOIN_BROADCAST)
      OLAP_TRY(join_broadcast(ctx, (const uint64_t*)build_keys, n_build, (const uint64_t*)probe_keys, n_probe, out_pairs, max_matches, d_count, s));
    else
      OLAP_TRY(join_radix(ctx, (const uint64_t*)build_keys, n_build, (const uint64_t*)probe_keys, n_probe, out_pairs, max_matches, d_count, s));
  }
  phase_end(ctx, s);
  if (host_count) {
    OLAP_CUDA(cudaMemcpyAsync(host_count, d_count, 8, cudaMemcpyDeviceToHost, s));
    OLAP_CUDA(cudaStreamSynchronize(s));
    if (out_pairs && *host_count > max_matches) {
      set_error("join produced %llu pairs but max_matches is %llu (count is exact, pairs truncated)", (unsigned long long)*host_count,
                (unsigned long long)max_matches);
      return OLAP_ERR_OVERFLOW;
    }
  }
  return OLAP_OK;
}

olap_status olap_radix_partition_i64(olap_ctx* ctx, const int64_t* keys, const uint32_t* row_ids, uint64_t n_rows, int32_t bits,
                                     int64_t* out_keys, uint32_t* out_row_ids, uint32_t* out_part_begin, void* stream) {
  OLAP_REQUIRE(ctx, "ctx is NULL");
  OLAP_REQUIRE(bits >= 1 && bits <= kMaxPartBits, "bits must be in [1, 8]");
  OLAP_REQUIRE(n_rows < 0xFFFFFFFFULL, "row ids are u32");
  OLAP_REQUIRE(out_part_begin && (n_rows == 0 || (keys && out_keys && out_row_ids)), "NULL pointer");
  std::lock_guard<std::mutex> g(ctx->mu);
  OLAP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream_of(stream);
  const uint32_t P = 1u << bits;
  if (n_rows == 0) { OLAP_CUDA(cudaMemsetAsync(out_part_begin, 0, (P + 1) * 4, s)); return OLAP_OK; }
  Scratch sc(ctx, s);
  uint32_t* hist = nullptr;
  OLAP_TRY(sc.get<uint32_t>((uint64_t)P * ctx->sm_count * 2, &hist));
  PartArgs a;
  memset(&a, 0, sizeof(a));
  a.bits = bits; a.keys = (const uint64_t*)keys; a.ids_in = row_ids; a.n = n_rows; a.hist = hist;
  a.out_keys = (uint64_t*)out_keys; a.out_ids = out_row_ids;
  return run_partition(ctx, a, 0, true, out_part_begin, s);
}

}  // extern "C"
