// Key-range-sharded hash join across the GPUs of one box (the reference has no multi-GPU path, README.md:414).
//
// When the build keys are dense (the same test as the 1-GPU direct join, on the GLOBAL min / max) the key domain is cut into
// one contiguous slice per rank, each slice into Q sub-ranges of F x 2^15 keys.  A rank's shuffle is the 1-GPU partition pass
// with a different destination: the rows of sub-range j go straight into region (source rank, j % Q) of the OWNING rank's
// receive slab, over NVLink, run by run (djoin.cu: rpart_kernel<raw, ., PEER>) -- already in the packed 8-byte form
// {key - slice start, source row id} and already split Q ways, so the owner only runs the second partition level (F ranges per
// region) and the shared-memory build + probe.  The slab regions have a fixed capacity and a single writer each: no count pass,
// no count collective, no remote atomics; the row counts are published by a one-CTA kernel behind the shuffle kernel.
// Two ways to move a chunk (OLAP_SHARD_MODE, fixed when the shard is created):
//   "ce"  (default) the partition kernel writes the rows of the OTHER ranks into a local outbox laid out like the receivers'
//         regions, and one peer cudaMemcpyAsync per destination (a copy engine, no SM) moves each block of Q regions; the rows the
//         rank keeps go straight into its own slab.  The SMs only partition (at the 1-GPU coarse pass's speed) and join, NVLink
//         runs at the copy engines' rate next to them.
//   "tma" one kernel does both: shuffle_kernel (djoin.cu) stages each tile in shared memory and stores every run into the peer's
//         slab with cp.async.bulk, one CTA per SM next to the join kernels.  Measured at 2 GPUs: 540 GB/s per direction, and the
//         join kernels slow down by a third while it runs -- the SMs, not NVLink or HBM, are what both sides need.
// The probe side moves in chunks: chunk k + 1 crosses NVLink (stream "comm") while chunk k is partitioned and joined (stream
// "comp").  The only cross-rank synchronisation is "every rank's shuffle of chunk k has finished": stream events inside one
// process (olap_sharded_hash_join_i64), or a collective of the host language between the per-phase calls (sharded.py: NCCL).
//
// Skew (Zipf): heavy-hitter keys are agreed on by all ranks (olap_shard_hot_*), their build rows are replicated (a <= 2048-entry
// table), and probe rows carrying them never leave their source rank: the shuffle kernel joins them on the spot.
#include "common.cuh"
#include "djoin.cuh"

#include <algorithm>
#include <unordered_map>

namespace olap {
olap_status key_min_max(olap_ctx* ctx, const uint64_t* keys, uint64_t n, cudaStream_t s, int64_t* mn, int64_t* mx);  // join.cu
}
using namespace olap;

constexpr int kShardMaxWorld = 8;
constexpr uint32_t kShardRegions = 256;       // regions per slab side (>= world * Q)
constexpr uint32_t kShardSlackPct = 30;       // words reserved per side above the rows it can receive
constexpr uint32_t kHotCand = 16384;          // heavy-hitter candidates a rank reports

struct olap_shard {
  olap_ctx* ctx = nullptr;
  int rank = 0, world = 1;
  uint64_t build_cap = 0, probe_cap = 0, chunk_cap = 0;
  int chunks = 1;
  // slab layout (identical on every rank): [build regions][probe chunk 0 regions] ... [counts: u32 x 256 x (1 + chunks)]
  uint64_t words_b = 0, words_p = 0, off_counts = 0, off_sync = 0, slab_bytes = 0;
  uint32_t epoch = 0, epoch_waited = 0;   // signals sent / waits issued so far (olap_shard_signal / _wait)
  void* slabs[kShardMaxWorld] = {};
  void* own_slab_alloc = nullptr;
  bool ce_mode = true;          // copy-engine exchange through a local outbox (else: TMA stores from the shuffle kernel)
  cudaEvent_t ce_fork = nullptr, ce_join[8] = {};
  uint64_t* outbox = nullptr;   // ce: [build regions][probe chunk regions ...], laid out like a slab
  // geometry of the current join
  bool planned = false;
  uint32_t Q = 0, F = 0, subw = 0, bcap = 0, pcap = 0;
  int64_t gmin = 0;
  uint64_t grange = 0, slice = 0;
  // local device state: sender cursors [256 x (1 + chunks)] | flags [kFlagWords x (2 + chunks)] | pair count (u64) | hot count (u64)
  uint32_t* local = nullptr;
  uint64_t local_words = 0;
  DJBuild dj;
  bool built = false;
  DJProbe pg;                   // the current group of probe chunks: added chunk by chunk, joined once per group
  int group = 1, last_group = 1;   // chunks per group; the LAST group is short: its join is the tail nothing overlaps
  // heavy hitters: hot = [kHotSlots keys | kHotSlots ids | u64 merged[kHotSlots] | sampling scratch]
  uint32_t* hot = nullptr;
  bool hot_on = false, hot_ready = false, skew_room = false;   // skew_room: plan(hot != 0) -- wider regions, with (1) or without (2) a heavy-hitter table
  olap_pair* hot_pairs = nullptr; uint64_t hot_max = 0;
  uint32_t* hot_keys() const { return hot; }
  uint32_t* hot_ids() const { return hot + hot_table_slots(); }
  uint64_t* hot_glob() const { return (uint64_t*)(hot + 2 * hot_table_slots()); }
  uint32_t* hot_scratch() const { return hot + 6 * hot_table_slots(); }

  uint64_t* outbox_b() const { return outbox; }
  uint64_t* outbox_p(int c) const { return outbox + words_b + (uint64_t)c * words_p; }
  uint64_t* regions_b(int r) const { return (uint64_t*)slabs[r]; }
  uint64_t* regions_p(int r, int c) const { return (uint64_t*)slabs[r] + words_b + (uint64_t)c * words_p; }
  uint32_t* counts_b(int r) const { return (uint32_t*)((char*)slabs[r] + off_counts); }
  uint32_t* counts_p(int r, int c) const { return counts_b(r) + kShardRegions * (1 + c); }
  uint32_t* sync_words(int r) const { return (uint32_t*)((char*)slabs[r] + off_sync); }
  uint32_t* cursor(int block) const { return local + kShardRegions * block; }
  uint32_t* flags(int block) const { return local + kShardRegions * (1 + chunks) + kFlagWords * block; }
  unsigned long long* pair_count() const { return (unsigned long long*)(local + kShardRegions * (1 + chunks) + kFlagWords * (2 + chunks)); }
  unsigned long long* hot_count() const { return pair_count() + 1; }
};

// Geometry of the range sharding for a dense key domain of `grange` keys over `world` ranks: one contiguous slice per rank, cut into Q
// sub-ranges of F x 2^15 keys (F <= 256, world x Q <= 256 regions per slab side).  Host arithmetic only.
struct ShardGeometry { uint32_t Q = 0, F = 0, subw = 0; uint64_t slice = 0; };
static olap_status shard_geometry(int world, uint64_t grange, ShardGeometry* g) {
  const uint64_t W = (uint64_t)world;
  const uint64_t slice_min = (grange + W - 1) / W;
  const uint64_t pl = std::max<uint64_t>((slice_min + (1ULL << 15) - 1) >> 15, 1);
  const uint64_t Q = (pl + 255) / 256, F = (pl + Q - 1) / Q;
  if (W * Q > kShardRegions) { set_error("key domain too wide for the range-sharded join: %llu sub-ranges per rank", (unsigned long long)Q); return OLAP_ERR_UNSUPPORTED; }
  g->Q = (uint32_t)Q; g->F = (uint32_t)F; g->subw = (uint32_t)(F << 15); g->slice = Q * F << 15;
  if (g->slice * W >= 0xFFFFFFFFULL) { set_error("key domain too wide for 32-bit offsets"); return OLAP_ERR_UNSUPPORTED; }
  return OLAP_OK;
}

static uint64_t reserve_words(uint64_t rows) { return (rows + rows * kShardSlackPct / 100 + kShardRegions * 4096ULL + 3) & ~3ULL; }

static unsigned shuffle_grid(olap_ctx* ctx) {
  const char* v = getenv("OLAP_SHARD_SHUFFLE_CTAS");
  return v && *v ? (unsigned)strtoul(v, nullptr, 10) : (unsigned)ctx->sm_count;
}
// CTAs of the local partition pass while a shuffle CTA of the next chunk sits on every SM: 2 per SM (2 x 20 K + 24 K registers);
// nonzero also selects the 768-thread join kernel.  OLAP_SHARD_LOCAL_CTAS=0: the local kernels take the whole GPU.
static unsigned local_grid(olap_ctx* ctx, bool ce_mode) {
  const char* v = getenv("OLAP_SHARD_LOCAL_CTAS");
  if (v && *v) return (unsigned)strtoul(v, nullptr, 10);
  return ce_mode ? 0u : (unsigned)ctx->sm_count * 2;   // ce: nothing else wants the SMs while a local kernel runs
}

// (left, right) = positions in the receive slabs -> global row ids (source rank's row offset + source row id)
__global__ void shard_pairs_to_global_kernel(const olap_pair* __restrict__ pairs, uint64_t n, const uint64_t* __restrict__ rb, uint32_t bcap,
                                             const uint64_t* __restrict__ rp, uint64_t words_p, uint32_t pcap, uint32_t q,
                                             const uint64_t* __restrict__ boff, const uint64_t* __restrict__ poff, uint64_t* __restrict__ left,
                                             uint64_t* __restrict__ right) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const olap_pair p = pairs[i];
    const uint32_t gb = p.left_row / bcap;
    left[i] = boff[gb / q] + (uint32_t)rb[p.left_row];
    const uint64_t c = p.right_row / words_p, rem = p.right_row - c * words_p;
    right[i] = poff[(uint32_t)(rem / pcap) / q] + (uint32_t)rp[p.right_row];
  }
}

// Cross-GPU synchronisation through the slabs, in two halves that may sit on different streams:
//   signal  one warp tells every peer "my work queued on this stream before this kernel is done" (a release store of the epoch
//           into the peer's slab);
//   wait    one warp spins until every peer has said so for the same epoch (acquire loads of its own words).
// The scatter stream only ever signals -- it never waits for a peer, so copies of chunk k + 1 follow those of chunk k back to
// back -- and the stream that consumes a chunk waits.  32 threads and a handful of registers: resident next to anything (an NCCL
// all-reduce used as the barrier could not start while the join kernel held the SM's registers and stalled the shuffle stream
// for the length of a join, 2 ms per group).  One process per GPU only: ranks that share a GPU must not spin on each other (the
// one-process driver uses stream events).
__global__ void shard_signal_kernel(uint32_t epoch, uint32_t rank, uint32_t world, uint32_t* p0, uint32_t* p1, uint32_t* p2, uint32_t* p3, uint32_t* p4,
                                    uint32_t* p5, uint32_t* p6, uint32_t* p7) {
  uint32_t* peers[8] = {p0, p1, p2, p3, p4, p5, p6, p7};
  const uint32_t t = threadIdx.x;
  if (t < world) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peers[t] + rank), "r"(epoch) : "memory");
  }
}
__global__ void shard_wait_kernel(uint32_t epoch, uint32_t world, uint32_t* mine, uint32_t* local_flags) {
  const uint32_t t = threadIdx.x;
  if (t < world) {
    const long long t0 = clock64();
    for (;;) {
      uint32_t v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine + t) : "memory");
      if ((int32_t)(v - epoch) >= 0) break;
      if (clock64() - t0 > 8000000000LL) { local_flags[3] = 1; break; }  // ~4 s: a peer never arrived; give up instead of hanging the GPU
      __nanosleep(200);
    }
    __threadfence_system();
  }
}

extern "C" {

// "my work queued on `stream` before this call has completed", told to every rank.  Every rank issues the same sequence of
// signals; olap_shard_wait number n returns once every rank's signal number n has arrived.
olap_status olap_shard_signal(olap_shard* sh, void* stream) {
  OLAP_REQUIRE(sh && sh->slabs[0], "olap_shard_attach first");
  olap_ctx* ctx = sh->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  OLAP_CUDA(cudaSetDevice(ctx->device));
  uint32_t* p[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  for (int r = 0; r < sh->world; r++) p[r] = sh->sync_words(r);
  sh->epoch++;
  shard_signal_kernel<<<1, 32, 0, ctx->stream_of(stream)>>>(sh->epoch, (uint32_t)sh->rank, (uint32_t)sh->world, p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7]);
  count_launch(ctx);
  OLAP_CUDA(cudaGetLastError());
  return OLAP_OK;
}

olap_status olap_shard_wait(olap_shard* sh, void* stream) {
  OLAP_REQUIRE(sh && sh->slabs[0], "olap_shard_attach first");
  olap_ctx* ctx = sh->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  OLAP_CUDA(cudaSetDevice(ctx->device));
  sh->epoch_waited++;
  shard_wait_kernel<<<1, 32, 0, ctx->stream_of(stream)>>>(sh->epoch_waited, (uint32_t)sh->world, sh->sync_words(sh->rank), sh->flags(0));
  count_launch(ctx);
  OLAP_CUDA(cudaGetLastError());
  return OLAP_OK;
}

// the BARRIER of the per-phase protocol on one stream: signal, then wait
olap_status olap_shard_barrier(olap_shard* sh, void* stream) {
  OLAP_TRY(olap_shard_signal(sh, stream));
  return olap_shard_wait(sh, stream);
}

olap_status olap_shard_create(olap_ctx* ctx, int32_t rank, int32_t world, uint64_t build_rows_cap, uint64_t probe_rows_cap, int32_t probe_chunks,
                              olap_shard** out) {
  OLAP_REQUIRE(ctx && out, "NULL argument");
  OLAP_REQUIRE(world >= 1 && world <= kShardMaxWorld && rank >= 0 && rank < world, "1 <= world <= 8, 0 <= rank < world");
  OLAP_REQUIRE(probe_chunks >= 1 && probe_chunks <= 64, "1 <= probe_chunks <= 64");
  auto* sh = new olap_shard();
  sh->ctx = ctx; sh->rank = rank; sh->world = world; sh->chunks = probe_chunks;
  sh->build_cap = build_rows_cap; sh->probe_cap = probe_rows_cap;
  sh->chunk_cap = (probe_rows_cap + probe_chunks - 1) / probe_chunks;
  sh->words_b = reserve_words(build_rows_cap);
  sh->words_p = reserve_words(sh->chunk_cap);
  if (sh->words_b >= 0xFFFFFFFFULL || sh->words_p * probe_chunks >= 0xFFFFFFFFULL) {
    delete sh; set_error("row positions are u32: a rank receives < 2^32 rows per side"); return OLAP_ERR_BAD_ARG;
  }
  sh->off_counts = (sh->words_b + sh->words_p * probe_chunks) * 8;
  sh->off_sync = sh->off_counts + (uint64_t)kShardRegions * (1 + probe_chunks) * 4;
  sh->slab_bytes = sh->off_sync + 64;   // one arrival word per rank (+ one error word), written by the peers
  std::lock_guard<std::mutex> g(ctx->mu);
  OLAP_CUDA(cudaSetDevice(ctx->device));
  sh->local_words = (uint64_t)kShardRegions * (1 + probe_chunks) + kFlagWords * (2 + probe_chunks) + 8;
  {
    const char* m = getenv("OLAP_SHARD_MODE");
    sh->ce_mode = !(m && (m[0] == 't' || m[0] == 'T'));
  }
  olap_status st = pool_alloc(ctx, sh->local_words * 4, ctx->streams[0], (void**)&sh->local);
  if (st == OLAP_OK && sh->ce_mode && world > 1) st = pool_alloc(ctx, (sh->words_b + sh->words_p * probe_chunks) * 8, ctx->streams[0], (void**)&sh->outbox);
  if (st == OLAP_OK) st = pool_alloc(ctx, ((uint64_t)6 * hot_table_slots() + 2 * (uint64_t)hot_sample_slots() + 2 * kHotCand + 16) * 4, ctx->streams[0], (void**)&sh->hot);
  if (st != OLAP_OK) { delete sh; return st; }
  *out = sh;
  return OLAP_OK;
}

uint64_t olap_shard_slab_bytes(const olap_shard* sh) { return sh ? sh->slab_bytes : 0; }

olap_status olap_shard_attach(olap_shard* sh, void* const* peer_slabs) {
  OLAP_REQUIRE(sh && peer_slabs, "NULL argument");
  for (int r = 0; r < sh->world; r++) { OLAP_REQUIRE(peer_slabs[r], "NULL slab pointer"); sh->slabs[r] = peer_slabs[r]; }
  // this rank's arrival words start at zero; the caller lets no rank call olap_shard_barrier before every rank has attached
  olap_ctx* ctx = sh->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  OLAP_CUDA(cudaSetDevice(ctx->device));
  OLAP_CUDA(cudaMemsetAsync(sh->sync_words(sh->rank), 0, 64, ctx->streams[0]));
  OLAP_CUDA(cudaStreamSynchronize(ctx->streams[0]));
  sh->epoch = 0; sh->epoch_waited = 0;
  return OLAP_OK;
}

olap_status olap_shard_destroy(olap_shard* sh) {
  if (!sh) return OLAP_OK;
  olap_ctx* ctx = sh->ctx;
  {
    std::lock_guard<std::mutex> g(ctx->mu);
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->streams[0]);
    if (sh->built) { dj_probe_end(sh->pg, ctx->streams[0]); dj_build_free(sh->dj, ctx->streams[0]); }
    if (sh->local) pool_free(ctx, sh->local, ctx->streams[0]);
    if (sh->outbox) pool_free(ctx, sh->outbox, ctx->streams[0]);
    if (sh->ce_fork) { cudaEventDestroy(sh->ce_fork); for (auto e : sh->ce_join) if (e) cudaEventDestroy(e); }
    if (sh->hot) pool_free(ctx, sh->hot, ctx->streams[0]);
    if (sh->own_slab_alloc) cudaFree(sh->own_slab_alloc);
  }
  delete sh;
  return OLAP_OK;
}

// Geometry for one join from the GLOBAL min / max of the build keys and the global build row count (the caller reduces them over
// the ranks; every rank must pass the same values).  OLAP_ERR_UNSUPPORTED: the domain is not dense (range > 5 x rows) or too wide
// for this path -- the caller then takes the hash-sharded path.  Resets the per-join device state on `stream`.
olap_status olap_shard_plan(olap_shard* sh, int64_t global_min, int64_t global_max, uint64_t global_build_rows, int32_t hot, void* stream) {
  OLAP_REQUIRE(sh, "shard is NULL");
  olap_ctx* ctx = sh->ctx;
  sh->planned = false;
  const uint64_t span = (uint64_t)global_max - (uint64_t)global_min;
  if (global_max < global_min || span >= 0xFFFFFFF0ULL || span + 1 > 5 * global_build_rows) {
    set_error("range sharding needs a dense build-key domain (range <= 5 x rows, < 2^32)"); return OLAP_ERR_UNSUPPORTED;
  }
  const uint64_t grange = span + 1, W = (uint64_t)sh->world;
  ShardGeometry geo;
  OLAP_TRY(shard_geometry(sh->world, grange, &geo));
  const uint64_t Q = geo.Q;
  sh->Q = geo.Q; sh->F = geo.F; sh->subw = geo.subw; sh->slice = geo.slice;
  sh->gmin = global_min; sh->grange = grange;
  const uint64_t regions = W * Q;
  {
    // region capacity: slack % above an even share (every byte of a region's capacity crosses NVLink in ce mode), within the reserve
    const char* v = getenv("OLAP_SHARD_SLACK_PCT");
    // 5 %: a uniform chunk region holds ~2^19 rows +- 0.14 % (one sigma); what the slack buys is tolerance for mild non-uniformity, what
    // it costs is NVLink time -- at 8 GPUs the exchange, not the SMs, bounds the step (7/8 of every row leaves its rank)
    const uint64_t slack = v && *v ? strtoull(v, nullptr, 10) : (hot ? 25 : 5);
    auto cap_of = [&](uint64_t rows, uint64_t words) {
      // a region takes the source's rows of ONE sub-range: rows x (sub-range width / domain width) under uniform keys -- not
      // rows / regions: the slices are whole multiples of 2^15 keys, so on a small domain the first ranks own more than 1 / W of it
      const uint64_t mean = (uint64_t)((double)rows * (double)sh->subw / (double)grange) + 1;
      const uint64_t want = (mean + mean * slack / 100 + 2048 + 3) & ~3ULL;
      return (uint32_t)std::min<uint64_t>(std::min<uint64_t>(want, (words / regions) & ~3ULL), 0xFFFFFFF0ULL);
    };
    sh->bcap = cap_of(sh->build_cap, sh->words_b);
    sh->pcap = cap_of(sh->chunk_cap, sh->words_p);
  }
  sh->hot_on = hot == 1; sh->skew_room = hot != 0; sh->hot_ready = false; sh->hot_pairs = nullptr; sh->hot_max = 0;
  std::lock_guard<std::mutex> g(ctx->mu);
  OLAP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream_of(stream);
  OLAP_CUDA(cudaMemsetAsync(sh->local, 0, sh->local_words * 4, s));
  if (sh->built) { dj_probe_end(sh->pg, s); dj_build_free(sh->dj, s); sh->built = false; }
  {
    // chunks 0 .. K - last - 1 are joined in groups of `group`, the last `last` chunks form the final group
    const char* v = getenv("OLAP_SHARD_JOIN_GROUPS");
    const char* w = getenv("OLAP_SHARD_LAST_GROUP");
    const int K = sh->chunks;
    const int groups = std::max(1, std::min(K, v && *v ? atoi(v) : 2));
    sh->last_group = groups == 1 ? K : std::max(1, std::min(K - 1, w && *w ? atoi(w) : (K + 3) / 4));
    const int head = K - sh->last_group;
    sh->group = groups == 1 ? K : std::max(1, (head + groups - 2) / (groups - 1));
  }
  phase_begin(ctx, s);
  sh->planned = true;
  return OLAP_OK;
}

// side 0: the whole build side (chunk ignored); side 1: probe chunk `chunk`.  row_base = local row id of keys[0].
// After this call returns the stream holds: shuffle kernel, count publication.  The receivers may read once EVERY rank's call
// for this (side, chunk) has completed -- the caller provides that barrier.
olap_status olap_shard_scatter(olap_shard* sh, int32_t side, int32_t chunk, const int64_t* keys, uint64_t n_rows, uint32_t row_base, void* stream) {
  OLAP_REQUIRE(sh && sh->planned, "olap_shard_plan first");
  OLAP_REQUIRE(side == 0 || (side == 1 && chunk >= 0 && chunk < sh->chunks), "side 0 / 1, chunk < probe_chunks");
  OLAP_REQUIRE(n_rows == 0 || keys, "keys is NULL");
  OLAP_REQUIRE(n_rows <= (side ? sh->chunk_cap : sh->build_cap), "more rows than the shard was created for");
  olap_ctx* ctx = sh->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  OLAP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream_of(stream);
  const uint32_t cap = side ? sh->pcap : sh->bcap;
  uint64_t* regions[kShardMaxWorld];
  for (int r = 0; r < sh->world; r++) {
    if (sh->ce_mode && r != sh->rank) {
      // the outbox block of destination r holds what will be regions [rank * Q, rank * Q + Q) over there: bias the base so that the
      // kernel's region index (rank * Q + sub) lands on outbox region (r * Q + sub)
      uint64_t* ob = side ? sh->outbox_p(chunk) : sh->outbox_b();
      regions[r] = ob + ((int64_t)r - (int64_t)sh->rank) * (int64_t)sh->Q * (int64_t)cap;
    } else regions[r] = side ? sh->regions_p(r, chunk) : sh->regions_b(r);
  }
  uint32_t* cur = sh->cursor(side ? 1 + chunk : 0);
  HotRef hr;
  const bool use_hot = side && sh->hot_on;
  if (use_hot) {
    OLAP_REQUIRE(sh->hot_ready, "hot = 1: olap_shard_hot_set / olap_shard_hot_resolve / olap_shard_hot_finalize come before the probe scatter");
    hr.hot_keys = sh->hot_keys(); hr.hot_ids = sh->hot_ids();
    hr.pairs = sh->hot_pairs; hr.max_matches = sh->hot_pairs ? sh->hot_max : 0; hr.count = sh->hot_count();
  }
  if (n_rows)
    OLAP_TRY(range_shuffle(ctx, (const uint64_t*)keys, n_rows, row_base, sh->gmin, sh->grange, sh->Q, sh->subw, sh->rank, sh->world, regions, cap, cur,
                           sh->flags(0), side == 0, use_hot ? &hr : nullptr, sh->ce_mode ? 0u : shuffle_grid(ctx), !sh->ce_mode, s));
  return OLAP_OK;
}

// Second half of a scatter, on a stream the caller has ordered behind it: ce mode -- one peer copy per destination rank (copy
// engines: the outbox block becomes regions [rank * Q, rank * Q + Q) of that rank's slab); both modes -- this rank's row counts
// written into the owners' count arrays.  Then comes the BARRIER.
olap_status olap_shard_send(olap_shard* sh, int32_t side, int32_t chunk, void* stream) {
  OLAP_REQUIRE(sh && sh->planned, "olap_shard_plan first");
  OLAP_REQUIRE(side == 0 || (side == 1 && chunk >= 0 && chunk < sh->chunks), "side 0 / 1, chunk < probe_chunks");
  olap_ctx* ctx = sh->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  OLAP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream_of(stream);
  const uint32_t cap = side ? sh->pcap : sh->bcap;
  if (sh->ce_mode && sh->world > 1) {
    // one copy per peer, spread over the context's spare streams so that several copy engines work at once (a single stream's
    // copies run one after the other: 447 GB/s per direction at 4 GPUs); s waits for all of them before the counts go out
    const uint64_t* ob = side ? sh->outbox_p(chunk) : sh->outbox_b();
    const uint64_t block = (uint64_t)sh->Q * cap;
    const size_t ns = ctx->streams.size();
    const int lanes = ns > 3 ? (int)std::min<size_t>(ns - 3, (size_t)sh->world - 1) : 0;
    if (!sh->ce_fork) { OLAP_CUDA(cudaEventCreateWithFlags(&sh->ce_fork, cudaEventDisableTiming)); for (auto& e : sh->ce_join) OLAP_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); }
    if (lanes) OLAP_CUDA(cudaEventRecord(sh->ce_fork, s));
    for (int l = 0; l < lanes; l++) OLAP_CUDA(cudaStreamWaitEvent(ctx->streams[3 + l], sh->ce_fork, 0));
    for (int k = 1; k < sh->world; k++) {
      const int r = (sh->rank + k) % sh->world;   // every rank starts with a different peer
      uint64_t* dst = (side ? sh->regions_p(r, chunk) : sh->regions_b(r)) + (uint64_t)sh->rank * block;
      cudaStream_t cs = lanes ? ctx->streams[3 + (k - 1) % lanes] : s;
      OLAP_CUDA(cudaMemcpyAsync(dst, ob + (uint64_t)r * block, block * 8, cudaMemcpyDeviceToDevice, cs));
    }
    for (int l = 0; l < lanes; l++) {
      OLAP_CUDA(cudaEventRecord(sh->ce_join[l], ctx->streams[3 + l]));
      OLAP_CUDA(cudaStreamWaitEvent(s, sh->ce_join[l], 0));
    }
  }
  uint32_t* counts[kShardMaxWorld];
  for (int r = 0; r < sh->world; r++) counts[r] = side ? sh->counts_p(r, chunk) : sh->counts_b(r);
  OLAP_TRY(range_shuffle_publish(ctx, sh->cursor(side ? 1 + chunk : 0), cap, sh->Q, sh->rank, sh->world, counts, s));
  return OLAP_OK;
}

// the received build rows -> key ranges of 2^15 keys (second partition level); after the barrier behind olap_shard_scatter(side 0)
olap_status olap_shard_build(olap_shard* sh, void* stream) {
  OLAP_REQUIRE(sh && sh->planned, "olap_shard_plan first");
  olap_ctx* ctx = sh->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  OLAP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream_of(stream);
  if (sh->built) { dj_probe_end(sh->pg, s); dj_build_free(sh->dj, s); sh->built = false; }
  RInput in;
  in.regions = sh->regions_b(sh->rank); in.reg_cur = sh->counts_b(sh->rank); in.reg_cap = sh->bcap; in.n_reg = sh->world * sh->Q;
  in.reg_q = sh->Q; in.reg_parts = sh->F;
  OLAP_TRY(dj_build_prepare(ctx, in, 0, sh->slice, local_grid(ctx, sh->ce_mode), s, &sh->dj));
  sh->built = true;
  OLAP_CUDA(cudaMemcpyAsync(sh->flags(1), sh->dj.bctl + sh->dj.PP, kFlagWords * 4, cudaMemcpyDeviceToDevice, s));   // the build side's flags
  // rows this rank receives per chunk: its slice's share of the domain of every source's chunk; with heavy hitters pulled out the
  // warm keys still load single key ranges unevenly, so the ranges get three times the room (hot = 2: 200 % slack)
  const double share = std::min(1.0, (double)sh->slice / (double)sh->grange);
  const uint64_t per_chunk = (uint64_t)((double)sh->chunk_cap * sh->world * share) + 1;
  return dj_probe_begin(sh->dj, per_chunk * (uint64_t)std::max(sh->group, sh->last_group), sh->skew_room ? 2 : 0, s, &sh->pg);
}

// join probe chunk `chunk` against the shard's build side; pairs are APPENDED behind those of the earlier chunks of this join
// (one running device counter per join).  pair.left_row / right_row index this rank's receive slabs (build / probe side);
// olap_shard_pairs_to_global turns them into global row ids.
olap_status olap_shard_probe(olap_shard* sh, int32_t chunk, olap_pair* out_pairs, uint64_t max_matches, void* stream) {
  OLAP_REQUIRE(sh && sh->planned && sh->built, "olap_shard_plan / olap_shard_build first");
  OLAP_REQUIRE(chunk >= 0 && chunk < sh->chunks, "chunk < probe_chunks");
  olap_ctx* ctx = sh->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  OLAP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream_of(stream);
  RInput in;
  in.regions = sh->regions_p(sh->rank, chunk); in.reg_cur = sh->counts_p(sh->rank, chunk); in.reg_cap = sh->pcap; in.n_reg = sh->world * sh->Q;
  in.reg_q = sh->Q; in.reg_parts = sh->F; in.id_base = (uint32_t)((uint64_t)chunk * sh->words_p);
  if (!out_pairs) max_matches = 0;
  // the chunk's rows join the current group's regions; the shared-memory tables are built and probed once per group
  OLAP_TRY(dj_probe_add(sh->pg, in, out_pairs, max_matches, sh->pair_count(), local_grid(ctx, sh->ce_mode), s, sh->flags(2 + chunk)));
  const int head = sh->chunks - sh->last_group;
  if (chunk + 1 == sh->chunks || chunk + 1 == head || (chunk < head && (chunk + 1) % sh->group == 0))
    OLAP_TRY(dj_probe_join(sh->pg, out_pairs, max_matches, sh->pair_count(), local_grid(ctx, sh->ce_mode), s, sh->flags(2 + chunk)));
  return OLAP_OK;
}

// synchronises `stream`; count = pairs found so far in this join (exact even past max_matches); flags: OR over the shuffle and
// every chunk (kFlag*: [0] a region overflowed -- skew, [1] a build key outside the planned range, [2] a repeated build key).
// Any flag set on ANY rank (the caller ORs them across ranks) means the result must be discarded and the join redone on the
// hash-sharded path.
olap_status olap_shard_finish(olap_shard* sh, uint64_t* host_count, uint64_t* host_hot_count, uint32_t* host_flags, void* stream) {
  OLAP_REQUIRE(sh && host_count && host_flags, "NULL argument");
  olap_ctx* ctx = sh->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  OLAP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream_of(stream);
  std::vector<uint32_t> h(kFlagWords * (2 + sh->chunks) + 4);
  OLAP_CUDA(cudaMemcpyAsync(h.data(), sh->flags(0), h.size() * 4, cudaMemcpyDeviceToHost, s));
  OLAP_CUDA(cudaStreamSynchronize(s));
  for (int f = 0; f < kFlagWords; f++) host_flags[f] = 0;
  for (int b = 0; b < 2 + sh->chunks; b++)
    for (int f = 0; f < kFlagWords; f++) host_flags[f] |= h[kFlagWords * b + f] ? 1u : 0u;
  memcpy(host_count, h.data() + kFlagWords * (2 + sh->chunks), 8);
  if (host_hot_count) memcpy(host_hot_count, h.data() + kFlagWords * (2 + sh->chunks) + 2, 8);
  return OLAP_OK;
}

olap_status olap_shard_pairs_to_global(olap_shard* sh, const olap_pair* pairs, uint64_t n_pairs, const uint64_t* build_row_offsets_dptr,
                                       const uint64_t* probe_row_offsets_dptr, uint64_t* out_left, uint64_t* out_right, void* stream) {
  OLAP_REQUIRE(sh && sh->planned, "olap_shard_plan first");
  if (!n_pairs) return OLAP_OK;
  olap_ctx* ctx = sh->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  OLAP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream_of(stream);
  shard_pairs_to_global_kernel<<<(unsigned)ctx->sm_count * 8, 256, 0, s>>>(pairs, n_pairs, sh->regions_b(sh->rank), sh->bcap, sh->regions_p(sh->rank, 0),
                                                                             sh->words_p, sh->pcap, sh->Q, build_row_offsets_dptr, probe_row_offsets_dptr,
                                                                             out_left, out_right);
  count_launch(ctx);
  OLAP_CUDA(cudaGetLastError());
  return OLAP_OK;
}

// layout of the receive slab and geometry of the planned join, for callers that inspect the slabs (bench / tests):
// out[0] words reserved for the build regions, [1] words per probe chunk, [2] build region capacity, [3] probe region capacity,
// [4] Q (sub-ranges per rank), [5] F (key ranges of 2^15 keys per sub-range), [6] byte offset of the count arrays, [7] chunks
olap_status olap_shard_layout(const olap_shard* sh, uint64_t* out8) {
  OLAP_REQUIRE(sh && out8, "NULL argument");
  out8[0] = sh->words_b; out8[1] = sh->words_p; out8[2] = sh->bcap; out8[3] = sh->pcap; out8[4] = sh->Q; out8[5] = sh->F;
  out8[6] = sh->off_counts; out8[7] = (uint64_t)sh->chunks;
  return OLAP_OK;
}

// rows this rank stored into every rank's slab during the current join (both sides, all chunks so far); synchronises `stream`
olap_status olap_shard_sent_rows(olap_shard* sh, uint64_t* host_rows_per_dest /* [world] */, void* stream) {
  OLAP_REQUIRE(sh && sh->planned && host_rows_per_dest, "olap_shard_plan first");
  olap_ctx* ctx = sh->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  OLAP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream_of(stream);
  std::vector<uint32_t> h((size_t)kShardRegions * (1 + sh->chunks));
  OLAP_CUDA(cudaMemcpyAsync(h.data(), sh->local, h.size() * 4, cudaMemcpyDeviceToHost, s));
  OLAP_CUDA(cudaStreamSynchronize(s));
  for (int r = 0; r < sh->world; r++) host_rows_per_dest[r] = 0;
  for (int b = 0; b < 1 + sh->chunks; b++) {
    const uint32_t cap = b ? sh->pcap : sh->bcap;
    for (uint32_t j = 0; j < (uint32_t)sh->world * sh->Q; j++) host_rows_per_dest[j / sh->Q] += std::min(h[(size_t)kShardRegions * b + j], cap);
  }
  return OLAP_OK;
}

// ---- heavy hitters (Zipf-skewed probe keys).  With olap_shard_plan(hot = 1) the order of a join is:
//   olap_shard_hot_candidates (every rank, its probe shard)  ->  the host language gathers the lists, sums the counts per key,
//   keeps the <= 2048 most frequent  ->  olap_shard_hot_set (same keys on every rank)
//   scatter + BARRIER + olap_shard_build (build side, as before)
//   olap_shard_hot_resolve -> element-wise MAX of the 8192 x u64 table over the ranks -> olap_shard_hot_finalize
//   probe scatters: rows carrying a hot key stay on their rank; pair {hot slot, local probe row} goes to hot_pairs, the slot's
//   build row is (merged[slot] - 1) = source rank << 32 | source row.
// keys are passed as (key - global_min + 1) in u32, the form the tables hold.
olap_status olap_shard_hot_candidates(olap_shard* sh, const int64_t* probe_keys, uint64_t n_rows, uint32_t* host_keys_plus1, uint32_t* host_counts,
                                      uint32_t capacity, uint32_t* host_n, void* stream) {
  OLAP_REQUIRE(sh && sh->planned && host_keys_plus1 && host_counts && host_n, "olap_shard_plan first; non-NULL outputs");
  olap_ctx* ctx = sh->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  OLAP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream_of(stream);
  uint32_t *dk, *dc, *dn;
  OLAP_TRY(hot_candidates(ctx, (const uint64_t*)probe_keys, n_rows, sh->gmin, sh->grange, sh->hot_scratch(), kHotCand, &dk, &dc, &dn, s));
  uint32_t n = 0;
  OLAP_CUDA(cudaMemcpyAsync(&n, dn, 4, cudaMemcpyDeviceToHost, s));
  OLAP_CUDA(cudaStreamSynchronize(s));
  n = std::min(n, std::min(capacity, kHotCand));
  if (n) {
    OLAP_CUDA(cudaMemcpyAsync(host_keys_plus1, dk, n * 4, cudaMemcpyDeviceToHost, s));
    OLAP_CUDA(cudaMemcpyAsync(host_counts, dc, n * 4, cudaMemcpyDeviceToHost, s));
    OLAP_CUDA(cudaStreamSynchronize(s));
  }
  *host_n = n;
  return OLAP_OK;
}

olap_status olap_shard_hot_set(olap_shard* sh, const uint32_t* keys_plus1, uint32_t n_keys, void* stream) {
  OLAP_REQUIRE(sh && sh->planned && (n_keys == 0 || keys_plus1), "olap_shard_plan first");
  OLAP_REQUIRE(n_keys <= hot_max_keys(), "at most 2048 hot keys (a quarter of the table's slots)");
  // the table is built here, sequentially over the SORTED keys: every rank ends up with the same slot for the same key
  std::vector<uint32_t> keys(keys_plus1, keys_plus1 + n_keys), tab(hot_table_slots(), 0);
  std::sort(keys.begin(), keys.end());
  keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
  hot_table_build_host(keys.data(), (uint32_t)keys.size(), tab.data());
  olap_ctx* ctx = sh->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  OLAP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream_of(stream);
  OLAP_CUDA(cudaMemcpyAsync(sh->hot_keys(), tab.data(), tab.size() * 4, cudaMemcpyHostToDevice, s));
  OLAP_CUDA(cudaStreamSynchronize(s));  // tab is a local
  return OLAP_OK;
}

// after olap_shard_build.  out_dptr: 8192 x u64 on the device, to be MAX-reduced over the ranks by the caller
olap_status olap_shard_hot_resolve(olap_shard* sh, uint64_t* out_dptr, void* stream) {
  OLAP_REQUIRE(sh && sh->planned && sh->built && out_dptr, "olap_shard_build first");
  olap_ctx* ctx = sh->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  OLAP_CUDA(cudaSetDevice(ctx->device));
  return hot_resolve_global(ctx, sh->hot_keys(), out_dptr, sh->dj, sh->regions_b(sh->rank), sh->bcap, sh->Q, sh->slice, sh->rank, ctx->stream_of(stream));
}

// merged_dptr: the reduced table; hot_pairs (nullable): where this rank's {slot, probe row} pairs go
olap_status olap_shard_hot_finalize(olap_shard* sh, const uint64_t* merged_dptr, olap_pair* hot_pairs, uint64_t max_hot_pairs, void* stream) {
  OLAP_REQUIRE(sh && sh->planned && merged_dptr, "olap_shard_plan first");
  olap_ctx* ctx = sh->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  OLAP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream_of(stream);
  OLAP_CUDA(cudaMemcpyAsync(sh->hot_glob(), merged_dptr, (size_t)hot_table_slots() * 8, cudaMemcpyDeviceToDevice, s));
  OLAP_TRY(hot_finalize(ctx, sh->hot_keys(), sh->hot_ids(), sh->hot_glob(), s));
  sh->hot_pairs = hot_pairs; sh->hot_max = max_hot_pairs; sh->hot_ready = true;
  return OLAP_OK;
}

// the merged hot table of the current join: host_merged[slot] = 0, or (source rank << 32 | source row) + 1 of the slot's build row
olap_status olap_shard_hot_table(olap_shard* sh, uint64_t* host_merged /* [8192] */, void* stream) {
  OLAP_REQUIRE(sh && host_merged, "NULL argument");
  olap_ctx* ctx = sh->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  OLAP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream_of(stream);
  OLAP_CUDA(cudaMemcpyAsync(host_merged, sh->hot_glob(), (size_t)hot_table_slots() * 8, cudaMemcpyDeviceToHost, s));
  OLAP_CUDA(cudaStreamSynchronize(s));
  return OLAP_OK;
}

// this rank's min / max of a key column (remembered per column like the 1-GPU join's); the caller reduces them over the ranks
olap_status olap_shard_key_range(olap_shard* sh, const int64_t* keys, uint64_t n_rows, int64_t* out_min, int64_t* out_max, void* stream) {
  OLAP_REQUIRE(sh && out_min && out_max, "NULL argument");
  olap_ctx* ctx = sh->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  OLAP_CUDA(cudaSetDevice(ctx->device));
  if (n_rows == 0) { *out_min = INT64_MAX; *out_max = INT64_MIN; return OLAP_OK; }
  return key_min_max(ctx, (const uint64_t*)keys, n_rows, ctx->stream_of(stream), out_min, out_max);
}

// Host-only: the heavy-hitter table as olap_shard_hot_set builds it (sorted, de-duplicated keys_plus1 = key - global_min + 1; 2048
// buckets of 4 slots, second-choice bucket when the first is full) into out_table[8192], and the slot the kernels' lookup rule finds
// for each probe_plus1[i] (0xFFFFFFFF: not a heavy hitter).  *out_placed = keys that found a slot (a key whose two buckets are both
// full is simply not a heavy hitter).  No GPU: the CPU suite checks builder and lookup against each other.
olap_status olap_shard_hot_table_host(const uint32_t* keys_plus1, uint32_t n_keys, uint32_t* out_table, uint32_t* out_placed,
                                      const uint32_t* probe_plus1, uint32_t n_probe, uint32_t* out_slots) {
  OLAP_REQUIRE(out_table && (n_keys == 0 || keys_plus1) && (n_probe == 0 || (probe_plus1 && out_slots)), "NULL argument");
  OLAP_REQUIRE(n_keys <= hot_max_keys(), "at most 2048 hot keys (a quarter of the table's slots)");
  std::vector<uint32_t> keys(keys_plus1, keys_plus1 + n_keys);
  std::sort(keys.begin(), keys.end());
  keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
  memset(out_table, 0, (size_t)hot_table_slots() * 4);
  const uint32_t placed = hot_table_build_host(keys.data(), (uint32_t)keys.size(), out_table);
  if (out_placed) *out_placed = placed;
  for (uint32_t i = 0; i < n_probe; i++) out_slots[i] = (probe_plus1[i] == 0u || probe_plus1[i] == 0xFFFFFFFFu) ? 0xFFFFFFFFu : hot_lookup_host(out_table, probe_plus1[i]);
  return OLAP_OK;
}

// Host-only: where the shuffle sends a key.  For every keys[i] (host array) in the planned domain [global_min, global_max]: the owning
// rank, the sub-range (region index modulo Q) inside that rank's slab side and the offset the row carries there (key - slice start);
// keys outside the domain get rank -1 (the shuffle drops them: they cannot match).  Runs the SAME routing arithmetic as the kernels
// (route_part / route_dest, djoin.cuh: exact division by multiplication) -- no GPU needed, which is how the CPU suite checks that
// arithmetic against plain integer division.  out_geometry (nullable): {Q, F, sub-range width, slice width}.
olap_status olap_shard_route_i64(int32_t world, int64_t global_min, int64_t global_max, const int64_t* keys, uint64_t n, int32_t* out_rank,
                                 uint32_t* out_sub_range, uint32_t* out_offset, uint64_t* out_geometry) {
  OLAP_REQUIRE(world >= 1 && world <= kShardMaxWorld, "1 <= world <= 8");
  OLAP_REQUIRE(n == 0 || (keys && out_rank && out_sub_range && out_offset), "NULL argument");
  const uint64_t span = (uint64_t)global_max - (uint64_t)global_min;
  OLAP_REQUIRE(global_max >= global_min && span < 0xFFFFFFF0ULL, "the key domain must be non-empty and narrower than 2^32");
  ShardGeometry g;
  OLAP_TRY(shard_geometry(world, span + 1, &g));
  if (out_geometry) { out_geometry[0] = g.Q; out_geometry[1] = g.F; out_geometry[2] = g.subw; out_geometry[3] = g.slice; }
  const uint64_t mf = route_mul_f(g.F);
  const uint32_t mq = route_mul_q(g.Q);
  for (uint64_t i = 0; i < n; i++) {
    const uint64_t off = (uint64_t)keys[i] - (uint64_t)global_min;
    if (off > span) { out_rank[i] = -1; out_sub_range[i] = 0; out_offset[i] = 0; continue; }
    const uint32_t d = (uint32_t)off, part = route_part(d, mf), dest = route_dest(part, mq);
    out_rank[i] = (int32_t)dest;
    out_sub_range[i] = part - dest * g.Q;
    out_offset[i] = d - dest * (g.Q * g.subw);
  }
  return OLAP_OK;
}

// ---------------------------------------------------------------------------------------------------- one process, all ranks
// n ranks driven by one host thread: contexts on `devices[r]` (a device may repeat: all ranks on one GPU is how the exchange
// logic is tested on a 1-GPU box), slabs allocated here, peer access enabled.  Wraps what SURVEY 8(b) calls shard_ctx_create_all.
olap_status olap_shard_ctx_create_all(int32_t n_ranks, const int32_t* devices, uint64_t max_gpu_memory, uint64_t build_rows_cap,
                                      uint64_t probe_rows_cap, int32_t probe_chunks, olap_ctx** out_ctxs, olap_shard** out_shards) {
  OLAP_REQUIRE(n_ranks >= 1 && n_ranks <= kShardMaxWorld && devices && out_ctxs && out_shards, "1 <= n_ranks <= 8, non-NULL arrays");
  for (int r = 0; r < n_ranks; r++) { out_ctxs[r] = nullptr; out_shards[r] = nullptr; }
  olap_status st = OLAP_OK;
  for (int r = 0; r < n_ranks && st == OLAP_OK; r++) {
    st = olap_ctx_create(devices[r], max_gpu_memory, 8, &out_ctxs[r]);
    if (st == OLAP_OK) st = olap_shard_create(out_ctxs[r], r, n_ranks, build_rows_cap, probe_rows_cap, probe_chunks, &out_shards[r]);
  }
  void* slabs[kShardMaxWorld] = {};
  for (int r = 0; r < n_ranks && st == OLAP_OK; r++) {
    cudaSetDevice(devices[r]);
    for (int q = 0; q < n_ranks; q++) {
      if (devices[q] == devices[r]) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, devices[r], devices[q]);
      if (!can) { set_error("device %d cannot access device %d", devices[r], devices[q]); st = OLAP_ERR_UNSUPPORTED; break; }
      cudaError_t e = cudaDeviceEnablePeerAccess(devices[q], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { st = cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__); break; }
      cudaGetLastError();
    }
    if (st != OLAP_OK) break;
    cudaError_t e = cudaMalloc(&slabs[r], out_shards[r]->slab_bytes);
    if (e != cudaSuccess) { st = cuda_fail(e, "cudaMalloc(slab)", __FILE__, __LINE__); break; }
    out_shards[r]->own_slab_alloc = slabs[r];
  }
  for (int r = 0; r < n_ranks && st == OLAP_OK; r++) st = olap_shard_attach(out_shards[r], slabs);
  if (st != OLAP_OK) {
    for (int r = 0; r < n_ranks; r++) {
      if (out_shards[r]) { olap_shard_destroy(out_shards[r]); out_shards[r] = nullptr; }
      if (out_ctxs[r]) { olap_ctx_destroy(out_ctxs[r]); out_ctxs[r] = nullptr; }
    }
  }
  return st;
}

// The whole sharded join from one host thread.  Rank r holds build_keys[r][0 .. n_build[r]) and probe_keys[r][0 .. n_probe[r]) on
// its device; its share of the result lands in out_pairs[r] (positions in rank r's receive slabs: olap_shard_pairs_to_global).
// host_counts[r] = pairs of rank r.  OLAP_ERR_UNSUPPORTED: not a dense key domain, or a flag was raised (skew, repeated key).
olap_status olap_sharded_hash_join_i64(olap_shard* const* shards, int32_t n_ranks, const int64_t* const* build_keys, const uint64_t* n_build,
                                       const int64_t* const* probe_keys, const uint64_t* n_probe, olap_pair* const* out_pairs,
                                       const uint64_t* max_matches, uint64_t* host_counts, olap_pair* const* hot_pairs, const uint64_t* max_hot_pairs,
                                       uint64_t* host_hot_counts) {
  OLAP_REQUIRE(shards && n_ranks >= 1 && n_ranks <= kShardMaxWorld && build_keys && n_build && probe_keys && n_probe && out_pairs && max_matches && host_counts,
               "NULL argument");
  const int W = n_ranks;
  int64_t gmin = INT64_MAX, gmax = INT64_MIN;
  uint64_t total_build = 0;
  for (int r = 0; r < W; r++) {
    int64_t mn, mx;
    OLAP_TRY(olap_shard_key_range(shards[r], build_keys[r], n_build[r], &mn, &mx, nullptr));
    gmin = std::min(gmin, mn); gmax = std::max(gmax, mx);
    total_build += n_build[r];
  }
  if (total_build == 0) { for (int r = 0; r < W; r++) host_counts[r] = 0; return OLAP_OK; }
  const bool hot = hot_pairs != nullptr && max_hot_pairs != nullptr && host_hot_counts != nullptr;
  std::vector<cudaStream_t> comm(W), comp(W);
  for (int r = 0; r < W; r++) {
    olap_ctx* c = shards[r]->ctx;
    comm[r] = c->streams[1 % c->streams.size()]; comp[r] = c->streams[0];
    OLAP_TRY(olap_shard_plan(shards[r], gmin, gmax, total_build, hot ? 1 : 0, comm[r]));
  }
  if (hot) {  // the ranks' candidate lists, merged here (what the host language does between the calls in the multi-process flow)
    std::unordered_map<uint32_t, uint64_t> freq;
    std::vector<uint32_t> k(kHotCand), c(kHotCand);
    for (int r = 0; r < W; r++) {
      uint32_t n = 0;
      OLAP_TRY(olap_shard_hot_candidates(shards[r], probe_keys[r], n_probe[r], k.data(), c.data(), kHotCand, &n, comm[r]));
      for (uint32_t i = 0; i < n; i++) freq[k[i]] += c[i];
    }
    std::vector<std::pair<uint64_t, uint32_t>> byc;
    for (auto& kv : freq) byc.emplace_back(kv.second, kv.first);
    std::sort(byc.begin(), byc.end(), [](const std::pair<uint64_t, uint32_t>& a, const std::pair<uint64_t, uint32_t>& b) { return a.first != b.first ? a.first > b.first : a.second < b.second; });
    std::vector<uint32_t> keys;
    for (size_t i = 0; i < byc.size() && keys.size() < hot_max_keys(); i++) keys.push_back(byc[i].second);
    for (int r = 0; r < W; r++) OLAP_TRY(olap_shard_hot_set(shards[r], keys.data(), (uint32_t)keys.size(), comm[r]));
  }
  const int K = shards[0]->chunks;
  // per rank: part = partition kernels of the scatters, comm = copies / count publication (what the peers wait for), comp = local work
  std::vector<cudaStream_t> part(W);
  for (int r = 0; r < W; r++) { olap_ctx* c = shards[r]->ctx; part[r] = shards[r]->ce_mode ? c->streams[2 % c->streams.size()] : comm[r]; }
  std::vector<cudaEvent_t> ev((size_t)W * (1 + K)), pev((size_t)W * (1 + K));
  auto cleanup = [&]() { for (auto e : ev) if (e) cudaEventDestroy(e); for (auto e : pev) if (e) cudaEventDestroy(e); };
  for (auto& e : ev) e = nullptr;
  for (auto& e : pev) e = nullptr;
  olap_status st = OLAP_OK;
  // scatter on `part`, then (behind an event) copies + counts on `comm`, then the event the peers wait for
  auto scatter_send = [&](int r, int phase, int side, int chunk, const int64_t* keys, uint64_t n, uint32_t row_base) -> olap_status {
    cudaSetDevice(shards[r]->ctx->device);
    if (phase == 0 && part[r] != comm[r]) {  // the reset of the cursors (olap_shard_plan, on comm) precedes the first partition kernel
      cudaEvent_t e0;
      OLAP_CUDA(cudaEventCreateWithFlags(&e0, cudaEventDisableTiming));
      OLAP_CUDA(cudaEventRecord(e0, comm[r]));
      OLAP_CUDA(cudaStreamWaitEvent(part[r], e0, 0));
      OLAP_CUDA(cudaEventDestroy(e0));
    }
    OLAP_TRY(olap_shard_scatter(shards[r], side, chunk, keys, n, row_base, part[r]));
    if (part[r] != comm[r]) {
      cudaEvent_t& pe = pev[(size_t)phase * W + r];
      OLAP_CUDA(cudaEventCreateWithFlags(&pe, cudaEventDisableTiming));
      OLAP_CUDA(cudaEventRecord(pe, part[r]));
      OLAP_CUDA(cudaStreamWaitEvent(comm[r], pe, 0));
    }
    OLAP_TRY(olap_shard_send(shards[r], side, chunk, comm[r]));
    cudaEvent_t& e = ev[(size_t)phase * W + r];
    OLAP_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    OLAP_CUDA(cudaEventRecord(e, comm[r]));
    return OLAP_OK;
  };
  auto wait_all = [&](int r, int phase) -> olap_status {
    cudaSetDevice(shards[r]->ctx->device);
    for (int q = 0; q < W; q++) OLAP_CUDA(cudaStreamWaitEvent(comp[r], ev[(size_t)phase * W + q], 0));
    return OLAP_OK;
  };
  for (int r = 0; r < W && st == OLAP_OK; r++) st = scatter_send(r, 0, 0, 0, build_keys[r], n_build[r], 0);
  for (int r = 0; r < W && st == OLAP_OK; r++) {
    st = wait_all(r, 0);
    if (st == OLAP_OK) st = olap_shard_build(shards[r], comp[r]);
  }
  if (hot && st == OLAP_OK) {  // the hot keys' build rows: resolved by their owners, MAX-merged, handed back to every rank
    const uint32_t S = hot_table_slots();
    std::vector<uint64_t> merged(S, 0), partial(S);
    for (int r = 0; r < W && st == OLAP_OK; r++) {
      st = olap_shard_hot_resolve(shards[r], shards[r]->hot_glob(), comp[r]);
      if (st != OLAP_OK) break;
      cudaSetDevice(shards[r]->ctx->device);
      if (cudaMemcpyAsync(partial.data(), shards[r]->hot_glob(), S * 8, cudaMemcpyDeviceToHost, comp[r]) != cudaSuccess || cudaStreamSynchronize(comp[r]) != cudaSuccess) {
        st = cuda_fail(cudaGetLastError(), "hot table download", __FILE__, __LINE__); break;
      }
      for (uint32_t i = 0; i < S; i++) merged[i] = std::max(merged[i], partial[i]);
    }
    for (int r = 0; r < W && st == OLAP_OK; r++) {
      cudaSetDevice(shards[r]->ctx->device);
      uint64_t* tmp = (uint64_t*)shards[r]->hot_scratch();   // the sampling scratch is free again
      if (cudaMemcpyAsync(tmp, merged.data(), S * 8, cudaMemcpyHostToDevice, part[r]) != cudaSuccess || cudaStreamSynchronize(part[r]) != cudaSuccess) {
        st = cuda_fail(cudaGetLastError(), "hot table upload", __FILE__, __LINE__); break;
      }
      st = olap_shard_hot_finalize(shards[r], tmp, hot_pairs[r], max_hot_pairs[r], part[r]);
    }
  }
  for (int k = 0; k < K && st == OLAP_OK; k++) {
    for (int r = 0; r < W && st == OLAP_OK; r++) {
      const uint64_t cc = shards[r]->chunk_cap, lo = std::min<uint64_t>((uint64_t)k * cc, n_probe[r]), hi = std::min<uint64_t>(lo + cc, n_probe[r]);
      st = scatter_send(r, 1 + k, 1, k, probe_keys[r] + lo, hi - lo, (uint32_t)lo);
    }
    for (int r = 0; r < W && st == OLAP_OK; r++) {
      st = wait_all(r, 1 + k);
      if (st == OLAP_OK) st = olap_shard_probe(shards[r], k, out_pairs[r], max_matches[r], comp[r]);
    }
  }
  bool bad = false;
  for (int r = 0; r < W; r++) {
    uint32_t hf[kFlagWords] = {0, 0, 0, 0};
    cudaSetDevice(shards[r]->ctx->device);
    cudaStreamSynchronize(part[r]);
    cudaStreamSynchronize(comm[r]);
    olap_status s2 = olap_shard_finish(shards[r], &host_counts[r], hot ? &host_hot_counts[r] : nullptr, hf, comp[r]);
    if (st == OLAP_OK) st = s2;
    for (int f = 0; f < kFlagWords; f++) bad = bad || hf[f];
  }
  cleanup();
  OLAP_TRY(st);
  if (bad) { set_error("range-sharded join: a region overflowed (skewed keys) or a build key repeats; use the hash-sharded path"); return OLAP_ERR_UNSUPPORTED; }
  return OLAP_OK;
}

}  // extern "C"
