// Key-sharded GROUP BY across the GPUs of one box, one process driving all ranks (SURVEY 8(b) "multi-GPU": sharded_group_by_i64;
// 8(e): local pre-aggregation -> all-to-all of (key, partials) by key radix -> per-shard merge).  The reference has no multi-GPU
// path (README.md:414) and its aggregate ignores GROUP BY (gpu-olap-core/src/executor.rs:297-355); the aggregate semantics are
// AggregateSpec / AggFunc's (gpu-kernel-compiler/src/aggregate_kernel.rs:17-58): AVG travels as (sum, count), :22-23.
//
//   1. every rank aggregates its own rows into PARTIALS with the 1-GPU kernels (groupby.cu): SUM / MIN / MAX as they are, every
//      COUNT and AVG shares one row count, AVG adds a SUM of its column -- so a rank ships at most G x (n_aggs + 1) values whatever
//      its row count;
//   2. the partial rows are partitioned by owner rank = the top bits of mix64(key ^ salt) (the hash-sharded join's shuffle pass,
//      join.cu) and each (source, owner) run goes to the owner with one peer copy per column (copy engines over NVLink);
//   3. the owner merges what it received with the same group-by kernel (SUM and COUNT partials add, MIN / MAX combine) and
//      finalises: AVG = sum / count (0.0 when the count is 0, executor.rs:339-345).
// Rank r ends up with exactly the groups whose key it owns; the union over the ranks is the result.  The multi-process flow
// (sharded.py, one process per GPU) runs the same three steps with an NCCL all-to-all in place of the peer copies.
#include "common.cuh"

#include <algorithm>
#include <vector>

using namespace olap;

namespace {

// out[g] = count > 0 ? sum / count : 0.0; the sum is an i64 or an f64 bit pattern
__global__ void __launch_bounds__(256) avg_finalize_kernel(const uint64_t* __restrict__ sum, const uint64_t* __restrict__ cnt, int sum_is_f64,
                                                           uint64_t n, double* __restrict__ out) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x, st = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) {
    const double s = sum_is_f64 ? __longlong_as_double((long long)sum[i]) : (double)(long long)sum[i];
    const double c = (double)(long long)cnt[i];
    out[i] = cnt[i] ? s / c : 0.0;
  }
}

struct DevMem {   // pool memory of one context, freed in stream order on destruction
  olap_ctx* ctx = nullptr; void* p = nullptr;
  DevMem() = default;
  DevMem(const DevMem&) = delete;
  DevMem& operator=(const DevMem&) = delete;
  DevMem(DevMem&& o) noexcept : ctx(o.ctx), p(o.p) { o.p = nullptr; }
  DevMem& operator=(DevMem&& o) noexcept { if (this != &o) { release(); ctx = o.ctx; p = o.p; o.p = nullptr; } return *this; }
  ~DevMem() { release(); }
  void release() { if (p) { olap_pool_free(ctx, p, nullptr); p = nullptr; } }
  olap_status get(olap_ctx* c, uint64_t bytes) { ctx = c; return olap_pool_alloc(c, bytes ? bytes : 8, nullptr, &p); }
  template <class T> T* as() const { return (T*)p; }
};

// What crosses from one rank to another lives in plain cudaMalloc memory: a stream-ordered pool's allocations are private to their
// device unless cudaMemPoolSetAccess says otherwise, and these buffers are small and short-lived.
struct RawMem {
  int dev = 0; void* p = nullptr;
  RawMem() = default;
  RawMem(const RawMem&) = delete;
  RawMem& operator=(const RawMem&) = delete;
  ~RawMem() { if (p) { cudaSetDevice(dev); cudaFree(p); } }   // cudaFree waits for the device: nothing still reads it
  olap_status get(int device, uint64_t bytes) {
    dev = device;
    OLAP_CUDA(cudaSetDevice(device));
    OLAP_CUDA(cudaMalloc(&p, bytes ? bytes : 8));
    return OLAP_OK;
  }
  template <class T> T* as() const { return (T*)p; }
};

// device-to-device copy on the source's stream; across devices it is a peer copy (copy engines over NVLink)
static cudaError_t copy_to_rank(void* dst, int dst_dev, const void* src, int src_dev, size_t bytes, cudaStream_t s) {
  if (dst_dev == src_dev) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s);
  return cudaMemcpyPeerAsync(dst, dst_dev, src, src_dev, bytes, s);
}

struct Partial { int func; int type; int src; };   // src: index of the user aggregate whose column it reads (-1: the row count)

}  // namespace

extern "C" olap_status olap_sharded_group_by_i64(olap_ctx* const* ctxs, int32_t n_ranks, const int64_t* const* keys, const olap_agg_spec* const* specs,
                                                 int32_t n_aggs, const uint64_t* n_rows, uint64_t n_groups_hint, int64_t* const* out_keys,
                                                 void* const* out_aggs, const uint64_t* out_capacity, uint64_t* host_n_groups) {
  OLAP_REQUIRE(ctxs && keys && specs && n_rows && out_keys && out_aggs && out_capacity && host_n_groups, "NULL argument");
  OLAP_REQUIRE(n_ranks == 1 || n_ranks == 2 || n_ranks == 4 || n_ranks == 8, "n_ranks must be 1, 2, 4 or 8 (owner rank = top bits of the key hash)");
  OLAP_REQUIRE(n_aggs >= 1 && n_aggs <= 15, "1 <= n_aggs <= 15 (the partials add one shared row count)");
  const int W = n_ranks;
  int bits = 0;
  while ((1 << bits) < W) bits++;
  for (int r = 0; r < W; r++) {
    OLAP_REQUIRE(ctxs[r] && specs[r], "NULL context / spec array");
    OLAP_REQUIRE(n_rows[r] == 0 || keys[r], "keys is NULL (a sharded aggregate is grouped; use olap_group_by_i64 + a reduce for a global one)");
    OLAP_REQUIRE(n_rows[r] < 0xFFFFFFFFULL, "row ids are u32");
    for (int i = 0; i < n_aggs; i++) {
      OLAP_REQUIRE(specs[r][i].func == specs[0][i].func && specs[r][i].value_type == specs[0][i].value_type, "every rank passes the same aggregates");
      OLAP_REQUIRE(specs[r][i].func >= OLAP_SUM && specs[r][i].func <= OLAP_AVG, "unknown aggregate function");
      OLAP_REQUIRE(specs[r][i].func == OLAP_COUNT || n_rows[r] == 0 || specs[r][i].values, "values is NULL");
    }
    host_n_groups[r] = 0;
  }
  // ---- the partials: what each user aggregate reads after the merge
  std::vector<Partial> parts;
  std::vector<int> part_of(n_aggs, -1), count_of(n_aggs, -1);
  int count_part = -1;
  auto shared_count = [&]() { if (count_part < 0) { count_part = (int)parts.size(); parts.push_back({OLAP_COUNT, OLAP_I64, -1}); } return count_part; };
  for (int i = 0; i < n_aggs; i++) {
    const int f = specs[0][i].func, t = specs[0][i].value_type;
    if (f == OLAP_COUNT) { part_of[i] = shared_count(); continue; }
    const int pf = f == OLAP_AVG ? OLAP_SUM : f;
    int found = -1;   // AVG(x) and SUM(x) share one sum: same function, type and column on rank 0 (ranks pass the same shape)
    for (size_t q = 0; q < parts.size(); q++)
      if (parts[q].src >= 0 && parts[q].func == pf && parts[q].type == t && specs[0][parts[q].src].values == specs[0][i].values) { found = (int)q; break; }
    if (found < 0) { found = (int)parts.size(); parts.push_back({pf, t, i}); }
    part_of[i] = found;
    if (f == OLAP_AVG) count_of[i] = shared_count();
  }
  const int np = (int)parts.size();

  // ---- 1. local pre-aggregation
  std::vector<DevMem> pk(W), pa(W);
  std::vector<uint64_t> g(W, 0), pcap(W, 0);
  for (int r = 0; r < W; r++) {
    if (n_rows[r] == 0) continue;
    std::vector<olap_agg_spec> ps(np);
    for (int q = 0; q < np; q++) { ps[q].func = parts[q].func; ps[q].value_type = parts[q].type; ps[q].values = parts[q].src >= 0 ? specs[r][parts[q].src].values : nullptr; }
    uint64_t cap = std::max<uint64_t>(std::min<uint64_t>(n_groups_hint ? n_groups_hint : 1, n_rows[r]), 1);
    for (int attempt = 0;; attempt++) {
      pk[r] = DevMem(); pa[r] = DevMem();
      OLAP_TRY(pk[r].get(ctxs[r], cap * 8));
      OLAP_TRY(pa[r].get(ctxs[r], (uint64_t)np * cap * 8));
      const olap_status st = olap_group_by_i64(ctxs[r], keys[r], ps.data(), np, n_rows[r], n_groups_hint, nullptr, 0, pk[r].as<int64_t>(), pa[r].p, cap, nullptr,
                                               &g[r], nullptr);
      if (st == OLAP_ERR_CAPACITY && attempt == 0) { cap = g[r]; continue; }   // the count is exact: size the partials and redo
      OLAP_TRY(st);
      break;
    }
    pcap[r] = cap;
  }

  // ---- 2. route every partial row to the rank that owns its key
  // per source: rows in owner order (16-row aligned regions), their positions in the partial table, begin / count per owner
  std::vector<RawMem> sk(W);
  std::vector<DevMem> sid(W), smeta(W);
  std::vector<std::vector<uint32_t>> begin(W, std::vector<uint32_t>(W + 1, 0)), count(W, std::vector<uint32_t>(W, 0));
  if (W > 1) {
    for (int r = 0; r < W; r++) {
      if (!g[r]) continue;
      OLAP_TRY(sk[r].get(ctxs[r]->device, (g[r] + 16ULL * W) * 8));
      OLAP_TRY(sid[r].get(ctxs[r], (g[r] + 16ULL * W) * 4));
      OLAP_TRY(smeta[r].get(ctxs[r], (2ULL * W + 1) * 4));
      OLAP_TRY(olap_shuffle_partition_i64(ctxs[r], pk[r].as<int64_t>(), nullptr, g[r], bits, sk[r].as<int64_t>(), sid[r].as<uint32_t>(), smeta[r].as<uint32_t>(),
                                          smeta[r].as<uint32_t>() + W + 1, nullptr));
      std::vector<uint32_t> m(2 * W + 1);
      OLAP_TRY(olap_memcpy_d2h(ctxs[r], m.data(), smeta[r].p, m.size() * 4, nullptr));
      OLAP_TRY(olap_stream_synchronize(ctxs[r], nullptr));
      for (int d = 0; d <= W; d++) begin[r][d] = m[d];
      for (int d = 0; d < W; d++) count[r][d] = m[W + 1 + d];
    }
  } else {
    count[0][0] = (uint32_t)g[0];
  }
  // per owner: receive buffers (keys + np partial columns), filled source by source
  std::vector<uint64_t> total(W, 0);
  for (int d = 0; d < W; d++) for (int r = 0; r < W; r++) total[d] += count[r][d];
  std::vector<RawMem> rk(W), ra(W);
  for (int d = 0; d < W && W > 1; d++) {
    if (!total[d]) continue;
    OLAP_TRY(rk[d].get(ctxs[d]->device, total[d] * 8));
    OLAP_TRY(ra[d].get(ctxs[d]->device, (uint64_t)np * total[d] * 8));
  }
  if (W > 1) {
    std::vector<RawMem> send(W);   // per source: one partial column gathered into owner order, reused column by column
    for (int r = 0; r < W; r++) {
      if (!g[r]) continue;
      olap_ctx* cs = ctxs[r];
      OLAP_CUDA(cudaSetDevice(cs->device));
      cudaStream_t s = cs->streams[0];
      std::vector<uint64_t> at(W, 0);   // where source r's run starts in owner d's buffers
      for (int d = 0; d < W; d++) for (int q = 0; q < r; q++) at[d] += count[q][d];
      for (int d = 0; d < W; d++) {
        if (!count[r][d]) continue;
        OLAP_CUDA(copy_to_rank(rk[d].as<uint64_t>() + at[d], ctxs[d]->device, sk[r].as<uint64_t>() + begin[r][d], cs->device, (size_t)count[r][d] * 8, s));
      }
      OLAP_TRY(send[r].get(cs->device, (uint64_t)np * (g[r] + 16ULL * W) * 8));
      for (int q = 0; q < np; q++) {
        uint64_t* col = send[r].as<uint64_t>() + (uint64_t)q * (g[r] + 16ULL * W);
        for (int d = 0; d < W; d++) {
          if (!count[r][d]) continue;
          OLAP_TRY(olap_gather_8b(cs, pa[r].as<uint64_t>() + (uint64_t)q * pcap[r], sid[r].as<uint32_t>() + begin[r][d], count[r][d], col + begin[r][d], nullptr));
          OLAP_CUDA(cudaSetDevice(cs->device));
          OLAP_CUDA(copy_to_rank(ra[d].as<uint64_t>() + (uint64_t)q * total[d] + at[d], ctxs[d]->device, col + begin[r][d], cs->device, (size_t)count[r][d] * 8, s));
        }
      }
    }
    for (int r = 0; r < W; r++) OLAP_TRY(olap_stream_synchronize(ctxs[r], nullptr));   // every run has landed before any owner merges
  }

  // ---- 3. merge per owner, finalise into the caller's layout
  olap_status result = OLAP_OK;
  for (int d = 0; d < W; d++) {
    olap_ctx* c = ctxs[d];
    const uint64_t n_in = W > 1 ? total[d] : g[0];
    if (!n_in) continue;
    const int64_t* in_keys = W > 1 ? rk[d].as<int64_t>() : pk[0].as<int64_t>();
    const uint64_t* in_cols = W > 1 ? ra[d].as<uint64_t>() : pa[0].as<uint64_t>();
    const uint64_t in_stride = W > 1 ? total[d] : pcap[0];
    const uint64_t cap = out_capacity[d];
    DevMem mk, ma;
    uint64_t mcap = cap, mg = 0;
    if (W > 1) {
      std::vector<olap_agg_spec> ms(np);
      for (int q = 0; q < np; q++) {
        ms[q].func = parts[q].func == OLAP_COUNT ? OLAP_SUM : parts[q].func;   // counts add
        ms[q].value_type = parts[q].type; ms[q].values = in_cols + (uint64_t)q * in_stride;
      }
      for (int attempt = 0;; attempt++) {
        mk = DevMem(); ma = DevMem();
        OLAP_TRY(mk.get(c, std::max<uint64_t>(mcap, 1) * 8));
        OLAP_TRY(ma.get(c, (uint64_t)np * std::max<uint64_t>(mcap, 1) * 8));
        const olap_status st = olap_group_by_i64(c, in_keys, ms.data(), np, n_in, std::max<uint64_t>(n_groups_hint, 1), nullptr, 0, mk.as<int64_t>(), ma.p,
                                                 std::max<uint64_t>(mcap, 1), nullptr, &mg, nullptr);
        if (st == OLAP_ERR_CAPACITY && attempt == 0) { mcap = mg; continue; }   // still merge, so that the count reported is exact
        OLAP_TRY(st);
        break;
      }
      in_keys = mk.as<int64_t>(); in_cols = ma.as<uint64_t>();
    } else {
      mg = g[0]; mcap = pcap[0];
    }
    const uint64_t stride = W > 1 ? std::max<uint64_t>(mcap, 1) : pcap[0];
    host_n_groups[d] = mg;
    if (mg > cap) {
      set_error("sharded group-by: rank %d owns %llu groups but out_capacity is %llu", d, (unsigned long long)mg, (unsigned long long)cap);
      result = OLAP_ERR_CAPACITY;
      continue;
    }
    OLAP_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->streams[0];
    OLAP_CUDA(cudaMemcpyAsync(out_keys[d], in_keys, mg * 8, cudaMemcpyDeviceToDevice, s));
    for (int i = 0; i < n_aggs; i++) {
      uint64_t* dst = (uint64_t*)out_aggs[d] + (uint64_t)i * cap;
      const uint64_t* src = in_cols + (uint64_t)part_of[i] * stride;
      if (specs[0][i].func == OLAP_AVG) {
        const unsigned grid = (unsigned)std::min<uint64_t>((mg + 255) / 256, (uint64_t)c->sm_count * 8);
        avg_finalize_kernel<<<grid ? grid : 1, 256, 0, s>>>(src, in_cols + (uint64_t)count_of[i] * stride, specs[0][i].value_type == OLAP_F64, mg, (double*)dst);
        count_launch(c);
        OLAP_CUDA(cudaGetLastError());
      } else {
        OLAP_CUDA(cudaMemcpyAsync(dst, src, mg * 8, cudaMemcpyDeviceToDevice, s));
      }
    }
    OLAP_TRY(olap_stream_synchronize(c, nullptr));   // mk / ma go back to the pool when this scope ends
  }
  return result;
}
