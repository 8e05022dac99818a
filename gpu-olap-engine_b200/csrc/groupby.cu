// Hash group-by aggregation (SUM / COUNT / MIN / MAX / AVG over int64 or float64 values, int64 group key), with an
// optional fused filter conjunction.  Replaces execute_aggregate (gpu-olap-core/src/executor.rs:297-355),
// AggregateKernel::aggregate (gpu-kernel-compiler/src/aggregate_kernel.rs:72-125) and hash_aggregate_kernel
// (kernels/join_kernels.cuh:216-240: direct addressing, SUM only, one global atomic per row).
//
// Regimes (chosen from n_groups_hint; a wrong hint costs speed, never correctness):
//   no key      per-thread register accumulators -> warp shuffle -> one global atomic per CTA per cell
//   shared      per-CTA open-addressing table in shared memory, replicated R times for tiny cardinalities so
//               that smem-atomic contention stays low; rows whose key does not fit fall through to the global
//               table; tables are merged into the global table at the end
//   global      global open-addressing table (L2-resident up to ~1M groups), native 64-bit L2 atomics,
//               read-before-atomic for MIN/MAX
//   partitioned rows are radix-partitioned by key hash (join.cu's partitioner) so that every partition's groups
//               fit the shared regime -- see group_by_partitioned below
// Roofline: HBM; algorithmic bytes per row = 8 B key + 8 B per distinct value column (+ 8 B per predicate column).
#include "common.cuh"

#include <cfloat>

namespace olap {

constexpr int kMaxCells = 8;
constexpr uint64_t kEmptyKey = 0x8000000000000000ULL;  // INT64_MIN is the empty marker; the real key lives in slot `cap`

enum CellOp : int8_t { ADD_I64 = 0, ADD_F64 = 1, MIN_I64 = 2, MAX_I64 = 3, MIN_F64 = 4, MAX_F64 = 5 };

struct GbParams {
  const uint64_t* keys;  // NULL: single global group
  uint64_t n;
  TermPack terms;
  int n_cells;
  const uint64_t* cell_col[kMaxCells];
  int8_t cell_op[kMaxCells];
  // global table: SoA, cap + 1 entries each (the extra one holds key == INT64_MIN)
  uint64_t* g_keys;
  unsigned long long* g_count;
  uint64_t* g_cells;  // cell c at g_cells + c * (g_cap + 1)
  uint64_t g_cap;     // power of two
  uint32_t* g_overflow;
  // shared regime
  uint32_t s_slots;   // power of two, per copy
  uint32_t s_copies;  // power of two
};

// order-preserving f64 <-> i64 (involution); lets MIN/MAX of doubles use integer atomics
__host__ __device__ __forceinline__ int64_t f64_ord(uint64_t bits) {
  int64_t b = (int64_t)bits;
  return b >= 0 ? b : (int64_t)(bits ^ 0x7FFFFFFFFFFFFFFFULL);
}

__host__ __device__ __forceinline__ uint64_t cell_identity(int op) {
  union { double d; uint64_t u; } mx, mn;
  mx.d = DBL_MAX; mn.d = -DBL_MAX;
  switch (op) {
    case MIN_I64: return (uint64_t)INT64_MAX;
    case MAX_I64: return (uint64_t)INT64_MIN;
    case MIN_F64: return (uint64_t)f64_ord(mx.u);
    case MAX_F64: return (uint64_t)f64_ord(mn.u);
    default: return 0;  // +0.0 and 0
  }
}

// value as stored in a cell (f64 min/max use the ordered encoding); valid=false for NaN in min/max
__device__ __forceinline__ uint64_t cell_encode(int op, uint64_t raw, bool& valid) {
  valid = true;
  if (op == MIN_F64 || op == MAX_F64) {
    double d = __longlong_as_double((long long)raw);
    valid = d == d;
    return (uint64_t)f64_ord(raw);
  }
  return raw;
}

__device__ __forceinline__ uint64_t cell_combine(int op, uint64_t a, uint64_t b) {
  switch (op) {
    case ADD_I64: return a + b;
    case ADD_F64: return (uint64_t)__double_as_longlong(__longlong_as_double((long long)a) + __longlong_as_double((long long)b));
    case MIN_I64: case MIN_F64: return (uint64_t)min((long long)a, (long long)b);
    default: return (uint64_t)max((long long)a, (long long)b);
  }
}

__device__ __forceinline__ uint64_t ld_cg_u64(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// atomic update of a global cell
__device__ __forceinline__ void cell_atomic_global(int op, uint64_t* cell, uint64_t v) {
  switch (op) {
    case ADD_I64: atomicAdd((unsigned long long*)cell, (unsigned long long)v); break;
    case ADD_F64: atomicAdd((double*)cell, __longlong_as_double((long long)v)); break;
    case MIN_I64: case MIN_F64:
      if ((long long)v < (long long)ld_cg_u64(cell)) atomicMin((long long*)cell, (long long)v);
      break;
    default:
      if ((long long)v > (long long)ld_cg_u64(cell)) atomicMax((long long*)cell, (long long)v);
  }
}
__device__ __forceinline__ void cell_atomic_shared(int op, uint64_t* cell, uint64_t v) {
  switch (op) {
    case ADD_I64: atomicAdd((unsigned long long*)cell, (unsigned long long)v); break;
    case ADD_F64: atomicAdd((double*)cell, __longlong_as_double((long long)v)); break;
    case MIN_I64: case MIN_F64:
      if ((long long)v < *(volatile long long*)cell) atomicMin((long long*)cell, (long long)v);
      break;
    default:
      if ((long long)v > *(volatile long long*)cell) atomicMax((long long*)cell, (long long)v);
  }
}

// find-or-insert in the global table; returns the slot, or ~0 when the table is full (overflow flag raised)
__device__ __forceinline__ uint64_t global_slot(const GbParams& p, uint64_t key) {
  if (key == kEmptyKey) return p.g_cap;
  const uint64_t mask = p.g_cap - 1;
  uint64_t s = mix64(key) & mask;
  for (uint64_t probes = 0; probes <= mask; probes++) {
    uint64_t k = ld_cg_u64(p.g_keys + s);
    if (k == key) return s;
    if (k == kEmptyKey) {
      uint64_t old = atomicCAS((unsigned long long*)(p.g_keys + s), (unsigned long long)kEmptyKey, (unsigned long long)key);
      if (old == kEmptyKey || old == key) return s;
    }
    s = (s + 1) & mask;
  }
  atomicExch(p.g_overflow, 1u);
  return ~0ULL;
}

__device__ __forceinline__ void global_update_row(const GbParams& p, uint64_t key, const uint64_t* vals, const bool* valid) {
  uint64_t s = global_slot(p, key);
  if (s == ~0ULL) return;
  atomicAdd(p.g_count + s, 1ULL);
#pragma unroll 1
  for (int c = 0; c < p.n_cells; c++)
    if (valid[c]) cell_atomic_global(p.cell_op[c], p.g_cells + (uint64_t)c * (p.g_cap + 1) + s, vals[c]);
}

constexpr int kGbThreads = 512;
constexpr int kGbRows = 4;  // rows per thread per iteration (independent loads in flight)

// ------------------------------------------------------------------------------------------- no-key regime
__global__ void __launch_bounds__(kGbThreads) gb_nokey_kernel(GbParams p) {
  uint64_t acc[kMaxCells];
  unsigned long long cnt = 0;
#pragma unroll
  for (int c = 0; c < kMaxCells; c++) acc[c] = c < p.n_cells ? cell_identity(p.cell_op[c]) : 0;
  const uint64_t stride = (uint64_t)gridDim.x * kGbThreads;
  for (uint64_t row = blockIdx.x * (uint64_t)kGbThreads + threadIdx.x; row < p.n; row += stride) {
    if (p.terms.n && !eval_terms_row(p.terms, row)) continue;
    cnt++;
#pragma unroll
    for (int c = 0; c < kMaxCells; c++) {
      if (c < p.n_cells) {
        bool valid;
        uint64_t v = cell_encode(p.cell_op[c], ld_stream_u64(p.cell_col[c] + row), valid);
        if (valid) acc[c] = cell_combine(p.cell_op[c], acc[c], v);
      }
    }
  }
  // warp reduce, then one atomic per warp (f64 sums: the order varies run to run; see DESIGN.md on f64 parity)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
#pragma unroll
    for (int c = 0; c < kMaxCells; c++)
      if (c < p.n_cells) acc[c] = cell_combine(p.cell_op[c], acc[c], __shfl_xor_sync(0xffffffffu, acc[c], o));
  }
  __shared__ uint64_t s_acc[kGbThreads / 32][kMaxCells + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    s_acc[warp][kMaxCells] = cnt;
    for (int c = 0; c < p.n_cells; c++) s_acc[warp][c] = acc[c];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long tc = 0;
    for (int w = 0; w < kGbThreads / 32; w++) tc += s_acc[w][kMaxCells];
    if (tc) atomicAdd(p.g_count, tc);
    for (int c = 0; c < p.n_cells; c++) {
      uint64_t a = s_acc[0][c];
      for (int w = 1; w < kGbThreads / 32; w++) a = cell_combine(p.cell_op[c], a, s_acc[w][c]);
      if (tc) cell_atomic_global(p.cell_op[c], p.g_cells + (uint64_t)c * (p.g_cap + 1), a);
    }
  }
}

// ------------------------------------------------------------------------------------------- global regime
__global__ void __launch_bounds__(kGbThreads) gb_global_kernel(GbParams p) {
  const uint64_t tile = (uint64_t)kGbThreads * kGbRows;
  for (uint64_t base = blockIdx.x * tile; base < p.n; base += (uint64_t)gridDim.x * tile) {
    uint64_t key[kGbRows];
    bool pass[kGbRows];
#pragma unroll
    for (int j = 0; j < kGbRows; j++) {
      uint64_t row = base + (uint64_t)j * kGbThreads + threadIdx.x;
      pass[j] = row < p.n;
      key[j] = pass[j] ? ld_stream_u64(p.keys + row) : 0;
    }
#pragma unroll
    for (int j = 0; j < kGbRows; j++) {
      uint64_t row = base + (uint64_t)j * kGbThreads + threadIdx.x;
      if (pass[j] && p.terms.n) pass[j] = eval_terms_row(p.terms, row);
      if (!pass[j]) continue;
      uint64_t vals[kMaxCells];
      bool valid[kMaxCells];
#pragma unroll
      for (int c = 0; c < kMaxCells; c++)
        if (c < p.n_cells) vals[c] = cell_encode(p.cell_op[c], ld_stream_u64(p.cell_col[c] + row), valid[c]);
      global_update_row(p, key[j], vals, valid);
    }
  }
}

// ------------------------------------------------------------------------------------------- shared regime
// Dynamic smem layout per copy: keys[s_slots + 1] | count[s_slots + 1] (u32) | cells[n_cells][s_slots + 1]
__global__ void __launch_bounds__(kGbThreads) gb_shared_kernel(GbParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t slots1 = p.s_slots + 1, total = slots1 * p.s_copies;
  uint64_t* s_keys = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* s_cells = s_keys + total;                                        // [c][copy][slot]
  uint32_t* s_count = reinterpret_cast<uint32_t*>(s_cells + (uint64_t)p.n_cells * total);
  for (uint32_t i = threadIdx.x; i < total; i += kGbThreads) {
    s_keys[i] = kEmptyKey;
    s_count[i] = 0;
    for (int c = 0; c < p.n_cells; c++) s_cells[(uint64_t)c * total + i] = cell_identity(p.cell_op[c]);
  }
  __syncthreads();
  const uint32_t copy_base = (threadIdx.x & (p.s_copies - 1)) * slots1;
  const uint32_t smask = p.s_slots - 1;
  // leave a quarter of each copy free so that probe sequences stay short
  const uint32_t max_probe = p.s_slots;

  const uint64_t tile = (uint64_t)kGbThreads * kGbRows;
  for (uint64_t base = blockIdx.x * tile; base < p.n; base += (uint64_t)gridDim.x * tile) {
    uint64_t key[kGbRows];
    bool pass[kGbRows];
#pragma unroll
    for (int j = 0; j < kGbRows; j++) {
      uint64_t row = base + (uint64_t)j * kGbThreads + threadIdx.x;
      pass[j] = row < p.n;
      key[j] = pass[j] ? ld_stream_u64(p.keys + row) : 0;
    }
#pragma unroll
    for (int j = 0; j < kGbRows; j++) {
      uint64_t row = base + (uint64_t)j * kGbThreads + threadIdx.x;
      if (pass[j] && p.terms.n) pass[j] = eval_terms_row(p.terms, row);
      if (!pass[j]) continue;
      uint64_t vals[kMaxCells];
      bool valid[kMaxCells];
#pragma unroll
      for (int c = 0; c < kMaxCells; c++)
        if (c < p.n_cells) vals[c] = cell_encode(p.cell_op[c], ld_stream_u64(p.cell_col[c] + row), valid[c]);
      // find-or-insert in this thread's copy
      uint32_t slot = 0xFFFFFFFFu;
      if (key[j] == kEmptyKey) slot = copy_base + p.s_slots;
      else {
        uint32_t s = (uint32_t)(mix64(key[j]) >> 32) & smask;
        for (uint32_t probes = 0; probes < max_probe; probes++) {
          uint64_t k = *(volatile uint64_t*)(s_keys + copy_base + s);
          if (k == key[j]) { slot = copy_base + s; break; }
          if (k == kEmptyKey) {
            uint64_t old = atomicCAS((unsigned long long*)(s_keys + copy_base + s), (unsigned long long)kEmptyKey, (unsigned long long)key[j]);
            if (old == kEmptyKey || old == key[j]) { slot = copy_base + s; break; }
          }
          s = (s + 1) & smask;
        }
      }
      if (slot == 0xFFFFFFFFu) { global_update_row(p, key[j], vals, valid); continue; }  // copy is full: spill to the global table
      atomicAdd(s_count + slot, 1u);
#pragma unroll 1
      for (int c = 0; c < p.n_cells; c++)
        if (valid[c]) cell_atomic_shared(p.cell_op[c], s_cells + (uint64_t)c * total + slot, vals[c]);
    }
  }
  __syncthreads();
  // merge every copy into the global table
  for (uint32_t i = threadIdx.x; i < total; i += kGbThreads) {
    const uint32_t cnt = s_count[i];
    if (!cnt) continue;
    const uint64_t key = (i % slots1 == p.s_slots) ? kEmptyKey : s_keys[i];
    uint64_t s = global_slot(p, key);
    if (s == ~0ULL) continue;
    atomicAdd(p.g_count + s, (unsigned long long)cnt);
    for (int c = 0; c < p.n_cells; c++) cell_atomic_global(p.cell_op[c], p.g_cells + (uint64_t)c * (p.g_cap + 1) + s, s_cells[(uint64_t)c * total + i]);
  }
}

// ------------------------------------------------------------------------------------------- table init / extraction
__global__ void __launch_bounds__(256) gb_init_kernel(GbParams p) {
  const uint64_t n = p.g_cap + 1, stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    p.g_keys[i] = kEmptyKey;
    p.g_count[i] = 0;
    for (int c = 0; c < p.n_cells; c++) p.g_cells[(uint64_t)c * n + i] = cell_identity(p.cell_op[c]);
  }
}

struct OutSpec {
  int n_aggs;
  int8_t func[16];
  int8_t type[16];
  int8_t cell[16];  // cell index feeding this aggregate (-1: count only)
};

__global__ void __launch_bounds__(256) gb_extract_kernel(GbParams p, OutSpec o, bool emit_empty_global, int64_t* __restrict__ out_keys,
                                                         uint64_t* __restrict__ out_aggs, uint64_t out_cap,
                                                         unsigned long long* __restrict__ out_n) {
  const uint64_t n = p.g_cap + 1, stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i0 = blockIdx.x * (uint64_t)blockDim.x; i0 < n; i0 += stride) {
    const uint64_t i = i0 + threadIdx.x;
    unsigned long long cnt = i < n ? p.g_count[i] : 0;
    bool occ = cnt != 0 || (emit_empty_global && i == 0);
    unsigned m = __ballot_sync(0xffffffffu, occ);
    if (!m) continue;
    const int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(out_n, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (!occ) continue;
    const uint64_t g = base + __popc(m & ((1u << lane) - 1));
    if (g >= out_cap) continue;
    out_keys[g] = p.keys ? (i == p.g_cap ? INT64_MIN : (int64_t)p.g_keys[i]) : 0;
    for (int a = 0; a < o.n_aggs; a++) {
      uint64_t r;
      const uint64_t cell = o.cell[a] >= 0 ? p.g_cells[(uint64_t)o.cell[a] * n + i] : 0;
      if (o.func[a] == OLAP_COUNT) r = cnt;
      else if (o.func[a] == OLAP_AVG) {
        double s = o.type[a] == OLAP_F64 ? __longlong_as_double((long long)cell) : (double)(long long)cell;
        r = (uint64_t)__double_as_longlong(cnt ? s / (double)cnt : 0.0);
      } else if ((o.func[a] == OLAP_MIN || o.func[a] == OLAP_MAX) && o.type[a] == OLAP_F64) r = (uint64_t)f64_ord(cell);
      else r = cell;
      out_aggs[(uint64_t)a * out_cap + g] = r;
    }
  }
}

// ------------------------------------------------------------------------------------------- host driver
static olap_status build_cells(const olap_agg_spec* specs, int n_aggs, GbParams* p, OutSpec* o) {
  OLAP_REQUIRE(n_aggs >= 0 && n_aggs <= 16, "n_aggs must be in [0, 16]");
  p->n_cells = 0;
  o->n_aggs = n_aggs;
  for (int a = 0; a < n_aggs; a++) {
    const int f = specs[a].func, t = specs[a].value_type;
    OLAP_REQUIRE(f >= OLAP_SUM && f <= OLAP_AVG, "unknown aggregate function");
    OLAP_REQUIRE(t == OLAP_I64 || t == OLAP_F64, "value_type must be OLAP_I64 or OLAP_F64");
    o->func[a] = (int8_t)f; o->type[a] = (int8_t)t; o->cell[a] = -1;
    if (f == OLAP_COUNT) continue;
    OLAP_REQUIRE(specs[a].values, "aggregate value column is NULL");
    int8_t op = (f == OLAP_SUM || f == OLAP_AVG) ? (t == OLAP_F64 ? ADD_F64 : ADD_I64)
                : f == OLAP_MIN ? (t == OLAP_F64 ? MIN_F64 : MIN_I64) : (t == OLAP_F64 ? MAX_F64 : MAX_I64);
    int found = -1;
    for (int c = 0; c < p->n_cells; c++)
      if (p->cell_op[c] == op && p->cell_col[c] == (const uint64_t*)specs[a].values) found = c;
    if (found < 0) {
      OLAP_REQUIRE(p->n_cells < kMaxCells, "more than 8 distinct (function, column) accumulators");
      found = p->n_cells++;
      p->cell_op[found] = op;
      p->cell_col[found] = (const uint64_t*)specs[a].values;
    }
    o->cell[a] = (int8_t)found;
  }
  return OLAP_OK;
}

static uint64_t pow2_at_least(uint64_t x) { uint64_t p = 1; while (p < x) p <<= 1; return p; }

// defined in join.cu: partitioned group-by support
olap_status group_by_partitioned(olap_ctx* ctx, GbParams& p, const OutSpec& o, uint64_t n_groups_hint, int64_t* out_keys,
                                 uint64_t* out_aggs, uint64_t out_cap, unsigned long long* d_n, cudaStream_t s, bool* handled);

}  // namespace olap

using namespace olap;

extern "C" olap_status olap_group_by_i64(olap_ctx* ctx, const int64_t* keys, const olap_agg_spec* specs, int32_t n_aggs, uint64_t n_rows,
                                         uint64_t n_groups_hint, const olap_filter_term* terms, int32_t n_terms, int64_t* out_keys,
                                         void* out_aggs, uint64_t out_capacity, uint64_t* out_n_groups_dptr, uint64_t* host_n_groups,
                                         void* stream) {
  OLAP_REQUIRE(ctx, "ctx is NULL");
  OLAP_REQUIRE(n_aggs == 0 || specs, "specs is NULL");
  OLAP_REQUIRE(out_capacity == 0 || (out_keys && (n_aggs == 0 || out_aggs)), "output pointer is NULL");
  std::lock_guard<std::mutex> g(ctx->mu);
  OLAP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream_of(stream);

  GbParams p;
  memset(&p, 0, sizeof(p));
  OutSpec o;
  memset(&o, 0, sizeof(o));
  OLAP_TRY(build_cells(specs, n_aggs, &p, &o));
  OLAP_TRY(pack_terms(terms, n_terms, &p.terms));
  p.keys = (const uint64_t*)keys;
  p.n = n_rows;
  if (!keys) n_groups_hint = 1;
  if (n_groups_hint < 1) n_groups_hint = 1;

  Scratch sc(ctx, s);
  unsigned long long* d_n = nullptr;
  uint32_t* d_overflow = nullptr;
  OLAP_TRY(sc.get<unsigned long long>(1, &d_n));
  OLAP_TRY(sc.get<uint32_t>(1, &d_overflow));
  phase_begin(ctx, s);

  uint64_t n_groups = 0;
  for (int attempt = 0;; attempt++) {
    Scratch tsc(ctx, s);
    p.g_cap = keys ? pow2_at_least(2 * n_groups_hint < 1024 ? 1024 : 2 * n_groups_hint) : 1;
    const uint64_t n1 = p.g_cap + 1;
    OLAP_TRY(tsc.get<uint64_t>(n1, &p.g_keys));
    OLAP_TRY(tsc.get<unsigned long long>(n1, &p.g_count));
    OLAP_TRY(tsc.get<uint64_t>(n1 * (p.n_cells ? p.n_cells : 1), &p.g_cells));
    p.g_overflow = d_overflow;
    OLAP_CUDA(cudaMemsetAsync(d_n, 0, 8, s));
    OLAP_CUDA(cudaMemsetAsync(d_overflow, 0, 4, s));
    const unsigned init_grid = (unsigned)(((n1 + 255) / 256) < (uint64_t)ctx->sm_count * 8 ? ((n1 + 255) / 256) : (uint64_t)ctx->sm_count * 8);
    gb_init_kernel<<<init_grid, 256, 0, s>>>(p);
    count_launch(ctx);
    phase_mark(ctx, s, "init");

    bool handled = false;
    if (n_rows > 0) {
      const uint64_t tiles = (n_rows + (uint64_t)kGbThreads * kGbRows - 1) / ((uint64_t)kGbThreads * kGbRows);
      if (!keys) {
        unsigned grid = (unsigned)(tiles < (uint64_t)ctx->sm_count * 4 ? tiles : (uint64_t)ctx->sm_count * 4);
        gb_nokey_kernel<<<grid, kGbThreads, 0, s>>>(p);
        count_launch(ctx);
        handled = true;
      } else {
        // shared regime if a table with <= 75 % load fits the per-CTA budget
        const uint64_t slot_bytes = 8 + 4 + 8ULL * p.n_cells;
        const uint64_t budget = 96 * 1024;  // two CTAs per SM
        uint64_t need = pow2_at_least((n_groups_hint * 4 + 2) / 3);
        if (need < 16) need = 16;
        if ((need + 1) * slot_bytes <= budget && attempt == 0) {
          uint64_t copies = 1;
          while (copies < 32 && (need + 1) * slot_bytes * copies * 2 <= budget) copies <<= 1;
          p.s_slots = (uint32_t)need;
          p.s_copies = (uint32_t)copies;
          const size_t smem = (size_t)((need + 1) * copies * slot_bytes + 16);
          OLAP_CUDA(cudaFuncSetAttribute(gb_shared_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          unsigned grid = (unsigned)(tiles < (uint64_t)ctx->sm_count * 2 ? tiles : (uint64_t)ctx->sm_count * 2);
          gb_shared_kernel<<<grid, kGbThreads, smem, s>>>(p);
          count_launch(ctx);
          handled = true;
        } else if (attempt == 0) {
          OLAP_TRY(group_by_partitioned(ctx, p, o, n_groups_hint, out_keys, (uint64_t*)out_aggs, out_capacity, d_n, s, &handled));
        }
        if (!handled) {
          unsigned grid = (unsigned)(tiles < (uint64_t)ctx->sm_count * 4 ? tiles : (uint64_t)ctx->sm_count * 4);
          gb_global_kernel<<<grid, kGbThreads, 0, s>>>(p);
          count_launch(ctx);
        }
      }
      OLAP_CUDA(cudaGetLastError());
    }
    phase_mark(ctx, s, "aggregate");
    const uint64_t n1b = p.g_cap + 1;
    const unsigned ex_grid = (unsigned)(((n1b + 255) / 256) < (uint64_t)ctx->sm_count * 8 ? ((n1b + 255) / 256) : (uint64_t)ctx->sm_count * 8);
    gb_extract_kernel<<<ex_grid, 256, 0, s>>>(p, o, /*emit_empty_global=*/keys == nullptr, out_keys, (uint64_t*)out_aggs, out_capacity, d_n);
    count_launch(ctx);
    OLAP_CUDA(cudaGetLastError());
    phase_mark(ctx, s, "extract");
    uint32_t h_overflow = 0;
    OLAP_CUDA(cudaMemcpyAsync(&h_overflow, d_overflow, 4, cudaMemcpyDeviceToHost, s));
    OLAP_CUDA(cudaMemcpyAsync(&n_groups, d_n, 8, cudaMemcpyDeviceToHost, s));
    OLAP_CUDA(cudaStreamSynchronize(s));
    if (!h_overflow) break;
    // the caller's hint was too small for the global table: grow and redo (correctness never depends on the hint)
    n_groups_hint = n_groups_hint * 8 > n_rows ? (n_rows ? n_rows : 1) : n_groups_hint * 8;
    OLAP_REQUIRE(attempt < 24, "group-by table kept overflowing");
  }
  phase_end(ctx, s);
  if (out_n_groups_dptr) OLAP_CUDA(cudaMemcpyAsync(out_n_groups_dptr, d_n, 8, cudaMemcpyDeviceToDevice, s));
  if (host_n_groups) *host_n_groups = n_groups;
  if (n_groups > out_capacity) {
    set_error("group-by produced %llu groups but out_capacity is %llu", (unsigned long long)n_groups, (unsigned long long)out_capacity);
    return OLAP_ERR_CAPACITY;
  }
  return OLAP_OK;
}
