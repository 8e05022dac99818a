// Hash group-by aggregation (SUM / COUNT / MIN / MAX / AVG over int64 or float64 values, int64 group key), with an
// optional fused filter conjunction.  Replaces execute_aggregate (gpu-olap-core/src/executor.rs:297-355),
// AggregateKernel::aggregate (gpu-kernel-compiler/src/aggregate_kernel.rs:72-125) and hash_aggregate_kernel
// (kernels/join_kernels.cuh:216-240: direct addressing, SUM only, one global atomic per row).
//
// Regimes (chosen from n_groups_hint; a wrong hint costs speed, never correctness):
//   no key       per-thread register accumulators -> warp shuffle -> one global atomic per CTA per cell
//   shared       per-CTA open-addressing table in shared memory, replicated R times for tiny cardinalities so that
//                same-address smem-atomic contention stays low; rows whose key does not fit spill to the global table;
//                tables are merged into the global table at the end
//   partitioned  rows (key + distinct value columns, filter applied) are radix-partitioned by key hash with join.cu's
//                write-combining partitioner, so that every partition's groups fit a shared-memory table (one CTA per
//                partition -- or per hash sub-range of a partition, up to 4 passes over its rows -- results written
//                straight to the output) or, for very high cardinalities, an L2-resident global table reused
//                partition after partition
//   global       one global open-addressing table (native 64-bit L2 atomics, read-before-atomic MIN/MAX): the
//                universal fallback (also the spill target of the other regimes)
// Roofline: HBM; algorithmic bytes per row = 8 B key + 8 B per distinct value column (+ 8 B per predicate column).
#include "common.cuh"

#include <algorithm>
#include <cfloat>

namespace olap {

constexpr int kMaxCells = 8;
constexpr int kMaxVcols = 6;
constexpr uint64_t kEmptyKey = 0x8000000000000000ULL;  // INT64_MIN is the empty marker; the real key lives in slot `cap`

enum CellOp : int8_t { ADD_I64 = 0, ADD_F64 = 1, MIN_I64 = 2, MAX_I64 = 3, MIN_F64 = 4, MAX_F64 = 5 };

struct OutSpec {
  int n_aggs;
  int8_t func[16];
  int8_t type[16];
  int8_t cell[16];  // cell index feeding this aggregate (-1: count only)
};

struct GbParams {
  const uint64_t* keys;  // NULL: single global group
  uint64_t n;
  TermPack terms;
  int n_vcols;                             // distinct value columns, each loaded once per row
  const uint64_t* vcol[kMaxVcols];
  ExprProg prog[kMaxVcols];                // prog[q].n > 0: value column q is computed, not loaded
  ExprProg key_prog;                       // n > 0: the group key is computed from other columns (keys is then a non-NULL sentinel, never read)
  bool prog_vec;                           // the programs' columns are 32-byte aligned
  int n_cells;
  int8_t cell_op[kMaxCells];
  int8_t cell_vcol[kMaxCells];
  // global table: SoA, cap + 1 entries each (the extra one holds key == INT64_MIN)
  uint64_t* g_keys;
  unsigned long long* g_count;
  uint64_t* g_cells;  // cell c at g_cells + c * (g_cap + 1)
  uint64_t g_cap;     // power of two
  uint32_t* g_overflow;
  bool vec;           // keys and value columns are 32-byte aligned: 256-bit loads
  bool terms_vec;     // so are the fused filter's columns
  // shared-memory tables
  uint32_t s_slots;   // power of two, per copy
  uint32_t s_copies;  // power of two
  int s_hskip;        // hash bits already used to choose the radix partition (0 in the shared regime)
  int s_hshift;       // 64 - log2(s_slots)
  int s_sub_bits;     // partitioned regime: each partition is aggregated in 2^s_sub_bits passes over disjoint hash sub-ranges
  // partitioned input (keys / vcol point at the partitioned copies)
  const uint32_t* part_begin;
  const uint32_t* part_end;
  uint32_t n_parts;
  // direct output (partitioned regime)
  OutSpec out;
  int64_t* out_keys;
  uint64_t* out_aggs;
  uint64_t out_cap;
  unsigned long long* out_n;
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
// Shared memory has no native 64-bit atomic add (ptxas emits an ATOMS.CAST.SPIN.64 loop that collapses under contention):
// an i64 sum is kept as two 32-bit halves updated with native ATOMS.ADD, the carry detected from the returned low word.
__device__ __forceinline__ void cell_atomic_shared(int op, uint64_t* cell, uint64_t v) {
  switch (op) {
    case ADD_I64: {
      uint32_t* w = reinterpret_cast<uint32_t*>(cell);
      const uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
      const uint32_t old = atomicAdd(w, lo);
      const uint32_t carry = (uint32_t)(old + lo) < lo;
      if (hi + carry) atomicAdd(w + 1, hi + carry);
      break;
    }
    case ADD_F64: atomicAdd((double*)cell, __longlong_as_double((long long)v)); break;
    case MIN_I64: case MIN_F64:
      if ((long long)v < *(volatile long long*)cell) atomicMin((long long*)cell, (long long)v);
      break;
    default:
      if ((long long)v > *(volatile long long*)cell) atomicMax((long long*)cell, (long long)v);
  }
}

// static-index pick from a small register array (a dynamic index would push the array to local memory)
template <int NV>
__device__ __forceinline__ uint64_t pick(const uint64_t (&v)[NV], int i) {
  uint64_t r = v[0];
#pragma unroll
  for (int k = 1; k < NV; k++) r = (i == k) ? v[k] : r;
  return r;
}

// find-or-insert in the global table; returns the slot, or ~0 when the table is full (overflow flag raised).
// Warp-uniform loop: EVERY lane of the warp calls this (active = false for lanes without a row) and iterates until no
// lane is pending, so the warp is converged again when it returns.  A per-lane `for (...) break` loop left the warps
// split for the rest of the kernel (ncu: 2.8 active threads per instruction, 11x the instructions).
__device__ __forceinline__ uint64_t global_slot(const GbParams& p, uint64_t key, bool active) {
  const uint64_t mask = p.g_cap - 1;
  uint64_t s = mix64(key) & mask, slot = ~0ULL, probes = 0;
  bool pending = active;
  if (pending && key == kEmptyKey) { slot = p.g_cap; pending = false; }
  while (__any_sync(0xffffffffu, pending)) {
    if (pending) {
      const uint64_t k = ld_cg_u64(p.g_keys + s);
      if (k == key) { slot = s; pending = false; }
      else {
        if (k == kEmptyKey) {
          const uint64_t old = atomicCAS((unsigned long long*)(p.g_keys + s), (unsigned long long)kEmptyKey, (unsigned long long)key);
          if (old == kEmptyKey || old == key) { slot = s; pending = false; }
        }
        if (pending) {
          s = (s + 1) & mask;
          if (++probes > mask) { atomicExch(p.g_overflow, 1u); pending = false; }
        }
      }
    }
  }
  return slot;
}

template <int NV>
__device__ __forceinline__ void global_update_row(const GbParams& p, uint64_t key, const uint64_t (&v)[NV], bool active) {
  const uint64_t s = global_slot(p, key, active);
  if (s == ~0ULL) return;
  atomicAdd(p.g_count + s, 1ULL);
#pragma unroll
  for (int c = 0; c < kMaxCells; c++)
    if (c < p.n_cells) {
      bool valid;
      const uint64_t x = cell_encode(p.cell_op[c], pick<NV>(v, p.cell_vcol[c]), valid);
      if (valid) cell_atomic_global(p.cell_op[c], p.g_cells + (uint64_t)c * (p.g_cap + 1) + s, x);
    }
}

// Four rows per thread into the global table: same structure as smem_update_rows (first probe of all rows up front, one
// shared insert / collision loop, operator switch outside the row loop).  Every lane calls it.
template <int NV>
__device__ __forceinline__ void global_update_rows(const GbParams& p, const uint64_t (&key)[4], const uint64_t (&v)[4][NV], const bool (&pass)[4]) {
  const uint64_t mask = p.g_cap - 1, kNone = ~0ULL;
  uint64_t s[4], slot[4];
  uint32_t pend = 0;
#pragma unroll
  for (int j = 0; j < 4; j++) { s[j] = mix64(key[j]) & mask; slot[j] = kNone; }
  uint64_t k0[4];
#pragma unroll
  for (int j = 0; j < 4; j++) k0[j] = pass[j] ? ld_cg_u64(p.g_keys + s[j]) : 0;  // four L2 reads in flight
#pragma unroll
  for (int j = 0; j < 4; j++)
    if (pass[j]) {
      if (key[j] == kEmptyKey) slot[j] = p.g_cap;
      else if (k0[j] == key[j]) slot[j] = s[j];
      else pend |= 1u << j;
    }
  uint64_t probes = 0;
  while (__any_sync(0xffffffffu, pend != 0)) {
#pragma unroll
    for (int j = 0; j < 4; j++)
      if (pend & (1u << j)) {
        uint64_t k = ld_cg_u64(p.g_keys + s[j]);
        if (k == kEmptyKey) k = atomicCAS((unsigned long long*)(p.g_keys + s[j]), (unsigned long long)kEmptyKey, (unsigned long long)key[j]);
        if (k == key[j] || k == kEmptyKey) { slot[j] = s[j]; pend &= ~(1u << j); }
        else s[j] = (s[j] + 1) & mask;
      }
    if (++probes > mask) { if (pend) atomicExch(p.g_overflow, 1u); break; }
  }
#pragma unroll
  for (int j = 0; j < 4; j++) if (slot[j] != kNone) atomicAdd(p.g_count + slot[j], 1ULL);
#pragma unroll 1
  for (int c = 0; c < p.n_cells; c++) {
    const int op = p.cell_op[c], q = p.cell_vcol[c];
    uint64_t* cells = p.g_cells + (uint64_t)c * (p.g_cap + 1);
    uint64_t x[4];
#pragma unroll
    for (int j = 0; j < 4; j++) x[j] = pick<NV>(v[j], q);
    switch (op) {
      case ADD_I64:
#pragma unroll
        for (int j = 0; j < 4; j++) if (slot[j] != kNone) atomicAdd((unsigned long long*)(cells + slot[j]), (unsigned long long)x[j]);
        break;
      case ADD_F64:
#pragma unroll
        for (int j = 0; j < 4; j++) if (slot[j] != kNone) atomicAdd((double*)(cells + slot[j]), __longlong_as_double((long long)x[j]));
        break;
      case MIN_I64: case MAX_I64: {
        long long cur[4];
#pragma unroll
        for (int j = 0; j < 4; j++) cur[j] = slot[j] != kNone ? (long long)ld_cg_u64(cells + slot[j]) : 0;  // read before the atomic
#pragma unroll
        for (int j = 0; j < 4; j++)
          if (slot[j] != kNone) {
            if (op == MIN_I64) { if ((long long)x[j] < cur[j]) atomicMin((long long*)(cells + slot[j]), (long long)x[j]); }
            else if ((long long)x[j] > cur[j]) atomicMax((long long*)(cells + slot[j]), (long long)x[j]);
          }
        break;
      }
      default: {  // MIN_F64 / MAX_F64 in the ordered encoding; NaN never wins
        long long cur[4];
#pragma unroll
        for (int j = 0; j < 4; j++) cur[j] = slot[j] != kNone ? (long long)ld_cg_u64(cells + slot[j]) : 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const double d = __longlong_as_double((long long)x[j]);
          const long long o = f64_ord(x[j]);
          if (slot[j] != kNone && d == d) {
            if (op == MIN_F64) { if (o < cur[j]) atomicMin((long long*)(cells + slot[j]), o); }
            else if (o > cur[j]) atomicMax((long long*)(cells + slot[j]), o);
          }
        }
      }
    }
  }
}

// finalise one group into the output (COUNT, AVG = sum / count, f64 MIN/MAX decoding); slot layout aggregate_kernel.rs:98
__device__ __forceinline__ void emit_group(const OutSpec& o, int64_t* out_keys, uint64_t* out_aggs, uint64_t out_cap, uint64_t g, int64_t key,
                                           unsigned long long cnt, const uint64_t (&cells)[kMaxCells]) {
  out_keys[g] = key;
#pragma unroll 1
  for (int a = 0; a < o.n_aggs; a++) {
    uint64_t cell = 0;
#pragma unroll
    for (int c = 0; c < kMaxCells; c++) cell = (o.cell[a] == c) ? cells[c] : cell;
    uint64_t r;
    if (o.func[a] == OLAP_COUNT) r = cnt;
    else if (o.func[a] == OLAP_AVG) {
      const double s = o.type[a] == OLAP_F64 ? __longlong_as_double((long long)cell) : (double)(long long)cell;
      r = (uint64_t)__double_as_longlong(cnt ? s / (double)cnt : 0.0);
    } else if ((o.func[a] == OLAP_MIN || o.func[a] == OLAP_MAX) && o.type[a] == OLAP_F64) r = (uint64_t)f64_ord(cell);
    else r = cell;
    out_aggs[(uint64_t)a * out_cap + g] = r;
  }
}

constexpr int kGbThreads = 512;
constexpr int kGbRows = 4;  // rows per thread per iteration: all of their loads are in flight together

// loads the keys and the distinct value columns of this thread's kGbRows CONSECUTIVE rows (row0 is a multiple of 4, so
// one 256-bit evict-first load per column when the columns are 32-byte aligned), masks rows outside [begin, end), applies
// the filter
template <int NV, bool KEYED, bool PROG = false>
__device__ __forceinline__ void load_rows(const GbParams& p, uint64_t row0, uint64_t begin, uint64_t end, uint64_t (&key)[kGbRows],
                                          uint64_t (&v)[kGbRows][NV], bool (&pass)[kGbRows]) {
  if (p.vec && row0 >= begin && row0 + 3 < end) {
    if (KEYED && !(PROG && p.key_prog.n)) { const u64x4 k = ld_stream_u64x4(p.keys + row0); key[0] = k.a; key[1] = k.b; key[2] = k.c; key[3] = k.d; }
#pragma unroll
    for (int q = 0; q < NV; q++) {
      if (q < p.n_vcols && (!PROG || p.prog[q].n == 0)) { const u64x4 x = ld_stream_u64x4(p.vcol[q] + row0); v[0][q] = x.a; v[1][q] = x.b; v[2][q] = x.c; v[3][q] = x.d; }
      else { v[0][q] = v[1][q] = v[2][q] = v[3][q] = 0; }
    }
#pragma unroll
    for (int j = 0; j < kGbRows; j++) { pass[j] = true; if (!KEYED) key[j] = 0; }
  } else {
#pragma unroll
    for (int j = 0; j < kGbRows; j++) {
      const uint64_t row = row0 + j;
      pass[j] = row >= begin && row < end;
      key[j] = (KEYED && pass[j] && !(PROG && p.key_prog.n)) ? ld_stream_u64(p.keys + row) : 0;
#pragma unroll
      for (int q = 0; q < NV; q++) v[j][q] = (pass[j] && q < p.n_vcols && (!PROG || p.prog[q].n == 0)) ? ld_stream_u64(p.vcol[q] + row) : 0;
    }
  }
#pragma unroll
  for (int q = 0; q < NV; q++)
    if (PROG && q < p.n_vcols && p.prog[q].n) {  // computed value column: evaluate its program on these four rows
      uint64_t r[4];
      if (p.prog_vec) eval_prog4<true>(p.prog[q], row0, begin, end, r); else eval_prog4<false>(p.prog[q], row0, begin, end, r);
#pragma unroll
      for (int j = 0; j < kGbRows; j++) v[j][q] = r[j];
    }
  if (PROG && KEYED && p.key_prog.n) {  // computed group key (GROUP BY returnflag * 2 + linestatus): no projection pass, no key column
    uint64_t r[4];
    if (p.prog_vec) eval_prog4<true>(p.key_prog, row0, begin, end, r); else eval_prog4<false>(p.key_prog, row0, begin, end, r);
#pragma unroll
    for (int j = 0; j < kGbRows; j++) key[j] = r[j];
  }
  if (p.terms.n) {  // the fused filter: 256-bit loads of the predicate columns, operator switch once per four rows
    bool f[4];
    if (p.terms_vec) eval_quad<true>(p.terms, row0, end, f); else eval_quad<false>(p.terms, row0, end, f);
#pragma unroll
    for (int j = 0; j < kGbRows; j++) pass[j] = pass[j] && f[j];
  }
}

// ------------------------------------------------------------------------------------------- no-key regime
// Branch-free row body: every value column keeps all six accumulators (i64 / f64 sum, min, max) in registers -- ~25
// instructions per row and column -- and the requested (op, column) cells are picked once at the end.  (A per-row
// runtime dispatch over the cell list cost ~100 instructions per row and left the kernel issue-bound at 26 % of HBM.)
template <int NV>
struct AllAcc {
  uint64_t add_i[NV]; double add_f[NV]; long long min_i[NV], max_i[NV], min_f[NV], max_f[NV];
};

template <int NV, bool PROG = false>
__global__ void __launch_bounds__(kGbThreads) gb_nokey_kernel(GbParams p) {
  AllAcc<NV> a;
#pragma unroll
  for (int q = 0; q < NV; q++) {
    a.add_i[q] = 0; a.add_f[q] = 0.0;
    a.min_i[q] = INT64_MAX; a.max_i[q] = INT64_MIN;
    a.min_f[q] = (long long)cell_identity(MIN_F64); a.max_f[q] = (long long)cell_identity(MAX_F64);
  }
  unsigned long long cnt = 0;
  constexpr int U = NV <= 2 ? 2 : 1;  // independent 256-bit loads per column in flight per thread
  const uint64_t tile = (uint64_t)kGbThreads * kGbRows;
  for (uint64_t base = blockIdx.x * tile * U; base < p.n; base += (uint64_t)gridDim.x * tile * U) {
    uint64_t key[U][kGbRows], v[U][kGbRows][NV];
    bool pass[U][kGbRows];
#pragma unroll
    for (int u = 0; u < U; u++) load_rows<NV, false, PROG>(p, base + u * tile + (uint64_t)threadIdx.x * kGbRows, 0, p.n, key[u], v[u], pass[u]);
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
      for (int j = 0; j < kGbRows; j++) {
        const bool ok = pass[u][j];
        cnt += ok;
#pragma unroll
        for (int q = 0; q < NV; q++) {
          const uint64_t x = v[u][j][q];
          const double d = __longlong_as_double((long long)x);
          const long long o = f64_ord(x);
          const bool fok = ok && d == d;  // NaN never wins a MIN / MAX (executor.rs:319-324 compares with < and >)
          a.add_i[q] += ok ? x : 0;
          a.add_f[q] += ok ? d : 0.0;
          a.min_i[q] = ok ? min(a.min_i[q], (long long)x) : a.min_i[q];
          a.max_i[q] = ok ? max(a.max_i[q], (long long)x) : a.max_i[q];
          a.min_f[q] = fok ? min(a.min_f[q], o) : a.min_f[q];
          a.max_f[q] = fok ? max(a.max_f[q], o) : a.max_f[q];
        }
      }
  }
  // pick the requested cells
  uint64_t acc[kMaxCells];
#pragma unroll
  for (int c = 0; c < kMaxCells; c++) {
    uint64_t r = 0;
    if (c < p.n_cells) {
#pragma unroll
      for (int q = 0; q < NV; q++)
        if (p.cell_vcol[c] == q) {
          switch (p.cell_op[c]) {
            case ADD_I64: r = a.add_i[q]; break;
            case ADD_F64: r = (uint64_t)__double_as_longlong(a.add_f[q]); break;
            case MIN_I64: r = (uint64_t)a.min_i[q]; break;
            case MAX_I64: r = (uint64_t)a.max_i[q]; break;
            case MIN_F64: r = (uint64_t)a.min_f[q]; break;
            default: r = (uint64_t)a.max_f[q];
          }
        }
    }
    acc[c] = r;
  }
  // warp reduce, then one atomic per CTA (f64 sums: the order varies run to run; see DESIGN.md on f64 parity)
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
#pragma unroll
    for (int c = 0; c < kMaxCells; c++) s_acc[warp][c] = acc[c];
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
// begin/end: row range (whole input, or one partition of the partitioned copies)
template <int NV, bool PROG = false>
__global__ void __launch_bounds__(kGbThreads) gb_global_kernel(GbParams p, uint64_t begin, uint64_t end) {
  const uint64_t tile = (uint64_t)kGbThreads * kGbRows;
  for (uint64_t base = (begin & ~3ULL) + blockIdx.x * tile; base < end; base += (uint64_t)gridDim.x * tile) {
    uint64_t key[kGbRows], v[kGbRows][NV];
    bool pass[kGbRows];
    load_rows<NV, true, PROG>(p, base + (uint64_t)threadIdx.x * kGbRows, begin, end, key, v, pass);
    global_update_rows<NV>(p, key, v, pass);
  }
}

// ------------------------------------------------------------------------------------------- shared-memory tables
// Dynamic smem layout: keys[total] | cells[n_cells][total] | count[total] (u32), total = (s_slots + 1) * s_copies
struct SmemTable {
  uint64_t* keys; uint64_t* cells; uint32_t* count; uint32_t slots1, total;
};
__device__ __forceinline__ SmemTable smem_table(const GbParams& p, unsigned char* raw) {
  SmemTable t;
  t.slots1 = p.s_slots + 1; t.total = t.slots1 * p.s_copies;
  t.keys = reinterpret_cast<uint64_t*>(raw);
  t.cells = t.keys + t.total;
  t.count = reinterpret_cast<uint32_t*>(t.cells + (uint64_t)p.n_cells * t.total);
  return t;
}
__device__ __forceinline__ void smem_table_clear(const GbParams& p, const SmemTable& t) {
  for (uint32_t i = threadIdx.x; i < t.total; i += kGbThreads) {
    t.keys[i] = kEmptyKey;
    t.count[i] = 0;
    for (int c = 0; c < p.n_cells; c++) t.cells[(uint64_t)c * t.total + i] = cell_identity(p.cell_op[c]);
  }
}

// Home slot of a key in a shared-memory table: one multiply (Fibonacci hashing of the folded key), the bits right below
// the ones that chose the radix partition (s_hskip = partition bits, 0 in the shared regime).
__device__ __forceinline__ uint32_t smem_home(const GbParams& p, uint64_t key) {
  const uint64_t h = (key ^ (key >> 29)) * 0x9E3779B97F4A7C15ULL;
  return (uint32_t)((h << (p.s_hskip + p.s_sub_bits)) >> p.s_hshift);  // the TOP remaining bits: a lower window of a multiplicative hash clusters
}

// hash sub-range of a key inside its partition (only called when s_sub_bits > 0)
__device__ __forceinline__ uint32_t smem_sub(const GbParams& p, uint64_t key) {
  const uint64_t h = (key ^ (key >> 29)) * 0x9E3779B97F4A7C15ULL;
  return (uint32_t)((h << p.s_hskip) >> (64 - p.s_sub_bits));
}

// Aggregate this thread's kGbRows rows into the thread's copy of the table.  Every lane calls it (pass = false without a row).
// The first version handled one row at a time -- a warp-uniform probe loop per row and a runtime switch over the cell list
// per row and cell: 322 warp instructions per row (ncu), issue-bound at 13-15 % of HBM.  Here the home slots of all four
// rows are probed up front (a hit, the steady state, never enters the loop), the insert / collision loop is shared by the
// four rows, and the cell list is walked once per four rows with the operator switch outside the row loop.
// Rows whose key finds no room fall through to the global table.
template <int NV>
__device__ __forceinline__ void smem_update_rows(const GbParams& p, const SmemTable& t, uint32_t copy_base, const uint64_t (&key)[kGbRows],
                                                 const uint64_t (&v)[kGbRows][NV], const bool (&pass)[kGbRows], uint32_t* spill_flag) {
  constexpr uint32_t kNone = 0xFFFFFFFFu;
  const uint32_t smask = p.s_slots - 1;
  uint32_t s[kGbRows], slot[kGbRows], pend = 0;
#pragma unroll
  for (int j = 0; j < kGbRows; j++) {
    s[j] = smem_home(p, key[j]);
    slot[j] = kNone;
    if (pass[j]) {
      if (key[j] == kEmptyKey) slot[j] = copy_base + p.s_slots;
      else if (*(volatile uint64_t*)(t.keys + copy_base + s[j]) == key[j]) slot[j] = copy_base + s[j];
      else pend |= 1u << j;
    }
  }
  uint32_t probes = 0;
  while (__any_sync(0xffffffffu, pend != 0)) {  // first sight of a key, or a collision
#pragma unroll
    for (int j = 0; j < kGbRows; j++)
      if (pend & (1u << j)) {
        uint64_t* kp = t.keys + copy_base + s[j];
        uint64_t k = *(volatile uint64_t*)kp;
        if (k == kEmptyKey) k = atomicCAS((unsigned long long*)kp, (unsigned long long)kEmptyKey, (unsigned long long)key[j]);
        if (k == key[j] || k == kEmptyKey) { slot[j] = copy_base + s[j]; pend &= ~(1u << j); }
        else s[j] = (s[j] + 1) & smask;
      }
    if (++probes > p.s_slots) break;  // the copy is full (same trip count in every lane): what is still pending spills
  }
  if (__any_sync(0xffffffffu, pend != 0)) {
    if (pend && spill_flag) *spill_flag = 1;
#pragma unroll
    for (int j = 0; j < kGbRows; j++) global_update_row<NV>(p, key[j], v[j], (pend >> j) & 1u);
  }
#pragma unroll
  for (int j = 0; j < kGbRows; j++) if (slot[j] != kNone) atomicAdd(t.count + slot[j], 1u);
#pragma unroll 1
  for (int c = 0; c < p.n_cells; c++) {
    const int op = p.cell_op[c], q = p.cell_vcol[c];
    uint64_t* cells = t.cells + (uint64_t)c * t.total;
    uint64_t x[kGbRows];
#pragma unroll
    for (int j = 0; j < kGbRows; j++) x[j] = pick<NV>(v[j], q);
    switch (op) {
      case ADD_I64:
#pragma unroll
        for (int j = 0; j < kGbRows; j++)
          if (slot[j] != kNone) {
            uint32_t* w = reinterpret_cast<uint32_t*>(cells + slot[j]);
            const uint32_t lo = (uint32_t)x[j], hi = (uint32_t)(x[j] >> 32);
            const uint32_t old = atomicAdd(w, lo);
            const uint32_t carry = (uint32_t)(old + lo) < lo;
            if (hi + carry) atomicAdd(w + 1, hi + carry);
          }
        break;
      case ADD_F64:
#pragma unroll
        for (int j = 0; j < kGbRows; j++) if (slot[j] != kNone) atomicAdd((double*)(cells + slot[j]), __longlong_as_double((long long)x[j]));
        break;
      case MIN_I64:
#pragma unroll
        for (int j = 0; j < kGbRows; j++)
          if (slot[j] != kNone && (long long)x[j] < *(volatile long long*)(cells + slot[j])) atomicMin((long long*)(cells + slot[j]), (long long)x[j]);
        break;
      case MAX_I64:
#pragma unroll
        for (int j = 0; j < kGbRows; j++)
          if (slot[j] != kNone && (long long)x[j] > *(volatile long long*)(cells + slot[j])) atomicMax((long long*)(cells + slot[j]), (long long)x[j]);
        break;
      case MIN_F64:
#pragma unroll
        for (int j = 0; j < kGbRows; j++) {
          const double d = __longlong_as_double((long long)x[j]);
          const long long o = f64_ord(x[j]);
          if (slot[j] != kNone && d == d && o < *(volatile long long*)(cells + slot[j])) atomicMin((long long*)(cells + slot[j]), o);
        }
        break;
      default:
#pragma unroll
        for (int j = 0; j < kGbRows; j++) {
          const double d = __longlong_as_double((long long)x[j]);
          const long long o = f64_ord(x[j]);
          if (slot[j] != kNone && d == d && o > *(volatile long long*)(cells + slot[j])) atomicMax((long long*)(cells + slot[j]), o);
        }
    }
  }
}

// folds one pre-aggregated shared-memory slot into the global table (warp-uniform)
__device__ __forceinline__ void global_merge_slot(const GbParams& p, const SmemTable& t, uint64_t key, uint32_t cnt, uint32_t src, bool active) {
  const uint64_t s = global_slot(p, key, active);
  if (!active || s == ~0ULL) return;
  atomicAdd(p.g_count + s, (unsigned long long)cnt);
  for (int c = 0; c < p.n_cells; c++) cell_atomic_global(p.cell_op[c], p.g_cells + (uint64_t)c * (p.g_cap + 1) + s, t.cells[(uint64_t)c * t.total + src]);
}

// folds one pre-aggregated slot (count + cell values of copy c > 0) into copy 0; warp-uniform like smem_update_row
__device__ __forceinline__ void smem_merge_slot(const GbParams& p, const SmemTable& t, uint64_t key, uint32_t cnt, uint32_t src, bool active) {
  const uint32_t smask = p.s_slots - 1;
  uint32_t s = smem_home(p, key), slot = 0xFFFFFFFFu, probes = 0;
  bool pending = active;
  if (pending && key == kEmptyKey) { slot = p.s_slots; pending = false; }
  while (__any_sync(0xffffffffu, pending)) {
    if (pending) {
      const uint64_t k = *(volatile uint64_t*)(t.keys + s);
      if (k == key) { slot = s; pending = false; }
      else {
        if (k == kEmptyKey) {
          const uint64_t old = atomicCAS((unsigned long long*)(t.keys + s), (unsigned long long)kEmptyKey, (unsigned long long)key);
          if (old == kEmptyKey || old == key) { slot = s; pending = false; }
        }
        if (pending) { s = (s + 1) & smask; if (++probes >= p.s_slots) pending = false; }
      }
    }
  }
  // copy 0 is full (the union of the copies' keys exceeds it): this key lives in the global table only
  const bool spill = active && slot == 0xFFFFFFFFu;
  if (__any_sync(0xffffffffu, spill)) global_merge_slot(p, t, key, cnt, src, spill);
  if (!active || spill) return;
  atomicAdd(t.count + slot, cnt);
  for (int c = 0; c < p.n_cells; c++) cell_atomic_shared(p.cell_op[c], t.cells + (uint64_t)c * t.total + slot, t.cells[(uint64_t)c * t.total + src]);
}

template <int NV, bool PROG = false>
__global__ void __launch_bounds__(kGbThreads) gb_shared_kernel(GbParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const SmemTable t = smem_table(p, smem_raw);
  smem_table_clear(p, t);
  __syncthreads();
  const uint32_t copy_base = (threadIdx.x & (p.s_copies - 1)) * t.slots1;
  const uint64_t tile = (uint64_t)kGbThreads * kGbRows, stride = (uint64_t)gridDim.x * tile;
  // software pipeline: the next tile's rows are in flight while this tile's dependent shared-memory updates run
  uint64_t nkey[kGbRows], nv[kGbRows][NV];
  bool npass[kGbRows];
  uint64_t base = blockIdx.x * tile;
  if (base < p.n) load_rows<NV, true, PROG>(p, base + (uint64_t)threadIdx.x * kGbRows, 0, p.n, nkey, nv, npass);
  for (; base < p.n; base += stride) {
    uint64_t key[kGbRows], v[kGbRows][NV];
    bool pass[kGbRows];
#pragma unroll
    for (int j = 0; j < kGbRows; j++) {
      key[j] = nkey[j]; pass[j] = npass[j];
#pragma unroll
      for (int q = 0; q < NV; q++) v[j][q] = nv[j][q];
    }
    if (base + stride < p.n) load_rows<NV, true, PROG>(p, base + stride + (uint64_t)threadIdx.x * kGbRows, 0, p.n, nkey, nv, npass);
    smem_update_rows<NV>(p, t, copy_base, key, v, pass, nullptr);
  }
  __syncthreads();
  // merge every copy into the global table (warp-uniform trip count: global_slot is a convergent call)
  for (uint32_t i0 = 0; i0 < t.total; i0 += kGbThreads) {
    const uint32_t i = i0 + threadIdx.x;
    const uint32_t cnt = i < t.total ? t.count[i] : 0;
    const uint64_t key = (cnt && i % t.slots1 != p.s_slots) ? t.keys[i] : kEmptyKey;
    global_merge_slot(p, t, key, cnt, i, cnt != 0);
  }
}

// one CTA per partition (grid-stride over partitions): every partition's groups are disjoint from all others', so the
// table goes straight to the output; rows that do not fit the table spill to the global table (extracted afterwards)
template <int NV>
__global__ void __launch_bounds__(kGbThreads) gb_part_smem_kernel(GbParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ unsigned long long s_base;
  __shared__ uint32_t s_groups, s_spill;
  const SmemTable t = smem_table(p, smem_raw);
  const uint64_t tile = (uint64_t)kGbThreads * kGbRows;
  // work unit = (partition, hash sub-range): a partition whose groups do not fit one table is read 2^s_sub_bits times, each
  // pass aggregating the keys of one sub-range -- still far cheaper than L2 atomics on a global table
  const uint32_t n_units = p.n_parts << p.s_sub_bits;
  for (uint32_t unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
    const uint32_t part = unit >> p.s_sub_bits, sub = unit & ((1u << p.s_sub_bits) - 1);
    smem_table_clear(p, t);
    if (threadIdx.x == 0) { s_groups = 0; s_spill = 0; }
    __syncthreads();
    const uint64_t begin = p.part_begin[part], end = p.part_end[part];
    const uint32_t copy_base = (threadIdx.x & (p.s_copies - 1)) * t.slots1;
    uint64_t nkey[kGbRows], nv[kGbRows][NV];
    bool npass[kGbRows];
    uint64_t base = begin & ~3ULL;
    if (base < end) load_rows<NV, true>(p, base + (uint64_t)threadIdx.x * kGbRows, begin, end, nkey, nv, npass);
    for (; base < end; base += tile) {
      uint64_t key[kGbRows], v[kGbRows][NV];
      bool pass[kGbRows];
#pragma unroll
      for (int j = 0; j < kGbRows; j++) {
        key[j] = nkey[j]; pass[j] = npass[j];
#pragma unroll
        for (int q = 0; q < NV; q++) v[j][q] = nv[j][q];
      }
      if (base + tile < end) load_rows<NV, true>(p, base + tile + (uint64_t)threadIdx.x * kGbRows, begin, end, nkey, nv, npass);
      if (p.s_sub_bits) {
#pragma unroll
        for (int j = 0; j < kGbRows; j++) pass[j] = pass[j] && smem_sub(p, key[j]) == sub;
      }
      smem_update_rows<NV>(p, t, copy_base, key, v, pass, &s_spill);
    }
    __syncthreads();
    if (s_spill) {  // some row went to the global table while its key may sit in another copy: this partition goes there entirely
      for (uint32_t i0 = 0; i0 < t.total; i0 += kGbThreads) {
        const uint32_t i = i0 + threadIdx.x;
        const uint32_t cnt = i < t.total ? t.count[i] : 0;
        const uint64_t key = (cnt && i % t.slots1 != p.s_slots) ? t.keys[i] : kEmptyKey;
        global_merge_slot(p, t, key, cnt, i, cnt != 0);
      }
      __syncthreads();
      continue;
    }
    // fold copies 1.. into copy 0; a slot that finds copy 0 full goes (pre-aggregated) to the global table instead, so the
    // direct output below and the final extraction of the global table stay disjoint
    for (uint32_t i0 = t.slots1; i0 < t.total; i0 += kGbThreads) {
      const uint32_t i = i0 + threadIdx.x;
      const uint32_t cnt = i < t.total ? t.count[i] : 0;
      const uint64_t key = (cnt && i % t.slots1 != p.s_slots) ? t.keys[i] : kEmptyKey;
      smem_merge_slot(p, t, key, cnt, i, cnt != 0);
      __syncwarp();
    }
    __syncthreads();
    // count the groups, reserve a dense output range with one global atomic, write them out
    uint32_t mine = 0;
    for (uint32_t i = threadIdx.x; i < t.slots1; i += kGbThreads) mine += t.count[i] != 0;
    uint32_t my_off = 0;
    if (mine) my_off = atomicAdd(&s_groups, mine);
    __syncthreads();
    if (threadIdx.x == 0 && s_groups) s_base = atomicAdd(p.out_n, (unsigned long long)s_groups);
    __syncthreads();
    uint64_t g = s_base + my_off;
    for (uint32_t i = threadIdx.x; i < t.slots1; i += kGbThreads) {
      const uint32_t cnt = t.count[i];
      if (!cnt) continue;
      if (g < p.out_cap) {
        uint64_t cells[kMaxCells];
#pragma unroll
        for (int c = 0; c < kMaxCells; c++) cells[c] = c < p.n_cells ? t.cells[(uint64_t)c * t.total + i] : 0;
        emit_group(p.out, p.out_keys, p.out_aggs, p.out_cap, g, i == p.s_slots ? INT64_MIN : (int64_t)t.keys[i], cnt, cells);
      }
      g++;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------- dense-key regime
// Group keys that fill a range (max - min + 1 <= 16 x the group-count hint: surrogate keys, dates, the benchmark's uniform
// [0, G)) need no hash table at all: the row's slot is key - min.  What they need above ~7 K groups is a table that fits shared
// memory, so the rows are first split by key RANGE -- the join's partition pass (djoin.cu) with 12-byte rows {key - min : u32,
// value : u64} kept as two arrays -- into ranges of 2^12 keys (one pass up to 2^20 keys, two levels up to 2^28), and one CTA then
// aggregates a range in a 4096-slot direct-address table in shared memory and writes its groups straight to the output.
// Measured before (profiles/r01_bench_ops.jsonl, 2^30 rows): 10^5 groups 15.4 ms (hash partition of 16-byte rows + hash tables),
// 10^6 30 ms, 10^8 57 ms (per-partition L2 tables, L2-atomic bound).
constexpr int kGrKeysLog2 = 12;
constexpr uint32_t kGrKeys = 1u << kGrKeysLog2;
constexpr int kGrThreads = 256;   // rows per thread and tile (ROWS): 16 at 3 CTAs per SM, or 8 at 5 CTAs per SM -- see launch_gr_part
enum { kGrFlagOverflow = 0, kGrFlagOutOfRange = 1, kGrFlagWords = 4 };

struct GrPartArgs {
  const uint64_t* keys; const uint64_t* vals; uint64_t n;        // level 0: raw columns (vals may be NULL: COUNT only)
  uint64_t rmin, rrange;
  const uint32_t* in_d; const uint64_t* in_v;                     // level 1: the regions of level 0
  const uint32_t* reg_cur; uint32_t reg_cap, n_reg, reg_parts;
  int shift; uint32_t cap;
  uint32_t* cursor; uint32_t* out_d; uint64_t* out_v; uint32_t* flags;
};

template <bool REGIONS, int ROWS>
__global__ void __launch_bounds__(kGrThreads, ROWS == 16 ? 3 : 5) gr_part_kernel(const __grid_constant__ GrPartArgs a) {
  constexpr int R = ROWS, NW = kGrThreads / 32, kGrTile = kGrThreads * ROWS;
  extern __shared__ __align__(16) unsigned char gr_smem[];
  uint64_t* stage_v = reinterpret_cast<uint64_t*>(gr_smem);               // [kGrTile]
  uint32_t* stage_d = reinterpret_cast<uint32_t*>(stage_v + kGrTile);     // [kGrTile]
  __shared__ uint32_t cnt[kGrThreads], gbase[kGrThreads], s_warp[NW];
  __shared__ uint32_t s_total, s_over[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  cnt[tid] = 0;
  if (tid == 0) s_over[0] = s_over[1] = 0;
  __syncthreads();
  const uint32_t tiles_per_reg = REGIONS ? (a.reg_cap + kGrTile - 1) / kGrTile : 1u;
  const uint32_t tiles = REGIONS ? tiles_per_reg * a.n_reg : (uint32_t)((a.n + kGrTile - 1) / kGrTile);
  int par = 1;
  for (uint32_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    uint64_t n = a.n, row0 = (uint64_t)t * kGrTile, base = 0;
    uint32_t pbase = 0;
    if (REGIONS) {
      const uint32_t g = t / tiles_per_reg;
      n = min(a.reg_cur[g], a.reg_cap);
      row0 = (uint64_t)(t - g * tiles_per_reg) * kGrTile;
      base = (uint64_t)g * a.reg_cap;
      pbase = g * a.reg_parts;
      if (row0 >= n) continue;
    }
    par ^= 1;
    uint32_t d[R], rk[R];
    uint64_t v[R];
    bool oob = false;
#pragma unroll
    for (int k = 0; k < R; k++) {
      const uint64_t row = row0 + (uint64_t)k * kGrThreads + tid;
      const bool in = row < n;
      if (REGIONS) {
        d[k] = in ? a.in_d[base + row] : 0u;
        v[k] = in ? ld_stream_u64(a.in_v + base + row) : 0;
        rk[k] = in ? 0u : 0xFFFFFFFFu;
      } else {
        const uint64_t off = (in ? ld_stream_u64(a.keys + row) : 0) - a.rmin;
        v[k] = (in && a.vals) ? ld_stream_u64(a.vals + row) : 0;
        const bool ok = in && off < a.rrange;
        oob = oob || (in && !ok);
        d[k] = (uint32_t)off;
        rk[k] = ok ? 0u : 0xFFFFFFFFu;
      }
    }
#pragma unroll
    for (int k = 0; k < R; k++) if (rk[k] == 0u) rk[k] = atomicAdd(&cnt[(d[k] >> a.shift) - pbase], 1u);
    if (!REGIONS && oob) a.flags[kGrFlagOutOfRange] = 1;  // the remembered key range is stale: the caller recomputes it
    __syncthreads();
    const uint32_t c = cnt[tid];
    const uint32_t got = c ? atomicAdd(&a.cursor[pbase + tid], c) : 0u;
    if (tid == 0) s_over[par ^ 1] = 0;
    uint32_t inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += x; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    {
      uint32_t wbase = 0;
#pragma unroll
      for (int w = 0; w < NW; w++) wbase += w < warp ? s_warp[w] : 0u;
      const uint32_t off = wbase + inc - c;
      cnt[tid] = off;
      gbase[tid] = (pbase + (uint32_t)tid) * a.cap + got - off;
      if (got + c > a.cap) { s_over[par] = 1; a.flags[kGrFlagOverflow] = 1; }
      if (tid == kGrThreads - 1) s_total = off + c;
    }
    __syncthreads();
    const uint32_t total = s_total;
#pragma unroll
    for (int k = 0; k < R; k++) {
      if (rk[k] == 0xFFFFFFFFu) continue;
      const uint32_t pos = cnt[(d[k] >> a.shift) - pbase] + rk[k];
      stage_d[pos] = d[k];
      stage_v[pos] = v[k];
    }
    __syncthreads();
    cnt[tid] = 0;
    const bool over = s_over[par] != 0;
    for (uint32_t i = tid; i < total; i += kGrThreads) {
      const uint32_t dd = stage_d[i], part = (dd >> a.shift) - pbase, dst = gbase[part] + i;
      if (over && dst - (pbase + part) * a.cap >= a.cap) continue;  // the region is full: dropped, flagged, the caller takes another regime
      a.out_d[dst] = dd;
      a.out_v[dst] = stage_v[i];
    }
    __syncthreads();
  }
}

struct GrAggArgs {
  const uint32_t* d; const uint64_t* v; const uint32_t* cur; uint32_t cap, P;
  uint32_t S;                                  // CTAs that share one range (few ranges: 10^4 groups are 3 ranges for 296 CTAs); S > 1: merge
  unsigned long long* m_cnt; uint64_t* m_cells;  // S > 1: global direct table [P * 4096] (+ cells [n_cells][P * 4096]) the CTAs' tables are merged into
  uint32_t* ticket;
  int n_cells; int8_t cell_op[kMaxCells];
  OutSpec out; int64_t* out_keys; uint64_t* out_aggs; uint64_t out_cap; unsigned long long* out_n;
  int64_t rmin;
};

constexpr int kGrAggThreads = 512;

__global__ void __launch_bounds__(kGrAggThreads, 2) gr_agg_kernel(const __grid_constant__ GrAggArgs a) {
  extern __shared__ __align__(16) unsigned char gr_smem[];
  uint64_t* cells = reinterpret_cast<uint64_t*>(gr_smem);                          // [n_cells][kGrKeys]
  uint32_t* cnt = reinterpret_cast<uint32_t*>(cells + (size_t)a.n_cells * kGrKeys);  // [kGrKeys]
  __shared__ uint32_t s_part;
  const int tid = threadIdx.x, lane = tid & 31;
  for (;;) {
    __syncthreads();
    if (tid == 0) s_part = atomicAdd(a.ticket, 1u);
    __syncthreads();
    const uint32_t item = s_part;
    if (item >= a.P * a.S) break;
    const uint32_t p = item / a.S, slice = item - p * a.S;
    const uint32_t np = min(a.cur[p], a.cap);
    const uint32_t per = ((np + a.S - 1) / a.S + 3) & ~3u, r0 = min(np, slice * per), n = min(np, r0 + per) - r0;
    if (n == 0) continue;
    for (uint32_t i = tid; i < kGrKeys; i += kGrAggThreads) cnt[i] = 0;
#pragma unroll 1
    for (int c = 0; c < a.n_cells; c++) {
      const uint64_t id = cell_identity(a.cell_op[c]);
      for (uint32_t i = tid; i < kGrKeys; i += kGrAggThreads) cells[(size_t)c * kGrKeys + i] = id;
    }
    __syncthreads();
    const uint32_t* dd = a.d + (uint64_t)p * a.cap + r0;
    const uint64_t* vv = a.v + (uint64_t)p * a.cap + r0;
    for (uint32_t i0 = 0; i0 < n; i0 += kGrAggThreads * 8) {
      uint32_t k[8]; uint64_t x[8];
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const uint32_t i = i0 + j * kGrAggThreads + tid;
        k[j] = i < n ? dd[i] & (kGrKeys - 1) : 0xFFFFFFFFu;
        x[j] = i < n ? ld_stream_u64(vv + i) : 0;
      }
#pragma unroll
      for (int j = 0; j < 8; j++) if (k[j] != 0xFFFFFFFFu) atomicAdd(&cnt[k[j]], 1u);
#pragma unroll 1
      for (int c = 0; c < a.n_cells; c++) {
        const int op = a.cell_op[c];
        uint64_t* cc = cells + (size_t)c * kGrKeys;
#pragma unroll
        for (int j = 0; j < 8; j++) {
          if (k[j] == 0xFFFFFFFFu) continue;
          bool valid;
          const uint64_t e = cell_encode(op, x[j], valid);
          if (valid) cell_atomic_shared(op, cc + k[j], e);
        }
      }
    }
    __syncthreads();
    if (a.S > 1) {  // several CTAs per range: this one's partial groups go into the global direct table; gr_emit_kernel writes the output
      const uint64_t slots = (uint64_t)a.P << kGrKeysLog2;
      for (uint32_t i = tid; i < kGrKeys; i += kGrAggThreads) {
        const uint32_t cn = cnt[i];
        if (!cn) continue;
        const uint64_t g = ((uint64_t)p << kGrKeysLog2) + i;
        atomicAdd(a.m_cnt + g, (unsigned long long)cn);
#pragma unroll 1
        for (int c = 0; c < a.n_cells; c++) cell_atomic_global(a.cell_op[c], a.m_cells + (uint64_t)c * slots + g, cells[(size_t)c * kGrKeys + i]);
      }
      continue;
    }
    // every non-empty slot is a group: its key is the range's first key plus the slot
    for (uint32_t i0 = 0; i0 < kGrKeys; i0 += kGrAggThreads) {
      const uint32_t i = i0 + tid;
      const uint32_t cn = cnt[i];
      const uint32_t m = __ballot_sync(0xffffffffu, cn != 0);
      if (m == 0) continue;
      unsigned long long g0 = 0;
      const int leader = __ffs(m) - 1;
      if (lane == leader) g0 = atomicAdd(a.out_n, (unsigned long long)__popc(m));
      g0 = __shfl_sync(0xffffffffu, g0, leader);
      if (cn) {
        const unsigned long long g = g0 + __popc(m & ((1u << lane) - 1));
        if (g < a.out_cap) {
          uint64_t cv[kMaxCells];
#pragma unroll
          for (int c = 0; c < kMaxCells; c++) cv[c] = c < a.n_cells ? cells[(size_t)c * kGrKeys + i] : 0;
          emit_group(a.out, a.out_keys, a.out_aggs, a.out_cap, g, a.rmin + (int64_t)(((uint64_t)p << kGrKeysLog2) + i), cn, cv);
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256) gr_merge_init_kernel(GrAggArgs a) {
  const uint64_t slots = (uint64_t)a.P << kGrKeysLog2;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < slots; i += (uint64_t)gridDim.x * blockDim.x) {
    a.m_cnt[i] = 0;
    for (int c = 0; c < a.n_cells; c++) a.m_cells[(uint64_t)c * slots + i] = cell_identity(a.cell_op[c]);
  }
}

__global__ void __launch_bounds__(256) gr_emit_kernel(GrAggArgs a) {
  const uint64_t slots = (uint64_t)a.P << kGrKeysLog2;
  const int lane = threadIdx.x & 31;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i0 = blockIdx.x * (uint64_t)blockDim.x; i0 < slots; i0 += stride) {   // whole warps stay together for the ballot
    const uint64_t i = i0 + threadIdx.x;
    const unsigned long long cn = i < slots ? a.m_cnt[i] : 0ULL;
    const uint32_t m = __ballot_sync(0xffffffffu, cn != 0);
    if (m == 0) continue;
    unsigned long long g0 = 0;
    const int leader = __ffs(m) - 1;
    if (lane == leader) g0 = atomicAdd(a.out_n, (unsigned long long)__popc(m));
    g0 = __shfl_sync(0xffffffffu, g0, leader);
    if (cn) {
      const unsigned long long g = g0 + __popc(m & ((1u << lane) - 1));
      if (g < a.out_cap) {
        uint64_t cv[kMaxCells];
#pragma unroll
        for (int c = 0; c < kMaxCells; c++) cv[c] = c < a.n_cells ? a.m_cells[(uint64_t)c * slots + i] : 0;
        emit_group(a.out, a.out_keys, a.out_aggs, a.out_cap, g, a.rmin + (int64_t)i, cn, cv);
      }
    }
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

__global__ void __launch_bounds__(256) gb_extract_kernel(GbParams p, bool emit_empty_global) {
  const uint64_t n = p.g_cap + 1, stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i0 = blockIdx.x * (uint64_t)blockDim.x; i0 < n; i0 += stride) {
    const uint64_t i = i0 + threadIdx.x;
    const unsigned long long cnt = i < n ? p.g_count[i] : 0;
    const bool occ = cnt != 0 || (emit_empty_global && i == 0);
    const unsigned m = __ballot_sync(0xffffffffu, occ);
    if (!m) continue;
    const int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(p.out_n, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (!occ) continue;
    const uint64_t g = base + __popc(m & ((1u << lane) - 1));
    if (g >= p.out_cap) continue;
    uint64_t cells[kMaxCells];
#pragma unroll
    for (int c = 0; c < kMaxCells; c++) cells[c] = c < p.n_cells ? p.g_cells[(uint64_t)c * n + i] : 0;
    emit_group(p.out, p.out_keys, p.out_aggs, p.out_cap, g, p.keys ? (i == p.g_cap ? INT64_MIN : (int64_t)p.g_keys[i]) : 0, cnt, cells);
  }
}

// ------------------------------------------------------------------------------------------- host driver
static olap_status convert_prog(const olap_expr_prog* g, bool f64, ExprProg* out) {
  OLAP_REQUIRE(g->n_ops >= 1 && g->n_ops <= kMaxProgOps, "expression program: 1..OLAP_PROG_MAX_OPS entries");
  ExprProg& e = *out;
  e.n = (int8_t)g->n_ops; e.f64 = f64;
  int depth = 0;
  for (int k = 0; k < g->n_ops; k++) {
    const int op = g->op[k];
    e.op[k] = (int8_t)op; e.col[k] = (const uint64_t*)g->column[k]; e.lit[k] = (uint64_t)g->literal_bits[k];
    if (op == 'c') { OLAP_REQUIRE(g->column[k], "expression program: NULL column"); depth++; }
    else if (op == 'l') depth++;
    else { OLAP_REQUIRE(op == '+' || op == '-' || op == '*' || op == '/', "expression program: unknown op"); OLAP_REQUIRE(depth >= 2, "expression program: stack underflow"); depth--; }
    OLAP_REQUIRE(depth <= 3, "expression program: stack depth > 3");
  }
  OLAP_REQUIRE(depth == 1, "expression program must leave exactly one value");
  return OLAP_OK;
}

static olap_status build_cells(const olap_agg_spec* specs, const olap_expr_prog* const* progs, int n_aggs, GbParams* p) {
  OLAP_REQUIRE(n_aggs >= 0 && n_aggs <= 16, "n_aggs must be in [0, 16]");
  OutSpec* o = &p->out;
  p->n_cells = 0; p->n_vcols = 0;
  o->n_aggs = n_aggs;
  for (int a = 0; a < n_aggs; a++) {
    const int f = specs[a].func, t = specs[a].value_type;
    OLAP_REQUIRE(f >= OLAP_SUM && f <= OLAP_AVG, "unknown aggregate function");
    OLAP_REQUIRE(t == OLAP_I64 || t == OLAP_F64, "value_type must be OLAP_I64 or OLAP_F64");
    o->func[a] = (int8_t)f; o->type[a] = (int8_t)t; o->cell[a] = -1;
    if (f == OLAP_COUNT) continue;
    const olap_expr_prog* g = progs ? progs[a] : nullptr;
    // a value column is identified by its buffer, or by its program object when it is computed
    const uint64_t* id = g ? (const uint64_t*)g : (const uint64_t*)specs[a].values;
    OLAP_REQUIRE(id, "aggregate value column is NULL");
    int vc = -1;
    for (int q = 0; q < p->n_vcols; q++) if (p->vcol[q] == id) vc = q;
    if (vc < 0) {
      OLAP_REQUIRE(p->n_vcols < kMaxVcols, "more than 6 distinct value columns in one aggregation");
      vc = p->n_vcols++;
      p->vcol[vc] = id;
      if (g) OLAP_TRY(convert_prog(g, t == OLAP_F64, &p->prog[vc]));
    }
    const int8_t op = (f == OLAP_SUM || f == OLAP_AVG) ? (t == OLAP_F64 ? ADD_F64 : ADD_I64)
                      : f == OLAP_MIN ? (t == OLAP_F64 ? MIN_F64 : MIN_I64) : (t == OLAP_F64 ? MAX_F64 : MAX_I64);
    int found = -1;
    for (int c = 0; c < p->n_cells; c++) if (p->cell_op[c] == op && p->cell_vcol[c] == vc) found = c;
    if (found < 0) {
      OLAP_REQUIRE(p->n_cells < kMaxCells, "more than 8 distinct (function, column) accumulators");
      found = p->n_cells++;
      p->cell_op[found] = op;
      p->cell_vcol[found] = (int8_t)vc;
    }
    o->cell[a] = (int8_t)found;
  }
  return OLAP_OK;
}

static uint64_t pow2_at_least(uint64_t x) { uint64_t p = 1; while (p < x) p <<= 1; return p; }
static int ilog2_u64(uint64_t x) { int l = 0; while ((1ULL << l) < x) l++; return l; }
static unsigned capped_grid(uint64_t want, uint64_t cap) { return (unsigned)(want < 1 ? 1 : (want > cap ? cap : want)); }

#define GB_DISPATCH_NV(nv, ...)                                 \
  switch (nv) {                                                 \
    case 0: case 1: { constexpr int NV = 1; __VA_ARGS__; } break; \
    case 2: { constexpr int NV = 2; __VA_ARGS__; } break;       \
    case 3: { constexpr int NV = 3; __VA_ARGS__; } break;       \
    case 4: { constexpr int NV = 4; __VA_ARGS__; } break;       \
    case 5: { constexpr int NV = 5; __VA_ARGS__; } break;       \
    default: { constexpr int NV = 6; __VA_ARGS__; } break;      \
  }

// join.cu: radix-partition (key, value columns) with the fused filter; returns the partition boundaries on the device
uint64_t group_by_partition_capacity(uint64_t n, int bits);
olap_status group_by_partition_rows(olap_ctx* ctx, const uint64_t* keys, uint64_t n, const TermPack& terms, int n_vcols,
                                    const uint64_t* const* vcols, int bits, uint64_t* out_keys, uint64_t* const* out_vcols,
                                    uint32_t* bounds, uint32_t* h_bounds, cudaStream_t s);

// shared-memory budget per CTA for group tables
static uint64_t slot_bytes(const GbParams& p) { return 8 + 4 + 8ULL * p.n_cells; }

static olap_status run_partitioned(olap_ctx* ctx, GbParams& p, uint64_t n_groups_hint, Scratch& sc, cudaStream_t s) {
  const int bits = 8;
  const uint32_t P = 1u << bits;
  uint64_t* pk = nullptr;
  uint64_t* pv[kMaxVcols] = {};
  uint32_t* bounds = nullptr;  // rows of partition q are [bounds[q], bounds[P + q])
  std::vector<uint32_t> h(2 * (size_t)P);
  const uint64_t rows_cap = group_by_partition_capacity(p.n, bits);
  OLAP_TRY(sc.get<uint64_t>(rows_cap + 4, &pk));
  for (int q = 0; q < p.n_vcols; q++) OLAP_TRY(sc.get<uint64_t>(rows_cap + 4, &pv[q]));
  OLAP_TRY(sc.get<uint32_t>(2 * (uint64_t)P, &bounds));
  OLAP_TRY(group_by_partition_rows(ctx, p.keys, p.n, p.terms, p.n_vcols, p.vcol, bits, pk, pv, bounds, h.data(), s));
  phase_mark(ctx, s, "partition");
  GbParams q = p;  // aggregate the partitioned copies; the filter has been applied by the partition pass
  q.keys = pk;
  for (int i = 0; i < p.n_vcols; i++) q.vcol[i] = pv[i];
  q.terms.n = 0;
  q.vec = true;  // pool allocations are 256-byte aligned
  q.part_begin = bounds; q.part_end = bounds + P; q.n_parts = P;
  const uint64_t per_part = (n_groups_hint + P - 1) / P;
  uint64_t need = pow2_at_least((per_part * 3 + 1) / 2);  // <= 67 % load even if the partitions are a little uneven
  if (need < 64) need = 64;
  // a partition whose table exceeds shared memory is aggregated in up to 4 passes over disjoint hash sub-ranges
  int sub_bits = 0;
  while (sub_bits < 2 && need > 64 && (need + 1) * slot_bytes(p) + 16 > 200 * 1024) { need >>= 1; sub_bits++; }
  const uint64_t bytes = (need + 1) * slot_bytes(p) + 16;
  if (bytes <= 200 * 1024) {
    uint64_t copies = 1;  // replicate small tables so that 512 threads do not hammer a few dozen slots (2 CTAs per SM: 100 KB each)
    while (copies < 16 && ((need + 1) * slot_bytes(p) * copies * 2 + 16) <= 100 * 1024) copies <<= 1;
    q.s_slots = (uint32_t)need; q.s_copies = (uint32_t)copies; q.s_hskip = bits; q.s_sub_bits = sub_bits; q.s_hshift = 64 - ilog2_u64(need);
    const uint64_t bytes = (need + 1) * slot_bytes(p) * copies + 16;
    const unsigned per_sm = bytes <= 100 * 1024 ? 2 : 1;
    GB_DISPATCH_NV(p.n_vcols, {
      OLAP_CUDA(cudaFuncSetAttribute(gb_part_smem_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
      gb_part_smem_kernel<NV><<<capped_grid((uint64_t)P << sub_bits, (uint64_t)ctx->sm_count * per_sm), kGbThreads, bytes, s>>>(q);
    });
    count_launch(ctx);
  } else {
    // very high cardinality: every partition aggregates into an L2-resident slice-sized global table, reused in turn
    GbParams t = q;
    t.g_cap = pow2_at_least(std::max<uint64_t>(2 * per_part + per_part / 2, 1024));
    const uint64_t n1 = t.g_cap + 1;
    OLAP_TRY(sc.get<uint64_t>(n1, &t.g_keys));
    OLAP_TRY(sc.get<unsigned long long>(n1, &t.g_count));
    OLAP_TRY(sc.get<uint64_t>(n1 * std::max(p.n_cells, 1), &t.g_cells));
    const unsigned tgrid = capped_grid((n1 + 255) / 256, (uint64_t)ctx->sm_count * 8);
    const uint64_t tile = (uint64_t)kGbThreads * kGbRows;
    for (uint32_t part = 0; part < P; part++) {
      const uint64_t b0 = h[part], b1 = h[P + part];
      if (b1 == b0) continue;
      gb_init_kernel<<<tgrid, 256, 0, s>>>(t);
      GB_DISPATCH_NV(p.n_vcols, (gb_global_kernel<NV><<<capped_grid((b1 - b0 + tile - 1) / tile, (uint64_t)ctx->sm_count * 4), kGbThreads, 0, s>>>(t, b0, b1)));
      gb_extract_kernel<<<tgrid, 256, 0, s>>>(t, false);
      count_launch(ctx, 3);
    }
  }
  OLAP_CUDA(cudaGetLastError());
  return OLAP_OK;
}

olap_status key_min_max(olap_ctx* ctx, const uint64_t* keys, uint64_t n, cudaStream_t s, int64_t* mn, int64_t* mx);  // join.cu

static uint32_t gr_region_cap(uint64_t n, uint64_t parts) {
  const uint64_t mean = (n + parts - 1) / parts;
  return (uint32_t)((mean + mean / 4 + 2048 + 3) & ~3ULL);
}

// *flags_out: 0 = the groups are in the output; bit kGrFlag* set = nothing usable was produced (region overflow: skewed keys;
// out of range: the remembered min / max was stale) and the caller takes another regime / refreshes the range
static olap_status run_dense_range(olap_ctx* ctx, const GbParams& p, int64_t rmin, uint64_t rrange, Scratch& sc, cudaStream_t s, uint32_t* flags_out) {
  const uint64_t P = (rrange + kGrKeys - 1) >> kGrKeysLog2;
  const bool two = P > (uint64_t)kGrThreads;
  const uint64_t C = two ? (P + kGrThreads - 1) / kGrThreads : 0;   // coarse regions of 256 ranges each
  const uint64_t PP = ((two ? C * kGrThreads : P) + 255) / 256 * 256 + 256;
  const uint32_t cap = gr_region_cap(p.n, P), cap1 = two ? gr_region_cap(p.n, C) : 0;
  if (P * cap >= 0xFFFFFFFFULL || (two && C * cap1 >= 0xFFFFFFFFULL)) { *flags_out = 1u << kGrFlagOverflow; return OLAP_OK; }
  uint32_t *d1 = nullptr, *d2 = nullptr, *ctl = nullptr;
  uint64_t *v1 = nullptr, *v2 = nullptr;
  const uint64_t ctl_words = PP + (two ? ((C + 255) / 256 * 256 + 256) : 0) + kGrFlagWords + 4;
  OLAP_TRY(sc.get<uint32_t>(P * cap, &d2));
  OLAP_TRY(sc.get<uint64_t>(P * cap, &v2));
  if (two) { OLAP_TRY(sc.get<uint32_t>(C * cap1, &d1)); OLAP_TRY(sc.get<uint64_t>(C * cap1, &v1)); }
  OLAP_TRY(sc.get<uint32_t>(ctl_words, &ctl));
  OLAP_CUDA(cudaMemsetAsync(ctl, 0, ctl_words * 4, s));
  uint32_t *cur2 = ctl, *cur1 = ctl + PP, *flags = ctl + ctl_words - kGrFlagWords - 4, *ticket = flags + kGrFlagWords;
  // Tile shape of the partition passes.  ncu on the 16-rows-per-thread kernel (profiles/r02_ncu_groupby_details.txt: 2^30 rows,
  // G = 1e5) shows a latency-bound pass -- 3.6 TB/s, L1TEX 40 % busy, issue 27 %, 37 % of the warp slots occupied (80 registers: 3
  // CTAs per SM) -- whose phases (load, rank, scan, stage, write-out) are separated by barriers, so only other CTAs fill the gaps.
  // 8 rows per thread halve the registers held across the barriers (48: 5 CTAs per SM).  Measured on 2^30 rows (B200): a pass that
  // fans out to few partitions gains a fifth (3 partitions 8.9 -> 7.1 ms, 25: 8.2 -> 6.7, 96: 8.5 -> 7.1); one that fans out to ~250
  // loses (8.7 -> 9.3, second level 6.8 -> 9.2 ms) because a 2048-row tile leaves 8-row runs and twice the run reservations.  So the
  // shape follows the pass's fan-out; OLAP_GR_ROWS = 8 / 16 forces one.
  const char* rv = getenv("OLAP_GR_ROWS");
  const int forced_rows = rv && *rv ? atoi(rv) : 0;
  auto launch_gr_part = [&](bool regions, const GrPartArgs& ga, uint64_t fan_out, uint64_t n_tiles_of_4096) -> cudaError_t {
    const bool r8 = forced_rows ? forced_rows == 8 : fan_out <= 128;
    const size_t smem = (size_t)kGrThreads * (r8 ? 8 : 16) * 12;
    const unsigned grid = capped_grid(n_tiles_of_4096 * (r8 ? 2 : 1), (uint64_t)ctx->sm_count * (r8 ? 5 : 3));
#define GR_LAUNCH(REG, RW)                                                                                                        \
  {                                                                                                                               \
    cudaError_t e = cudaFuncSetAttribute(gr_part_kernel<REG, RW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);       \
    if (e != cudaSuccess) return e;                                                                                               \
    gr_part_kernel<REG, RW><<<grid, kGrThreads, smem, s>>>(ga);                                                                   \
  }
    if (regions) { if (r8) GR_LAUNCH(true, 8) else GR_LAUNCH(true, 16) }
    else { if (r8) GR_LAUNCH(false, 8) else GR_LAUNCH(false, 16) }
#undef GR_LAUNCH
    return cudaGetLastError();
  };
  GrPartArgs a;
  memset(&a, 0, sizeof(a));
  a.keys = p.keys; a.vals = p.n_vcols ? p.vcol[0] : nullptr; a.n = p.n; a.rmin = (uint64_t)rmin; a.rrange = rrange; a.flags = flags;
  const uint64_t tiles0 = (p.n + 4095) / 4096;
  if (!two) {
    a.shift = kGrKeysLog2; a.cap = cap; a.cursor = cur2; a.out_d = d2; a.out_v = v2;
    OLAP_CUDA(launch_gr_part(false, a, P, tiles0));
    count_launch(ctx, 2);
  } else {
    a.shift = kGrKeysLog2 + 8; a.cap = cap1; a.cursor = cur1; a.out_d = d1; a.out_v = v1;
    OLAP_CUDA(launch_gr_part(false, a, C, tiles0));
    phase_mark(ctx, s, "partition_coarse");
    GrPartArgs f;
    memset(&f, 0, sizeof(f));
    f.in_d = d1; f.in_v = v1; f.reg_cur = cur1; f.reg_cap = cap1; f.n_reg = (uint32_t)C; f.reg_parts = kGrThreads;
    f.shift = kGrKeysLog2; f.cap = cap; f.cursor = cur2; f.out_d = d2; f.out_v = v2; f.flags = flags;
    const uint64_t tiles1 = (uint64_t)((cap1 + 4095) / 4096) * C;
    OLAP_CUDA(launch_gr_part(true, f, kGrThreads, tiles1));
    count_launch(ctx, 3);
  }
  phase_mark(ctx, s, "partition");
  GrAggArgs g;
  memset(&g, 0, sizeof(g));
  g.d = d2; g.v = v2; g.cur = cur2; g.cap = cap; g.P = (uint32_t)P; g.ticket = ticket;
  g.n_cells = p.n_cells;
  for (int c = 0; c < kMaxCells; c++) g.cell_op[c] = p.cell_op[c];
  g.out = p.out; g.out_keys = p.out_keys; g.out_aggs = p.out_aggs; g.out_cap = p.out_cap; g.out_n = p.out_n; g.rmin = rmin;
  const size_t asmem = (size_t)kGrKeys * (4 + 8 * (size_t)p.n_cells);
  OLAP_CUDA(cudaFuncSetAttribute(gr_agg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)asmem));
  const uint64_t ctas = (uint64_t)ctx->sm_count * 2;
  g.S = P >= ctas ? 1u : (uint32_t)((ctas + P - 1) / P);
  if (g.S > 1) {
    const uint64_t slots = P << kGrKeysLog2;
    OLAP_TRY(sc.get<unsigned long long>(slots, &g.m_cnt));
    OLAP_TRY(sc.get<uint64_t>(slots * std::max(p.n_cells, 1), &g.m_cells));
    gr_merge_init_kernel<<<capped_grid((slots + 255) / 256, (uint64_t)ctx->sm_count * 8), 256, 0, s>>>(g);
  }
  gr_agg_kernel<<<capped_grid(P * g.S, ctas), kGrAggThreads, asmem, s>>>(g);
  if (g.S > 1) gr_emit_kernel<<<capped_grid(((P << kGrKeysLog2) + 255) / 256, (uint64_t)ctx->sm_count * 8), 256, 0, s>>>(g);
  count_launch(ctx, 3);
  OLAP_CUDA(cudaGetLastError());
  uint32_t hf[kGrFlagWords];
  OLAP_CUDA(cudaMemcpyAsync(hf, flags, sizeof(hf), cudaMemcpyDeviceToHost, s));
  OLAP_CUDA(cudaStreamSynchronize(s));
  *flags_out = 0;
  for (int i = 0; i < kGrFlagWords; i++) if (hf[i]) *flags_out |= 1u << i;
  return OLAP_OK;
}

}  // namespace olap

using namespace olap;

static olap_status group_by_impl(olap_ctx* ctx, const olap_expr_prog* key_prog, const int64_t* keys, const olap_agg_spec* specs, const olap_expr_prog* const* progs, int32_t n_aggs,
                                 uint64_t n_rows, uint64_t n_groups_hint, const olap_filter_term* terms, int32_t n_terms, int64_t* out_keys,
                                 void* out_aggs, uint64_t out_capacity, uint64_t* out_n_groups_dptr, uint64_t* host_n_groups, void* stream) {
  OLAP_REQUIRE(ctx, "ctx is NULL");
  OLAP_REQUIRE(n_aggs == 0 || specs, "specs is NULL");
  OLAP_REQUIRE(out_capacity == 0 || (out_keys && (n_aggs == 0 || out_aggs)), "output pointer is NULL");
  OLAP_REQUIRE(n_rows < 0xFFFFFFFFULL || !keys, "grouped aggregation partitions with u32 offsets: n_rows must be < 2^32 - 1");
  std::lock_guard<std::mutex> g(ctx->mu);
  OLAP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream_of(stream);

  GbParams p;
  memset(&p, 0, sizeof(p));
  OLAP_TRY(build_cells(specs, progs, n_aggs, &p));
  OLAP_TRY(pack_terms(terms, n_terms, &p.terms));
  bool any_prog = false;
  p.prog_vec = true;
  if (key_prog) {  // computed Int64 group key: every "keys ?" below reads "grouped ?", and the PROG kernels never load through it
    OLAP_REQUIRE(!keys, "pass either a key column or a key program");
    OLAP_TRY(convert_prog(key_prog, false, &p.key_prog));
    for (int k = 0; k < p.key_prog.n; k++)
      if (p.key_prog.op[k] == 'c') p.prog_vec = p.prog_vec && (((uintptr_t)p.key_prog.col[k]) & 31) == 0;
    keys = reinterpret_cast<const int64_t*>(uintptr_t(0x100));
    any_prog = true;
  }
  p.keys = (const uint64_t*)keys;
  p.n = n_rows;
  p.vec = (((uintptr_t)keys) & 31) == 0;
  for (int q = 0; q < p.n_vcols; q++) {
    if (p.prog[q].n == 0) { p.vec = p.vec && (((uintptr_t)p.vcol[q]) & 31) == 0; continue; }
    any_prog = true;
    for (int k = 0; k < p.prog[q].n; k++)
      if (p.prog[q].op[k] == 'c') p.prog_vec = p.prog_vec && (((uintptr_t)p.prog[q].col[k]) & 31) == 0;
  }
  p.terms_vec = true;
  for (int k = 0; k < p.terms.n; k++)
    p.terms_vec = p.terms_vec && (((uintptr_t)p.terms.lhs[k]) & 31) == 0 && (!p.terms.rhs[k] || (((uintptr_t)p.terms.rhs[k]) & 31) == 0);
  if (!keys) n_groups_hint = 1;
  if (n_groups_hint < 1) n_groups_hint = 1;

  Scratch sc(ctx, s);
  unsigned long long* d_n = nullptr;
  uint32_t* d_overflow = nullptr;
  OLAP_TRY(sc.get<unsigned long long>(1, &d_n));
  OLAP_TRY(sc.get<uint32_t>(1, &d_overflow));
  p.out_keys = out_keys; p.out_aggs = (uint64_t*)out_aggs; p.out_cap = out_capacity; p.out_n = d_n;
  p.g_overflow = d_overflow;
  phase_begin(ctx, s);

  const uint64_t tile = (uint64_t)kGbThreads * kGbRows;
  const uint64_t tiles = (n_rows + tile - 1) / tile;
  uint64_t n_groups = 0;
  bool dense_failed = false;
  for (int attempt = 0;; attempt++) {
    Scratch tsc(ctx, s);
    // regime
    const uint64_t sbytes = slot_bytes(p), budget = 96 * 1024;  // shared regime: two CTAs per SM (a 50 KB budget / 4 CTAs measured no faster)
    uint64_t need = pow2_at_least((n_groups_hint * 4 + 2) / 3);
    if (need < 16) need = 16;
    // up to 96 KB: two CTAs per SM; up to 200 KB: one CTA per SM, still ahead of the partitioned regime's extra pass
    const bool shared = keys && attempt == 0 && (need + 1) * sbytes + 16 <= 200 * 1024;
    // dense keys (key range <= 16 x the hint, <= 2^28): direct-address tables over key ranges, no hashing (run_dense_range)
    bool dense = false;
    int64_t dmin = 0; uint64_t drange = 0;
    if (keys && attempt == 0 && !dense_failed && !shared && !any_prog && p.terms.n == 0 && p.n_vcols <= 1 && n_rows >= (1u << 20) &&
        getenv("OLAP_GB_DENSE") == nullptr) {
      int64_t mn, mx;
      OLAP_TRY(key_min_max(ctx, (const uint64_t*)keys, n_rows, s, &mn, &mx));
      const uint64_t span = (uint64_t)mx - (uint64_t)mn;
      if (mx >= mn && span < (1ULL << 28) && span + 1 <= 16 * std::max<uint64_t>(n_groups_hint, 1)) { dense = true; dmin = mn; drange = span + 1; }
    }
    const bool partitioned = keys && attempt == 0 && !shared && !dense && n_rows >= (1u << 20) && p.n_vcols <= 4;  // the partitioner carries <= 4 payload columns
    if (any_prog && partitioned) {
      set_error("computed aggregate inputs are fused in the no-key / shared / global regimes only; materialise them for %llu groups",
                (unsigned long long)n_groups_hint);
      return OLAP_ERR_UNSUPPORTED;
    }
    // global table: the spill target of the shared/partitioned regimes (small), or the whole table
    uint64_t gh = (shared || partitioned || dense) ? std::min<uint64_t>(n_groups_hint, 4096) : n_groups_hint;
    p.g_cap = keys ? pow2_at_least(std::max<uint64_t>(2 * gh, 1024)) : 1;
    const uint64_t n1 = p.g_cap + 1;
    OLAP_TRY(tsc.get<uint64_t>(n1, &p.g_keys));
    OLAP_TRY(tsc.get<unsigned long long>(n1, &p.g_count));
    OLAP_TRY(tsc.get<uint64_t>(n1 * (p.n_cells ? p.n_cells : 1), &p.g_cells));
    OLAP_CUDA(cudaMemsetAsync(d_n, 0, 8, s));
    OLAP_CUDA(cudaMemsetAsync(d_overflow, 0, 4, s));
    const unsigned tgrid = capped_grid((n1 + 255) / 256, (uint64_t)ctx->sm_count * 8);
    gb_init_kernel<<<tgrid, 256, 0, s>>>(p);
    count_launch(ctx);
    phase_mark(ctx, s, "init");
    if (n_rows > 0) {
      if (!keys) {
        if (any_prog) { GB_DISPATCH_NV(p.n_vcols, (gb_nokey_kernel<NV, true><<<capped_grid((tiles + 1) / 2, (uint64_t)ctx->sm_count * 4), kGbThreads, 0, s>>>(p))); }
        else { GB_DISPATCH_NV(p.n_vcols, (gb_nokey_kernel<NV><<<capped_grid((tiles + 1) / 2, (uint64_t)ctx->sm_count * 4), kGbThreads, 0, s>>>(p))); }
        count_launch(ctx);
      } else if (shared) {
        uint64_t copies = 1;
        while (copies < 64 && (need + 1) * sbytes * copies * 2 + 16 <= budget) copies <<= 1;
        p.s_slots = (uint32_t)need; p.s_copies = (uint32_t)copies; p.s_hskip = 0; p.s_hshift = 64 - ilog2_u64(need);
        const size_t smem = (size_t)((need + 1) * copies * sbytes + 16);
        if (any_prog) {
          GB_DISPATCH_NV(p.n_vcols, {
            OLAP_CUDA((cudaFuncSetAttribute(gb_shared_kernel<NV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
            gb_shared_kernel<NV, true><<<capped_grid(tiles, (uint64_t)ctx->sm_count * (smem <= budget ? 2 : 1)), kGbThreads, smem, s>>>(p);
          });
        } else {
          GB_DISPATCH_NV(p.n_vcols, {
            OLAP_CUDA(cudaFuncSetAttribute(gb_shared_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            gb_shared_kernel<NV><<<capped_grid(tiles, (uint64_t)ctx->sm_count * (smem <= budget ? 2 : 1)), kGbThreads, smem, s>>>(p);
          });
        }
        count_launch(ctx);
      } else if (dense) {
        uint32_t df = 0;
        OLAP_TRY(run_dense_range(ctx, p, dmin, drange, tsc, s, &df));
        if (df) {  // skewed keys overflowed a range's region, or the remembered key range was stale: redo without this regime
          if (df & (1u << kGrFlagOutOfRange)) ctx->shard_key_stats.erase(keys);
          else dense_failed = true;
          attempt = -1;
          continue;
        }
      } else if (partitioned) {
        OLAP_TRY(run_partitioned(ctx, p, n_groups_hint, tsc, s));
      } else {
        if (any_prog) { GB_DISPATCH_NV(p.n_vcols, (gb_global_kernel<NV, true><<<capped_grid(tiles, (uint64_t)ctx->sm_count * 4), kGbThreads, 0, s>>>(p, 0, n_rows))); }
        else { GB_DISPATCH_NV(p.n_vcols, (gb_global_kernel<NV><<<capped_grid(tiles, (uint64_t)ctx->sm_count * 4), kGbThreads, 0, s>>>(p, 0, n_rows))); }
        count_launch(ctx);
      }
      OLAP_CUDA(cudaGetLastError());
    }
    phase_mark(ctx, s, "aggregate");
    gb_extract_kernel<<<tgrid, 256, 0, s>>>(p, /*emit_empty_global=*/keys == nullptr);
    count_launch(ctx);
    OLAP_CUDA(cudaGetLastError());
    phase_mark(ctx, s, "extract");
    uint32_t h_overflow = 0;
    OLAP_CUDA(cudaMemcpyAsync(&h_overflow, d_overflow, 4, cudaMemcpyDeviceToHost, s));
    OLAP_CUDA(cudaMemcpyAsync(&n_groups, d_n, 8, cudaMemcpyDeviceToHost, s));
    OLAP_CUDA(cudaStreamSynchronize(s));
    if (!h_overflow) break;
    // the caller's hint was too small for the spill/global table: redo everything on one big global table
    n_groups_hint = std::max<uint64_t>(n_groups_hint * 8, 1u << 16);
    if (n_groups_hint > n_rows) n_groups_hint = n_rows ? n_rows : 1;
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

extern "C" olap_status olap_group_by_i64(olap_ctx* ctx, const int64_t* keys, const olap_agg_spec* specs, int32_t n_aggs, uint64_t n_rows,
                                         uint64_t n_groups_hint, const olap_filter_term* terms, int32_t n_terms, int64_t* out_keys,
                                         void* out_aggs, uint64_t out_capacity, uint64_t* out_n_groups_dptr, uint64_t* host_n_groups,
                                         void* stream) {
  return group_by_impl(ctx, nullptr, keys, specs, nullptr, n_aggs, n_rows, n_groups_hint, terms, n_terms, out_keys, out_aggs, out_capacity,
                       out_n_groups_dptr, host_n_groups, stream);
}

extern "C" olap_status olap_group_by_prog_i64(olap_ctx* ctx, const int64_t* keys, const olap_agg_spec* specs, const olap_expr_prog* const* progs,
                                              int32_t n_aggs, uint64_t n_rows, uint64_t n_groups_hint, const olap_filter_term* terms,
                                              int32_t n_terms, int64_t* out_keys, void* out_aggs, uint64_t out_capacity,
                                              uint64_t* out_n_groups_dptr, uint64_t* host_n_groups, void* stream) {
  return group_by_impl(ctx, nullptr, keys, specs, progs, n_aggs, n_rows, n_groups_hint, terms, n_terms, out_keys, out_aggs, out_capacity,
                       out_n_groups_dptr, host_n_groups, stream);
}

extern "C" olap_status olap_group_by_expr_i64(olap_ctx* ctx, const olap_expr_prog* key_prog, const olap_agg_spec* specs, const olap_expr_prog* const* progs,
                                              int32_t n_aggs, uint64_t n_rows, uint64_t n_groups_hint, const olap_filter_term* terms,
                                              int32_t n_terms, int64_t* out_keys, void* out_aggs, uint64_t out_capacity,
                                              uint64_t* out_n_groups_dptr, uint64_t* host_n_groups, void* stream) {
  OLAP_REQUIRE(key_prog, "key_prog is NULL (olap_group_by_prog_i64 takes a key column)");
  return group_by_impl(ctx, key_prog, nullptr, specs, progs, n_aggs, n_rows, n_groups_hint, terms, n_terms, out_keys, out_aggs, out_capacity,
                       out_n_groups_dptr, host_n_groups, stream);
}
