// Shared host/device helpers of the B200 execution core.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/olap_b200.h"

namespace olap {

constexpr int kSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

// --------------------------------------------------------------------------------------------- errors
void set_error(const char* fmt, ...);
olap_status cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define OLAP_CUDA(expr)                                                        \
  do {                                                                         \
    cudaError_t _e = (expr);                                                   \
    if (_e != cudaSuccess) return ::olap::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)
#define OLAP_TRY(expr)                         \
  do {                                         \
    olap_status _s = (expr);                   \
    if (_s != OLAP_OK) return _s;              \
  } while (0)
#define OLAP_REQUIRE(cond, msg)                                  \
  do {                                                           \
    if (!(cond)) { ::olap::set_error("bad argument: %s", msg); return OLAP_ERR_BAD_ARG; } \
  } while (0)

}  // namespace olap

// --------------------------------------------------------------------------------------------- context
struct olap_ctx {
  int device = 0;
  int sm_count = olap::kSMs;
  size_t smem_optin = 0;
  cudaMemPool_t pool = nullptr;
  std::vector<cudaStream_t> streams;
  std::mutex mu;
  std::unordered_map<void*, uint64_t> live;  // pool allocations -> bytes
  uint64_t used = 0, high_water = 0, max_bytes = 0;
  uint64_t launches = 0;
  // profiling of the last multi-phase operator
  bool profiling = false;
  std::vector<std::pair<std::string, cudaEvent_t>> marks;
  std::string phase_names;
  std::vector<float> phase_ms;
  // cached scratch that survives across calls (hash-table arena reused by the join, zeroed once)
  void* table_arena = nullptr;
  uint64_t table_arena_bytes = 0;
  uint32_t table_generation = 0;
  cudaEvent_t fork_ev = nullptr, join_ev[8] = {};
  cudaEvent_t arena_ev = nullptr;   // recorded behind every use of table_arena; the next user (on any stream) waits for it first
  bool arena_used = false;
  // column statistics remembered per (device pointer, rows): min / max of a join's build keys (verified by the build pass
  // itself: a key outside the remembered range raises a flag and the range is recomputed) and whether a probe column needed
  // heavy-hitter handling
  struct KeyStats { uint64_t n = 0; int64_t min = 0, max = 0; bool have_range = false; bool skewed = false; };
  std::unordered_map<const void*, KeyStats> key_stats;
  std::unordered_map<const void*, KeyStats> shard_key_stats;  // per-rank shards of a multi-GPU join (shard.cu)
  // operator layer state (in-memory catalog of device-resident tables), owned by operators.cu
  void* op_state = nullptr;
  void (*op_state_free)(void*) = nullptr;

  cudaStream_t stream_of(void* s) const { return s ? (cudaStream_t)s : streams[0]; }
};

namespace olap {

// RAII scratch allocation from the context's stream-ordered pool (freed on the same stream).
olap_status pool_alloc(olap_ctx* ctx, uint64_t bytes, cudaStream_t s, void** out);
olap_status pool_free(olap_ctx* ctx, void* p, cudaStream_t s);

struct Scratch {
  olap_ctx* ctx; cudaStream_t s; std::vector<void*> ptrs;
  Scratch(olap_ctx* c, cudaStream_t st) : ctx(c), s(st) {}
  ~Scratch() { for (void* p : ptrs) pool_free(ctx, p, s); }
  template <class T> olap_status get(uint64_t count, T** out) {
    void* p = nullptr;
    olap_status st = pool_alloc(ctx, count * sizeof(T) + 16, s, &p);
    if (st == OLAP_OK) { ptrs.push_back(p); *out = (T*)p; }
    return st;
  }
};

void phase_begin(olap_ctx* ctx, cudaStream_t s);
void phase_mark(olap_ctx* ctx, cudaStream_t s, const char* name);
void phase_end(olap_ctx* ctx, cudaStream_t s);

inline void count_launch(olap_ctx* ctx, uint64_t n = 1) { ctx->launches += n; }

// --------------------------------------------------------------------------------------------- device helpers
// murmur3 fmix64: a bijection on 64-bit words.  High bits pick the radix partition, low bits the table slot.
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return x;
}

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t seed, uint64_t i) {
  uint64_t z = seed + (i + 1) * 0x9e3779b97f4a7c15ULL;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}

// Streaming loads: read-only path, no L1 allocation.  The 256-bit form (sm_100a LDG.E.256) additionally marks the
// line evict-first in L2 so that scans do not push the L2-resident hash tables / partition write fronts out of
// the 126 MB L2 (ptxas only accepts .L2::evict_first on 256-bit accesses).  Needs 32-byte alignment.
struct u64x4 { uint64_t a, b, c, d; };
__device__ __forceinline__ u64x4 ld_stream_u64x4(const void* p) {
  u64x4 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.u64 {%0, %1, %2, %3}, [%4];"
               : "=l"(v.a), "=l"(v.b), "=l"(v.c), "=l"(v.d) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream_u64x4(void* p, u64x4 v) {
  asm volatile("st.global.L1::no_allocate.v4.u64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(v.a), "l"(v.b), "l"(v.c), "l"(v.d) : "memory");
}
__device__ __forceinline__ ulonglong2 ld_stream_u64x2(const void* p) {
  ulonglong2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ uint64_t ld_stream_u64(const void* p) {
  uint64_t v;
  asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint4 ld_stream_u32x4(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t ld_stream_u32(const void* p) {
  uint32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
// Streaming stores: written once, not re-read by this kernel.
__device__ __forceinline__ void st_stream_u64(void* p, uint64_t v) {
  asm volatile("st.global.L1::no_allocate.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_stream_u32(void* p, uint32_t v) {
  asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_stream_u64x2(void* p, ulonglong2 v) {
  asm volatile("st.global.L1::no_allocate.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(v.x), "l"(v.y) : "memory");
}

// comparison shared by filter / fused kernels; op outside EQ..GE matches nothing (executor.rs:464)
template <class T>
__device__ __forceinline__ bool cmp_op(T l, T r, int op) {
  switch (op) {
    case OLAP_EQ: return l == r;
    case OLAP_NE: return l != r;
    case OLAP_LT: return l < r;
    case OLAP_LE: return l <= r;
    case OLAP_GT: return l > r;
    case OLAP_GE: return l >= r;
    default: return false;
  }
}

// Filter conjunction passed by value to kernels (up to 8 terms).
constexpr int kMaxTerms = 8;
struct TermPack {
  const uint64_t* lhs[kMaxTerms];
  const uint64_t* rhs[kMaxTerms];
  int64_t lit[kMaxTerms];
  int8_t op[kMaxTerms];
  int8_t type[kMaxTerms];
  int n;
};

__device__ __forceinline__ bool eval_term(const TermPack& t, int k, uint64_t l, uint64_t r) {
  if (t.type[k] == OLAP_F64) return cmp_op<double>(__longlong_as_double((long long)l), __longlong_as_double((long long)r), t.op[k]);
  return cmp_op<long long>((long long)l, (long long)r, t.op[k]);
}
__device__ __forceinline__ bool eval_terms_row(const TermPack& t, uint64_t row) {
  bool pass = true;
#pragma unroll 1
  for (int k = 0; k < t.n; k++) {
    uint64_t l = ld_stream_u64(t.lhs[k] + row);
    uint64_t r = t.rhs[k] ? ld_stream_u64(t.rhs[k] + row) : (uint64_t)t.lit[k];
    pass = pass && eval_term(t, k, l, r);
  }
  return pass;
}

// The comparison operator is dispatched once per 4 (or 8) rows, not per row: the switch sits outside the row loop.
template <class T, int N>
__device__ __forceinline__ void and_cmp(int op, const T (&l)[N], const T (&r)[N], bool (&f)[N]) {
  switch (op) {
    case OLAP_EQ:
#pragma unroll
      for (int i = 0; i < N; i++) f[i] = f[i] && l[i] == r[i];
      break;
    case OLAP_NE:
#pragma unroll
      for (int i = 0; i < N; i++) f[i] = f[i] && l[i] != r[i];
      break;
    case OLAP_LT:
#pragma unroll
      for (int i = 0; i < N; i++) f[i] = f[i] && l[i] < r[i];
      break;
    case OLAP_LE:
#pragma unroll
      for (int i = 0; i < N; i++) f[i] = f[i] && l[i] <= r[i];
      break;
    case OLAP_GT:
#pragma unroll
      for (int i = 0; i < N; i++) f[i] = f[i] && l[i] > r[i];
      break;
    case OLAP_GE:
#pragma unroll
      for (int i = 0; i < N; i++) f[i] = f[i] && l[i] >= r[i];
      break;
    default:  // executor.rs:464: an unknown operator matches no row
#pragma unroll
      for (int i = 0; i < N; i++) f[i] = false;
  }
}

// evaluates the conjunction on the 4 consecutive rows starting at row0 (a multiple of 4); rows >= n never pass
template <bool VEC>
__device__ __forceinline__ void eval_quad(const TermPack& terms, uint64_t row0, uint64_t n, bool (&f)[4]) {
#pragma unroll
  for (int e = 0; e < 4; e++) f[e] = row0 + e < n;
#pragma unroll 1
  for (int k = 0; k < terms.n; k++) {
    uint64_t l[4], r[4];
    if (VEC && row0 + 3 < n) {
      const u64x4 a = ld_stream_u64x4(terms.lhs[k] + row0);
      l[0] = a.a; l[1] = a.b; l[2] = a.c; l[3] = a.d;
      if (terms.rhs[k]) { const u64x4 b = ld_stream_u64x4(terms.rhs[k] + row0); r[0] = b.a; r[1] = b.b; r[2] = b.c; r[3] = b.d; }
      else { r[0] = r[1] = r[2] = r[3] = (uint64_t)terms.lit[k]; }
    } else {
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const bool in = row0 + e < n;
        l[e] = in ? ld_stream_u64(terms.lhs[k] + row0 + e) : 0;
        r[e] = (in && terms.rhs[k]) ? ld_stream_u64(terms.rhs[k] + row0 + e) : (uint64_t)terms.lit[k];
      }
    }
    if (terms.type[k] == OLAP_F64) {
      double dl[4], dr[4];
#pragma unroll
      for (int e = 0; e < 4; e++) { dl[e] = __longlong_as_double((long long)l[e]); dr[e] = __longlong_as_double((long long)r[e]); }
      and_cmp<double, 4>(terms.op[k], dl, dr, f);
    } else {
      long long il[4], ir[4];
#pragma unroll
      for (int e = 0; e < 4; e++) { il[e] = (long long)l[e]; ir[e] = (long long)r[e]; }
      and_cmp<long long, 4>(terms.op[k], il, ir, f);
    }
  }
}

olap_status pack_terms(const olap_filter_term* terms, int32_t n_terms, TermPack* out);


// eval_binary_expr (gpu-olap-core/src/executor.rs:411-441): i64 arithmetic wraps, x / 0 = 0, any other operator yields 0;
// f64 is this core's extension (IEEE)
__device__ __forceinline__ int64_t binop_i64(int64_t a, int64_t b, char op) {
  switch (op) {
    case '+': return (int64_t)((uint64_t)a + (uint64_t)b);
    case '-': return (int64_t)((uint64_t)a - (uint64_t)b);
    case '*': return (int64_t)((uint64_t)a * (uint64_t)b);
    case '/': return b == 0 ? 0 : ((a == INT64_MIN && b == -1) ? INT64_MIN : a / b);
    default: return 0;
  }
}
__device__ __forceinline__ double binop_f64(double a, double b, char op) {
  switch (op) {
    case '+': return a + b;
    case '-': return a - b;
    case '*': return a * b;
    case '/': return a / b;
    default: return 0.0;
  }
}


// A value column computed on the fly: arithmetic over 8-byte columns and literals of ONE type in reverse Polish notation,
// stack depth <= 3 (olap_expr_prog).  Lets the group-by kernels aggregate SUM(price * (1 - discount)) without the
// projection passes that would materialise every temporary as a column.
constexpr int kMaxProgOps = OLAP_PROG_MAX_OPS;
struct ExprProg {
  int8_t n;                        // 0: the value column is read from memory
  int8_t f64;
  int8_t op[kMaxProgOps];          // 'c' push column, 'l' push literal, '+' '-' '*' '/'
  const uint64_t* col[kMaxProgOps];
  uint64_t lit[kMaxProgOps];
};

// the program's value for the four consecutive rows starting at row0 (dead rows get an unspecified value)
template <bool VEC>
__device__ __forceinline__ void eval_prog4(const ExprProg& g, uint64_t row0, uint64_t begin, uint64_t end, uint64_t (&out)[4]) {
  uint64_t s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
  const bool whole = VEC && row0 >= begin && row0 + 3 < end;
#pragma unroll 1
  for (int k = 0; k < g.n; k++) {
    const int op = g.op[k];
    if (op == 'c' || op == 'l') {
#pragma unroll
      for (int e = 0; e < 4; e++) { s2[e] = s1[e]; s1[e] = s0[e]; }
      if (op == 'l') { s0[0] = s0[1] = s0[2] = s0[3] = g.lit[k]; }
      else if (whole) { const u64x4 v = ld_stream_u64x4(g.col[k] + row0); s0[0] = v.a; s0[1] = v.b; s0[2] = v.c; s0[3] = v.d; }
      else {
#pragma unroll
        for (int e = 0; e < 4; e++) s0[e] = (row0 + e >= begin && row0 + e < end) ? ld_stream_u64(g.col[k] + row0 + e) : 0;
      }
    } else {
      if (g.f64) {
#pragma unroll
        for (int e = 0; e < 4; e++)
          s0[e] = (uint64_t)__double_as_longlong(binop_f64(__longlong_as_double((long long)s1[e]), __longlong_as_double((long long)s0[e]), (char)op));
      } else {
#pragma unroll
        for (int e = 0; e < 4; e++) s0[e] = (uint64_t)binop_i64((int64_t)s1[e], (int64_t)s0[e], (char)op);
      }
#pragma unroll
      for (int e = 0; e < 4; e++) s1[e] = s2[e];
    }
  }
#pragma unroll
  for (int e = 0; e < 4; e++) out[e] = s0[e];
}

}  // namespace olap
