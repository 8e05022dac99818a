/*
 * oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of gpu-olap-engine's data-parallel hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * file (as the checker / the timed CPU baseline).  The product path (gpu-olap-engine_b200/csrc) never does.
 *
 * What is restated (citations are path:line under the reference checkout):
 *   filter       gpu-olap-core/src/executor.rs:443-476 (mask), :478-507 (gather)
 *   projection   gpu-olap-core/src/executor.rs:411-441
 *   hash         gpu-kernel-compiler/kernels/join_kernels.cuh:26-41
 *   partition    gpu-kernel-compiler/kernels/join_kernels.cuh:45-76
 *   build/probe  gpu-kernel-compiler/kernels/join_kernels.cuh:80-111, :115-166
 *   hash join    gpu-olap-core/src/executor.rs:238-253 (build) + the probe the reference describes at :236 but
 *                omits; match semantics are those of join_kernels.cuh:146-165
 *   aggregate    gpu-olap-core/src/executor.rs:308-346 (global, f64 accumulators), identities
 *                gpu-kernel-compiler/src/aggregate_kernel.rs:27-33, output layout :98
 *   direct-address SUM   gpu-kernel-compiler/kernels/join_kernels.cuh:216-240
 *   arrow layout arrow-interop/src/record_batch_convert.rs:102-130 (widening, FNV-1a)
 *
 * Parity status: the reference's Rust cannot be built here (no cargo/rustc, crates un-vendored, and the
 * workspace has source-level compile errors -- see SURVEY.md F10), and its own tests hold exactly one data
 * vector for this path (the filter mask at executor.rs:553-570) plus boundary constants.  Those are pinned in
 * tests/test_oracle_golden.py.  The CUDA header join_kernels.cuh DOES compile; oracle/ref_driver.cu builds it
 * (from where it lies, into oracle/_ref/) and the -m gpu tests run the reference's own kernels on the B200 to
 * pin hash / partition / build / probe / direct-address-SUM semantics.  Grouped aggregation beyond that is a
 * completion of an unfinished stub (executor.rs:307) along the rules in SURVEY.md 8(c): "parity unpinned" for
 * GROUP BY keys, MIN/MAX/AVG and f64 -- there is no reference execution to compare with.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <pthread.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------------------- */
/* Synthetic data (counter-based so any shard can generate its slice).                     */
/* ------------------------------------------------------------------------------------------------------- */
ORC_API uint64_t orc_splitmix64(uint64_t seed, uint64_t i) {
  uint64_t z = seed + (i + 1) * 0x9e3779b97f4a7c15ULL;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}

static inline uint64_t perm_round(uint64_t x, int bits, uint64_t seed) {
  const uint64_t mask = bits == 64 ? ~0ULL : ((1ULL << bits) - 1);
  static const uint64_t M[4] = {0x9e3779b97f4a7c15ULL, 0xbf58476d1ce4e5b9ULL, 0x94d049bb133111ebULL, 0xff51afd7ed558ccdULL};
  for (int r = 0; r < 4; r++) {
    x = (x * M[r] + orc_splitmix64(seed, (uint64_t)r)) & mask; /* odd multiplier: bijection mod 2^bits */
    x ^= x >> ((bits >> 1) + (r & 1) + (bits < 2));             /* xor-shift-right by >=1: bijection       */
  }
  return x;
}

/* bijection on [0, domain): cycle-walk the power-of-two permutation */
ORC_API uint64_t orc_perm(uint64_t x, uint64_t domain, uint64_t seed) {
  int bits = 1;
  while (bits < 64 && (1ULL << bits) < domain) bits++;
  do { x = perm_round(x, bits, seed); } while (x >= domain);
  return x;
}

ORC_API void orc_gen_uniform_i64(int64_t* out, uint64_t n, uint64_t seed, uint64_t row0, uint64_t domain) {
  for (uint64_t i = 0; i < n; i++) out[i] = (int64_t)(orc_splitmix64(seed, row0 + i) % domain);
}
ORC_API void orc_gen_perm_i64(int64_t* out, uint64_t n, uint64_t seed, uint64_t row0, uint64_t domain) {
  for (uint64_t i = 0; i < n; i++) out[i] = (int64_t)orc_perm(row0 + i, domain, seed);
}
/* dyadic f64: (u % domain) / 1024 -- every partial sum exactly representable (SURVEY.md 8(d) f64 protocol) */
ORC_API void orc_gen_dyadic_f64(double* out, uint64_t n, uint64_t seed, uint64_t row0, uint64_t domain) {
  for (uint64_t i = 0; i < n; i++) out[i] = (double)(orc_splitmix64(seed, row0 + i) % domain) / 1024.0;
}
/* uniform f64 in [0,1) with 53 random bits */
ORC_API void orc_gen_unit_f64(double* out, uint64_t n, uint64_t seed, uint64_t row0) {
  for (uint64_t i = 0; i < n; i++) out[i] = (double)(orc_splitmix64(seed, row0 + i) >> 11) * (1.0 / 9007199254740992.0);
}

/* ------------------------------------------------------------------------------------------------------- */
/* Filter: executor.rs:443-476.  op codes follow CmpOp's declaration order (codegen.rs:10).                 */
/* ------------------------------------------------------------------------------------------------------- */
enum { ORC_EQ = 0, ORC_NE = 1, ORC_LT = 2, ORC_LE = 3, ORC_GT = 4, ORC_GE = 5 };

static inline uint8_t cmp_i64(int64_t l, int64_t r, int op) {
  switch (op) {
    case ORC_EQ: return l == r;
    case ORC_NE: return l != r;
    case ORC_LT: return l < r;
    case ORC_LE: return l <= r;
    case ORC_GT: return l > r;
    case ORC_GE: return l >= r;
    default: return 0; /* executor.rs:464 -- unknown operator maps every row to false */
  }
}
/* rhs == NULL: compare with the broadcast literal (executor.rs:390-394 materialises it as a column) */
ORC_API void orc_filter_mask_i64(const int64_t* lhs, const int64_t* rhs, int64_t literal, uint64_t n, int op, uint8_t* mask) {
  for (uint64_t i = 0; i < n; i++) mask[i] = cmp_i64(lhs[i], rhs ? rhs[i] : literal, op);
}
/* non-Int64 operands: executor.rs:471-472 passes every row */
ORC_API void orc_filter_mask_all(uint64_t n, uint8_t* mask) { memset(mask, 1, n); }
/* conjunction completion rule: AND of the term masks, not the reference's all-false bug */
ORC_API void orc_mask_and(uint8_t* acc, const uint8_t* other, uint64_t n) {
  for (uint64_t i = 0; i < n; i++) acc[i] &= other[i];
}
/* executor.rs:478-507: kept indices in row order, then gather 8-byte values (Int64 and Float64 alike) */
ORC_API uint64_t orc_mask_to_indices(const uint8_t* mask, uint64_t n, uint64_t* idx) {
  uint64_t k = 0;
  for (uint64_t i = 0; i < n; i++) if (mask[i]) idx[k++] = i;
  return k;
}
ORC_API void orc_take_8b(const uint64_t* col, const uint64_t* idx, uint64_t k, uint64_t* out) {
  for (uint64_t i = 0; i < k; i++) out[i] = col[idx[i]];
}

/* ------------------------------------------------------------------------------------------------------- */
/* Projection: executor.rs:411-441.  Release-mode Rust wraps; x/0 -> 0; unknown op -> 0.                     */
/* ------------------------------------------------------------------------------------------------------- */
ORC_API void orc_project_i64(const int64_t* l, const int64_t* r, int64_t literal, uint64_t n, char op, int64_t* out) {
  for (uint64_t i = 0; i < n; i++) {
    uint64_t a = (uint64_t)l[i], b = (uint64_t)(r ? r[i] : literal);
    uint64_t v;
    switch (op) {
      case '+': v = a + b; break;
      case '-': v = a - b; break;
      case '*': v = a * b; break;
      case '/':
        if (b == 0) v = 0;
        else if ((int64_t)a == INT64_MIN && (int64_t)b == -1) v = (uint64_t)INT64_MIN; /* wrapping_div */
        else v = (uint64_t)((int64_t)a / (int64_t)b);
        break;
      default: v = 0;
    }
    out[i] = (int64_t)v;
  }
}

/* ------------------------------------------------------------------------------------------------------- */
/* Reference hash functions: join_kernels.cuh:26-41.                                                        */
/* ------------------------------------------------------------------------------------------------------- */
ORC_API uint32_t orc_hash_int32(int32_t key) {
  uint32_t h = (uint32_t)key;
  h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
  return h;
}
ORC_API uint32_t orc_hash_int64(int64_t key) {
  uint32_t h = (uint32_t)((uint64_t)key ^ (uint64_t)(key >> 32)); /* arithmetic shift, as the C++ does */
  return orc_hash_int32((int32_t)h);
}
ORC_API uint32_t orc_partition_id(int64_t key, int radix_shift) { return (orc_hash_int64(key) >> radix_shift) & 255u; }

/* join_kernels.cuh:45-76 with the histogram + exclusive scan the reference never wrote.  Within a partition
 * the kernel's order is whatever the atomics give; this is the serialisation tid = 0..n-1 (stable).         */
ORC_API void orc_radix_partition_i64(const int64_t* keys, const uint32_t* row_ids, uint32_t n, int radix_shift,
                                     uint32_t* part_begin /*[257]*/, int64_t* out_keys, uint32_t* out_ids) {
  uint32_t cur[256];
  memset(part_begin, 0, 257 * sizeof(uint32_t));
  for (uint32_t i = 0; i < n; i++) part_begin[orc_partition_id(keys[i], radix_shift) + 1]++;
  for (int p = 0; p < 256; p++) part_begin[p + 1] += part_begin[p];
  memcpy(cur, part_begin, sizeof(cur));
  for (uint32_t i = 0; i < n; i++) {
    uint32_t pos = cur[orc_partition_id(keys[i], radix_shift)]++;
    out_keys[pos] = keys[i];
    out_ids[pos] = row_ids ? row_ids[i] : i;
  }
}

/* join_kernels.cuh:8-12 */
typedef struct { uint32_t hash, row_id, next; } orc_hash_entry;

/* join_kernels.cuh:80-111, serialised tid = 0..n-1; buckets must arrive filled with 0xFFFFFFFF (:146) */
ORC_API void orc_ref_build(const int64_t* keys, const uint32_t* row_ids, uint32_t n, orc_hash_entry* entries,
                           uint32_t* buckets, uint32_t capacity) {
  for (uint32_t t = 0; t < n; t++) {
    uint32_t h = orc_hash_int64(keys[t]);
    uint32_t b = h % capacity;
    entries[t].hash = h;
    entries[t].row_id = row_ids ? row_ids[t] : t;
    entries[t].next = buckets[b];
    buckets[b] = t;
  }
}
/* join_kernels.cuh:115-166; build_keys is indexed by ORIGINAL row id (:152).  Returns the exact match count;
 * pairs beyond max_matches are dropped silently as in :157.                                                 */
ORC_API uint64_t orc_ref_probe(const int64_t* probe_keys, const uint32_t* probe_ids, uint32_t n_probe,
                               const orc_hash_entry* entries, const uint32_t* buckets, uint32_t capacity,
                               const int64_t* build_keys, uint32_t* left, uint32_t* right, uint64_t max_matches) {
  uint64_t count = 0;
  for (uint32_t t = 0; t < n_probe; t++) {
    int64_t pk = probe_keys[t];
    uint32_t h = orc_hash_int64(pk);
    for (uint32_t e = buckets[h % capacity]; e != 0xFFFFFFFFu; e = entries[e].next) {
      if (entries[e].hash == h && build_keys[entries[e].row_id] == pk) {
        if (count < max_matches) { left[count] = entries[e].row_id; right[count] = probe_ids ? probe_ids[t] : t; }
        count++;
      }
    }
  }
  return count;
}

/* join_kernels.cuh:216-240 for an integer T of 8 bytes: direct addressing, SUM only */
ORC_API void orc_ref_hash_aggregate_u64(const uint32_t* group_keys, const uint64_t* values, uint32_t n, uint64_t* aggregates,
                                        uint32_t num_groups) {
  for (uint32_t i = 0; i < n; i++) aggregates[group_keys[i] % num_groups] += values[i];
}

/* ------------------------------------------------------------------------------------------------------- */
/* Hash join as the executor lays it down: executor.rs:238-253 builds HashMap<i64, Vec<usize>> over the LEFT */
/* side with a global row offset across batches; the probe (right side) is described at :236.  Output order  */
/* here is probe-major with build rows in insertion order, i.e. already sorted by (right_id, left_id).       */
/* ------------------------------------------------------------------------------------------------------- */
static inline uint64_t mix_u64(uint64_t x) {
  x ^= x >> 32; x *= 0xd6e8feb86659fd93ULL; x ^= x >> 32; x *= 0xd6e8feb86659fd93ULL; x ^= x >> 32;
  return x;
}
typedef struct {
  uint64_t cap_mask;
  int64_t* keys;    /* distinct keys */
  uint32_t* head;   /* first build row of the key's list, +1 (0 = empty slot) */
  uint32_t* tail;
  uint32_t* next;   /* per build row: next row with the same key, +1 */
} orc_map;

static int map_init(orc_map* m, uint64_t n_build) {
  uint64_t cap = 16;
  while (cap < 2 * n_build) cap <<= 1;
  m->cap_mask = cap - 1;
  m->keys = (int64_t*)malloc(cap * sizeof(int64_t));
  m->head = (uint32_t*)calloc(cap, sizeof(uint32_t));
  m->tail = (uint32_t*)malloc(cap * sizeof(uint32_t));
  m->next = (uint32_t*)calloc(n_build ? n_build : 1, sizeof(uint32_t));
  return m->keys && m->head && m->tail && m->next;
}
static void map_free(orc_map* m) { free(m->keys); free(m->head); free(m->tail); free(m->next); }
static inline void map_push(orc_map* m, int64_t k, uint32_t row) { /* entry(k).or_default().push(row) */
  uint64_t s = mix_u64((uint64_t)k) & m->cap_mask;
  while (m->head[s] && m->keys[s] != k) s = (s + 1) & m->cap_mask;
  if (!m->head[s]) { m->keys[s] = k; m->head[s] = row + 1; }
  else m->next[m->tail[s] - 1] = row + 1;
  m->tail[s] = row + 1;
}
static inline uint32_t map_find(const orc_map* m, int64_t k) {
  uint64_t s = mix_u64((uint64_t)k) & m->cap_mask;
  while (m->head[s] && m->keys[s] != k) s = (s + 1) & m->cap_mask;
  return m->head[s];
}

ORC_API int orc_hash_join_i64(const int64_t* build, uint64_t n_build, const int64_t* probe, uint64_t n_probe,
                              uint32_t* out_left, uint32_t* out_right, uint64_t max_matches, uint64_t* out_count) {
  orc_map m;
  if (!map_init(&m, n_build)) return -1;
  for (uint64_t i = 0; i < n_build; i++) map_push(&m, build[i], (uint32_t)i);
  uint64_t count = 0;
  for (uint64_t j = 0; j < n_probe; j++) {
    for (uint32_t r = map_find(&m, probe[j]); r; r = m.next[r - 1]) {
      if (count < max_matches && out_left) { out_left[count] = r - 1; out_right[count] = (uint32_t)j; }
      count++;
    }
  }
  *out_count = count;
  map_free(&m);
  return 0;
}

/* ------------------------------------------------------------------------------------------------------- */
/* Aggregation.                                                                                             */
/* ------------------------------------------------------------------------------------------------------- */
enum { ORC_SUM = 0, ORC_COUNT = 1, ORC_MIN = 2, ORC_MAX = 3, ORC_AVG = 4 }; /* AggFunc order, aggregate_kernel.rs:17-25 */
enum { ORC_I64 = 0, ORC_F64 = 1 };

/* aggregate_kernel.rs:27-33 */
ORC_API int64_t orc_identity_value(int func) {
  switch (func) { case ORC_MIN: return INT64_MAX; case ORC_MAX: return INT64_MIN; default: return 0; }
}

/* executor.rs:308-346, one aggregate expression: state = {sum, count, min, max} as the reference keeps them.
 * Call orc_agg_global_init once, orc_agg_global_update per batch in batch order, orc_agg_global_final last.  */
typedef struct { double sum; uint64_t count; double min, max; } orc_global_state;
ORC_API void orc_agg_global_init(orc_global_state* s) { s->sum = 0.0; s->count = 0; s->min = DBL_MAX; s->max = -DBL_MAX; }
ORC_API void orc_agg_global_update(orc_global_state* s, const void* col, int type, uint64_t n) {
  for (uint64_t i = 0; i < n; i++) {
    double v = type == ORC_F64 ? ((const double*)col)[i] : (double)((const int64_t*)col)[i]; /* executor.rs:520-523 */
    s->count += 1;
    s->sum += v;
    if (v < s->min) s->min = v;
    if (v > s->max) s->max = v;
  }
}
ORC_API double orc_agg_global_final(const orc_global_state* s, int func) {
  switch (func) {
    case ORC_SUM: return s->sum;
    case ORC_COUNT: return (double)s->count;
    case ORC_MIN: return s->min;
    case ORC_MAX: return s->max;
    default: return s->count ? s->sum / (double)s->count : 0.0;
  }
}

/* Grouped completion: group key = flat int64 column (aggregate_kernel.rs:48-49); one output
 * row per distinct key, sorted by key; output agg-major, 8-byte slots, stride = n_groups (aggregate_kernel.rs:98).
 * Typed results: COUNT -> i64; SUM/MIN/MAX of i64 -> i64 (wrapping SUM); of f64 -> f64 accumulated in row
 * order; AVG -> f64 = sum/count (i64 input: exact i64 sum converted once).  compensated != 0 sums f64 in
 * long double (x87, 64-bit mantissa) -- the yardstick for the <= 1e-12 tolerance on random f64 data.         */
typedef struct { int64_t key; uint64_t count; } grp_hdr;
typedef union { int64_t i; double d; long double ld; } grp_acc;

typedef struct {
  uint64_t cap_mask, n_groups, n_aggs;
  uint8_t* used;
  grp_hdr* hdr;
  grp_acc* acc; /* [slot][agg][2]: primary accumulator, plus f64 running sum for AVG */
} grp_table;

static int grp_init(grp_table* t, uint64_t cap, uint64_t n_aggs) {
  t->cap_mask = cap - 1; t->n_groups = 0; t->n_aggs = n_aggs;
  t->used = (uint8_t*)calloc(cap, 1);
  t->hdr = (grp_hdr*)malloc(cap * sizeof(grp_hdr));
  t->acc = (grp_acc*)malloc(cap * (n_aggs ? n_aggs : 1) * sizeof(grp_acc));
  return t->used && t->hdr && t->acc;
}
static void grp_free(grp_table* t) { free(t->used); free(t->hdr); free(t->acc); }

static void acc_identity(grp_acc* a, int func, int type, int comp) {
  if (func == ORC_MIN) { if (type == ORC_F64) a->d = DBL_MAX; else a->i = INT64_MAX; }
  else if (func == ORC_MAX) { if (type == ORC_F64) a->d = -DBL_MAX; else a->i = INT64_MIN; }
  else if (type == ORC_F64 && (func == ORC_SUM || func == ORC_AVG)) { if (comp) a->ld = 0.0L; else a->d = 0.0; }
  else a->i = 0;
}
static inline void acc_update(grp_acc* a, int func, int type, int comp, const void* col, uint64_t row) {
  if (func == ORC_COUNT) return; /* header count */
  if (type == ORC_F64) {
    double v = ((const double*)col)[row];
    if (func == ORC_MIN) { if (v < a->d) a->d = v; }
    else if (func == ORC_MAX) { if (v > a->d) a->d = v; }
    else if (comp) a->ld += (long double)v;
    else a->d += v;
  } else {
    int64_t v = ((const int64_t*)col)[row];
    if (func == ORC_MIN) { if (v < a->i) a->i = v; }
    else if (func == ORC_MAX) { if (v > a->i) a->i = v; }
    else a->i = (int64_t)((uint64_t)a->i + (uint64_t)v);
  }
}
static uint64_t acc_final(const grp_acc* a, int func, int type, int comp, uint64_t count) {
  union { uint64_t u; int64_t i; double d; } r;
  if (func == ORC_COUNT) r.i = (int64_t)count;
  else if (func == ORC_AVG) {
    double s = type == ORC_F64 ? (comp ? (double)a->ld : a->d) : (double)a->i;
    if (type == ORC_F64 && comp) r.d = count ? (double)(a->ld / (long double)count) : 0.0;
    else r.d = count ? s / (double)count : 0.0;
  } else if (type == ORC_F64 && func == ORC_SUM) r.d = comp ? (double)a->ld : a->d;
  else if (type == ORC_F64) r.d = a->d;
  else r.i = a->i;
  return r.u;
}

static uint64_t grp_slot(grp_table* t, int64_t key, const int* funcs, const int* types, int comp) {
  uint64_t s = mix_u64((uint64_t)key) & t->cap_mask;
  while (t->used[s] && t->hdr[s].key != key) s = (s + 1) & t->cap_mask;
  if (!t->used[s]) {
    t->used[s] = 1; t->hdr[s].key = key; t->hdr[s].count = 0; t->n_groups++;
    for (uint64_t a = 0; a < t->n_aggs; a++) acc_identity(&t->acc[s * t->n_aggs + a], funcs[a], types[a], comp);
  }
  return s;
}

static int cmp_keyslot(const void* a, const void* b) {
  int64_t x = ((const int64_t*)a)[0], y = ((const int64_t*)b)[0];
  return x < y ? -1 : x > y;
}

/* Returns the number of groups (or (uint64_t)-1 on allocation failure / out_cap overflow).
 * keys == NULL: one global group with key 0, emitted even for n == 0 (executor.rs emits one row always).     */
ORC_API uint64_t orc_group_by_i64(const int64_t* keys, uint64_t n, uint64_t n_aggs, const void* const* value_cols,
                                  const int* types, const int* funcs, int compensated, uint64_t out_cap,
                                  int64_t* out_keys, uint64_t* out_aggs) {
  uint64_t cap = 1024;
  grp_table t;
  if (!grp_init(&t, cap, n_aggs)) return (uint64_t)-1;
  if (!keys) grp_slot(&t, 0, funcs, types, compensated);
  for (uint64_t i = 0; i < n; i++) {
    if (2 * (t.n_groups + 1) > cap) { /* grow: rehash */
      grp_table g;
      if (!grp_init(&g, cap * 4, n_aggs)) { grp_free(&t); return (uint64_t)-1; }
      for (uint64_t s = 0; s < cap; s++) if (t.used[s]) {
        uint64_t d = mix_u64((uint64_t)t.hdr[s].key) & g.cap_mask;
        while (g.used[d]) d = (d + 1) & g.cap_mask;
        g.used[d] = 1; g.hdr[d] = t.hdr[s]; g.n_groups++;
        memcpy(&g.acc[d * n_aggs], &t.acc[s * n_aggs], n_aggs * sizeof(grp_acc));
      }
      grp_free(&t); t = g; cap *= 4;
    }
    uint64_t s = grp_slot(&t, keys ? keys[i] : 0, funcs, types, compensated);
    t.hdr[s].count++;
    for (uint64_t a = 0; a < n_aggs; a++) acc_update(&t.acc[s * n_aggs + a], funcs[a], types[a], compensated, value_cols[a], i);
  }
  uint64_t G = t.n_groups;
  if (G > out_cap) { grp_free(&t); return (uint64_t)-1; }
  int64_t* order = (int64_t*)malloc((G ? G : 1) * 2 * sizeof(int64_t));
  uint64_t k = 0;
  for (uint64_t s = 0; s < cap; s++) if (t.used[s]) { order[2 * k] = t.hdr[s].key; order[2 * k + 1] = (int64_t)s; k++; }
  qsort(order, G, 2 * sizeof(int64_t), cmp_keyslot);
  for (uint64_t g = 0; g < G; g++) {
    uint64_t s = (uint64_t)order[2 * g + 1];
    out_keys[g] = t.hdr[s].key;
    for (uint64_t a = 0; a < n_aggs; a++)
      out_aggs[a * out_cap + g] = acc_final(&t.acc[s * n_aggs + a], funcs[a], types[a], compensated, t.hdr[s].count);
  }
  free(order);
  grp_free(&t);
  return G;
}

/* ------------------------------------------------------------------------------------------------------- */
/* Arrow layout helpers: record_batch_convert.rs:102-130.                                                    */
/* ------------------------------------------------------------------------------------------------------- */
ORC_API uint64_t orc_fnv1a64(const uint8_t* s, uint64_t len) {
  uint64_t h = 14695981039346656037ULL;
  for (uint64_t i = 0; i < len; i++) { h ^= s[i]; h *= 1099511628211ULL; }
  return h;
}
/* src_type: 0 i8, 1 i16, 2 i32, 3 i64, 4 u8, 5 u16, 6 u32, 7 u64, 8 f32, 9 f64; validity: 1 byte per row or NULL;
 * nulls become 0 / 0.0 (record_batch_convert.rs:103, :108) */
ORC_API int orc_widen_column(const void* src, int src_type, const uint8_t* validity, uint64_t n, uint64_t* out) {
  for (uint64_t i = 0; i < n; i++) {
    union { uint64_t u; int64_t s; double d; } v; v.u = 0;
    if (!validity || validity[i]) switch (src_type) {
      case 0: v.s = ((const int8_t*)src)[i]; break;
      case 1: v.s = ((const int16_t*)src)[i]; break;
      case 2: v.s = ((const int32_t*)src)[i]; break;
      case 3: v.s = ((const int64_t*)src)[i]; break;
      case 4: v.s = ((const uint8_t*)src)[i]; break;
      case 5: v.s = ((const uint16_t*)src)[i]; break;
      case 6: v.s = ((const uint32_t*)src)[i]; break;
      case 7: v.u = ((const uint64_t*)src)[i]; break;
      case 8: v.d = (double)((const float*)src)[i]; break;
      case 9: v.d = ((const double*)src)[i]; break;
      default: return -1;
    }
    out[i] = v.u;
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------------------- */
/* refN: the same results on all host threads, for the CPU baseline in bench.py.     */
/* ------------------------------------------------------------------------------------------------------- */
#define MT_BITS 10
#define MT_PARTS (1u << MT_BITS)

typedef struct {
  int tid, nthreads;
  const int64_t *build, *probe;
  uint64_t n_build, n_probe;
  uint64_t* hist;            /* [2][nthreads][MT_PARTS] */
  uint64_t* begin;           /* [2][MT_PARTS+1] */
  int64_t* pk[2]; uint32_t* pid[2];
  uint64_t* part_count;      /* matches per partition */
  uint64_t* part_out;        /* exclusive scan of part_count */
  uint32_t *out_left, *out_right; uint64_t max_matches;
  pthread_barrier_t* bar;
} mt_join_arg;

static inline uint32_t mt_part(int64_t k) { return (uint32_t)(mix_u64((uint64_t)k) >> (64 - MT_BITS)); }

static void mt_join_partition(mt_join_arg* a, uint32_t p, int write) {
  uint64_t b0 = a->begin[p], b1 = a->begin[p + 1];
  uint64_t p0 = a->begin[MT_PARTS + 1 + p], p1 = a->begin[MT_PARTS + 1 + p + 1];
  uint64_t nb = b1 - b0, cap = 16;
  while (cap < 2 * nb) cap <<= 1;
  uint32_t* slot = (uint32_t*)calloc(cap, sizeof(uint32_t)); /* build position + 1 */
  const int64_t* bk = a->pk[0] + b0; const uint32_t* bid = a->pid[0] + b0;
  for (uint64_t i = 0; i < nb; i++) {
    uint64_t s = (mix_u64((uint64_t)bk[i]) >> 7) & (cap - 1);
    while (slot[s]) s = (s + 1) & (cap - 1);
    slot[s] = (uint32_t)i + 1;
  }
  uint64_t cnt = 0, o = write ? a->part_out[p] : 0;
  for (uint64_t j = p0; j < p1; j++) {
    int64_t k = a->pk[1][j];
    uint64_t s = (mix_u64((uint64_t)k) >> 7) & (cap - 1);
    for (; slot[s]; s = (s + 1) & (cap - 1)) {
      if (bk[slot[s] - 1] == k) {
        if (write && o + cnt < a->max_matches) { a->out_left[o + cnt] = bid[slot[s] - 1]; a->out_right[o + cnt] = a->pid[1][j]; }
        cnt++;
      }
    }
  }
  if (!write) a->part_count[p] = cnt;
  free(slot);
}

static void* mt_join_worker(void* vp) {
  mt_join_arg* a = (mt_join_arg*)vp;
  const int T = a->nthreads, t = a->tid;
  for (int side = 0; side < 2; side++) {
    const int64_t* src = side ? a->probe : a->build;
    uint64_t n = side ? a->n_probe : a->n_build;
    uint64_t lo = n * (uint64_t)t / T, hi = n * (uint64_t)(t + 1) / T;
    uint64_t* h = a->hist + ((uint64_t)side * T + t) * MT_PARTS;
    memset(h, 0, MT_PARTS * sizeof(uint64_t));
    for (uint64_t i = lo; i < hi; i++) h[mt_part(src[i])]++;
  }
  pthread_barrier_wait(a->bar);
  if (t == 0) {
    for (int side = 0; side < 2; side++) {
      uint64_t run = 0;
      uint64_t* beg = a->begin + (uint64_t)side * (MT_PARTS + 1);
      for (uint32_t p = 0; p < MT_PARTS; p++) {
        beg[p] = run;
        for (int u = 0; u < T; u++) { uint64_t* h = a->hist + ((uint64_t)side * T + u) * MT_PARTS; uint64_t c = h[p]; h[p] = run; run += c; }
      }
      beg[MT_PARTS] = run;
    }
  }
  pthread_barrier_wait(a->bar);
  for (int side = 0; side < 2; side++) {
    const int64_t* src = side ? a->probe : a->build;
    uint64_t n = side ? a->n_probe : a->n_build;
    uint64_t lo = n * (uint64_t)t / T, hi = n * (uint64_t)(t + 1) / T;
    uint64_t* h = a->hist + ((uint64_t)side * T + t) * MT_PARTS;
    for (uint64_t i = lo; i < hi; i++) { uint64_t pos = h[mt_part(src[i])]++; a->pk[side][pos] = src[i]; a->pid[side][pos] = (uint32_t)i; }
  }
  pthread_barrier_wait(a->bar);
  for (uint32_t p = (uint32_t)t; p < MT_PARTS; p += (uint32_t)T) mt_join_partition(a, p, 0);
  pthread_barrier_wait(a->bar);
  if (t == 0) { uint64_t run = 0; for (uint32_t p = 0; p < MT_PARTS; p++) { a->part_out[p] = run; run += a->part_count[p]; } a->part_out[MT_PARTS] = run; }
  pthread_barrier_wait(a->bar);
  if (a->out_left) for (uint32_t p = (uint32_t)t; p < MT_PARTS; p += (uint32_t)T) mt_join_partition(a, p, 1);
  return NULL;
}

/* Radix-partitioned multi-threaded join; same match multiset as orc_hash_join_i64 (order differs). */
ORC_API int orc_hash_join_i64_mt(const int64_t* build, uint64_t n_build, const int64_t* probe, uint64_t n_probe,
                                 uint32_t* out_left, uint32_t* out_right, uint64_t max_matches, uint64_t* out_count, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  const int T = nthreads;
  mt_join_arg* args = (mt_join_arg*)calloc((size_t)T, sizeof(mt_join_arg));
  pthread_t* th = (pthread_t*)malloc((size_t)T * sizeof(pthread_t));
  uint64_t* hist = (uint64_t*)malloc((size_t)2 * T * MT_PARTS * sizeof(uint64_t));
  uint64_t* begin = (uint64_t*)malloc(2 * (MT_PARTS + 1) * sizeof(uint64_t));
  uint64_t* pc = (uint64_t*)calloc(MT_PARTS + 1, sizeof(uint64_t));
  uint64_t* po = (uint64_t*)calloc(MT_PARTS + 1, sizeof(uint64_t));
  int64_t* pk0 = (int64_t*)malloc((n_build ? n_build : 1) * 8); uint32_t* pi0 = (uint32_t*)malloc((n_build ? n_build : 1) * 4);
  int64_t* pk1 = (int64_t*)malloc((n_probe ? n_probe : 1) * 8); uint32_t* pi1 = (uint32_t*)malloc((n_probe ? n_probe : 1) * 4);
  pthread_barrier_t bar;
  pthread_barrier_init(&bar, NULL, (unsigned)T);
  for (int t = 0; t < T; t++) {
    mt_join_arg a = {t, T, build, probe, n_build, n_probe, hist, begin, {pk0, pk1}, {pi0, pi1}, pc, po, out_left, out_right, max_matches, &bar};
    args[t] = a;
    pthread_create(&th[t], NULL, mt_join_worker, &args[t]);
  }
  for (int t = 0; t < T; t++) pthread_join(th[t], NULL);
  *out_count = po[MT_PARTS];
  pthread_barrier_destroy(&bar);
  free(args); free(th); free(hist); free(begin); free(pc); free(po); free(pk0); free(pi0); free(pk1); free(pi1);
  return 0;
}

/* Multi-threaded grouped aggregation for i64 values: thread-local tables over row ranges, merged by thread 0.
 * Only order-independent aggregates (i64 SUM/COUNT/MIN/MAX, AVG of i64) so results equal orc_group_by_i64.   */
typedef struct {
  int tid, nthreads; const int64_t* keys; uint64_t n, n_aggs; const void* const* cols; const int* types; const int* funcs;
  grp_table table; uint64_t cap; int ok;
} mt_grp_arg;

static void* mt_grp_worker(void* vp) {
  mt_grp_arg* a = (mt_grp_arg*)vp;
  uint64_t lo = a->n * (uint64_t)a->tid / a->nthreads, hi = a->n * (uint64_t)(a->tid + 1) / a->nthreads;
  a->cap = 1024;
  a->ok = grp_init(&a->table, a->cap, a->n_aggs);
  if (!a->ok) return NULL;
  for (uint64_t i = lo; i < hi; i++) {
    if (2 * (a->table.n_groups + 1) > a->cap) {
      grp_table g;
      if (!grp_init(&g, a->cap * 4, a->n_aggs)) { a->ok = 0; return NULL; }
      for (uint64_t s = 0; s < a->cap; s++) if (a->table.used[s]) {
        uint64_t d = mix_u64((uint64_t)a->table.hdr[s].key) & g.cap_mask;
        while (g.used[d]) d = (d + 1) & g.cap_mask;
        g.used[d] = 1; g.hdr[d] = a->table.hdr[s]; g.n_groups++;
        memcpy(&g.acc[d * a->n_aggs], &a->table.acc[s * a->n_aggs], a->n_aggs * sizeof(grp_acc));
      }
      grp_free(&a->table); a->table = g; a->cap *= 4;
    }
    uint64_t s = grp_slot(&a->table, a->keys[i], a->funcs, a->types, 0);
    a->table.hdr[s].count++;
    for (uint64_t q = 0; q < a->n_aggs; q++) acc_update(&a->table.acc[s * a->n_aggs + q], a->funcs[q], a->types[q], 0, a->cols[q], i);
  }
  return NULL;
}

ORC_API uint64_t orc_group_by_i64_mt(const int64_t* keys, uint64_t n, uint64_t n_aggs, const void* const* value_cols,
                                     const int* types, const int* funcs, uint64_t out_cap, int64_t* out_keys,
                                     uint64_t* out_aggs, int nthreads) {
  for (uint64_t a = 0; a < n_aggs; a++) if (types[a] != ORC_I64) return (uint64_t)-1;
  if (nthreads < 1) nthreads = 1;
  const int T = nthreads;
  mt_grp_arg* args = (mt_grp_arg*)calloc((size_t)T, sizeof(mt_grp_arg));
  pthread_t* th = (pthread_t*)malloc((size_t)T * sizeof(pthread_t));
  for (int t = 0; t < T; t++) {
    args[t].tid = t; args[t].nthreads = T; args[t].keys = keys; args[t].n = n; args[t].n_aggs = n_aggs;
    args[t].cols = value_cols; args[t].types = types; args[t].funcs = funcs;
    pthread_create(&th[t], NULL, mt_grp_worker, &args[t]);
  }
  for (int t = 0; t < T; t++) pthread_join(th[t], NULL);
  /* merge into thread 0's table */
  uint64_t result = (uint64_t)-1;
  grp_table* m = &args[0].table; uint64_t mcap = args[0].cap;
  int ok = 1;
  for (int t = 0; t < T; t++) ok &= args[t].ok;
  if (ok) {
    for (int t = 1; t < T; t++) {
      grp_table* s = &args[t].table;
      for (uint64_t i = 0; i < args[t].cap; i++) if (s->used[i]) {
        if (2 * (m->n_groups + 1) > mcap) {
          grp_table g; grp_init(&g, mcap * 4, n_aggs);
          for (uint64_t q = 0; q < mcap; q++) if (m->used[q]) {
            uint64_t d = mix_u64((uint64_t)m->hdr[q].key) & g.cap_mask;
            while (g.used[d]) d = (d + 1) & g.cap_mask;
            g.used[d] = 1; g.hdr[d] = m->hdr[q]; g.n_groups++;
            memcpy(&g.acc[d * n_aggs], &m->acc[q * n_aggs], n_aggs * sizeof(grp_acc));
          }
          grp_free(m); *m = g; mcap *= 4;
        }
        uint64_t d = grp_slot(m, s->hdr[i].key, funcs, types, 0);
        m->hdr[d].count += s->hdr[i].count;
        for (uint64_t q = 0; q < n_aggs; q++) {
          grp_acc* x = &m->acc[d * n_aggs + q]; const grp_acc* y = &s->acc[i * n_aggs + q];
          if (funcs[q] == ORC_MIN) { if (y->i < x->i) x->i = y->i; }
          else if (funcs[q] == ORC_MAX) { if (y->i > x->i) x->i = y->i; }
          else if (funcs[q] != ORC_COUNT) x->i = (int64_t)((uint64_t)x->i + (uint64_t)y->i);
        }
      }
    }
    uint64_t G = m->n_groups;
    if (G <= out_cap) {
      int64_t* order = (int64_t*)malloc((G ? G : 1) * 2 * sizeof(int64_t));
      uint64_t k = 0;
      for (uint64_t s = 0; s < mcap; s++) if (m->used[s]) { order[2 * k] = m->hdr[s].key; order[2 * k + 1] = (int64_t)s; k++; }
      qsort(order, G, 2 * sizeof(int64_t), cmp_keyslot);
      for (uint64_t g = 0; g < G; g++) {
        uint64_t s = (uint64_t)order[2 * g + 1];
        out_keys[g] = m->hdr[s].key;
        for (uint64_t q = 0; q < n_aggs; q++) out_aggs[q * out_cap + g] = acc_final(&m->acc[s * n_aggs + q], funcs[q], types[q], 0, m->hdr[s].count);
      }
      free(order);
      result = G;
    }
  }
  for (int t = 0; t < T; t++) if (args[t].ok) grp_free(&args[t].table);
  free(args); free(th);
  return result;
}
