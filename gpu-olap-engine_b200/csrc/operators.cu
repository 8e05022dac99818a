// Operator layer: PhysicalPlan interpreter over device-resident column batches, exchanged with the host through the
// Arrow C Data Interface.  Host-side C++ mirror of gpu-olap-core/src/executor.rs (GpuExecutor::execute :37-104 and the
// per-operator methods :110-355) with the stubs completed per SURVEY.md 8(c); every data-touching step is one of the
// sm_100a kernels behind the kernel-level C ABI (filter.cu / groupby.cu / join.cu).  No CPU compute path.
#include "common.cuh"

#include <cerrno>
#include <cfloat>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <cmath>
#include <map>

namespace {

using namespace olap;

// ------------------------------------------------------------------------------------------- device batches
struct DevBuf {
  olap_ctx* ctx; void* p = nullptr;
  DevBuf(olap_ctx* c) : ctx(c) {}
  ~DevBuf() { if (p) olap_pool_free(ctx, p, nullptr); }
};
using BufPtr = std::shared_ptr<DevBuf>;

static olap_status new_buf(olap_ctx* ctx, uint64_t bytes, BufPtr* out) {
  BufPtr b = std::make_shared<DevBuf>(ctx);
  OLAP_TRY(olap_pool_alloc(ctx, bytes, nullptr, &b->p));
  *out = b;
  return OLAP_OK;
}

struct Col {
  std::string name;
  int type = OLAP_I64;  // OLAP_I64 / OLAP_F64 (record_batch_convert.rs: everything is one of the two 8-byte layouts)
  BufPtr buf;
  uint64_t off = 0;     // first row inside buf: the batches of a consolidated table are views into one allocation per column
  uint64_t* data() const { return (uint64_t*)buf->p + off; }
  // optional validity, one byte per row, 1 = valid (column_buffer.rs:37-41).  Set by olap_register_batch for columns that arrive
  // with nulls and by outer joins for the NULL-extended side; carried by scans, exported as an Arrow bitmap.  Every other operator
  // reads a null as the 0 / 0.0 stored in its place (record_batch_convert.rs:103, :108) and emits non-null columns.
  BufPtr valid;
  uint8_t* validity() const { return valid ? (uint8_t*)valid->p : nullptr; }
};
struct Batch { std::vector<Col> cols; uint64_t n = 0; };
using Batches = std::vector<Batch>;

struct Catalog { std::mutex mu; std::map<std::string, Batches> tables; };
static void catalog_free(void* p) { delete (Catalog*)p; }
static Catalog* catalog_of(olap_ctx* ctx) {
  std::lock_guard<std::mutex> g(ctx->mu);
  if (!ctx->op_state) { ctx->op_state = new Catalog(); ctx->op_state_free = catalog_free; }
  return (Catalog*)ctx->op_state;
}

// ------------------------------------------------------------------------------------------- small kernels
__global__ void __launch_bounds__(256) fill_kernel(uint64_t* out, uint64_t n, uint64_t v) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x, st = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) out[i] = v;
}
// out[i] = col[pairs[2 * i + side]]: materialises one join output column from the row-id pairs
// Four pairs per thread and iteration: the four row ids are loaded first, then the four gathers are in flight together (one
// dependent load chain per value left the DRAM latency of every gather exposed).
__global__ void __launch_bounds__(256) gather_pairs_kernel(const uint64_t* __restrict__ col, const uint32_t* __restrict__ pairs, int side,
                                                           uint64_t n, uint64_t* __restrict__ out) {
  const uint64_t st = (uint64_t)gridDim.x * blockDim.x;
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  for (; i + 3 * st < n; i += 4 * st) {
    uint32_t id[4];
#pragma unroll
    for (int k = 0; k < 4; k++) id[k] = ld_stream_u32(pairs + 2 * (i + k * st) + side);
    uint64_t v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = __ldg(col + id[k]);
#pragma unroll
    for (int k = 0; k < 4; k++) out[i + k * st] = v[k];
  }
  for (; i < n; i += st) out[i] = __ldg(col + ld_stream_u32(pairs + 2 * i + side));
}
// flags[pairs[i].side] = 1: the rows of one join side that found a partner (8-byte flags: the ordered selection pass reads them
// as an Int64 column)
__global__ void __launch_bounds__(256) mark_pairs_kernel(const uint32_t* __restrict__ pairs, int side, uint64_t n, uint64_t* __restrict__ flags) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x, st = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) flags[ld_stream_u32(pairs + 2 * i + side)] = 1;
}
// out[i] = col[i * stride]: an evenly spaced sample of a column
__global__ void __launch_bounds__(256) sample_kernel(const uint64_t* __restrict__ col, uint64_t stride, uint64_t m, uint64_t* __restrict__ out) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x, st = (uint64_t)gridDim.x * blockDim.x;
  for (; i < m; i += st) out[i] = col[i * stride];
}
static unsigned small_grid(olap_ctx* ctx, uint64_t n) {
  uint64_t g = (n + 255) / 256, cap = (uint64_t)ctx->sm_count * 8;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

// ------------------------------------------------------------------------------------------- IR
enum { EXPR_COLUMN = 0, EXPR_LITERAL = 1, EXPR_BINARY = 2 };
struct olap_expr {
  int kind; uint64_t index = 0; std::string text;  // column name / literal text / operator
  olap_expr *l = nullptr, *r = nullptr;
  ~olap_expr() { delete l; delete r; }
};
enum { PLAN_SCAN = 0, PLAN_FILTER, PLAN_PROJECTION, PLAN_HASH_JOIN, PLAN_AGGREGATE };
struct Field { std::string name; bool is_float; };
struct olap_plan {
  int kind;
  std::string table; std::vector<Field> schema; bool has_schema = false, has_proj = false; std::vector<int64_t> proj;
  olap_plan *in = nullptr, *right = nullptr;
  olap_expr* pred = nullptr;
  std::vector<olap_expr*> exprs, lkeys, rkeys, group_by, agg_in;
  std::vector<int> agg_funcs;
  int join_type = OLAP_JOIN_INNER;
  ~olap_plan() {
    delete in; delete right; delete pred;
    for (auto v : {&exprs, &lkeys, &rkeys, &group_by, &agg_in}) for (auto e : *v) delete e;
  }
};
struct olap_result { Batches batches; };

namespace {

// expr_name, executor.rs:529-537
static std::string expr_name(const olap_expr* e) {
  switch (e->kind) {
    case EXPR_COLUMN: return e->text;
    case EXPR_LITERAL: return "lit(" + e->text + ")";
    default: return "(" + expr_name(e->l) + " " + e->text + " " + expr_name(e->r) + ")";
  }
}

// value of an expression on one batch: a device column or a broadcast literal
struct Value { bool is_lit = false; int type = OLAP_I64; uint64_t bits = 0; BufPtr buf; uint64_t off = 0; uint64_t* data() const { return (uint64_t*)buf->p + off; } };

static bool parse_literal(const std::string& s, Value* v) {  // executor.rs:390-401: i64 first, then f64
  if (s.empty()) return false;
  char* end = nullptr;
  errno = 0;
  long long i = strtoll(s.c_str(), &end, 10);
  if (errno == 0 && end && *end == 0) { v->is_lit = true; v->type = OLAP_I64; v->bits = (uint64_t)i; return true; }
  errno = 0;
  double d = strtod(s.c_str(), &end);
  if (end && *end == 0 && end != s.c_str()) { v->is_lit = true; v->type = OLAP_F64; memcpy(&v->bits, &d, 8); return true; }
  return false;
}

static olap_status materialise(olap_ctx* ctx, const Value& v, uint64_t n, Value* out) {
  if (!v.is_lit) { *out = v; return OLAP_OK; }
  Value r; r.type = v.type;
  OLAP_TRY(new_buf(ctx, n * 8, &r.buf));
  if (n) { fill_kernel<<<small_grid(ctx, n), 256, 0, ctx->streams[0]>>>(r.data(), n, v.bits); count_launch(ctx); OLAP_CUDA(cudaGetLastError()); }
  *out = r;
  return OLAP_OK;
}

static olap_status eval_expr(olap_ctx* ctx, const olap_expr* e, const Batch& b, Value* out) {
  if (e->kind == EXPR_COLUMN) {  // by name first, then by index (executor.rs:381-387)
    for (const Col& c : b.cols)
      if (c.name == e->text) { out->is_lit = false; out->type = c.type; out->buf = c.buf; out->off = c.off; return OLAP_OK; }
    if (e->index < b.cols.size()) { const Col& c = b.cols[e->index]; out->is_lit = false; out->type = c.type; out->buf = c.buf; out->off = c.off; return OLAP_OK; }
    set_error("Column '%s' not found in batch", e->text.c_str());
    return OLAP_ERR_BAD_ARG;
  }
  if (e->kind == EXPR_LITERAL) {
    if (!parse_literal(e->text, out)) { set_error("Cannot parse literal '%s'", e->text.c_str()); return OLAP_ERR_BAD_ARG; }
    return OLAP_OK;
  }
  Value l, r;
  OLAP_TRY(eval_expr(ctx, e->l, b, &l));
  OLAP_TRY(eval_expr(ctx, e->r, b, &r));
  // eval_binary_expr (executor.rs:411-441) is Int64-only and errors otherwise; Float64 x Float64 is this core's extension
  if (l.type != r.type) { set_error("Binary expr: operands of '%s' have different types (Int64 vs Float64)", e->text.c_str()); return OLAP_ERR_UNSUPPORTED; }
  const char op = e->text.size() == 1 ? e->text[0] : '?';  // any other operator yields 0 (executor.rs:435)
  Value lm;
  OLAP_TRY(materialise(ctx, l, b.n, &lm));
  Value res; res.type = l.type;
  OLAP_TRY(new_buf(ctx, b.n * 8, &res.buf));
  if (l.type == OLAP_I64)
    OLAP_TRY(olap_project_i64(ctx, (const int64_t*)lm.data(), r.is_lit ? nullptr : (const int64_t*)r.data(), (int64_t)r.bits, op, b.n, (int64_t*)res.data(), nullptr));
  else {
    double lit; memcpy(&lit, &r.bits, 8);
    OLAP_TRY(olap_project_f64(ctx, (const double*)lm.data(), r.is_lit ? nullptr : (const double*)r.data(), lit, op, b.n, (double*)res.data(), nullptr));
  }
  *out = res;
  return OLAP_OK;
}

static int cmp_code(const std::string& op) {  // executor.rs:457-465
  if (op == "=" || op == "==") return OLAP_EQ;
  if (op == "!=" || op == "<>") return OLAP_NE;
  if (op == ">") return OLAP_GT;
  if (op == ">=") return OLAP_GE;
  if (op == "<") return OLAP_LT;
  if (op == "<=") return OLAP_LE;
  return -1;  // unknown operator: no row passes
}
static int flip(int op) { switch (op) { case OLAP_LT: return OLAP_GT; case OLAP_LE: return OLAP_GE; case OLAP_GT: return OLAP_LT; case OLAP_GE: return OLAP_LE; default: return op; } }

// predicate -> conjunction terms.  keep: temporaries that must outlive the kernel call.
static olap_status collect_terms(olap_ctx* ctx, const olap_expr* e, const Batch& b, std::vector<olap_filter_term>* terms, std::vector<Value>* keep) {
  if (e->kind != EXPR_BINARY) return OLAP_OK;  // executor.rs:474: not a binary expression -> every row passes
  if (e->text == "AND" || e->text == "and" || e->text == "&&") {  // a real conjunction, not the all-false bug
    OLAP_TRY(collect_terms(ctx, e->l, b, terms, keep));
    return collect_terms(ctx, e->r, b, terms, keep);
  }
  Value l, r;
  OLAP_TRY(eval_expr(ctx, e->l, b, &l));
  OLAP_TRY(eval_expr(ctx, e->r, b, &r));
  if (l.type != r.type) return OLAP_OK;  // executor.rs:471-472: operands are not both Int64 -> pass all
  int op = cmp_code(e->text);
  if (l.is_lit && !r.is_lit) { std::swap(l, r); op = flip(op); }
  if (l.is_lit) { Value m; OLAP_TRY(materialise(ctx, l, b.n, &m)); l = m; }
  olap_filter_term t;
  t.lhs = l.data(); t.rhs = r.is_lit ? nullptr : r.data(); t.literal_bits = (int64_t)r.bits; t.op = op; t.type = l.type;
  terms->push_back(t);
  keep->push_back(l); keep->push_back(r);
  return OLAP_OK;
}

static olap_status exec(olap_ctx* ctx, const olap_plan* p, Batches* out);

// execute_table_scan, executor.rs:110-155 -- wired to the in-memory catalog instead of returning nothing
// A table registered batch by batch is moved, the first time it is scanned, into ONE allocation per column; its batches
// become views (same boundaries, same order).  Operators that need all rows of a column at once (aggregate, join keys) then
// take a zero-copy view instead of concatenating on every query.
static olap_status consolidate(olap_ctx* ctx, Batches* t) {
  if (t->size() < 2) return OLAP_OK;
  const size_t nc = (*t)[0].cols.size();
  uint64_t total = 0;
  bool done = true;
  for (const Batch& b : *t) {
    if (b.cols.size() != nc) return OLAP_OK;  // ragged schemas stay as they are
    for (size_t c = 0; c < nc; c++) done = done && b.cols[c].buf == (*t)[0].cols[c].buf && b.cols[c].off == total;
    total += b.n;
  }
  if (done || total == 0) return OLAP_OK;
  for (size_t c = 0; c < nc; c++) {
    BufPtr arena;
    OLAP_TRY(new_buf(ctx, total * 8, &arena));
    uint64_t off = 0;
    for (Batch& b : *t) {
      if (b.n) OLAP_CUDA(cudaMemcpyAsync((uint64_t*)arena->p + off, b.cols[c].data(), b.n * 8, cudaMemcpyDeviceToDevice, ctx->streams[0]));
      off += b.n;
    }
    OLAP_CUDA(cudaStreamSynchronize(ctx->streams[0]));  // the sources are released below
    off = 0;
    for (Batch& b : *t) { b.cols[c].buf = arena; b.cols[c].off = off; off += b.n; }
  }
  return OLAP_OK;
}

static olap_status exec_scan(olap_ctx* ctx, const olap_plan* p, Batches* out) {
  Catalog* cat = catalog_of(ctx);
  Batches src;
  {
    std::lock_guard<std::mutex> g(cat->mu);
    auto it = cat->tables.find(p->table);
    if (it != cat->tables.end()) {
      OLAP_TRY(consolidate(ctx, &it->second));
      src = it->second;
    }
  }
  auto project = [&](const Batch& b, Batch* o) -> olap_status {
    o->n = b.n;
    if (!p->has_proj) { o->cols = b.cols; return OLAP_OK; }
    for (int64_t i : p->proj) {
      if (i < 0 || (size_t)i >= b.cols.size()) { set_error("projection index %lld out of range", (long long)i); return OLAP_ERR_BAD_ARG; }
      o->cols.push_back(b.cols[i]);
    }
    return OLAP_OK;
  };
  if (!src.empty()) { for (const Batch& b : src) { Batch o; OLAP_TRY(project(b, &o)); out->push_back(o); } return OLAP_OK; }
  if (!p->has_schema) { set_error("table '%s' is not registered and the scan carries no schema", p->table.c_str()); return OLAP_ERR_BAD_ARG; }
  Batch empty;  // float fields -> Float64, everything else -> Int64 (executor.rs:138-149)
  for (const Field& f : p->schema) { Col c; c.name = f.name; c.type = f.is_float ? OLAP_F64 : OLAP_I64; OLAP_TRY(new_buf(ctx, 16, &c.buf)); empty.cols.push_back(c); }
  Batch o; OLAP_TRY(project(empty, &o));
  out->push_back(o);
  return OLAP_OK;
}

// execute_filter, executor.rs:161-178 + filter_record_batch :478-494, fused into one pass per batch
static olap_status exec_filter(olap_ctx* ctx, const olap_plan* p, Batches* out) {
  Batches in;
  OLAP_TRY(exec(ctx, p->in, &in));
  for (const Batch& b : in) {
    std::vector<olap_filter_term> terms; std::vector<Value> keep;
    OLAP_TRY(collect_terms(ctx, p->pred, b, &terms, &keep));
    if (terms.size() > 8) { set_error("more than 8 conjuncts in one filter"); return OLAP_ERR_UNSUPPORTED; }
    if (terms.empty()) { if (b.n) out->push_back(b); continue; }
    Batch o; o.cols.resize(b.cols.size());
    uint64_t kept = 0;
    for (size_t i = 0; i < b.cols.size(); i++) { o.cols[i].name = b.cols[i].name; o.cols[i].type = b.cols[i].type; OLAP_TRY(new_buf(ctx, b.n * 8, &o.cols[i].buf)); }
    for (size_t c0 = 0; c0 < b.cols.size() || c0 == 0; c0 += 8) {
      const int nc = (int)std::min<size_t>(8, b.cols.size() - c0);
      const void* ins[8]; void* outs[8];
      for (int i = 0; i < nc; i++) { ins[i] = b.cols[c0 + i].data(); outs[i] = o.cols[c0 + i].data(); }
      OLAP_TRY(olap_filter_compact(ctx, terms.data(), (int32_t)terms.size(), b.n, ins, nc, outs, nullptr, &kept, nullptr));
      if (b.cols.size() <= 8) break;
    }
    o.n = kept;
    if (kept > 0) out->push_back(o);  // executor.rs:173: batches that become empty are dropped
  }
  return OLAP_OK;
}

// execute_projection, executor.rs:184-214
static olap_status exec_projection(olap_ctx* ctx, const olap_plan* p, Batches* out) {
  Batches in;
  OLAP_TRY(exec(ctx, p->in, &in));
  for (const Batch& b : in) {
    Batch o; o.n = b.n;
    for (const olap_expr* e : p->exprs) {
      Value v, m;
      OLAP_TRY(eval_expr(ctx, e, b, &v));
      OLAP_TRY(materialise(ctx, v, b.n, &m));
      Col c; c.name = expr_name(e); c.type = m.type; c.buf = m.buf; c.off = m.off;
      if (e->kind == EXPR_COLUMN) {  // a column passed through keeps its validity (resolved like eval_expr: name first, then index)
        const Col* src = nullptr;
        for (const Col& x : b.cols) if (x.name == e->text) { src = &x; break; }
        if (!src && e->index < b.cols.size()) src = &b.cols[e->index];
        if (src) c.valid = src->valid;
      }
      o.cols.push_back(c);
    }
    out->push_back(o);
  }
  return OLAP_OK;
}

// all batches of one side as single columns (global row offset across batches, executor.rs:239-253)

static olap_status eval_over_batches(olap_ctx* ctx, const olap_expr* e, const Batches& in, uint64_t total, Value* out) {
  if (in.size() == 1) { Value v; OLAP_TRY(eval_expr(ctx, e, in[0], &v)); return materialise(ctx, v, in[0].n, out); }
  Value r; bool first = true;
  uint64_t off = 0;
  for (const Batch& b : in) {
    Value v, m;
    OLAP_TRY(eval_expr(ctx, e, b, &v));
    OLAP_TRY(materialise(ctx, v, b.n, &m));
    if (first) { r.type = m.type; OLAP_TRY(new_buf(ctx, total * 8, &r.buf)); first = false; }
    if (m.type != r.type) { set_error("expression changes type between batches"); return OLAP_ERR_BAD_ARG; }
    if (b.n) OLAP_CUDA(cudaMemcpyAsync(r.data() + off, m.data(), b.n * 8, cudaMemcpyDeviceToDevice, ctx->streams[0]));
    off += b.n;
  }
  if (first) OLAP_TRY(new_buf(ctx, 16, &r.buf));
  *out = r;
  return OLAP_OK;
}

// execute_hash_join, executor.rs:220-265, with the probe and the output the reference leaves out
// positions (in the input's column order) of the columns an expression reads, resolved like eval_expr does
static void referenced_columns(const olap_expr* e, const Batch& b, std::vector<bool>* used) {
  if (!e) return;
  if (e->kind == EXPR_BINARY) { referenced_columns(e->l, b, used); referenced_columns(e->r, b, used); return; }
  if (e->kind != EXPR_COLUMN) return;
  for (size_t i = 0; i < b.cols.size(); i++)
    if (b.cols[i].name == e->text) { (*used)[i] = true; return; }
  if (e->index < b.cols.size()) (*used)[e->index] = true;
}

// One batch over all input rows holding (only) the columns in `used`; a single input batch is taken as is.
static olap_status concat_columns(olap_ctx* ctx, const Batches& in, const std::vector<bool>& used, uint64_t total, Batch* all) {
  if (in.size() == 1) { *all = in[0]; return OLAP_OK; }
  all->n = total;
  all->cols.resize(in[0].cols.size());
  for (size_t c = 0; c < in[0].cols.size(); c++) {
    all->cols[c].name = in[0].cols[c].name; all->cols[c].type = in[0].cols[c].type;
    if (!used[c]) continue;
    bool contiguous = true;  // consolidated catalog tables: consecutive views into one allocation
    uint64_t next = in[0].cols[c].off;
    for (const Batch& b : in) {
      contiguous = contiguous && b.cols.size() == in[0].cols.size() && b.cols[c].buf == in[0].cols[c].buf && b.cols[c].off == next;
      next += b.n;
    }
    if (contiguous) { all->cols[c].buf = in[0].cols[c].buf; all->cols[c].off = in[0].cols[c].off; continue; }
    OLAP_TRY(new_buf(ctx, std::max<uint64_t>(total, 2) * 8, &all->cols[c].buf));
    uint64_t off = 0;
    for (const Batch& b : in) {
      if (b.cols.size() != in[0].cols.size() || b.cols[c].type != in[0].cols[c].type) { set_error("batches of one input disagree on their schema"); return OLAP_ERR_BAD_ARG; }
      if (b.n) OLAP_CUDA(cudaMemcpyAsync(all->cols[c].data() + off, b.cols[c].data(), b.n * 8, cudaMemcpyDeviceToDevice, ctx->streams[0]));
      off += b.n;
    }
  }
  return OLAP_OK;
}

static olap_status exec_hash_join(olap_ctx* ctx, const olap_plan* p, Batches* out) {
  Batches left, right;
  OLAP_TRY(exec(ctx, p->in, &left));
  OLAP_TRY(exec(ctx, p->right, &right));
  if (left.empty() || right.empty()) return OLAP_OK;  // executor.rs:231-233
  const int jt = p->join_type;
  if (jt < OLAP_JOIN_INNER || jt > OLAP_JOIN_FULL) { set_error("unknown join type %d (logical_plan.rs:52-58)", jt); return OLAP_ERR_BAD_ARG; }
  if (p->lkeys.empty() || p->rkeys.empty()) { set_error("hash join needs one key expression per side"); return OLAP_ERR_BAD_ARG; }
  uint64_t nl = 0, nr = 0;
  for (const Batch& b : left) nl += b.n;
  for (const Batch& b : right) nr += b.n;
  Batch L, R;  // every column of both sides over all rows (views when the inputs are consolidated tables)
  OLAP_TRY(concat_columns(ctx, left, std::vector<bool>(left[0].cols.size(), true), nl, &L));
  OLAP_TRY(concat_columns(ctx, right, std::vector<bool>(right[0].cols.size(), true), nr, &R));
  L.n = nl; R.n = nr;
  Value lk, rk;  // first key expression of each side only (executor.rs:245)
  OLAP_TRY(eval_over_batches(ctx, p->lkeys[0], Batches{L}, nl, &lk));
  OLAP_TRY(eval_over_batches(ctx, p->rkeys[0], Batches{R}, nr, &rk));
  if (lk.type != OLAP_I64 || rk.type != OLAP_I64) { set_error("hash join keys must be Int64 (executor.rs:246)"); return OLAP_ERR_UNSUPPORTED; }
  uint64_t cap = std::max<uint64_t>(nr, 1), count = 0;
  BufPtr pairs;
  for (int attempt = 0; attempt < 2; attempt++) {
    OLAP_TRY(new_buf(ctx, cap * 8, &pairs));
    olap_status st = olap_hash_join_i64(ctx, (const int64_t*)lk.data(), nl, (const int64_t*)rk.data(), nr, (olap_pair*)pairs->p, cap, nullptr, &count,
                                        OLAP_JOIN_AUTO, nullptr);
    if (st == OLAP_ERR_OVERFLOW && attempt == 0) { cap = count; continue; }  // count is exact: size the output and redo
    OLAP_TRY(st);
    break;
  }
  // Left / Right / Full (logical_plan.rs:52-58): the preserved side's rows without a partner follow the inner pairs, in row
  // order -- unmatched left rows first, then unmatched right rows -- with the other side's columns NULL (value 0, validity 0)
  uint64_t ul = 0, ur = 0;
  BufPtr lsel, rsel;
  if (jt == OLAP_JOIN_LEFT || jt == OLAP_JOIN_FULL) {
    OLAP_TRY(new_buf(ctx, std::max<uint64_t>(nl, 1) * 4, &lsel));
    OLAP_TRY(olap_join_unmatched_rows(ctx, (const olap_pair*)pairs->p, count, 0, nl, (uint32_t*)lsel->p, nullptr, &ul, nullptr));
  }
  if (jt == OLAP_JOIN_RIGHT || jt == OLAP_JOIN_FULL) {
    OLAP_TRY(new_buf(ctx, std::max<uint64_t>(nr, 1) * 4, &rsel));
    OLAP_TRY(olap_join_unmatched_rows(ctx, (const olap_pair*)pairs->p, count, 1, nr, (uint32_t*)rsel->p, nullptr, &ur, nullptr));
  }
  const uint64_t total = count + ul + ur;
  cudaStream_t s0 = ctx->streams[0];
  Batch o; o.n = total;
  for (int side = 0; side < 2; side++) {
    const Batch& S = side ? R : L;
    const uint64_t own = side ? ur : ul, own_at = side ? count + ul : count;      // this side's unmatched rows and where they go
    const uint64_t null_n = side ? ul : ur, null_at = side ? count : count + ul;  // rows where this side is NULL
    for (const Col& c : S.cols) {  // left columns, then right columns, gathered by the row-id pairs
      Col oc; oc.name = c.name; oc.type = c.type;
      OLAP_TRY(new_buf(ctx, std::max<uint64_t>(total, 1) * 8, &oc.buf));
      if (count) { gather_pairs_kernel<<<small_grid(ctx, count), 256, 0, s0>>>(c.data(), (const uint32_t*)pairs->p, side, count, oc.data()); count_launch(ctx); }
      if (own) OLAP_TRY(olap_gather_8b(ctx, c.data(), (const uint32_t*)(side ? rsel : lsel)->p, own, oc.data() + own_at, nullptr));
      if (null_n) {
        OLAP_CUDA(cudaMemsetAsync(oc.data() + null_at, 0, null_n * 8, s0));
        OLAP_TRY(new_buf(ctx, total, &oc.valid));
        OLAP_CUDA(cudaMemsetAsync(oc.validity(), 1, total, s0));
        OLAP_CUDA(cudaMemsetAsync(oc.validity() + null_at, 0, null_n, s0));
      }
      o.cols.push_back(oc);
    }
  }
  OLAP_CUDA(cudaGetLastError());
  OLAP_CUDA(cudaStreamSynchronize(ctx->streams[0]));
  out->push_back(o);
  return OLAP_OK;
}

static const char* agg_name(int f) { static const char* n[] = {"SUM", "COUNT", "MIN", "MAX", "AVG"}; return n[f]; }

static olap_status upload_scalar(olap_ctx* ctx, uint64_t bits, BufPtr* out) {
  OLAP_TRY(new_buf(ctx, 16, out));
  OLAP_CUDA(cudaMemcpyAsync((*out)->p, &bits, 8, cudaMemcpyHostToDevice, ctx->streams[0]));
  OLAP_CUDA(cudaStreamSynchronize(ctx->streams[0]));
  return OLAP_OK;
}

// execute_aggregate, executor.rs:297-355.  No GROUP BY: exactly the reference's execution -- one row, every aggregate
// emitted as Float64 (:347), identities for empty input (:310-311, :339-345).  GROUP BY: the completion rule of
// key = first group_by expression as Int64, one row per distinct key, typed aggregates.
// PhysicalExpr tree -> olap_expr_prog over the columns of batch b (reverse Polish, depth <= 3, one value type).
// false: the tree does not fit a program (mixed types, too long / deep, a bare column or literal) -- evaluate it with
// the projection kernels instead.
static bool compile_prog_rec(const olap_expr* e, const Batch& b, olap_expr_prog* g, int* type, int* depth, int* max_depth) {
  if (g->n_ops >= OLAP_PROG_MAX_OPS) return false;
  if (e->kind == EXPR_COLUMN) {
    const Col* c = nullptr;
    for (const Col& x : b.cols) if (x.name == e->text) { c = &x; break; }
    if (!c && e->index < b.cols.size()) c = &b.cols[e->index];
    if (!c || !c->buf) return false;
    if (*type >= 0 && *type != c->type) return false;
    *type = c->type;
    g->op[g->n_ops] = 'c'; g->column[g->n_ops] = c->data(); g->literal_bits[g->n_ops] = 0; g->n_ops++;
    *max_depth = std::max(*max_depth, ++*depth);
    return true;
  }
  if (e->kind == EXPR_LITERAL) {
    Value v;
    if (!parse_literal(e->text, &v)) return false;
    if (*type >= 0 && *type != v.type) return false;
    *type = v.type;
    g->op[g->n_ops] = 'l'; g->column[g->n_ops] = nullptr; g->literal_bits[g->n_ops] = (int64_t)v.bits; g->n_ops++;
    *max_depth = std::max(*max_depth, ++*depth);
    return true;
  }
  if (e->text.size() != 1 || !strchr("+-*/", e->text[0])) return false;
  if (!compile_prog_rec(e->l, b, g, type, depth, max_depth)) return false;
  if (!compile_prog_rec(e->r, b, g, type, depth, max_depth)) return false;
  if (g->n_ops >= OLAP_PROG_MAX_OPS) return false;
  g->op[g->n_ops] = (int8_t)e->text[0]; g->column[g->n_ops] = nullptr; g->literal_bits[g->n_ops] = 0; g->n_ops++;
  --*depth;
  return true;
}
// The literal's type must be known before the column it is combined with is seen ("1.0 - l_discount"): find the tree's type first.
static int expr_type(const olap_expr* e, const Batch& b) {
  if (e->kind == EXPR_COLUMN) {
    for (const Col& x : b.cols) if (x.name == e->text) return x.type;
    return e->index < b.cols.size() ? b.cols[e->index].type : -1;
  }
  if (e->kind == EXPR_LITERAL) { Value v; return parse_literal(e->text, &v) ? v.type : -1; }
  const int l = expr_type(e->l, b), r = expr_type(e->r, b);
  return l == r ? l : -1;  // eval_binary_expr refuses mixed operands; the materialising path reports that error
}
static bool compile_prog(const olap_expr* e, const Batch& b, olap_expr_prog* g, int* type) {
  if (e->kind != EXPR_BINARY) return false;
  memset(g, 0, sizeof(*g));
  *type = expr_type(e, b);
  if (*type < 0) return false;
  int depth = 0, max_depth = 0;
  return compile_prog_rec(e, b, g, type, &depth, &max_depth) && depth == 1 && max_depth <= 3;
}

// Number of groups of a key column, estimated from the distinct keys d among m <= 2^17 evenly spaced rows: for G equally likely
// groups E[d] = G (1 - exp(-m / G)).  The group-by kernels take it as a hint: too low or too high costs speed (a shared
// table that spills, or L2 atomics where shared memory would do), never correctness.
// keys: the key column, or NULL with key_prog: a computed key, whose input columns are sampled at the same rows and the program run
// on the samples.
static olap_status estimate_groups(olap_ctx* ctx, const uint64_t* keys, const olap_expr_prog* key_prog, uint64_t n, uint64_t* est) {
  const uint64_t m = std::min<uint64_t>(n, 1u << 17), stride = n / m;  // 2^17 rows resolve cardinalities up to ~10^6 (d/m = 0.94 there) and cost ~50 us
  BufPtr ok, oa;
  OLAP_TRY(new_buf(ctx, m * 8, &ok));
  OLAP_TRY(new_buf(ctx, m * 8, &oa));
  olap_agg_spec cnt; cnt.values = nullptr; cnt.func = OLAP_COUNT; cnt.value_type = OLAP_I64;
  uint64_t d = 0;
  std::vector<BufPtr> samples;
  auto sample = [&](const uint64_t* col, const uint64_t** out) -> olap_status {
    BufPtr sk;
    OLAP_TRY(new_buf(ctx, m * 8, &sk));
    sample_kernel<<<small_grid(ctx, m), 256, 0, ctx->streams[0]>>>(col, stride, m, (uint64_t*)sk->p);
    count_launch(ctx);
    OLAP_CUDA(cudaGetLastError());
    samples.push_back(sk);
    *out = (const uint64_t*)sk->p;
    return OLAP_OK;
  };
  if (key_prog) {
    olap_expr_prog sp = *key_prog;
    for (int k = 0; k < sp.n_ops; k++)
      if (sp.op[k] == 'c') { const uint64_t* c = nullptr; OLAP_TRY(sample((const uint64_t*)key_prog->column[k], &c)); sp.column[k] = c; }
    OLAP_TRY(olap_group_by_expr_i64(ctx, &sp, &cnt, nullptr, 1, m, std::min<uint64_t>(m, 1u << 16), nullptr, 0, (int64_t*)ok->p, oa->p, m, nullptr, &d, nullptr));
  } else {
    const uint64_t* sk = nullptr;
    OLAP_TRY(sample(keys, &sk));
    OLAP_TRY(olap_group_by_i64(ctx, (const int64_t*)sk, &cnt, 1, m, std::min<uint64_t>(m, 1u << 16), nullptr, 0, (int64_t*)ok->p, oa->p, m, nullptr, &d, nullptr));
  }
  double g;
  if (m == n) g = (double)d;                        // the sample is the column
  else if ((double)d > 0.98 * (double)m) g = (double)n;  // (almost) every sampled key is new: no information beyond "very many"
  else {
    double lo = (double)d, hi = (double)n;          // d = G (1 - exp(-m / G)) is increasing in G
    for (int it = 0; it < 60; it++) {
      const double mid = 0.5 * (lo + hi), dd = mid * (1.0 - exp(-(double)m / mid));
      if (dd < (double)d) lo = mid; else hi = mid;
    }
    g = hi;
  }
  // No floor beyond the "+ 16": a handful of groups (Q1 has six) must be told as such -- the shared-memory regime turns a small hint
  // into MORE table copies per CTA, and with a floor of 64 groups Q1 ran on 8 copies of a 128-slot table: lanes 0, 8, 16, 24 of every
  // warp shared a copy, and the f64 atomicAdd on shared memory (a CAS loop) went round 3.6 times per add (ncu: 48 % of the stall
  // samples, profiles/r02_ncu_q1_before.txt).
  g = g * 1.25 + 16;
  *est = (uint64_t)std::min<double>((double)n, g);
  return OLAP_OK;
}

static olap_status exec_aggregate(olap_ctx* ctx, const olap_plan* p, Batches* out) {
  // Aggregate directly over a filter: the conjunction becomes fused filter terms of the group-by kernel, evaluated on the
  // unfiltered input -- the filter's output columns are never materialised (same rows aggregated, so same result).
  const olap_plan* src = p->in;
  const olap_expr* fused = nullptr;
  if (src->kind == PLAN_FILTER) { fused = src->pred; src = src->in; }
  Batches raw;
  OLAP_TRY(exec(ctx, src, &raw));
  uint64_t n = 0;
  for (const Batch& b : raw) n += b.n;
  const int na = (int)p->agg_in.size();
  std::vector<olap_filter_term> terms; std::vector<Value> keep;
  Batches in;  // zero or one batch: every input row, only the columns this node reads
  if (!raw.empty() && n > 0) {
    std::vector<bool> used(raw[0].cols.size(), false);
    referenced_columns(fused, raw[0], &used);
    for (const olap_expr* e : p->agg_in) referenced_columns(e, raw[0], &used);
    for (const olap_expr* e : p->group_by) referenced_columns(e, raw[0], &used);
    Batch all;
    OLAP_TRY(concat_columns(ctx, raw, used, n, &all));
    if (fused) {
      OLAP_TRY(collect_terms(ctx, fused, all, &terms, &keep));
      if (terms.size() > 8) {  // more conjuncts than the kernels fuse: materialise the filter after all
        terms.clear(); keep.clear(); raw.clear();
        OLAP_TRY(exec(ctx, p->in, &raw));
        n = 0;
        for (const Batch& b : raw) n += b.n;
        std::vector<bool> every(raw.empty() ? 0 : raw[0].cols.size(), true);
        if (!raw.empty() && n > 0) { OLAP_TRY(concat_columns(ctx, raw, every, n, &all)); in.push_back(all); }
      } else in.push_back(all);
    } else in.push_back(all);
  }
  if (in.empty()) n = 0;
  std::vector<Value> vals(na);
  std::vector<olap_agg_spec> specs(na + 1);
  std::vector<olap_expr_prog> prog_store(na);
  std::vector<const olap_expr_prog*> progs(na + 1, nullptr);
  std::map<std::string, int> seen;  // identical input expressions are evaluated once (SUM(x), AVG(x), MIN(x) share one column)
  // Arithmetic inputs are handed to the aggregation kernel as small programs over the input columns (no projection passes);
  // fuse_exprs = false materialises them with the projection kernels, as the reference's executor would.
  auto bind_inputs = [&](bool fuse_exprs) -> olap_status {
    seen.clear();
    for (int a = 0; a < na; a++) {
      progs[a] = nullptr;
      const std::string ename = expr_name(p->agg_in[a]);
      auto it = seen.find(ename);
      if (it != seen.end()) { vals[a] = vals[it->second]; progs[a] = progs[it->second]; }
      else if (!in.empty()) {
        int type = -1;
        if (fuse_exprs && p->agg_funcs[a] != OLAP_COUNT && compile_prog(p->agg_in[a], in[0], &prog_store[a], &type)) {
          progs[a] = &prog_store[a]; vals[a] = Value(); vals[a].type = type;
        } else OLAP_TRY(eval_over_batches(ctx, p->agg_in[a], in, n, &vals[a]));
        seen[ename] = a;
      } else { vals[a].type = OLAP_I64; OLAP_TRY(new_buf(ctx, 16, &vals[a].buf)); }
      specs[a].values = progs[a] ? nullptr : vals[a].data(); specs[a].func = p->agg_funcs[a]; specs[a].value_type = vals[a].type;
    }
    return OLAP_OK;
  };
  // grouped aggregation is bound by its table updates, so the extra arithmetic is free; the no-key kernel is a pure stream
  // and measured faster on a materialised column (Q6 shape: 0.78 ms vs 0.96 ms fused)
  OLAP_TRY(bind_inputs(!p->group_by.empty()));
  // one more COUNT: the number of rows that pass the fused filter (the reference's "empty input" cases depend on it)
  specs[na].values = nullptr; specs[na].func = OLAP_COUNT; specs[na].value_type = OLAP_I64;
  const bool grouped = !p->group_by.empty();
  Value key;
  // An arithmetic group key (Q1: GROUP BY returnflag * 2 + linestatus) becomes a program of the aggregation kernel like the value
  // expressions: no projection passes, no key column written and read back.  Keys that do not fit a program, and cardinalities
  // that need the partitioned regime (which carries plain columns), are materialised as before.
  olap_expr_prog key_prog;
  bool key_fused = false;
  auto materialise_key = [&]() -> olap_status {
    key_fused = false;
    OLAP_TRY(eval_over_batches(ctx, p->group_by[0], in, n, &key));
    if (key.type != OLAP_I64) { set_error("GROUP BY key must be Int64 (aggregate_kernel.rs:48-49)"); return OLAP_ERR_UNSUPPORTED; }
    return OLAP_OK;
  };
  if (grouped) {
    if (in.empty() || n == 0) return OLAP_OK;  // grouped aggregation of nothing: no rows
    int ktype = -1;
    if (compile_prog(p->group_by[0], in[0], &key_prog, &ktype) && ktype == OLAP_I64) key_fused = true;
    else OLAP_TRY(materialise_key());
  }
  const int ns = na < 16 ? na + 1 : na;  // the kernel seam takes at most 16 aggregates
  uint64_t hint = 1;
  if (grouped) OLAP_TRY(estimate_groups(ctx, key_fused ? nullptr : key.data(), key_fused ? &key_prog : nullptr, n, &hint));
  uint64_t cap = hint, groups = 0;
  BufPtr ok, oa;
  bool refused = false;  // at most one UNSUPPORTED retry
  for (int attempt = 0; attempt < 2; attempt++) {
    OLAP_TRY(new_buf(ctx, cap * 8, &ok));
    OLAP_TRY(new_buf(ctx, (uint64_t)ns * cap * 8, &oa));
    olap_status st;
    if (key_fused)
      st = olap_group_by_expr_i64(ctx, &key_prog, specs.data(), progs.data(), ns, n, hint, terms.empty() ? nullptr : terms.data(), (int32_t)terms.size(),
                                  (int64_t*)ok->p, oa->p, cap, nullptr, &groups, nullptr);
    else
      st = olap_group_by_prog_i64(ctx, grouped ? (const int64_t*)key.data() : nullptr, specs.data(), progs.data(), ns, n, hint,
                                  terms.empty() ? nullptr : terms.data(), (int32_t)terms.size(), (int64_t*)ok->p, oa->p, cap, nullptr, &groups, nullptr);
    if (st == OLAP_ERR_UNSUPPORTED && !refused) {  // too many groups for the regimes that fuse expressions: materialise them
      refused = true;
      OLAP_TRY(bind_inputs(false));
      if (key_fused) OLAP_TRY(materialise_key());
      attempt = -1;
      continue;
    }
    if (st == OLAP_ERR_CAPACITY && attempt == 0) { cap = hint = groups; continue; }
    OLAP_TRY(st);
    break;
  }
  if (grouped && groups == 0) return OLAP_OK;  // every row was filtered out: like a filter that drops all its batches
  Batch o; o.n = groups;
  if (grouped) { Col kc; kc.name = expr_name(p->group_by[0]); kc.type = OLAP_I64; kc.buf = ok; o.cols.push_back(kc); }
  std::vector<uint64_t> host;
  uint64_t passing = n;
  if (!grouped) {
    host.resize((size_t)ns * cap);
    OLAP_CUDA(cudaMemcpyAsync(host.data(), oa->p, host.size() * 8, cudaMemcpyDeviceToHost, ctx->streams[0]));
    OLAP_CUDA(cudaStreamSynchronize(ctx->streams[0]));
    if (ns > na) passing = host[(size_t)na * cap];
    else if (!terms.empty())  // 16 user aggregates leave no room for the hidden COUNT: count the rows that pass the fused filter separately
      OLAP_TRY(olap_filter(ctx, terms.data(), (int32_t)terms.size(), n, nullptr, nullptr, nullptr, &passing, nullptr));
  }
  for (int a = 0; a < na; a++) {
    Col c; c.name = std::string(agg_name(p->agg_funcs[a])) + "(" + expr_name(p->agg_in[a]) + ")";
    const int f = p->agg_funcs[a];
    const bool is_f = f == OLAP_AVG || (f != OLAP_COUNT && vals[a].type == OLAP_F64);
    if (grouped) {  // a view into the agg-major output (aggregate_kernel.rs:98)
      c.type = is_f ? OLAP_F64 : OLAP_I64;
      OLAP_TRY(new_buf(ctx, std::max<uint64_t>(groups, 1) * 8, &c.buf));
      if (groups) OLAP_CUDA(cudaMemcpyAsync(c.buf->p, (uint64_t*)oa->p + (uint64_t)a * cap, groups * 8, cudaMemcpyDeviceToDevice, ctx->streams[0]));
    } else {        // reference view: Float64 for everything (executor.rs:330-349)
      uint64_t bits = host[(size_t)a * cap];
      double d;
      if (is_f) memcpy(&d, &bits, 8); else d = (double)(int64_t)bits;
      if (passing == 0 && f == OLAP_MIN) d = DBL_MAX;    // f64::MAX (executor.rs:310)
      if (passing == 0 && f == OLAP_MAX) d = -DBL_MAX;   // f64::MIN (executor.rs:311)
      memcpy(&bits, &d, 8);
      c.type = OLAP_F64;
      OLAP_TRY(upload_scalar(ctx, bits, &c.buf));
    }
    o.cols.push_back(c);
  }
  OLAP_CUDA(cudaStreamSynchronize(ctx->streams[0]));
  out->push_back(o);
  return OLAP_OK;
}

static olap_status exec(olap_ctx* ctx, const olap_plan* p, Batches* out) {  // GpuExecutor::execute, executor.rs:37-104
  switch (p->kind) {
    case PLAN_SCAN: return exec_scan(ctx, p, out);
    case PLAN_FILTER: return exec_filter(ctx, p, out);
    case PLAN_PROJECTION: return exec_projection(ctx, p, out);
    case PLAN_HASH_JOIN: return exec_hash_join(ctx, p, out);
    case PLAN_AGGREGATE: return exec_aggregate(ctx, p, out);
  }
  set_error("unknown plan node");
  return OLAP_ERR_BAD_ARG;
}

// ------------------------------------------------------------------------------------------- Arrow C Data
static uint64_t fnv1a64(const uint8_t* s, uint64_t len) {  // record_batch_convert.rs:123-130
  uint64_t h = 14695981039346656037ULL;
  for (uint64_t i = 0; i < len; i++) { h ^= s[i]; h *= 1099511628211ULL; }
  return h;
}

static int src_type_of(const char* fmt) {
  if (!fmt || !fmt[0] || fmt[1]) return -1;
  switch (fmt[0]) {
    case 'c': return OLAP_SRC_I8; case 's': return OLAP_SRC_I16; case 'i': return OLAP_SRC_I32; case 'l': return OLAP_SRC_I64;
    case 'C': return OLAP_SRC_U8; case 'S': return OLAP_SRC_U16; case 'I': return OLAP_SRC_U32; case 'L': return OLAP_SRC_U64;
    case 'f': return OLAP_SRC_F32; case 'g': return OLAP_SRC_F64;
    default: return -1;
  }
}

struct ExportPriv { std::vector<void*> owned; std::vector<ArrowArray*> child_arrays; std::vector<ArrowArray> children; std::vector<const void*> bufs; };
static void release_array(ArrowArray* a) {
  if (!a || !a->release) return;
  ExportPriv* p = (ExportPriv*)a->private_data;
  if (p) { for (void* q : p->owned) free(q); delete p; }
  a->release = nullptr;
}
static void release_child(ArrowArray* a) { if (a) a->release = nullptr; }  // buffers are owned by the parent
struct SchemaPriv { std::vector<std::string> names; std::vector<ArrowSchema> children; std::vector<ArrowSchema*> ptrs; };
static void release_schema(ArrowSchema* s) {
  if (!s || !s->release) return;
  delete (SchemaPriv*)s->private_data;
  s->release = nullptr;
}
static void release_child_schema(ArrowSchema* s) { if (s) s->release = nullptr; }

}  // namespace

extern "C" {

olap_expr* olap_expr_column(uint64_t index, const char* name) { auto* e = new olap_expr{EXPR_COLUMN}; e->index = index; e->text = name ? name : ""; return e; }
olap_expr* olap_expr_literal(const char* value) { auto* e = new olap_expr{EXPR_LITERAL}; e->text = value ? value : ""; return e; }
olap_expr* olap_expr_binary(olap_expr* left, const char* op, olap_expr* right) {
  if (!left || !right) { delete left; delete right; return nullptr; }
  auto* e = new olap_expr{EXPR_BINARY}; e->text = op ? op : ""; e->l = left; e->r = right; return e;
}
void olap_expr_free(olap_expr* e) { delete e; }

olap_plan* olap_plan_table_scan(const char* table_name, const ArrowSchema* schema, const int64_t* projection, int32_t n_projection) {
  auto* p = new olap_plan{PLAN_SCAN};
  p->table = table_name ? table_name : "";
  if (schema) {
    p->has_schema = true;
    for (int64_t i = 0; i < schema->n_children; i++) {
      const ArrowSchema* c = schema->children[i];
      const bool fl = c->format && (c->format[0] == 'f' || c->format[0] == 'g' || c->format[0] == 'e') && !c->format[1];
      p->schema.push_back(Field{c->name ? c->name : "", fl});
    }
  }
  if (projection && n_projection >= 0) { p->has_proj = true; p->proj.assign(projection, projection + n_projection); }
  return p;
}
olap_plan* olap_plan_filter(olap_plan* input, olap_expr* predicate) {
  if (!input || !predicate) { delete input; delete predicate; return nullptr; }
  auto* p = new olap_plan{PLAN_FILTER}; p->in = input; p->pred = predicate; return p;
}
olap_plan* olap_plan_projection(olap_plan* input, olap_expr* const* exprs, int32_t n) {
  if (!input) return nullptr;
  auto* p = new olap_plan{PLAN_PROJECTION}; p->in = input; p->exprs.assign(exprs, exprs + (n > 0 ? n : 0)); return p;
}
olap_plan* olap_plan_hash_join(olap_plan* left, olap_plan* right, olap_expr* const* lk, int32_t nl, olap_expr* const* rk, int32_t nr, int32_t join_type) {
  if (!left || !right) { delete left; delete right; return nullptr; }
  auto* p = new olap_plan{PLAN_HASH_JOIN}; p->in = left; p->right = right; p->join_type = join_type;
  p->lkeys.assign(lk, lk + (nl > 0 ? nl : 0)); p->rkeys.assign(rk, rk + (nr > 0 ? nr : 0));
  return p;
}
olap_plan* olap_plan_aggregate(olap_plan* input, olap_expr* const* group_by, int32_t ng, const int32_t* funcs, olap_expr* const* inputs, int32_t na) {
  if (!input) return nullptr;
  for (int i = 0; i < na; i++) if (funcs[i] < OLAP_SUM || funcs[i] > OLAP_AVG) { delete input; return nullptr; }
  auto* p = new olap_plan{PLAN_AGGREGATE}; p->in = input;
  p->group_by.assign(group_by, group_by + (ng > 0 ? ng : 0));
  p->agg_in.assign(inputs, inputs + (na > 0 ? na : 0));
  p->agg_funcs.assign(funcs, funcs + (na > 0 ? na : 0));
  return p;
}
void olap_plan_free(olap_plan* p) { delete p; }

olap_status olap_join_gather_8b(olap_ctx* ctx, const void* col, const olap_pair* pairs, int32_t side, uint64_t n_pairs, void* out, void* stream) {
  OLAP_REQUIRE(ctx && (side == 0 || side == 1), "ctx is NULL / side is 0 (left) or 1 (right)");
  OLAP_REQUIRE(n_pairs == 0 || (col && pairs && out), "NULL pointer");
  OLAP_CUDA(cudaSetDevice(ctx->device));
  if (n_pairs == 0) return OLAP_OK;
  const uint64_t blocks = (n_pairs + 1023) / 1024, cap = (uint64_t)ctx->sm_count * 16;
  gather_pairs_kernel<<<(unsigned)std::min<uint64_t>(blocks, cap), 256, 0, ctx->stream_of(stream)>>>((const uint64_t*)col, (const uint32_t*)pairs, side, n_pairs,
                                                                                                     (uint64_t*)out);
  count_launch(ctx);
  OLAP_CUDA(cudaGetLastError());
  return OLAP_OK;
}

// Rows of one join side that appear in no pair, ascending: mark the partners, then the ordered selection pass of the filter
// (filter.cu) over the flags.  Composed of public calls, so no context lock is held here.
olap_status olap_join_unmatched_rows(olap_ctx* ctx, const olap_pair* pairs, uint64_t n_pairs, int32_t side, uint64_t n_rows, uint32_t* out_row_ids,
                                     uint64_t* out_count_dptr, uint64_t* host_count, void* stream) {
  OLAP_REQUIRE(ctx, "ctx is NULL");
  OLAP_REQUIRE(side == 0 || side == 1, "side is 0 (left / build rows) or 1 (right / probe rows)");
  OLAP_REQUIRE(n_pairs == 0 || pairs, "pairs is NULL");
  OLAP_REQUIRE(n_rows < 0xFFFFFFFFULL, "row ids are u32");
  OLAP_REQUIRE(n_rows == 0 || out_row_ids, "out_row_ids is NULL");
  OLAP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream_of(stream);
  if (n_rows == 0) {
    if (out_count_dptr) OLAP_CUDA(cudaMemsetAsync(out_count_dptr, 0, 8, s));
    if (host_count) { *host_count = 0; OLAP_CUDA(cudaStreamSynchronize(s)); }
    return OLAP_OK;
  }
  void* flags = nullptr;
  OLAP_TRY(olap_pool_alloc(ctx, n_rows * 8, stream, &flags));
  olap_status st = OLAP_OK;
  cudaError_t e = cudaMemsetAsync(flags, 0, n_rows * 8, s);
  if (e == cudaSuccess && n_pairs) {
    mark_pairs_kernel<<<small_grid(ctx, n_pairs), 256, 0, s>>>((const uint32_t*)pairs, side, n_pairs, (uint64_t*)flags);
    count_launch(ctx);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) st = cuda_fail(e, "marking the matched rows", __FILE__, __LINE__);
  if (st == OLAP_OK) st = olap_filter_i64(ctx, (const int64_t*)flags, n_rows, OLAP_EQ, 0, nullptr, out_row_ids, out_count_dptr, host_count, stream);
  olap_pool_free(ctx, flags, stream);   // stream-ordered: behind the selection pass
  return st;
}

static olap_status register_batch_impl(olap_ctx* ctx, const char* table_name, ArrowArray* batch, ArrowSchema* schema);

olap_status olap_register_batch(olap_ctx* ctx, const char* table_name, ArrowArray* batch, ArrowSchema* schema) {
  OLAP_REQUIRE(ctx && table_name && batch && schema, "NULL argument");
  const olap_status st = register_batch_impl(ctx, table_name, batch, schema);
  // the array is consumed whatever happened: the copies above were synchronised (or never started), so the producer may free it
  if (st != OLAP_OK) cudaStreamSynchronize(ctx->streams[0]);
  if (batch->release) batch->release(batch);
  return st;
}

static olap_status register_batch_impl(olap_ctx* ctx, const char* table_name, ArrowArray* batch, ArrowSchema* schema) {
  OLAP_REQUIRE(schema->format && strcmp(schema->format, "+s") == 0, "a RecordBatch is a struct-typed ArrowArray (format \"+s\")");
  OLAP_REQUIRE(batch->n_children == schema->n_children, "array/schema children mismatch");
  Batch b; b.n = (uint64_t)batch->length;
  for (int64_t i = 0; i < batch->n_children; i++) {
    const ArrowArray* a = batch->children[i];
    const ArrowSchema* s = schema->children[i];
    Col c; c.name = s->name ? s->name : "";
    OLAP_REQUIRE((uint64_t)a->length == b.n && batch->offset == 0, "child length / parent offset not supported");
    const uint8_t* validity = a->null_count != 0 && a->n_buffers > 0 ? (const uint8_t*)a->buffers[0] : nullptr;
    OLAP_TRY(new_buf(ctx, b.n * 8, &c.buf));
    const int st = src_type_of(s->format);
    if (st >= 0) {
      c.type = st >= OLAP_SRC_F32 ? OLAP_F64 : OLAP_I64;
      if (validity && b.n) OLAP_TRY(new_buf(ctx, b.n, &c.valid));
      OLAP_TRY(olap_column_upload(ctx, a->buffers[1], b.n, st, validity, (uint64_t)a->offset, c.buf->p, c.validity(), nullptr));
    } else if (s->format && (strcmp(s->format, "u") == 0 || strcmp(s->format, "U") == 0)) {
      // Utf8 -> FNV-1a-64 as i64, nulls -> 0 (record_batch_convert.rs:93-97, :112-121); hashed on the host: strings are not on the hot path
      c.type = OLAP_I64;
      std::vector<int64_t> h(b.n);
      std::vector<uint8_t> hv(validity ? b.n : 0);
      const bool large = s->format[0] == 'U';
      const uint8_t* data = (const uint8_t*)a->buffers[2];
      for (uint64_t r = 0; r < b.n; r++) {
        const uint64_t idx = r + (uint64_t)a->offset;
        const bool valid = !validity || ((validity[idx >> 3] >> (idx & 7)) & 1);
        int64_t o0 = large ? ((const int64_t*)a->buffers[1])[idx] : ((const int32_t*)a->buffers[1])[idx];
        int64_t o1 = large ? ((const int64_t*)a->buffers[1])[idx + 1] : ((const int32_t*)a->buffers[1])[idx + 1];
        h[r] = valid ? (int64_t)fnv1a64(data + o0, (uint64_t)(o1 - o0)) : 0;
        if (validity) hv[r] = valid ? 1 : 0;
      }
      OLAP_TRY(olap_memcpy_h2d(ctx, c.buf->p, h.data(), b.n * 8, nullptr));
      if (validity && b.n) {
        OLAP_TRY(new_buf(ctx, b.n, &c.valid));
        OLAP_TRY(olap_memcpy_h2d(ctx, c.valid->p, hv.data(), b.n, nullptr));
      }
      OLAP_TRY(olap_stream_synchronize(ctx, nullptr));
    } else {
      set_error("Unsupported column type for GPU conversion: format '%s' (column '%s')", s->format ? s->format : "?", c.name.c_str());
      return OLAP_ERR_UNSUPPORTED;
    }
    b.cols.push_back(c);
  }
  OLAP_TRY(olap_stream_synchronize(ctx, nullptr));  // the Arrow buffers may go away once the caller releases the array
  Catalog* cat = catalog_of(ctx);
  std::lock_guard<std::mutex> g(cat->mu);
  cat->tables[table_name].push_back(b);
  return OLAP_OK;
}

olap_status olap_register_device_batch(olap_ctx* ctx, const char* table_name, int32_t n_cols, const char* const* names, const int32_t* types,
                                       const void* const* columns_dptr, const uint8_t* const* validity_dptr, uint64_t n_rows) {
  OLAP_REQUIRE(ctx && table_name && n_cols >= 1 && names && types && columns_dptr, "NULL argument");
  OLAP_CUDA(cudaSetDevice(ctx->device));
  Batch b; b.n = n_rows;
  for (int i = 0; i < n_cols; i++) {
    OLAP_REQUIRE(names[i] && (n_rows == 0 || columns_dptr[i]), "NULL column name / pointer");
    OLAP_REQUIRE(types[i] == OLAP_I64 || types[i] == OLAP_F64, "column type must be OLAP_I64 or OLAP_F64");
    Col c; c.name = names[i]; c.type = types[i];
    OLAP_TRY(new_buf(ctx, std::max<uint64_t>(n_rows, 2) * 8, &c.buf));
    if (n_rows) OLAP_CUDA(cudaMemcpyAsync(c.buf->p, columns_dptr[i], n_rows * 8, cudaMemcpyDeviceToDevice, ctx->streams[0]));
    if (validity_dptr && validity_dptr[i] && n_rows) {
      OLAP_TRY(new_buf(ctx, n_rows, &c.valid));
      OLAP_CUDA(cudaMemcpyAsync(c.valid->p, validity_dptr[i], n_rows, cudaMemcpyDeviceToDevice, ctx->streams[0]));
    }
    b.cols.push_back(c);
  }
  OLAP_CUDA(cudaStreamSynchronize(ctx->streams[0]));  // the caller may reuse its buffers on return
  Catalog* cat = catalog_of(ctx);
  std::lock_guard<std::mutex> g(cat->mu);
  cat->tables[table_name].push_back(b);
  return OLAP_OK;
}

olap_status olap_drop_table(olap_ctx* ctx, const char* table_name) {
  OLAP_REQUIRE(ctx && table_name, "NULL argument");
  Catalog* cat = catalog_of(ctx);
  std::lock_guard<std::mutex> g(cat->mu);
  cat->tables.erase(table_name);
  return OLAP_OK;
}

olap_status olap_execute(olap_ctx* ctx, const olap_plan* plan, olap_result** out) {
  OLAP_REQUIRE(ctx && plan && out, "NULL argument");
  OLAP_CUDA(cudaSetDevice(ctx->device));
  auto* r = new olap_result();
  olap_status st = exec(ctx, plan, &r->batches);
  if (st == OLAP_OK) { cudaError_t e = cudaStreamSynchronize(ctx->streams[0]); if (e != cudaSuccess) st = cuda_fail(e, "sync", __FILE__, __LINE__); }
  if (st != OLAP_OK) { delete r; return st; }
  *out = r;
  return OLAP_OK;
}

int32_t olap_result_num_batches(const olap_result* r) { return r ? (int32_t)r->batches.size() : 0; }
int64_t olap_result_num_rows(const olap_result* r, int32_t i) { return (r && i >= 0 && (size_t)i < r->batches.size()) ? (int64_t)r->batches[i].n : -1; }

olap_status olap_result_export(olap_result* r, int32_t i, ArrowArray* out_array, ArrowSchema* out_schema) {
  OLAP_REQUIRE(r && out_array && out_schema && i >= 0 && (size_t)i < r->batches.size(), "bad batch index");
  const Batch& b = r->batches[i];
  const size_t nc = b.cols.size();
  auto* ap = new ExportPriv(); auto* sp = new SchemaPriv();
  ap->children.resize(nc); ap->child_arrays.resize(nc); ap->bufs.resize(2 * nc + 1);
  sp->children.resize(nc); sp->ptrs.resize(nc); sp->names.resize(nc);
  for (size_t c = 0; c < nc; c++) {
    void* host = malloc(b.n ? b.n * 8 : 8);
    ap->owned.push_back(host);
    if (b.n) {
      olap_ctx* ctx = b.cols[c].buf->ctx;
      cudaError_t e = cudaMemcpyAsync(host, b.cols[c].data(), b.n * 8, cudaMemcpyDeviceToHost, ctx->streams[0]);
      if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->streams[0]);
      if (e != cudaSuccess) { for (void* q : ap->owned) free(q); delete ap; delete sp; return cuda_fail(e, "export copy", __FILE__, __LINE__); }
    }
    ap->bufs[2 * c] = nullptr; ap->bufs[2 * c + 1] = host;
    int64_t nulls = 0;
    if (b.cols[c].valid && b.n) {  // byte-per-row validity -> Arrow bitmap (LSB first), packed on the host
      olap_ctx* ctx = b.cols[c].buf->ctx;
      std::vector<uint8_t> bytes(b.n);
      cudaError_t e = cudaMemcpyAsync(bytes.data(), b.cols[c].validity(), b.n, cudaMemcpyDeviceToHost, ctx->streams[0]);
      if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->streams[0]);
      if (e != cudaSuccess) { for (void* q : ap->owned) free(q); delete ap; delete sp; return cuda_fail(e, "export copy", __FILE__, __LINE__); }
      uint8_t* bits = (uint8_t*)calloc((b.n + 7) / 8 + 8, 1);
      ap->owned.push_back(bits);
      for (uint64_t r = 0; r < b.n; r++) { if (bytes[r]) bits[r >> 3] |= (uint8_t)(1u << (r & 7)); else nulls++; }
      if (nulls) ap->bufs[2 * c] = bits;
    }
    ArrowArray& a = ap->children[c];
    a.length = (int64_t)b.n; a.null_count = nulls; a.offset = 0; a.n_buffers = 2; a.n_children = 0;
    a.buffers = &ap->bufs[2 * c]; a.children = nullptr; a.dictionary = nullptr; a.release = release_child; a.private_data = nullptr;
    ap->child_arrays[c] = &a;
    sp->names[c] = b.cols[c].name;
    ArrowSchema& s = sp->children[c];
    s.format = b.cols[c].type == OLAP_F64 ? "g" : "l"; s.name = sp->names[c].c_str(); s.metadata = nullptr; s.flags = b.cols[c].valid ? 2 /* ARROW_FLAG_NULLABLE */ : 0; s.n_children = 0;
    s.children = nullptr; s.dictionary = nullptr; s.release = release_child_schema; s.private_data = nullptr;
    sp->ptrs[c] = &s;
  }
  ap->bufs[2 * nc] = nullptr;
  out_array->length = (int64_t)b.n; out_array->null_count = 0; out_array->offset = 0; out_array->n_buffers = 1; out_array->n_children = (int64_t)nc;
  out_array->buffers = &ap->bufs[2 * nc]; out_array->children = nc ? ap->child_arrays.data() : nullptr; out_array->dictionary = nullptr;
  out_array->release = release_array; out_array->private_data = ap;
  out_schema->format = "+s"; out_schema->name = ""; out_schema->metadata = nullptr; out_schema->flags = 0; out_schema->n_children = (int64_t)nc;
  out_schema->children = nc ? sp->ptrs.data() : nullptr; out_schema->dictionary = nullptr; out_schema->release = release_schema; out_schema->private_data = sp;
  return OLAP_OK;
}

void olap_result_free(olap_result* r) { delete r; }

}  // extern "C"
