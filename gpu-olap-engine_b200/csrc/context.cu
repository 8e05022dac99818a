// Context lifecycle, stream-ordered device pool, transfers, Arrow-buffer upload with on-device widening,
// synthetic data generation.  Replaces gpu-memory-manager (SlabAllocator, TransferQueue) and the host-side
// widening copy of arrow-interop/src/record_batch_convert.rs:102-110.
#include "common.cuh"

#include <cstdarg>

namespace olap {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

olap_status cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
  return e == cudaErrorMemoryAllocation ? OLAP_ERR_OOM : OLAP_ERR_CUDA;
}

olap_status pool_alloc(olap_ctx* ctx, uint64_t bytes, cudaStream_t s, void** out) {
  if (bytes == 0) bytes = 16;
  if (ctx->max_bytes && ctx->used + bytes > ctx->max_bytes) {
    set_error("device pool cap exceeded: %llu in use + %llu requested > max_gpu_memory %llu", (unsigned long long)ctx->used,
              (unsigned long long)bytes, (unsigned long long)ctx->max_bytes);
    return OLAP_ERR_OOM;
  }
  void* p = nullptr;
  cudaError_t e = cudaMallocFromPoolAsync(&p, bytes, ctx->pool, s);
  if (e != cudaSuccess) { cudaGetLastError(); return cuda_fail(e, "cudaMallocFromPoolAsync", __FILE__, __LINE__); }
  ctx->live[p] = bytes;
  ctx->used += bytes;
  if (ctx->used > ctx->high_water) ctx->high_water = ctx->used;
  *out = p;
  return OLAP_OK;
}

olap_status pool_free(olap_ctx* ctx, void* p, cudaStream_t s) {
  if (!p) return OLAP_OK;
  auto it = ctx->live.find(p);
  if (it == ctx->live.end()) { set_error("pool_free: pointer %p was not allocated from this context", p); return OLAP_ERR_BAD_ARG; }
  ctx->used -= it->second;
  ctx->live.erase(it);
  ctx->key_stats.erase(p);  // statistics remembered for a column die with its allocation
  ctx->shard_key_stats.erase(p);
  OLAP_CUDA(cudaFreeAsync(p, s));
  return OLAP_OK;
}

void phase_begin(olap_ctx* ctx, cudaStream_t s) {
  if (!ctx->profiling) return;
  for (auto& m : ctx->marks) cudaEventDestroy(m.second);
  ctx->marks.clear();
  phase_mark(ctx, s, "begin");
}
void phase_mark(olap_ctx* ctx, cudaStream_t s, const char* name) {
  if (!ctx->profiling) return;
  cudaEvent_t ev;
  cudaEventCreate(&ev);
  cudaEventRecord(ev, s);
  ctx->marks.emplace_back(name, ev);
}
void phase_end(olap_ctx* ctx, cudaStream_t s) {
  if (!ctx->profiling) return;
  phase_mark(ctx, s, "end");
}

olap_status pack_terms(const olap_filter_term* terms, int32_t n_terms, TermPack* out) {
  OLAP_REQUIRE(n_terms >= 0 && n_terms <= kMaxTerms, "n_terms must be in [0, 8]");
  OLAP_REQUIRE(n_terms == 0 || terms, "terms is NULL");
  memset(out, 0, sizeof(*out));
  out->n = n_terms;
  for (int k = 0; k < n_terms; k++) {
    OLAP_REQUIRE(terms[k].lhs, "filter term lhs column is NULL");
    OLAP_REQUIRE(terms[k].type == OLAP_I64 || terms[k].type == OLAP_F64, "filter term type must be OLAP_I64 or OLAP_F64");
    out->lhs[k] = (const uint64_t*)terms[k].lhs;
    out->rhs[k] = (const uint64_t*)terms[k].rhs;
    out->lit[k] = terms[k].literal_bits;
    out->op[k] = (terms[k].op >= OLAP_EQ && terms[k].op <= OLAP_GE) ? (int8_t)terms[k].op : (int8_t)-1;
    out->type[k] = (int8_t)terms[k].type;
  }
  return OLAP_OK;
}

// ------------------------------------------------------------------------------------------- widening kernel
// One pass over an uploaded raw Arrow values buffer: widen to 8 bytes, zero nulls, expand validity bits to bytes.
template <class Src, class Dst>
__global__ void __launch_bounds__(256) widen_kernel(const Src* __restrict__ src, const uint8_t* __restrict__ bitmap,
                                                    uint64_t bit_offset, uint64_t n, Dst* __restrict__ out,
                                                    uint8_t* __restrict__ out_valid) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    bool valid = true;
    if (bitmap) { uint64_t b = bit_offset + i; valid = (bitmap[b >> 3] >> (b & 7)) & 1; }
    out[i] = valid ? (Dst)src[i] : (Dst)0;
    if (out_valid) out_valid[i] = valid;
  }
}

__device__ __forceinline__ uint64_t perm_pow2(uint64_t x, int bits, uint64_t k0, uint64_t k1, uint64_t k2, uint64_t k3) {
  const uint64_t mask = bits == 64 ? ~0ULL : ((1ULL << bits) - 1);
  const int sh = (bits >> 1) + (bits < 2);
  x = (x * 0x9e3779b97f4a7c15ULL + k0) & mask; x ^= x >> sh;
  x = (x * 0xbf58476d1ce4e5b9ULL + k1) & mask; x ^= x >> (sh + 1);
  x = (x * 0x94d049bb133111ebULL + k2) & mask; x ^= x >> sh;
  x = (x * 0xff51afd7ed558ccdULL + k3) & mask; x ^= x >> (sh + 1);
  return x;
}

__global__ void __launch_bounds__(256) generate_kernel(uint64_t* __restrict__ out, uint64_t n, int kind, uint64_t seed,
                                                       uint64_t row0, uint64_t domain, int bits) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t k0 = splitmix64(seed, 0), k1 = splitmix64(seed, 1), k2 = splitmix64(seed, 2), k3 = splitmix64(seed, 3);
  for (; i < n; i += stride) {
    uint64_t r = row0 + i, v;
    if (kind == 0) v = splitmix64(seed, r) % domain;
    else if (kind == 1) { v = r; do { v = perm_pow2(v, bits, k0, k1, k2, k3); } while (v >= domain); }
    else if (kind == 2) v = (uint64_t)__double_as_longlong((double)(splitmix64(seed, r) % domain) / 1024.0);
    else v = (uint64_t)__double_as_longlong((double)(splitmix64(seed, r) >> 11) * (1.0 / 9007199254740992.0));
    out[i] = v;
  }
}

// Foreign keys into a permutation-keyed build side (kind 1 with the same seed / domain): row r of the probe side references
// build row b(r) and carries that row's key perm(b).  b is uniform (zipf_s = 0) or b + 1 ~ Zipf(zipf_s) over the build rows, drawn
// with the continuous inverse CDF ((N^(1-s) - 1) u + 1)^(1 / (1 - s)) (N^u for s = 1): the hot build rows are 0, 1, 2, ..., whose
// keys the permutation spreads over the whole domain ("Zipf(s) over ranks mapped through a fixed permutation").
__global__ void __launch_bounds__(256) generate_fk_kernel(uint64_t* __restrict__ out, uint64_t n, uint64_t perm_seed, uint64_t domain, int bits,
                                                          uint64_t fk_seed, uint64_t row0, double zipf_s) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t k0 = splitmix64(perm_seed, 0), k1 = splitmix64(perm_seed, 1), k2 = splitmix64(perm_seed, 2), k3 = splitmix64(perm_seed, 3);
  const double N = (double)domain, e = 1.0 - zipf_s;
  const double a = zipf_s > 0.0 ? (fabs(e) < 1e-9 ? log(N) : pow(N, e) - 1.0) : 0.0;
  for (; i < n; i += stride) {
    const uint64_t h = splitmix64(fk_seed, row0 + i);
    uint64_t b;
    if (zipf_s <= 0.0) b = h % domain;
    else {
      const double u = (double)(h >> 11) * (1.0 / 9007199254740992.0);
      const double x = fabs(e) < 1e-9 ? exp(u * a) : pow(a * u + 1.0, 1.0 / e);
      b = (uint64_t)x;
      b = b ? b - 1 : 0;
      if (b >= domain) b = domain - 1;
    }
    uint64_t v = b;
    do { v = perm_pow2(v, bits, k0, k1, k2, k3); } while (v >= domain);
    out[i] = v;
  }
}

static inline unsigned grid_for(uint64_t n, int block, int sms, int per_sm = 8) {
  uint64_t g = (n + block - 1) / block;
  uint64_t cap = (uint64_t)sms * per_sm;
  if (g > cap) g = cap;
  return (unsigned)(g ? g : 1);
}

}  // namespace olap

using namespace olap;

extern "C" {

const char* olap_last_error(void) { return g_err; }
const char* olap_version(void) { return "olap_b200 0.1 (sm_100a)"; }

olap_status olap_ctx_create(int32_t device, uint64_t max_gpu_memory, int32_t num_streams, olap_ctx** out) {
  OLAP_REQUIRE(out, "out is NULL");
  int count = 0;
  OLAP_CUDA(cudaGetDeviceCount(&count));
  if (device < 0 || device >= count) { set_error("device ordinal %d out of range (%d devices)", device, count); return OLAP_ERR_BAD_ARG; }
  OLAP_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  OLAP_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("olap_b200 is built for sm_100a only; device %d is sm_%d%d (%s) and there is no fallback path", device, prop.major,
              prop.minor, prop.name);
    return OLAP_ERR_UNSUPPORTED;
  }
  olap_ctx* ctx = new olap_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->smem_optin = prop.sharedMemPerBlockOptin;
  ctx->max_bytes = max_gpu_memory;
  cudaMemPoolProps pp;
  memset(&pp, 0, sizeof(pp));
  pp.allocType = cudaMemAllocationTypePinned;
  pp.handleTypes = cudaMemHandleTypeNone;
  pp.location.type = cudaMemLocationTypeDevice;
  pp.location.id = device;
  cudaError_t e = cudaMemPoolCreate(&ctx->pool, &pp);
  if (e != cudaSuccess) { delete ctx; return cuda_fail(e, "cudaMemPoolCreate", __FILE__, __LINE__); }
  uint64_t keep = ~0ULL;  // never trim: freed blocks stay in the pool for the next operator
  cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &keep);
  if (num_streams < 1) num_streams = 1;
  if (num_streams > 64) num_streams = 64;
  ctx->streams.resize(num_streams);
  for (auto& s : ctx->streams) {
    e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete ctx; return cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__); }
  }
  cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->arena_ev, cudaEventDisableTiming);
  for (auto& ev : ctx->join_ev) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
  *out = ctx;
  return OLAP_OK;
}

olap_status olap_ctx_destroy(olap_ctx* ctx) {
  if (!ctx) return OLAP_OK;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  if (ctx->op_state && ctx->op_state_free) { ctx->op_state_free(ctx->op_state); ctx->op_state = nullptr; }  // frees its pool buffers first
  for (auto& m : ctx->marks) cudaEventDestroy(m.second);
  if (ctx->table_arena) cudaFree(ctx->table_arena);
  for (auto& kv : ctx->live) cudaFree(kv.first);
  for (auto s : ctx->streams) cudaStreamDestroy(s);
  if (ctx->fork_ev) cudaEventDestroy(ctx->fork_ev);
  if (ctx->arena_ev) cudaEventDestroy(ctx->arena_ev);
  for (auto ev : ctx->join_ev) if (ev) cudaEventDestroy(ev);
  if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
  delete ctx;
  return OLAP_OK;
}

olap_status olap_ctx_synchronize(olap_ctx* ctx) {
  OLAP_REQUIRE(ctx, "ctx is NULL");
  OLAP_CUDA(cudaSetDevice(ctx->device));
  OLAP_CUDA(cudaDeviceSynchronize());
  return OLAP_OK;
}

void* olap_ctx_stream(olap_ctx* ctx, int32_t stream_id) {
  if (!ctx) return nullptr;
  uint32_t n = (uint32_t)ctx->streams.size();
  return (void*)ctx->streams[(uint32_t)stream_id % n];
}

olap_status olap_pool_alloc(olap_ctx* ctx, uint64_t bytes, void* stream, void** out_dptr) {
  OLAP_REQUIRE(ctx && out_dptr, "ctx/out is NULL");
  std::lock_guard<std::mutex> g(ctx->mu);
  OLAP_CUDA(cudaSetDevice(ctx->device));
  return pool_alloc(ctx, bytes, ctx->stream_of(stream), out_dptr);
}

olap_status olap_pool_free(olap_ctx* ctx, void* dptr, void* stream) {
  OLAP_REQUIRE(ctx, "ctx is NULL");
  std::lock_guard<std::mutex> g(ctx->mu);
  OLAP_CUDA(cudaSetDevice(ctx->device));
  return pool_free(ctx, dptr, ctx->stream_of(stream));
}

olap_status olap_pool_stats(olap_ctx* ctx, uint64_t* used, uint64_t* high, uint64_t* reserved) {
  OLAP_REQUIRE(ctx, "ctx is NULL");
  std::lock_guard<std::mutex> g(ctx->mu);
  if (used) *used = ctx->used;
  if (high) *high = ctx->high_water;
  if (reserved) { uint64_t r = 0; cudaMemPoolGetAttribute(ctx->pool, cudaMemPoolAttrReservedMemCurrent, &r); *reserved = r; }
  return OLAP_OK;
}

olap_status olap_host_alloc(uint64_t bytes, void** out) {
  OLAP_REQUIRE(out, "out is NULL");
  OLAP_CUDA(cudaHostAlloc(out, bytes ? bytes : 16, cudaHostAllocDefault));
  return OLAP_OK;
}
olap_status olap_host_free(void* p) {
  if (p) OLAP_CUDA(cudaFreeHost(p));
  return OLAP_OK;
}

olap_status olap_memcpy_h2d(olap_ctx* ctx, void* dptr, const void* hptr, uint64_t bytes, void* stream) {
  OLAP_REQUIRE(ctx && (bytes == 0 || (dptr && hptr)), "NULL pointer");
  OLAP_CUDA(cudaSetDevice(ctx->device));
  if (bytes) OLAP_CUDA(cudaMemcpyAsync(dptr, hptr, bytes, cudaMemcpyHostToDevice, ctx->stream_of(stream)));
  return OLAP_OK;
}
olap_status olap_memcpy_d2h(olap_ctx* ctx, void* hptr, const void* dptr, uint64_t bytes, void* stream) {
  OLAP_REQUIRE(ctx && (bytes == 0 || (dptr && hptr)), "NULL pointer");
  OLAP_CUDA(cudaSetDevice(ctx->device));
  if (bytes) OLAP_CUDA(cudaMemcpyAsync(hptr, dptr, bytes, cudaMemcpyDeviceToHost, ctx->stream_of(stream)));
  return OLAP_OK;
}
olap_status olap_stream_synchronize(olap_ctx* ctx, void* stream) {
  OLAP_REQUIRE(ctx, "ctx is NULL");
  OLAP_CUDA(cudaStreamSynchronize(ctx->stream_of(stream)));
  return OLAP_OK;
}

olap_status olap_column_upload(olap_ctx* ctx, const void* host_values, uint64_t n, int32_t src_type, const uint8_t* bitmap,
                               uint64_t bit_offset, void* out_dptr, uint8_t* out_valid, void* stream) {
  OLAP_REQUIRE(ctx && (n == 0 || (host_values && out_dptr)), "NULL pointer");
  OLAP_REQUIRE(src_type >= OLAP_SRC_I8 && src_type <= OLAP_SRC_F64, "unsupported Arrow source type");
  std::lock_guard<std::mutex> g(ctx->mu);
  OLAP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream_of(stream);
  if (n == 0) return OLAP_OK;
  static const int width[] = {1, 2, 4, 8, 1, 2, 4, 8, 4, 8};
  const int w = width[src_type];
  const char* src = (const char*)host_values + bit_offset * w;
  const bool straight = (w == 8) && !bitmap;
  if (straight) {  // Int64 / UInt64 / Float64 without nulls: the Arrow buffer IS the GPU layout
    OLAP_CUDA(cudaMemcpyAsync(out_dptr, src, n * 8, cudaMemcpyHostToDevice, s));
    if (out_valid) OLAP_CUDA(cudaMemsetAsync(out_valid, 1, n, s));
    return OLAP_OK;
  }
  Scratch sc(ctx, s);
  char* raw = nullptr;
  uint8_t* dbits = nullptr;
  OLAP_TRY(sc.get<char>(n * w, &raw));
  OLAP_CUDA(cudaMemcpyAsync(raw, src, n * w, cudaMemcpyHostToDevice, s));
  uint64_t bit_lo = 0;
  if (bitmap) {
    uint64_t byte0 = bit_offset >> 3, byte1 = (bit_offset + n + 7) >> 3;
    OLAP_TRY(sc.get<uint8_t>(byte1 - byte0, &dbits));
    OLAP_CUDA(cudaMemcpyAsync(dbits, bitmap + byte0, byte1 - byte0, cudaMemcpyHostToDevice, s));
    bit_lo = bit_offset & 7;
  }
  unsigned grid = grid_for(n, 256, ctx->sm_count);
#define W(S, D) widen_kernel<S, D><<<grid, 256, 0, s>>>((const S*)raw, dbits, bit_lo, n, (D*)out_dptr, out_valid)
  switch (src_type) {
    case OLAP_SRC_I8: W(int8_t, int64_t); break;
    case OLAP_SRC_I16: W(int16_t, int64_t); break;
    case OLAP_SRC_I32: W(int32_t, int64_t); break;
    case OLAP_SRC_I64: W(int64_t, int64_t); break;
    case OLAP_SRC_U8: W(uint8_t, int64_t); break;
    case OLAP_SRC_U16: W(uint16_t, int64_t); break;
    case OLAP_SRC_U32: W(uint32_t, int64_t); break;
    case OLAP_SRC_U64: W(uint64_t, uint64_t); break;  // u64 as i64 reinterprets (record_batch_convert.rs:80)
    case OLAP_SRC_F32: W(float, double); break;
    default: W(double, double); break;
  }
#undef W
  count_launch(ctx);
  OLAP_CUDA(cudaGetLastError());
  return OLAP_OK;
}

olap_status olap_generate(olap_ctx* ctx, void* out, uint64_t n, int32_t kind, uint64_t seed, uint64_t row0, uint64_t domain,
                          void* stream) {
  OLAP_REQUIRE(ctx && (n == 0 || out), "NULL pointer");
  OLAP_REQUIRE(kind >= 0 && kind <= 3, "kind must be 0..3");
  OLAP_REQUIRE(kind == 3 || domain > 0, "domain must be > 0");
  std::lock_guard<std::mutex> g(ctx->mu);
  OLAP_CUDA(cudaSetDevice(ctx->device));
  if (n == 0) return OLAP_OK;
  int bits = 1;
  while (bits < 64 && (1ULL << bits) < domain) bits++;
  generate_kernel<<<grid_for(n, 256, ctx->sm_count), 256, 0, ctx->stream_of(stream)>>>((uint64_t*)out, n, kind, seed, row0, domain, bits);
  count_launch(ctx);
  OLAP_CUDA(cudaGetLastError());
  return OLAP_OK;
}

olap_status olap_generate_fk(olap_ctx* ctx, void* out, uint64_t n, uint64_t perm_seed, uint64_t domain, uint64_t fk_seed, uint64_t row0, double zipf_s,
                             void* stream) {
  OLAP_REQUIRE(ctx && (n == 0 || out), "NULL pointer");
  OLAP_REQUIRE(domain > 0 && zipf_s >= 0.0, "domain must be > 0, zipf_s >= 0");
  std::lock_guard<std::mutex> g(ctx->mu);
  OLAP_CUDA(cudaSetDevice(ctx->device));
  if (n == 0) return OLAP_OK;
  int bits = 1;
  while (bits < 64 && (1ULL << bits) < domain) bits++;
  generate_fk_kernel<<<grid_for(n, 256, ctx->sm_count), 256, 0, ctx->stream_of(stream)>>>((uint64_t*)out, n, perm_seed, domain, bits, fk_seed, row0, zipf_s);
  count_launch(ctx);
  OLAP_CUDA(cudaGetLastError());
  return OLAP_OK;
}

olap_status olap_set_profiling(olap_ctx* ctx, int32_t enabled) {
  OLAP_REQUIRE(ctx, "ctx is NULL");
  ctx->profiling = enabled != 0;
  return OLAP_OK;
}

olap_status olap_last_phase_times(olap_ctx* ctx, const char** names, const float** ms, int32_t* n) {
  OLAP_REQUIRE(ctx && names && ms && n, "NULL pointer");
  std::lock_guard<std::mutex> g(ctx->mu);
  ctx->phase_names.clear();
  ctx->phase_ms.clear();
  for (size_t i = 1; i < ctx->marks.size(); i++) {
    OLAP_CUDA(cudaEventSynchronize(ctx->marks[i].second));
    float t = 0;
    OLAP_CUDA(cudaEventElapsedTime(&t, ctx->marks[i - 1].second, ctx->marks[i].second));
    if (!ctx->phase_names.empty()) ctx->phase_names += ";";
    ctx->phase_names += ctx->marks[i].first;
    ctx->phase_ms.push_back(t);
  }
  *names = ctx->phase_names.c_str();
  *ms = ctx->phase_ms.data();
  *n = (int32_t)ctx->phase_ms.size();
  return OLAP_OK;
}

uint64_t olap_launch_count(olap_ctx* ctx) { return ctx ? ctx->launches : 0; }

int64_t olap_agg_identity(int32_t func) {
  switch (func) { case OLAP_MIN: return INT64_MAX; case OLAP_MAX: return INT64_MIN; default: return 0; }
}

void olap_compute_launch_params(uint64_t n_rows, uint32_t* grid, uint32_t* block) {
  const uint64_t B = 256;
  uint64_t g = (n_rows + B - 1) / B;
  if (g < 1) g = 1;
  if (g > 0x7fffffffULL) g = 0x7fffffffULL;
  if (grid) *grid = (uint32_t)g;
  if (block) *block = (uint32_t)B;
}

int32_t olap_choose_join_strategy(uint64_t n_build) { return n_build <= 1000000ULL ? OLAP_JOIN_BROADCAST : OLAP_JOIN_RADIX; }

uint32_t olap_partition_of(int64_t key, int32_t bits) {
  if (bits <= 0) return 0;
  const uint64_t k = (uint64_t)key;  // must match part_hash(key, 0) in join.cu
  return (uint32_t)(((k ^ (k >> 29)) * 0x9E3779B97F4A7C15ULL) >> (64 - bits));
}

}  // extern "C"
