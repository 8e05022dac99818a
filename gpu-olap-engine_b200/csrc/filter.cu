// Fused predicate + mask + order-preserving stream compaction (one pass, decoupled look-back), gather and
// projection kernels.  Replaces eval_filter_mask / filter_record_batch / take_int64 / eval_binary_expr
// (gpu-olap-core/src/executor.rs:411-507) and the JIT'd filter_kernel (gpu-kernel-compiler/src/codegen.rs:36-47).
//
// Roofline: HBM.  Algorithmic bytes per row = 8 B per distinct predicate column read + 1 B mask (if asked) +
// 4 B per passing row for the selection vector (+ 16 B per passing row per compacted column).
#include "common.cuh"

namespace olap {

constexpr int kFilterThreads = 256;
constexpr int kFilterPairs = 4;                                   // 128-bit loads (2 rows) per thread per tile
constexpr int kFilterTile = kFilterThreads * kFilterPairs * 2;    // 2048 rows per tile
constexpr int kMaxCompactCols = 8;

struct CompactCols {
  const uint64_t* in[kMaxCompactCols];
  uint64_t* out[kMaxCompactCols];
  int n;
};

// tile descriptor for the look-back: [63:62] status (0 = not ready, 1 = tile aggregate, 2 = inclusive prefix)
constexpr uint64_t kStAgg = 1ULL << 62, kStPrefix = 2ULL << 62, kStMask = (1ULL << 62) - 1;

__device__ __forceinline__ uint64_t ld_volatile_u64(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Exclusive prefix of this tile's aggregate over all earlier tiles; called by warp 0, result valid in all its lanes.
__device__ __forceinline__ uint64_t lookback_exclusive(uint64_t* status, uint64_t tile, uint64_t aggregate, int lane) {
  if (lane == 0) {
    uint64_t w = (tile == 0 ? kStPrefix : kStAgg) | aggregate;
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(status + tile), "l"(w) : "memory");
  }
  uint64_t excl = 0;
  int64_t base = (int64_t)tile - 1;
  while (base >= 0) {
    int64_t idx = base - lane;
    uint64_t st = kStPrefix;  // tiles before 0 contribute a zero prefix
    if (idx >= 0) {
      do { st = ld_volatile_u64(status + idx); } while ((st >> 62) == 0);
    }
    unsigned pm = __ballot_sync(0xffffffffu, (st >> 62) == 2);
    int first = pm ? (__ffs(pm) - 1) : 32;
    uint64_t v = (lane <= first) ? (st & kStMask) : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    excl += v;
    if (pm) break;
    base -= 32;
  }
  if (lane == 0 && tile != 0) {
    uint64_t w = kStPrefix | (excl + aggregate);
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(status + tile), "l"(w) : "memory");
  }
  return excl;
}

// work[0] = tile ticket, work[1..] = tile descriptors.  VEC: all predicate columns are 16-byte aligned.
template <bool VEC>
__global__ void __launch_bounds__(kFilterThreads) filter_scan_kernel(TermPack terms, uint64_t n, uint64_t n_tiles,
                                                                     uint8_t* __restrict__ out_mask, uint32_t* __restrict__ out_sel,
                                                                     CompactCols cc, uint64_t* __restrict__ out_count,
                                                                     uint64_t* __restrict__ work) {
  __shared__ uint64_t s_tile;
  __shared__ uint32_t s_warp[kFilterPairs][kFilterThreads / 32];
  __shared__ uint64_t s_excl;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint64_t* status = work + 1;

  for (;;) {
    if (tid == 0) s_tile = atomicAdd((unsigned long long*)work, 1ULL);  // ticket order = start order: look-back cannot deadlock
    __syncthreads();
    const uint64_t tile = s_tile;
    if (tile >= n_tiles) return;
    const uint64_t tile_base = tile * (uint64_t)kFilterTile;

    bool f[kFilterPairs][2];
#pragma unroll
    for (int j = 0; j < kFilterPairs; j++) { f[j][0] = true; f[j][1] = true; }
#pragma unroll 1
    for (int k = 0; k < terms.n; k++) {
      uint64_t l[kFilterPairs][2], r[kFilterPairs][2];
#pragma unroll
      for (int j = 0; j < kFilterPairs; j++) {
        const uint64_t row = tile_base + (uint64_t)(j * kFilterThreads + tid) * 2;
        l[j][0] = l[j][1] = 0; r[j][0] = r[j][1] = (uint64_t)terms.lit[k];
        if (VEC && row + 1 < n) {
          ulonglong2 v = ld_stream_u64x2(terms.lhs[k] + row);
          l[j][0] = v.x; l[j][1] = v.y;
          if (terms.rhs[k]) { ulonglong2 w = ld_stream_u64x2(terms.rhs[k] + row); r[j][0] = w.x; r[j][1] = w.y; }
        } else {
          if (row < n) { l[j][0] = ld_stream_u64(terms.lhs[k] + row); if (terms.rhs[k]) r[j][0] = ld_stream_u64(terms.rhs[k] + row); }
          if (row + 1 < n) { l[j][1] = ld_stream_u64(terms.lhs[k] + row + 1); if (terms.rhs[k]) r[j][1] = ld_stream_u64(terms.rhs[k] + row + 1); }
        }
      }
#pragma unroll
      for (int j = 0; j < kFilterPairs; j++) {
        f[j][0] = f[j][0] && eval_term(terms, k, l[j][0], r[j][0]);
        f[j][1] = f[j][1] && eval_term(terms, k, l[j][1], r[j][1]);
      }
    }
    // rows past the end never pass; write the mask
    uint32_t rank[kFilterPairs];
#pragma unroll
    for (int j = 0; j < kFilterPairs; j++) {
      const uint64_t row = tile_base + (uint64_t)(j * kFilterThreads + tid) * 2;
      f[j][0] = f[j][0] && row < n;
      f[j][1] = f[j][1] && row + 1 < n;
      if (out_mask) {
        if (row + 1 < n) *reinterpret_cast<uchar2*>(out_mask + row) = make_uchar2(f[j][0], f[j][1]);
        else if (row < n) out_mask[row] = f[j][0];
      }
      const unsigned b0 = __ballot_sync(0xffffffffu, f[j][0]), b1 = __ballot_sync(0xffffffffu, f[j][1]);
      const unsigned lt = (1u << lane) - 1;
      rank[j] = __popc(b0 & lt) + __popc(b1 & lt);  // passing rows of earlier lanes (both elements precede)
      if (lane == 0) s_warp[j][warp] = __popc(b0) + __popc(b1);
    }
    __syncthreads();
    if (warp == 0) {  // scan the 32 (pair, warp) totals in row order, then look back across tiles
      constexpr int W = kFilterThreads / 32;
      uint32_t c = lane < kFilterPairs * W ? s_warp[lane / W][lane % W] : 0;
      uint32_t inc = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
      const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
      if (lane < kFilterPairs * W) s_warp[lane / W][lane % W] = inc - c;
      const uint64_t excl = lookback_exclusive(status, tile, total, lane);
      if (lane == 0) {
        s_excl = excl;
        if (tile == n_tiles - 1 && out_count) *out_count = excl + total;
      }
    }
    __syncthreads();
    if (out_sel || cc.n) {
      const uint64_t excl = s_excl;
#pragma unroll
      for (int j = 0; j < kFilterPairs; j++) {
        const uint64_t row = tile_base + (uint64_t)(j * kFilterThreads + tid) * 2;
        uint64_t dst = excl + s_warp[j][warp] + rank[j];
#pragma unroll
        for (int e = 0; e < 2; e++) {
          if (f[j][e]) {
            if (out_sel) out_sel[dst] = (uint32_t)(row + e);
#pragma unroll 1
            for (int c = 0; c < cc.n; c++) cc.out[c][dst] = ld_stream_u64(cc.in[c] + row + e);
            dst++;
          }
        }
      }
    }
    __syncthreads();  // s_tile / s_warp reused by the next tile
  }
}

__global__ void __launch_bounds__(256) gather_kernel(const uint64_t* __restrict__ col, const uint32_t* __restrict__ sel, uint64_t n,
                                                     uint64_t* __restrict__ out) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    uint32_t s0 = ld_stream_u32(sel + i), s1 = ld_stream_u32(sel + i + stride), s2 = ld_stream_u32(sel + i + 2 * stride),
             s3 = ld_stream_u32(sel + i + 3 * stride);
    uint64_t v0 = __ldg(col + s0), v1 = __ldg(col + s1), v2 = __ldg(col + s2), v3 = __ldg(col + s3);
    out[i] = v0; out[i + stride] = v1; out[i + 2 * stride] = v2; out[i + 3 * stride] = v3;
  }
  for (; i < n; i += stride) out[i] = __ldg(col + ld_stream_u32(sel + i));
}

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

template <bool F64>
__global__ void __launch_bounds__(256) project_kernel(const uint64_t* __restrict__ lhs, const uint64_t* __restrict__ rhs, uint64_t lit,
                                                      char op, uint64_t n, uint64_t* __restrict__ out, bool vec) {
  const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
  auto apply = [&](uint64_t a, uint64_t b) -> uint64_t {
    if (F64) return (uint64_t)__double_as_longlong(binop_f64(__longlong_as_double((long long)a), __longlong_as_double((long long)b), op));
    return (uint64_t)binop_i64((int64_t)a, (int64_t)b, op);
  };
  if (vec) {
    const uint64_t n2 = n >> 1;
    for (uint64_t i = t; i < n2; i += stride) {
      ulonglong2 a = ld_stream_u64x2(lhs + 2 * i), b = make_ulonglong2(lit, lit);
      if (rhs) b = ld_stream_u64x2(rhs + 2 * i);
      st_stream_u64x2(out + 2 * i, make_ulonglong2(apply(a.x, b.x), apply(a.y, b.y)));
    }
    if (t == 0 && (n & 1)) out[n - 1] = apply(lhs[n - 1], rhs ? rhs[n - 1] : lit);
  } else {
    for (uint64_t i = t; i < n; i += stride) out[i] = apply(lhs[i], rhs ? rhs[i] : lit);
  }
}

static inline bool aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

static olap_status run_filter(olap_ctx* ctx, const olap_filter_term* terms, int32_t n_terms, uint64_t n, uint8_t* out_mask,
                              uint32_t* out_sel, const CompactCols& cc, uint64_t* out_count, uint64_t* host_count, cudaStream_t s) {
  TermPack tp;
  OLAP_TRY(pack_terms(terms, n_terms, &tp));
  OLAP_REQUIRE(n <= 0xFFFFFFFEULL || !out_sel, "selection vectors hold u32 row ids: n_rows must be < 2^32");
  Scratch sc(ctx, s);
  uint64_t* dcount = out_count;
  if (!dcount && host_count) OLAP_TRY(sc.get<uint64_t>(1, &dcount));
  if (n == 0) {
    if (dcount) OLAP_CUDA(cudaMemsetAsync(dcount, 0, 8, s));
  } else {
    const uint64_t n_tiles = (n + kFilterTile - 1) / kFilterTile;
    uint64_t* work = nullptr;
    OLAP_TRY(sc.get<uint64_t>(n_tiles + 1, &work));
    OLAP_CUDA(cudaMemsetAsync(work, 0, (n_tiles + 1) * 8, s));
    bool vec = true;
    for (int k = 0; k < tp.n; k++) vec = vec && aligned16(tp.lhs[k]) && (!tp.rhs[k] || aligned16(tp.rhs[k]));
    unsigned grid = (unsigned)(n_tiles < (uint64_t)ctx->sm_count * 8 ? n_tiles : (uint64_t)ctx->sm_count * 8);
    if (vec) filter_scan_kernel<true><<<grid, kFilterThreads, 0, s>>>(tp, n, n_tiles, out_mask, out_sel, cc, dcount, work);
    else filter_scan_kernel<false><<<grid, kFilterThreads, 0, s>>>(tp, n, n_tiles, out_mask, out_sel, cc, dcount, work);
    count_launch(ctx);
    OLAP_CUDA(cudaGetLastError());
  }
  if (host_count) {
    OLAP_CUDA(cudaMemcpyAsync(host_count, dcount, 8, cudaMemcpyDeviceToHost, s));
    OLAP_CUDA(cudaStreamSynchronize(s));
  }
  return OLAP_OK;
}

}  // namespace olap

using namespace olap;

extern "C" {

olap_status olap_filter(olap_ctx* ctx, const olap_filter_term* terms, int32_t n_terms, uint64_t n_rows, uint8_t* out_mask,
                        uint32_t* out_sel, uint64_t* out_count_dptr, uint64_t* host_count, void* stream) {
  OLAP_REQUIRE(ctx, "ctx is NULL");
  std::lock_guard<std::mutex> g(ctx->mu);
  OLAP_CUDA(cudaSetDevice(ctx->device));
  CompactCols cc;
  memset(&cc, 0, sizeof(cc));
  return run_filter(ctx, terms, n_terms, n_rows, out_mask, out_sel, cc, out_count_dptr, host_count, ctx->stream_of(stream));
}

olap_status olap_filter_i64(olap_ctx* ctx, const int64_t* col, uint64_t n_rows, int32_t op, int64_t literal, uint8_t* out_mask,
                            uint32_t* out_sel, uint64_t* out_count_dptr, uint64_t* host_count, void* stream) {
  olap_filter_term t;
  t.lhs = col; t.rhs = nullptr; t.literal_bits = literal; t.op = op; t.type = OLAP_I64;
  return olap_filter(ctx, &t, 1, n_rows, out_mask, out_sel, out_count_dptr, host_count, stream);
}

olap_status olap_filter_compact(olap_ctx* ctx, const olap_filter_term* terms, int32_t n_terms, uint64_t n_rows,
                                const void* const* cols, int32_t n_cols, void* const* out_cols, uint64_t* out_count_dptr,
                                uint64_t* host_count, void* stream) {
  OLAP_REQUIRE(ctx, "ctx is NULL");
  OLAP_REQUIRE(n_cols >= 0 && n_cols <= kMaxCompactCols, "n_cols must be in [0, 8]");
  std::lock_guard<std::mutex> g(ctx->mu);
  OLAP_CUDA(cudaSetDevice(ctx->device));
  CompactCols cc;
  memset(&cc, 0, sizeof(cc));
  cc.n = n_cols;
  for (int c = 0; c < n_cols; c++) {
    OLAP_REQUIRE(n_rows == 0 || (cols[c] && out_cols[c]), "column pointer is NULL");
    cc.in[c] = (const uint64_t*)cols[c];
    cc.out[c] = (uint64_t*)out_cols[c];
  }
  return run_filter(ctx, terms, n_terms, n_rows, nullptr, nullptr, cc, out_count_dptr, host_count, ctx->stream_of(stream));
}

olap_status olap_gather_8b(olap_ctx* ctx, const void* col, const uint32_t* sel, uint64_t n_sel, void* out, void* stream) {
  OLAP_REQUIRE(ctx && (n_sel == 0 || (col && sel && out)), "NULL pointer");
  std::lock_guard<std::mutex> g(ctx->mu);
  OLAP_CUDA(cudaSetDevice(ctx->device));
  if (n_sel == 0) return OLAP_OK;
  uint64_t blocks = (n_sel + 1023) / 1024;
  unsigned grid = (unsigned)(blocks < (uint64_t)ctx->sm_count * 8 ? (blocks ? blocks : 1) : (uint64_t)ctx->sm_count * 8);
  gather_kernel<<<grid, 256, 0, ctx->stream_of(stream)>>>((const uint64_t*)col, sel, n_sel, (uint64_t*)out);
  count_launch(ctx);
  OLAP_CUDA(cudaGetLastError());
  return OLAP_OK;
}

static olap_status project(olap_ctx* ctx, bool f64, const void* lhs, const void* rhs, uint64_t lit, char op, uint64_t n, void* out,
                           void* stream) {
  OLAP_REQUIRE(ctx && (n == 0 || (lhs && out)), "NULL pointer");
  std::lock_guard<std::mutex> g(ctx->mu);
  OLAP_CUDA(cudaSetDevice(ctx->device));
  if (n == 0) return OLAP_OK;
  const bool vec = aligned16(lhs) && aligned16(out) && (!rhs || aligned16(rhs));
  uint64_t blocks = (n / 2 + 255) / 256;
  unsigned grid = (unsigned)(blocks < (uint64_t)ctx->sm_count * 8 ? (blocks ? blocks : 1) : (uint64_t)ctx->sm_count * 8);
  cudaStream_t s = ctx->stream_of(stream);
  if (f64) project_kernel<true><<<grid, 256, 0, s>>>((const uint64_t*)lhs, (const uint64_t*)rhs, lit, op, n, (uint64_t*)out, vec);
  else project_kernel<false><<<grid, 256, 0, s>>>((const uint64_t*)lhs, (const uint64_t*)rhs, lit, op, n, (uint64_t*)out, vec);
  count_launch(ctx);
  OLAP_CUDA(cudaGetLastError());
  return OLAP_OK;
}

olap_status olap_project_i64(olap_ctx* ctx, const int64_t* lhs, const int64_t* rhs, int64_t literal, char op, uint64_t n_rows,
                             int64_t* out, void* stream) {
  return project(ctx, false, lhs, rhs, (uint64_t)literal, op, n_rows, out, stream);
}

olap_status olap_project_f64(olap_ctx* ctx, const double* lhs, const double* rhs, double literal, char op, uint64_t n_rows, double* out,
                             void* stream) {
  uint64_t bits;
  memcpy(&bits, &literal, 8);
  return project(ctx, true, lhs, rhs, bits, op, n_rows, out, stream);
}

}  // extern "C"
