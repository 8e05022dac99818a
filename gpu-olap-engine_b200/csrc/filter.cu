// Fused predicate + mask + order-preserving stream compaction (one pass, decoupled look-back), gather and
// projection kernels.  Replaces eval_filter_mask / filter_record_batch / take_int64 / eval_binary_expr
// (gpu-olap-core/src/executor.rs:411-507) and the JIT'd filter_kernel (gpu-kernel-compiler/src/codegen.rs:36-47).
//
// Roofline: HBM.  Algorithmic bytes per row = 8 B per distinct predicate column read + 1 B mask (if asked) +
// 4 B per passing row for the selection vector (+ 16 B per passing row per compacted column).
#include "common.cuh"

namespace olap {

constexpr int kFilterThreads = 512;
constexpr int kFilterQuads = 2;                                   // 256-bit loads (4 consecutive rows) per thread per tile
constexpr int kFilterTile = kFilterThreads * kFilterQuads * 4;    // 4096 rows per tile
constexpr int kScanQuads = 8;                                     // evaluated in batches of kScanBatch quads (registers), results kept as bits
constexpr int kScanBatch = 4;
constexpr int kScanTile = kFilterThreads * kScanQuads * 4;        // look-back kernel: 16384-row tiles, selected row offsets (u16) staged in smem
constexpr int kMaxCompactCols = 8;

struct CompactCols {
  const uint64_t* in[kMaxCompactCols];
  uint64_t* out[kMaxCompactCols];
  int n;
};

__device__ __forceinline__ void store_mask4(uint8_t* out_mask, uint64_t row0, uint64_t n, const bool (&f)[4]) {
  if (row0 + 3 < n && ((uintptr_t)(out_mask + row0) & 3) == 0)
    *reinterpret_cast<uint32_t*>(out_mask + row0) = (uint32_t)f[0] | ((uint32_t)f[1] << 8) | ((uint32_t)f[2] << 16) | ((uint32_t)f[3] << 24);
  else {
#pragma unroll
    for (int e = 0; e < 4; e++) if (row0 + e < n) out_mask[row0 + e] = f[e];
  }
}

// Mask and/or count only (the FilterParams contract, filter_kernel.rs:17-28): a pure streaming kernel -- no ranking,
// no inter-tile dependency; one global atomic per CTA for the count.
template <bool VEC>
__global__ void __launch_bounds__(kFilterThreads) filter_mask_kernel(TermPack terms, uint64_t n, uint8_t* __restrict__ out_mask,
                                                                     unsigned long long* __restrict__ out_count) {
  unsigned cnt = 0;
  const uint64_t stride = (uint64_t)gridDim.x * kFilterTile;
  for (uint64_t base = blockIdx.x * (uint64_t)kFilterTile; base < n; base += stride) {
    bool f[kFilterQuads][4];
#pragma unroll
    for (int q = 0; q < kFilterQuads; q++) eval_quad<VEC>(terms, base + (uint64_t)(q * kFilterThreads + threadIdx.x) * 4, n, f[q]);
#pragma unroll
    for (int q = 0; q < kFilterQuads; q++) {
      const uint64_t row0 = base + (uint64_t)(q * kFilterThreads + threadIdx.x) * 4;
      if (out_mask && row0 < n) store_mask4(out_mask, row0, n, f[q]);
      cnt += (unsigned)f[q][0] + f[q][1] + f[q][2] + f[q][3];
    }
  }
  if (out_count) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    __shared__ unsigned s_c[kFilterThreads / 32];
    if ((threadIdx.x & 31) == 0) s_c[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long t = 0;
      for (int w = 0; w < kFilterThreads / 32; w++) t += s_c[w];
      if (t) atomicAdd(out_count, t);
    }
  }
}

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

// Selection vector and/or fused compaction of up to 8 columns, row order preserved: single pass with a decoupled
// look-back over 16384-row tiles.  work[0] = tile ticket, work[1..] = tile descriptors.
// Every tile costs a chain of dependent round trips (ticket atomic, column loads, look-back loads, write-out); with
// 4096-row tiles that chain, not HBM, set the speed (5.4 ms at 2^30 rows against 1.7 ms for the mask-only kernel), so the
// tile is 16384 rows: predicates are evaluated in two batches of four 256-bit loads per thread and kept as one bit per row.
template <bool VEC>
__global__ void __launch_bounds__(kFilterThreads, 3) filter_scan_kernel(TermPack terms, uint64_t n, uint64_t n_tiles,
                                                                     uint8_t* __restrict__ out_mask, uint32_t* __restrict__ out_sel,
                                                                     CompactCols cc, uint64_t* __restrict__ out_count,
                                                                     uint64_t* __restrict__ work) {
  constexpr int W = kFilterThreads / 32;
  static_assert(kScanTile <= 65536, "staged row offsets are u16");
  static_assert(kScanQuads * 4 <= 32, "one flag bit per row of a thread");
  __shared__ uint64_t s_tile;
  __shared__ uint32_t s_warp[kScanQuads * W];
  __shared__ uint64_t s_excl;
  __shared__ uint32_t s_total;
  __shared__ uint16_t s_sel[kScanTile];  // the tile's passing row offsets, compacted: written out with fully coalesced stores
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint64_t* status = work + 1;

  for (;;) {
    if (tid == 0) s_tile = atomicAdd((unsigned long long*)work, 1ULL);  // ticket order = start order: look-back cannot deadlock
    __syncthreads();
    const uint64_t tile = s_tile;
    if (tile >= n_tiles) return;
    const uint64_t tile_base = tile * (uint64_t)kScanTile;

    uint32_t flags = 0;          // bit q * 4 + e: row e of quad q passes
    uint32_t rank[kScanQuads];
#pragma unroll
    for (int q0 = 0; q0 < kScanQuads; q0 += kScanBatch) {
      bool f[kScanBatch][4];
#pragma unroll
      for (int q = 0; q < kScanBatch; q++) eval_quad<VEC>(terms, tile_base + (uint64_t)((q0 + q) * kFilterThreads + tid) * 4, n, f[q]);
#pragma unroll
      for (int q = 0; q < kScanBatch; q++) {
        const uint64_t row0 = tile_base + (uint64_t)((q0 + q) * kFilterThreads + tid) * 4;
        if (out_mask && row0 < n) store_mask4(out_mask, row0, n, f[q]);
        // rows of a quad are consecutive, lanes follow each other: an inclusive warp scan of the per-thread counts ranks them
        const uint32_t c = (uint32_t)f[q][0] + f[q][1] + f[q][2] + f[q][3];
        flags |= ((uint32_t)f[q][0] | ((uint32_t)f[q][1] << 1) | ((uint32_t)f[q][2] << 2) | ((uint32_t)f[q][3] << 3)) << ((q0 + q) * 4);
        uint32_t inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        rank[q0 + q] = inc - c;
        if (lane == 31) s_warp[(q0 + q) * W + warp] = inc;
      }
    }
    __syncthreads();
    if (warp == 0) {  // scan the (quad, warp) totals in row order (PER per lane), then look back across tiles
      constexpr int PER = kScanQuads * W / 32;
      uint32_t v[PER], sum = 0;
#pragma unroll
      for (int i = 0; i < PER; i++) { v[i] = s_warp[lane * PER + i]; sum += v[i]; }
      uint32_t inc = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
      const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
      uint32_t run = inc - sum;
#pragma unroll
      for (int i = 0; i < PER; i++) { s_warp[lane * PER + i] = run; run += v[i]; }
      if (lane == 0) s_total = total;
    }
    __syncthreads();
    if (warp == 0) {  // warp 0 walks back over the earlier tiles while the other warps stage this tile's row offsets
      const uint32_t total = s_total;
      const uint64_t excl = lookback_exclusive(status, tile, total, lane);
      if (lane == 0) {
        s_excl = excl;
        if (tile == n_tiles - 1 && out_count) *out_count = excl + total;
      }
    }
#pragma unroll
    for (int q = 0; q < kScanQuads; q++) {
      const uint32_t row0 = (uint32_t)(q * kFilterThreads + tid) * 4;
      uint32_t pos = s_warp[q * W + warp] + rank[q];
#pragma unroll
      for (int e = 0; e < 4; e++)
        if ((flags >> (q * 4 + e)) & 1u) s_sel[pos++] = (uint16_t)(row0 + e);
    }
    __syncthreads();
    {
      // a warp scattering ~16 four-byte ids over 8 half-filled sectors made this kernel L2-write-transaction bound
      // (5.7 ms at 2^30 rows); from shared memory every store instruction fills whole sectors
      const uint64_t excl = s_excl;
      const uint32_t total = s_total;
      for (uint32_t i = tid; i < total; i += kFilterThreads) {
        const uint64_t row = tile_base + s_sel[i];
        if (out_sel) out_sel[excl + i] = (uint32_t)row;
#pragma unroll 1
        for (int c = 0; c < cc.n; c++) cc.out[c][excl + i] = ld_stream_u64(cc.in[c] + row);
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
static inline bool aligned32(const void* p) { return (((uintptr_t)p) & 31) == 0; }

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
    bool vec = true;
    for (int k = 0; k < tp.n; k++) vec = vec && aligned32(tp.lhs[k]) && (!tp.rhs[k] || aligned32(tp.rhs[k]));
    if (!out_sel && cc.n == 0) {  // mask / count only: streaming kernel, no inter-tile dependency
      if (dcount) OLAP_CUDA(cudaMemsetAsync(dcount, 0, 8, s));
      unsigned grid = (unsigned)(n_tiles < (uint64_t)ctx->sm_count * 4 ? n_tiles : (uint64_t)ctx->sm_count * 4);
      if (vec) filter_mask_kernel<true><<<grid, kFilterThreads, 0, s>>>(tp, n, out_mask, (unsigned long long*)dcount);
      else filter_mask_kernel<false><<<grid, kFilterThreads, 0, s>>>(tp, n, out_mask, (unsigned long long*)dcount);
    } else {
      const uint64_t s_tiles = (n + kScanTile - 1) / kScanTile;
      uint64_t* work = nullptr;
      OLAP_TRY(sc.get<uint64_t>(s_tiles + 1, &work));
      OLAP_CUDA(cudaMemsetAsync(work, 0, (s_tiles + 1) * 8, s));
      unsigned grid = (unsigned)(s_tiles < (uint64_t)ctx->sm_count * 3 ? s_tiles : (uint64_t)ctx->sm_count * 3);
      if (vec) filter_scan_kernel<true><<<grid, kFilterThreads, 0, s>>>(tp, n, s_tiles, out_mask, out_sel, cc, dcount, work);
      else filter_scan_kernel<false><<<grid, kFilterThreads, 0, s>>>(tp, n, s_tiles, out_mask, out_sel, cc, dcount, work);
    }
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
