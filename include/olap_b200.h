/*
 * olap_b200.h -- C ABI of the B200-native execution core for gpu-olap-engine's data-parallel hot path.
 *
 * This is the seam the reference's Rust executor binds (see INTEGRATION.md for the -sys crate stub).  Every
 * entry point names the reference interface it replaces (path:line under the reference checkout).  Plain C:
 * pointers, sizes, POD structs; no C++/torch types; no exceptions cross it.  Every function returns an
 * olap_status; olap_last_error() gives the thread-local message (maps 1:1 onto anyhow::Error in Rust).
 *
 * Conventions
 *   - "dptr" arguments are device addresses (what the reference passes as u64 in JoinParams/FilterParams/
 *     AggregateParams); all columns are packed little-endian 8-byte values (arrow-interop column_buffer.rs:19-21).
 *   - stream is a cudaStream_t passed as void* (NULL = the context's stream 0).  Filters, projections, gathers, uploads,
 *     copies, partition / shuffle passes and the olap_shard_* phase calls are asynchronous on that stream unless a host
 *     out-parameter (uint64_t* host_*) is non-NULL, in which case the call synchronises the stream before returning so
 *     the value is valid.  The join (olap_hash_join_i64, olap_join_build_create, olap_join_probe) and the group-by are
 *     ADAPTIVE operators: they read a few words back (fast-path flags, partition fill, the group count) to pick or to
 *     abandon a regime, so they synchronise their stream internally whether or not a host out-parameter is given; on
 *     return their results are complete in stream order either way.
 *   - One olap_ctx per GPU.  Calls on one ctx from several host threads are serialised by an internal mutex (internal
 *     waits included); distinct ctxs are independent, so concurrent queries take one ctx each.
 *   - There is no CPU fallback anywhere behind this header.
 */
#ifndef OLAP_B200_H
#define OLAP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OLAP_API __attribute__((visibility("default")))

typedef struct olap_ctx olap_ctx;
typedef int32_t olap_status;

enum {
  OLAP_OK = 0,
  OLAP_ERR_BAD_ARG = 1,
  OLAP_ERR_OOM = 2,
  OLAP_ERR_OVERFLOW = 3, /* join produced more pairs than max_matches (count is still exact) */
  OLAP_ERR_CUDA = 4,
  OLAP_ERR_UNSUPPORTED = 5,
  OLAP_ERR_CAPACITY = 6  /* more groups than out_capacity */
};

/* CmpOp, gpu-kernel-compiler/src/codegen.rs:10 (declaration order) */
enum { OLAP_EQ = 0, OLAP_NE = 1, OLAP_LT = 2, OLAP_LE = 3, OLAP_GT = 4, OLAP_GE = 5 };
/* AggFunc, gpu-kernel-compiler/src/aggregate_kernel.rs:17-25 */
enum { OLAP_SUM = 0, OLAP_COUNT = 1, OLAP_MIN = 2, OLAP_MAX = 3, OLAP_AVG = 4 };
/* GpuDataType, arrow-interop/src/column_buffer.rs:8-15 (the two arithmetic layouts) */
enum { OLAP_I64 = 0, OLAP_F64 = 1 };
/* JoinStrategy, gpu-kernel-compiler/src/join_kernel.rs:13-18; AUTO applies choose_strategy (:71-77) */
enum { OLAP_JOIN_AUTO = -1, OLAP_JOIN_RADIX = 0, OLAP_JOIN_SORT_MERGE = 1, OLAP_JOIN_BROADCAST = 2 };
/* Arrow source types accepted by olap_column_upload (record_batch_convert.rs:42-98) */
enum { OLAP_SRC_I8 = 0, OLAP_SRC_I16, OLAP_SRC_I32, OLAP_SRC_I64, OLAP_SRC_U8, OLAP_SRC_U16, OLAP_SRC_U32, OLAP_SRC_U64,
       OLAP_SRC_F32, OLAP_SRC_F64 };

/* ---------------------------------------------------------------------------------------------------------
 * Lifecycle.  Replaces EngineConfig (gpu-olap-core/src/lib.rs:20-43) + TransferQueue::new
 * (gpu-memory-manager/src/transfer_queue.rs:36-52).  max_gpu_memory caps the stream-ordered pool (0 = no cap);
 * use_unified_memory has no equivalent: device pools replace managed memory by design.
 * ------------------------------------------------------------------------------------------------------- */
OLAP_API olap_status olap_ctx_create(int32_t device_ordinal, uint64_t max_gpu_memory, int32_t num_streams, olap_ctx** out);
OLAP_API olap_status olap_ctx_destroy(olap_ctx* ctx);
OLAP_API olap_status olap_ctx_synchronize(olap_ctx* ctx);               /* TransferQueue::synchronize, :133-139 */
OLAP_API void* olap_ctx_stream(olap_ctx* ctx, int32_t stream_id);       /* stream_id % num_streams, :72 */
OLAP_API const char* olap_last_error(void);
OLAP_API const char* olap_version(void);

/* ---------------------------------------------------------------------------------------------------------
 * Memory.  Replaces SlabAllocator::allocate/free (gpu-memory-manager/src/slab_allocator.rs:50-93) with a
 * stream-ordered cudaMemPool, and TransferQueue::enqueue_transfer (transfer_queue.rs:55-79).
 * ------------------------------------------------------------------------------------------------------- */
OLAP_API olap_status olap_pool_alloc(olap_ctx* ctx, uint64_t bytes, void* stream, void** out_dptr);
OLAP_API olap_status olap_pool_free(olap_ctx* ctx, void* dptr, void* stream);
OLAP_API olap_status olap_pool_stats(olap_ctx* ctx, uint64_t* used_bytes, uint64_t* high_water_bytes, uint64_t* reserved_bytes);
OLAP_API olap_status olap_host_alloc(uint64_t bytes, void** out_hptr);  /* pinned */
OLAP_API olap_status olap_host_free(void* hptr);
OLAP_API olap_status olap_memcpy_h2d(olap_ctx* ctx, void* dptr, const void* hptr, uint64_t bytes, void* stream);
OLAP_API olap_status olap_memcpy_d2h(olap_ctx* ctx, void* hptr, const void* dptr, uint64_t bytes, void* stream);
OLAP_API olap_status olap_stream_synchronize(olap_ctx* ctx, void* stream);

/* record_batch_to_gpu_buffers / column_to_buffer (arrow-interop/src/record_batch_convert.rs:22-110): copy an
 * Arrow values buffer to the device and widen it to 8 bytes per row THERE.  Int64/Float64 without nulls is a
 * straight async copy from the Arrow buffer.  validity_bitmap is Arrow's bit-per-row bitmap (LSB first) or
 * NULL; null rows become 0 / 0.0 (:103, :108); out_validity_bytes (nullable dptr, 1 byte per row, :37-41).
 * bit_offset is the Arrow array offset applied to both buffers.                                            */
OLAP_API olap_status olap_column_upload(olap_ctx* ctx, const void* host_values, uint64_t n_rows, int32_t src_type,
                                        const uint8_t* validity_bitmap, uint64_t bit_offset, void* out_dptr,
                                        uint8_t* out_validity_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Filter.  Replaces FilterParams/FilterKernel::launch (gpu-kernel-compiler/src/filter_kernel.rs:17-72) and the
 * per-predicate PTX JIT (codegen.rs:36-47): (op, literal) are kernel arguments of one AOT kernel.  Semantics:
 * eval_filter_mask (gpu-olap-core/src/executor.rs:443-476).  A conjunction is evaluated in one pass.
 * ------------------------------------------------------------------------------------------------------- */
typedef struct {
  const void* lhs;      /* dptr, 8-byte column */
  const void* rhs;      /* dptr of a second column, or NULL to compare with the literal */
  int64_t literal_bits; /* i64 literal, or the bit pattern of an f64 literal when type == OLAP_F64 */
  int32_t op;           /* OLAP_EQ..OLAP_GE; any other value matches no row (executor.rs:464) */
  int32_t type;         /* OLAP_I64 (reference behaviour) or OLAP_F64 (extension) */
} olap_filter_term;

/* out_mask: u8 per row, 1 = pass (filter_kernel.rs:3-5), nullable.  out_sel: u32 row ids of passing rows in
 * row order (the "prefix-sum + scatter" pass, fused), nullable.  out_count_dptr: u64 on device, nullable.
 * host_count: nullable (synchronises).                                                                      */
OLAP_API olap_status olap_filter(olap_ctx* ctx, const olap_filter_term* terms, int32_t n_terms, uint64_t n_rows,
                                 uint8_t* out_mask, uint32_t* out_sel, uint64_t* out_count_dptr, uint64_t* host_count,
                                 void* stream);
/* FilterPredicate{col, op, literal:i64} convenience form (codegen.rs:23-28) */
OLAP_API olap_status olap_filter_i64(olap_ctx* ctx, const int64_t* col, uint64_t n_rows, int32_t op, int64_t literal,
                                     uint8_t* out_mask, uint32_t* out_sel, uint64_t* out_count_dptr, uint64_t* host_count,
                                     void* stream);
/* take_int64 (executor.rs:496-507): out[i] = col[sel[i]] for any 8-byte column */
OLAP_API olap_status olap_gather_8b(olap_ctx* ctx, const void* col, const uint32_t* sel, uint64_t n_sel, void* out, void* stream);
/* filter_record_batch (executor.rs:478-494) fused: evaluate the conjunction and compact n_cols 8-byte columns
 * in one pass, row order preserved.  out_cols[i] must hold n_rows values.                                   */
OLAP_API olap_status olap_filter_compact(olap_ctx* ctx, const olap_filter_term* terms, int32_t n_terms, uint64_t n_rows,
                                         const void* const* cols, int32_t n_cols, void* const* out_cols,
                                         uint64_t* out_count_dptr, uint64_t* host_count, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Projection.  eval_binary_expr (executor.rs:411-441): op in "+-*" wraps, '/' by zero gives 0, any other op 0.
 * rhs == NULL uses the literal.                                                                             */
OLAP_API olap_status olap_project_i64(olap_ctx* ctx, const int64_t* lhs, const int64_t* rhs, int64_t literal, char op,
                                      uint64_t n_rows, int64_t* out, void* stream);
/* f64 arithmetic extension (the reference errors at executor.rs:416-424) */
OLAP_API olap_status olap_project_f64(olap_ctx* ctx, const double* lhs, const double* rhs, double literal, char op,
                                      uint64_t n_rows, double* out, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Hash join.  Replaces JoinParams/JoinKernel::join (gpu-kernel-compiler/src/join_kernel.rs:20-68) and the
 * kernels radix_partition_kernel / build_hash_table_kernel / probe_hash_table_kernel
 * (kernels/join_kernels.cuh:45-166).  Inner equi-join on int64 keys; build side = left (executor.rs:238-253).
 * Output: (left_row, right_row) pairs in unspecified order (compare sorted), exact match count.
 * ------------------------------------------------------------------------------------------------------- */
typedef struct { uint32_t left_row, right_row; } olap_pair;

/* out_pairs may be NULL (count only).  If count > max_matches the first max_matches reservations are written,
 * the count stays exact (join_kernels.cuh:155-160) and -- when host_count is given -- the call returns
 * OLAP_ERR_OVERFLOW instead of the reference's silent truncation.                                           */
OLAP_API olap_status olap_hash_join_i64(olap_ctx* ctx, const int64_t* build_keys, uint64_t n_build,
                                        const int64_t* probe_keys, uint64_t n_probe, olap_pair* out_pairs,
                                        uint64_t max_matches, uint64_t* out_count_dptr, uint64_t* host_count,
                                        int32_t strategy, void* stream);
/* Build once, probe per batch: the reference executor receives the right side as a Vec<RecordBatch> (executor.rs:66-69,
 * EngineConfig.batch_size), so the probe side can stream through while host<->device copies of the neighbouring batches
 * overlap.  probe_row_base is added to every right_row of this batch; pairs/count are per call.                   */
typedef struct olap_join_build olap_join_build;
OLAP_API olap_status olap_join_build_create(olap_ctx* ctx, const int64_t* build_keys, uint64_t n_build, int32_t strategy,
                                            void* stream, olap_join_build** out);
OLAP_API olap_status olap_join_probe(olap_join_build* build, const int64_t* probe_keys, uint64_t n_probe, uint32_t probe_row_base,
                                     olap_pair* out_pairs, uint64_t max_matches, uint64_t* out_count_dptr, uint64_t* host_count,
                                     void* stream);
OLAP_API olap_status olap_join_build_destroy(olap_join_build* build);
/* Join output materialisation (the RecordBatch execute_hash_join promises and never builds, executor.rs:255-264): out[i] =
 * col[pairs[i].left_row] (side 0) or col[pairs[i].right_row] (side 1) for any 8-byte column -- the payload columns of both sides in
 * pair order.  Random 8-byte gathers: a 32-byte DRAM sector per value once the column is larger than the L2.                  */
OLAP_API olap_status olap_join_gather_8b(olap_ctx* ctx, const void* col, const olap_pair* pairs, int32_t side, uint64_t n_pairs, void* out,
                                         void* stream);
/* Outer joins (JoinType::{Left, Right, Full}, gpu-olap-core/src/logical_plan.rs:52-58; the reference never executes them,
 * README.md:410-411): the row ids of one side that appear in NO pair, in ascending row order.  side 0 = left / build rows,
 * 1 = right / probe rows; n_rows = rows of that side; out_row_ids holds up to n_rows u32.  The operator layer appends these rows,
 * NULL-extended, behind the inner pairs.  Asynchronous unless host_count is given.                                          */
OLAP_API olap_status olap_join_unmatched_rows(olap_ctx* ctx, const olap_pair* pairs, uint64_t n_pairs, int32_t side, uint64_t n_rows,
                                              uint32_t* out_row_ids, uint64_t* out_count_dptr, uint64_t* host_count, void* stream);
/* JoinKernel::choose_strategy (join_kernel.rs:71-77) */
OLAP_API int32_t olap_choose_join_strategy(uint64_t n_build);

/* radix_partition_kernel (join_kernels.cuh:45-76) as a unit: partitions (key, row id) into 2^bits partitions by
 * the high bits of the core's own 64-bit mixer (partition ids are not externally visible).
 * row_ids NULL = implicit 0..n-1.  out_part_begin: dptr to 2^bits + 1 u32 offsets.                          */
OLAP_API olap_status olap_radix_partition_i64(olap_ctx* ctx, const int64_t* keys, const uint32_t* row_ids, uint64_t n_rows,
                                              int32_t bits, int64_t* out_keys, uint32_t* out_row_ids,
                                              uint32_t* out_part_begin, void* stream);
/* the partition a key lands in (host helper for tests) */
OLAP_API uint32_t olap_partition_of(int64_t key, int32_t bits);

/* Multi-GPU key sharding (no reference analogue: README.md:414 "No multi-GPU support yet").  Same kernel as
 * olap_radix_partition_i64 but the destination (one of 2^bits ranks) comes from mix64(key ^ salt), independent of
 * the bits the local join partitions on.  Partition p is the contiguous send region for rank p, so the output feeds
 * an NCCL all-to-all with no staging copy.  out_part_count (dptr, 2^bits u32, nullable): when given, every region
 * starts on a multiple of 16 rows (NCCL peer copies need 16-byte aligned segments to run at full NVLink speed), the
 * row count of region p is out_part_count[p], and the output arrays must hold n_rows + 16 * 2^bits rows.          */
OLAP_API olap_status olap_shuffle_partition_i64(olap_ctx* ctx, const int64_t* keys, const uint32_t* row_ids, uint64_t n_rows,
                                                int32_t bits, int64_t* out_keys, uint32_t* out_row_ids, uint32_t* out_part_begin,
                                                uint32_t* out_part_count, void* stream);
OLAP_API uint32_t olap_shuffle_dest_of(int64_t key, int32_t bits);

/* The same shuffle FUSED with the exchange over NVSwitch peer memory (2, 4 or 8 ranks: bits in [1, 3]).
 *   1. olap_shuffle_count_i64: rows this rank holds for every destination -> out_part_count (dptr, 2^bits u32).
 *   2. the caller all-gathers the count matrix and derives, for every destination p, the first row this rank owns in
 *      rank p's receive buffers (sum of the counts of the lower ranks).
 *   3. olap_shuffle_scatter_i64: the partition kernel stores (key, source row id) of every row directly at
 *      dest_keys[p][dest_row_offset[p] + ...] / dest_row_ids[p][...]: pointers into the PEERS' memory (e.g. torch symmetric
 *      memory buffer_ptrs), written run by run over NVLink -- no send buffer, no collective for the payload.
 * scratch (dptr): olap_shuffle_scratch_words(ctx, bits) u32, the same buffer for both calls, untouched in between.
 * The caller orders the stores against the readers (a barrier across ranks before and after step 3).            */
OLAP_API uint64_t olap_shuffle_scratch_words(olap_ctx* ctx, int32_t bits);
OLAP_API olap_status olap_shuffle_count_i64(olap_ctx* ctx, const int64_t* keys, uint64_t n_rows, int32_t bits, uint32_t* scratch,
                                            uint32_t* out_part_count, void* stream);
OLAP_API olap_status olap_shuffle_scatter_i64(olap_ctx* ctx, const int64_t* keys, uint64_t n_rows, int32_t bits, uint32_t* scratch,
                                              int64_t* const* dest_keys, uint32_t* const* dest_row_ids,
                                              const uint64_t* dest_row_offset, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Key-range-sharded join across the GPUs of one box (SURVEY 8(b) "multi-GPU" group; no reference analogue: README.md:414).
 * For a DENSE build-key domain (global max - min + 1 <= 5 x global build rows) the domain is cut into one contiguous slice per
 * rank; every rank's partition kernel stores its rows -- as packed 8-byte words {key - slice start, source row id} -- straight
 * into the owning rank's receive slab over NVLink (peer-mapped pointers), so the owner runs only the second partition level and
 * the shared-memory build + probe of the 1-GPU direct join.  Fixed-capacity, single-writer slab regions: no count pass, no count
 * collective.  The probe side moves in `probe_chunks` chunks so that chunk k + 1 is on the wire while chunk k is joined.
 *
 * One olap_shard per rank.  Two ways to drive a group:
 *   (a) one process, all ranks: olap_shard_ctx_create_all + olap_sharded_hash_join_i64 (stream events order the ranks; the
 *       devices may repeat, which is how the exchange is tested on a 1-GPU box);
 *   (b) one process per GPU: olap_shard_create on every rank, each rank provides a slab of olap_shard_slab_bytes() bytes that
 *       its peers can address (CUDA IPC / torch symmetric memory), olap_shard_attach(all ranks' pointers), then per join:
 *         olap_shard_key_range -> reduce min / max / rows over the ranks -> olap_shard_plan (same values on every rank)
 *         olap_shard_scatter(side 0) -> olap_shard_send -> BARRIER -> olap_shard_build
 *         for every chunk k: olap_shard_scatter(side 1, k) -> olap_shard_send -> BARRIER -> olap_shard_probe(k)   (separate streams: they overlap)
 *         olap_shard_finish -> OR the flags over the ranks; any flag set = discard, take the hash-sharded path
 *       BARRIER = "every rank's scatter of that (side, chunk) has completed": olap_shard_barrier on the scatter stream (a
 *       one-warp kernel signalling through the slabs), or any collective of the host language.  Before the first scatter of a join every rank must have finished
 *       reading the slabs of the previous join (same barrier).
 * Pairs index the rank's receive slabs; olap_shard_pairs_to_global turns them into global row ids
 * (row offset of the source rank + source row id) for both sides.
 * ------------------------------------------------------------------------------------------------------- */
typedef struct olap_shard olap_shard;
OLAP_API olap_status olap_shard_create(olap_ctx* ctx, int32_t rank, int32_t world, uint64_t build_rows_cap, uint64_t probe_rows_cap,
                                       int32_t probe_chunks, olap_shard** out);
OLAP_API olap_status olap_shard_destroy(olap_shard* shard);
OLAP_API uint64_t olap_shard_slab_bytes(const olap_shard* shard);
OLAP_API olap_status olap_shard_attach(olap_shard* shard, void* const* peer_slabs /* [world], [rank] = own */);
/* BARRIER for groups with one process per GPU: a one-warp kernel on `stream` that signals every peer through its slab and waits for
 * every peer's signal.  Every rank calls it the same number of times in the same order.  Not for ranks that share a GPU.       */
OLAP_API olap_status olap_shard_barrier(olap_shard* shard, void* stream);
/* the two halves of the barrier, for different streams: the scatter stream signals ("everything queued here before is done") and
 * goes on without waiting for anybody; the stream that consumes the chunk waits.  Wait number n matches every rank's signal n.  */
OLAP_API olap_status olap_shard_signal(olap_shard* shard, void* stream);
OLAP_API olap_status olap_shard_wait(olap_shard* shard, void* stream);
OLAP_API olap_status olap_shard_key_range(olap_shard* shard, const int64_t* keys, uint64_t n_rows, int64_t* out_min, int64_t* out_max, void* stream);
/* hot: 0 = uniform-ish probe keys (regions 5 % above an even share); 1 = heavy hitters: the olap_shard_hot_* sequence below is
 * mandatory and the regions get 25 % of room for the warm keys; 2 = the same room WITHOUT a heavy-hitter table -- for skew so mild
 * (no key carries more than ~0.02 % of the probe rows) that the lookups would only cost time.                                */
OLAP_API olap_status olap_shard_plan(olap_shard* shard, int64_t global_min, int64_t global_max, uint64_t global_build_rows, int32_t hot, void* stream);
OLAP_API olap_status olap_shard_scatter(olap_shard* shard, int32_t side, int32_t chunk, const int64_t* keys, uint64_t n_rows, uint32_t row_base,
                                        void* stream);
/* second half of a scatter, on a stream ordered behind it: the peer copies of the copy-engine mode (OLAP_SHARD_MODE=ce, default:
 * one cudaMemcpyAsync per destination rank, no SM involved) and the publication of this rank's row counts; then the BARRIER     */
OLAP_API olap_status olap_shard_send(olap_shard* shard, int32_t side, int32_t chunk, void* stream);
OLAP_API olap_status olap_shard_build(olap_shard* shard, void* stream);
OLAP_API olap_status olap_shard_probe(olap_shard* shard, int32_t chunk, olap_pair* out_pairs, uint64_t max_matches, void* stream);
OLAP_API olap_status olap_shard_finish(olap_shard* shard, uint64_t* host_count, uint64_t* host_hot_count /* nullable */, uint32_t* host_flags /* [4] */,
                                       void* stream);
/* Heavy hitters (Zipf-skewed probe keys), olap_shard_plan(hot = 1).  Order within a join:
 *   olap_shard_hot_candidates on every rank (samples its probe shard; keys as key - global_min + 1 in u32, with sample counts)
 *   -> the host language gathers the lists, sums the counts per key, keeps the <= 2048 most frequent -> olap_shard_hot_set
 *   build side as before (scatter, BARRIER, olap_shard_build)
 *   olap_shard_hot_resolve (8192 x u64 on the device) -> element-wise MAX over the ranks -> olap_shard_hot_finalize
 *   probe scatters: a row carrying a hot key never leaves its rank; its pair {hot slot, local probe row} goes to hot_pairs and the
 *   slot's build row is merged[slot] - 1 = (source rank << 32 | source row).                                                    */
OLAP_API olap_status olap_shard_hot_candidates(olap_shard* shard, const int64_t* probe_keys, uint64_t n_rows, uint32_t* host_keys_plus1,
                                               uint32_t* host_counts, uint32_t capacity, uint32_t* host_n, void* stream);
OLAP_API olap_status olap_shard_hot_set(olap_shard* shard, const uint32_t* keys_plus1, uint32_t n_keys, void* stream);
OLAP_API olap_status olap_shard_hot_resolve(olap_shard* shard, uint64_t* out_dptr /* 8192 x u64 */, void* stream);
OLAP_API olap_status olap_shard_hot_finalize(olap_shard* shard, const uint64_t* merged_dptr, olap_pair* hot_pairs, uint64_t max_hot_pairs, void* stream);
OLAP_API olap_status olap_shard_pairs_to_global(olap_shard* shard, const olap_pair* pairs, uint64_t n_pairs, const uint64_t* build_row_offsets_dptr,
                                                const uint64_t* probe_row_offsets_dptr, uint64_t* out_left, uint64_t* out_right, void* stream);
/* out8: words reserved for the build regions, words per probe chunk, build / probe region capacity, Q, F, byte offset of the
 * count arrays, chunks -- for callers that inspect the slabs */
OLAP_API olap_status olap_shard_layout(const olap_shard* shard, uint64_t* out8);
/* rows this rank stored into every rank's slab during the current join (the NVLink volume); synchronises the stream */
OLAP_API olap_status olap_shard_sent_rows(olap_shard* shard, uint64_t* host_rows_per_dest /* [world] */, void* stream);
/* Host-only (no GPU, no context): the heavy-hitter table exactly as olap_shard_hot_set lays it out (out_table: 8192 u32), the number
 * of keys placed, and the slot the kernels' lookup rule finds for every probe_plus1[i] (0xFFFFFFFF = not a heavy hitter).       */
OLAP_API olap_status olap_shard_hot_table_host(const uint32_t* keys_plus1, uint32_t n_keys, uint32_t* out_table, uint32_t* out_placed,
                                               const uint32_t* probe_plus1, uint32_t n_probe, uint32_t* out_slots);
/* Host-only (no GPU, no context): where the range shuffle sends keys of the dense domain [global_min, global_max] shared by `world`
 * ranks -- owning rank (-1: outside the domain, dropped), sub-range inside the owner's slab side, offset carried (key - slice start).
 * The same routing arithmetic as the kernels; out_geometry (nullable, 4 x u64): {Q, F, sub-range width in keys, slice width in keys}. */
OLAP_API olap_status olap_shard_route_i64(int32_t world, int64_t global_min, int64_t global_max, const int64_t* keys, uint64_t n,
                                          int32_t* out_rank, uint32_t* out_sub_range, uint32_t* out_offset, uint64_t* out_geometry);
OLAP_API olap_status olap_shard_ctx_create_all(int32_t n_ranks, const int32_t* devices, uint64_t max_gpu_memory, uint64_t build_rows_cap,
                                               uint64_t probe_rows_cap, int32_t probe_chunks, olap_ctx** out_ctxs, olap_shard** out_shards);
/* hot_pairs / max_hot_pairs / host_hot_counts: all NULL = no heavy-hitter handling; else per-rank buffers for the {hot slot, local
 * probe row} pairs of probe rows that carry a heavy-hitter key (olap_shard_hot_table gives the slots' build rows).             */
OLAP_API olap_status olap_sharded_hash_join_i64(olap_shard* const* shards, int32_t n_ranks, const int64_t* const* build_keys, const uint64_t* n_build,
                                                const int64_t* const* probe_keys, const uint64_t* n_probe, olap_pair* const* out_pairs,
                                                const uint64_t* max_matches, uint64_t* host_counts, olap_pair* const* hot_pairs,
                                                const uint64_t* max_hot_pairs, uint64_t* host_hot_counts);
OLAP_API olap_status olap_shard_hot_table(olap_shard* shard, uint64_t* host_merged /* [8192] */, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Aggregation.  Replaces AggregateParams/AggregateSpec/AggregateKernel::aggregate
 * (gpu-kernel-compiler/src/aggregate_kernel.rs:37-125) and hash_aggregate_kernel (join_kernels.cuh:216-240).
 * keys == NULL: one global group (what execute_aggregate does today, executor.rs:307).  Output: out_keys[g],
 * and agg-major 8-byte slots out_aggs[i * out_capacity + g] (aggregate_kernel.rs:98) typed per aggregate:
 * COUNT i64; SUM/MIN/MAX in the value's type (i64 SUM wraps); AVG finalised f64 (sum/count, 0.0 if count = 0).
 * Groups come out in unspecified order.  n_groups_hint is the caller's upper bound (AggregateParams.n_groups);
 * a wrong hint costs speed, never correctness.  Optional filter terms are fused into the same pass.
 * ------------------------------------------------------------------------------------------------------- */
typedef struct {
  const void* values;  /* dptr, 8-byte column (ignored for COUNT; may be NULL) */
  int32_t func;        /* OLAP_SUM.. */
  int32_t value_type;  /* OLAP_I64 / OLAP_F64 */
} olap_agg_spec;

OLAP_API olap_status olap_group_by_i64(olap_ctx* ctx, const int64_t* keys, const olap_agg_spec* specs, int32_t n_aggs,
                                       uint64_t n_rows, uint64_t n_groups_hint, const olap_filter_term* terms,
                                       int32_t n_terms, int64_t* out_keys, void* out_aggs, uint64_t out_capacity,
                                       uint64_t* out_n_groups_dptr, uint64_t* host_n_groups, void* stream);
/* Aggregate inputs computed on the fly (no reference analogue: the reference materialises every projection first,
 * executor.rs:184-214).  progs[i] == NULL: aggregate i reads specs[i].values as before.  Otherwise specs[i].values is
 * ignored and the value is the result of a small arithmetic program in reverse Polish notation over 8-byte device columns
 * and literals, all of type specs[i].value_type, evaluated inside the aggregation kernel with eval_binary_expr's semantics
 * (i64 wraps, x / 0 = 0; f64 IEEE):  op 'c' pushes column[k], 'l' pushes literal_bits[k], '+' '-' '*' '/' replace the two
 * topmost entries (second-from-top OP top).  At most OLAP_PROG_MAX_OPS entries, stack depth <= 3.  Aggregates that share
 * one program pointer share one evaluation.  Returns OLAP_ERR_UNSUPPORTED when the cardinality hint needs the partitioned
 * regime (which carries plain columns only): materialise the expression and call olap_group_by_i64 instead.          */
#define OLAP_PROG_MAX_OPS 10
typedef struct {
  int32_t n_ops;
  int8_t op[OLAP_PROG_MAX_OPS];
  const void* column[OLAP_PROG_MAX_OPS];
  int64_t literal_bits[OLAP_PROG_MAX_OPS];
} olap_expr_prog;
OLAP_API olap_status olap_group_by_prog_i64(olap_ctx* ctx, const int64_t* keys, const olap_agg_spec* specs,
                                            const olap_expr_prog* const* progs, int32_t n_aggs, uint64_t n_rows,
                                            uint64_t n_groups_hint, const olap_filter_term* terms, int32_t n_terms,
                                            int64_t* out_keys, void* out_aggs, uint64_t out_capacity,
                                            uint64_t* out_n_groups_dptr, uint64_t* host_n_groups, void* stream);
/* ... and a computed GROUP BY key: key_prog (Int64 program, same form) replaces the key column -- "GROUP BY returnflag * 2 +
 * linestatus" without the projection passes that would write the key column and read it back.  Same regimes and the same
 * OLAP_ERR_UNSUPPORTED rule as olap_group_by_prog_i64; progs may be NULL.                                             */
OLAP_API olap_status olap_group_by_expr_i64(olap_ctx* ctx, const olap_expr_prog* key_prog, const olap_agg_spec* specs,
                                            const olap_expr_prog* const* progs, int32_t n_aggs, uint64_t n_rows,
                                            uint64_t n_groups_hint, const olap_filter_term* terms, int32_t n_terms,
                                            int64_t* out_keys, void* out_aggs, uint64_t out_capacity,
                                            uint64_t* out_n_groups_dptr, uint64_t* host_n_groups, void* stream);
/* AggFunc::identity_value (aggregate_kernel.rs:27-33) */
OLAP_API int64_t olap_agg_identity(int32_t func);
/* compute_launch_params (filter_kernel.rs:75-79, aggregate_kernel.rs:127-131), without the u32 truncation bug */
OLAP_API void olap_compute_launch_params(uint64_t n_rows, uint32_t* grid, uint32_t* block);

/* GROUP BY across the GPUs of one box, one process driving all ranks (SURVEY.md 8(b) "multi-GPU", 8(e); no reference analogue,
 * README.md:414).  ctxs[r] is rank r's context (olap_ctx_create / olap_shard_ctx_create_all; the devices may repeat), keys[r] /
 * specs[r][n_aggs] / n_rows[r] its rows -- every rank passes the same aggregates (function, value type), each over its own column.
 * Local pre-aggregation into partials (AVG travels as sum + count, aggregate_kernel.rs:22-23; every COUNT / AVG shares one row
 * count), the partial rows go to the rank that owns their key (olap_shuffle_dest_of(key, log2 n_ranks)) with one peer copy per
 * (source, owner, column), the owner merges and finalises.  Rank r's out_keys[r] / out_aggs[r] (agg-major, stride out_capacity[r],
 * typed like olap_group_by_i64's) receive the host_n_groups[r] groups it owns; the union over the ranks is the result.  n_ranks is
 * 1, 2, 4 or 8.  OLAP_ERR_CAPACITY when some rank owns more groups than its out_capacity (host_n_groups is exact for every rank).
 * Synchronous: the results are complete on return.                                                                          */
OLAP_API olap_status olap_sharded_group_by_i64(olap_ctx* const* ctxs, int32_t n_ranks, const int64_t* const* keys,
                                               const olap_agg_spec* const* specs, int32_t n_aggs, const uint64_t* n_rows,
                                               uint64_t n_groups_hint, int64_t* const* out_keys, void* const* out_aggs,
                                               const uint64_t* out_capacity, uint64_t* host_n_groups);


/* ---------------------------------------------------------------------------------------------------------
 * Operator API: the drop-in boundary proper.  Mirrors PhysicalPlan / PhysicalExpr / AggregateExpr
 * (gpu-olap-core/src/physical_plan.rs:11-84) and executor::execute (gpu-olap-core/src/executor.rs:15-24) over
 * the Arrow C Data Interface -- the same ArrowArray / ArrowSchema structs arrow-rs's `arrow::ffi` and pyarrow
 * exchange, so the Rust side is a mechanical wrapper.  Builders take ownership of their
 * children; olap_plan_free / olap_expr_free release a whole tree.
 * ------------------------------------------------------------------------------------------------------- */
#ifndef ARROW_C_DATA_INTERFACE
#define ARROW_C_DATA_INTERFACE
struct ArrowSchema {
  const char* format; const char* name; const char* metadata; int64_t flags; int64_t n_children;
  struct ArrowSchema** children; struct ArrowSchema* dictionary;
  void (*release)(struct ArrowSchema*); void* private_data;
};
struct ArrowArray {
  int64_t length; int64_t null_count; int64_t offset; int64_t n_buffers; int64_t n_children;
  const void** buffers; struct ArrowArray** children; struct ArrowArray* dictionary;
  void (*release)(struct ArrowArray*); void* private_data;
};
#endif

typedef struct olap_expr olap_expr;
typedef struct olap_plan olap_plan;
typedef struct olap_result olap_result;

/* JoinType, gpu-olap-core/src/logical_plan.rs:52-58.  The reference executes none of them (README.md:410-411 lists the outer
 * ones as missing); here Left / Right / Full append the unmatched rows of the preserved side(s) behind the inner pairs, the other
 * side's columns NULL there (Arrow validity bitmaps in the exported batch).                                                    */
enum { OLAP_JOIN_INNER = 0, OLAP_JOIN_LEFT = 1, OLAP_JOIN_RIGHT = 2, OLAP_JOIN_FULL = 3 };

/* PhysicalExpr::{Column{index,name}, Literal{value}, BinaryExpr{left,op,right}} (physical_plan.rs:66-75).  Columns
 * resolve by name first, then by index (executor.rs:381-387); literals parse as i64, else f64 (:390-401).       */
OLAP_API olap_expr* olap_expr_column(uint64_t index, const char* name);
OLAP_API olap_expr* olap_expr_literal(const char* value);
OLAP_API olap_expr* olap_expr_binary(olap_expr* left, const char* op, olap_expr* right);
OLAP_API void olap_expr_free(olap_expr* e);

/* PhysicalPlan::GpuTableScan{table_name, schema, projection} (physical_plan.rs:13-17).  Rows come from the batches
 * registered under table_name (the in-memory cache of catalog.rs:83-87); with none registered the scan yields one
 * EMPTY batch typed from `schema` exactly as execute_table_scan does (executor.rs:110-155).  schema may be NULL when
 * the table is registered.  projection: column indices or NULL.                                               */
OLAP_API olap_plan* olap_plan_table_scan(const char* table_name, const struct ArrowSchema* schema, const int64_t* projection,
                                         int32_t n_projection);
OLAP_API olap_plan* olap_plan_filter(olap_plan* input, olap_expr* predicate);                         /* GpuFilter     :20-23 */
OLAP_API olap_plan* olap_plan_projection(olap_plan* input, olap_expr* const* exprs, int32_t n_exprs); /* GpuProjection :26-29 */
OLAP_API olap_plan* olap_plan_hash_join(olap_plan* left, olap_plan* right, olap_expr* const* left_keys, int32_t n_left_keys,
                                        olap_expr* const* right_keys, int32_t n_right_keys, int32_t join_type); /* GpuHashJoin :32-39 */
/* GpuAggregate{input, group_by, aggr_exprs} (:52-57); aggr_exprs[i] = AggregateExpr::<func>{input} (:77-84) */
OLAP_API olap_plan* olap_plan_aggregate(olap_plan* input, olap_expr* const* group_by, int32_t n_group_by, const int32_t* agg_funcs,
                                        olap_expr* const* agg_inputs, int32_t n_aggs);
OLAP_API void olap_plan_free(olap_plan* p);

/* Catalog::get_table_data analogue: import one RecordBatch (struct-typed ArrowArray) into device memory and append
 * it to table_name.  Numeric columns are widened to 8 bytes ON DEVICE, Utf8 becomes FNV-1a-64 as i64, nulls become 0
 * (arrow-interop/src/record_batch_convert.rs:35-130) and the column keeps a byte-per-row validity (:37-41) that scans carry and
 * olap_result_export turns back into an Arrow bitmap; filters, projections, joins and aggregates read a null as 0 / 0.0 like the
 * reference does (:103, :108).  The array is released in every case, also when the call fails.                              */
OLAP_API olap_status olap_register_batch(olap_ctx* ctx, const char* table_name, struct ArrowArray* batch, struct ArrowSchema* schema);
/* The same for columns that are ALREADY device-resident (an upstream operator's output, a device-side loader, synthetic data):
 * n_cols 8-byte columns of n_rows rows each, types[i] = OLAP_I64 / OLAP_F64, are copied device-to-device (stream-ordered, no host
 * round trip) into one batch appended to table_name.  validity[i] (nullable array; entries nullable): one byte per row, 1 = valid. */
OLAP_API olap_status olap_register_device_batch(olap_ctx* ctx, const char* table_name, int32_t n_cols, const char* const* names,
                                                const int32_t* types, const void* const* columns_dptr,
                                                const uint8_t* const* validity_dptr, uint64_t n_rows);
OLAP_API olap_status olap_drop_table(olap_ctx* ctx, const char* table_name);

/* executor::execute(plan, config) -> Vec<RecordBatch> (executor.rs:15-24).  Results are exported as Arrow C Data
 * (host memory owned by the exported array; its release callback frees it).                                    */
OLAP_API olap_status olap_execute(olap_ctx* ctx, const olap_plan* plan, olap_result** out);
OLAP_API int32_t olap_result_num_batches(const olap_result* r);
OLAP_API int64_t olap_result_num_rows(const olap_result* r, int32_t batch);
OLAP_API olap_status olap_result_export(olap_result* r, int32_t batch, struct ArrowArray* out_array, struct ArrowSchema* out_schema);
OLAP_API void olap_result_free(olap_result* r);

/* ---------------------------------------------------------------------------------------------------------
 * Synthetic data on the device (counter-based splitmix64(seed, row); any shard can generate
 * its slice).  kind: 0 uniform i64 in [0, domain); 1 permutation of [0, domain) evaluated at row0 + i;
 * 2 dyadic f64 = (u % domain) / 1024; 3 uniform f64 in [0, 1).                                              */
OLAP_API olap_status olap_generate(olap_ctx* ctx, void* out_dptr, uint64_t n_rows, int32_t kind, uint64_t seed,
                                   uint64_t row0, uint64_t domain, void* stream);

/* Probe-side foreign keys for a build side generated with kind 1 (perm_seed, domain): row r carries the key of build row b(r),
 * b uniform over [0, domain) for zipf_s = 0, else b + 1 ~ Zipf(zipf_s) (continuous inverse-CDF sampler) -- BASELINE config 5's
 * skew: the hot build rows are rows 0, 1, 2, ... and the permutation spreads their keys over the domain.                  */
OLAP_API olap_status olap_generate_fk(olap_ctx* ctx, void* out_dptr, uint64_t n_rows, uint64_t perm_seed, uint64_t domain, uint64_t fk_seed,
                                      uint64_t row0, double zipf_s, void* stream);

/* per-phase device timings of the last join / group-by on this ctx (ms, CUDA events on the call's stream);
 * names is a ';'-separated list valid until the next call.  Enabled by olap_set_profiling(ctx, 1).          */
OLAP_API olap_status olap_set_profiling(olap_ctx* ctx, int32_t enabled);
OLAP_API olap_status olap_last_phase_times(olap_ctx* ctx, const char** names, const float** ms, int32_t* n);
/* number of kernels launched by this ctx since creation (bench.py's gpu_launches) */
OLAP_API uint64_t olap_launch_count(olap_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* OLAP_B200_H */
