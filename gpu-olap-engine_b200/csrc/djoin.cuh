// Interface of djoin.cu: the direct-address join over key ranges with shared-memory tables (dense, unique build keys).
#pragma once
#include "common.cuh"

namespace olap {

enum { kFlagOverflow = 0, kFlagOutOfRange = 1, kFlagDup = 2, kFlagWords = 4 };

// the rows of one side of the join: a column of raw int64 keys, or -- multi-GPU -- the regions the shuffle filled
struct RInput {
  const uint64_t* keys = nullptr; uint64_t n = 0; uint32_t id_base = 0;   // raw keys; row id = id_base + row
  // n_reg regions of reg_cap packed words {key - rmin : u32 hi, source row id : u32 lo}, row counts on the device; region g
  // holds keys of sub-range g % reg_q, reg_parts * 2^15 keys wide.  Row ids become positions id_base + g * reg_cap + row.
  const uint64_t* regions = nullptr; const uint32_t* reg_cur = nullptr; uint32_t reg_cap = 0, n_reg = 0, reg_q = 0, reg_parts = 0;
};

// The build side, partitioned by key range once and probed by one or many batches.
struct DJBuild {
  olap_ctx* ctx = nullptr;
  int64_t rmin = 0; uint64_t rrange = 0;
  uint32_t P = 0, PP = 0;
  uint64_t* brows = nullptr;   // [P][2^15] packed rows {key - rmin : u32 hi, row id : u32 lo}
  uint32_t* bctl = nullptr;    // [PP] cursors | kFlagWords flags
};

bool djoin_eligible(uint64_t rrange);
uint32_t djoin_region_keys_log2();
olap_status dj_build_prepare(olap_ctx* ctx, const RInput& in, int64_t rmin, uint64_t rrange, unsigned grid_limit, cudaStream_t s, DJBuild* b);
olap_status dj_probe(const DJBuild& b, const RInput& in, uint64_t np_cap, olap_pair* out, uint64_t max_matches, unsigned long long* d_count,
                     int hot, unsigned grid_limit, cudaStream_t s, uint32_t* d_flags);
void dj_build_free(DJBuild& b, cudaStream_t s);

// a group of probe batches joined with one build of every shared-memory table (dj_probe_begin / _add ... / _join / _end)
struct DJProbe {
  const DJBuild* b = nullptr;
  int hot = 0; uint32_t slack = 25, pcap = 0; uint64_t np_cap = 0;
  uint64_t* prows = nullptr;   // [P][pcap] packed probe rows
  uint32_t* ctl = nullptr;     // [PP] cursors | ticket (4) | hot keys | hot ids | hot_n (4)
};
olap_status dj_probe_begin(const DJBuild& b, uint64_t np_cap, int hot, cudaStream_t s, DJProbe* p);
olap_status dj_probe_add(DJProbe& p, const RInput& in, olap_pair* out, uint64_t max_matches, unsigned long long* d_count, unsigned grid_limit,
                         cudaStream_t s, uint32_t* d_flags);
olap_status dj_probe_join(DJProbe& p, olap_pair* out, uint64_t max_matches, unsigned long long* d_count, unsigned grid_limit, cudaStream_t s,
                          uint32_t* d_flags);
void dj_probe_end(DJProbe& p, cudaStream_t s);

// Routing arithmetic of the multi-GPU shuffle, shared by the kernels (djoin.cu) and the host-side olap_shard_route (shard.cu), which
// the CPU suite checks against plain integer division.  Exact division by multiplication: sub-ranges are F x 2^15 keys wide with
// F <= 256 and offsets < 2^32, so x = offset >> 15 < 2^17 and (x * ceil(2^32 / F)) >> 32 == x / F (the error term x / 2^32 < 2^-15 < 1 / F);
// partitions < 256: (p * ceil(2^24 / Q)) >> 24 == p / Q.
__host__ __device__ __forceinline__ uint64_t route_mul_f(uint64_t F) { return ((1ULL << 32) + F - 1) / F; }
__host__ __device__ __forceinline__ uint32_t route_mul_q(uint32_t Q) { return (uint32_t)(((1u << 24) + Q - 1) / Q); }
__host__ __device__ __forceinline__ uint32_t route_part(uint32_t offset, uint64_t mf) { return (uint32_t)(((uint64_t)(offset >> 15) * mf) >> 32); }
__host__ __device__ __forceinline__ uint32_t route_dest(uint32_t part, uint32_t mq) { return (part * mq) >> 24; }

// heavy-hitter table handed to a pass over raw probe keys: rows whose key is in it are joined on the spot
struct HotRef { const uint32_t* hot_keys = nullptr; const uint32_t* hot_ids = nullptr; olap_pair* pairs = nullptr; uint64_t max_matches = 0; unsigned long long* count = nullptr; };
uint32_t hot_table_slots();
uint32_t hot_max_keys();   // at most a quarter of the slots hold a key
uint32_t hot_table_build_host(const uint32_t* keys_plus1, uint32_t n, uint32_t* tab);
uint32_t hot_lookup_host(const uint32_t* tab, uint32_t want);
uint32_t hot_sample_slots();
olap_status hot_candidates(olap_ctx* ctx, const uint64_t* keys, uint64_t n, int64_t gmin, uint64_t grange, uint32_t* scratch, uint32_t cap,
                           uint32_t** out_key, uint32_t** out_cnt, uint32_t** out_n, cudaStream_t s);
olap_status hot_resolve_global(olap_ctx* ctx, const uint32_t* hot_keys, uint64_t* out, const DJBuild& b, const uint64_t* slab_b, uint32_t slab_bcap,
                               uint32_t q, uint64_t slice, int rank, cudaStream_t s);
olap_status hot_finalize(olap_ctx* ctx, uint32_t* hot_keys, uint32_t* hot_ids, const uint64_t* merged, cudaStream_t s);
olap_status range_shuffle(olap_ctx* ctx, const uint64_t* keys, uint64_t n, uint32_t id_base, int64_t gmin, uint64_t grange, uint32_t q, uint32_t subw,
                          int rank, int world, uint64_t* const* peer_regions, uint32_t cap, uint32_t* cursor, uint32_t* flags, bool build_side,
                          const HotRef* hot, unsigned grid_limit, bool tma, cudaStream_t s);
olap_status range_shuffle_publish(olap_ctx* ctx, const uint32_t* cursor, uint32_t cap, uint32_t q, int rank, int world, uint32_t* const* peer_counts,
                                  cudaStream_t s);

}  // namespace olap
