// Interface of djoin.cu: the direct-address join over key ranges with shared-memory tables (dense, unique build keys).
#pragma once
#include "common.cuh"

namespace olap {

enum { kFlagOverflow = 0, kFlagOutOfRange = 1, kFlagDup = 2, kFlagWords = 4 };

// the rows of one side of the join: a column of raw int64 keys, or -- multi-GPU -- the regions the shuffle filled
struct RInput {
  const uint64_t* keys = nullptr; uint64_t n = 0; uint32_t id_base = 0;   // raw keys; row id = id_base + row
  // n_reg regions of reg_cap packed words {key - rmin : u32 hi, source row id : u32 lo}, row counts on the device; region g
  // holds keys of sub-range (g & reg_mask), 2^(15 + reg_bits) keys wide.  Row ids become positions g * reg_cap + row.
  const uint64_t* regions = nullptr; const uint32_t* reg_cur = nullptr; uint32_t reg_cap = 0, n_reg = 0, reg_mask = 0; int reg_bits = 0;
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

}  // namespace olap
