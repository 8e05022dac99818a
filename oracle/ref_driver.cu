// TEST INFRASTRUCTURE ONLY.  Host driver around the reference's own CUDA kernels
// (gpu-kernel-compiler/kernels/join_kernels.cuh, included from the read-only reference checkout via -I; nothing is
// copied).  Built by oracle/Makefile into oracle/_ref/libref_kernels.so.  The reference never wrote the host side
// (histogram + exclusive scan for partition_offsets, 0xFFFFFFFF bucket fill, launch configs): that glue is here.
// All pointers are HOST pointers; every call allocates, runs and frees on device 0.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "join_kernels.cuh"

#define RK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "ref_driver: %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e_)); return -1; } } while (0)
#define REF_API extern "C" __attribute__((visibility("default")))

__global__ void ref_hash_probe_kernel(const int64_t* k64, uint32_t n64, uint32_t* o64, const int32_t* k32, uint32_t n32, uint32_t* o32) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n64) o64[t] = hash_int64(k64[t]);
  if (t < n32) o32[t] = hash_int32(k32[t]);
}

template <class T> struct DevBuf {
  T* p = nullptr;
  cudaError_t alloc(size_t n) { return cudaMalloc(&p, (n ? n : 1) * sizeof(T)); }
  ~DevBuf() { if (p) cudaFree(p); }
};

static inline unsigned blocks(uint32_t n) { return n ? (n + 255) / 256 : 1; }  // "one thread per row", block 256 (README.md:107)

REF_API int ref_hash_vectors(const int64_t* k64, uint32_t n64, uint32_t* o64, const int32_t* k32, uint32_t n32, uint32_t* o32) {
  DevBuf<int64_t> a; DevBuf<int32_t> b; DevBuf<uint32_t> oa, ob;
  RK(a.alloc(n64)); RK(b.alloc(n32)); RK(oa.alloc(n64)); RK(ob.alloc(n32));
  RK(cudaMemcpy(a.p, k64, n64 * 8, cudaMemcpyHostToDevice));
  RK(cudaMemcpy(b.p, k32, n32 * 4, cudaMemcpyHostToDevice));
  uint32_t n = n64 > n32 ? n64 : n32;
  ref_hash_probe_kernel<<<blocks(n), 256>>>(a.p, n64, oa.p, b.p, n32, ob.p);
  RK(cudaDeviceSynchronize());
  RK(cudaMemcpy(o64, oa.p, n64 * 4, cudaMemcpyDeviceToHost));
  RK(cudaMemcpy(o32, ob.p, n32 * 4, cudaMemcpyDeviceToHost));
  return 0;
}

// radix_partition_kernel<int64_t> with partition_offsets pre-seeded by an exclusive scan of the histogram.
// part_begin[257] receives the partition boundaries.
REF_API int ref_radix_partition(const int64_t* keys, const uint32_t* row_ids, uint32_t n, int radix_shift, uint32_t* part_begin,
                                int64_t* out_keys, uint32_t* out_ids) {
  DevBuf<int64_t> dk, ok; DevBuf<uint32_t> di, oi, dh, off;
  RK(dk.alloc(n)); RK(ok.alloc(n)); RK(di.alloc(n)); RK(oi.alloc(n)); RK(dh.alloc(n)); RK(off.alloc(256));
  RK(cudaMemcpy(dk.p, keys, (size_t)n * 8, cudaMemcpyHostToDevice));
  if (row_ids) RK(cudaMemcpy(di.p, row_ids, (size_t)n * 4, cudaMemcpyHostToDevice));
  ref_hash_probe_kernel<<<blocks(n), 256>>>(dk.p, n, dh.p, nullptr, 0, nullptr);
  std::vector<uint32_t> h(n);
  RK(cudaMemcpy(h.data(), dh.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
  memset(part_begin, 0, 257 * 4);
  for (uint32_t i = 0; i < n; i++) part_begin[((h[i] >> radix_shift) & (NUM_RADIX_PARTITIONS - 1)) + 1]++;
  for (int p = 0; p < 256; p++) part_begin[p + 1] += part_begin[p];
  RK(cudaMemcpy(off.p, part_begin, 256 * 4, cudaMemcpyHostToDevice));
  radix_partition_kernel<int64_t><<<blocks(n), 256>>>(dk.p, row_ids ? di.p : nullptr, n, off.p, ok.p, oi.p, radix_shift);
  RK(cudaDeviceSynchronize());
  RK(cudaMemcpy(out_keys, ok.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
  RK(cudaMemcpy(out_ids, oi.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
  return 0;
}

// build_hash_table_kernel + probe_hash_table_kernel over the whole (unpartitioned) inputs.
// capacity = 0 applies the README sizing rule 1.5 x build rows (README.md:114).
REF_API int ref_hash_join(const int64_t* build, uint32_t nb, const int64_t* probe, uint32_t np, uint32_t capacity, uint32_t* left,
                          uint32_t* right, uint32_t max_matches, uint32_t* count) {
  if (!capacity) capacity = nb + nb / 2 + 1;
  DevBuf<int64_t> db, dp; DevBuf<HashEntry> ent; DevBuf<uint32_t> bkt, cnt, l, r;
  RK(db.alloc(nb)); RK(dp.alloc(np)); RK(ent.alloc(nb)); RK(bkt.alloc(capacity)); RK(cnt.alloc(1)); RK(l.alloc(max_matches)); RK(r.alloc(max_matches));
  RK(cudaMemcpy(db.p, build, (size_t)nb * 8, cudaMemcpyHostToDevice));
  RK(cudaMemcpy(dp.p, probe, (size_t)np * 8, cudaMemcpyHostToDevice));
  RK(cudaMemset(bkt.p, 0xFF, (size_t)capacity * 4));
  RK(cudaMemset(cnt.p, 0, 4));
  if (nb) build_hash_table_kernel<int64_t><<<blocks(nb), 256>>>(db.p, nullptr, nb, ent.p, bkt.p, capacity);
  if (np) probe_hash_table_kernel<int64_t><<<blocks(np), 256>>>(dp.p, nullptr, np, ent.p, bkt.p, capacity, db.p, cnt.p, l.p, r.p, max_matches);
  RK(cudaDeviceSynchronize());
  RK(cudaMemcpy(count, cnt.p, 4, cudaMemcpyDeviceToHost));
  uint32_t k = *count < max_matches ? *count : max_matches;
  RK(cudaMemcpy(left, l.p, (size_t)k * 4, cudaMemcpyDeviceToHost));
  RK(cudaMemcpy(right, r.p, (size_t)k * 4, cudaMemcpyDeviceToHost));
  return 0;
}

// The full intended radix flow (README.md:92-137): partition both sides with radix_partition_kernel, then per
// partition build + probe; build_keys passed to the probe is the ORIGINAL column (join_kernels.cuh:152).
REF_API int ref_radix_hash_join(const int64_t* build, uint32_t nb, const int64_t* probe, uint32_t np, uint32_t* left, uint32_t* right,
                                uint32_t max_matches, uint32_t* count) {
  std::vector<uint32_t> bb(257), pb(257), bi(nb ? nb : 1), pi(np ? np : 1);
  std::vector<int64_t> bk(nb ? nb : 1), pk(np ? np : 1);
  if (ref_radix_partition(build, nullptr, nb, 0, bb.data(), bk.data(), bi.data())) return -1;
  if (ref_radix_partition(probe, nullptr, np, 0, pb.data(), pk.data(), pi.data())) return -1;
  DevBuf<int64_t> dorig, dbk, dpk; DevBuf<uint32_t> dbi, dpi, bkt, cnt, l, r; DevBuf<HashEntry> ent;
  uint32_t max_part = 0;
  for (int p = 0; p < 256; p++) if (bb[p + 1] - bb[p] > max_part) max_part = bb[p + 1] - bb[p];
  const uint32_t max_cap = max_part + max_part / 2 + 1;
  RK(dorig.alloc(nb)); RK(dbk.alloc(nb)); RK(dpk.alloc(np)); RK(dbi.alloc(nb)); RK(dpi.alloc(np)); RK(ent.alloc(max_part));
  RK(bkt.alloc(max_cap)); RK(cnt.alloc(1)); RK(l.alloc(max_matches)); RK(r.alloc(max_matches));
  RK(cudaMemcpy(dorig.p, build, (size_t)nb * 8, cudaMemcpyHostToDevice));
  RK(cudaMemcpy(dbk.p, bk.data(), (size_t)nb * 8, cudaMemcpyHostToDevice));
  RK(cudaMemcpy(dpk.p, pk.data(), (size_t)np * 8, cudaMemcpyHostToDevice));
  RK(cudaMemcpy(dbi.p, bi.data(), (size_t)nb * 4, cudaMemcpyHostToDevice));
  RK(cudaMemcpy(dpi.p, pi.data(), (size_t)np * 4, cudaMemcpyHostToDevice));
  RK(cudaMemset(cnt.p, 0, 4));
  for (int p = 0; p < 256; p++) {
    const uint32_t n_b = bb[p + 1] - bb[p], n_p = pb[p + 1] - pb[p];
    if (!n_b || !n_p) continue;
    const uint32_t cap = n_b + n_b / 2 + 1;
    RK(cudaMemset(bkt.p, 0xFF, (size_t)cap * 4));
    build_hash_table_kernel<int64_t><<<blocks(n_b), 256>>>(dbk.p + bb[p], dbi.p + bb[p], n_b, ent.p, bkt.p, cap);
    probe_hash_table_kernel<int64_t><<<blocks(n_p), 256>>>(dpk.p + pb[p], dpi.p + pb[p], n_p, ent.p, bkt.p, cap, dorig.p, cnt.p, l.p, r.p, max_matches);
  }
  RK(cudaDeviceSynchronize());
  RK(cudaMemcpy(count, cnt.p, 4, cudaMemcpyDeviceToHost));
  uint32_t k = *count < max_matches ? *count : max_matches;
  RK(cudaMemcpy(left, l.p, (size_t)k * 4, cudaMemcpyDeviceToHost));
  RK(cudaMemcpy(right, r.p, (size_t)k * 4, cudaMemcpyDeviceToHost));
  return 0;
}

// hash_aggregate_kernel<unsigned long long>: direct-address SUM (join_kernels.cuh:216-240)
REF_API int ref_hash_aggregate_u64(const uint32_t* group_keys, const unsigned long long* values, uint32_t n, uint32_t num_groups,
                                   unsigned long long* aggregates) {
  DevBuf<uint32_t> gk, gid; DevBuf<unsigned long long> v, agg;
  RK(gk.alloc(n)); RK(gid.alloc(n)); RK(v.alloc(n)); RK(agg.alloc(num_groups));
  RK(cudaMemcpy(gk.p, group_keys, (size_t)n * 4, cudaMemcpyHostToDevice));
  RK(cudaMemcpy(v.p, values, (size_t)n * 8, cudaMemcpyHostToDevice));
  RK(cudaMemset(agg.p, 0, (size_t)num_groups * 8));
  if (n) hash_aggregate_kernel<unsigned long long><<<blocks(n), 256>>>(gk.p, v.p, n, gid.p, agg.p, num_groups);
  RK(cudaDeviceSynchronize());
  RK(cudaMemcpy(aggregates, agg.p, (size_t)num_groups * 8, cudaMemcpyDeviceToHost));
  return 0;
}
