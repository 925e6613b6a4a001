// k3_probe.cuh -- device side of one hash-table lookup (shared by the probe kernel of the replicated
// index, k3_table.cu, and the owner-side probe of the key-partitioned index, partition.cu).
#pragma once
#include "k3_table.cuh"
#include "hrm_common.cuh"

namespace hrm {

// owner of a key in a G-way key partition: multiply-shift reduction of the HIGH hash word (the home bucket
// uses the low word, so shard and bucket are independent)
HRM_HD uint32_t key_owner(uint64_t key, uint32_t world)
{
    const uint64_t h = murmur64(key + 0x5ad0dedULL);
    return (uint32_t)(((uint64_t)(uint32_t)(h >> 32) * (uint64_t)world) >> 32);
}

#if defined(__CUDACC__)
// home bucket of a key: multiply-shift range reduction of the low hash word (nbuckets need not be a
// power of two); the probe sequence continues linearly over buckets
__device__ __forceinline__ uint32_t home_bucket(uint64_t key, uint32_t nbuckets)
{
    const uint64_t h = murmur64(key + 0x5ad0dedULL);
    return __umulhi((uint32_t)h, nbuckets);
}
__device__ __forceinline__ uint32_t next_bucket(uint32_t b, uint32_t nbuckets) { return b + 1 == nbuckets ? 0u : b + 1; }

// the 64 bytes of one bucket: two 256-bit loads that bypass L1 allocation (no reuse between lookups)
__device__ __forceinline__ void load_bucket(const Slot* bucket, uint64_t (&w)[8])
{
    asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(w[0]), "=l"(w[1]), "=l"(w[2]), "=l"(w[3])
                 : "l"(bucket));
    asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(w[4]), "=l"(w[5]), "=l"(w[6]), "=l"(w[7])
                 : "l"(bucket + 2));
}

// One lookup by one thread.  Returns (off, cnt); visited += buckets read.
__device__ __forceinline__ uint2 probe_bucket_sequence(const Slot* __restrict__ slots, uint32_t nbuckets, uint64_t key,
                                           uint32_t max_results, uint32_t& visited)
{
    uint2 res = make_uint2(0u, 0u);
    if (key == SLOT_EMPTY) return res; // invalid signature (len < k)
    uint32_t b = home_bucket(key, nbuckets);
    for (uint32_t probe = 0; probe < nbuckets; probe++) {
        uint64_t w[8];
        load_bucket(slots + (size_t)b * BUCKET_SLOTS, w);
        visited += 1;
        bool hit = false, empty = false;
        uint64_t pay = 0;
#pragma unroll
        for (int i = 0; i < BUCKET_SLOTS; i++) {
            if (w[2 * i] == key) {
                hit = true;
                pay = w[2 * i + 1]; // off | cnt << 32
            }
            empty |= w[2 * i] == SLOT_EMPTY;
        }
        if (hit) {
            const uint32_t cnt = (uint32_t)(pay >> 32);
            if (cnt <= max_results) res = make_uint2((uint32_t)pay, cnt); // ref: fakegpuminhasher.cuh:280-285
            break;
        }
        if (empty) break; // a free slot in the bucket ends the probe sequence: key absent
        b = next_bucket(b, nbuckets);
    }
    return res;
}

#endif // __CUDACC__

} // namespace hrm
