// partition.cuh -- internal interface of the key-partitioned index (partition.cu).
#pragma once
#include <nccl.h>
#include "k3_probe.cuh"
#include "mapper.hpp"

struct hrm_comm;

namespace hrm {
int comm_rank(const hrm_comm* c);
int comm_world(const hrm_comm* c);
hrm_status comm_max_i64(hrm_comm* c, int64_t* v, cudaStream_t s);
hrm_status comm_agree(hrm_comm* c, hrm_status local, cudaStream_t s);
// count + retrieve of one batch through the partitioned tables (collective over the communicator):
// d_num_per_seq [n], d_offsets [n + 1], `values` allocated to *h_total entries in table order
hrm_status partitioned_query(hrm_comm* c, hrm_minhasher* mh, const uint64_t* d_sigs, int n, int32_t* d_num_per_seq,
                             int32_t* d_offsets, int64_t* h_total, Scratch& values, StageTimer& T, cudaStream_t s,
                             uint2* d_ranges = nullptr); // d_ranges [n][H]: (offset in values, count) per (read, table)
} // namespace hrm
