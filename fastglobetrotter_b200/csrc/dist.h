// dist.h -- the two exchange steps of a sharded call, over NCCL:
//   * one all-gather of the per-(chromosome, individual) pair counts, so that every shard can place its draws in the
//     shared rand() stream (reference order: chromosome-major, individual-minor; C:338, C:419) -- the offsets are then
//     computed ON THE DEVICE (k_dist_offsets) and feed the evaluation kernel without a host round trip;
//   * one all-reduce (sum) of the S^2 x G partial curves.
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy torch already loaded in a Python process, the system one
// under R), so the library itself has no link-time dependency on it.  Two topologies:
//   single process, several devices ("devices" option): one communicator per device from ncclCommInitAll;
//   one process per device (torchrun): fgt_dist_unique_id / fgt_dist_init with the id broadcast by the launcher.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace fgt {

typedef struct ncclComm *NcclComm;
struct NcclUniqueId { char internal[128]; };

struct NcclApi {
    bool loaded = false;
    const char *why = "";
    int (*GetUniqueId)(NcclUniqueId *) = nullptr;
    int (*CommInitRank)(NcclComm *, int, NcclUniqueId, int) = nullptr;
    int (*CommInitAll)(NcclComm *, int, const int *) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, NcclComm, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int *) = nullptr;
};
constexpr int kNcclInt64 = 4, kNcclFloat64 = 8, kNcclSum = 0;

NcclApi &nccl();                 // loads on first use; check .loaded

// what a sharded call needs to know about its place among the shards
struct DistCtx {
    int rank = 0, world = 1;
    NcclComm comm = nullptr;
    const uint32_t *h0 = nullptr;    // [31] generator state at call entry, shared by all shards (nullptr: the engine's own)
    bool store_state = true;         // this shard writes the advanced generator state back (one shard per process does)
};

// launch wrappers (kernels.cu)
void launch_dist_offsets(const long long *all /*[world][n_max+1]*/, int world, int rank, int n_max, long long *stream_pos,
                         long long *pair_base /*[n_local]*/, cudaStream_t st);

}  // namespace fgt
