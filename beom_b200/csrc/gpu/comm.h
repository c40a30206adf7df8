// comm.h -- NCCL plumbing for the y-slab decomposition (one process per GPU).  NCCL is loaded with
// dlopen at first use so that a single-GPU run needs no NCCL at all and a process that already has
// torch's NCCL loaded shares it.
#ifndef BEOM_COMM_H
#define BEOM_COMM_H
#include <cuda_runtime.h>

#include <string>

namespace beom {
int comm_unique_id(char id[128], std::string *err);
int comm_init(const char id[128], int rank, int nranks, int device, std::string *err);
int comm_finalize();
bool comm_ready();
int comm_rank();
int comm_size();
// exchange `count` doubles with the lower (rank-1) and upper (rank+1) neighbour; peers < 0 are skipped
int comm_exchange(const double *send_lo, double *recv_lo, int peer_lo, const double *send_hi, double *recv_hi, int peer_hi,
                  size_t count, cudaStream_t s, std::string *err);
int comm_allreduce_sum(double *buf, size_t count, cudaStream_t s, std::string *err);
int comm_allreduce_max(double *buf, size_t count, cudaStream_t s, std::string *err);
}  // namespace beom
#endif
