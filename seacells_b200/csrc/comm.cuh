// Collective plumbing of the row-sharded fit (one rank per GPU).
//
// The engine needs exactly one primitive: an in-place all-gather of equal slabs (slab q of `buf` is written by rank q and
// read by everyone).  Every exchanged quantity is produced whole by one rank - compressed A columns, per-cell RSS terms,
// per-archetype (value, index) argmin partials, rows of the Gram matrix - and no floating-point value is ever reduced
// across ranks, which is what makes the sharded fit bitwise identical to the single-GPU fit.
//
// Two transports behind the same interface:
//   NcclComm   ncclAllGather over NVLink / NVSwitch on the library's own stream (one process per GPU; libnccl.so.2 is
//              resolved with dlopen at run time, so single-GPU use has no NCCL dependency)
//   LocalComm  ranks are threads of one process (one engine and stream each, on the same or on different GPUs); slabs move
//              with peer cudaMemcpyAsync between two thread barriers.  This is how the sharded logic is tested on a
//              single GPU, and it is a valid transport for one process driving several GPUs.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include <string>

struct sc_comm {
  int rank = 0, world = 1;
  std::string err;
  virtual ~sc_comm() {}
  // in place: slab q occupies [buf + q * bytes_per_rank, buf + (q + 1) * bytes_per_rank); returns 0 or SC_ERR_NCCL
  virtual int allgather(void* buf, size_t bytes_per_rank, cudaStream_t stream) = 0;
};

struct sc_comm_group;  // shared state of a LocalComm group (one per emulated job)
sc_comm_group* sc_local_group_create(int world);
void sc_local_group_destroy(sc_comm_group* g);
void sc_local_group_break(sc_comm_group* g);  // wake every waiting rank with an error
sc_comm* sc_local_comm_create(sc_comm_group* g, int rank);

int sc_nccl_load(const char* path, std::string* err);  // 0 when libnccl is available
int sc_nccl_get_unique_id(void* out128, std::string* err);
sc_comm* sc_nccl_comm_create(int rank, int world, const void* id128, std::string* err);
