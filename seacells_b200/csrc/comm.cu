// Transports of the one collective the engine uses (comm.cuh).
#include "comm.cuh"

#include <dlfcn.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <condition_variable>
#include <mutex>
#include <vector>

#include "common.cuh"

int sc_ctx::rank() const { return comm ? comm->rank : 0; }
int sc_ctx::world() const { return comm ? comm->world : 1; }

// ------------------------------------------------------------------------------------------------ threads of one process
struct sc_comm_group {
  int world = 1;
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0;
  unsigned long long generation = 0;
  std::vector<void*> bufs;
  bool broken = false;
  // false when a peer did not arrive within 10 minutes (it failed between two exchanges): the job is abandoned
  bool barrier() {
    std::unique_lock<std::mutex> lk(mu);
    if (broken) return false;
    const unsigned long long gen = generation;
    if (++arrived == world) {
      arrived = 0;
      generation++;
      cv.notify_all();
      return true;
    }
    if (!cv.wait_for(lk, std::chrono::minutes(10), [&] { return generation != gen || broken; }) || broken) {
      broken = true;
      cv.notify_all();
      return false;
    }
    return true;
  }
};

namespace {

struct LocalComm : sc_comm {
  sc_comm_group* g;
  int allgather(void* buf, size_t bytes, cudaStream_t stream) override {
    if (world == 1 || bytes == 0) return SC_OK;
    // my slab is complete once my stream has drained; then everyone publishes its buffer and pulls the other slabs
    if (cudaStreamSynchronize(stream) != cudaSuccess) {
      err = "LocalComm: stream synchronisation failed";
      return SC_ERR_NCCL;
    }
    g->bufs[rank] = buf;
    if (!g->barrier()) {
      err = "LocalComm: a peer rank did not reach the exchange";
      return SC_ERR_NCCL;
    }
    cudaError_t e = cudaSuccess;
    for (int q = 0; q < world && e == cudaSuccess; ++q) {
      if (q == rank) continue;
      e = cudaMemcpyAsync((char*)buf + (size_t)q * bytes, (const char*)g->bufs[q] + (size_t)q * bytes, bytes,
                          cudaMemcpyDefault, stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    const bool all = g->barrier();  // nobody overwrites its slab before every peer has read it
    if (!all) {
      err = "LocalComm: a peer rank did not reach the exchange";
      return SC_ERR_NCCL;
    }
    if (e != cudaSuccess) {
      err = std::string("LocalComm: peer copy failed: ") + cudaGetErrorString(e);
      return SC_ERR_NCCL;
    }
    return SC_OK;
  }
};

}  // namespace

sc_comm_group* sc_local_group_create(int world) {
  if (world < 1) return nullptr;
  sc_comm_group* g = new sc_comm_group();
  g->world = world;
  g->bufs.assign(world, nullptr);
  return g;
}
void sc_local_group_destroy(sc_comm_group* g) { delete g; }
void sc_local_group_break(sc_comm_group* g) {
  if (!g) return;
  std::lock_guard<std::mutex> lk(g->mu);
  g->broken = true;
  g->cv.notify_all();
}
sc_comm* sc_local_comm_create(sc_comm_group* g, int rank) {
  if (!g || rank < 0 || rank >= g->world) return nullptr;
  LocalComm* c = new LocalComm();
  c->g = g;
  c->rank = rank;
  c->world = g->world;
  return c;
}

// ------------------------------------------------------------------------------------------------ NCCL (dlopen)
// Minimal declarations of the NCCL C API (stable since 2.x); the library is resolved at run time.
namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct {
  char internal[128];
} ncclUniqueId;
typedef int ncclResult_t;  // ncclSuccess == 0
constexpr int kNcclInt8 = 0;

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok() const { return handle && GetUniqueId && CommInitRank && CommDestroy && AllGather && GetErrorString; }
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

bool try_open(const char* path) {
  if (!path || !*path) return false;
  void* h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
  if (!h) return false;
  NcclApi a;
  a.handle = h;
  a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  a.CommInitRank = (decltype(a.CommInitRank))dlsym(h, "ncclCommInitRank");
  a.CommDestroy = (decltype(a.CommDestroy))dlsym(h, "ncclCommDestroy");
  a.AllGather = (decltype(a.AllGather))dlsym(h, "ncclAllGather");
  a.GetErrorString = (decltype(a.GetErrorString))dlsym(h, "ncclGetErrorString");
  if (!a.ok()) {
    dlclose(h);
    return false;
  }
  g_nccl = a;
  return true;
}

struct NcclComm : sc_comm {
  ncclComm_t comm = nullptr;
  ~NcclComm() override {
    if (comm) g_nccl.CommDestroy(comm);
  }
  int allgather(void* buf, size_t bytes, cudaStream_t stream) override {
    if (world == 1 || bytes == 0) return SC_OK;
    const ncclResult_t r = g_nccl.AllGather((const char*)buf + (size_t)rank * bytes, buf, bytes, kNcclInt8, comm, stream);
    if (r != 0) {
      err = std::string("ncclAllGather: ") + g_nccl.GetErrorString(r);
      return SC_ERR_NCCL;
    }
    return SC_OK;
  }
};
}  // namespace

int sc_nccl_load(const char* path, std::string* err) {
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  if (g_nccl.ok()) return SC_OK;
  // the copy torch already mapped (same SONAME) is preferred: one NCCL instance per process
  if (try_open(getenv("SC_NCCL_LIB")) || try_open(path) || try_open("libnccl.so.2") || try_open("libnccl.so")) return SC_OK;
  if (err) *err = "libnccl.so.2 not found (set SC_NCCL_LIB or pass its path to sc_nccl_load)";
  return SC_ERR_NCCL;
}

int sc_nccl_get_unique_id(void* out128, std::string* err) {
  SC_TRY(sc_nccl_load(nullptr, err));
  ncclUniqueId id;
  const ncclResult_t r = g_nccl.GetUniqueId(&id);
  if (r != 0) {
    if (err) *err = std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r);
    return SC_ERR_NCCL;
  }
  memcpy(out128, id.internal, 128);
  return SC_OK;
}

sc_comm* sc_nccl_comm_create(int rank, int world, const void* id128, std::string* err) {
  if (sc_nccl_load(nullptr, err) != SC_OK) return nullptr;
  ncclUniqueId id;
  memcpy(id.internal, id128, 128);
  NcclComm* c = new NcclComm();
  c->rank = rank;
  c->world = world;
  const ncclResult_t r = g_nccl.CommInitRank(&c->comm, world, id, rank);
  if (r != 0) {
    if (err) *err = std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r);
    c->comm = nullptr;
    delete c;
    return nullptr;
  }
  return c;
}
