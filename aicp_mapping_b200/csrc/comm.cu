// comm.cu -- multi-GPU single registration: the reading is sharded over the ranks, the reference index is replicated.
//
// BASELINE.json config 4 ("120k-pt reading vs 10M-pt fixed map, reading points sharded over 8 GPUs with NCCL JtJ
// allreduce").  Per ICP iteration the ranks exchange, on the compute stream and without host synchronisation:
//   3 x ncclAllReduce(sum, uint32[2048])   the three radix-select digit histograms -- the trimmed-distance threshold is a
//                                          quantile over ALL reading points, so it cannot be picked per shard
//   1 x ncclAllReduce(sum, uint64[113])    the 27 normal-equation partials + inlier count as 32-bit limbs, + a status word
// Everything exchanged is an integer, so the reduction is exact and every rank solves bit-identical normal equations:
// the sharded result equals the single-GPU result bit for bit, whatever the rank count.
//
// NCCL is dlopen()ed at aicp_b200_comm_init (libnccl.so.2 -- the copy torch already loaded when the caller is a torch
// process), so libaicp_b200.so has no link-time dependency on it and loads on machines without NCCL.
#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <string>

#include "handle.cuh"

namespace aicp {

// the subset of nccl.h that is used (NCCL 2.x ABI)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclSuccess_ = 0 };
enum { ncclSum_ = 0 };
enum { ncclUint32_ = 3, ncclUint64_ = 5 };

struct Comm {
  void* lib = nullptr;
  ncclComm_t comm = nullptr;
  int rank = 0, n_ranks = 1;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  DevBuf<unsigned long long> limbs;       // 4*AICP_NSUM + 1 exchange words, + 1 word for the reading size
  long long n_read_total = 0;
};

static void* open_nccl(std::string* err) {
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    void* lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (lib) return lib;
  }
  *err = dlerror() ? dlerror() : "libnccl.so.2 not found";
  return nullptr;
}

static int nccl_fail(Handle* h, Comm* c, ncclResult_t r, const char* what) {
  return fail(h, AICP_B200_ERR_COMM, "NCCL %s failed: %s", what, c && c->GetErrorString ? c->GetErrorString(r) : "?");
}

int comm_allreduce_u32(Handle* h, unsigned int* buf, size_t count) {
  Comm* c = h->comm;
  ncclResult_t r = c->AllReduce(buf, buf, count, ncclUint32_, ncclSum_, c->comm, h->stream);
  if (r != ncclSuccess_) return nccl_fail(h, c, r, "AllReduce(uint32)");
  h->launches += 1;
  return AICP_B200_OK;
}

int comm_allreduce_u64(Handle* h, unsigned long long* buf, size_t count) {
  Comm* c = h->comm;
  ncclResult_t r = c->AllReduce(buf, buf, count, ncclUint64_, ncclSum_, c->comm, h->stream);
  if (r != ncclSuccess_) return nccl_fail(h, c, r, "AllReduce(uint64)");
  h->launches += 1;
  return AICP_B200_OK;
}

unsigned long long* comm_limbs(Handle* h) { return h->comm->limbs.p; }
long long comm_total_reading(Handle* h) { return h->comm->n_read_total; }

bool comm_peer_view(Handle* h, PeerView* pv) {
  memset(pv, 0, sizeof(*pv));
  pv->n_ranks = 1;
  return false;
}

// total reading size over the ranks (denominator of getWeightedPointUsedRatio)
int comm_begin_registration(Handle* h, long long n_read_local) {
  Comm* c = h->comm;
  unsigned long long* slot = c->limbs.p + 4 * AICP_NSUM + 1;
  unsigned long long v = (unsigned long long)n_read_local;
  CUDA_TRY(cudaMemcpyAsync(slot, &v, sizeof(v), cudaMemcpyHostToDevice, h->stream));
  int rc = comm_allreduce_u64(h, slot, 1);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(&v, slot, sizeof(v), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  c->n_read_total = (long long)v;
  return AICP_B200_OK;
}

}  // namespace aicp

using namespace aicp;

extern "C" {

int aicp_b200_comm_unique_id(uint8_t id_out[128]) {
  if (!id_out) return AICP_B200_ERR_BAD_ARG;
  std::string err;
  void* lib = open_nccl(&err);
  if (!lib) return fail(nullptr, AICP_B200_ERR_COMM, "cannot load NCCL: %s", err.c_str());
  auto get = reinterpret_cast<ncclResult_t (*)(ncclUniqueId*)>(dlsym(lib, "ncclGetUniqueId"));
  if (!get) return fail(nullptr, AICP_B200_ERR_COMM, "ncclGetUniqueId not found in libnccl");
  ncclUniqueId id;
  ncclResult_t r = get(&id);
  if (r != ncclSuccess_) return fail(nullptr, AICP_B200_ERR_COMM, "ncclGetUniqueId failed (%d)", r);
  memcpy(id_out, id.internal, 128);
  return AICP_B200_OK;
}

int aicp_b200_comm_init(aicp_b200_handle* hh, const uint8_t nccl_unique_id[128], int rank, int n_ranks) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || !nccl_unique_id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(h, AICP_B200_ERR_BAD_ARG, "comm_init: bad arguments");
  if (h->comm) return fail(h, AICP_B200_ERR_BAD_ARG, "comm_init: communicator already initialised");
  CUDA_TRY(cudaSetDevice(h->device));
  std::string err;
  Comm* c = new Comm();
  c->lib = open_nccl(&err);
  if (!c->lib) { delete c; return fail(h, AICP_B200_ERR_COMM, "cannot load NCCL: %s", err.c_str()); }
  c->CommInitRank = reinterpret_cast<decltype(c->CommInitRank)>(dlsym(c->lib, "ncclCommInitRank"));
  c->CommDestroy = reinterpret_cast<decltype(c->CommDestroy)>(dlsym(c->lib, "ncclCommDestroy"));
  c->AllReduce = reinterpret_cast<decltype(c->AllReduce)>(dlsym(c->lib, "ncclAllReduce"));
  c->GetErrorString = reinterpret_cast<decltype(c->GetErrorString)>(dlsym(c->lib, "ncclGetErrorString"));
  if (!c->CommInitRank || !c->CommDestroy || !c->AllReduce) { delete c; return fail(h, AICP_B200_ERR_COMM, "libnccl lacks a required symbol"); }
  ncclUniqueId id;
  memcpy(id.internal, nccl_unique_id, 128);
  ncclResult_t r = c->CommInitRank(&c->comm, n_ranks, id, rank);
  if (r != ncclSuccess_) { int rc = nccl_fail(h, c, r, "CommInitRank"); delete c; return rc; }
  c->rank = rank; c->n_ranks = n_ranks;
  if (c->limbs.reserve(4 * AICP_NSUM + 2) != cudaSuccess) { c->CommDestroy(c->comm); delete c; return fail(h, AICP_B200_ERR_CUDA, "comm_init: allocation failed"); }
  h->comm = c;
  return AICP_B200_OK;
}

int aicp_b200_comm_info(aicp_b200_handle* hh, char* buf, int len) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || !buf || len < 1) return AICP_B200_ERR_BAD_ARG;
  if (!h->comm) { snprintf(buf, (size_t)len, "no communicator"); return AICP_B200_OK; }
  snprintf(buf, (size_t)len, "%d ranks; per iteration 3 x ncclAllReduce(uint32[2048]) for the trimmed quantile + 1 x ncclAllReduce(uint64[113]) "
           "for the normal equations, 9 launches", h->comm->n_ranks);
  return AICP_B200_OK;
}

int aicp_b200_comm_destroy(aicp_b200_handle* hh) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || !h->comm) return AICP_B200_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  Comm* c = h->comm;
  if (c->comm) c->CommDestroy(c->comm);
  c->limbs.release();
  delete c;
  h->comm = nullptr;
  return AICP_B200_OK;
}

}
