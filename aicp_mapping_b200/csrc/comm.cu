// comm.cu -- multi-GPU single registration: the reading is sharded over the ranks, the reference index is replicated.
//
// BASELINE.json config 4 ("120k-pt reading vs 10M-pt fixed map, reading points sharded over 8 GPUs with NCCL JtJ
// allreduce").  Per ICP iteration the ranks exchange, on the compute stream and without host synchronisation:
// Two carriers of the same exchange:
//   PEER MEMORY (default).  Every rank owns an inbox in its HBM that all peers map (CUDA IPC between processes, peer access
//   inside one; layout in common.cuh).  The persistent loop kernel (icp.cu, k_icp_loop) stores its digit-1 histogram, the
//   candidate keys of the picked bin and the 28 partial sums straight into every peer's inbox over NVLink and spins on
//   sequence-stamped flags: three exchanges per iteration INSIDE one kernel, no launch and no host involvement.
//   NCCL (fallback when the inboxes cannot be mapped, or AICP_B200_COMM=nccl).  Per iteration, on the compute stream:
//   3 x ncclAllReduce(sum, uint32[2048])   the three radix-select digit histograms -- the trimmed-distance threshold is a
//                                          quantile over ALL reading points, so it cannot be picked per shard
//   1 x ncclAllReduce(sum, uint64[113])    the 27 normal-equation partials + inlier count as 32-bit limbs, + a status word
// Everything exchanged is an integer, so the reduction is exact and every rank solves bit-identical normal equations:
// the sharded result equals the single-GPU result bit for bit, whatever the rank count and whichever carrier.
//
// NCCL is dlopen()ed at aicp_b200_comm_init (libnccl.so.2 -- the copy torch already loaded when the caller is a torch
// process), so libaicp_b200.so has no link-time dependency on it and loads on machines without NCCL.
#include <dlfcn.h>
#include <unistd.h>

#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>

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
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  DevBuf<unsigned long long> limbs;       // 4*AICP_NSUM + 1 exchange words, + 1 word for the reading size
  long long n_read_total = 0;
  // peer-memory carrier
  bool peer = false;                      // the inboxes are mapped
  bool peer_now = false;                  // ... and the registration in progress uses them
  unsigned char* inbox = nullptr;         // this rank's inbox (cudaMalloc)
  size_t cand_stride = 0;
  unsigned int cand_cap = 0;              // candidate keys per source
  unsigned char* peer_inbox[AICP_MAX_RANKS] = {};   // every rank's inbox as mapped here; [rank] == inbox
  bool peer_ipc[AICP_MAX_RANKS] = {};     // mapping came from cudaIpcOpenMemHandle (must be closed)
  unsigned long long epoch = 0;           // registrations started on this communicator (the same on every rank)
};

// what the ranks tell each other about their inboxes
struct InboxCard {
  cudaIpcMemHandle_t handle;
  unsigned long long ptr;
  int pid, device;
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

// Setup of a sharded registration: the reference is replicated, so its SurfaceNormal filter (the most expensive setup stage)
// is split -- rank r computes the normals of Morton positions [r * per, (r + 1) * per) and the slices are all-gathered
int comm_ranks(Handle* h) { return h->comm ? h->comm->n_ranks : 1; }

int comm_slice(Handle* h, int n, int* q0, int* q1, int* per) {
  Comm* c = h->comm;
  const int p = ((n + c->n_ranks - 1) / c->n_ranks + 31) / 32 * 32;
  *per = p;
  *q0 = c->rank * p < n ? c->rank * p : n;
  *q1 = (c->rank + 1) * p < n ? (c->rank + 1) * p : n;
  return AICP_B200_OK;
}

int comm_allgather_bytes(Handle* h, void* buf, size_t bytes_per_rank) {
  Comm* c = h->comm;
  if (!c->AllGather) return fail(h, AICP_B200_ERR_COMM, "libnccl lacks ncclAllGather");
  ncclResult_t r = c->AllGather(static_cast<unsigned char*>(buf) + (size_t)c->rank * bytes_per_rank, buf, bytes_per_rank, 0 /* ncclInt8 */, c->comm, h->stream);
  if (r != ncclSuccess_) return nccl_fail(h, c, r, "AllGather");
  h->launches += 1;
  return AICP_B200_OK;
}

unsigned long long* comm_limbs(Handle* h) { return h->comm->limbs.p; }
long long comm_total_reading(Handle* h) { return h->comm->n_read_total; }

bool comm_peer_view(Handle* h, PeerView* pv) {
  memset(pv, 0, sizeof(*pv));
  pv->n_ranks = 1;
  Comm* c = h->comm;
  if (!c || !c->peer_now) return false;
  pv->rank = c->rank; pv->n_ranks = c->n_ranks; pv->epoch = c->epoch;
  for (int r = 0; r < c->n_ranks; ++r) pv->inbox[r] = c->peer_inbox[r];
  pv->cand_stride = c->cand_stride; pv->cand_cap = c->cand_cap;
  return true;
}

static void peer_teardown(Comm* c) {
  for (int r = 0; r < AICP_MAX_RANKS; ++r) {
    if (c->peer_ipc[r] && c->peer_inbox[r]) cudaIpcCloseMemHandle(c->peer_inbox[r]);
    c->peer_inbox[r] = nullptr; c->peer_ipc[r] = false;
  }
  if (c->inbox) cudaFree(c->inbox);
  c->inbox = nullptr; c->peer = false;
}

// Allocate this rank's inbox, exchange the inbox cards over NCCL and map every peer's inbox.  Collective: every rank
// ends with the same c->peer (a rank that cannot map a peer makes all ranks fall back to the NCCL carrier).
static int peer_setup(Handle* h, Comm* c, unsigned int cand_cap) {
  const int G = c->n_ranks;
  if (G > AICP_MAX_RANKS || !c->AllGather) return AICP_B200_OK;
  cudaStream_t s = h->stream;
  c->cand_cap = cand_cap;
  c->cand_stride = (8 + (size_t)cand_cap * 8 + 255) / 256 * 256;
  const size_t bytes = AICP_INBOX_CAND_OFF + (size_t)G * c->cand_stride;
  int ok = 1;
  InboxCard mine;
  memset(&mine, 0, sizeof(mine));
  if (cudaMalloc((void**)&c->inbox, bytes) != cudaSuccess || cudaMemsetAsync(c->inbox, 0, bytes, s) != cudaSuccess ||
      cudaIpcGetMemHandle(&mine.handle, c->inbox) != cudaSuccess) { cudaGetLastError(); ok = 0; }
  mine.ptr = (unsigned long long)(uintptr_t)c->inbox; mine.pid = (int)getpid(); mine.device = h->device;
  // cards of all ranks (device staging: NCCL moves device memory)
  unsigned char* stage = nullptr;
  CUDA_TRY(cudaMalloc((void**)&stage, sizeof(InboxCard) * (size_t)(G + 1) + 16));
  CUDA_TRY(cudaMemcpyAsync(stage, &mine, sizeof(mine), cudaMemcpyHostToDevice, s));
  ncclResult_t r = c->AllGather(stage, stage + sizeof(InboxCard), sizeof(InboxCard), 0 /* ncclInt8 */, c->comm, s);
  if (r != ncclSuccess_) { cudaFree(stage); return nccl_fail(h, c, r, "AllGather(inbox cards)"); }
  std::vector<InboxCard> cards((size_t)G);
  CUDA_TRY(cudaMemcpyAsync(cards.data(), stage + sizeof(InboxCard), sizeof(InboxCard) * (size_t)G, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  for (int p = 0; p < G && ok; ++p) {
    if (p == c->rank) { c->peer_inbox[p] = c->inbox; continue; }
    if (cards[(size_t)p].ptr == 0) { ok = 0; break; }
    if (cards[(size_t)p].pid == mine.pid) {
      // same process, another device: plain peer access to the other rank's allocation
      int can = 0;
      if (cards[(size_t)p].device != h->device) {
        cudaDeviceCanAccessPeer(&can, h->device, cards[(size_t)p].device);
        if (can) { cudaError_t e = cudaDeviceEnablePeerAccess(cards[(size_t)p].device, 0); if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) can = 0; cudaGetLastError(); }
      }
      if (!can) { ok = 0; break; }
      c->peer_inbox[p] = reinterpret_cast<unsigned char*>((uintptr_t)cards[(size_t)p].ptr);
    } else {
      void* mapped = nullptr;
      if (cudaIpcOpenMemHandle(&mapped, cards[(size_t)p].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
      c->peer_inbox[p] = reinterpret_cast<unsigned char*>(mapped); c->peer_ipc[p] = true;
    }
  }
  // unanimous or not at all
  unsigned int* flag = reinterpret_cast<unsigned int*>(stage);
  unsigned int v = ok ? 1u : 0u;
  CUDA_TRY(cudaMemcpyAsync(flag, &v, sizeof(v), cudaMemcpyHostToDevice, s));
  r = c->AllReduce(flag, flag, 1, ncclUint32_, ncclSum_, c->comm, s);
  if (r != ncclSuccess_) { cudaFree(stage); return nccl_fail(h, c, r, "AllReduce(inbox mapping)"); }
  CUDA_TRY(cudaMemcpyAsync(&v, flag, sizeof(v), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  cudaFree(stage);
  if ((int)v == G) c->peer = true;
  else peer_teardown(c);
  return AICP_B200_OK;
}

// total reading size over the ranks (denominator of getWeightedPointUsedRatio)
int comm_begin_registration(Handle* h, long long n_read_local, bool want_peer) {
  Comm* c = h->comm;
  // peer carrier: nothing to exchange up front -- the reading size travels with the sums inside the loop kernel
  c->peer_now = c->peer && want_peer;
  if (c->peer_now) {
    ++c->epoch; c->n_read_total = 0;
    if ((c->epoch & 0xFFFFFull) == 0) {
      // the 32-bit stamps (2048 per registration) come round after 2^21 registrations: long before that every rank clears
      // its inbox, and nobody proceeds until all have (the all-reduce is the barrier) -- no stale word can ever match
      const size_t bytes = AICP_INBOX_CAND_OFF + (size_t)c->n_ranks * c->cand_stride;
      CUDA_TRY(cudaMemsetAsync(c->inbox, 0, bytes, h->stream));
      unsigned int* w = reinterpret_cast<unsigned int*>(c->limbs.p);
      ncclResult_t r = c->AllReduce(w, w, 1, ncclUint32_, ncclSum_, c->comm, h->stream);
      if (r != ncclSuccess_) return nccl_fail(h, c, r, "AllReduce(inbox reset)");
      CUDA_TRY(cudaStreamSynchronize(h->stream));
    }
    return AICP_B200_OK;
  }
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
  c->AllGather = reinterpret_cast<decltype(c->AllGather)>(dlsym(c->lib, "ncclAllGather"));
  c->GetErrorString = reinterpret_cast<decltype(c->GetErrorString)>(dlsym(c->lib, "ncclGetErrorString"));
  if (!c->CommInitRank || !c->CommDestroy || !c->AllReduce) { delete c; return fail(h, AICP_B200_ERR_COMM, "libnccl lacks a required symbol"); }
  ncclUniqueId id;
  memcpy(id.internal, nccl_unique_id, 128);
  ncclResult_t r = c->CommInitRank(&c->comm, n_ranks, id, rank);
  if (r != ncclSuccess_) { int rc = nccl_fail(h, c, r, "CommInitRank"); delete c; return rc; }
  c->rank = rank; c->n_ranks = n_ranks;
  if (c->limbs.reserve(4 * AICP_NSUM + 2) != cudaSuccess) { c->CommDestroy(c->comm); delete c; return fail(h, AICP_B200_ERR_CUDA, "comm_init: allocation failed"); }
  h->comm = c;
  const char* mode = getenv("AICP_B200_COMM");
  if (!(mode && strcmp(mode, "nccl") == 0)) {
    unsigned int cap = 1u << 20;                       // candidate keys per source rank (4 MiB each)
    if (const char* e = getenv("AICP_B200_COMM_CAND_CAP")) { long v = atol(e); if (v >= 1024 && v <= (1l << 28)) cap = (unsigned int)v; }
    int rc = peer_setup(h, c, cap);
    if (rc) { aicp_b200_comm_destroy(hh); return rc; }
  }
  return AICP_B200_OK;
}

int aicp_b200_comm_info(aicp_b200_handle* hh, char* buf, int len) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || !buf || len < 1) return AICP_B200_ERR_BAD_ARG;
  if (!h->comm) { snprintf(buf, (size_t)len, "no communicator"); return AICP_B200_OK; }
  if (h->comm->peer) {
    snprintf(buf, (size_t)len, "%d ranks; peer-mapped inboxes over NVLink (CUDA IPC): per iteration 3 exchanges inside the "
             "persistent loop kernel, flag-in-data stores (digit-1 histogram, candidate keys of the picked bin, 28 x 128-bit sums), 0 NCCL "
             "calls, 1 launch for the whole loop", h->comm->n_ranks);
    return AICP_B200_OK;
  }
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
  peer_teardown(c);
  if (c->comm) c->CommDestroy(c->comm);
  c->limbs.release();
  delete c;
  h->comm = nullptr;
  return AICP_B200_OK;
}

}
