// C ABI layer 5: the multi-GPU exchange that replaces ps-lite's KVWorker slicing + ZeroMQ Van
// (ps-lite/include/ps/kv_app.h:405-460, ps-lite/src/zmq_van.h).  One process per GPU; NCCL over
// NVLink / NVSwitch.  The communicator id is created by rank 0 and distributed out of band by the
// launcher (bench.py uses torch.distributed for that and nothing else).
#include <nccl.h>
#include <stdio.h>
#include <string.h>

#include "internal.h"

#define XF_NCCL_TRY(expr)                                                                        \
  do {                                                                                           \
    ncclResult_t _r = (expr);                                                                    \
    if (_r != ncclSuccess) {                                                                     \
      xf_set_error("NCCL error at %s:%d: %s", __FILE__, __LINE__, ncclGetErrorString(_r));       \
      return XF_ERR_COMM;                                                                        \
    }                                                                                            \
  } while (0)

struct xf_comm {
  ncclComm_t nccl = nullptr;
  int rank = 0, nranks = 1, device = 0;
};

static_assert(sizeof(ncclUniqueId) <= XF_COMM_ID_BYTES, "ncclUniqueId does not fit XF_COMM_ID_BYTES");

XF_DLL int xf_comm_get_id(uint8_t id[XF_COMM_ID_BYTES]) {
  if (!id) return XF_ERR_ARG;
  ncclUniqueId uid;
  XF_NCCL_TRY(ncclGetUniqueId(&uid));
  memset(id, 0, XF_COMM_ID_BYTES);
  memcpy(id, &uid, sizeof(uid));
  return XF_OK;
}

XF_DLL int xf_comm_create(xf_comm** out, const uint8_t id[XF_COMM_ID_BYTES], int rank, int nranks, int device) {
  if (!out || !id || nranks < 1 || rank < 0 || rank >= nranks) return XF_ERR_ARG;
  XF_CUDA_TRY(cudaSetDevice(device));
  xf_comm* c = new xf_comm;
  c->rank = rank;
  c->nranks = nranks;
  c->device = device;
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  ncclResult_t r = ncclCommInitRank(&c->nccl, nranks, uid, rank);
  if (r != ncclSuccess) {
    xf_set_error("ncclCommInitRank failed: %s", ncclGetErrorString(r));
    delete c;
    return XF_ERR_COMM;
  }
  *out = c;
  return XF_OK;
}

XF_DLL int xf_comm_destroy(xf_comm* c) {
  if (!c) return XF_OK;
  if (c->nccl) ncclCommDestroy(c->nccl);
  delete c;
  return XF_OK;
}

int xf_comm_nranks(xf_comm* c) { return c ? c->nranks : 1; }
int xf_comm_rank(xf_comm* c) { return c ? c->rank : 0; }

XF_DLL int xf_comm_barrier(xf_comm* c) {
  if (!c) return XF_ERR_ARG;
  XF_CUDA_TRY(cudaSetDevice(c->device));
  int* d = nullptr;
  XF_CUDA_TRY(cudaMalloc(&d, sizeof(int)));
  XF_CUDA_TRY(cudaMemset(d, 0, sizeof(int)));
  XF_NCCL_TRY(ncclAllReduce(d, d, 1, ncclInt, ncclSum, c->nccl, 0));
  XF_CUDA_TRY(cudaStreamSynchronize(0));
  cudaFree(d);
  return XF_OK;
}

// ---- sharded step: implemented in a later section of this file ----
int xf_mg_create(xf_trainer* tr) {
  (void)tr;
  xf_set_error("multi-GPU step not built yet");
  return XF_ERR_STATE;
}
void xf_mg_destroy(xf_trainer* tr) { (void)tr; }
int xf_mg_step(xf_trainer*, const uint32_t*, const uint64_t*, const uint8_t*, uint32_t, uint32_t, int, float*) {
  xf_set_error("multi-GPU step not built yet");
  return XF_ERR_STATE;
}
