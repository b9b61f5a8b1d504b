// C ABI layer 5: the multi-GPU exchange that replaces ps-lite's KVWorker slicing + ZeroMQ Van
// (ps-lite/include/ps/kv_app.h:405-460, ps-lite/src/zmq_van.h).  One process per GPU; NCCL over
// NVLink / NVSwitch.  The communicator id is created by rank 0 and distributed out of band by the
// launcher (bench.py uses torch.distributed for that and nothing else).
//
// Every rank is at once a WORKER (its own CSR batch) and the SERVER of one key range
// (shard = min(key / floor((2^64-1)/S), S-1), postoffice.cc:134-143).  One step =
//
//   worker  dedup the batch's keys into the per-batch work set (workset.cuh): every unique key gets a
//           number u inside its owner's bucket                            [xf_k_ws_dedup]
//   all     exchange bucket sizes (ncclAllGather), then keys              [all-to-all #1 = the Pull request]
//   owner   probe/insert the received keys in its shard, gather w (,v)    [xf_k_probe, xf_k_gather_v]
//   all     values back, straight into the work set's compact arrays      [all-to-all #2 = the Pull response]
//   worker  the fused step against the work set (same arithmetic as the single-GPU kernel: forward,
//           residual, per-key gradient accumulation), then g / rows       [xf_k_step_ws, xf_k_ws_grads]
//   all     gradients to the owners                                       [all-to-all #3 = the Push]
//   owner   one FTRL/SGD step per (source rank, key), source ranks applied in rank order — one legal
//           schedule of the reference's asynchronous multi-worker run (every worker pulled before any
//           push of the round), reproducible by the oracle                [xf_k_update<.,false> x S]
//
// The three all-to-alls are reads from peer memory (cudaIpc-mapped buffers + DMA copies behind a 4-byte
// NCCL all-reduce, see xf_mg_setup_p2p / xf_mg_pull); grouped ncclSend/ncclRecv on the table's stream is
// the fallback (XFLOW_P2P=0 or no IPC).  NVSwitch gives every pair the same bandwidth, so a flat
// all-to-all is the right schedule; at S = 1 none of this runs.
#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only: the library itself is bound at run time, see XfNccl
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "internal.h"

// NCCL is bound with dlopen at the first xf_comm_* call instead of at link time.  A process that also
// hosts PyTorch already has PyTorch's own (newer) libnccl.so.2 mapped; linking ours against the system
// copy made whichever library loaded second fail on missing symbols.  RTLD_NOLOAD first reuses an
// already-mapped libnccl.so.2, otherwise the system one is loaded (plain C++ hosts).
struct XfNccl {
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  decltype(&ncclCommSplit) CommSplit = nullptr;  // optional (NCCL >= 2.18)
  bool ok = false;
};
static XfNccl g_nccl;

static int xf_nccl_bind() {
  if (g_nccl.ok) return XF_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    xf_set_error("cannot load libnccl.so.2: %s", dlerror());
    return XF_ERR_COMM;
  }
#define XF_BIND(name)                                                         \
  g_nccl.name = reinterpret_cast<decltype(g_nccl.name)>(dlsym(h, "nccl" #name)); \
  if (!g_nccl.name) {                                                         \
    xf_set_error("libnccl lacks nccl" #name);                                 \
    return XF_ERR_COMM;                                                       \
  }
  XF_BIND(GetUniqueId) XF_BIND(CommInitRank) XF_BIND(CommDestroy) XF_BIND(AllReduce) XF_BIND(AllGather)
  XF_BIND(Send) XF_BIND(Recv) XF_BIND(GroupStart) XF_BIND(GroupEnd) XF_BIND(GetErrorString)
#undef XF_BIND
  g_nccl.CommSplit = reinterpret_cast<decltype(g_nccl.CommSplit)>(dlsym(h, "ncclCommSplit"));
  g_nccl.ok = true;
  return XF_OK;
}

#define XF_NCCL_TRY(expr)                                                                        \
  do {                                                                                           \
    ncclResult_t _r = (expr);                                                                    \
    if (_r != ncclSuccess) {                                                                     \
      xf_set_error("NCCL error at %s:%d: %s", __FILE__, __LINE__, g_nccl.GetErrorString(_r));       \
      return XF_ERR_COMM;                                                                        \
    }                                                                                            \
  } while (0)

struct xf_comm {
  ncclComm_t nccl = nullptr;
  ncclComm_t nccl2 = nullptr;  // same ranks, independent resources: bucket-size allgather of the NEXT batch
  int rank = 0, nranks = 1, device = 0;
};

static_assert(sizeof(ncclUniqueId) <= XF_COMM_ID_BYTES, "ncclUniqueId does not fit XF_COMM_ID_BYTES");

XF_DLL int xf_comm_get_id(uint8_t id[XF_COMM_ID_BYTES]) {
  if (!id) return XF_ERR_ARG;
  XF_TRY(xf_nccl_bind());
  ncclUniqueId uid;
  XF_NCCL_TRY(g_nccl.GetUniqueId(&uid));
  memset(id, 0, XF_COMM_ID_BYTES);
  memcpy(id, &uid, sizeof(uid));
  return XF_OK;
}

XF_DLL int xf_comm_create(xf_comm** out, const uint8_t id[XF_COMM_ID_BYTES], int rank, int nranks, int device) {
  if (!out || !id || nranks < 1 || rank < 0 || rank >= nranks) return XF_ERR_ARG;
  XF_TRY(xf_nccl_bind());
  XF_CUDA_TRY(cudaSetDevice(device));
  xf_comm* c = new xf_comm;
  c->rank = rank;
  c->nranks = nranks;
  c->device = device;
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  ncclResult_t r = g_nccl.CommInitRank(&c->nccl, nranks, uid, rank);
  if (r != ncclSuccess) {
    xf_set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
    delete c;
    return XF_ERR_COMM;
  }
  if (g_nccl.CommSplit && nranks > 1) {
    if (g_nccl.CommSplit(c->nccl, 0, rank, &c->nccl2, nullptr) != ncclSuccess) c->nccl2 = nullptr;
  }
  *out = c;
  return XF_OK;
}

XF_DLL int xf_comm_destroy(xf_comm* c) {
  if (!c) return XF_OK;
  if (c->nccl2) g_nccl.CommDestroy(c->nccl2);
  if (c->nccl) g_nccl.CommDestroy(c->nccl);
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
  XF_NCCL_TRY(g_nccl.AllReduce(d, d, 1, ncclInt, ncclSum, c->nccl, 0));
  XF_CUDA_TRY(cudaStreamSynchronize(0));
  cudaFree(d);
  return XF_OK;
}

// -------------------------------------------------------------------------------------------------
// worker-side kernels of the sharded step
// -------------------------------------------------------------------------------------------------
#define XF_MG_MAX_SHARDS 16
#define XF_WS_BLOCK 1024
#define XF_WS_TOK 4  // tokens per thread

__device__ __forceinline__ int xf_dev_shard_of(uint64_t key, uint64_t width, int S) {
  const uint64_t s = key / width;
  return (int)(s < (uint64_t)S ? s : (uint64_t)S - 1);
}

// Insert every token's key into the work set; the token that claims a key first numbers it inside
// its owner's bucket (block-aggregated reservation: one global atomic per block and shard) and lists
// it in ws.keys.  Buckets are contiguous: u = shard * cap + position.
__global__ void __launch_bounds__(XF_WS_BLOCK)
xf_k_ws_dedup(XfWorkSet ws, const uint64_t* __restrict__ keys, uint32_t nnz, uint64_t width, int S,
              uint32_t* __restrict__ bucket_cnt) {
  __shared__ unsigned int s_cnt[XF_MG_MAX_SHARDS];
  __shared__ unsigned int s_base[XF_MG_MAX_SHARDS];
  if (threadIdx.x < XF_MG_MAX_SHARDS) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t j_base = blockIdx.x * (XF_WS_BLOCK * XF_WS_TOK) + threadIdx.x;
  uint64_t my_key[XF_WS_TOK];
  uint64_t my_slot[XF_WS_TOK];
  uint32_t my_local[XF_WS_TOK];  // position inside the block's share of the bucket, or NONE
#pragma unroll
  for (int i = 0; i < XF_WS_TOK; ++i) {
    const uint32_t j = j_base + i * XF_WS_BLOCK;
    my_local[i] = 0xFFFFFFFFu;
    my_key[i] = (j < nnz) ? __ldcs(keys + j) : 0xFFFFFFFFFFFFFFFFull;
  }
#pragma unroll
  for (int i = 0; i < XF_WS_TOK; ++i) {
    const uint64_t key = my_key[i];
    if (key == 0xFFFFFFFFFFFFFFFFull) continue;
    uint64_t s = xf_ws_hash(key, ws.log2cap);
    for (int probes = 0; probes < 8192; ++probes) {
      unsigned long long* kp = reinterpret_cast<unsigned long long*>(ws.set + s * 16);
      unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(kp);
      if (cur == 0xFFFFFFFFFFFFFFFFull) cur = atomicCAS(kp, 0xFFFFFFFFFFFFFFFFull, (unsigned long long)key);
      if (cur == 0xFFFFFFFFFFFFFFFFull || cur == key) break;
      s = (s + 1) & ws.mask;
    }
    // claim: exactly one token per key sees "unclaimed"
    const unsigned int old = atomicExch(reinterpret_cast<unsigned int*>(ws.set + s * 16 + 12), 0u);
    if (old == 0xFFFFFFFFu) {
      my_slot[i] = s;
      my_local[i] = atomicAdd(&s_cnt[xf_dev_shard_of(key, width, S)], 1u);
    }
  }
  __syncthreads();
  if (threadIdx.x < S && s_cnt[threadIdx.x]) s_base[threadIdx.x] = atomicAdd(bucket_cnt + threadIdx.x, s_cnt[threadIdx.x]);
  __syncthreads();
#pragma unroll
  for (int i = 0; i < XF_WS_TOK; ++i) {
    if (my_local[i] == 0xFFFFFFFFu) continue;
    const int q = xf_dev_shard_of(my_key[i], width, S);
    const uint32_t u = (uint32_t)q * ws.cap + s_base[q] + my_local[i];
    ws.keys[u] = my_key[i];
    *reinterpret_cast<uint32_t*>(ws.set + my_slot[i] * 16 + 8) = u;
  }
}

// per unique key: gradient sums / rows -> Push payload; accumulators back to zero (streaming pass)
__global__ void xf_k_ws_grads(XfWorkSet ws, XfBucketCounts cnt, double rows, int want_grads,
                              float* __restrict__ grad_w, float* __restrict__ grad_v) {
  const int q = blockIdx.y;
  const uint32_t n = cnt.c[q];
  const int K = ws.K;
  const uint64_t per = (uint64_t)K + 1;
  const uint64_t total = (uint64_t)n * per;
  for (uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (uint64_t)gridDim.x * blockDim.x) {
    // coordinate-major inside the bucket keeps both arrays coalesced
    if (x < n) {
      const uint64_t u = (uint64_t)q * ws.cap + x;
      if (want_grads) grad_w[u] = xf_div_rows((float)ws.gw[u], rows);  // lr_worker.cc:116-118
      ws.gw[u] = 0.0;
    } else {
      const uint64_t y = x - n;  // in [0, n*K)
      const uint64_t idx = (uint64_t)q * ws.cap * K + y;
      // latent gradient from the factorised sums: gv[u,k] = Aq[u] - v[u,k] * L[u]  (fm_worker.cc:141-142,154-156);
      // the accumulators are cleared by the caller once every coordinate has read them
      const uint64_t ul = y / (uint64_t)K;
      const double2 a = __ldcg(reinterpret_cast<const double2*>(ws.acc) + ((uint64_t)q * ws.cap + ul));
      if (want_grads) grad_v[idx] = xf_div_rows((float)(a.y - (double)ws.v[idx] * a.x), rows);
    }
  }
}

// -------------------------------------------------------------------------------------------------
// host orchestration
// -------------------------------------------------------------------------------------------------
struct XfMg {
  int S = 1, rank = 0;
  uint64_t width = 0;
  // Two work sets: batch b+1 is deduplicated (second stream, second communicator for its bucket-size
  // allgather) while batch b's gradients travel and its owner updates run on the table stream.
  XfWorkSet ws2[2];
  size_t set_bytes = 0;
  XfDevBuf d_set[2], d_keys[2], d_w[2], d_v[2], d_gw[2], d_acc[2], grad_w[2], grad_v[2], bucket_cnt[2], all_counts;
  cudaStream_t st2 = nullptr;
  cudaEvent_t ev_free[2] = {nullptr, nullptr};  // work set no longer read by the table stream
  uint64_t step_no = 0;
  XfDevBuf recv_keys, recv_slots, resp_w, resp_v, rgrad_w, rgrad_v;   // owner side, grouped by source
  // Peer-memory exchange (default when cudaIpc works; XFLOW_P2P=0 keeps the NCCL send/recv groups):
  // every buffer a peer reads is exported once with cudaIpcGetMemHandle and mapped by all ranks; an
  // exchange is then a handful of DMA reads from peer memory (cudaMemcpyAsync) behind a tiny NCCL
  // all-reduce that orders "producer finished" across ranks on the table stream.
  bool p2p = false;
  enum { PEER_KEYS0 = 0, PEER_KEYS1, PEER_GW0, PEER_GW1, PEER_GV0, PEER_GV1, PEER_RESP_W, PEER_RESP_V, PEER_NBUF };
  void* peer[XF_MG_MAX_SHARDS][PEER_NBUF];
  int* d_barrier = nullptr;
  uint32_t* h_counts = nullptr;        // pinned S*S
  std::vector<uint64_t> send_off, recv_off, send_cnt, recv_cnt, own_off, resp_off;
  // XFLOW_MG_TRACE=1: CUDA events at the phase boundaries of every step, averages printed at destroy
  bool trace = false;
  std::vector<cudaEvent_t> tev;
  double tsum[16] = {0};
  uint64_t tsteps = 0;
};
static const char* kMgPhase[] = {"set clear + dedup", "counts allgather+sync", "a2a keys", "owner pull", "a2a values",
                                 "(unused)", "fused step (work set)", "grads", "a2a grads", "owner updates"};
#define XF_MG_TRACE_STEPS 512
#define XF_MG_MARK(i) do { if (mg->trace && mg->tsteps < XF_MG_TRACE_STEPS) cudaEventRecord(mg->tev[mg->tsteps * 11 + (i)], st); } while (0)

// ---- peer-memory exchange ------------------------------------------------------------------------
static int xf_mg_barrier(xf_comm* c, XfMg* mg, cudaStream_t st) {
  XF_NCCL_TRY(g_nccl.AllReduce(mg->d_barrier, mg->d_barrier, 1, ncclInt, ncclSum, c->nccl, st));
  return XF_OK;
}

static void* xf_mg_local_buf(XfMg* mg, int i) {
  switch (i) {
    case XfMg::PEER_KEYS0: return mg->d_keys[0].p;
    case XfMg::PEER_KEYS1: return mg->d_keys[1].p;
    case XfMg::PEER_GW0: return mg->grad_w[0].p;
    case XfMg::PEER_GW1: return mg->grad_w[1].p;
    case XfMg::PEER_GV0: return mg->grad_v[0].p;
    case XfMg::PEER_GV1: return mg->grad_v[1].p;
    case XfMg::PEER_RESP_W: return mg->resp_w.p;
    case XfMg::PEER_RESP_V: return mg->resp_v.p;
  }
  return nullptr;
}

// Export this rank's exchange buffers, import everybody else's.  Any failure on any rank (no IPC in this
// container, no peer access) leaves p2p off everywhere: the decision is agreed through an all-reduce.
static int xf_mg_setup_p2p(xf_trainer* tr, XfMg* mg, size_t tot, size_t K) {
  xf_comm* c = tr->comm;
  const int S = mg->S;
  memset(mg->peer, 0, sizeof(mg->peer));
  const char* e = getenv("XFLOW_P2P");
  const bool want = !(e && *e == '0');
  cudaStream_t st = tr->table->stream;
  XF_CUDA_TRY(cudaMalloc(&mg->d_barrier, sizeof(int)));
  XF_CUDA_TRY(cudaMemsetAsync(mg->d_barrier, 0, sizeof(int), st));
  // peers read the Pull responses in place: fixed size, never reallocated
  XF_TRY(mg->resp_w.ensure(tot * 4 + 4));
  if (K) XF_TRY(mg->resp_v.ensure(tot * 4 * K + 4));
  struct Pack { cudaIpcMemHandle_t h[XfMg::PEER_NBUF]; int ok; int pad[15]; };
  std::vector<Pack> all((size_t)S);
  Pack mine;
  memset(&mine, 0, sizeof(mine));
  mine.ok = want ? 1 : 0;
  for (int i = 0; i < XfMg::PEER_NBUF && mine.ok; ++i) {
    void* p = xf_mg_local_buf(mg, i);
    if (p && cudaIpcGetMemHandle(&mine.h[i], p) != cudaSuccess) { cudaGetLastError(); mine.ok = 0; }
  }
  XfDevBuf d_all;
  XF_TRY(d_all.ensure(sizeof(Pack) * (size_t)S));
  XF_CUDA_TRY(cudaMemcpyAsync((char*)d_all.p + sizeof(Pack) * (size_t)mg->rank, &mine, sizeof(Pack), cudaMemcpyHostToDevice, st));
  XF_NCCL_TRY(g_nccl.AllGather((char*)d_all.p + sizeof(Pack) * (size_t)mg->rank, d_all.p, sizeof(Pack), ncclChar, c->nccl, st));
  XF_CUDA_TRY(cudaMemcpyAsync(all.data(), d_all.p, sizeof(Pack) * (size_t)S, cudaMemcpyDeviceToHost, st));
  XF_CUDA_TRY(cudaStreamSynchronize(st));
  int ok = 1;
  for (int q = 0; q < S; ++q) ok &= all[q].ok;
  if (ok) {
    for (int q = 0; q < S && ok; ++q) {
      if (q == mg->rank) continue;
      for (int i = 0; i < XfMg::PEER_NBUF && ok; ++i) {
        if (!xf_mg_local_buf(mg, i)) continue;  // same set of buffers on every rank
        if (cudaIpcOpenMemHandle(&mg->peer[q][i], all[q].h[i], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
          cudaGetLastError();
          mg->peer[q][i] = nullptr;
          ok = 0;
        }
      }
    }
  }
  // agree on the outcome
  int flag = ok ? 0 : 1;
  XF_CUDA_TRY(cudaMemcpyAsync(mg->d_barrier, &flag, sizeof(int), cudaMemcpyHostToDevice, st));
  XF_TRY(xf_mg_barrier(c, mg, st));
  XF_CUDA_TRY(cudaMemcpyAsync(&flag, mg->d_barrier, sizeof(int), cudaMemcpyDeviceToHost, st));
  XF_CUDA_TRY(cudaStreamSynchronize(st));
  XF_CUDA_TRY(cudaMemsetAsync(mg->d_barrier, 0, sizeof(int), st));
  d_all.release();
  mg->p2p = (flag == 0);
  if (!mg->p2p) {
    for (int q = 0; q < S; ++q)
      for (int i = 0; i < XfMg::PEER_NBUF; ++i)
        if (mg->peer[q][i]) { cudaIpcCloseMemHandle(mg->peer[q][i]); mg->peer[q][i] = nullptr; }
  }
  if (getenv("XFLOW_MG_TRACE")) fprintf(stderr, "[xflow mg] rank %d: peer-memory exchange %s\n", mg->rank, mg->p2p ? "on" : "off (NCCL send/recv)");
  return XF_OK;
}

// One exchange as peer reads: segment q of `dst` (element offset doff[q], cnt[q] keys) is copied from rank
// q's exported buffer `which`, where it starts at element offset soff[q].  The caller has already ordered
// "rank q finished writing" before this point of the stream (bucket-size allgather or xf_mg_barrier).
static int xf_mg_pull(XfMg* mg, int which, const void* local_src, const std::vector<uint64_t>& soff, void* dst,
                      const std::vector<uint64_t>& doff, const std::vector<uint64_t>& cnt, size_t bytes_per_key,
                      cudaStream_t st) {
  for (int q = 0; q < mg->S; ++q) {
    if (!cnt[q]) continue;
    const char* src = (q == mg->rank) ? (const char*)local_src : (const char*)mg->peer[q][which];
    XF_CUDA_TRY(cudaMemcpyAsync((char*)dst + doff[q] * bytes_per_key, src + soff[q] * bytes_per_key, cnt[q] * bytes_per_key,
                                cudaMemcpyDeviceToDevice, st));
  }
  return XF_OK;
}

int xf_mg_create(xf_trainer* tr) {
  xf_comm* c = tr->comm;
  XfMg* mg = new XfMg;
  mg->S = c->nranks;
  mg->rank = c->rank;
  if (mg->S > XF_MG_MAX_SHARDS) {
    xf_set_error("at most %d shards supported", XF_MG_MAX_SHARDS);
    delete mg;
    return XF_ERR_ARG;
  }
  const int S = mg->S;
  mg->width = 0xFFFFFFFFFFFFFFFFull / (uint64_t)S;
  const uint32_t nnz = tr->cfg.max_nnz;
  const size_t K = (size_t)tr->table->view.K;
  uint64_t cap_set = 1024;
  while (cap_set < 2ull * nnz) cap_set <<= 1;  // load factor <= 0.5
  uint32_t lg = 0;
  while ((1ull << lg) < cap_set) ++lg;
  mg->set_bytes = cap_set * 16;
  const size_t tot = (size_t)S * nnz;  // bucket-major arrays, bucket stride = max_nnz
  cudaStream_t st = tr->table->stream;
  XF_CUDA_TRY(cudaStreamCreateWithFlags(&mg->st2, cudaStreamNonBlocking));
  for (int b = 0; b < 2; ++b) {
    XF_TRY(mg->d_set[b].ensure(mg->set_bytes));
    XF_TRY(mg->d_keys[b].ensure(tot * 8));
    XF_TRY(mg->d_w[b].ensure(tot * 4));
    XF_TRY(mg->d_gw[b].ensure(tot * 8));
    XF_TRY(mg->grad_w[b].ensure(tot * 4));
    if (K) {
      XF_TRY(mg->d_v[b].ensure(tot * 4 * K));
      XF_TRY(mg->d_acc[b].ensure(tot * 16));
      XF_TRY(mg->grad_v[b].ensure(tot * 4 * K));
    }
    XF_TRY(mg->bucket_cnt[b].ensure((size_t)S * 4));
    XF_CUDA_TRY(cudaMemsetAsync(mg->d_gw[b].p, 0, tot * 8, st));
    if (K) XF_CUDA_TRY(cudaMemsetAsync(mg->d_acc[b].p, 0, tot * 16, st));
    XF_CUDA_TRY(cudaEventCreateWithFlags(&mg->ev_free[b], cudaEventDisableTiming));
    XfWorkSet& w = mg->ws2[b];
    w.set = mg->d_set[b].as<uint8_t>();
    w.mask = cap_set - 1;
    w.log2cap = lg;
    w.cap = nnz;
    w.K = (int)K;
    w.keys = mg->d_keys[b].as<uint64_t>();
    w.w = mg->d_w[b].as<float>();
    w.v = K ? mg->d_v[b].as<float>() : nullptr;
    w.gw = mg->d_gw[b].as<double>();
    w.acc = K ? mg->d_acc[b].as<double>() : nullptr;
  }
  XF_TRY(mg->all_counts.ensure((size_t)S * S * 4));
  XF_CUDA_TRY(cudaStreamSynchronize(st));
  XF_CUDA_TRY(cudaHostAlloc(&mg->h_counts, (size_t)S * S * 4, cudaHostAllocDefault));
  mg->send_off.resize(S + 1);
  mg->recv_off.resize(S + 1);
  mg->send_cnt.resize(S);
  mg->recv_cnt.resize(S);
  mg->own_off.resize(S);
  mg->resp_off.resize(S);
  const char* tenv = getenv("XFLOW_MG_TRACE");
  mg->trace = tenv && *tenv == '1';
  if (mg->trace) {
    mg->tev.resize(11 * XF_MG_TRACE_STEPS);
    for (auto& e : mg->tev) cudaEventCreate(&e);
  }
  tr->mg = mg;
  int rc = xf_mg_setup_p2p(tr, mg, tot, K);
  if (rc != XF_OK) { xf_mg_destroy(tr); return rc; }
  return XF_OK;
}

void xf_mg_destroy(xf_trainer* tr) {
  XfMg* mg = (XfMg*)tr->mg;
  if (!mg) return;
  if (mg->trace && mg->tsteps) {
    cudaDeviceSynchronize();
    // skip the first steps (population / allocation); events were recorded without any extra sync
    const uint64_t skip = mg->tsteps > 40 ? 30 : 0;
    for (uint64_t t = skip; t < mg->tsteps; ++t)
      for (int i = 0; i < 10; ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, mg->tev[t * 11 + i], mg->tev[t * 11 + i + 1]);
        mg->tsum[i] += ms;
      }
    const double n = (double)(mg->tsteps - skip);
    fprintf(stderr, "[xflow mg trace] rank %d, steps %llu..%llu, mean ms per phase:\n", mg->rank,
            (unsigned long long)skip, (unsigned long long)mg->tsteps);
    double tot = 0;
    for (int i = 0; i < 10; ++i) { fprintf(stderr, "    %-24s %8.4f\n", kMgPhase[i], mg->tsum[i] / n); tot += mg->tsum[i] / n; }
    fprintf(stderr, "    %-24s %8.4f\n", "total", tot);
  }
  cudaStreamSynchronize(mg->st2);
  if (mg->p2p) {
    // unmap the peers' buffers, then make sure every rank has done so before anybody frees its own
    cudaStreamSynchronize(tr->table->stream);
    for (int q = 0; q < mg->S; ++q)
      for (int i = 0; i < XfMg::PEER_NBUF; ++i)
        if (mg->peer[q][i]) cudaIpcCloseMemHandle(mg->peer[q][i]);
    if (xf_mg_barrier(tr->comm, mg, tr->table->stream) == XF_OK) cudaStreamSynchronize(tr->table->stream);
  }
  if (mg->d_barrier) cudaFree(mg->d_barrier);
  for (int b = 0; b < 2; ++b) {
    XfDevBuf* pb[] = {&mg->d_set[b], &mg->d_keys[b], &mg->d_w[b], &mg->d_v[b], &mg->d_gw[b], &mg->d_acc[b],
                      &mg->grad_w[b], &mg->grad_v[b], &mg->bucket_cnt[b]};
    for (XfDevBuf* x : pb) x->release();
    if (mg->ev_free[b]) cudaEventDestroy(mg->ev_free[b]);
  }
  XfDevBuf* bufs[] = {&mg->all_counts, &mg->recv_keys, &mg->recv_slots, &mg->resp_w, &mg->resp_v, &mg->rgrad_w, &mg->rgrad_v};
  for (XfDevBuf* b : bufs) b->release();
  if (mg->st2) cudaStreamDestroy(mg->st2);
  if (mg->h_counts) cudaFreeHost(mg->h_counts);
  for (auto& e : mg->tev) cudaEventDestroy(e);
  delete mg;
  tr->mg = nullptr;
}

// grouped all-to-all: bucket q of `send` (element offset soff[q], scnt[q] keys) goes to rank q, segment q
// of `recv` comes from rank q.  `scale` = items of `elem_bytes` per key.
static int xf_all_to_all(xf_comm* c, const void* send, const std::vector<uint64_t>& soff,
                         const std::vector<uint64_t>& scnt, void* recv, const std::vector<uint64_t>& roff,
                         const std::vector<uint64_t>& rcnt, size_t elem_bytes, size_t scale, cudaStream_t st) {
  XF_NCCL_TRY(g_nccl.GroupStart());
  for (int q = 0; q < c->nranks; ++q) {
    if (scnt[q])
      XF_NCCL_TRY(g_nccl.Send((const char*)send + soff[q] * scale * elem_bytes, scnt[q] * scale * elem_bytes, ncclChar, q,
                              c->nccl, st));
    if (rcnt[q])
      XF_NCCL_TRY(g_nccl.Recv((char*)recv + roff[q] * scale * elem_bytes, rcnt[q] * scale * elem_bytes, ncclChar, q,
                              c->nccl, st));
  }
  XF_NCCL_TRY(g_nccl.GroupEnd());
  return XF_OK;
}

// Where everything of one step's three exchanges lives, from the S x S matrix counts[p*S+q] = number of
// unique keys of worker p's batch owned by q (every rank holds the whole matrix after the allgather).
// Element offsets ("keys"); all bucket-major arrays of a worker have bucket stride `cap`.
//   send_off/send_cnt[q]  my bucket q: what I request from / push to owner q (and where its answers land)
//   recv_off/recv_cnt[q]  where source q's keys / gradients land in my owner-side arrays (grouped by source)
//   own_off[q]            where my share starts inside WORKER q's arrays   (peer reads of keys and gradients)
//   resp_off[q]           where my answers start inside OWNER q's response arrays (peer reads of values)
XF_DLL int xf_exchange_plan(const uint32_t* counts, int S, int rank, uint64_t cap, uint64_t* send_off,
                            uint64_t* send_cnt, uint64_t* recv_off, uint64_t* recv_cnt, uint64_t* own_off,
                            uint64_t* resp_off) {
  if (!counts || S < 1 || rank < 0 || rank >= S || !send_off || !send_cnt || !recv_off || !recv_cnt || !own_off ||
      !resp_off)
    return XF_ERR_ARG;
  uint64_t n_recv = 0;
  for (int q = 0; q < S; ++q) {
    send_cnt[q] = counts[rank * S + q];  // my keys owned by q
    recv_cnt[q] = counts[q * S + rank];  // q's keys owned by me
    send_off[q] = (uint64_t)q * cap;     // bucket-major work-set arrays
    recv_off[q] = n_recv;
    n_recv += recv_cnt[q];
    own_off[q] = (uint64_t)rank * cap;   // bucket `rank` of worker q
    uint64_t before = 0;                 // owner q answers its sources in rank order
    for (int p = 0; p < rank; ++p) before += counts[p * S + q];
    resp_off[q] = before;
  }
  return XF_OK;
}

int xf_mg_step(xf_trainer* tr, const uint32_t* d_row_ptr, const uint64_t* d_keys, const uint8_t* d_labels,
               uint32_t rows, uint32_t nnz, int mode, float* d_abs, cudaEvent_t* pm) {
  XfMg* mg = (XfMg*)tr->mg;
  xf_comm* c = tr->comm;
  xf_table* t = tr->table;
  cudaStream_t st = t->stream;
  const int S = mg->S;
  const size_t K = (size_t)t->view.K;
  const int cur = (int)(mg->step_no++ & 1);
  const XfWorkSet& ws = mg->ws2[cur];
  // With a second communicator the dedup of this batch and its bucket-size allgather run on their own
  // stream, concurrently with whatever the previous step still has queued on the table stream
  // (gradient exchange, owner updates).  Without it everything stays on the table stream.
  const bool overlap = c->nccl2 != nullptr;
  cudaStream_t sd = overlap ? mg->st2 : st;
  ncclComm_t cd = overlap ? c->nccl2 : c->nccl;

  // ---- worker: clear the set (one streaming memset), dedup + number + bucket by owner
  if (mg->trace && mg->tsteps < XF_MG_TRACE_STEPS) cudaEventRecord(mg->tev[mg->tsteps * 11 + 0], sd);
  if (overlap) {
    XF_CUDA_TRY(cudaStreamWaitEvent(sd, mg->ev_free[cur], 0));
    // the batch itself may still be on its way (host path: H2D on the trainer's copy stream)
    if (tr->input_ready) XF_CUDA_TRY(cudaStreamWaitEvent(sd, tr->input_ready, 0));
  }
  tr->input_ready = nullptr;
  XF_CUDA_TRY(cudaMemsetAsync(mg->d_set[cur].p, 0xFF, mg->set_bytes, sd));
  XF_CUDA_TRY(cudaMemsetAsync(mg->bucket_cnt[cur].p, 0, (size_t)S * 4, sd));
  if (nnz) {
    const uint32_t per_block = XF_WS_BLOCK * XF_WS_TOK;
    xf_k_ws_dedup<<<(nnz + per_block - 1) / per_block, XF_WS_BLOCK, 0, sd>>>(ws, d_keys, nnz, mg->width, S,
                                                                              mg->bucket_cnt[cur].as<uint32_t>());
    ++tr->launches;
  }
  // ---- bucket sizes of every rank (the only host sync of the step)
  if (mg->trace && mg->tsteps < XF_MG_TRACE_STEPS) cudaEventRecord(mg->tev[mg->tsteps * 11 + 1], sd);
  XF_NCCL_TRY(g_nccl.AllGather(mg->bucket_cnt[cur].p, mg->all_counts.p, (size_t)S, ncclUint32, cd, sd));
  XF_CUDA_TRY(cudaMemcpyAsync(mg->h_counts, mg->all_counts.p, (size_t)S * S * 4, cudaMemcpyDeviceToHost, sd));
  XF_CUDA_TRY(cudaStreamSynchronize(sd));
  XF_TRY(xf_exchange_plan(mg->h_counts, S, mg->rank, ws.cap, mg->send_off.data(), mg->send_cnt.data(), mg->recv_off.data(),
                          mg->recv_cnt.data(), mg->own_off.data(), mg->resp_off.data()));
  uint64_t n_send = 0, n_recv = 0;
  XfBucketCounts bc;
  memset(&bc, 0, sizeof(bc));
  uint32_t max_bucket = 0;
  for (int q = 0; q < S; ++q) {
    bc.c[q] = (uint32_t)mg->send_cnt[q];
    if (bc.c[q] > max_bucket) max_bucket = bc.c[q];
    n_send += mg->send_cnt[q];
    n_recv += mg->recv_cnt[q];
  }
  XF_TRY(mg->recv_keys.ensure(n_recv * 8 + 8));
  XF_TRY(mg->recv_slots.ensure(n_recv * 4 + 4));
  XF_TRY(mg->resp_w.ensure(n_recv * 4 + 4));
  XF_TRY(mg->rgrad_w.ensure(n_recv * 4 + 4));
  if (K) {
    XF_TRY(mg->resp_v.ensure(n_recv * 4 * K + 4));
    XF_TRY(mg->rgrad_v.ensure(n_recv * 4 * K + 4));
  }

  std::vector<uint64_t>& own_off = mg->own_off;
  std::vector<uint64_t>& resp_off = mg->resp_off;

  // ---- all-to-all #1: keys to their owners (the Pull request, kv_app.h:147-165)
  XF_MG_MARK(2);
  if (mg->p2p) {
    // every rank's dedup finished before its bucket sizes left (allgather above): read the keys in place
    XF_TRY(xf_mg_pull(mg, cur ? XfMg::PEER_KEYS1 : XfMg::PEER_KEYS0, ws.keys, own_off, mg->recv_keys.p, mg->recv_off,
                      mg->recv_cnt, 8, st));
  } else {
    XF_TRY(xf_all_to_all(c, ws.keys, mg->send_off, mg->send_cnt, mg->recv_keys.p, mg->recv_off, mg->recv_cnt, 8, 1, st));
  }
  // ---- owner: Pull handler on this shard (insert-on-pull, ftrl.h:56,114-120)
  XF_MG_MARK(3);
  if (n_recv) {
    XF_TRY(t->ensure_room(n_recv));
    xf_launch_probe(t->view, mg->recv_keys.as<uint64_t>(), n_recv, true, mg->recv_slots.as<uint32_t>(),
                    mg->resp_w.as<float>(), st);
    ++tr->launches;
    if (K) {
      xf_launch_gather_v(t->view, mg->recv_slots.as<uint32_t>(), mg->recv_keys.as<uint64_t>(), n_recv,
                         mg->resp_v.as<float>(), st);
      ++tr->launches;
    }
  }
  // ---- all-to-all #2: values back, straight into the work set (the Pull response)
  XF_MG_MARK(4);
  if (mg->p2p) {
    XF_TRY(xf_mg_barrier(c, mg, st));  // every owner has answered
    XF_TRY(xf_mg_pull(mg, XfMg::PEER_RESP_W, mg->resp_w.p, resp_off, ws.w, mg->send_off, mg->send_cnt, 4, st));
    if (K) XF_TRY(xf_mg_pull(mg, XfMg::PEER_RESP_V, mg->resp_v.p, resp_off, ws.v, mg->send_off, mg->send_cnt, 4 * K, st));
  } else {
    XF_TRY(xf_all_to_all(c, mg->resp_w.p, mg->recv_off, mg->recv_cnt, ws.w, mg->send_off, mg->send_cnt, 4, 1, st));
    if (K) XF_TRY(xf_all_to_all(c, mg->resp_v.p, mg->recv_off, mg->recv_cnt, ws.v, mg->send_off, mg->send_cnt, 4, K, st));
  }

  // ---- worker: forward / residual / gradient accumulation against the work set
  XF_MG_MARK(5);
  XF_MG_MARK(6);
  if (pm) XF_CUDA_TRY(cudaEventRecord(pm[0], st));
  xf_launch_step_ws(ws, d_row_ptr, d_keys, d_labels, (int)rows, mode,
                    (mode == 0 && tr->cfg.keep_loss) ? tr->loss.as<float>() : nullptr,
                    mode == 1 ? tr->pctr.as<float>() : nullptr, d_abs, st);
  if (pm) XF_CUDA_TRY(cudaEventRecord(pm[1], st));
  ++tr->launches;
  XF_MG_MARK(7);
  if (mode != 0) {
    // peers may still be reading this rank's responses: nobody starts the next Pull before all are done
    if (mg->p2p) XF_TRY(xf_mg_barrier(c, mg, st));
    XF_CUDA_TRY(cudaEventRecord(mg->ev_free[cur], st));
    XF_CUDA_TRY(cudaGetLastError());
    return XF_OK;  // forward only: nothing was accumulated
  }
  if (n_send) {
    dim3 grid((unsigned)xf_grid_for((uint64_t)max_bucket * (K + 1), 256, 4), (unsigned)S);
    xf_k_ws_grads<<<grid, 256, 0, st>>>(ws, bc, (double)rows, 1, mg->grad_w[cur].as<float>(),
                                        K ? mg->grad_v[cur].as<float>() : nullptr);
    ++tr->launches;
    // {L, Aq} of the used part of every bucket back to zero (one strided memset)
    if (K) XF_CUDA_TRY(cudaMemset2DAsync(ws.acc, (size_t)ws.cap * 16, 0, (size_t)max_bucket * 16, (size_t)S, st));
  }

  // ---- all-to-all #3: gradients to the owners (the Push, kv_app.h:110-118)
  XF_MG_MARK(8);
  if (mg->p2p) {
    XF_TRY(xf_mg_barrier(c, mg, st));  // every worker's gradients are final
    XF_TRY(xf_mg_pull(mg, cur ? XfMg::PEER_GW1 : XfMg::PEER_GW0, mg->grad_w[cur].p, own_off, mg->rgrad_w.p, mg->recv_off,
                      mg->recv_cnt, 4, st));
    if (K)
      XF_TRY(xf_mg_pull(mg, cur ? XfMg::PEER_GV1 : XfMg::PEER_GV0, mg->grad_v[cur].p, own_off, mg->rgrad_v.p, mg->recv_off,
                        mg->recv_cnt, 4 * K, st));
  } else {
    XF_TRY(xf_all_to_all(c, mg->grad_w[cur].p, mg->send_off, mg->send_cnt, mg->rgrad_w.p, mg->recv_off, mg->recv_cnt, 4, 1, st));
    if (K)
      XF_TRY(xf_all_to_all(c, mg->grad_v[cur].p, mg->send_off, mg->send_cnt, mg->rgrad_v.p, mg->recv_off, mg->recv_cnt, 4, K, st));
  }
  XF_CUDA_TRY(cudaEventRecord(mg->ev_free[cur], st));  // work set `cur` and its gradient buffers are free again
  // ---- owner: Push handler, one optimizer step per (source, key), sources in rank order
  XF_MG_MARK(9);
  if (pm) XF_CUDA_TRY(cudaEventRecord(pm[2], st));
  for (int q = 0; q < S; ++q) {
    if (!mg->recv_cnt[q]) continue;
    const uint64_t off = mg->recv_off[q];
    xf_launch_update_pushed(t->view, mg->recv_slots.as<uint32_t>() + off, mg->recv_cnt[q],
                            mg->rgrad_w.as<float>() + off, K ? mg->rgrad_v.as<float>() + off * K : nullptr, st);
    ++tr->launches;
  }
  if (pm) XF_CUDA_TRY(cudaEventRecord(pm[3], st));
  if (mg->trace && mg->tsteps < XF_MG_TRACE_STEPS) {
    cudaEventRecord(mg->tev[mg->tsteps * 11 + 10], st);
    ++mg->tsteps;
  }
  tr->host_unique += n_send;  // statistics: unique keys of this rank's batch
  XF_CUDA_TRY(cudaGetLastError());
  return XF_OK;
}
