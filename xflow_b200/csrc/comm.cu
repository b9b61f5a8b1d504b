// C ABI layer 5: the multi-GPU exchange that replaces ps-lite's KVWorker slicing + ZeroMQ Van
// (ps-lite/include/ps/kv_app.h:405-460, ps-lite/src/zmq_van.h).  One process per GPU; NCCL over
// NVLink / NVSwitch.  The communicator id is created by rank 0 and distributed out of band by the
// launcher (bench.py uses torch.distributed for that and nothing else).
//
// Every rank is at once a WORKER (its own CSR batch) and the SERVER of one key range
// (shard = min(key / floor((2^64-1)/S), S-1), postoffice.cc:134-143).  One step =
//
//   worker  dedup the batch's keys into a small per-batch cache table (same row layout as the main
//           table), bucket the unique keys by owner                       [xf_k_mg_dedup, _scan, _scatter]
//   all     exchange bucket sizes (ncclAllGather), then keys              [all-to-all #1 = the Pull request]
//   owner   probe/insert the received keys in its shard, gather w (,v)    [xf_k_probe, xf_k_gather_v]
//   all     values back                                                   [all-to-all #2 = the Pull response]
//   worker  load them into the cache rows; run THE SAME fused step kernel as the single-GPU path
//           against the cache table (forward, residual, per-key gradient accumulation)
//           gather g / rows per unique key, clear the cache rows          [xf_k_mg_fill, xf_k_step, xf_k_mg_grads]
//   all     gradients to the owners                                       [all-to-all #3 = the Push]
//   owner   one FTRL/SGD step per (source rank, key), source ranks applied in rank order — one legal
//           schedule of the reference's asynchronous multi-worker run (every worker pulled before any
//           push of the round), reproducible by the oracle                [xf_k_update<.,false> x S]
//
// The three all-to-alls are grouped ncclSend/ncclRecv on the table's stream.  NVSwitch gives every
// pair the same bandwidth, so a flat all-to-all is the right schedule; at S = 1 none of this runs.
#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only: the library itself is bound at run time, see XfNccl
#include <stdio.h>
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
  *out = c;
  return XF_OK;
}

XF_DLL int xf_comm_destroy(xf_comm* c) {
  if (!c) return XF_OK;
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
#define XF_MG_BLOCK 256
#define XF_MG_MAX_SHARDS 16
#define XF_NONE 0xFFFFFFFFu

__device__ __forceinline__ int xf_dev_shard_of(uint64_t key, uint64_t width, int S) {
  const uint64_t s = key / width;
  return (int)(s < (uint64_t)S ? s : (uint64_t)S - 1);
}

// Token j -> cache row of its key (inserted on first sight).  first[j] = cache slot if token j is the
// one that created the entry (so unique keys are exactly the tokens with first[j] != NONE), and the
// block counts its creators per owner shard.  Block b owns tokens [b*chunk, (b+1)*chunk).
__global__ void __launch_bounds__(XF_MG_BLOCK)
xf_k_mg_dedup(XfTableView wc, const uint64_t* __restrict__ keys, uint32_t nnz, uint32_t chunk, uint64_t width, int S,
              uint32_t* __restrict__ first, uint32_t* __restrict__ blk_counts) {
  __shared__ unsigned int s_cnt[XF_MG_MAX_SHARDS];
  if (threadIdx.x < XF_MG_MAX_SHARDS) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t lo = blockIdx.x * chunk;
  const uint32_t hi = min(nnz, lo + chunk);
  for (uint32_t j = lo + threadIdx.x; j < hi; j += blockDim.x) {
    const uint64_t key = keys[j];
    const uint64_t p = xf_slot_hash(key, wc.log2cap);
    XfHead h = xf_load_head(xf_row(wc, p));
    const int64_t s = xf_probe_from<true>(wc, key, p, h);
    uint32_t mine = XF_NONE;
    if (s >= 0) {
      // claim "I list this key": flags bit 1, exactly one token wins per key and batch
      unsigned int* fl = reinterpret_cast<unsigned int*>(xf_row(wc, (uint64_t)s) + XF_OFF_FLAGS);
      const unsigned int old = atomicOr(fl, 2u);
      if ((old & 2u) == 0u) {
        mine = (uint32_t)s;
        atomicAdd(&s_cnt[xf_dev_shard_of(key, width, S)], 1u);
      }
    }
    first[j] = mine;
  }
  __syncthreads();
  if (threadIdx.x < S) blk_counts[blockIdx.x * S + threadIdx.x] = s_cnt[threadIdx.x];
}

// one block: exclusive scan of blk_counts over blocks, per shard; bucket totals and bucket bases
__global__ void xf_k_mg_scan(const uint32_t* __restrict__ blk_counts, uint32_t nblk, int S,
                             uint32_t* __restrict__ blk_offsets, uint32_t* __restrict__ send_counts) {
  __shared__ uint32_t s_tot[XF_MG_MAX_SHARDS];
  const int sh = threadIdx.x;
  if (sh < S) {
    uint32_t run = 0;
    for (uint32_t b = 0; b < nblk; ++b) {
      const uint32_t c = blk_counts[b * S + sh];
      blk_offsets[b * S + sh] = run;
      run += c;
    }
    s_tot[sh] = run;
    send_counts[sh] = run;
  }
  __syncthreads();
  if (sh < S) {
    uint32_t base = 0;
    for (int q = 0; q < sh; ++q) base += s_tot[q];
    for (uint32_t b = 0; b < nblk; ++b) blk_offsets[b * S + sh] += base;
  }
}

// creators write (key, cache slot) into their owner's bucket of the send list
__global__ void __launch_bounds__(XF_MG_BLOCK)
xf_k_mg_scatter(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ first, uint32_t nnz, uint32_t chunk,
                uint64_t width, int S, const uint32_t* __restrict__ blk_offsets, uint64_t* __restrict__ send_keys,
                uint32_t* __restrict__ send_slot) {
  __shared__ unsigned int s_pos[XF_MG_MAX_SHARDS];
  if (threadIdx.x < S) s_pos[threadIdx.x] = blk_offsets[blockIdx.x * S + threadIdx.x];
  __syncthreads();
  const uint32_t lo = blockIdx.x * chunk;
  const uint32_t hi = min(nnz, lo + chunk);
  for (uint32_t j = lo + threadIdx.x; j < hi; j += blockDim.x) {
    const uint32_t s = first[j];
    if (s == XF_NONE) continue;
    const uint64_t key = keys[j];
    const uint32_t pos = atomicAdd(&s_pos[xf_dev_shard_of(key, width, S)], 1u);
    send_keys[pos] = key;
    send_slot[pos] = s;
  }
}

// pulled values -> cache rows (w into the head, v block + V_READY)
__global__ void xf_k_mg_fill(XfTableView wc, const uint32_t* __restrict__ send_slot, uint32_t n,
                             const float* __restrict__ w, const float* __restrict__ v) {
  const int K = wc.K;
  const uint64_t per = (uint64_t)K + 1;
  const uint64_t total = (uint64_t)n * per;
  for (uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; x < total;
       x += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t i = x / per;
    const int c = (int)(x % per);
    uint8_t* rowp = xf_row(wc, send_slot[i]);
    if (c == 0) {
      *reinterpret_cast<float*>(rowp + 8) = w[i];
      *reinterpret_cast<uint32_t*>(rowp + XF_OFF_FLAGS) = 2u | (K > 0 ? XF_FLAG_V_READY : 0u);
    } else {
      xf_row_v(rowp)[c - 1] = v[i * K + (c - 1)];
    }
  }
}

// per unique key: gradient sums / rows -> send arrays (what the worker Pushes), then reset the cache row
__global__ void xf_k_mg_grads(XfTableView wc, const uint32_t* __restrict__ send_slot, uint32_t n, double rows,
                              int want_grads, float* __restrict__ gw, float* __restrict__ gv) {
  const int K = wc.K;
  const uint64_t per = (uint64_t)K + 1;
  const uint64_t total = (uint64_t)n * per;
  for (uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; x < total;
       x += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t i = x / per;
    const int c = (int)(x % per);
    uint8_t* rowp = xf_row(wc, send_slot[i]);
    if (c == 0) {
      if (want_grads) {
        const double g = *reinterpret_cast<const double*>(rowp + 24);
        gw[i] = xf_div_rows((float)g, rows);  // lr_worker.cc:116-118
      }
    } else {
      float* gp = xf_row_gv(rowp, K) + (c - 1);
      if (want_grads) gv[i * K + (c - 1)] = xf_div_rows(*gp, rows);  // fm_worker.cc:154-156
      *gp = 0.f;
    }
  }
}

// second pass of the reset (separate kernel: the head may only be cleared once every coordinate
// thread of xf_k_mg_grads has read it)
__global__ void xf_k_mg_clear(XfTableView wc, const uint32_t* __restrict__ send_slot, uint32_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    XfHead h;
    h.key = XF_EMPTY_KEY; h.flags = 0; h.w = 0.f; h.n = 0.f; h.z = 0.f; h.g = -0.0;
    xf_store_head(xf_row(wc, send_slot[i]), h);
  }
}

// -------------------------------------------------------------------------------------------------
// host orchestration
// -------------------------------------------------------------------------------------------------
struct XfMg {
  int S = 1, rank = 0;
  uint64_t width = 0;
  xf_table* wc = nullptr;              // per-batch worker cache table
  uint32_t chunk = 0, nblk = 0;
  XfDevBuf first, blk_counts, blk_offsets, send_counts, all_counts;
  XfDevBuf send_keys, send_slot, pull_w, pull_v, grad_w, grad_v;      // worker side, bucketed by owner
  XfDevBuf recv_keys, recv_slots, resp_w, resp_v, rgrad_w, rgrad_v;   // owner side, grouped by source
  uint32_t* h_counts = nullptr;        // pinned S*S
  std::vector<uint64_t> send_off, recv_off, send_cnt, recv_cnt;
};

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
  mg->width = 0xFFFFFFFFFFFFFFFFull / (uint64_t)mg->S;
  xf_table_config cfg = tr->table->cfg;
  cfg.optimizer = XF_OPTIMIZER_SGD;  // layout with v and gv blocks only; no optimizer ever runs on the cache
  cfg.v_init = XF_VINIT_ZERO;
  cfg.capacity = 4ull * tr->cfg.max_nnz;  // load factor <= 0.25, never grows
  cfg.shard_index = 0;
  cfg.num_shards = 1;
  int r = xf_table_create(&mg->wc, &cfg);
  if (r != XF_OK) { delete mg; return r; }
  const uint32_t nnz = tr->cfg.max_nnz;
  mg->chunk = 4096;
  mg->nblk = (nnz + mg->chunk - 1) / mg->chunk;
  const size_t K = (size_t)tr->table->view.K;
  const int S = mg->S;
  XF_TRY(mg->first.ensure((size_t)nnz * 4));
  XF_TRY(mg->blk_counts.ensure((size_t)mg->nblk * S * 4));
  XF_TRY(mg->blk_offsets.ensure((size_t)mg->nblk * S * 4));
  XF_TRY(mg->send_counts.ensure((size_t)S * 4));
  XF_TRY(mg->all_counts.ensure((size_t)S * S * 4));
  XF_TRY(mg->send_keys.ensure((size_t)nnz * 8));
  XF_TRY(mg->send_slot.ensure((size_t)nnz * 4));
  XF_TRY(mg->pull_w.ensure((size_t)nnz * 4));
  XF_TRY(mg->grad_w.ensure((size_t)nnz * 4));
  if (K) {
    XF_TRY(mg->pull_v.ensure((size_t)nnz * 4 * K));
    XF_TRY(mg->grad_v.ensure((size_t)nnz * 4 * K));
  }
  XF_CUDA_TRY(cudaHostAlloc(&mg->h_counts, (size_t)S * S * 4, cudaHostAllocDefault));
  mg->send_off.resize(S + 1);
  mg->recv_off.resize(S + 1);
  mg->send_cnt.resize(S);
  mg->recv_cnt.resize(S);
  tr->mg = mg;
  return XF_OK;
}

void xf_mg_destroy(xf_trainer* tr) {
  XfMg* mg = (XfMg*)tr->mg;
  if (!mg) return;
  XfDevBuf* bufs[] = {&mg->first, &mg->blk_counts, &mg->blk_offsets, &mg->send_counts, &mg->all_counts,
                      &mg->send_keys, &mg->send_slot, &mg->pull_w, &mg->pull_v, &mg->grad_w, &mg->grad_v,
                      &mg->recv_keys, &mg->recv_slots, &mg->resp_w, &mg->resp_v, &mg->rgrad_w, &mg->rgrad_v};
  for (XfDevBuf* b : bufs) b->release();
  if (mg->h_counts) cudaFreeHost(mg->h_counts);
  xf_table_destroy(mg->wc);
  delete mg;
  tr->mg = nullptr;
}

int xf_grid_for(uint64_t work_items, int block, int blocks_per_sm);

// grouped all-to-all of `elem_bytes`-sized items: bucket q of `send` (send_off/cnt) goes to rank q,
// segment q of `recv` (recv_off/cnt) comes from rank q.  scale = items per key (1 or K).
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

int xf_mg_step(xf_trainer* tr, const uint32_t* d_row_ptr, const uint64_t* d_keys, const uint8_t* d_labels,
               uint32_t rows, uint32_t nnz, int mode, float* d_abs) {
  XfMg* mg = (XfMg*)tr->mg;
  xf_comm* c = tr->comm;
  xf_table* t = tr->table;
  cudaStream_t st = t->stream;
  const int S = mg->S;
  const size_t K = (size_t)t->view.K;
  const XfTableView wc = mg->wc->view;
  const uint32_t nblk = (nnz + mg->chunk - 1) / mg->chunk;

  // ---- worker: dedup + bucket by owner
  if (nnz) {
    xf_k_mg_dedup<<<nblk, XF_MG_BLOCK, 0, st>>>(wc, d_keys, nnz, mg->chunk, mg->width, S, mg->first.as<uint32_t>(),
                                                 mg->blk_counts.as<uint32_t>());
    xf_k_mg_scan<<<1, 32, 0, st>>>(mg->blk_counts.as<uint32_t>(), nblk, S, mg->blk_offsets.as<uint32_t>(),
                                   mg->send_counts.as<uint32_t>());
    xf_k_mg_scatter<<<nblk, XF_MG_BLOCK, 0, st>>>(d_keys, mg->first.as<uint32_t>(), nnz, mg->chunk, mg->width, S,
                                                   mg->blk_offsets.as<uint32_t>(), mg->send_keys.as<uint64_t>(),
                                                   mg->send_slot.as<uint32_t>());
    tr->launches += 3;
  } else {
    XF_CUDA_TRY(cudaMemsetAsync(mg->send_counts.p, 0, (size_t)S * 4, st));
  }
  // ---- bucket sizes of every rank (the only host sync of the step)
  XF_NCCL_TRY(g_nccl.AllGather(mg->send_counts.p, mg->all_counts.p, (size_t)S, ncclUint32, c->nccl, st));
  XF_CUDA_TRY(cudaMemcpyAsync(mg->h_counts, mg->all_counts.p, (size_t)S * S * 4, cudaMemcpyDeviceToHost, st));
  XF_CUDA_TRY(cudaStreamSynchronize(st));
  uint64_t n_send = 0, n_recv = 0;
  for (int q = 0; q < S; ++q) {
    mg->send_cnt[q] = mg->h_counts[mg->rank * S + q];   // my keys owned by q
    mg->recv_cnt[q] = mg->h_counts[q * S + mg->rank];   // q's keys owned by me
    mg->send_off[q] = n_send;
    mg->recv_off[q] = n_recv;
    n_send += mg->send_cnt[q];
    n_recv += mg->recv_cnt[q];
  }
  mg->send_off[S] = n_send;
  mg->recv_off[S] = n_recv;
  XF_TRY(mg->recv_keys.ensure(n_recv * 8 + 8));
  XF_TRY(mg->recv_slots.ensure(n_recv * 4 + 4));
  XF_TRY(mg->resp_w.ensure(n_recv * 4 + 4));
  XF_TRY(mg->rgrad_w.ensure(n_recv * 4 + 4));
  if (K) {
    XF_TRY(mg->resp_v.ensure(n_recv * 4 * K + 4));
    XF_TRY(mg->rgrad_v.ensure(n_recv * 4 * K + 4));
  }

  // ---- all-to-all #1: keys to their owners (the Pull request, kv_app.h:147-165)
  XF_TRY(xf_all_to_all(c, mg->send_keys.p, mg->send_off, mg->send_cnt, mg->recv_keys.p, mg->recv_off, mg->recv_cnt,
                       8, 1, st));
  // ---- owner: Pull handler on this shard (insert-on-pull, ftrl.h:56,114-120)
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
  // ---- all-to-all #2: values back (the Pull response)
  XF_TRY(xf_all_to_all(c, mg->resp_w.p, mg->recv_off, mg->recv_cnt, mg->pull_w.p, mg->send_off, mg->send_cnt, 4, 1, st));
  if (K)
    XF_TRY(xf_all_to_all(c, mg->resp_v.p, mg->recv_off, mg->recv_cnt, mg->pull_v.p, mg->send_off, mg->send_cnt, 4, K, st));

  // ---- worker: forward / residual / gradient accumulation on the cache table, same kernel as S = 1
  if (n_send) {
    xf_k_mg_fill<<<xf_grid_for(n_send * (K + 1), 256, 8), 256, 0, st>>>(wc, mg->send_slot.as<uint32_t>(),
                                                                         (uint32_t)n_send, mg->pull_w.as<float>(),
                                                                         K ? mg->pull_v.as<float>() : nullptr);
    ++tr->launches;
  }
  xf_launch_step(wc, d_row_ptr, d_keys, d_labels, (int)rows, mode, tr->touched.as<uint32_t>(),
                 (mode == 0 && tr->cfg.keep_loss) ? tr->loss.as<float>() : nullptr,
                 mode == 1 ? tr->pctr.as<float>() : nullptr, d_abs, st);
  ++tr->launches;
  if (n_send) {
    xf_k_mg_grads<<<xf_grid_for(n_send * (K + 1), 256, 8), 256, 0, st>>>(
        wc, mg->send_slot.as<uint32_t>(), (uint32_t)n_send, (double)rows, mode == 0 ? 1 : 0, mg->grad_w.as<float>(),
        K ? mg->grad_v.as<float>() : nullptr);
    xf_k_mg_clear<<<xf_grid_for(n_send, 256, 8), 256, 0, st>>>(wc, mg->send_slot.as<uint32_t>(), (uint32_t)n_send);
    tr->launches += 2;
  }
  if (mode != 0) {
    XF_CUDA_TRY(cudaGetLastError());
    return XF_OK;
  }

  // ---- all-to-all #3: gradients to the owners (the Push, kv_app.h:110-118)
  XF_TRY(xf_all_to_all(c, mg->grad_w.p, mg->send_off, mg->send_cnt, mg->rgrad_w.p, mg->recv_off, mg->recv_cnt, 4, 1, st));
  if (K)
    XF_TRY(xf_all_to_all(c, mg->grad_v.p, mg->send_off, mg->send_cnt, mg->rgrad_v.p, mg->recv_off, mg->recv_cnt, 4, K, st));
  // ---- owner: Push handler, one optimizer step per (source, key), sources in rank order
  for (int q = 0; q < S; ++q) {
    if (!mg->recv_cnt[q]) continue;
    const uint64_t off = mg->recv_off[q];
    xf_launch_update_pushed(t->view, mg->recv_slots.as<uint32_t>() + off, mg->recv_cnt[q],
                            mg->rgrad_w.as<float>() + off, K ? mg->rgrad_v.as<float>() + off * K : nullptr, st);
    ++tr->launches;
  }
  tr->host_unique += n_send;  // statistics: unique keys of this rank's batch
  XF_CUDA_TRY(cudaGetLastError());
  return XF_OK;
}
