// Per-batch WORK SET of the sharded (multi-GPU) worker: the unique keys of one batch, numbered, with
// their pulled parameters and gradient accumulators in COMPACT arrays (L2-resident for LR), instead of
// rows scattered through a hash table.  It is what LRWorker::update holds between Pull and Push —
// unique_keys / w / push_gradient (lr_worker.cc:147-175) — laid out for the GPU:
//
//   set     open-addressing hash set, 16-byte entries { u64 key ; u32 u ; u32 claimed }
//           (EMPTY key = 2^64-1, claimed = 0xFFFFFFFF until the first token of the key claims it);
//           cleared with one streaming memset per batch
//   u       = owner_shard * cap + position inside that owner's bucket  (bucket-major, so bucket q of
//           every array is one contiguous NCCL send / receive)
//   keys[u] the unique keys (Pull request)      w[u], v[u*K+k]  pulled values (Pull response)
//   gw[u] (f64), acc[2u] = {L, Aq} (f64) gradient accumulators -> grad_w / grad_v (Push payload); the
//           latent gradient is factorised like in the table (table.cuh): gv[u,k] = Aq[u] - v[u,k] * L[u]
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct XfWorkSet {
  uint8_t* set;
  uint64_t mask;
  uint32_t log2cap;
  uint32_t cap;      // stride between owner buckets
  int K;
  uint64_t* keys;
  float* w;
  float* v;
  double* gw;
  double* acc;  // {L, Aq} per key (FM only)
};

struct XfBucketCounts {
  uint32_t c[16];
};

#ifdef __CUDACC__
__device__ __forceinline__ uint64_t xf_ws_hash(uint64_t key, uint32_t log2cap) {
  return (key * 0x9E3779B97F4A7C15ull) >> (64 - log2cap);
}
// key -> u for a key that is known to be in the set
__device__ __forceinline__ uint32_t xf_ws_find(const XfWorkSet& ws, uint64_t key) {
  uint64_t s = xf_ws_hash(key, ws.log2cap);
  for (int probes = 0; probes < 8192; ++probes) {
    // through L1: the set is read-only once the dedup kernel has finished, and hot keys stay SM-local
    const uint4 e = __ldca(reinterpret_cast<const uint4*>(ws.set + s * 16));
    const uint64_t k = (uint64_t)e.x | ((uint64_t)e.y << 32);
    if (k == key) return e.z;
    if (k == 0xFFFFFFFFFFFFFFFFull) return 0xFFFFFFFFu;
    s = (s + 1) & ws.mask;
  }
  return 0xFFFFFFFFu;
}
#endif
