// Sharded (multi-GPU) step: device-side layout shared by mg_kernels.cu and comm.cu.
//
// One process per GPU; every rank is a WORKER (its own CSR batch) and the OWNER of one key range
// (shard = min(key / floor((2^64-1)/S), S-1), ps-lite/src/postoffice.cc:134-143).  Instead of ps-lite's
// KVWorker slicing + ZeroMQ Van (ps-lite/include/ps/kv_app.h:405-460) the ranks exchange through ONE
// cudaMalloc'ed "slab" per rank that every peer maps with cudaIpc: producers store straight into the
// consumer's slab over NVLink from inside their kernels, and order "data is there" with 64-bit step
// counters (flags) written to the consumer after a system-scope fence.  No NCCL call, no host
// synchronisation and no count exchange through the host inside a step.
//
// Per step t (parity p = t & 1), see comm.cu for the stream schedule:
//   worker  xf_k_route        every token's key and row number go to the owner of the key:
//                             in_keys[p][me][.], in_rows[p][me][.] of that owner      (the Pull request)
//   owner   xf_k_pull_tokens  probe/insert each received token's key, answer with w (FM: w, sum_k v,
//                             sum_k v^2) straight into the worker's vals[] array      (the Pull response)
//   worker  xf_k_rows         per-row sums, sigmoid, residual; the per-row residual (FM: and S) is
//                             broadcast to every owner: in_rowv[me][row]              (the Push payload,
//                             factorised: 4 or 8 bytes per ROW instead of 4(1+K) bytes per key)
//   owner   xf_k_push_*       per source rank, in rank order: every token adds its row's residual to its
//                             key's accumulators and the optimizer step is applied once per (source, key)
//                             — the reference's Push of that worker (ftrl.h:54-79, sgd.h:46-52)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define XF_MG_MAX_SHARDS 16

enum { XF_F_KEYS = 0, XF_F_VALS = 1, XF_F_ROWV = 2, XF_F_DONE = 3, XF_NFLAG = 4 };

// byte offsets inside a rank's slab; identical on every rank (same S, cap, max_rows, model)
struct XfSlabLayout {
  uint64_t off_flags;    // u64 [XF_NFLAG][XF_MG_MAX_SHARDS]  step counters written by peers
  uint64_t off_meta;     // u32 [2][XF_MG_MAX_SHARDS][4]      {tokens, rows, -, -} of source s, parity p
  uint64_t off_uniq;     // u64                               unique keys of MY batches, added by the owners
  uint64_t off_in_keys;  // u64 [2][S][cap]                   routed tokens: key
  uint64_t off_in_rows;  // u32 [2][S][cap]                   routed tokens: row number in the source's batch
  uint64_t off_in_rowv;  // f32 (FM: float2) [S][max_rows]    per-row residual (FM: {residual, S}) of source s
  uint64_t off_vals;     // f32 (FM: float4) [S][cap]         Pull responses, bucket-major: owner q writes segment q
  uint64_t total;
};

struct XfPeers {
  uint8_t* slab[XF_MG_MAX_SHARDS];
};

static inline uint64_t xf_align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

static inline XfSlabLayout xf_slab_layout(int S, uint64_t cap, uint64_t max_rows, bool fm) {
  XfSlabLayout L;
  uint64_t o = 0;
  L.off_flags = o; o = xf_align_up(o + (uint64_t)XF_NFLAG * XF_MG_MAX_SHARDS * 8, 4096);
  L.off_meta = o;  o = xf_align_up(o + 2ull * XF_MG_MAX_SHARDS * 4 * 4, 4096);
  L.off_uniq = o;  o = xf_align_up(o + 8, 4096);
  L.off_in_keys = o; o = xf_align_up(o + 2ull * S * cap * 8, 4096);
  L.off_in_rows = o; o = xf_align_up(o + 2ull * S * cap * 4, 4096);
  L.off_in_rowv = o; o = xf_align_up(o + (uint64_t)S * max_rows * (fm ? 8 : 4), 4096);
  L.off_vals = o;  o = xf_align_up(o + (uint64_t)S * cap * (fm ? 16 : 4), 4096);
  L.total = o;
  return L;
}
