// Device-side layout and primitives of the GPU-resident parameter table.
//
// The table replaces the reference's server-side state — the two
// std::unordered_map<ps::Key, FTRLEntry/SGDEntry> stores behind KV apps 0 (w) and 1 (v)
// (src/optimizer/ftrl.h:27-36,84,87-96,151 ; src/optimizer/sgd.h:23-28,61,67-72,105) —
// with ONE open-addressing hash table in HBM whose row holds everything a key owns:
//
//   byte  0  u64  key            (EMPTY = 2^64-1; never a legal key in the reference either:
//                                 it lies outside every server range, postoffice.cc:134-143)
//   byte  8  f64  g              per-batch gradient accumulator of w; -0.0 == "untouched this batch".
//                                Double, so that the sum over a key's occurrences is exact to float
//                                precision whatever order the L2 atomics land in (deterministic, and
//                                more accurate than the reference's own sequential float sum).
//                                LAZY tables (LR, "update on next touch"): the same 8 bytes hold the sum as a
//                                64-bit FIXED-POINT integer (scale 2^40): integer adds are exactly associative
//                                (deterministic) and exactly invertible, which the open protocol needs.
//   byte 16  f32  w              app-0 weight                         (FTRLEntry_w::w / SGDEntry_w::w)
//   byte 20  f32  n              FTRL accumulator of w (unused by SGD)
//   byte 24  f32  z              FTRL accumulator of w (unused by SGD)
//   byte 28  u32  flags          bit0 = latent block materialised (V_READY); lazy tables: the batch TAG.
//                                {w, n, z, flags} are one aligned 16-byte word: lazy tables claim AND
//                                publish a row with a single 128-bit compare-and-swap on it (step_lazy.cu)
//   ---- 32 B = one DRAM sector: an LR pull, gradient accumulate or update touches exactly one ----
//   byte 32            f32 v[K]    app-1 latent row
//   byte A             f64 L, f64 Aq   per-batch latent-gradient accumulators, A = round_up(32 + 4K, 16).
//                                  The reference's gv[i,k] = sum_occ loss_s * (S_s - v[i,k])
//                                  (fm_worker.cc:141-142) factorises as Aq_i - v[i,k] * L_i with
//                                  L_i = sum_occ loss_s and Aq_i = sum_occ loss_s * S_s: two f64
//                                  accumulators per key instead of K float ones (3 atomics per token
//                                  instead of 1 + K/4 vector ones, and no second read of v).
//   byte A + 16        f32 nv[K]   (FTRL only)
//   byte A + 16 + 4K   f32 zv[K]   (FTRL only)
//   row stride = round_up(A + 16 + (FTRL ? 8K : 0), 32): K = 16 FTRL -> 256 B.
//
// Missing keys are inserted on first touch by a pull OR a push, like `store[key]`
// (ftrl.h:56,114-120 ; sgd.h:48,92).  Default contents: w = n = z = 0 ; v per init mode.  The latent
// block is materialised lazily by the first update of the key (a row is only ever written by
// kernels that own the key exclusively); until then readers compute the same deterministic
// initial value from (key, k, seed), so "insert on pull" is observably identical.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define XF_EMPTY_KEY 0xFFFFFFFFFFFFFFFFull
#define XF_FLAG_V_READY 1u
#define XF_NEG_ZERO_BITS 0x80000000u
#define XF_NEG_ZERO_BITS64 0x8000000000000000ull
#define XF_MAX_PROBE 8192

enum { XF_OPT_FTRL = 0, XF_OPT_SGD = 1 };
enum { XF_INIT_DEFAULT = 0, XF_INIT_COUNTER = 1, XF_INIT_ZERO = 3 };

struct XfTableView {
  uint8_t* base;
  uint64_t mask;      // capacity - 1
  uint32_t log2cap;
  uint32_t bshift;    // log2 of the slots per probing bucket (see xf_probe_slot)
  uint32_t stride;    // bytes per row
  int K;
  int opt;
  int v_init;         // resolved: 0 const(v_const), 1 counter normal, 3 zero
  float v_const;
  uint64_t seed;
  float alpha, beta, lambda1, lambda2, learning_rate;
  unsigned long long* size;   // number of keys present
  int* error;                 // set to 1 when a probe sequence overflows (table full)
  // "update on next touch" (K == 0 tables, see step.cu): the flags word of a row is then a TAG =
  // sequence number of the batch whose residual sum sits in g (0: nothing pending, XF_TAG_LOCKED: a
  // thread is folding the pending step into the row right now); rows_by_seq[tag] is that batch's row
  // count (the divisor of lr_worker.cc:116-118).  Every reader applies a pending step on the fly.
  int lazy;
  const uint32_t* rows_by_seq;
};
#define XF_TAG_LOCKED 0xFFFFFFFFu  // never a batch number (the sequence ring is far smaller)
#define XF_FIX_SCALE 1099511627776.0        // 2^40: residual sums of lazy tables are integers of this unit
#define XF_FIX_INV 9.094947017729282379e-13  // 2^-40

__host__ __device__ inline uint32_t xf_acc_off(int K) { return (32u + 4u * (uint32_t)K + 15u) & ~15u; }
__host__ __device__ inline uint32_t xf_row_stride(int K, int opt) {
  if (K <= 0) return 32u;
  uint32_t bytes = xf_acc_off(K) + 16u + ((opt == XF_OPT_FTRL) ? 8u * (uint32_t)K : 0u);
  return (bytes + 31u) & ~31u;
}

#ifdef __CUDACC__

// Probe sequence: BUCKETISED linear probing.  Slots are grouped in aligned buckets of B = 2^bshift slots
// that share one 128-byte line (LR: 4 rows of 32 B; FM rows are longer than a line: B = 1 = plain linear
// probing).  Probe i of a key visits its home bucket first — starting at a key-dependent slot and wrapping
// inside the bucket — and then the following buckets slot by slot.  Why: a warp waits for the slowest of its
// lanes, and every step of a linear-probing chain used to be one more DEPENDENT DRAM access (measured on the
// 1e8-id table at load 0.37: the longest chain among the 64 tokens a warp has in flight averaged 5 DRAM round
// trips, ncu long-scoreboard 51 cycles per issue).  Inside a bucket the further probes hit the line the first
// one fetched (L2 fetch granularity = the bucket, xf_table_create); chains that leave the home bucket are rare
// (1.6 % of the keys at load 0.37 with B = 4, simulated; expected longest chain among 64 tokens 1.7 lines).
__device__ __forceinline__ uint64_t xf_probe_slot(const XfTableView& t, uint64_t key, uint32_t i) {
  const uint64_t m = key * 0x9E3779B97F4A7C15ull;
  const uint32_t bs = t.bshift;
  const uint64_t hb = m >> (64 - (t.log2cap - bs));  // home bucket
  const uint32_t b1 = (1u << bs) - 1u;
  const uint32_t j0 = (uint32_t)(m >> 9) & b1;       // where the walk through the home bucket starts
  const uint32_t k = i >> bs;
  const uint32_t j = (k == 0) ? ((j0 + i) & b1) : (i & b1);
  return (((hb + k) << bs) | j) & t.mask;
}
__device__ __forceinline__ uint64_t xf_home_slot(const XfTableView& t, uint64_t key) { return xf_probe_slot(t, key, 0); }

// ---- counter-based initial value of the latent table; must stay bit-identical to
// ---- oracle/xflow_oracle.cc: xo_counter_normal (integer ops + two exactly rounded float ops)
__host__ __device__ __forceinline__ uint64_t xf_splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__device__ __forceinline__ float xf_counter_normal(uint64_t key, uint32_t k, uint64_t seed) {
  uint64_t base = xf_splitmix64(key ^ xf_splitmix64(seed + 0x632BE59BD9B4E019ull * (uint64_t)(k + 1)));
  uint32_t sum = 0;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    uint64_t x = xf_splitmix64(base + (uint64_t)r);
    sum += (uint32_t)(x & 0xFFFF) + (uint32_t)((x >> 16) & 0xFFFF) + (uint32_t)((x >> 32) & 0xFFFF) +
           (uint32_t)((x >> 48) & 0xFFFF);
  }
  float u = __fmul_rn((float)((int32_t)sum - 6 * 65535), 1.0f / 65536.0f);
  return __fmul_rn(u, 1e-2f);
}
__device__ __forceinline__ float xf_v_init(const XfTableView& t, uint64_t key, uint32_t k) {
  if (t.v_init == XF_INIT_COUNTER) return xf_counter_normal(key, k, t.seed);
  if (t.v_init == XF_INIT_ZERO) return 0.0f;
  return t.v_const;
}

// ---- row accessors
__device__ __forceinline__ uint8_t* xf_row(const XfTableView& t, uint64_t slot) {
  return t.base + slot * (uint64_t)t.stride;
}
__device__ __forceinline__ double* xf_row_g(uint8_t* row) { return reinterpret_cast<double*>(row + 8); }
#define XF_OFF_STATE 16  // {w, n, z, flags}: one aligned 16-byte word
#define XF_OFF_FLAGS 28
__device__ __forceinline__ float* xf_row_v(uint8_t* row) { return reinterpret_cast<float*>(row + 32); }
__device__ __forceinline__ double* xf_row_acc(uint8_t* row, int K) { return reinterpret_cast<double*>(row + xf_acc_off(K)); }
__device__ __forceinline__ float* xf_row_nv(uint8_t* row, int K) { return reinterpret_cast<float*>(row + xf_acc_off(K) + 16); }
__device__ __forceinline__ float* xf_row_zv(uint8_t* row, int K) { return xf_row_nv(row, K) + K; }

struct XfHead {
  uint64_t key;
  uint32_t flags;
  float w, n, z;
  double g;
};

__device__ __forceinline__ XfHead xf_load_head(const uint8_t* row) {
  // one 32-byte sector = ONE 256-bit L2 load (sm_100 LDG.E.256; L1 is useless for random rows).
  // Two 16-byte loads cost two L2 requests and, measured with ncu, up to two DRAM fetches.
  uint64_t q0, q1, q2, q3;
  asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(q0), "=l"(q1), "=l"(q2), "=l"(q3) : "l"(row));
  XfHead h;
  h.key = q0;
  h.g = __longlong_as_double((long long)q1);
  h.w = __uint_as_float((uint32_t)q2);
  h.n = __uint_as_float((uint32_t)(q2 >> 32));
  h.z = __uint_as_float((uint32_t)q3);
  h.flags = (uint32_t)(q3 >> 32);
  return h;
}

// Same sector through L1 (ld.global.ca).  Only for kernels in which w / n / z / flags of existing rows do
// not change while the kernel runs (the eager step: it adds to g / gv with L2 atomics and inserts keys,
// nothing else).  A stale copy can then differ from L2 only by showing EMPTY for a slot that was claimed
// during this kernel — the insert CAS resolves that (xf_probe_from), and a row inserted during the kernel
// still has its default parameters.  With skewed ids this keeps each SM's reads of the hot rows local:
// through L2 every token of the hottest key queues on one slice (measured).
__device__ __forceinline__ XfHead xf_load_head_l1(const uint8_t* row) {
  uint64_t q0, q1, q2, q3;
  asm volatile("ld.global.ca.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(q0), "=l"(q1), "=l"(q2), "=l"(q3) : "l"(row));
  XfHead h;
  h.key = q0;
  h.g = __longlong_as_double((long long)q1);
  h.w = __uint_as_float((uint32_t)q2);
  h.n = __uint_as_float((uint32_t)(q2 >> 32));
  h.z = __uint_as_float((uint32_t)q3);
  h.flags = (uint32_t)(q3 >> 32);
  return h;
}

// eight consecutive floats of a latent row with ONE 256-bit load through L1 (read-only-for-the-kernel data,
// see xf_load_head_l1): a row access costs per instruction, not per byte (tools/membench.cu)
__device__ __forceinline__ void xf_ld8_l1(const float* p, float (&o)[8]) {
  uint64_t q0, q1, q2, q3;
  asm volatile("ld.global.ca.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(q0), "=l"(q1), "=l"(q2), "=l"(q3) : "l"(p));
  o[0] = __uint_as_float((uint32_t)q0); o[1] = __uint_as_float((uint32_t)(q0 >> 32));
  o[2] = __uint_as_float((uint32_t)q1); o[3] = __uint_as_float((uint32_t)(q1 >> 32));
  o[4] = __uint_as_float((uint32_t)q2); o[5] = __uint_as_float((uint32_t)(q2 >> 32));
  o[6] = __uint_as_float((uint32_t)q3); o[7] = __uint_as_float((uint32_t)(q3 >> 32));
}

// full-sector store of the head (one 256-bit STG: no partial-sector write, no read-for-fill)
__device__ __forceinline__ void xf_store_head(uint8_t* row, const XfHead& h) {
  const uint64_t q1 = (uint64_t)__double_as_longlong(h.g);
  const uint64_t q2 = (uint64_t)__float_as_uint(h.w) | ((uint64_t)__float_as_uint(h.n) << 32);
  const uint64_t q3 = (uint64_t)__float_as_uint(h.z) | ((uint64_t)h.flags << 32);
  asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(row), "l"(h.key), "l"(q1), "l"(q2), "l"(q3) : "memory");
}

// Find `key` starting at its home slot `s` (= xf_home_slot) whose head `h` the caller has already loaded; if
// INSERT, claim an empty slot for it when absent (store[key] semantics).  Returns the slot index, or -1 (not found
// without INSERT, or probe overflow -> *t.error = 1).  On return `h` is the row's first sector as it
// was when the key matched (or the default contents on insert).
template <bool INSERT>
__device__ __forceinline__ int64_t xf_probe_from(const XfTableView& t, uint64_t key, uint64_t s, XfHead& h) {
  for (int probes = 0; probes < XF_MAX_PROBE; ++probes) {
    if (h.key == key) return (int64_t)s;
    if (h.key == XF_EMPTY_KEY) {
      if (!INSERT) return -1;
      unsigned long long old =
          atomicCAS(reinterpret_cast<unsigned long long*>(xf_row(t, s)), (unsigned long long)XF_EMPTY_KEY,
                    (unsigned long long)key);
      if (old == XF_EMPTY_KEY) {
        // we created the entry: count it (warp-aggregated) and report default contents
        unsigned m = __activemask();
        int leader = __ffs(m) - 1;
        if ((int)(threadIdx.x & 31) == leader) atomicAdd(t.size, (unsigned long long)__popc(m));
        h.key = key; h.flags = 0; h.w = 0.f; h.n = 0.f; h.z = 0.f;
        h.g = t.lazy ? 0.0 : -0.0;  // what xf_k_fill left in the row (lazy: the integer 0)
        return (int64_t)s;
      }
      if (old == key) {
        // raced with another inserter of the same key: the parameter fields are still defaults
        h.key = key;
        return (int64_t)s;
      }
      // a different key took the slot: fall through to the next one
    }
    s = xf_probe_slot(t, key, (uint32_t)probes + 1u);
    h = xf_load_head(xf_row(t, s));
  }
  *t.error = 1;
  return -1;
}

template <bool INSERT>
__device__ __forceinline__ int64_t xf_probe(const XfTableView& t, uint64_t key, XfHead* head) {
  const uint64_t s = xf_home_slot(t, key);
  XfHead h = xf_load_head(xf_row(t, s));
  const int64_t r = xf_probe_from<INSERT>(t, key, s, h);
  *head = h;
  return r;
}

// ---- arithmetic restated from the reference, IEEE-rounded op by op (no FMA contraction) ----

// Base::sigmoid  src/base/base.h:54-63
__device__ __forceinline__ float xf_sigmoid(float x) {
  if (x < -30.f) return (float)1e-6;
  if (x > 30.f) return 1.0f;
  // pow(2.718281828, x) as exp(x * ln 2.718281828) in double: same value to ~3e-15 relative (far below
  // the final float rounding), a third of the registers and instructions of the generic double pow
  const double ex = exp((double)x * 0.9999999998311266);
  return (float)(ex / (1.0 + ex));
}

// FTRL-proximal coordinate update  src/optimizer/ftrl.h:59-74 (== :126-141)
__device__ __forceinline__ void xf_ftrl_coord(const XfTableView& t, float g, float& w, float& n, float& z) {
  float old_n = n;
  float nn = __fadd_rn(old_n, __fmul_rn(g, g));
  float sig = __fdiv_rn(__fsub_rn(__fsqrt_rn(nn), __fsqrt_rn(old_n)), t.alpha);
  z = __fadd_rn(z, __fsub_rn(g, __fmul_rn(sig, w)));
  n = nn;
  if (fabsf(z) <= t.lambda1) {
    w = 0.0f;
  } else {
    float tmpr = 0.0f;
    if (z > 0.0f) tmpr = __fsub_rn(z, t.lambda1);
    if (z < 0.0f) tmpr = __fadd_rn(z, t.lambda1);
    float tmpl = -__fadd_rn(__fdiv_rn(__fadd_rn(t.beta, __fsqrt_rn(n)), t.alpha), t.lambda2);
    w = __fdiv_rn(tmpr, tmpl);
  }
}

// SGD coordinate update  src/optimizer/sgd.h:52,96
__device__ __forceinline__ void xf_sgd_coord(const XfTableView& t, float g, float& w) {
  w = __fsub_rn(w, __fmul_rn(t.learning_rate, g));
}

__device__ __forceinline__ void xf_opt_coord(const XfTableView& t, float g, float& w, float& n, float& z) {
  if (t.opt == XF_OPT_FTRL) xf_ftrl_coord(t, g, w, n, z);
  else xf_sgd_coord(t, g, w);
}

// push_gradient[i] /= 1.0 * loss.size()  lr_worker.cc:116-118 ; fm_worker.cc:150-156 (double divide)
__device__ __forceinline__ float xf_div_rows_plain(float g, double rows) { return (float)((double)g / rows); }
__device__ __forceinline__ float xf_div_rows(float g, double rows) {
  // A power-of-two row count (the usual batch size) makes the quotient an exact scaling: multiplying by
  // the exact reciprocal gives the same double, hence the same float, without a double division (the
  // update kernel was instruction-bound on it, profiles/r01_ncu_full_fm_ftrl.md).
  const long long b = __double_as_longlong(rows);
  if ((b & 0x000FFFFFFFFFFFFFll) == 0ll && b > 0ll) {
    const double inv = __longlong_as_double((2046ll << 52) - b);  // 2^-e for rows = 2^e
    return (float)((double)g * inv);
  }
  return (float)((double)g / rows);
}

// Lazy tables: the row as the reference's server would hold it — i.e. with the pending optimizer step
// (the Push of the batch named by the tag) applied.  Pure function of the snapshot; writes nothing.
__device__ __forceinline__ bool xf_has_pending(const XfTableView& t, const XfHead& h) {
  return t.lazy && h.flags != 0u && h.flags != XF_TAG_LOCKED;
}
// lazy tables: the residual sum of a row (fixed point, see XF_FIX_SCALE) and a residual in that unit
__device__ __forceinline__ long long xf_head_gfix(const XfHead& h) { return __double_as_longlong(h.g); }
__device__ __forceinline__ long long xf_fix_of(float residual) { return __double2ll_rn((double)residual * XF_FIX_SCALE); }
// the pending optimizer step of a lazy row: gradient = (float)(residual sum) / rows   (lr_worker.cc:116-118)
__device__ __forceinline__ void xf_fold_pending(const XfTableView& t, long long gfix, uint32_t rows, float& w, float& n,
                                                float& z) {
  // the residual sum is rounded to float once (push_gradient is a float vector), then divided in double
  const float g = xf_div_rows_plain((float)((double)gfix * XF_FIX_INV), (double)rows);
  xf_opt_coord(t, g, w, n, z);
}
__device__ __forceinline__ void xf_apply_pending(const XfTableView& t, XfHead& h) {
  if (!xf_has_pending(t, h)) return;
  xf_fold_pending(t, xf_head_gfix(h), __ldg(t.rows_by_seq + h.flags), h.w, h.n, h.z);
  h.flags = 0u;
  h.g = 0.0;  // all-zero bits: also the fixed-point zero
}

// ---- lazy tables: claim AND publish a row with ONE 128-bit compare-and-swap on {w, n, z, tag} -------------
// Measured on B200 (tools/membench.cu, profiles/r02_membench.md): on a multi-GB table every instruction that
// touches a random row costs about the same whatever it is — load, store, CAS or RED, hit or miss (the
// translation / request path saturates near 36 G requests/s) — so the number of row-touching instructions
// per token is what sets the speed of these kernels.  The round-1 protocol used four (load, CAS on the tag,
// 256-bit store, RED); this one uses three (load, CAS.128, RED) and has no LOCKED state to poll.
struct XfState {
  float w, n, z;
  uint32_t flags;
};
__device__ __forceinline__ bool xf_cas_state(uint8_t* rowp, const XfState& expect, const XfState& desired, XfState& found) {
  const uint64_t e0 = (uint64_t)__float_as_uint(expect.w) | ((uint64_t)__float_as_uint(expect.n) << 32);
  const uint64_t e1 = (uint64_t)__float_as_uint(expect.z) | ((uint64_t)expect.flags << 32);
  const uint64_t d0 = (uint64_t)__float_as_uint(desired.w) | ((uint64_t)__float_as_uint(desired.n) << 32);
  const uint64_t d1 = (uint64_t)__float_as_uint(desired.z) | ((uint64_t)desired.flags << 32);
  uint64_t o0, o1;
  asm volatile(
      "{\n .reg .b128 cmp, swp, old;\n mov.b128 cmp, {%2, %3};\n mov.b128 swp, {%4, %5};\n"
      " atom.global.cas.b128 old, [%6], cmp, swp;\n mov.b128 {%0, %1}, old;\n}"
      : "=l"(o0), "=l"(o1)
      : "l"(e0), "l"(e1), "l"(d0), "l"(d1), "l"(rowp + XF_OFF_STATE)
      : "memory");
  found.w = __uint_as_float((uint32_t)o0);
  found.n = __uint_as_float((uint32_t)(o0 >> 32));
  found.z = __uint_as_float((uint32_t)o1);
  found.flags = (uint32_t)(o1 >> 32);
  return o0 == e0 && o1 == e1;
}
// "Open" a row for batch `seq` from the snapshot `h` a token has loaded: the first token of the batch that gets
// there folds the pending optimizer step of the row's previous batch in and stamps the row with `seq`, all in the
// one CAS; everybody else finds (or is handed back by the failed CAS) the published weight.  Returns the weight
// the batch pulls.  `won` = this token opened the row; it then OWES the row the removal of the consumed residual
// sum: `pend` = the raw 64 bits of g in its snapshot, to be subtracted in the same RED that adds the token's own
// residual (integer arithmetic mod 2^64: exact, whatever lands in between).  g of the snapshot is final when the
// CAS succeeds: residuals of batch `seq` are only added to rows that were seen open, i.e. after this CAS.
// `stale` (optional): the caller's snapshot may be OLDER than this batch (the sharded owner keeps the snapshot its
// Pull took, mg_kernels.cu); a CAS that fails against a row which is not open for `seq` then reports *stale = true
// instead of an error, and the caller reloads the row and tries again.
__device__ __forceinline__ float xf_lazy_open(const XfTableView& t, uint8_t* rowp, const XfHead& h, uint32_t seq,
                                              bool& won, unsigned long long& pend, bool* stale = nullptr) {
  won = false;
  pend = 0ull;
  if (stale) *stale = false;
  if (h.flags == seq) return h.w;
  XfState e{h.w, h.n, h.z, h.flags}, d = e, f;
  if (h.flags != 0u) xf_fold_pending(t, xf_head_gfix(h), __ldg(t.rows_by_seq + h.flags), d.w, d.n, d.z);
  d.flags = seq;
  if (xf_cas_state(rowp, e, d, f)) {
    won = true;
    pend = (unsigned long long)__double_as_longlong(h.g);
    return d.w;
  }
  if (f.flags != seq) {
    if (stale) *stale = true;
    else *t.error = 2;  // inside a batch a row only ever goes from "pending" to "open for seq"
  }
  return f.w;
}
__device__ __forceinline__ void xf_lazy_add(uint8_t* rowp, unsigned long long fix) {
  atomicAdd(reinterpret_cast<unsigned long long*>(rowp + 8), fix);
}

__device__ __forceinline__ float xf_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

#endif  // __CUDACC__
