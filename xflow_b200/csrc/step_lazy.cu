// LR with "update on next touch" (lazy tables, K == 0): ONE kernel per batch, no optimizer kernel.
//
// The optimizer step of batch p for a key is not applied when batch p ends; the residual sum stays in
// the row (g, tagged p) and is folded in by the first token of a later batch b that touches the row
// ("opening" the row for b).  Every other reader applies it on the fly (xf_apply_pending, table.cuh), so
// the observable table is the reference's at every batch boundary.  Why: both kernels of the eager path
// run at the DRAM random-request ceiling (43 G sectors/s, profiles/r01_randsector.md); the only lever
// left is the number of requests, and this removes the update kernel's read + write and the step's
// dirty write-back (17.1 M -> 9.5 M DRAM sector requests per cfg2 batch, measured with ncu).  The
// gradient accumulation must stay in the same kernel as the open: run as a second kernel it finds the
// rows evicted again (measured: +180 us).
//
// Protocol per row, batch b (tag = the row's flags word):
//   tag == b              open for b: w is current, g accumulates batch b.
//   tag == p (0 < p < b)  pending: exactly one token wins atomicCAS(tag, p, LOCKED), applies
//   tag == 0              FTRL/SGD(g / rows[p]) (nothing for 0), and publishes {w,n,z, tag = b, g = 0}
//                         with one 256-bit store.  Nobody else writes the row while it is LOCKED.
//   tag == LOCKED         another token is opening: poll until tag == b, then read w.
// Rules that keep it deadlock-free: a winner publishes immediately (it never waits while holding a
// claim, for either of its two tokens); inside a warp, tokens with the same slot elect one lane
// (__match_any_sync) so a warp never waits on itself.  The same groups let one lane add
// count x residual in phase B.  Waits are bounded (error code 2 instead of a hung GPU).
//
// The kernel is dependency-chain bound, not bandwidth bound (ncu: DRAM 10 %, L2 24 %, issue 15 % busy;
// ablation: read-only pass 190 us, + claim/publish 130 us, + accumulate 75 us): see DESIGN.md section 6.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"
#include "table.cuh"

#define XF_NO_SLOT 0xFFFFFFFFu

__device__ __forceinline__ uint32_t xf_ld_tag(const uint8_t* rowp) {
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(rowp + XF_OFF_FLAGS));
  return v;
}

// volatile 256-bit re-read of a row whose tag this thread has just observed to be `seq`
__device__ __forceinline__ float xf_reload_w(const uint8_t* rowp) {
  uint64_t q0, q1, q2, q3;
  asm volatile("ld.volatile.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(q0), "=l"(q1), "=l"(q2), "=l"(q3) : "l"(rowp));
  (void)q0; (void)q2; (void)q3;
  return __uint_as_float((uint32_t)q1);
}

// poll until the row is open for `seq` (its opener never waits, so this ends within a store latency)
__device__ __forceinline__ void xf_wait_open(const XfTableView& t, const uint8_t* rowp, uint32_t seq) {
  for (int spin = 0; xf_ld_tag(rowp) != seq; ++spin) {
    if (spin > (1 << 22)) { *t.error = 2; break; }
  }
}

// fold the pending step (batch h.flags, `rows` rows) into the snapshot and stamp it open for `seq`
__device__ __forceinline__ void xf_open_snapshot(const XfTableView& t, XfHead& h, uint32_t rows, uint32_t seq) {
  if (h.flags != 0u) {
    const float g = xf_div_rows_plain((float)h.g, (double)rows);  // lr_worker.cc:116-118 (the fast path spills here)
    xf_opt_coord(t, g, h.w, h.n, h.z);                       // ftrl.h:59-74 / sgd.h:52
  }
  h.flags = seq;
  h.g = 0.0;
}

__global__ void __launch_bounds__(256)
xf_k_step_lr_lazy(XfTableView t, const uint32_t* __restrict__ row_ptr, const uint64_t* __restrict__ keys,
                  const uint8_t* __restrict__ labels, int B, int mode, uint32_t seq, uint32_t* rows_by_seq,
                  float* __restrict__ loss_out, float* __restrict__ pctr_out, float* __restrict__ abs_loss_sum,
                  unsigned long long* __restrict__ unique_total) {
  __shared__ float s_abs[8];
  __shared__ unsigned int s_open;
  if (threadIdx.x == 0) s_open = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0 && mode == 0) rows_by_seq[seq] = (uint32_t)B;  // read by later batches only
  __syncthreads();
  float abs_acc = 0.f;
  unsigned int open_acc = 0;
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int gwarp = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * warps_per_block;

  for (int row = gwarp; row < B; row += nwarps) {
    const uint32_t beg = __ldg(row_ptr + row);
    const uint32_t end = __ldg(row_ptr + row + 1);
    const int chunks = (int)((end - beg + 63u) >> 6);
    float wsum = 0.f;
    uint32_t lead_s0 = XF_NO_SLOT, lead_s1 = XF_NO_SLOT;  // first chunk: slot if this lane leads its group
    uint32_t cnt_c = 0;                                    //              and the group sizes (8 bits each)

    // ---------------- phase A: pull (and, in training, open) every token's row
    for (int ch = 0; ch < chunks; ++ch) {
      const uint32_t j0 = beg + (uint32_t)ch * 64u + (uint32_t)lane;
      const uint32_t j1 = j0 + 32u;
      const bool v0 = j0 < end, v1 = j1 < end;
      const uint64_t k0 = v0 ? __ldcs(keys + j0) : 0ull;
      const uint64_t k1 = v1 ? __ldcs(keys + j1) : 0ull;
      const uint64_t p0 = xf_slot_hash(k0, t.log2cap), p1 = xf_slot_hash(k1, t.log2cap);
      XfHead h0, h1;
      h0.key = h1.key = XF_EMPTY_KEY;
      h0.flags = h1.flags = 0u;
      if (v0) h0 = xf_load_head(xf_row(t, p0));
      if (v1) h1 = xf_load_head(xf_row(t, p1));
      uint32_t s0 = XF_NO_SLOT, s1 = XF_NO_SLOT;
      if (v0) { const int64_t r = xf_probe_from<true>(t, k0, p0, h0); if (r >= 0) s0 = (uint32_t)r; }
      if (v1) { const int64_t r = xf_probe_from<true>(t, k1, p1, h1); if (r >= 0) s1 = (uint32_t)r; }
      float w0 = 0.f, w1 = 0.f;
      uint32_t cnt0 = 0, cnt1 = 0;
      if (mode == 1) {
        // forward only: apply a pending step on the fly, write nothing
        if (s0 != XF_NO_SLOT) { xf_apply_pending(t, h0); w0 = h0.w; }
        if (s1 != XF_NO_SLOT) { xf_apply_pending(t, h1); w1 = h1.w; }
      } else {
        // lanes with the same slot elect their lowest lane; invalid lanes get unique dummy values
        const unsigned grp0 = __match_any_sync(0xffffffffu, (s0 != XF_NO_SLOT) ? s0 : (0xFFFFFF00u | (uint32_t)lane));
        const unsigned grp1 = __match_any_sync(0xffffffffu, (s1 != XF_NO_SLOT) ? s1 : (0xFFFFFF00u | (uint32_t)lane));
        const int lead0 = __ffs(grp0) - 1, lead1 = __ffs(grp1) - 1;
        const bool L0 = s0 != XF_NO_SLOT && lane == lead0, L1 = s1 != XF_NO_SLOT && lane == lead1;
        uint8_t* r0 = xf_row(t, s0);
        uint8_t* r1 = xf_row(t, s1);
        const bool try0 = L0 && h0.flags != seq && h0.flags != XF_TAG_LOCKED;
        const bool try1 = L1 && h1.flags != seq && h1.flags != XF_TAG_LOCKED;
        // divisors of the pending steps and both claims are issued before any result is consumed
        uint32_t prow0 = 1, prow1 = 1, old0 = 0, old1 = 0;
        if (try0 && h0.flags) prow0 = __ldg(t.rows_by_seq + h0.flags);
        if (try1 && h1.flags) prow1 = __ldg(t.rows_by_seq + h1.flags);
        if (try0) old0 = atomicCAS(reinterpret_cast<unsigned int*>(r0 + XF_OFF_FLAGS), h0.flags, XF_TAG_LOCKED);
        if (try1) old1 = atomicCAS(reinterpret_cast<unsigned int*>(r1 + XF_OFF_FLAGS), h1.flags, XF_TAG_LOCKED);
        // Winners publish at once, for both halves, BEFORE anybody waits: a claim is never held across a
        // wait (holding the half-1 claim while spinning on a half-0 row deadlocked two warps).
        const bool won0 = try0 && old0 == h0.flags;
        const bool won1 = try1 && old1 == h1.flags;
        if (won0) { xf_open_snapshot(t, h0, prow0, seq); xf_store_head(r0, h0); ++open_acc; }
        if (won1) { xf_open_snapshot(t, h1, prow1, seq); xf_store_head(r1, h1); ++open_acc; }
        __syncwarp();
        if (L0) {
          w0 = h0.w;
          if (!won0 && h0.flags != seq) { xf_wait_open(t, r0, seq); w0 = xf_reload_w(r0); }
          cnt0 = (uint32_t)__popc(grp0);
        }
        if (L1) {
          w1 = h1.w;
          if (!won1 && h1.flags != seq) { xf_wait_open(t, r1, seq); w1 = xf_reload_w(r1); }
          cnt1 = (uint32_t)__popc(grp1);
        }
        w0 = __shfl_sync(0xffffffffu, w0, lead0);
        w1 = __shfl_sync(0xffffffffu, w1, lead1);
        if (s0 == XF_NO_SLOT) w0 = 0.f;
        if (s1 == XF_NO_SLOT) w1 = 0.f;
      }
      wsum += w0;
      wsum += w1;
      if (ch == 0) {
        lead_s0 = cnt0 ? s0 : XF_NO_SLOT;
        lead_s1 = cnt1 ? s1 : XF_NO_SLOT;
        cnt_c = cnt0 | (cnt1 << 8);
      }
    }

    const float wx = xf_warp_sum(wsum);
    const float pctr = xf_sigmoid(wx);
    if (mode == 1) {
      if (lane == 0 && pctr_out) pctr_out[row] = pctr;
      continue;
    }
    const float loss = __fsub_rn(pctr, (float)labels[row]);  // lr_worker.cc:141
    if (lane == 0 && loss_out) loss_out[row] = loss;
    abs_acc += fabsf(loss);
    // ---------------- phase B: residual into the per-key sums (every row is open for `seq`); the rows
    // are L2-resident right now, which is why this is not a separate kernel
    const double gd = (double)loss;
    // one lane per distinct slot of the round adds count x residual (exact in double)
    if (lead_s0 != XF_NO_SLOT) atomicAdd(xf_row_g(xf_row(t, lead_s0)), gd * (double)(cnt_c & 0xFFu));
    if (lead_s1 != XF_NO_SLOT) atomicAdd(xf_row_g(xf_row(t, lead_s1)), gd * (double)(cnt_c >> 8));
    for (int ch = 1; ch < chunks; ++ch) {
      // long rows (> 64 tokens): slots are not cached; the rows were opened in phase A
      const uint32_t j0 = beg + (uint32_t)ch * 64u + (uint32_t)lane;
      const uint32_t j1 = j0 + 32u;
      XfHead h;
      if (j0 < end) { const int64_t r = xf_probe<false>(t, __ldg(keys + j0), &h); if (r >= 0) atomicAdd(xf_row_g(xf_row(t, (uint64_t)r)), gd); }
      if (j1 < end) { const int64_t r = xf_probe<false>(t, __ldg(keys + j1), &h); if (r >= 0) atomicAdd(xf_row_g(xf_row(t, (uint64_t)r)), gd); }
    }
  }
  if (mode == 0) {
    if (lane == 0) s_abs[threadIdx.x >> 5] = abs_acc;
    if (open_acc) atomicAdd(&s_open, open_acc);
    __syncthreads();
    if (threadIdx.x == 0) {
      if (abs_loss_sum != nullptr) {
        float tot = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s_abs[w];
        atomicAdd(abs_loss_sum, tot);
      }
      if (unique_total != nullptr && s_open) atomicAdd(unique_total, (unsigned long long)s_open);
    }
  }
}

void xf_launch_step_lr_lazy(const XfTableView& t, const uint32_t* row_ptr, const uint64_t* keys,
                            const uint8_t* labels, int B, int mode, uint32_t seq, uint32_t* rows_by_seq,
                            float* loss_out, float* pctr_out, float* abs_loss_sum, unsigned long long* unique_total,
                            cudaStream_t st) {
  if (B <= 0) return;
  const int grid = xf_grid_for((uint64_t)B * 32, 256, 8);
  xf_k_step_lr_lazy<<<grid, 256, 0, st>>>(t, row_ptr, keys, labels, B, mode, seq, rows_by_seq, loss_out, pctr_out,
                                           abs_loss_sum, unique_total);
}
