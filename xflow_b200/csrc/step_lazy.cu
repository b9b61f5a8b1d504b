// LR with "update on next touch" (lazy tables, K == 0): ONE kernel per batch, no optimizer kernel.
//
// The optimizer step of batch p for a key is not applied when batch p ends; the residual sum stays in
// the row (g, tagged p) and is folded in by the first token of a later batch b that touches the row
// ("opening" the row for b).  Every other reader applies it on the fly (xf_apply_pending, table.cuh), so
// the observable table is the reference's at every batch boundary.
//
// Why: on a multi-GB table every row-touching instruction costs about the same (~28 ps of chip time: load,
// store or atomic, hit or miss — tools/membench.cu, profiles/r02_membench.md), so the kernel's time is
// (row-touching instructions per token) x 28 ps x tokens.  The eager pair (step + update) needs 4.3 per token,
// round 1's lazy protocol (load, CAS on the tag, 256-bit store, RED) also 4.3 but in one launch; this one needs
// 3.3: load (1.3 with the collision probes), ONE 128-bit CAS that claims and publishes, one integer RED.
//
// Protocol per row, batch b (tag = the row's flags word; xf_lazy_open in table.cuh):
//   tag == b              open for b: w is current, g accumulates batch b.
//   tag == p (p < b)      pending (p == 0: nothing pending): the token computes the row's new state from its
//                         snapshot (FTRL/SGD step with g / rows[p]) and tries
//                         CAS.128({w,n,z,tag}: snapshot -> {w',n',z', b}).  Exactly one token of the batch
//                         succeeds; the others are handed the published state back by their failed CAS.
//                         Nobody ever waits: there is no locked state.
//   g                     64-bit fixed-point residual sum (scale 2^40).  The opener owes the row "- g_pending";
//                         it pays in the same RED that adds its own residual after the row reduction
//                         (integer adds commute exactly, so the order in which REDs land is irrelevant and the
//                         result is bit-reproducible).
// Inside a warp, tokens with the same slot elect one lane (__match_any_sync): one open and one RED of
// count x residual per distinct key of a 32-token group.
//
// Tried and measured on the 1e8-id table (profiles/r02_lazy_experiments.md): bucketised probing (collision
// probes inside one 128-byte line: -3 %, kept), L2 prefetch by dedicated warps running ahead (+23 % time: the
// prefetches are requests too, removed).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "kernels.h"
#include "table.cuh"

#define XF_NO_SLOT 0xFFFFFFFFu
#define XF_LAZY_CACHED 2  // 64-token chunks whose leader slots stay in registers between the two phases

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
    // first XF_LAZY_CACHED chunks (rows <= 128 tokens), per half: the slot if this lane leads its group, the group
    // size, and what the lane owes the row if it opened it
    uint32_t lead_s[2 * XF_LAZY_CACHED];
    uint32_t cnt_c[XF_LAZY_CACHED];  // 8 bits per half
    unsigned long long pend_c[2 * XF_LAZY_CACHED];
#pragma unroll
    for (int c = 0; c < XF_LAZY_CACHED; ++c) {
      lead_s[2 * c] = lead_s[2 * c + 1] = XF_NO_SLOT;
      pend_c[2 * c] = pend_c[2 * c + 1] = 0ull;
      cnt_c[c] = 0;
    }

    // ---------------- phase A: pull (and, in training, open) every token's row
    for (int ch = 0; ch < chunks; ++ch) {
      const uint32_t j0 = beg + (uint32_t)ch * 64u + (uint32_t)lane;
      const uint32_t j1 = j0 + 32u;
      const bool v0 = j0 < end, v1 = j1 < end;
      const uint64_t k0 = v0 ? __ldcs(keys + j0) : 0ull;
      const uint64_t k1 = v1 ? __ldcs(keys + j1) : 0ull;
      const uint64_t p0 = xf_home_slot(t, k0), p1 = xf_home_slot(t, k1);
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
      unsigned long long pend0 = 0ull, pend1 = 0ull;
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
        bool won0 = false, won1 = false;
        if (L0) { w0 = xf_lazy_open(t, xf_row(t, s0), h0, seq, won0, pend0); cnt0 = (uint32_t)__popc(grp0); }
        if (L1) { w1 = xf_lazy_open(t, xf_row(t, s1), h1, seq, won1, pend1); cnt1 = (uint32_t)__popc(grp1); }
        open_acc += (won0 ? 1u : 0u) + (won1 ? 1u : 0u);
        w0 = __shfl_sync(0xffffffffu, w0, lead0);
        w1 = __shfl_sync(0xffffffffu, w1, lead1);
        if (s0 == XF_NO_SLOT) w0 = 0.f;
        if (s1 == XF_NO_SLOT) w1 = 0.f;
        if (ch >= XF_LAZY_CACHED) {
          // long rows (> 128 tokens): nothing is remembered for phase B, the opener pays its debt at once
          if (won0) xf_lazy_add(xf_row(t, s0), 0ull - pend0);
          if (won1) xf_lazy_add(xf_row(t, s1), 0ull - pend1);
        }
      }
      wsum += w0;
      wsum += w1;
#pragma unroll
      for (int c = 0; c < XF_LAZY_CACHED; ++c)
        if (ch == c) {
          lead_s[2 * c] = cnt0 ? s0 : XF_NO_SLOT;
          lead_s[2 * c + 1] = cnt1 ? s1 : XF_NO_SLOT;
          pend_c[2 * c] = pend0;
          pend_c[2 * c + 1] = pend1;
          cnt_c[c] = cnt0 | (cnt1 << 8);
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
    // are L2-resident right now, which is why this is not a separate kernel.  One lane per distinct slot of a
    // group adds count x residual (minus what it owes as the row's opener): integer, exact.
    const unsigned long long lf = (unsigned long long)xf_fix_of(loss);
#pragma unroll
    for (int c = 0; c < XF_LAZY_CACHED; ++c) {
      if (lead_s[2 * c] != XF_NO_SLOT)
        xf_lazy_add(xf_row(t, lead_s[2 * c]), lf * (unsigned long long)(cnt_c[c] & 0xFFu) - pend_c[2 * c]);
      if (lead_s[2 * c + 1] != XF_NO_SLOT)
        xf_lazy_add(xf_row(t, lead_s[2 * c + 1]), lf * (unsigned long long)(cnt_c[c] >> 8) - pend_c[2 * c + 1]);
    }
    for (int ch = XF_LAZY_CACHED; ch < chunks; ++ch) {
      // long rows (> 128 tokens): slots are not cached; the rows were opened in phase A
      const uint32_t j0 = beg + (uint32_t)ch * 64u + (uint32_t)lane;
      const uint32_t j1 = j0 + 32u;
      XfHead h;
      if (j0 < end) { const int64_t r = xf_probe<false>(t, __ldg(keys + j0), &h); if (r >= 0) xf_lazy_add(xf_row(t, (uint64_t)r), lf); }
      if (j1 < end) { const int64_t r = xf_probe<false>(t, __ldg(keys + j1), &h); if (r >= 0) xf_lazy_add(xf_row(t, (uint64_t)r), lf); }
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
