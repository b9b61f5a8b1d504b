// Test-set metric of the reference, host side (not on the hot path): Base::calculate_auc
// (src/base/base.h:84-110).  Kept faithful to the reference's quirks because its printed numbers
// are the only accuracy figures xflow ever reports: base-2 log-likelihood that is NOT negated,
// accumulated in a float; rank-AUC with a float `area` accumulator; ties ordered by std::sort.
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <vector>

#include "../../include/xflow_b200.h"

namespace {
struct Scored {
  int label;
  float pctr;
};
}  // namespace

XF_DLL int xf_auc_logloss(const int32_t* labels, const float* pctr, uint64_t n, double out[4]) {
  if ((!labels || !pctr) && n) return XF_ERR_ARG;
  if (!out) return XF_ERR_ARG;
  std::vector<Scored> v((size_t)n);
  for (uint64_t i = 0; i < n; ++i) {
    v[i].label = labels[i];
    v[i].pctr = pctr[i];
  }
  std::sort(v.begin(), v.end(), [](const Scored& a, const Scored& b) { return a.pctr > b.pctr; });
  float area = 0.0f;   // base.h:89
  float ll = 0.0f;     // base.h:113 (member `logloss`)
  int positives = 0;
  for (const Scored& s : v) {
    if (s.label == 1) ++positives;
    else area += positives;
    // float term + double term, accumulated into a float (base.h:97-98)
    const float pos_term = s.label * log2f(s.pctr);
    const double neg_term = (1.0 - s.label) * log2(1.0 - s.pctr);
    ll = (float)((double)ll + ((double)pos_term + neg_term));
  }
  ll /= v.size();
  out[0] = ll;
  out[2] = positives;
  out[3] = (double)(v.size() - (size_t)positives);
  if (positives == 0 || (size_t)positives == v.size()) {
    out[1] = NAN;
  } else {
    area /= 1.0 * (positives * (v.size() - positives));
    out[1] = area;
  }
  return XF_OK;
}

// The same test-set metric without the reference's float accumulators: the
// float `area` stops counting at 2^24 per addend-sum and the base-2, un-negated log-likelihood is not the
// quantity papers quote.  out[0] = mean negative natural-log likelihood (probabilities clamped to
// [1e-15, 1-1e-15]), out[1] = rank AUC with ties counted 1/2 (Mann-Whitney U / (P*N)), integer arithmetic.
XF_DLL int xf_auc_logloss_exact(const int32_t* labels, const float* pctr, uint64_t n, double out[4]) {
  if ((!labels || !pctr) && n) return XF_ERR_ARG;
  if (!out) return XF_ERR_ARG;
  std::vector<Scored> v((size_t)n);
  double ll = 0.0;
  uint64_t positives = 0;
  for (uint64_t i = 0; i < n; ++i) {
    v[i].label = labels[i];
    v[i].pctr = pctr[i];
    const double p = std::min(std::max((double)pctr[i], 1e-15), 1.0 - 1e-15);
    ll -= labels[i] == 1 ? log(p) : log(1.0 - p);
    positives += labels[i] == 1;
  }
  std::sort(v.begin(), v.end(), [](const Scored& a, const Scored& b) { return a.pctr > b.pctr; });
  // walk groups of equal score: negatives in a group see the positives before it plus half of its own
  uint64_t twice_u = 0, pos_before = 0;
  for (size_t i = 0; i < v.size();) {
    size_t j = i;
    uint64_t gp = 0, gn = 0;
    while (j < v.size() && v[j].pctr == v[i].pctr) {
      if (v[j].label == 1) ++gp; else ++gn;
      ++j;
    }
    twice_u += gn * (2 * pos_before + gp);
    pos_before += gp;
    i = j;
  }
  const uint64_t negatives = n - positives;
  out[0] = n ? ll / (double)n : NAN;
  out[1] = (positives && negatives) ? (double)twice_u / (2.0 * (double)positives * (double)negatives) : NAN;
  out[2] = (double)positives;
  out[3] = (double)negatives;
  return XF_OK;
}
