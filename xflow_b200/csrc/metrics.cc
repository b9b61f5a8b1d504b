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
