// Compiled and run by tests/test_host.py::test_ps_compat_header: the ps-lite-shaped surface of
// include/xflow/ps_compat.h.  Part 1 needs no GPU (in-process transport with a plain CPU handle); part 2
// drives the device-table handles and must either work (GPU box) or fail loudly (no device).
#include <cmath>
#include <cstdio>
#include <map>

#include "xflow/ps_compat.h"

static int transport() {
  std::map<ps::Key, float> store;
  ps::KVServer<float> server(7);
  server.set_request_handle([&](const ps::KVMeta& m, const ps::KVPairs<float>& req, ps::KVServer<float>* s) {
    ps::KVPairs<float> res;
    if (m.push) {
      for (size_t i = 0; i < req.keys.size(); ++i) store[req.keys[i]] += req.vals[i];
    } else {
      res.keys = req.keys;
      for (ps::Key k : req.keys) res.vals.push_back(store[k]);
    }
    s->Response(m, res);
  });
  ps::KVWorker<float> w(7);
  std::vector<ps::Key> keys{3, 5, 9};
  std::vector<float> g{1, 2, 3}, out{42.f};  // Pull resizes a non-empty vector too
  int cb_calls = 0;
  w.Wait(w.Push(keys, g));
  w.Wait(w.Push(keys, g, {}, 0, [&] { ++cb_calls; }));
  const int ts = w.Pull(keys, &out);
  w.Wait(ts);
  if (!(out.size() == 3 && out[0] == 2 && out[1] == 4 && out[2] == 6 && cb_calls == 1 && ts == 2)) return 1;
  // a handle that never answers, and an app id nobody serves, are errors (not hangs)
  ps::KVServer<float> mute(8);
  mute.set_request_handle([](const ps::KVMeta&, const ps::KVPairs<float>&, ps::KVServer<float>*) {});
  ps::KVWorker<float> w8(8), w9(9);
  int caught = 0;
  try { w8.Pull(keys, &out); } catch (const std::runtime_error&) { ++caught; }
  try { w9.Pull(keys, &out); } catch (const std::runtime_error&) { ++caught; }
  return caught == 2 ? 0 : 2;
}

static int device_handles() {
  xflow::v_dim = 4;
  new xflow::Server(xflow::Optimizer::FTRL, 4);  // the process-wide server the handles resolve to
  ps::KVServer<float> server_w(0), server_v(1);  // server.h:22-31
  server_w.set_request_handle(xflow::FTRL::KVServerFTRLHandle_w());
  server_v.set_request_handle(xflow::FTRL::KVServerFTRLHandle_v());
  ps::KVWorker<float> kv_w(0), kv_v(1);
  std::vector<ps::Key> keys{11, 2000000000000ull, 17};
  std::vector<float> w, v;
  kv_w.Wait(kv_w.Pull(keys, &w));  // insert-on-pull, zeros
  kv_v.Wait(kv_v.Pull(keys, &v));
  if (w.size() != 3 || v.size() != 12 || w[0] != 0.f || w[1] != 0.f || w[2] != 0.f) return 1;
  const std::vector<float> v0 = v;
  std::vector<float> gw{0.5f, -0.25f, 0.f}, gv(12, 0.f);
  gv[5] = 0.125f;
  kv_w.Wait(kv_w.Push(keys, gw));
  kv_v.Wait(kv_v.Push(keys, gv));
  kv_w.Wait(kv_w.Pull(keys, &w));
  kv_v.Wait(kv_v.Pull(keys, &v));
  // first FTRL step from the zero state (ftrl.h:59-74): n = g^2, z = g, w = -(z - sgn(z) l1) / ((beta + |g|) / alpha + l2)
  auto first = [](float g) {
    if (std::fabs(g) <= xflow::lambda1) return 0.f;
    const float z = g, n = g * g;
    const float num = z > 0 ? z - xflow::lambda1 : z + xflow::lambda1;
    return num / -((xflow::beta + std::sqrt(n)) / xflow::alpha + xflow::lambda2);
  };
  for (int i = 0; i < 3; ++i)
    if (std::fabs(w[i] - first(gw[i])) > 1e-6f * std::fabs(first(gw[i])) + 1e-9f) return 2;
  // pushes with the wrong number of values are rejected, like the reference's CHECK_EQ (kv_app.h:414)
  try {
    kv_v.Push(keys, gw);
    return 3;
  } catch (const std::runtime_error&) {
  }
  // coordinates with zero gradient keep their value only up to the FTRL shrinkage; the touched one moved
  if (v[5] == v0[5]) return 4;
  // an SGD functor against this FTRL server is refused
  ps::KVServer<float> other(2);
  other.set_request_handle(xflow::SGD::KVServerSGDHandle_w());
  ps::KVWorker<float> k2(2);
  try {
    k2.Pull(keys, &w);
    return 5;
  } catch (const std::runtime_error&) {
  }
  return 0;
}

int main() {
  const int t = transport();
  if (t) { printf("transport FAILED (%d)\n", t); return 10 + t; }
  puts("transport ok");
  try {
    const int d = device_handles();
    if (d) { printf("device handles FAILED (%d)\n", d); return 20 + d; }
    puts("device handles ok");
  } catch (const std::exception& e) {
    printf("device handles failed loudly: %s\n", e.what());
    return 3;
  }
  return 0;
}
