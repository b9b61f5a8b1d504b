// TEST INFRASTRUCTURE ONLY — not part of the product.
//
// In-process stand-in for the ps-lite public API (ps-lite/include/ps/kv_app.h,
// ps.h, sarray.h) so that the reference's own src/model + src/optimizer compile
// and run UNMODIFIED inside one process with zero transport.  ps-lite itself is
// not buildable offline (needs zmq.h + protobuf, fetched from the network by
// ps-lite/make/deps.mk:3-27); none of the hot path's arithmetic lives in it.
//
// What is modelled (and nothing else):
//   ps::Key                       = uint64_t               (ps/base.h)
//   ps::KVPairs<V>{keys,vals,lens}                         (kv_app.h:33-42)
//   ps::KVMeta{cmd,push,sender,timestamp}                  (kv_app.h:272-281)
//   ps::KVServer<V>::set_request_handle / Response         (kv_app.h:307-320)
//   ps::KVWorker<V>::Push / Pull / Wait                    (kv_app.h:110-165)
//   ps::MyRank/IsServer/IsWorker/Start/Finalize            (ps.h, base.h)
//   CHECK_EQ                                               (dmlc/logging.h)
// Push/Pull invoke the installed request handle synchronously under one mutex
// per app id, which reproduces "each KV app's handle runs on one receive
// thread" (ps-lite/src/customer.cc:49-64).
#ifndef ORACLE_SHIM_PS_PS_H_
#define ORACLE_SHIM_PS_PS_H_

#include <stdint.h>
#include <stdlib.h>

#include <cmath>
#include <cstdio>
#include <functional>
#include <map>
#include <mutex>
#include <unordered_map>
#include <vector>

#define CHECK_EQ(a, b)                                                        \
  do {                                                                        \
    if (!((a) == (b))) {                                                      \
      fprintf(stderr, "CHECK_EQ failed %s:%d\n", __FILE__, __LINE__);         \
      abort();                                                                \
    }                                                                         \
  } while (0)

namespace ps {

typedef uint64_t Key;

// std::vector is enough: the handles only use size(), operator[], resize()
// and copy-assignment on these members.
template <typename Val>
struct KVPairs {
  std::vector<Key> keys;
  std::vector<Val> vals;
  std::vector<int> lens;
};

struct KVMeta {
  int cmd;
  bool push;
  int sender;
  int timestamp;
};

template <typename Val>
class KVServer;

namespace shim {
struct Registry {
  std::mutex mu;
  std::map<int, void*> servers;  // app id -> KVServer<float>*
  int rank = 0;
  bool is_server = true;
  bool is_worker = true;
  static Registry& Get() {
    static Registry r;
    return r;
  }
};
}  // namespace shim

template <typename Val>
class KVServer {
 public:
  using ReqHandle = std::function<void(const KVMeta& req_meta,
                                       const KVPairs<Val>& req_data,
                                       KVServer* server)>;
  explicit KVServer(int app_id) : app_id_(app_id) {
    std::lock_guard<std::mutex> lk(shim::Registry::Get().mu);
    shim::Registry::Get().servers[app_id] = this;
  }
  void set_request_handle(const ReqHandle& h) { handle_ = h; }
  void Response(const KVMeta& /*req*/, const KVPairs<Val>& res = KVPairs<Val>()) {
    if (pending_vals_ != nullptr) *pending_vals_ = res.vals;
  }
  // used by KVWorker below
  void Serve(const KVMeta& meta, const KVPairs<Val>& req, std::vector<Val>* out) {
    std::lock_guard<std::mutex> lk(mu_);
    pending_vals_ = out;
    handle_(meta, req, this);
    pending_vals_ = nullptr;
  }

 private:
  int app_id_;
  ReqHandle handle_;
  std::mutex mu_;
  std::vector<Val>* pending_vals_ = nullptr;
};

template <typename Val>
class KVWorker {
 public:
  explicit KVWorker(int app_id) : app_id_(app_id) {}
  int Push(const std::vector<Key>& keys, const std::vector<Val>& vals,
           const std::vector<int>& lens = {}, int cmd = 0) {
    KVPairs<Val> req;
    req.keys = keys;
    req.vals = vals;
    req.lens = lens;
    KVMeta m{cmd, true, 0, ts_};
    server()->Serve(m, req, nullptr);
    return ts_++;
  }
  int Pull(const std::vector<Key>& keys, std::vector<Val>* vals,
           std::vector<int>* lens = nullptr, int cmd = 0) {
    (void)lens;
    KVPairs<Val> req;
    req.keys = keys;
    KVMeta m{cmd, false, 0, ts_};
    server()->Serve(m, req, vals);
    return ts_++;
  }
  void Wait(int /*timestamp*/) {}

 private:
  KVServer<Val>* server() {
    auto& r = shim::Registry::Get();
    std::lock_guard<std::mutex> lk(r.mu);
    auto it = r.servers.find(app_id_);
    if (it == r.servers.end()) {
      fprintf(stderr, "ps shim: no server for app %d\n", app_id_);
      abort();
    }
    return reinterpret_cast<KVServer<Val>*>(it->second);
  }
  int app_id_;
  int ts_ = 0;
};

inline int MyRank() { return shim::Registry::Get().rank; }
inline bool IsServer() { return shim::Registry::Get().is_server; }
inline bool IsWorker() { return shim::Registry::Get().is_worker; }
inline void Start() {}
inline void Finalize() {}

}  // namespace ps
#endif  // ORACLE_SHIM_PS_PS_H_
