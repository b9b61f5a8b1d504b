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
//
// Benchmark-only extension (bench.py --impl reference): a KVServer may carry SEVERAL
// handle instances ("server shards", add_shard_handle), each with its own mutex and
// its own store, standing for S ps-lite server processes on the same box.  The worker
// then slices every request by key range exactly like KVWorker::DefaultSlicer
// (kv_app.h:405-460) over Postoffice::GetServerKeyRanges (postoffice.cc:134-143) and
// serves the slices one after another from the calling thread; with core_num worker
// threads in flight the shards work in parallel.  With one shard (the default, and
// what every parity fixture uses) nothing changes.
#ifndef ORACLE_SHIM_PS_PS_H_
#define ORACLE_SHIM_PS_PS_H_

#include <stdint.h>
#include <stdlib.h>

#include <cmath>
#include <cstdio>
#include <functional>
#include <algorithm>
#include <map>
#include <memory>
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
  void set_request_handle(const ReqHandle& h) {
    if (shards_.empty()) shards_.emplace_back(new Shard);
    shards_[0]->handle = h;
  }
  void add_shard_handle(const ReqHandle& h) {
    shards_.emplace_back(new Shard);
    shards_.back()->handle = h;
  }
  size_t num_shards() const { return shards_.size(); }
  // the handle answers through here; which shard is being served is thread-local
  void Response(const KVMeta& /*req*/, const KVPairs<Val>& res = KVPairs<Val>()) {
    std::vector<Val>*& pending = Pending();
    if (pending != nullptr) *pending = res.vals;
  }
  // used by KVWorker below
  void Serve(size_t shard, const KVMeta& meta, const KVPairs<Val>& req, std::vector<Val>* out) {
    Shard& sh = *shards_[shard];
    std::lock_guard<std::mutex> lk(sh.mu);
    Pending() = out;
    sh.handle(meta, req, this);
    Pending() = nullptr;
  }

 private:
  struct Shard {
    ReqHandle handle;
    std::mutex mu;
  };
  static std::vector<Val>*& Pending() {
    static thread_local std::vector<Val>* p = nullptr;
    return p;
  }
  int app_id_;
  std::vector<std::unique_ptr<Shard>> shards_;
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
    Dispatch(m, req, nullptr);
    return ts_++;
  }
  int Pull(const std::vector<Key>& keys, std::vector<Val>* vals,
           std::vector<int>* lens = nullptr, int cmd = 0) {
    (void)lens;
    KVPairs<Val> req;
    req.keys = keys;
    KVMeta m{cmd, false, 0, ts_};
    Dispatch(m, req, vals);
    return ts_++;
  }
  void Wait(int /*timestamp*/) {}

 private:
  void Dispatch(const KVMeta& m, const KVPairs<Val>& req, std::vector<Val>* out) {
    KVServer<Val>* sv = server();
    const size_t S = sv->num_shards();
    if (S <= 1) {
      sv->Serve(0, m, req, out);
      return;
    }
    // DefaultSlicer: keys are sorted; server i owns [width*i, width*(i+1)), width = floor((2^64-1)/S)
    const Key width = (Key)0xFFFFFFFFFFFFFFFFull / (Key)S;
    const size_t n = req.keys.size();
    const size_t dim = (n && !req.vals.empty()) ? req.vals.size() / n : 0;
    if (out) out->clear();
    size_t beg = 0;
    for (size_t i = 0; i < S && beg < n; ++i) {
      size_t end = n;
      if (i + 1 < S) end = std::lower_bound(req.keys.begin() + beg, req.keys.end(), width * (Key)(i + 1)) - req.keys.begin();
      if (end == beg) continue;
      KVPairs<Val> part;
      part.keys.assign(req.keys.begin() + beg, req.keys.begin() + end);
      if (dim) part.vals.assign(req.vals.begin() + beg * dim, req.vals.begin() + end * dim);
      std::vector<Val> got;
      sv->Serve(i, m, part, out ? &got : nullptr);
      if (out) out->insert(out->end(), got.begin(), got.end());
      beg = end;
    }
  }
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
