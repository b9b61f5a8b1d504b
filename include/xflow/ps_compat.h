// The reference's optimizer plug-in surface, kept for code that is written against ps-lite's KV API:
//
//   ps::Key, ps::KVPairs<V>, ps::KVMeta                       ps-lite/include/ps/kv_app.h:33-42,272-281
//   ps::KVServer<V>::set_request_handle / Response            kv_app.h:307-320,386-403
//   ps::KVWorker<V>::Push / Pull / Wait                       kv_app.h:110-165
//   xflow::FTRL::KVServerFTRLHandle_w / _v                    src/optimizer/ftrl.h:38-85,98-152
//   xflow::SGD::KVServerSGDHandle_w / _v                      src/optimizer/sgd.h:30-65,74-109
//
// so that the reference's wiring (src/model/server.h:22-31)
//     server_w_ = new ps::KVServer<float>(0);  server_w_->set_request_handle(FTRL::KVServerFTRLHandle_w());
//     server_v_ = new ps::KVServer<float>(1);  server_v_->set_request_handle(FTRL::KVServerFTRLHandle_v());
// and its call sites (kv_w_->Wait(kv_w_->Pull(unique_keys, &w)), lr_worker.cc:170) compile and run
// against the GPU-resident table: the handles forward to xf_table_pull / xf_table_push
// (include/xflow_b200.h).  There is no transport: worker and server live in one process (one per GPU),
// Push / Pull call the installed handle synchronously under one mutex per app id — "each KV app's handle
// runs on one receive thread" (ps-lite/src/customer.cc:49-64).
//
// Semantics kept (kv_app.h:110-165, ftrl.h:54-79,112-146, sgd.h:46-59,90-103): keys unique per request,
// values row-major keys x dim, Pull fills the caller's vector before Wait returns (resizing it), Push
// copies its inputs, a missing key is inserted with the optimizer's default contents by Pull or Push, one
// optimizer step per key per Push with the pushed gradient.  Errors surface as exceptions
// (CHECK -> LOG(FATAL) -> throw dmlc::Error in the reference, dmlc/logging.h:183-209).
//
// This is the migration path (one host round trip per call); the fast path is xf_trainer_step_*.
// Header only; link against libxflow_b200.so.
#ifndef XFLOW_PS_COMPAT_H_
#define XFLOW_PS_COMPAT_H_

#include <stdint.h>

#include <functional>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "xflow.h"

namespace ps {

typedef uint64_t Key;  // ps/base.h

template <typename Val>
struct KVPairs {  // kv_app.h:33-42 (SArray there; the handles only need size(), [], assignment)
  std::vector<Key> keys;
  std::vector<Val> vals;
  std::vector<int> lens;
};

struct KVMeta {  // kv_app.h:272-281
  int cmd = 0;
  bool push = false;
  int sender = 0;
  int timestamp = 0;
};

template <typename Val>
class KVServer;

namespace detail {
// app id -> the server object of this process
template <typename Val>
inline std::map<int, KVServer<Val>*>& registry() {
  static std::map<int, KVServer<Val>*> r;
  return r;
}
inline std::mutex& registry_mutex() {
  static std::mutex m;
  return m;
}
}  // namespace detail

template <typename Val>
class KVServer {
 public:
  using ReqHandle = std::function<void(const KVMeta& req_meta, const KVPairs<Val>& req_data, KVServer* server)>;

  explicit KVServer(int app_id) : app_id_(app_id) {
    std::lock_guard<std::mutex> lk(detail::registry_mutex());
    detail::registry<Val>()[app_id] = this;
  }
  ~KVServer() {
    std::lock_guard<std::mutex> lk(detail::registry_mutex());
    auto& r = detail::registry<Val>();
    auto it = r.find(app_id_);
    if (it != r.end() && it->second == this) r.erase(it);
  }
  void set_request_handle(const ReqHandle& request_handle) { handle_ = request_handle; }  // kv_app.h:312-315
  // kv_app.h:320: the handle answers every request exactly once
  void Response(const KVMeta& req, const KVPairs<Val>& res = KVPairs<Val>()) {
    (void)req;
    responded_ = true;
    response_ = res;
  }

  // what the Van + Customer thread do in ps-lite: deliver one request, return its response
  KVPairs<Val> Process(const KVMeta& meta, const KVPairs<Val>& req) {
    std::lock_guard<std::mutex> lk(mu_);
    if (!handle_) throw std::runtime_error("ps::KVServer: no request handle installed for app " + std::to_string(app_id_));
    responded_ = false;
    response_ = KVPairs<Val>();
    handle_(meta, req, this);
    if (!responded_) throw std::runtime_error("ps::KVServer: the request handle did not call Response");
    return response_;
  }

 private:
  int app_id_;
  ReqHandle handle_;
  std::mutex mu_;
  bool responded_ = false;
  KVPairs<Val> response_;
};

template <typename Val>
class KVWorker {
 public:
  using Callback = std::function<void()>;
  explicit KVWorker(int app_id, int customer_id = 0) : app_id_(app_id) { (void)customer_id; }

  // kv_app.h:110-118
  int Push(const std::vector<Key>& keys, const std::vector<Val>& vals, const std::vector<int>& lens = {}, int cmd = 0,
           const Callback& cb = nullptr) {
    KVMeta meta;
    meta.cmd = cmd;
    meta.push = true;
    meta.timestamp = ts_;
    KVPairs<Val> req;
    req.keys = keys;  // Push copies its inputs (sarray.h:138)
    req.vals = vals;
    req.lens = lens;
    server()->Process(meta, req);
    if (cb) cb();
    return ts_++;
  }
  // kv_app.h:147-165: fills *vals (resizing it) before Wait returns
  int Pull(const std::vector<Key>& keys, std::vector<Val>* vals, std::vector<int>* lens = nullptr, int cmd = 0,
           const Callback& cb = nullptr) {
    if (!vals) throw std::invalid_argument("ps::KVWorker::Pull: vals is null");
    KVMeta meta;
    meta.cmd = cmd;
    meta.push = false;
    meta.timestamp = ts_;
    KVPairs<Val> req;
    req.keys = keys;
    KVPairs<Val> res = server()->Process(meta, req);
    *vals = res.vals;
    if (lens) *lens = res.lens;
    if (cb) cb();
    return ts_++;
  }
  void Wait(int timestamp) { (void)timestamp; }  // the calls above are synchronous

 private:
  KVServer<Val>* server() {
    std::lock_guard<std::mutex> lk(detail::registry_mutex());
    auto& r = detail::registry<Val>();
    auto it = r.find(app_id_);
    if (it == r.end()) throw std::runtime_error("ps::KVWorker: no KVServer for app " + std::to_string(app_id_));
    return it->second;
  }
  int app_id_;
  int ts_ = 0;
};

}  // namespace ps

namespace xflow {

namespace detail {
// One handle body for the four functors: app 0 (w) and app 1 (v) of the reference address the scalar
// weight and the latent row of the same key, which here live in one row of one device table.
struct TableHandle {
  xf_table* table;   // null: the process-wide Server's FM-sized table (w and v of a key share a row)
  bool latent;       // false: app 0 (w), true: app 1 (v)
  Optimizer expect;  // the optimizer the functor's name promises

  xf_table* resolve() const {
    if (table) return table;
    Server* s = Server::Get();
    if (s->optimizer() != expect)
      throw std::runtime_error(std::string("xflow: the process-wide Server was created with ") +
                               (s->optimizer() == Optimizer::FTRL ? "FTRL" : "SGD") +
                               ", the installed request handle is for the other optimizer");
    return s->table_fm();
  }

  void operator()(const ps::KVMeta& meta, const ps::KVPairs<float>& req, ps::KVServer<float>* server) const {
    xf_table* t = resolve();
    const uint64_t n = req.keys.size();
    ps::KVPairs<float> res;
    // the row width is the TABLE's latent dimension (a Server(opt, K) or a caller-supplied table need not
    // agree with the global xflow::v_dim), never a guess: both directions are sized and checked with it
    int K = 0;
    if (xf_table_latent_dim(t, &K) != XF_OK) throw std::runtime_error(std::string("xf_table_latent_dim: ") + xf_last_error());
    if (latent && K <= 0) throw std::runtime_error("xflow handle: the latent (app 1) handle needs a table with latent_dim > 0");
    const uint64_t dim = latent ? (uint64_t)K : 1u;
    if (meta.push) {
      // ftrl.h:54-79 / 112-146, sgd.h:46-59 / 90-103: one optimizer step per key with the pushed gradient
      if (req.vals.size() != n * dim) throw std::runtime_error("xflow handle: Push of " + std::to_string(req.vals.size()) +
                                                                " values for " + std::to_string(n) + " keys");
      const int rc = latent ? xf_table_push(t, req.keys.data(), n, nullptr, req.vals.data())
                            : xf_table_push(t, req.keys.data(), n, req.vals.data(), nullptr);
      if (rc != XF_OK) throw std::runtime_error(std::string("xf_table_push: ") + xf_last_error());
    } else {
      // ftrl.h:75-77 / 142-144: res.keys = req.keys, res.vals = keys x dim (missing keys are inserted)
      res.keys = req.keys;
      res.vals.resize(n * dim);
      const int rc = latent ? xf_table_pull(t, req.keys.data(), n, nullptr, res.vals.data())
                            : xf_table_pull(t, req.keys.data(), n, res.vals.data(), nullptr);
      if (rc != XF_OK) throw std::runtime_error(std::string("xf_table_pull: ") + xf_last_error());
    }
    server->Response(meta, res);
  }
};
}  // namespace detail

// src/optimizer/ftrl.h:22-155
class FTRL {
 public:
  struct KVServerFTRLHandle_w : detail::TableHandle {
    explicit KVServerFTRLHandle_w(xf_table* t = nullptr) : detail::TableHandle{t, false, Optimizer::FTRL} {}
  };
  struct KVServerFTRLHandle_v : detail::TableHandle {
    explicit KVServerFTRLHandle_v(xf_table* t = nullptr) : detail::TableHandle{t, true, Optimizer::FTRL} {}
  };
};

// src/optimizer/sgd.h:18-112
class SGD {
 public:
  struct KVServerSGDHandle_w : detail::TableHandle {
    explicit KVServerSGDHandle_w(xf_table* t = nullptr) : detail::TableHandle{t, false, Optimizer::SGD} {}
  };
  struct KVServerSGDHandle_v : detail::TableHandle {
    explicit KVServerSGDHandle_v(xf_table* t = nullptr) : detail::TableHandle{t, true, Optimizer::SGD} {}
  };
};

}  // namespace xflow

#endif  // XFLOW_PS_COMPAT_H_
