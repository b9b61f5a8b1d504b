// TEST INFRASTRUCTURE ONLY — not part of the product.
//
// Harness that drives the UNMODIFIED reference sources
//   /root/reference/src/model/lr/lr_worker.cc, fm/fm_worker.cc,
//   src/io/load_data_from_disk.cc, src/optimizer/{ftrl,sgd}.h
// against the in-process ps shim (oracle/shim/ps/ps.h).  Built by
// oracle/Makefile into oracle/_ref/xflow_ref (git-ignored).  It is used to
//   (1) pin the CPU restatement in oracle/xflow_oracle.cc,
//   (2) generate the golden fixtures under tests/golden/ (tests/golden/make_golden.py),
//   (3) serve as the "reference" CPU baseline of bench.py.
//
// Determinism controls:
//   * LRWorker::core_num (= hardware_concurrency(), lr_worker.h:40) and the pool
//     built from it are overwritten after construction with what --core says
//     (interposing std::thread::hardware_concurrency is not possible when
//     libstdc++ is linked statically, as this image's g++ wrapper does);
//   * clock_gettime(CLOCK_REALTIME) can be pinned (--fix-time) so the
//     wall-clock seeded FTRL-v initialisation (base.h:33-44, ftrl.h:114-120)
//     is reproducible;
//   * private members (block_size, v_dim_, the handles' store) are reached with
//     `#define private public` in THIS translation unit only.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/syscall.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <fstream>
#include <functional>
#include <iostream>
#include <memory>
#include <mutex>
#include <random>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

static unsigned g_core_num = 1;
static bool g_fix_time = false;
static double g_fixed_time = 1.5e9;

extern "C" int clock_gettime(clockid_t clk, struct timespec* tp) {
  if (g_fix_time && clk == CLOCK_REALTIME) {
    tp->tv_sec = (time_t)g_fixed_time;
    tp->tv_nsec = 0;
    return 0;
  }
  return (int)syscall(SYS_clock_gettime, clk, tp);
}

#define private public
#include "ps/ps.h"
#include "src/optimizer/ftrl.h"
#include "src/optimizer/sgd.h"
#include "src/model/lr/lr_worker.h"
#include "src/model/fm/fm_worker.h"
#undef private

namespace {

double now_s() {
  struct timespec tp;
  syscall(SYS_clock_gettime, CLOCK_MONOTONIC, &tp);
  return tp.tv_sec + tp.tv_nsec * 1e-9;
}

struct Args {
  std::string model = "lr", opt = "ftrl", train, test, dump, preinit_dump;
  int epochs = 1, core = 1, block_mb = 2, vdim = 10;
  int servers = 1;      // benchmark only: key-range server shards in the shim (see oracle/shim/ps/ps.h)
  int warm_epochs = 0;  // benchmark only: after the timed cold run, time this many more epochs on the warm table
  bool no_predict = false;
};

// all distinct fids of "<prefix>-00000", through the reference's own loader
void collect_keys(const std::string& prefix, std::vector<ps::Key>* keys) {
  char path[1024];
  snprintf(path, sizeof(path), "%s-%05d", prefix.c_str(), 0);
  FILE* f = fopen(path, "r");
  if (!f) return;
  fseek(f, 0, SEEK_END);
  size_t sz = (size_t)ftell(f);
  fclose(f);
  if (sz == 0) return;
  xflow::LoadData ld(path, sz + 16);
  ld.load_minibatch_hash_data_fread();
  for (auto& row : ld.m_data.fea_matrix)
    for (auto& kv : row) keys->push_back(kv.fid);
}

void write_vec(FILE* f, const std::vector<float>& v) {
  if (!v.empty()) fwrite(v.data(), sizeof(float), v.size(), f);
}

}  // namespace

int main(int argc, char** argv) {
  Args a;
  for (int i = 1; i < argc; ++i) {
    std::string s = argv[i];
    auto next = [&]() -> const char* {
      if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", s.c_str()); exit(2); }
      return argv[++i];
    };
    if (s == "--model") a.model = next();
    else if (s == "--opt") a.opt = next();
    else if (s == "--train") a.train = next();
    else if (s == "--test") a.test = next();
    else if (s == "--epochs") a.epochs = atoi(next());
    else if (s == "--core") a.core = atoi(next());
    else if (s == "--block-mb") a.block_mb = atoi(next());
    else if (s == "--vdim") a.vdim = atoi(next());
    else if (s == "--alpha") xflow::alpha = (float)atof(next());
    else if (s == "--beta") xflow::beta = (float)atof(next());
    else if (s == "--l1") xflow::lambda1 = (float)atof(next());
    else if (s == "--l2") xflow::lambda2 = (float)atof(next());
    else if (s == "--lr") xflow::learning_rate = (float)atof(next());
    else if (s == "--dump") a.dump = next();
    else if (s == "--preinit-dump") a.preinit_dump = next();
    else if (s == "--no-predict") a.no_predict = true;
    else if (s == "--servers") a.servers = atoi(next());
    else if (s == "--warm-epochs") a.warm_epochs = atoi(next());
    else if (s == "--fix-time") { g_fix_time = true; g_fixed_time = atof(next()); }
    else { fprintf(stderr, "unknown arg %s\n", s.c_str()); return 2; }
  }
  g_core_num = (unsigned)a.core;
  const bool fm = (a.model == "fm");
  const bool ftrl = (a.opt == "ftrl");
  xflow::v_dim = a.vdim;

  // server side: what src/model/server.h:22-31 does, with the optimizer chosen
  // at run time instead of by commenting lines, and the functors kept
  // addressable so their `store` can be dumped.
  static xflow::FTRL::KVServerFTRLHandle_w ftrl_w;
  static xflow::FTRL::KVServerFTRLHandle_v ftrl_v;
  static xflow::SGD::KVServerSGDHandle_w sgd_w;
  static xflow::SGD::KVServerSGDHandle_v sgd_v;
  ps::KVServer<float>* server_w = new ps::KVServer<float>(0);
  ps::KVServer<float>* server_v = new ps::KVServer<float>(1);
  if (ftrl) {
    server_w->set_request_handle(std::ref(ftrl_w));
    server_v->set_request_handle(std::ref(ftrl_v));
  } else {
    server_w->set_request_handle(std::ref(sgd_w));
    server_v->set_request_handle(std::ref(sgd_v));
  }

  if (a.servers > 1) {
    if (!a.dump.empty() || !a.preinit_dump.empty()) { fprintf(stderr, "--servers > 1 cannot dump\n"); return 2; }
    for (int i = 1; i < a.servers; ++i) {
      if (ftrl) {
        server_w->add_shard_handle(xflow::FTRL::KVServerFTRLHandle_w());
        server_v->add_shard_handle(xflow::FTRL::KVServerFTRLHandle_v());
      } else {
        server_w->add_shard_handle(xflow::SGD::KVServerSGDHandle_w());
        server_v->add_shard_handle(xflow::SGD::KVServerSGDHandle_v());
      }
    }
  }

  std::vector<ps::Key> all_keys;
  if (!a.dump.empty() || !a.preinit_dump.empty()) {
    collect_keys(a.train, &all_keys);
    collect_keys(a.test, &all_keys);
    all_keys.push_back(0);  // the workers' "init push" key (lr_worker.cc:180-182)
    std::sort(all_keys.begin(), all_keys.end());
    all_keys.erase(std::unique(all_keys.begin(), all_keys.end()), all_keys.end());
  }

  // Optional: touch every key once (sorted order) BEFORE training so that the
  // randomly initialised FM latent table can be exported and replayed.
  if (!a.preinit_dump.empty()) {
    ps::KVWorker<float> kw(0), kv(1);
    std::vector<float> w, v;
    kw.Wait(kw.Pull(all_keys, &w));
    if (fm) {
      if (!ftrl) {
        // the SGD v-handle sizes rows from the pushed length (sgd.h:83), so its
        // global v_dim is already a.vdim here.
      }
      kv.Wait(kv.Pull(all_keys, &v));
    }
    FILE* f = fopen(a.preinit_dump.c_str(), "wb");
    uint64_t n = all_keys.size();
    uint32_t K = fm ? (uint32_t)a.vdim : 0, has_nz = 0;
    fwrite("XFTB", 1, 4, f);
    fwrite(&n, 8, 1, f); fwrite(&K, 4, 1, f); fwrite(&has_nz, 4, 1, f);
    fwrite(all_keys.data(), 8, n, f);
    write_vec(f, w);
    write_vec(f, v);
    std::vector<uint8_t> present(n, 1);
    fwrite(present.data(), 1, n, f);
    fclose(f);
  }

  double t_train = 0.0, t_warm = 0.0;
  if (fm) {
    xflow::FMWorker* wk = new xflow::FMWorker(a.train.c_str(), a.test.c_str());
    wk->epochs = a.epochs;
    wk->block_size = a.block_mb;
    wk->v_dim_ = a.vdim;
    wk->core_num = (int)g_core_num;
    wk->pool_ = new xflow::ThreadPool(g_core_num);
    if (a.no_predict) {
      wk->rank = ps::MyRank();
      snprintf(wk->train_data_path, 1024, "%s-%05d", wk->train_file_path, wk->rank);
      double t0 = now_s();
      wk->batch_training(wk->pool_);
      t_train = now_s() - t0;
      if (a.warm_epochs > 0) {  // the same shard again, every key already in the store
        wk->epochs = a.warm_epochs;
        t0 = now_s();
        wk->batch_training(wk->pool_);
        t_warm = now_s() - t0;
      }
    } else {
      double t0 = now_s();
      wk->train();
      t_train = now_s() - t0;
    }
  } else {
    xflow::LRWorker* wk = new xflow::LRWorker(a.train.c_str(), a.test.c_str());
    wk->epochs = a.epochs;
    wk->block_size = a.block_mb;
    wk->core_num = (int)g_core_num;
    wk->pool_ = new xflow::ThreadPool(g_core_num);
    if (a.no_predict) {
      wk->rank = ps::MyRank();
      snprintf(wk->train_data_path, 1024, "%s-%05d", wk->train_file_path, wk->rank);
      double t0 = now_s();
      wk->batch_training(wk->pool_);
      t_train = now_s() - t0;
      if (a.warm_epochs > 0) {  // the same shard again, every key already in the store
        wk->epochs = a.warm_epochs;
        t0 = now_s();
        wk->batch_training(wk->pool_);
        t_warm = now_s() - t0;
      }
    } else {
      double t0 = now_s();
      wk->train();
      t_train = now_s() - t0;
    }
  }
  printf("XFREF train_seconds %.6f\n", t_train);
  if (a.warm_epochs > 0) printf("XFREF warm_seconds %.6f\n", t_warm);

  if (!a.dump.empty()) {
    uint64_t n = all_keys.size();
    uint32_t K = fm ? (uint32_t)a.vdim : 0, has_nz = ftrl ? 1 : 0;
    std::vector<float> w(n), nw(n), zw(n), v((size_t)n * K), nv((size_t)n * K), zv((size_t)n * K);
    std::vector<uint8_t> present(n);
    for (size_t i = 0; i < n; ++i) {
      ps::Key k = all_keys[i];
      if (ftrl) {
        auto it = ftrl_w.store.find(k);
        present[i] = (it != ftrl_w.store.end());
        if (present[i]) { w[i] = it->second.w[0]; nw[i] = it->second.n[0]; zw[i] = it->second.z[0]; }
        if (fm) {
          auto iv = ftrl_v.store.find(k);
          if (iv != ftrl_v.store.end())
            for (uint32_t j = 0; j < K; ++j) {
              v[i * K + j] = iv->second.w[j];
              nv[i * K + j] = iv->second.n[j];
              zv[i * K + j] = iv->second.z[j];
            }
        }
      } else {
        auto it = sgd_w.store.find(k);
        present[i] = (it != sgd_w.store.end());
        if (present[i]) w[i] = it->second.w[0];
        if (fm) {
          auto iv = sgd_v.store.find(k);
          if (iv != sgd_v.store.end())
            for (uint32_t j = 0; j < K; ++j) v[i * K + j] = iv->second.w[j];
        }
      }
    }
    FILE* f = fopen(a.dump.c_str(), "wb");
    fwrite("XFTB", 1, 4, f);
    fwrite(&n, 8, 1, f); fwrite(&K, 4, 1, f); fwrite(&has_nz, 4, 1, f);
    fwrite(all_keys.data(), 8, n, f);
    write_vec(f, w);
    if (has_nz) { write_vec(f, nw); write_vec(f, zw); }
    write_vec(f, v);
    if (has_nz) { write_vec(f, nv); write_vec(f, zv); }
    fwrite(present.data(), 1, n, f);
    fclose(f);
  }
  fflush(stdout);
  _exit(0);  // the reference never joins its pools; skip static destructors
}
