// C ABI layer 1 and the C++ model/optimizer surface (include/xflow/xflow.h): the reference's
// LRWorker / FMWorker / Server call flow, re-hosted on the device table and the fused step.
// Host orchestration only; all arithmetic of the hot path runs in kernels.cu.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <iostream>
#include <mutex>
#include <stdexcept>

#include "../../include/xflow/xflow.h"

void xf_set_error(const char* fmt, ...);

namespace xflow {

int w_dim = 1;                // ftrl.h:15
int v_dim = 10;               // ftrl.h:16
float alpha = 5e-2;           // ftrl.h:17
float beta = 1.0;             // ftrl.h:18
float lambda1 = 5e-5;         // ftrl.h:19
float lambda2 = 10.0;         // ftrl.h:20
float learning_rate = 0.001;  // sgd.h:16

namespace {
std::mutex g_mu;
Server* g_server = nullptr;

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

// errors from the C ABI surface as exceptions inside the C++ façade (the reference's CHECK ->
// LOG(FATAL) -> throw dmlc::Error convention, dmlc/logging.h:183-209); XF* entry points catch them.
void must(int rc, const char* what) {
  if (rc != XF_OK) throw std::runtime_error(std::string(what) + ": " + xf_last_error());
}
}  // namespace

int MyRank() { return env_int("XFLOW_RANK", env_int("RANK", 0)); }
int NumWorkers() { return env_int("XFLOW_WORLD", env_int("WORLD_SIZE", 1)); }

// One process per GPU (the reference: one ps-lite worker + one server process each, local.sh).  With
// XFLOW_WORLD / WORLD_SIZE = N > 1 this process is worker `rank` of N AND the server of key range `rank`
// (postoffice.cc:134-143); the N processes find each other through a file (XFLOW_COMM_FILE).
static int LocalDevice() {
  const int d = env_int("XFLOW_DEVICE", env_int("LOCAL_RANK", -1));
  if (d >= 0) return d;
  const int n = xf_device_count();
  return n > 0 ? MyRank() % n : 0;
}
static std::string CommFile() {
  const char* f = getenv("XFLOW_COMM_FILE");
  if (f && *f) return f;
  const char* port = getenv("MASTER_PORT");
  return std::string("/tmp/xflow_b200_comm_") + (port && *port ? port : "default") + ".id";
}

// ------------------------------------------------------------------------------------------------
// Server  (src/model/server.h:20-35)
// ------------------------------------------------------------------------------------------------
Server::Server(Optimizer opt, int latent_dim, int device)
    : opt_(opt), latent_dim_(latent_dim), device_(device < 0 ? LocalDevice() : device) {
  rank_ = MyRank();
  world_ = NumWorkers();
  if (world_ < 1) world_ = 1;
  if (rank_ < 0 || rank_ >= world_)
    throw std::runtime_error("rank " + std::to_string(rank_) + " needs XFLOW_WORLD / WORLD_SIZE > rank: a worker with rank > 0 "
                             "has no servers to talk to on its own");
  if (world_ > 1) must(xf_comm_create_from_file(&comm_, CommFile().c_str(), rank_, world_, device_), "xf_comm_create_from_file");
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_server) g_server = this;
  std::cout << "init server success " << std::endl;  // server.h:30
}

Server::~Server() {
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_server == this) g_server = nullptr;
  }
  if (lr_) xf_table_destroy(lr_);
  if (fm_) xf_table_destroy(fm_);
  if (comm_) xf_comm_destroy(comm_);
}

Server* Server::Get() {
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_server) return g_server;
  }
  Optimizer opt = Optimizer::FTRL;
  const char* o = getenv("XFLOW_OPTIMIZER");
  if (o && (strcmp(o, "sgd") == 0 || strcmp(o, "SGD") == 0 || strcmp(o, "1") == 0)) opt = Optimizer::SGD;
  return new Server(opt);
}

static xf_table* make_table(Optimizer opt, int K, int device, int rank, int world) {
  xf_table_config cfg;
  xf_table_config_default(&cfg);
  cfg.device = device;
  cfg.shard_index = rank;
  cfg.num_shards = world;
  cfg.latent_dim = K;
  cfg.optimizer = (opt == Optimizer::FTRL) ? XF_OPTIMIZER_FTRL : XF_OPTIMIZER_SGD;
  cfg.alpha = alpha; cfg.beta = beta; cfg.lambda1 = lambda1; cfg.lambda2 = lambda2;
  cfg.learning_rate = learning_rate;
  cfg.seed = (uint64_t)env_int("XFLOW_SEED", 0);
  cfg.capacity = (uint64_t)1 << env_int("XFLOW_TABLE_LOG2", 20);
  xf_table* t = nullptr;
  must(xf_table_create(&t, &cfg), "xf_table_create");
  return t;
}

xf_table* Server::table_lr() {
  if (!lr_) lr_ = make_table(opt_, 0, device_, rank_, world_);
  return lr_;
}
xf_table* Server::table_fm() {
  if (!fm_) fm_ = make_table(opt_, latent_dim_ > 0 ? latent_dim_ : v_dim, device_, rank_, world_);
  return fm_;
}

// ------------------------------------------------------------------------------------------------
// workers
// ------------------------------------------------------------------------------------------------
WorkerBase::WorkerBase(const char* train_file, const char* test_file, int model)
    : model_(model), train_file_path(train_file ? train_file : ""), test_file_path(test_file ? test_file : "") {
  // the reference keeps the caller's pointers (lr_worker.h:81-82); we copy the strings
  core_num = env_int("XFLOW_CORE_NUM", 1);
  if (core_num < 1) core_num = 1;
  block_size = env_int("XFLOW_BLOCK_MB", 2);
  test_block_size = (model == XF_MODEL_LR) ? 4 : 2;
  Server* s = Server::Get();
  table_ = (model == XF_MODEL_LR) ? s->table_lr() : s->table_fm();
  comm_ = s->comm();
  if (comm_) core_num = 1;  // a sharded step is collective: one slice per block on every rank
  train_data_path[0] = test_data_path[0] = '\0';
}

WorkerBase::~WorkerBase() {
  if (trainer_) xf_trainer_destroy(trainer_);
  if (loader_) xf_loader_close(loader_);
}

LRWorker::LRWorker(const char* train_file, const char* test_file) : WorkerBase(train_file, test_file, XF_MODEL_LR) {}
FMWorker::FMWorker(const char* train_file, const char* test_file) : WorkerBase(train_file, test_file, XF_MODEL_FM) {}

void WorkerBase::ensure_trainer(uint32_t rows, uint32_t nnz) {
  if (trainer_ && rows <= trainer_rows_ && nnz <= trainer_nnz_) return;
  if (trainer_ && comm_)
    throw std::runtime_error("sharded worker: a block exceeds the trainer's limits (all ranks size their trainer from the "
                             "block size; rows without features can break that bound)");
  if (trainer_) {
    must(xf_trainer_sync(trainer_), "xf_trainer_sync");
    xf_trainer_destroy(trainer_);
    trainer_ = nullptr;
  }
  xf_trainer_config cfg;
  cfg.model = model_;
  cfg.max_rows = rows + rows / 4 + 16;
  cfg.max_nnz = nnz + nnz / 4 + 16;
  cfg.keep_loss = 0;
  must(xf_trainer_create(&trainer_, table_, comm_, &cfg), "xf_trainer_create");
  trainer_rows_ = cfg.max_rows;
  trainer_nnz_ = cfg.max_nnz;
}

// rows [start,end) of the current block, as one fused device step
void WorkerBase::update(int start, int end) {
  if (end <= start) return;
  const uint32_t base = cur_row_ptr_[start];
  const uint32_t rows = (uint32_t)(end - start);
  const uint32_t nnz = cur_row_ptr_[end] - base;
  const uint32_t* rp = cur_row_ptr_ + start;
  if (base != 0) {
    // slice offsets must start at 0 for the step's CSR view
    slice_row_ptr_.resize(rows + 1);
    for (uint32_t i = 0; i <= rows; ++i) slice_row_ptr_[i] = cur_row_ptr_[start + i] - base;
    rp = slice_row_ptr_.data();
  }
  ensure_trainer(rows, nnz);
  must(xf_trainer_step_host(trainer_, rp, cur_keys_ + base, cur_labels_ + start, rows, nnz, nullptr),
       "xf_trainer_step_host");
  rows_trained += rows;
}

// a well-formed text block of `bytes` bytes holds at most bytes/8 rows ("0\ta:b:c\n") and bytes/6 tokens
// ("a:b:c ")
void WorkerBase::ensure_trainer_for_block(uint64_t bytes) {
  // sharded: one trainer for the whole run (its exchange buffers are mapped by the peers): big enough for the
  // training blocks AND the prediction blocks
  if (comm_) bytes = std::max<uint64_t>(bytes, (uint64_t)std::max(block_size, test_block_size) << 20);
  ensure_trainer((uint32_t)(bytes / 8 + 2), (uint32_t)(bytes / 6 + 2));
}

// parse one raw block on the device; a block with rows that carry no features ("0\n") can hold more rows
// than ensure_trainer_for_block assumed: size the trainer for the absolute worst case and try once more
void WorkerBase::ingest_block(const char* text, uint64_t len, uint32_t* rows, uint32_t* nnz) {
  int rc = xf_trainer_ingest_text(trainer_, text, len, rows, nnz);
  if (rc == XF_ERR_ARG && !comm_ && (trainer_rows_ < len / 2 + 2 || trainer_nnz_ < len / 4 + 2)) {
    ensure_trainer((uint32_t)(len / 2 + 2), (uint32_t)(len / 4 + 2));
    rc = xf_trainer_ingest_text(trainer_, text, len, rows, nnz);
  }
  must(rc, "xf_trainer_ingest_text");
}

// open (or rewind) the loader of `path`; one loader serves every epoch
void WorkerBase::open_loader(const char* path, uint64_t block_bytes) {
  if (loader_ && loader_path_ == path && loader_block_ == block_bytes && xf_loader_rewind(loader_) == XF_OK) return;
  if (loader_) { xf_loader_close(loader_); loader_ = nullptr; }
  must(xf_loader_open(&loader_, path, block_bytes), "xf_loader_open");
  loader_path_ = path;
  loader_block_ = block_bytes;
}

// number of blocks the loader will form from `path` (one pass over the file, no parsing)
uint64_t WorkerBase::count_blocks(const char* path, uint64_t block_bytes) {
  open_loader(path, block_bytes);
  uint64_t n = 0;
  for (;;) {
    const char* text = nullptr;
    uint64_t len = 0;
    must(xf_loader_next_raw(loader_, &text, &len), "xf_loader_next_raw");
    if (len == 0) break;
    ++n;
  }
  return n;
}

// The device-parser loop shared by training and prediction.  Pipeline per block i:
//   host   read block i+1 from the file            (xf_loader_next_raw, second text buffer)
//   device H2D + parse of block i+1                (xf_trainer_ingest_begin, ingest stream)
//   device step / forward pass of block i          (table stream)
// all three overlap; the host only waits for a parse (xf_trainer_ingest_end), never for a step.
// Sharded: every rank runs `collective_blocks` iterations (the maximum over the ranks); a rank whose file
// has ended keeps taking part with empty blocks.
void WorkerBase::run_blocks(uint64_t collective_blocks, const std::function<void(uint32_t rows)>& on_block) {
  const char* text = nullptr;
  uint64_t len = 0;
  must(xf_loader_next_raw(loader_, &text, &len), "xf_loader_next_raw");
  bool pending = false;
  if (len) { must(xf_trainer_ingest_begin(trainer_, text, len), "xf_trainer_ingest_begin"); pending = true; }
  for (uint64_t blk = 0; comm_ ? blk < collective_blocks : pending; ++blk) {
    // read the following block while the device parses this one
    const char* next_text = nullptr;
    uint64_t next_len = 0;
    if (pending) must(xf_loader_next_raw(loader_, &next_text, &next_len), "xf_loader_next_raw");
    uint32_t rows = 0, nnz = 0;
    if (pending) {
      int rc = xf_trainer_ingest_end(trainer_, &rows, &nnz);
      if (rc == XF_ERR_ARG && !comm_ && (trainer_rows_ < len / 2 + 2 || trainer_nnz_ < len / 4 + 2)) {
        // rows without features: more rows than a well-formed block can hold; re-size and parse again
        must(xf_trainer_sync(trainer_), "xf_trainer_sync");
        ensure_trainer((uint32_t)(len / 2 + 2), (uint32_t)(len / 4 + 2));
        rc = xf_trainer_ingest_text(trainer_, text, len, &rows, &nnz);
      }
      must(rc, "xf_trainer_ingest_end");
    } else {
      // sharded, file exhausted: an empty block keeps this rank in the collective step
      must(xf_trainer_ingest_text(trainer_, "", 0, &rows, &nnz), "xf_trainer_ingest_text");
    }
    if (rows == 0 && !comm_) break;  // lr_worker.cc:189
    on_block(rows);
    text = next_text;
    len = next_len;
    pending = next_len != 0;
    if (pending) must(xf_trainer_ingest_begin(trainer_, text, len), "xf_trainer_ingest_begin");
  }
}

void WorkerBase::batch_training() {
  const bool host_parse = env_int("XFLOW_HOST_PARSE", 0) != 0 && !comm_;
  const uint64_t block_bytes = (uint64_t)block_size << 20;
  if (host_parse) ensure_trainer(1024, 65536);
  else ensure_trainer_for_block(block_bytes);
  must(xf_trainer_init_push(trainer_), "xf_trainer_init_push");  // lr_worker.cc:180-182
  uint64_t collective_blocks = 0;
  if (comm_) {
    collective_blocks = count_blocks(train_data_path, block_bytes);
    must(xf_comm_allreduce_max(comm_, &collective_blocks), "xf_comm_allreduce_max");
  }
  for (int epoch = 0; epoch < epochs; ++epoch) {
    open_loader(train_data_path, block_bytes);  // :184 (the reference re-opens the file every epoch)
    if (!host_parse) {
      // default path: the host only forms the block; parsing, hashing and the step run on the device
      run_blocks(collective_blocks, [&](uint32_t rows) {
        const uint32_t thread_size = rows / (uint32_t)core_num;  // :190 — remainder rows are dropped, as in the reference
        for (uint32_t i = 0; i < (uint32_t)core_num; ++i) {      // :192-196
          must(xf_trainer_step_ingested(trainer_, i * thread_size, (i + 1) * thread_size), "xf_trainer_step_ingested");
          rows_trained += thread_size;
        }
      });
    }
    while (host_parse) {
      // the loader alternates two output sets; before it overwrites one, the copies that read it
      // must have drained
      if (trainer_) must(xf_trainer_wait_uploads(trainer_), "xf_trainer_wait_uploads");
      uint32_t rows = 0, nnz = 0;
      must(xf_loader_next(loader_, &rows, &nnz), "xf_loader_next");
      if (rows == 0) break;  // :189
      must(xf_loader_batch(loader_, &cur_row_ptr_, &cur_keys_, &cur_labels_), "xf_loader_batch");
      const int thread_size = (int)rows / core_num;  // :190 — remainder rows are dropped, as in the reference
      for (int i = 0; i < core_num; ++i) update(i * thread_size, (i + 1) * thread_size);  // :192-196
    }
    must(xf_trainer_sync(trainer_), "xf_trainer_sync");
    if ((epoch + 1) % 30 == 0) std::cout << "epoch : " << epoch << std::endl;  // :202
  }
  cur_row_ptr_ = nullptr;
  cur_keys_ = nullptr;
  cur_labels_ = nullptr;
}

void WorkerBase::calculate_pctr(int start, int end) {
  if (end <= start) return;
  const uint32_t base = cur_row_ptr_[start];
  const uint32_t rows = (uint32_t)(end - start);
  const uint32_t nnz = cur_row_ptr_[end] - base;
  const uint32_t* rp = cur_row_ptr_ + start;
  if (base != 0) {
    slice_row_ptr_.resize(rows + 1);
    for (uint32_t i = 0; i <= rows; ++i) slice_row_ptr_[i] = cur_row_ptr_[start + i] - base;
    rp = slice_row_ptr_.data();
  }
  ensure_trainer(rows, nnz);
  std::vector<float> pctr(rows);
  must(xf_trainer_predict_host(trainer_, rp, cur_keys_ + base, rows, nnz, pctr.data()), "xf_trainer_predict_host");
  for (uint32_t i = 0; i < rows; ++i) {
    auc_key ak;
    ak.label = cur_labels_[start + i];
    ak.pctr = pctr[i];
    test_auc_vec.push_back(ak);
    md << pctr[i] << "\t" << 1 - ak.label << "\t" << ak.label << std::endl;  // lr_worker.cc:67
  }
}

// rank 0 predicts on <test>-00000 (lr_worker.cc:213-216).  Sharded: the forward pass needs the owners of
// the keys, so every rank runs the same number of collective forward steps; ranks > 0 feed empty blocks and
// write nothing.
void WorkerBase::predict(int rank_arg, int block) {
  const bool feeding = (rank_arg == 0);
  char buffer[1024];
  snprintf(buffer, 1024, "%d_%d", rank_arg, block);
  std::string filename = buffer;
  if (feeding) {
    md.open("pred_" + filename + ".txt");  // lr_worker.cc:77
    if (!md.is_open()) std::cout << "open pred file failure!" << std::endl;
  }
  snprintf(test_data_path, 1024, "%s-%05d", test_file_path.c_str(), rank_arg);
  const uint64_t block_bytes = (uint64_t)test_block_size << 20;
  test_auc_vec.clear();
  const bool host_parse = env_int("XFLOW_HOST_PARSE", 0) != 0 && !comm_;
  uint64_t collective_blocks = 0;
  if (comm_) {
    collective_blocks = feeding ? count_blocks(test_data_path, block_bytes) : 0;
    must(xf_comm_allreduce_max(comm_, &collective_blocks), "xf_comm_allreduce_max");
  }
  if (!host_parse) ensure_trainer_for_block(block_bytes);
  std::vector<float> pctr_buf;
  std::vector<uint8_t> label_buf;
  // XFLOW_DEVICE_METRIC=1: the metric is computed on the device (radix sort + integer rank sums, metric.cu) and the
  // printed numbers come from there.  Default: the host restatement of Base::calculate_auc, whose float
  // accumulators and std::sort tie order are what the reference's own printout (and the golden fixtures) contain;
  // the predictions come back to the host either way, the reference writes every one to pred_<rank>_<block>.txt.
  xf_metric* metric = nullptr;
  if (!host_parse && env_int("XFLOW_DEVICE_METRIC", 0) != 0)
    must(xf_metric_create(&metric, Server::Get()->device()), "xf_metric_create");
  if (!host_parse) {
    if (feeding) open_loader(test_data_path, block_bytes);
    else open_loader("/dev/null", block_bytes);  // nothing to feed: every block is empty
    run_blocks(collective_blocks, [&](uint32_t rows) {
      const uint32_t thread_size = rows / (uint32_t)core_num;
      pctr_buf.resize(thread_size + 1);
      label_buf.resize(thread_size + 1);
      for (uint32_t i = 0; i < (uint32_t)core_num; ++i) {
        if (metric)
          must(xf_trainer_predict_ingested_metric(trainer_, i * thread_size, (i + 1) * thread_size, metric,
                                                  feeding ? pctr_buf.data() : nullptr, feeding ? label_buf.data() : nullptr),
               "xf_trainer_predict_ingested_metric");
        else
          must(xf_trainer_predict_ingested(trainer_, i * thread_size, (i + 1) * thread_size, pctr_buf.data(), label_buf.data()),
               "xf_trainer_predict_ingested");
        if (!feeding) continue;
        for (uint32_t r = 0; r < thread_size; ++r) {
          auc_key ak;
          ak.label = label_buf[r];
          ak.pctr = pctr_buf[r];
          test_auc_vec.push_back(ak);
          md << pctr_buf[r] << "\t" << 1 - ak.label << "\t" << ak.label << "\n";  // lr_worker.cc:67
        }
      }
    });
  }
  if (host_parse) open_loader(test_data_path, block_bytes);
  while (host_parse) {
    uint32_t rows = 0, nnz = 0;
    must(xf_loader_next(loader_, &rows, &nnz), "xf_loader_next");
    if (rows == 0) break;
    must(xf_loader_batch(loader_, &cur_row_ptr_, &cur_keys_, &cur_labels_), "xf_loader_batch");
    const int thread_size = (int)rows / core_num;
    for (int i = 0; i < core_num; ++i) calculate_pctr(i * thread_size, (i + 1) * thread_size);
  }
  if (md.is_open()) md.close();
  cur_row_ptr_ = nullptr;
  cur_keys_ = nullptr;
  cur_labels_ = nullptr;
  double dm[6] = {0, 0, 0, 0, 0, 0};
  if (metric) {
    if (feeding) must(xf_metric_finish(metric, nullptr, dm), "xf_metric_finish");
    xf_metric_destroy(metric);
  }
  if (!feeding) return;
  if (metric) {
    // Base::calculate_auc's printout (base.h:101-109) from the device-side metric
    last_logloss = dm[0];
    last_auc = dm[1];
    std::cout << "logloss: " << (float)dm[0] << "\t";
    if (dm[2] == 0 || dm[3] == 0) {
      std::cout << "tp_n = " << (int)dm[2] << std::endl;
    } else {
      std::cout << "auc = " << (float)dm[1] << "\ttp = " << (int)dm[2] << " fp = " << (size_t)dm[3] << std::endl;
    }
    if (env_int("XFLOW_EXACT_METRIC", 0))
      std::cout << "exact: logloss(ln) = " << dm[4] << "\tauc = " << dm[5] << std::endl;
    return;
  }

  // Base::calculate_auc (base.h:84-110), same printout
  std::vector<int32_t> labels(test_auc_vec.size());
  std::vector<float> pctr(test_auc_vec.size());
  for (size_t i = 0; i < test_auc_vec.size(); ++i) {
    labels[i] = test_auc_vec[i].label;
    pctr[i] = test_auc_vec[i].pctr;
  }
  double m[4] = {0, 0, 0, 0};
  xf_auc_logloss(labels.data(), pctr.data(), labels.size(), m);
  last_logloss = m[0];
  last_auc = m[1];
  std::cout << "logloss: " << (float)m[0] << "\t";
  if (m[2] == 0 || m[3] == 0) {
    std::cout << "tp_n = " << (int)m[2] << std::endl;
  } else {
    std::cout << "auc = " << (float)m[1] << "\ttp = " << (int)m[2] << " fp = " << (size_t)m[3] << std::endl;
  }
  if (env_int("XFLOW_EXACT_METRIC", 0)) {
    // the same test set in exact arithmetic (natural-log logloss, tie-aware AUC); off by default so that
    // stdout stays the reference's
    double x[4] = {0, 0, 0, 0};
    xf_auc_logloss_exact(labels.data(), pctr.data(), labels.size(), x);
    std::cout << "exact: logloss(ln) = " << x[0] << "\tauc = " << x[1] << std::endl;
  }
}

void WorkerBase::train() {
  rank = MyRank();
  std::cout << "my rank is = " << rank << std::endl;
  snprintf(train_data_path, 1024, "%s-%05d", train_file_path.c_str(), rank);
  batch_training();
  if (rank == 0) {
    std::cout << model_name() << " AUC: " << std::endl;
    predict(rank, 0);
  } else if (comm_) {
    predict(rank, 0);  // takes part in rank 0's collective forward steps; feeds and prints nothing
  }
  std::cout << "train end......" << std::endl;
}

}  // namespace xflow

// ------------------------------------------------------------------------------------------------
// reference C API  (src/c_api/c_api.h:26-41, c_api.cc:10-20)
// ------------------------------------------------------------------------------------------------
namespace {
struct XFlowHandle {  // the reference's `class XFlow { LRWorker* lr_worker_; }`
  xflow::WorkerBase* worker = nullptr;
};

int xf_guard(const char* what, const std::function<void()>& fn) {
  try {
    fn();
    return XF_OK;
  } catch (const std::exception& e) {
    xf_set_error("%s: %s", what, e.what());
    return XF_ERR_STATE;
  } catch (...) {
    xf_set_error("%s: unknown exception", what);
    return XF_ERR_STATE;
  }
}
}  // namespace

XF_DLL int XFCreateEx(void** h, const char* train_path, const char* test_path, int model, int optimizer,
                      int latent_dim, int epochs) {
  if (!h || !train_path || !test_path) { xf_set_error("null argument"); return XF_ERR_ARG; }
  return xf_guard("XFCreate", [&]() {
    if (latent_dim > 0) xflow::v_dim = latent_dim;
    if (optimizer >= 0) {
      // make sure a server with the requested optimizer exists before the worker attaches
      bool need;
      {
        std::lock_guard<std::mutex> lk(xflow::g_mu);
        need = (xflow::g_server == nullptr);
      }
      if (need) new xflow::Server(optimizer == XF_OPTIMIZER_SGD ? xflow::Optimizer::SGD : xflow::Optimizer::FTRL);
    }
    XFlowHandle* xf = new XFlowHandle;
    if (model == XF_MODEL_FM) xf->worker = new xflow::FMWorker(train_path, test_path);
    else xf->worker = new xflow::LRWorker(train_path, test_path);
    if (epochs > 0) xf->worker->epochs = epochs;
    *h = xf;
  });
}

XF_DLL int XFCreate(void** h, const char* train_path, const char* test_path) {
  const char* e = getenv("XFLOW_EPOCHS");
  return XFCreateEx(h, train_path, test_path, XF_MODEL_LR, -1, 0, (e && *e) ? atoi(e) : 0);
}

XF_DLL int XFStartTrain(void** h) {
  if (!h || !*h) { xf_set_error("null handle"); return XF_ERR_ARG; }
  XFlowHandle* xf = reinterpret_cast<XFlowHandle*>(*h);
  return xf_guard("XFStartTrain", [&]() { xf->worker->train(); });
}

XF_DLL int XFDestroy(void** h) {
  if (!h || !*h) return XF_OK;
  XFlowHandle* xf = reinterpret_cast<XFlowHandle*>(*h);
  delete xf->worker;
  delete xf;
  *h = nullptr;
  return XF_OK;
}
