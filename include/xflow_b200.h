/*
 * xflow_b200 — C ABI of the B200-native drop-in for xflow's data-parallel hot path.
 *
 * Plain C: pointers, sizes and POD structs only (no torch / STL types).  Every entry point below is
 * what the reference's FFI for this path binds; the comment on each names the reference interface it
 * replaces (paths relative to the xswang/xflow tree).  All functions return 0 on success and a
 * negative code on failure (never throw across the ABI); xf_last_error() describes the last failure
 * of the calling thread.  There is NO CPU fallback: without a CUDA device every compute entry point
 * fails with XF_ERR_CUDA.
 *
 * Layers
 *   1. reference C API        XFCreate / XFStartTrain                    (src/c_api/c_api.h:26-29)
 *   2. parameter table        xf_table_*   = KVServer + FTRL/SGD handle  (src/model/server.h:20-35,
 *                                            src/optimizer/ftrl.h, sgd.h ; ps-lite kv_app.h:110-165)
 *   3. fused worker step      xf_trainer_* = LRWorker/FMWorker::update + predict
 *                                            (src/model/lr/lr_worker.cc:25-177, fm/fm_worker.cc:25-245)
 *   4. host ingest            xf_loader_*, xf_hash_* = LoadData::load_minibatch_hash_data_fread
 *                                            (src/io/load_data_from_disk.cc:103-210)
 *   5. multi-GPU exchange     xf_comm_*    = KVWorker slicing + Van transport
 *                                            (ps-lite kv_app.h:405-460, postoffice.cc:134-143)
 */
#ifndef XFLOW_B200_H_
#define XFLOW_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
#define XF_DLL extern "C" __attribute__((visibility("default")))
#else
#define XF_DLL __attribute__((visibility("default")))
#endif

enum {
  XF_OK = 0,
  XF_ERR_ARG = -1,       /* bad argument */
  XF_ERR_CUDA = -2,      /* CUDA runtime error (incl. "no device") */
  XF_ERR_FULL = -3,      /* table probe overflow and growth impossible */
  XF_ERR_IO = -4,        /* file open / read failure (reference: exit(1), io.h:33-36) */
  XF_ERR_COMM = -5,      /* NCCL failure */
  XF_ERR_STATE = -6
};

enum { XF_MODEL_LR = 0, XF_MODEL_FM = 1,              /* main.cc:26-39: '0' = LR, '1' = FM */
       XF_MODEL_FM_CANONICAL = 2,                     /* NOT the reference's model: the textbook FM, see below */
       XF_MODEL_MVM = 3 };                            /* a DEFINED multi-view machine (mvm_worker.cc is not), see below */
enum { XF_OPTIMIZER_FTRL = 0, XF_OPTIMIZER_SGD = 1 }; /* server.h:24-29 (comment toggle in the reference) */
enum {
  XF_VINIT_DEFAULT = 0,  /* FTRL: N(0,1)*1e-2 (ftrl.h:114-120, counter-based here); SGD: 0.001 (sgd.h:68-70) */
  XF_VINIT_COUNTER = 1,  /* counter-based N(0,1)*1e-2 keyed by (key,k,seed) for either optimizer */
  XF_VINIT_ZERO = 3      /* zeros (used before xf_table_import of a replayed table) */
};

XF_DLL const char* xf_last_error(void);
XF_DLL int xf_version(void);
/* number of CUDA devices visible (0 without a GPU; never fails) */
XF_DLL int xf_device_count(void);

/* ------------------------------------------------------------------------------------------------
 * 2. Parameter table
 * ---------------------------------------------------------------------------------------------- */
typedef struct xf_table xf_table;

typedef struct xf_table_config {
  int device;            /* CUDA ordinal */
  int latent_dim;        /* K: 0 = LR (app 0 only); >0 = FM (apps 0 and 1).  ftrl.h:16 / fm_worker.h:92 default 10 */
  int optimizer;         /* XF_OPTIMIZER_* */
  float alpha;           /* ftrl.h:17  default 5e-2 */
  float beta;            /* ftrl.h:18  default 1.0  */
  float lambda1;         /* ftrl.h:19  default 5e-5 */
  float lambda2;         /* ftrl.h:20  default 10.0 */
  float learning_rate;   /* sgd.h:16   default 1e-3 */
  int v_init;            /* XF_VINIT_* */
  uint64_t seed;
  uint64_t capacity;     /* initial slot count (rounded up to a power of two); 0 = 1<<20.  Grows on demand. */
  int shard_index;       /* this table owns keys of shard_index out of num_shards (postoffice.cc:134-143) */
  int num_shards;        /* 1 = whole key space */
  int canonical_fm;      /* 1: rows carry the accumulators of XF_MODEL_FM_CANONICAL / XF_MODEL_MVM (latent_dim in {4,8,16,32,64,128}) */
} xf_table_config;

/* fills *cfg with the reference's compile-time defaults (ftrl.h:15-20, sgd.h:16) */
XF_DLL int xf_table_config_default(xf_table_config* cfg);

/* replaces: new ps::KVServer<float>(0/1) + set_request_handle(FTRL/SGD handle)  (server.h:22-31) */
XF_DLL int xf_table_create(xf_table** out, const xf_table_config* cfg);
XF_DLL int xf_table_destroy(xf_table* t);

/* replaces: KVWorker<float>::Pull + Wait (kv_app.h:147-165) served by KVServerFTRLHandle_w/_v pull
 * branch (ftrl.h:49-52,75-77,108-111,142-144 ; sgd.h).  keys: n host u64 (any order, duplicates allowed
 * for pulls).  w_out: n floats or NULL.  v_out: n*K floats (row-major) or NULL.  Missing keys are
 * inserted with the optimizer's default contents, as `store[key]` does. */
XF_DLL int xf_table_pull(xf_table* t, const uint64_t* keys, uint64_t n, float* w_out, float* v_out);

/* replaces: KVWorker<float>::Push + Wait (kv_app.h:110-118) served by the handles' push branch
 * (ftrl.h:54-74,112-141 ; sgd.h:46-52,90-96).  keys must be unique, in any order (KVWorker::Push takes them
 * sorted and unique; a repeated key is refused with XF_ERR_ARG before anything is applied).  gw: n floats or
 * NULL (app 0); gv: n*K floats or NULL (app 1). */
XF_DLL int xf_table_push(xf_table* t, const uint64_t* keys, uint64_t n, const float* gw, const float* gv);

/* same two operations on DEVICE pointers, asynchronous on the table's stream (no host sync).  The device push
 * does not look for repeated keys (that would take a sort): unique keys are the caller's contract there. */
XF_DLL int xf_table_pull_device(xf_table* t, const uint64_t* d_keys, uint64_t n, float* d_w_out, float* d_v_out);
XF_DLL int xf_table_push_device(xf_table* t, const uint64_t* d_keys, uint64_t n, const float* d_gw, const float* d_gv);

/* Overwrite / read full optimizer state of given keys (host arrays; any pointer but keys may be NULL).
 * The reference has no checkpoint; these exist for parity replay and save/restore.
 * export does NOT insert: present[i] = 0 and zeros for unknown keys. */
XF_DLL int xf_table_import(xf_table* t, const uint64_t* keys, uint64_t n, const float* w, const float* nw,
                           const float* zw, const float* v, const float* nv, const float* zv);
XF_DLL int xf_table_export(xf_table* t, const uint64_t* keys, uint64_t n, float* w, float* nw, float* zw,
                           float* v, float* nv, float* zv, uint8_t* present);

XF_DLL int xf_table_size(xf_table* t, uint64_t* n_keys);        /* = store.size() */
XF_DLL int xf_table_capacity(xf_table* t, uint64_t* n_slots);
XF_DLL int xf_table_row_bytes(xf_table* t, uint32_t* bytes);
XF_DLL int xf_table_latent_dim(xf_table* t, int* latent_dim);   /* K of the table (0 = LR) */
/* make room for at least n_keys keys at load factor <= 0.5 (rehashes on device if needed) */
XF_DLL int xf_table_reserve(xf_table* t, uint64_t n_keys);
/* Pre-populate: make the keys of the integer feature ids [first_id, first_id + count) exist with default
 * contents, as a Pull of them would (store[key], ftrl.h:56,114-120); ids are hashed on the device as their
 * decimal strings (load_data_from_disk.cc:151).  A sharded table keeps only the keys of its own range.
 * Asynchronous on the table's stream. */
XF_DLL int xf_table_touch_decimal_ids(xf_table* t, uint64_t first_id, uint64_t count);
/* copy up to max_keys live keys to host; *n_out = number of live keys */
XF_DLL int xf_table_list_keys(xf_table* t, uint64_t* keys_out, uint64_t max_keys, uint64_t* n_out);
/* binary checkpoint of the whole shard (keys + full optimizer state) */
XF_DLL int xf_table_save(xf_table* t, const char* path);
XF_DLL int xf_table_load(xf_table* t, const char* path);
/* text model dump, one "<key>\t<w>[\t<v_0> ... <v_K-1>]" line per key in key order (weights only, %.9g);
 * nonzero_only skips keys whose weights are all exactly 0 (FTRL's L1 zeros).  *written = lines (may be NULL) */
XF_DLL int xf_table_dump_text(xf_table* t, const char* path, int nonzero_only, uint64_t* written);
/* use an external CUDA stream (cudaStream_t passed as void*) for all table work; NULL = own stream */
XF_DLL int xf_table_set_stream(xf_table* t, void* cuda_stream);
XF_DLL int xf_table_sync(xf_table* t);

/* bucketing rule of ps::Postoffice::GetServerKeyRanges (postoffice.cc:134-143) + DefaultSlicer
 * (kv_app.h:405-460): shard = min(key / floor((2^64-1)/S), S-1).  Pure host function. */
XF_DLL int xf_shard_of(uint64_t key, int num_shards);

/* ------------------------------------------------------------------------------------------------
 * 3. Fused worker step
 * ---------------------------------------------------------------------------------------------- */
typedef struct xf_trainer xf_trainer;
typedef struct xf_comm xf_comm;

typedef struct xf_trainer_config {
  int model;             /* XF_MODEL_* ; FM requires table latent_dim > 0 */
  uint32_t max_rows;     /* largest batch (rows) the trainer will be given */
  uint32_t max_nnz;      /* largest batch (tokens) */
  int keep_loss;         /* 1: keep the per-row residual (pctr - label) of the last batch for xf_trainer_get_loss */
} xf_trainer_config;

/* replaces: new xflow::LRWorker / FMWorker (lr_worker.h:34-42, fm_worker.h:33-42) bound to the table.
 * comm may be NULL (single GPU).  With a comm of N ranks the table must be shard `rank` of N. */
XF_DLL int xf_trainer_create(xf_trainer** out, xf_table* table, xf_comm* comm, const xf_trainer_config* cfg);
XF_DLL int xf_trainer_destroy(xf_trainer* tr);

/* replaces: LRWorker::update / FMWorker::update on one slice (lr_worker.cc:145-177, fm_worker.cc:204-245)
 * INCLUDING the server-side optimizer step the Push triggers.  CSR batch in HOST memory:
 *   row_ptr[rows+1] (u32 offsets into keys), keys[nnz] (u64 feature hashes), labels[rows] (0/1).
 * Copies the batch to the device (pinned staging, async), runs the step, and returns
 * *mean_abs_loss = mean |pctr - label| of the batch (one float read back; pass NULL to skip the
 * read-back and stay asynchronous). */
XF_DLL int xf_trainer_step_host(xf_trainer* tr, const uint32_t* row_ptr, const uint64_t* keys,
                                const uint8_t* labels, uint32_t rows, uint32_t nnz, float* mean_abs_loss);
/* same step on a batch already resident in device memory; asynchronous on the table's stream */
XF_DLL int xf_trainer_step_device(xf_trainer* tr, const uint32_t* d_row_ptr, const uint64_t* d_keys,
                                  const uint8_t* d_labels, uint32_t rows, uint32_t nnz);
/* replaces: calculate_pctr (lr_worker.cc:25-71, fm_worker.cc:25-96): forward only, pctr_out[rows] host */
XF_DLL int xf_trainer_predict_host(xf_trainer* tr, const uint32_t* row_ptr, const uint64_t* keys, uint32_t rows,
                                   uint32_t nnz, float* pctr_out);
/* The textbook factorisation machine with feature VALUES (the reference ignores `val` and collapses
 * the interaction over k, fm_worker.cc:177-196 — parity mode = XF_MODEL_FM):
 *     y = sum_i w_i x_i + 1/2 sum_k [ (sum_i v_ik x_i)^2 - sum_i (v_ik x_i)^2 ],   p = sigmoid(y)
 *     dL/dw_i = (p - label) x_i ,  dL/dv_ik = (p - label) x_i (S_k - v_ik x_i) ,  gradients / rows, one FTRL / SGD step
 * vals[nnz] are the tokens' values (NULL: all 1).  Needs a table created with canonical_fm = 1; single GPU. */
XF_DLL int xf_trainer_step_host_values(xf_trainer* tr, const uint32_t* row_ptr, const uint64_t* keys, const float* vals,
                                       const uint8_t* labels, uint32_t rows, uint32_t nnz, float* mean_abs_loss);
XF_DLL int xf_trainer_step_device_values(xf_trainer* tr, const uint32_t* d_row_ptr, const uint64_t* d_keys,
                                         const float* d_vals, const uint8_t* d_labels, uint32_t rows, uint32_t nnz);
XF_DLL int xf_trainer_predict_host_values(xf_trainer* tr, const uint32_t* row_ptr, const uint64_t* keys, const float* vals,
                                          uint32_t rows, uint32_t nnz, float* pctr_out);
/* A DEFINED multi-view machine.  src/model/mvm/mvm_worker.cc indexes its per-row field sums one past
 * their end and multiplies in the sums of fields a row does not have (:43,57,75,86-92,262): its output is not a
 * function of its input.  This is the model that code is reaching for, over the same table:
 *     s[f][k] = sum over the row's tokens of field f of v_ik x_i ,   y = sum_k prod_{f present in the row} s[f][k]
 *     p = sigmoid(y) ,  dL/dv_ik = (p - label) x_i prod_{f' present, f' != field(i)} s[f'][k] ,  gradients / rows, one
 *     FTRL / SGD step per touched key on v only (no linear term: mvm_worker.cc pulls and pushes v alone).
 * fields[nnz]: the tokens' field ids (libffm's fgid), each < 32.  vals[nnz] or NULL (all 1).  A row without
 * tokens predicts sigmoid(0).  Needs XF_MODEL_MVM on a table with canonical_fm = 1 and latent_dim in {4,8,16,32};
 * single GPU. */
XF_DLL int xf_trainer_step_host_fields(xf_trainer* tr, const uint32_t* row_ptr, const uint64_t* keys,
                                       const uint8_t* fields, const float* vals, const uint8_t* labels, uint32_t rows,
                                       uint32_t nnz, float* mean_abs_loss);
XF_DLL int xf_trainer_predict_host_fields(xf_trainer* tr, const uint32_t* row_ptr, const uint64_t* keys,
                                          const uint8_t* fields, const float* vals, uint32_t rows, uint32_t nnz,
                                          float* pctr_out);
/* the one-off "init push" of key 0 with zero gradient (lr_worker.cc:180-182, fm_worker.cc:248-252) */
XF_DLL int xf_trainer_init_push(xf_trainer* tr);
/* residuals (pctr - label) of the last step; needs keep_loss = 1 */
XF_DLL int xf_trainer_get_loss(xf_trainer* tr, float* loss_out, uint32_t rows);
/* counters since creation: steps, rows, tokens and unique keys summed over steps (device counter) */
XF_DLL int xf_trainer_stats(xf_trainer* tr, uint64_t* steps, uint64_t* rows, uint64_t* nnz, uint64_t* unique_keys);
/* number of kernels this library launched on behalf of the trainer/table since creation */
XF_DLL int xf_trainer_launches(xf_trainer* tr, uint64_t* launches);
XF_DLL int xf_trainer_sync(xf_trainer* tr);
/* block until every host->device batch copy issued so far has finished (the caller may then
 * overwrite host buffers it passed to xf_trainer_step_host with mean_abs_loss == NULL) */
XF_DLL int xf_trainer_wait_uploads(xf_trainer* tr);
/* Asynchronous variant of xf_trainer_step_host for pipelined callers: the batch arrays and
 * `pinned_abs_loss_sum` (one float, receives sum |pctr - label| of this batch by an asynchronous
 * device->host copy) must be page-locked host memory and stay untouched until xf_trainer_sync /
 * xf_trainer_wait_uploads.  Never blocks on the device. */
XF_DLL int xf_trainer_step_host_async(xf_trainer* tr, const uint32_t* row_ptr, const uint64_t* keys,
                                      const uint8_t* labels, uint32_t rows, uint32_t nnz,
                                      float* pinned_abs_loss_sum);
/* Same as xf_trainer_step_host_async for callers that hold integer feature ids instead of hashed keys:
 * ids[nnz] are u32 ids whose DECIMAL STRING is what the reference's loader would hash
 * (load_data_from_disk.cc:151); the device computes keys = std::hash(decimal string) itself (ingest.cu),
 * so only 4 bytes per token cross PCIe.  Buffers must be page-locked. */
XF_DLL int xf_trainer_step_host_ids_async(xf_trainer* tr, const uint32_t* row_ptr, const uint32_t* ids,
                                          const uint8_t* labels, uint32_t rows, uint32_t nnz,
                                          float* pinned_abs_loss_sum);
/* keys[i] = std::hash<std::string>(decimal string of ids[i]) on DEVICE arrays (stream: cudaStream_t or NULL) */
XF_DLL int xf_hash_decimal_ids_device(const uint32_t* d_ids, uint64_t n, uint64_t* d_keys, void* cuda_stream);
/* Device-side ingest of one text block (load_data_from_disk.cc:126-209 on the GPU, ingest.cu): uploads
 * `len` bytes of "<label>\t<fgid>:<fid>:<val> ...\n" rows, parses them into a CSR batch that stays on
 * the device, and reports its size.  xf_trainer_step_ingested / _predict_ingested then run the step on a
 * row range [row_start, row_end) of that block (the reference's per-thread slices, lr_worker.cc:190-196). */
XF_DLL int xf_trainer_ingest_text(xf_trainer* tr, const char* text, uint64_t len, uint32_t* rows, uint32_t* nnz);
/* The same in two phases, for pipelined callers: _begin enqueues the copy and the parse of the NEXT block on
 * the trainer's ingest stream (two device-side block buffers) and returns at once; `text` must stay untouched
 * until _end returns (page-locked memory is copied by DMA, pageable memory is staged first).  _end waits for
 * that parse only — not for training steps still running on the previous block — and makes the block
 * current.  Up to TWO blocks may be outstanding: the second _begin copies its text at once (into the buffer of
 * the block being trained on, whose text is no longer needed) and its parse is launched by the _end that
 * retires that block, so that with  begin(i+2); step(i); end(i+1)  the H2D of one block runs beside the parse
 * of the previous one and the training step of the one before.  _end always completes the OLDEST outstanding
 * block.  If _end fails (malformed or oversized block) every outstanding block is dropped. */
XF_DLL int xf_trainer_ingest_begin(xf_trainer* tr, const char* text, uint64_t len);
XF_DLL int xf_trainer_ingest_end(xf_trainer* tr, uint32_t* rows, uint32_t* nnz);
XF_DLL int xf_trainer_step_ingested(xf_trainer* tr, uint32_t row_start, uint32_t row_end);
/* copy the ingested block's CSR back to host arrays of rows+1 / nnz / rows elements (any may be NULL) */
XF_DLL int xf_trainer_ingested_export(xf_trainer* tr, uint32_t* row_ptr_out, uint64_t* keys_out, uint8_t* labels_out);
XF_DLL int xf_trainer_predict_ingested(xf_trainer* tr, uint32_t row_start, uint32_t row_end, float* pctr_out,
                                       uint8_t* labels_out);
/* Per-kernel device timing for roofline reporting.  on != 0: record CUDA events around the kernels
 * of every following step (on the table's stream).  xf_trainer_profile syncs and returns, summed
 * over the profiled steps since the last call: ms[0] = fused step kernel, ms[1] = optimizer kernel
 * (+ batch bookkeeping), and the number of steps. */
XF_DLL int xf_trainer_set_profile(xf_trainer* tr, int on);
XF_DLL int xf_trainer_profile(xf_trainer* tr, double ms[2], uint64_t* steps);
/* page-locked host memory for callers without a CUDA runtime of their own */
XF_DLL int xf_host_alloc(void** out, uint64_t bytes);
XF_DLL int xf_host_free(void* p);

/* Base::calculate_auc (base.h:84-110), host: out[0]=logloss (base-2, not negated, float accumulator)
 * out[1]=auc (float `area`; NaN when single-class) out[2]=tp out[3]=fp */
XF_DLL int xf_auc_logloss(const int32_t* labels, const float* pctr, uint64_t n, double out[4]);
/* the same metric in exact arithmetic (not the reference's numbers): out[0] = mean negative natural-log
 * likelihood, out[1] = AUC with ties counted 1/2, out[2]=positives out[3]=negatives */
XF_DLL int xf_auc_logloss_exact(const int32_t* labels, const float* pctr, uint64_t n, double out[4]);

/* The metric on the DEVICE (csrc/metric.cu): predictions and labels of every forward block are appended to a
 * device buffer, xf_metric_finish sorts them there (radix sort, stable) and reduces:
 *   out[0] the reference's logloss (base 2, not negated; double accumulator)   out[1] the reference's AUC (64-bit
 *   integer rank sum instead of the float `area` that stops counting at 2^24)  out[2] positives  out[3] negatives
 *   out[4] mean negative natural-log likelihood   out[5] AUC with ties counted 1/2          (base.h:84-110) */
typedef struct xf_metric xf_metric;
XF_DLL int xf_metric_create(xf_metric** out, int device);
XF_DLL int xf_metric_destroy(xf_metric* m);
XF_DLL int xf_metric_reset(xf_metric* m);
/* append n predictions / 0-1 labels living in device memory (copies run on `cuda_stream`, no sync) */
XF_DLL int xf_metric_add_device(xf_metric* m, const float* d_pctr, const uint8_t* d_labels, uint64_t n, void* cuda_stream);
XF_DLL int xf_metric_finish(xf_metric* m, void* cuda_stream, double out[6]);
XF_DLL int xf_auc_logloss_device(const float* d_pctr, const uint8_t* d_labels, uint64_t n, int device, void* cuda_stream,
                                 double out[6]);
/* forward pass over rows [row_start, row_end) of the current ingested block, appended to `m` without leaving the
 * device (asynchronous).  pctr_out / labels_out (optional host arrays of rows elements): the same values for a
 * caller that also writes them out; the call then waits for them. */
XF_DLL int xf_trainer_predict_ingested_metric(xf_trainer* tr, uint32_t row_start, uint32_t row_end, xf_metric* m,
                                              float* pctr_out, uint8_t* labels_out);

/* ------------------------------------------------------------------------------------------------
 * 4. Host ingest
 * ---------------------------------------------------------------------------------------------- */
/* std::hash<std::string> of libstdc++ (MurmurHash64A, seed 0xc70f6907) as used at
 * load_data_from_disk.cc:151,173,194 */
XF_DLL uint64_t xf_hash_bytes(const char* s, uint64_t len);
/* hashes of the decimal strings of ids[] ("%llu"): what the loader produces for numeric feature ids */
XF_DLL int xf_hash_decimal_ids(const uint64_t* ids, uint64_t n, uint64_t* out);

typedef struct xf_loader xf_loader;
/* replaces: xflow::LoadData(path, block_bytes) (load_data_from_disk.h:19-21) */
XF_DLL int xf_loader_open(xf_loader** out, const char* path, uint64_t block_bytes);
XF_DLL int xf_loader_close(xf_loader* l);
/* restart at the first byte (what re-opening the file at the top of every epoch does, lr_worker.cc:184) */
XF_DLL int xf_loader_rewind(xf_loader* l);
/* replaces: load_minibatch_hash_data_fread (load_data_from_disk.cc:103-210): parse the next block.
 * *rows = 0 at end of file.  The CSR arrays stay valid until the next call. */
XF_DLL int xf_loader_next(xf_loader* l, uint32_t* rows, uint32_t* nnz);
XF_DLL int xf_loader_batch(xf_loader* l, const uint32_t** row_ptr, const uint64_t** keys, const uint8_t** labels);
/* block formation only (load_data_from_disk.cc:108-124): the next block's raw text, for the device parser
 * (xf_trainer_ingest_text / _begin).  *len = 0 at end of file.  The loader alternates two (page-locked) text
 * buffers: a block's text stays valid until the call AFTER the next one.  Tab-less rows ("0\n") count as rows
 * without features on the device parser; the host parser (xf_loader_next), like the reference, scans on to the
 * next tab and merges them into the following row. */
XF_DLL int xf_loader_next_raw(xf_loader* l, const char** text, uint64_t* len);

/* ------------------------------------------------------------------------------------------------
 * 5. Multi-GPU exchange (one process per GPU over NVLink / NVSwitch).  A trainer created with a comm of
 *    N ranks runs the sharded step (csrc/comm.cu): every rank must create its trainer with the same
 *    max_rows / max_nnz / model and call the step / predict entry points the same number of times, in the
 *    same order (an empty batch is a valid step).  NCCL serves bootstrap only (exchange of cudaIpc handles);
 *    inside a step kernels store into the peers' memory directly.
 * ---------------------------------------------------------------------------------------------- */
#define XF_COMM_ID_BYTES 128
/* rank 0 creates the id and distributes it out of band (file, MPI, torch.distributed ...) */
XF_DLL int xf_comm_get_id(uint8_t id[XF_COMM_ID_BYTES]);
XF_DLL int xf_comm_create(xf_comm** out, const uint8_t id[XF_COMM_ID_BYTES], int rank, int nranks, int device);
/* the same with the id passed through a file: rank 0 writes it, the others wait for it (launchers that have no
 * other channel, e.g. the reference's CLI started once per GPU with XFLOW_RANK / XFLOW_WORLD) */
XF_DLL int xf_comm_create_from_file(xf_comm** out, const char* path, int rank, int nranks, int device);
/* max over the ranks of one host value (blocking; to agree on the number of collective steps) */
XF_DLL int xf_comm_allreduce_max(xf_comm* c, uint64_t* inout);
XF_DLL int xf_comm_destroy(xf_comm* c);
XF_DLL int xf_comm_barrier(xf_comm* c);
/* ------------------------------------------------------------------------------------------------
 * 1. Reference C API (src/c_api/c_api.h:26-29), unchanged signatures.
 *    XFCreate builds an LR worker on <train_path>-%05d / <test_path>-%05d (rank from XFLOW_RANK,
 *    default 0); paths are copied.  XFStartTrain trains `epochs` (default 60, lr_worker.h:63; env
 *    XFLOW_EPOCHS) and, on rank 0, predicts and prints logloss/auc like lr_worker.cc:207-217.
 *    Extensions: XFCreateEx picks model/optimizer/K; XFDestroy frees the handle.
 * ---------------------------------------------------------------------------------------------- */
XF_DLL int XFCreate(void** h, const char* train_path, const char* test_path);
XF_DLL int XFStartTrain(void** h);
XF_DLL int XFCreateEx(void** h, const char* train_path, const char* test_path, int model, int optimizer,
                      int latent_dim, int epochs);
XF_DLL int XFDestroy(void** h);

#endif /* XFLOW_B200_H_ */
