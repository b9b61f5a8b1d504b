"""Thin ctypes binding of libxflow_b200.so (include/xflow_b200.h) for tests and bench.py.

The product is the C ABI; this module only marshals numpy arrays / raw pointers into it.  It never
computes anything itself and has no fallback: if the shared library (or a CUDA device, for compute
calls) is missing, calls fail loudly.
"""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libxflow_b200.so")

MODEL_LR, MODEL_FM, MODEL_FM_CANONICAL, MODEL_MVM = 0, 1, 2, 3
OPT_FTRL, OPT_SGD = 0, 1
VINIT_DEFAULT, VINIT_COUNTER, VINIT_ZERO = 0, 1, 3
COMM_ID_BYTES = 128

_lib = None


class XflowError(RuntimeError):
    pass


class TableConfig(C.Structure):
    _fields_ = [("device", C.c_int), ("latent_dim", C.c_int), ("optimizer", C.c_int), ("alpha", C.c_float),
                ("beta", C.c_float), ("lambda1", C.c_float), ("lambda2", C.c_float),
                ("learning_rate", C.c_float), ("v_init", C.c_int), ("seed", C.c_uint64),
                ("capacity", C.c_uint64), ("shard_index", C.c_int), ("num_shards", C.c_int), ("canonical_fm", C.c_int)]


class TrainerConfig(C.Structure):
    _fields_ = [("model", C.c_int), ("max_rows", C.c_uint32), ("max_nnz", C.c_uint32), ("keep_loss", C.c_int)]


# name -> (restype, argtypes); also the list the symbol-export test checks against the header
_vp, _u64, _u32, _i, _f = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_float
SIGNATURES = {
    "xf_last_error": (C.c_char_p, []),
    "xf_version": (_i, []),
    "xf_device_count": (_i, []),
    "xf_table_config_default": (_i, [_vp]),
    "xf_table_create": (_i, [_vp, _vp]),
    "xf_table_destroy": (_i, [_vp]),
    "xf_table_pull": (_i, [_vp, _vp, _u64, _vp, _vp]),
    "xf_table_push": (_i, [_vp, _vp, _u64, _vp, _vp]),
    "xf_table_pull_device": (_i, [_vp, _vp, _u64, _vp, _vp]),
    "xf_table_push_device": (_i, [_vp, _vp, _u64, _vp, _vp]),
    "xf_table_import": (_i, [_vp, _vp, _u64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "xf_table_export": (_i, [_vp, _vp, _u64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "xf_table_size": (_i, [_vp, _vp]),
    "xf_table_capacity": (_i, [_vp, _vp]),
    "xf_table_row_bytes": (_i, [_vp, _vp]),
    "xf_table_latent_dim": (_i, [_vp, _vp]),
    "xf_table_reserve": (_i, [_vp, _u64]),
    "xf_table_list_keys": (_i, [_vp, _vp, _u64, _vp]),
    "xf_table_touch_decimal_ids": (_i, [_vp, _u64, _u64]),
    "xf_table_save": (_i, [_vp, C.c_char_p]),
    "xf_table_load": (_i, [_vp, C.c_char_p]),
    "xf_table_set_stream": (_i, [_vp, _vp]),
    "xf_table_sync": (_i, [_vp]),
    "xf_shard_of": (_i, [_u64, _i]),
    "xf_trainer_create": (_i, [_vp, _vp, _vp, _vp]),
    "xf_trainer_destroy": (_i, [_vp]),
    "xf_trainer_step_host": (_i, [_vp, _vp, _vp, _vp, _u32, _u32, _vp]),
    "xf_trainer_step_device": (_i, [_vp, _vp, _vp, _vp, _u32, _u32]),
    "xf_trainer_predict_host": (_i, [_vp, _vp, _vp, _u32, _u32, _vp]),
    "xf_trainer_step_host_values": (_i, [_vp, _vp, _vp, _vp, _vp, _u32, _u32, _vp]),
    "xf_trainer_step_device_values": (_i, [_vp, _vp, _vp, _vp, _vp, _u32, _u32]),
    "xf_trainer_predict_host_values": (_i, [_vp, _vp, _vp, _vp, _u32, _u32, _vp]),
    "xf_trainer_step_host_fields": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _u32, _u32, _vp]),
    "xf_trainer_predict_host_fields": (_i, [_vp, _vp, _vp, _vp, _vp, _u32, _u32, _vp]),
    "xf_trainer_init_push": (_i, [_vp]),
    "xf_trainer_get_loss": (_i, [_vp, _vp, _u32]),
    "xf_trainer_stats": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "xf_trainer_launches": (_i, [_vp, _vp]),
    "xf_trainer_sync": (_i, [_vp]),
    "xf_trainer_wait_uploads": (_i, [_vp]),
    "xf_trainer_step_host_async": (_i, [_vp, _vp, _vp, _vp, _u32, _u32, _vp]),
    "xf_trainer_step_host_ids_async": (_i, [_vp, _vp, _vp, _vp, _u32, _u32, _vp]),
    "xf_hash_decimal_ids_device": (_i, [_vp, _u64, _vp, _vp]),
    "xf_auc_logloss_exact": (_i, [_vp, _vp, _u64, _vp]),
    "xf_table_dump_text": (_i, [_vp, C.c_char_p, _i, _vp]),
    "xf_trainer_ingest_text": (_i, [_vp, _vp, _u64, _vp, _vp]),
    "xf_trainer_ingest_begin": (_i, [_vp, _vp, _u64]),
    "xf_trainer_ingest_end": (_i, [_vp, _vp, _vp]),
    "xf_trainer_step_ingested": (_i, [_vp, _u32, _u32]),
    "xf_trainer_ingested_export": (_i, [_vp, _vp, _vp, _vp]),
    "xf_trainer_predict_ingested": (_i, [_vp, _u32, _u32, _vp, _vp]),
    "xf_loader_next_raw": (_i, [_vp, _vp, _vp]),
    "xf_trainer_set_profile": (_i, [_vp, _i]),
    "xf_trainer_profile": (_i, [_vp, _vp, _vp]),
    "xf_host_alloc": (_i, [_vp, _u64]),
    "xf_host_free": (_i, [_vp]),
    "xf_auc_logloss": (_i, [_vp, _vp, _u64, _vp]),
    "xf_metric_create": (_i, [_vp, _i]),
    "xf_metric_destroy": (_i, [_vp]),
    "xf_metric_reset": (_i, [_vp]),
    "xf_metric_add_device": (_i, [_vp, _vp, _vp, _u64, _vp]),
    "xf_metric_finish": (_i, [_vp, _vp, _vp]),
    "xf_auc_logloss_device": (_i, [_vp, _vp, _u64, _i, _vp, _vp]),
    "xf_trainer_predict_ingested_metric": (_i, [_vp, _u32, _u32, _vp, _vp, _vp]),
    "xf_hash_bytes": (_u64, [C.c_char_p, _u64]),
    "xf_hash_decimal_ids": (_i, [_vp, _u64, _vp]),
    "xf_loader_open": (_i, [_vp, C.c_char_p, _u64]),
    "xf_loader_close": (_i, [_vp]),
    "xf_loader_rewind": (_i, [_vp]),
    "xf_loader_next": (_i, [_vp, _vp, _vp]),
    "xf_loader_batch": (_i, [_vp, _vp, _vp, _vp]),
    "xf_comm_get_id": (_i, [_vp]),
    "xf_comm_create": (_i, [_vp, _vp, _i, _i, _i]),
    "xf_comm_create_from_file": (_i, [_vp, C.c_char_p, _i, _i, _i]),
    "xf_comm_allreduce_max": (_i, [_vp, _vp]),
    "xf_comm_destroy": (_i, [_vp]),
    "xf_comm_barrier": (_i, [_vp]),
    "XFCreate": (_i, [_vp, C.c_char_p, C.c_char_p]),
    "XFStartTrain": (_i, [_vp]),
    "XFCreateEx": (_i, [_vp, C.c_char_p, C.c_char_p, _i, _i, _i, _i]),
    "XFDestroy": (_i, [_vp]),
}


def lib():
    """Load the shared library (building is __graft_entry__.build()'s job; missing .so is an error)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise XflowError("%s not built: run `python -m xflow_b200.build`" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise XflowError("xflow_b200 error %d: %s" % (rc, lib().xf_last_error().decode(errors="replace")))


def _p(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(int(a))  # raw address (e.g. torch tensor.data_ptr())


def device_count():
    return lib().xf_device_count()


def hash_bytes(s: bytes) -> int:
    return lib().xf_hash_bytes(s, len(s))


def hash_decimal_ids(ids):
    ids = np.ascontiguousarray(ids, np.uint64)
    out = np.empty_like(ids)
    _check(lib().xf_hash_decimal_ids(_p(ids), ids.size, _p(out)))
    return out


def shard_of(key, num_shards):
    return lib().xf_shard_of(int(key), int(num_shards))


def auc_logloss(labels, pctr):
    labels = np.ascontiguousarray(labels, np.int32)
    pctr = np.ascontiguousarray(pctr, np.float32)
    out = np.zeros(4, np.float64)
    _check(lib().xf_auc_logloss(_p(labels), _p(pctr), labels.size, _p(out)))
    return dict(logloss=float(out[0]), auc=float(out[1]), tp=int(out[2]), fp=int(out[3]))


def auc_logloss_exact(labels, pctr):
    """Exact-arithmetic test metric: natural-log logloss (negated) and tie-aware AUC."""
    labels = np.ascontiguousarray(labels, np.int32)
    pctr = np.ascontiguousarray(pctr, np.float32)
    out = np.zeros(4, np.float64)
    _check(lib().xf_auc_logloss_exact(_p(labels), _p(pctr), labels.size, _p(out)))
    return dict(logloss=float(out[0]), auc=float(out[1]), positives=int(out[2]), negatives=int(out[3]))


class Loader:
    """xflow::LoadData replacement: iterate CSR blocks of a text shard."""

    def __init__(self, path, block_bytes):
        self.h = C.c_void_p()
        _check(lib().xf_loader_open(C.byref(self.h), path.encode(), block_bytes))

    def close(self):
        if self.h:
            lib().xf_loader_close(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        self.close()

    def __iter__(self):
        return self

    def __next__(self):
        rows, nnz = C.c_uint32(), C.c_uint32()
        _check(lib().xf_loader_next(self.h, C.byref(rows), C.byref(nnz)))
        if rows.value == 0:
            raise StopIteration
        rp, kp, lp = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _check(lib().xf_loader_batch(self.h, C.byref(rp), C.byref(kp), C.byref(lp)))
        B, n = rows.value, nnz.value
        row_ptr = np.ctypeslib.as_array(C.cast(rp, C.POINTER(C.c_uint32)), (B + 1,)).copy()
        keys = np.ctypeslib.as_array(C.cast(kp, C.POINTER(C.c_uint64)), (max(n, 1),))[:n].copy()
        labels = np.ctypeslib.as_array(C.cast(lp, C.POINTER(C.c_uint8)), (B,)).copy()
        return row_ptr, keys, labels


    def next_raw(self):
        """Block formation only: the next block's raw text (b"" at end of file)."""
        text, n = C.c_void_p(), C.c_uint64()
        _check(lib().xf_loader_next_raw(self.h, C.byref(text), C.byref(n)))
        return C.string_at(text, n.value) if n.value else b""


class Table:
    def __init__(self, latent_dim=0, optimizer=OPT_FTRL, device=0, capacity=0, v_init=VINIT_DEFAULT, seed=0,
                 shard_index=0, num_shards=1, **hparams):
        L = lib()
        cfg = TableConfig()
        _check(L.xf_table_config_default(C.byref(cfg)))
        cfg.device, cfg.latent_dim, cfg.optimizer = device, latent_dim, optimizer
        cfg.capacity, cfg.v_init, cfg.seed = capacity, v_init, seed
        cfg.shard_index, cfg.num_shards = shard_index, num_shards
        for k, v in hparams.items():
            if not hasattr(cfg, k):
                raise TypeError("unknown hyper-parameter %s" % k)
            setattr(cfg, k, v)
        self.K = latent_dim
        self.cfg = cfg
        self.h = C.c_void_p()
        _check(L.xf_table_create(C.byref(self.h), C.byref(cfg)))

    def close(self):
        if getattr(self, "h", None):
            lib().xf_table_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def pull(self, keys, want_w=True, want_v=True):
        keys = np.ascontiguousarray(keys, np.uint64)
        w = np.empty(keys.size, np.float32) if want_w else None
        v = np.empty((keys.size, self.K), np.float32) if (want_v and self.K) else None
        _check(lib().xf_table_pull(self.h, _p(keys), keys.size, _p(w), _p(v)))
        return w, v

    def push(self, keys, gw=None, gv=None):
        keys = np.ascontiguousarray(keys, np.uint64)
        gw = None if gw is None else np.ascontiguousarray(gw, np.float32)
        gv = None if gv is None else np.ascontiguousarray(gv, np.float32)
        _check(lib().xf_table_push(self.h, _p(keys), keys.size, _p(gw), _p(gv)))

    def import_(self, keys, w=None, nw=None, zw=None, v=None, nv=None, zv=None):
        keys = np.ascontiguousarray(keys, np.uint64)
        arrs = [None if a is None else np.ascontiguousarray(a, np.float32) for a in (w, nw, zw, v, nv, zv)]
        _check(lib().xf_table_import(self.h, _p(keys), keys.size, *[_p(a) for a in arrs]))

    def export(self, keys):
        keys = np.ascontiguousarray(keys, np.uint64)
        n, K = keys.size, self.K
        out = dict(keys=keys, w=np.zeros(n, np.float32), nw=np.zeros(n, np.float32), zw=np.zeros(n, np.float32),
                   v=np.zeros((n, K), np.float32), nv=np.zeros((n, K), np.float32),
                   zv=np.zeros((n, K), np.float32), present=np.zeros(n, np.uint8))
        _check(lib().xf_table_export(self.h, _p(keys), n, _p(out["w"]), _p(out["nw"]), _p(out["zw"]),
                                     _p(out["v"]) if K else None, _p(out["nv"]) if K else None,
                                     _p(out["zv"]) if K else None, _p(out["present"])))
        return out

    def size(self):
        n = C.c_uint64()
        _check(lib().xf_table_size(self.h, C.byref(n)))
        return n.value

    def capacity(self):
        n = C.c_uint64()
        _check(lib().xf_table_capacity(self.h, C.byref(n)))
        return n.value

    def row_bytes(self):
        n = C.c_uint32()
        _check(lib().xf_table_row_bytes(self.h, C.byref(n)))
        return n.value

    def reserve(self, n_keys):
        _check(lib().xf_table_reserve(self.h, int(n_keys)))

    def touch_decimal_ids(self, first_id, count):
        _check(lib().xf_table_touch_decimal_ids(self.h, int(first_id), int(count)))

    def list_keys(self):
        n = self.size()
        keys = np.empty(max(n, 1), np.uint64)
        got = C.c_uint64()
        _check(lib().xf_table_list_keys(self.h, _p(keys), n, C.byref(got)))
        return keys[:min(n, got.value)]

    def save(self, path):
        _check(lib().xf_table_save(self.h, path.encode()))

    def dump_text(self, path, nonzero_only=False):
        n = C.c_uint64()
        _check(lib().xf_table_dump_text(self.h, path.encode(), int(nonzero_only), C.byref(n)))
        return n.value

    def load(self, path):
        _check(lib().xf_table_load(self.h, path.encode()))

    def set_stream(self, cuda_stream):
        _check(lib().xf_table_set_stream(self.h, C.c_void_p(int(cuda_stream)) if cuda_stream else None))

    def sync(self):
        _check(lib().xf_table_sync(self.h))


class Comm:
    @staticmethod
    def new_id():
        buf = np.zeros(COMM_ID_BYTES, np.uint8)
        _check(lib().xf_comm_get_id(_p(buf)))
        return buf

    def __init__(self, comm_id, rank, nranks, device):
        comm_id = np.ascontiguousarray(comm_id, np.uint8)
        self.rank, self.nranks = rank, nranks
        self.h = C.c_void_p()
        _check(lib().xf_comm_create(C.byref(self.h), _p(comm_id), rank, nranks, device))

    def barrier(self):
        _check(lib().xf_comm_barrier(self.h))

    def close(self):
        if getattr(self, "h", None):
            lib().xf_comm_destroy(self.h)
            self.h = None


class Trainer:
    def __init__(self, table, model=MODEL_LR, max_rows=65536, max_nnz=65536 * 64, keep_loss=False, comm=None):
        cfg = TrainerConfig(model, max_rows, max_nnz, 1 if keep_loss else 0)
        self.table = table
        self.comm = comm
        self.h = C.c_void_p()
        _check(lib().xf_trainer_create(C.byref(self.h), table.h, comm.h if comm else None, C.byref(cfg)))

    def close(self):
        if getattr(self, "h", None):
            lib().xf_trainer_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def init_push(self):
        _check(lib().xf_trainer_init_push(self.h))

    def step_host(self, row_ptr, keys, labels, want_loss=True):
        """One update() on host CSR arrays (numpy or raw pinned addresses with explicit sizes)."""
        row_ptr = np.ascontiguousarray(row_ptr, np.uint32)
        keys = np.ascontiguousarray(keys, np.uint64)
        labels = np.ascontiguousarray(labels, np.uint8)
        loss = C.c_float()
        _check(lib().xf_trainer_step_host(self.h, _p(row_ptr), _p(keys), _p(labels), labels.size, keys.size,
                                          C.byref(loss) if want_loss else None))
        return loss.value if want_loss else None

    def step_host_values(self, row_ptr, keys, vals, labels):
        """One step of the canonical FM (XF_MODEL_FM_CANONICAL) on host CSR arrays with feature values."""
        row_ptr = np.ascontiguousarray(row_ptr, np.uint32)
        keys = np.ascontiguousarray(keys, np.uint64)
        vals = None if vals is None else np.ascontiguousarray(vals, np.float32)
        labels = np.ascontiguousarray(labels, np.uint8)
        loss = C.c_float()
        _check(lib().xf_trainer_step_host_values(self.h, _p(row_ptr), _p(keys), _p(vals), _p(labels), labels.size, keys.size,
                                                 C.byref(loss)))
        return loss.value

    def predict_host_values(self, row_ptr, keys, vals):
        row_ptr = np.ascontiguousarray(row_ptr, np.uint32)
        keys = np.ascontiguousarray(keys, np.uint64)
        vals = None if vals is None else np.ascontiguousarray(vals, np.float32)
        rows = row_ptr.size - 1
        out = np.empty(rows, np.float32)
        _check(lib().xf_trainer_predict_host_values(self.h, _p(row_ptr), _p(keys), _p(vals), rows, keys.size, _p(out)))
        return out

    def step_host_fields(self, row_ptr, keys, fields, vals, labels):
        """One step of the defined multi-view machine (XF_MODEL_MVM): host CSR arrays + the tokens' field ids."""
        row_ptr = np.ascontiguousarray(row_ptr, np.uint32)
        keys = np.ascontiguousarray(keys, np.uint64)
        fields = np.ascontiguousarray(fields, np.uint8)
        vals = None if vals is None else np.ascontiguousarray(vals, np.float32)
        labels = np.ascontiguousarray(labels, np.uint8)
        loss = C.c_float()
        _check(lib().xf_trainer_step_host_fields(self.h, _p(row_ptr), _p(keys), _p(fields), _p(vals), _p(labels), labels.size,
                                                 keys.size, C.byref(loss)))
        return loss.value

    def predict_host_fields(self, row_ptr, keys, fields, vals):
        row_ptr = np.ascontiguousarray(row_ptr, np.uint32)
        keys = np.ascontiguousarray(keys, np.uint64)
        fields = np.ascontiguousarray(fields, np.uint8)
        vals = None if vals is None else np.ascontiguousarray(vals, np.float32)
        rows = row_ptr.size - 1
        out = np.empty(rows, np.float32)
        _check(lib().xf_trainer_predict_host_fields(self.h, _p(row_ptr), _p(keys), _p(fields), _p(vals), rows, keys.size, _p(out)))
        return out

    def step_host_raw(self, row_ptr_addr, keys_addr, labels_addr, rows, nnz, want_loss=True):
        loss = C.c_float()
        _check(lib().xf_trainer_step_host(self.h, _p(row_ptr_addr), _p(keys_addr), _p(labels_addr), rows, nnz,
                                          C.byref(loss) if want_loss else None))
        return loss.value if want_loss else None

    def step_host_async(self, row_ptr_addr, keys_addr, labels_addr, rows, nnz, out_addr=None):
        """Pipelined step on page-locked buffers given by address; never blocks on the device."""
        _check(lib().xf_trainer_step_host_async(self.h, _p(row_ptr_addr), _p(keys_addr), _p(labels_addr), rows, nnz,
                                                _p(out_addr) if out_addr else None))

    def step_host_ids_async(self, row_ptr_addr, ids_addr, labels_addr, rows, nnz, out_addr=None):
        """Like step_host_async but with u32 feature ids (hashed to keys on the device)."""
        _check(lib().xf_trainer_step_host_ids_async(self.h, _p(row_ptr_addr), _p(ids_addr), _p(labels_addr), rows,
                                                    nnz, _p(out_addr) if out_addr else None))

    def ingest_text(self, text: bytes):
        """Parse one text block on the device; returns (rows, nnz) of the CSR now resident there."""
        rows, nnz = C.c_uint32(), C.c_uint32()
        buf = C.create_string_buffer(text, len(text))
        _check(lib().xf_trainer_ingest_text(self.h, C.cast(buf, C.c_void_p), len(text), C.byref(rows), C.byref(nnz)))
        return rows.value, nnz.value

    def ingested_export(self, rows, nnz):
        rp = np.empty(rows + 1, np.uint32)
        keys = np.empty(nnz, np.uint64)
        lab = np.empty(rows, np.uint8)
        _check(lib().xf_trainer_ingested_export(self.h, _p(rp), _p(keys), _p(lab)))
        return rp, keys, lab

    def step_ingested(self, row_start, row_end):
        _check(lib().xf_trainer_step_ingested(self.h, row_start, row_end))

    def predict_ingested(self, row_start, row_end):
        n = row_end - row_start
        p = np.empty(n, np.float32)
        lab = np.empty(n, np.uint8)
        _check(lib().xf_trainer_predict_ingested(self.h, row_start, row_end, _p(p), _p(lab)))
        return p, lab

    def wait_uploads(self):
        _check(lib().xf_trainer_wait_uploads(self.h))

    def set_profile(self, on):
        _check(lib().xf_trainer_set_profile(self.h, 1 if on else 0))

    def profile(self):
        ms = (C.c_double * 2)()
        steps = C.c_uint64()
        _check(lib().xf_trainer_profile(self.h, ms, C.byref(steps)))
        return dict(step_ms=ms[0], update_ms=ms[1], steps=steps.value)

    def step_device(self, d_row_ptr, d_keys, d_labels, rows, nnz):
        _check(lib().xf_trainer_step_device(self.h, _p(d_row_ptr), _p(d_keys), _p(d_labels), rows, nnz))

    def predict_host(self, row_ptr, keys):
        row_ptr = np.ascontiguousarray(row_ptr, np.uint32)
        keys = np.ascontiguousarray(keys, np.uint64)
        rows = row_ptr.size - 1
        out = np.empty(rows, np.float32)
        _check(lib().xf_trainer_predict_host(self.h, _p(row_ptr), _p(keys), rows, keys.size, _p(out)))
        return out

    def get_loss(self, rows):
        out = np.empty(rows, np.float32)
        _check(lib().xf_trainer_get_loss(self.h, _p(out), rows))
        return out

    def stats(self):
        a, b, c, d = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(lib().xf_trainer_stats(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return dict(steps=a.value, rows=b.value, nnz=c.value, unique_keys=d.value)

    def launches(self):
        n = C.c_uint64()
        _check(lib().xf_trainer_launches(self.h, C.byref(n)))
        return n.value

    def sync(self):
        _check(lib().xf_trainer_sync(self.h))
