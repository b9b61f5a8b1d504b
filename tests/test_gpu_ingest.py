"""Device-side text ingest (pytest -m gpu): xf_trainer_ingest_text (ingest.cu) against the host parser
xf_loader_next (loader.cc), which the CPU suite pins to the reference's LoadData
(load_data_from_disk.cc:108-209) through the oracle.  Integer / byte work: bit-exact."""
import os
import subprocess
import sys

import numpy as np
import pytest

from common import GOLDEN, assert_close
from xflow_b200 import api, datagen

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TRAIN = os.path.join(GOLDEN, "data", "small_train-00000")
TEST = os.path.join(GOLDEN, "data", "small_test-00000")


def _blocks_host(path, block_bytes):
    return list(api.Loader(path, block_bytes))


def _blocks_device(path, block_bytes, trainer):
    out = []
    ld = api.Loader(path, block_bytes)
    while True:
        text = ld.next_raw()
        if not text:
            break
        rows, nnz = trainer.ingest_text(text)
        if rows == 0:
            break
        out.append(trainer.ingested_export(rows, nnz))
    return out


def _same_blocks(host, dev):
    assert len(host) == len(dev)
    for (rp, k, y), (rp2, k2, y2) in zip(host, dev):
        assert np.array_equal(rp, rp2)
        assert np.array_equal(k, k2)
        assert np.array_equal(y, y2)


@pytest.fixture(scope="module")
def trainer():
    t = api.Table(latent_dim=0, capacity=1 << 16)
    tr = api.Trainer(t, max_rows=1 << 19, max_nnz=1 << 21)
    yield tr
    tr.close()
    t.close()


@pytest.mark.parametrize("path", [TRAIN, TEST])
@pytest.mark.parametrize("block_bytes", [4096, 65536, 2 << 20])
def test_golden_files_parse_like_the_host_loader(trainer, path, block_bytes):
    host = _blocks_host(path, block_bytes)
    assert host, "fixture missing"
    _same_blocks(host, _blocks_device(path, block_bytes, trainer))


@pytest.mark.parametrize("path", [TRAIN, TEST])
@pytest.mark.parametrize("block_bytes", [4096, 65536, 2 << 20])
def test_golden_files_parse_like_the_reference_loader(trainer, path, block_bytes):
    """The device parser DIRECTLY against the oracle's restatement of LoadData::load_minibatch_hash_data_fread
    (which tests/test_oracle.py pins to the compiled reference): same blocks, rows, labels and std::hash keys —
    not only transitively through this repo's own host parser."""
    from oracle import oracle as O
    ref = [(rp.astype(np.uint32), k, y.astype(np.uint8)) for rp, k, y in O.load_blocks(path, block_bytes)]
    assert ref, "fixture missing"
    _same_blocks(ref, _blocks_device(path, block_bytes, trainer))


def test_two_phase_ingest_pipeline_equals_one_shot(trainer, tmp_path):
    """xf_trainer_ingest_begin / _end with a block in flight while the previous one is exported and trained on:
    the same CSR as the synchronous call, block after block (two device-side block buffers alternate)."""
    import ctypes as C
    row_ptr, ids, labels = datagen.make_ids(5, 6000, 20, 1 << 30, ragged=True)
    path = str(tmp_path / "pipe-00000")
    datagen.write_text(path, row_ptr, ids, labels)
    want = _blocks_device(path, 1 << 16, trainer)
    assert len(want) > 6
    lib = api.lib()
    ld = api.Loader(path, 1 << 16)
    text, ln, r, z = C.c_void_p(), C.c_uint64(), C.c_uint32(), C.c_uint32()
    assert lib.xf_loader_next_raw(ld.h, C.byref(text), C.byref(ln)) == 0
    assert lib.xf_trainer_ingest_begin(trainer.h, text, ln.value) == 0
    got = []
    while True:
        nt, nl = C.c_void_p(), C.c_uint64()
        assert lib.xf_loader_next_raw(ld.h, C.byref(nt), C.byref(nl)) == 0   # the other text buffer
        assert lib.xf_trainer_ingest_end(trainer.h, C.byref(r), C.byref(z)) == 0, lib.xf_last_error()
        trainer.step_ingested(0, r.value)                                    # asynchronous
        if nl.value:
            assert lib.xf_trainer_ingest_begin(trainer.h, nt, nl.value) == 0, lib.xf_last_error()
        got.append(trainer.ingested_export(r.value, z.value))               # the CURRENT block, while the next parses
        if not nl.value:
            break
    assert lib.xf_trainer_ingest_end(trainer.h, C.byref(r), C.byref(z)) != 0  # nothing in flight
    _same_blocks(want, got)


@pytest.mark.parametrize("pinned", [False, True], ids=["pageable", "page-locked"])
def test_two_blocks_in_flight(trainer, tmp_path, pinned):
    """Depth 2: begin(i+2); step(i); end(i+1).  The second outstanding block's text is copied into the buffer of
    the block being trained on, its parse is launched by the _end that retires that block; a third _begin is
    refused; every block arrives exactly as the synchronous call delivers it, and in order."""
    import ctypes as C
    row_ptr, ids, labels = datagen.make_ids(6, 7000, 20, 1 << 30, ragged=True)
    path = str(tmp_path / "pipe2-00000")
    datagen.write_text(path, row_ptr, ids, labels)
    want = _blocks_device(path, 1 << 16, trainer)
    assert len(want) > 6
    lib = api.lib()
    ld = api.Loader(path, 1 << 16)
    text, ln, r, z = C.c_void_p(), C.c_uint64(), C.c_uint32(), C.c_uint32()
    blocks, keep = [], []
    while True:
        assert lib.xf_loader_next_raw(ld.h, C.byref(text), C.byref(ln)) == 0
        if not ln.value:
            break
        raw = C.string_at(text, ln.value)
        if pinned:
            p = C.c_void_p()
            assert lib.xf_host_alloc(C.byref(p), len(raw) + 16) == 0
            C.memmove(p, raw, len(raw))
            blocks.append((p, len(raw)))
        else:
            keep.append(raw)
            blocks.append((C.c_char_p(raw), len(raw)))
    n = len(blocks)
    assert n == len(want)
    assert lib.xf_trainer_ingest_begin(trainer.h, blocks[0][0], blocks[0][1]) == 0
    assert lib.xf_trainer_ingest_begin(trainer.h, blocks[1][0], blocks[1][1]) == 0
    assert lib.xf_trainer_ingest_begin(trainer.h, blocks[2][0], blocks[2][1]) != 0      # two outstanding at most
    assert lib.xf_trainer_ingest_text(trainer.h, blocks[2][0], blocks[2][1], C.byref(r), C.byref(z)) != 0
    got = []
    for i in range(n):
        assert lib.xf_trainer_ingest_end(trainer.h, C.byref(r), C.byref(z)) == 0, lib.xf_last_error()
        trainer.step_ingested(0, r.value)                                    # asynchronous
        if i + 2 < n:
            assert lib.xf_trainer_ingest_begin(trainer.h, blocks[i + 2][0], blocks[i + 2][1]) == 0, lib.xf_last_error()
        got.append(trainer.ingested_export(r.value, z.value))               # block i, while i+1 parses and i+2 arrives
    assert lib.xf_trainer_ingest_end(trainer.h, C.byref(r), C.byref(z)) != 0  # nothing outstanding
    _same_blocks(want, got)
    if pinned:
        trainer.sync()
        for p, _ in blocks:
            lib.xf_host_free(p)


def test_synthetic_ragged_multi_block(trainer, tmp_path):
    row_ptr, ids, labels = datagen.make_ids(11, 20000, 24, 1 << 40, dist="zipf", ragged=True)
    path = str(tmp_path / "syn-00000")
    datagen.write_text(path, row_ptr, ids, labels)
    host = _blocks_host(path, 1 << 18)
    assert len(host) > 4
    _same_blocks(host, _blocks_device(path, 1 << 18, trainer))
    # and the hashes are the reference's std::hash of the decimal id strings
    keys = np.concatenate([k for _, k, _ in host])
    assert np.array_equal(keys, api.hash_decimal_ids(ids)[: keys.size])


EDGE_TEXTS = [
    b"",
    b"1\t3:17:1\n",
    b"1\t3:17:1",                                   # no trailing newline
    b"0\t1:a:1 2:bb:1 3:ccc:0.5\n1\t9:zz:1\n",
    b"1\t0:100:1 1:200:1\r\n0\t0:300:1\r\n",         # CRLF rows, as in the reference's sample data
    b"1.0\t0:5:1\n0.0\t0:6:1\n0.5\t0:7:1\n-1\t0:8:1\n1e-8\t0:9:1\n2e-7\t0:10:1\n",
    b"1\t0:5:1\n\n0\t0:6:1\n",                       # blank line between rows
    b"0\t" + b" ".join(b"%d:%d:1" % (i % 7, i * 2654435761 % (1 << 61)) for i in range(3000)) + b"\n",  # one long row
    b"".join(b"%d\t0:%d:1\n" % (i & 1, i) for i in range(5000)),                                        # many short rows
    b"1\t0::1 0:x:1\n",                              # empty fid hashes the empty string
    b"1\t0:" + b"k" * 700 + b":1 0:q:1\n",           # fid longer than a parser chunk
]


@pytest.mark.parametrize("i", range(len(EDGE_TEXTS)))
def test_edge_cases(trainer, tmp_path, i):
    text = EDGE_TEXTS[i]
    path = str(tmp_path / "edge-00000")
    with open(path, "wb") as f:
        f.write(text)
    host = _blocks_host(path, 1 << 20)
    dev = _blocks_device(path, 1 << 20, trainer)
    _same_blocks(host, dev)


def test_malformed_token_is_an_error(trainer):
    with pytest.raises(api.XflowError):
        trainer.ingest_text(b"1\t0:5:1 nocolon 0:6:1\n")
    # the trainer stays usable
    assert trainer.ingest_text(b"1\t0:5:1\n") == (1, 1)


def test_block_larger_than_trainer_limits_is_an_error():
    t = api.Table(latent_dim=0, capacity=1 << 12)
    tr = api.Trainer(t, max_rows=8, max_nnz=64)
    with pytest.raises(api.XflowError):
        tr.ingest_text(b"".join(b"1\t0:%d:1\n" % i for i in range(100)))
    tr.close()
    t.close()


@pytest.mark.parametrize("model,K", [(api.MODEL_LR, 0), (api.MODEL_FM, 4)])
@pytest.mark.parametrize("core_num", [1, 3])
def test_training_on_ingested_slices_matches_host_batches(model, K, core_num):
    """step_ingested on slices of a device-parsed block == step_host on the same slices."""
    tabs = []
    for ingest in (False, True):
        t = api.Table(latent_dim=K, capacity=1 << 16, v_init=api.VINIT_COUNTER, seed=5)
        tr = api.Trainer(t, model=model, max_rows=1 << 17, max_nnz=1 << 20)
        tr.init_push()
        ld = api.Loader(TRAIN, 1 << 16)
        while True:
            if ingest:
                text = ld.next_raw()
                if not text:
                    break
                rows, _ = tr.ingest_text(text)
                ts = rows // core_num
                for c in range(core_num):
                    tr.step_ingested(c * ts, (c + 1) * ts)
            else:
                try:
                    rp, keys, y = next(ld)
                except StopIteration:
                    break
                ts = (rp.size - 1) // core_num
                for c in range(core_num):
                    a, b = c * ts, (c + 1) * ts
                    tr.step_host((rp[a:b + 1] - rp[a]).astype(np.uint32), keys[rp[a]:rp[b]], y[a:b], want_loss=False)
        tr.sync()
        keys = np.sort(t.list_keys())
        tabs.append((keys, t.export(keys)))
        tr.close()
        t.close()
    (k0, e0), (k1, e1) = tabs
    assert np.array_equal(k0, k1)
    for name in e0:
        if e0[name] is None:
            continue
        assert_close(e1[name], e0[name], name)


def test_cli_device_ingest_equals_host_parse(tmp_path):
    """xflow_lr with the device parser (default) and with XFLOW_HOST_PARSE=1 write the same predictions."""
    exe = os.path.join(ROOT, "xflow_b200", "bin", "xflow_lr")
    outs = []
    for host_parse in ("0", "1"):
        d = tmp_path / ("hp" + host_parse)
        d.mkdir()
        env = dict(os.environ, XFLOW_HOST_PARSE=host_parse, XFLOW_CORE_NUM="2")
        r = subprocess.run([exe, os.path.join(GOLDEN, "data", "small_train"), os.path.join(GOLDEN, "data", "small_test"), "0", "3"],
                           cwd=str(d), env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        outs.append(((d / "pred_0_0.txt").read_text(), [l for l in r.stdout.splitlines() if "logloss" in l]))
    assert outs[0][0] == outs[1][0]
    assert outs[0][1] == outs[1][1]


@pytest.mark.parametrize("K", [0, 3])
def test_text_model_dump(tmp_path, K):
    """xf_table_dump_text: key-ordered "<key>\\t<w>[\\t<v...>]" lines that round-trip floats."""
    rng = np.random.default_rng(9)
    n = 1000
    keys = np.unique(rng.integers(1, 1 << 63, n, dtype=np.uint64))
    n = keys.size
    w = rng.standard_normal(n).astype(np.float32)
    w[::3] = 0.0
    v = rng.standard_normal((n, K)).astype(np.float32) if K else None
    if K:
        v[::3] = 0.0
        v[3::6] = 0.0   # some rows with w != 0 but v == 0
    t = api.Table(latent_dim=K, capacity=1 << 12)
    z = np.zeros(n, np.float32)
    t.import_(keys, w=w, nw=z, zw=z, v=v, nv=None if v is None else np.zeros_like(v), zv=None if v is None else np.zeros_like(v))
    for nonzero_only in (False, True):
        path = str(tmp_path / ("m%d.txt" % nonzero_only))
        lines = t.dump_text(path, nonzero_only=nonzero_only)
        got = [l.rstrip("\n").split("\t") for l in open(path)]
        assert lines == len(got)
        keep = np.ones(n, bool) if not nonzero_only else ((w != 0) | (v != 0).any(axis=1) if K else (w != 0))
        assert [int(g[0]) for g in got] == [int(k) for k in keys[keep]]        # key order
        assert np.array_equal(np.array([g[1] for g in got], np.float32), w[keep])
        if K:
            gv = np.array([[float(x) for x in g[2].split(" ")] for g in got], np.float32)
            assert np.array_equal(gv, v[keep])
    t.close()


def test_cli_blocks_with_featureless_rows(tmp_path):
    """Rows without features ("1\\t\\n", 3 bytes) put more rows into a block than the worker sized its
    trainer for (8 bytes per well-formed row): the worker re-sizes and parses again.  Same output as the
    host parser."""
    exe = os.path.join(ROOT, "xflow_b200", "bin", "xflow_lr")
    body = open(TRAIN, "rb").read()
    train = tmp_path / "sparse-00000"
    train.write_bytes(b"1\t\n" * 400000 + body + b"0\t\n" * 30000)  # first 1 MiB block: ~350 k rows
    test = tmp_path / "t-00000"
    test.write_bytes(b"0\t\n" * 5000 + open(TEST, "rb").read())
    outs = []
    for host_parse in ("0", "1"):
        d = tmp_path / ("hp" + host_parse)
        d.mkdir()
        env = dict(os.environ, XFLOW_HOST_PARSE=host_parse, XFLOW_BLOCK_MB="1")
        r = subprocess.run([exe, str(tmp_path / "sparse"), str(tmp_path / "t"), "0", "2"], cwd=str(d), env=env,
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        outs.append(((d / "pred_0_0.txt").read_text(), [l for l in r.stdout.splitlines() if "logloss" in l]))
    assert outs[0][1] and outs[0] == outs[1]
    assert outs[0][0].count("\n") == 5000 + 200
