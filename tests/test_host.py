"""CPU tests of the product's host side: the C-ABI library loads, exports every symbol the header
declares, and its host-only pieces (hashing, block loader, metric, shard rule) agree with the oracle.
No compute entry point is exercised here (there is no GPU); those fail loudly, which is also checked."""
import os
import re
import subprocess

import numpy as np
import pytest

from common import GOLDEN
from oracle import oracle as O
from xflow_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "xflow_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"XF_DLL\s+[\w\s\*]+?\b(\w+)\s*\(", txt)))


def test_library_exports_every_header_symbol():
    syms = _header_symbols()
    assert len(syms) >= 40
    out = subprocess.check_output(["nm", "-D", "--defined-only", api.LIB_PATH], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [s for s in syms if s not in exported]
    assert not missing, "declared in include/xflow_b200.h but not exported: %s" % missing
    # and the ctypes table covers the same set
    assert sorted(api.SIGNATURES) == syms


def test_library_loads_and_reports_version():
    L = api.lib()
    assert L.xf_version() >= 100
    assert api.device_count() >= 0


@pytest.mark.skipif(api.device_count() > 0, reason="checks the no-GPU failure mode")
def test_compute_calls_fail_loudly_without_gpu():
    with pytest.raises(api.XflowError):
        api.Table()


def test_hash_matches_oracle_and_std_hash():
    g = np.load(os.path.join(GOLDEN, "std_hash.npz"))
    for s, h in zip(g["strings"], g["hashes"]):
        assert api.hash_bytes(bytes(s)) == int(h)
    rng = np.random.default_rng(3)
    for n in list(range(0, 34)) + [100]:
        s = bytes(rng.integers(1, 255, n, dtype=np.uint8))
        assert api.hash_bytes(s) == O.std_hash(s)
    ids = rng.integers(0, 10 ** 15, 50000).astype(np.uint64)
    ids[:3] = [0, 9, 10]
    assert np.array_equal(api.hash_decimal_ids(ids), O.hash_decimal_ids(ids))


def test_shard_rule_matches_oracle():
    rng = np.random.default_rng(4)
    keys = rng.integers(0, 2 ** 64, 5000, dtype=np.uint64)
    edge = [0, 1, 2 ** 64 - 1, 2 ** 64 - 8, 2 ** 64 - 9, 0x1FFFFFFFFFFFFFFF, 0x1FFFFFFFFFFFFFFE, 0x2000000000000000]
    for S in (1, 2, 3, 4, 8):
        for k in list(keys[:500]) + edge:
            assert api.shard_of(int(k), S) == O.shard_of(int(k), S)


@pytest.mark.parametrize("block", [2 << 20, 1 << 20, 65536, 4096, 1000, 700])
def test_loader_blocks_match_oracle(block, syn_data):
    for path in (os.path.join(GOLDEN, "data", "small_train-00000"), syn_data[0] + "-00000"):
        if block < 4096 and "syn" in path:
            continue
        a = list(api.Loader(path, block))
        b = list(O.load_blocks(path, block))
        assert len(a) == len(b) and len(a) > 0
        for (rp, k, l), (rp2, k2, l2) in zip(a, b):
            assert np.array_equal(rp.astype(np.int64), rp2)
            assert np.array_equal(k, k2)
            assert np.array_equal(l.astype(np.int32), l2)


def test_loader_edge_cases(tmp_path):
    # CRLF line ends, float labels, trailing space, missing final newline, empty file
    p = tmp_path / "edge-00000"
    p.write_bytes(b"1\t0:12:0.5 1:34:0.5\r\n0.0\t2:56:1 \n1e-9\t3:7:1\n0.5\t4:8:1 5:9:1")
    a = list(api.Loader(str(p), 1 << 20))
    b = list(O.load_blocks(str(p), 1 << 20))
    assert len(a) == 1 and len(b) == 1
    rp, k, l = a[0]
    assert list(l) == [1, 0, 0, 1]
    assert list(rp) == [0, 2, 3, 4, 6]
    assert np.array_equal(k, b[0][1]) and np.array_equal(rp.astype(np.int64), b[0][0])
    assert int(k[0]) == O.std_hash(b"12") and int(k[5]) == O.std_hash(b"9")
    e = tmp_path / "empty-00000"
    e.write_bytes(b"")
    assert list(api.Loader(str(e), 1 << 20)) == []
    with pytest.raises(api.XflowError):
        api.Loader(str(tmp_path / "missing-00000"), 1 << 20)


def test_auc_logloss_matches_oracle():
    rng = np.random.default_rng(5)
    for n in (1, 2, 200, 5000):
        lab = (rng.random(n) < 0.3).astype(np.int32)
        p = rng.random(n).astype(np.float32) * 0.98 + 0.01
        p[: n // 4] = p[0]  # ties
        a, b = api.auc_logloss(lab, p), O.auc_logloss(lab, p)
        assert a["tp"] == b["tp"] and a["fp"] == b["fp"]
        assert a["logloss"] == b["logloss"]
        assert (np.isnan(a["auc"]) and np.isnan(b["auc"])) or a["auc"] == b["auc"]


def test_exact_metric_matches_sklearn():
    """xf_auc_logloss_exact: tie-aware AUC and natural-log logloss in exact arithmetic."""
    from sklearn.metrics import log_loss, roc_auc_score
    rng = np.random.default_rng(3)
    n = 20000
    y = (rng.random(n) < 0.3).astype(np.int32)
    p = np.clip(np.round(rng.random(n) * 0.6 + y * 0.2, 2), 0.01, 0.99).astype(np.float32)  # many ties
    m = api.auc_logloss_exact(y, p)
    assert m["positives"] == int(y.sum()) and m["negatives"] == int(n - y.sum())
    assert abs(m["auc"] - roc_auc_score(y, p)) < 1e-12
    assert abs(m["logloss"] - log_loss(y, p.astype(np.float64))) < 1e-9
    # single-class input has no AUC
    assert np.isnan(api.auc_logloss_exact(np.ones(5, np.int32), np.full(5, 0.5, np.float32))["auc"])
    # and the reference-faithful metric is a different quantity: base-2, not negated
    ref = api.auc_logloss(y, p)
    assert abs(ref["logloss"] * np.log(2) + m["logloss"]) < 1e-3


def test_loader_large_blocks_read_in_parallel_pieces(tmp_path):
    """Blocks of several MiB are filled by a few pread threads; the rows must not depend on that."""
    from xflow_b200 import datagen
    rp, ids, lab = datagen.make_ids(5, 30000, 30, 1 << 40, ragged=True)
    path = str(tmp_path / "big-00000")
    datagen.write_text(path, rp, ids, lab)
    assert os.path.getsize(path) > (8 << 20)
    def all_rows(block):
        ks, ys, lens = [], [], []
        for r, k, y in api.Loader(path, block):
            ks.append(k); ys.append(y); lens.append(np.diff(r))
        return np.concatenate(ks), np.concatenate(ys), np.concatenate(lens)
    small = all_rows(1 << 20)          # single-threaded reads
    for block in (5 << 20, 64 << 20):  # multi-threaded, with and without a carried tail
        big = all_rows(block)
        for a, b in zip(small, big):
            assert np.array_equal(a, b)
    assert np.array_equal(small[0], api.hash_decimal_ids(ids))
    assert np.array_equal(small[1], lab)


def _parse_block_py(text):
    """The reference's row / token rules (load_data_from_disk.cc:126-209) in plain Python."""
    rows = []
    for line in text.split(b"\n"):
        if not line:
            continue
        label, _, rest = line.partition(b"\t")
        try:
            y = 1 if np.float32(float(label.strip() or b"0")) > 1e-7 else 0
        except ValueError:
            y = 0
        toks = [t for t in rest.replace(b"\r", b"").split(b" ") if t]
        rows.append((y, [api.hash_bytes(t.split(b":")[1]) for t in toks]))
    return rows


def _parse_block_device_rules(t):
    """The per-byte rules of the device parser (ingest.cu: xf_is_line_start / xf_is_tok_start / label up to
    the tab / fid between the first two colons), restated byte by byte."""
    n, rows = len(t), []
    for p in range(n):
        c = t[p:p + 1]
        if (p == 0 or t[p - 1:p] == b"\n") and c != b"\n":
            q = p
            while q < n and t[q:q + 1] not in (b"\t", b"\n"):
                q += 1
            try:
                y = 1 if np.float32(float(t[p:q].strip() or b"0")) > 1e-7 else 0
            except ValueError:
                y = 0
            rows.append((y, []))
        if p > 0 and t[p - 1:p] in (b" ", b"\t") and c not in (b" ", b"\n", b"\t", b"\r"):
            q, colons = p, []
            while q < n and t[q:q + 1] not in (b" ", b"\n"):
                if t[q:q + 1] == b":":
                    colons.append(q)
                    if len(colons) == 2:
                        break
                q += 1
            assert len(colons) == 2
            rows[-1][1].append(api.hash_bytes(t[colons[0] + 1:colons[1]]))
    return rows


@pytest.mark.parametrize("block_bytes", [4096, 1 << 16, 1 << 22])
def test_raw_blocks_are_the_parsed_blocks(block_bytes):
    """xf_loader_next_raw (block formation only, what the device parser is fed) cuts the file exactly where
    xf_loader_next does: parsing each raw block in Python gives the rows of the corresponding CSR block."""
    path = os.path.join(GOLDEN, "data", "small_train-00000")
    parsed = list(api.Loader(path, block_bytes))
    raw = []
    ld = api.Loader(path, block_bytes)
    while True:
        t = ld.next_raw()
        if not t:
            break
        raw.append(t)
    assert len(raw) == len(parsed)
    for text, (rp, keys, lab) in zip(raw, parsed):
        rows = _parse_block_py(text)
        assert [y for y, _ in rows] == lab.tolist()
        assert [len(k) for _, k in rows] == np.diff(rp).tolist()
        assert [h for _, k in rows for h in k] == keys.tolist()
    # every byte of the file is in exactly one block (up to the newline at each cut)
    whole = open(path, "rb").read()
    assert b"".join(r.rstrip(b"\n") for r in raw).replace(b"\n", b"") == whole.replace(b"\n", b"")


def test_loader_rewind_and_streams(tmp_path):
    """One loader serves every epoch (xf_loader_rewind = re-opening the file, lr_worker.cc:184), and a FIFO —
    which reports size 0 to fstat — is read sequentially to EOF like the reference's fread does."""
    import threading
    path = os.path.join(GOLDEN, "data", "small_train-00000")
    first = [tuple(a.copy() for a in blk) for blk in api.Loader(path, 8192)]
    ld = api.Loader(path, 8192)
    again = []
    for _ in range(2):
        again.append([tuple(a.copy() for a in blk) for blk in ld])
        assert api.lib().xf_loader_rewind(ld.h) == 0
    for run in again:
        assert len(run) == len(first)
        for a, b in zip(run, first):
            assert all(np.array_equal(x, y) for x, y in zip(a, b))
    fifo = str(tmp_path / "pipe-00000")
    os.mkfifo(fifo)
    data = open(path, "rb").read()

    def feed():
        with open(fifo, "wb") as f:
            f.write(data)
    th = threading.Thread(target=feed)
    th.start()
    got = [tuple(a.copy() for a in blk) for blk in api.Loader(fifo, 8192)]
    th.join()
    assert len(got) == len(first)
    for a, b in zip(got, first):
        assert all(np.array_equal(x, y) for x, y in zip(a, b))


def test_loader_fuzz_against_python_parser(tmp_path):
    """Random well-formed shards (random labels incl. floats, 0..6 tokens per row, fids of random bytes and
    lengths, CRLF or LF, with / without final newline) cut at random block sizes: the rows, labels and
    hashes of xf_loader_next equal a plain-Python parse of the whole file."""
    from hypothesis import given, settings, strategies as st
    alphabet = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789_-."
    fid = st.text(alphabet=alphabet, min_size=1, max_size=24)
    tok = st.builds(lambda f, i, v: "%d:%s:%s" % (f, i, v), st.integers(0, 40), fid,
                    st.sampled_from(["1", "0.5", "3", "1e-3"]))
    label = st.sampled_from(["0", "1", "1.0", "0.0", "0.25", "-1", "1e-9", "2e-7"])
    row = st.tuples(label, st.lists(tok, min_size=0, max_size=6))
    path = str(tmp_path / "fuzz-00000")

    @settings(max_examples=60, deadline=None)
    @given(st.lists(row, min_size=1, max_size=60), st.booleans(), st.booleans(), st.integers(64, 2048))
    def check(rows, crlf, final_newline, block_bytes):
        eol = "\r\n" if crlf else "\n"
        text = eol.join("%s\t%s" % (l, " ".join(t)) for l, t in rows) + (eol if final_newline else "")
        data = text.encode()
        longest = max(len(x) for x in data.split(b"\n")) + 2
        if block_bytes <= longest:       # a block must be able to hold the longest row (as in the reference)
            block_bytes = longest + 1
        with open(path, "wb") as f:
            f.write(data)
        want = _parse_block_py(data)
        assert _parse_block_device_rules(data) == want      # host and device parsers implement one rule set
        got_y, got_len, got_k = [], [], []
        for rp, keys, lab in api.Loader(path, block_bytes):
            got_y += lab.tolist()
            got_len += np.diff(rp).tolist()
            got_k += keys.tolist()
        assert got_y == [y for y, _ in want]
        assert got_len == [len(k) for _, k in want]
        assert got_k == [h for _, k in want for h in k]

    check()


def test_ps_compat_header(tmp_path):
    """include/xflow/ps_compat.h (ps::KVServer / KVWorker / KVPairs / KVMeta and the reference's four optimizer
    functors over the device table): compiles as plain C++14 against the shipped library, its in-process
    transport behaves like ps-lite's request / response contract, and the device handles fail loudly when
    there is no GPU (on a GPU box the same program checks the first FTRL step through Push / Pull)."""
    from common import build_and_run_ps_compat
    r = build_and_run_ps_compat(tmp_path)
    assert "transport ok" in r.stdout, r.stdout + r.stderr
    if api.device_count() > 0:
        assert r.returncode == 0 and "device handles ok" in r.stdout, r.stdout + r.stderr
    else:
        assert r.returncode == 3 and "failed loudly" in r.stdout, r.stdout + r.stderr
