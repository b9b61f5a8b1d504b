"""The reference-facing entry points on the GPU (pytest -m gpu): XFCreate / XFStartTrain
(src/c_api/c_api.h:26-29), the xflow_lr CLI (src/model/main.cc) and the C++ worker classes behind them,
run as separate processes (one Server per process, like the reference) and compared with what the
reference itself printed and wrote for the same shards (tests/golden/*.npz)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from common import GOLDEN, golden
from xflow_b200 import api

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TRAIN = os.path.join(GOLDEN, "data", "small_train")
TEST = os.path.join(GOLDEN, "data", "small_test")

CAPI_SCRIPT = r"""
import ctypes, sys
lib = ctypes.CDLL(sys.argv[1])
h = ctypes.c_void_p()
model, opt, K, epochs = int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6]), int(sys.argv[7])
if model == 0 and opt < 0:
    rc = lib.XFCreate(ctypes.byref(h), sys.argv[2].encode(), sys.argv[3].encode())
else:
    rc = lib.XFCreateEx(ctypes.byref(h), sys.argv[2].encode(), sys.argv[3].encode(), model, opt, K, epochs)
assert rc == 0, rc
rc = lib.XFStartTrain(ctypes.byref(h))
lib.xf_last_error.restype = ctypes.c_char_p
assert rc == 0, lib.xf_last_error()
assert lib.XFDestroy(ctypes.byref(h)) == 0
"""


def _parse(stdout):
    m = re.search(r"logloss: (\S+)\s+auc = (\S+)\s+tp = (\d+) fp = (\d+)", stdout)
    assert m, stdout
    return float(m.group(1)), float(m.group(2)), int(m.group(3)), int(m.group(4))


def _check(case, stdout, cwd):
    g = golden(case)
    ll, auc, tp, fp = _parse(stdout)
    assert abs(ll - float(g["logloss"])) <= 2e-5 * abs(float(g["logloss"])) + 1e-6
    assert abs(auc - float(g["auc"])) <= 2e-5
    assert tp + fp == g["pred_label"].size and tp == int(g["pred_label"].sum())
    pred = np.loadtxt(os.path.join(cwd, "pred_0_0.txt"), ndmin=2)
    assert np.array_equal(pred[:, 2].astype(np.int32), g["pred_label"])
    assert np.all(np.abs(pred[:, 0] - g["pred_pctr"]) <= 2e-5 * np.abs(g["pred_pctr"]) + 1.1e-6)
    assert "train end......" in stdout and "my rank is = 0" in stdout


@pytest.mark.parametrize("case,model,opt,K,epochs", [
    ("small_lr_ftrl_e60", 0, -1, 0, 0),        # plain XFCreate: LR, FTRL, 60 epochs (the reference defaults)
    ("small_lr_sgd_e60", 0, 1, 0, 60),         # BASELINE configs[0]: LR + SGD on the bundled shards
    ("small_fm_sgd_k10_e60", 1, 1, 10, 60),
    ("small_lr_ftrl_e10", 0, 0, 0, 10),
])
def test_c_api_reproduces_reference_run(case, model, opt, K, epochs, tmp_path):
    env = dict(os.environ, XFLOW_RANK="0")
    r = subprocess.run([sys.executable, "-c", CAPI_SCRIPT, api.LIB_PATH, TRAIN, TEST, str(model), str(opt), str(K),
                        str(epochs)], cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    _check(case, r.stdout, str(tmp_path))


def test_cli_binary_matches_reference_lr_sgd(tmp_path):
    exe = os.path.join(ROOT, "xflow_b200", "bin", "xflow_lr")
    assert os.path.exists(exe), "xflow_lr not built"
    env = dict(os.environ, XFLOW_OPTIMIZER="sgd")
    r = subprocess.run([exe, TRAIN, TEST, "0", "60"], cwd=str(tmp_path), env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "start LR" in r.stdout
    _check("small_lr_sgd_e60", r.stdout, str(tmp_path))


def test_cli_slices_drop_remainder_like_reference(tmp_path):
    """core_num = 8 on a 200-row block: 8 slices of 25 rows (lr_worker.cc:190-196); the oracle's slice loop is the yardstick."""
    from oracle import oracle as O
    exe = os.path.join(ROOT, "xflow_b200", "bin", "xflow_lr")
    env = dict(os.environ, XFLOW_OPTIMIZER="ftrl", XFLOW_CORE_NUM="7")  # 200 // 7 = 28 rows per slice, 4 rows dropped
    r = subprocess.run([exe, TRAIN, TEST, "0", "5"], cwd=str(tmp_path), env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    ll, auc, tp, fp = _parse(r.stdout)
    t = O.Table()
    O.train_file(t, TRAIN + "-00000", 2 << 20, 5, slices=7)
    lab, p = O.predict_file(t, TEST + "-00000", 4 << 20, slices=7)
    m = O.auc_logloss(lab, p)
    assert tp + fp == lab.size == 196
    assert abs(ll - m["logloss"]) <= 2e-5 * abs(m["logloss"]) + 1e-6
    assert abs(auc - m["auc"]) <= 2e-5


def test_ps_lite_shaped_surface_on_the_device(tmp_path):
    """include/xflow/ps_compat.h: ps::KVWorker Push / Pull through the reference's optimizer functors
    (FTRL::KVServerFTRLHandle_w / _v installed on ps::KVServer 0 / 1, server.h:22-31) against the device table:
    insert-on-pull, the first FTRL step in closed form, size checks, optimizer mismatch (tests/cxx)."""
    from common import build_and_run_ps_compat
    r = build_and_run_ps_compat(tmp_path)
    assert r.returncode == 0 and "transport ok" in r.stdout and "device handles ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("model,K", [("0", 0), ("1", 10)])
def test_cli_two_ranks_train_against_the_sharded_table(model, K, tmp_path):
    """The reference's launch shape (local.sh: N worker processes, rank = ps::MyRank(), every rank trains
    <train>-%05d against the SHARED servers, main.cc:15-48, lr_worker.cc:207-217) on two GPUs: two xflow_lr processes
    with XFLOW_RANK / XFLOW_WORLD find each other through a file, each owns one key range of the table, rank 0
    predicts through the sharded forward pass.  Compared with the oracle's lock-step schedule on one table."""
    if api.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from oracle import oracle as O
    exe = os.path.join(ROOT, "xflow_b200", "bin", "xflow_lr")
    # two train shards: the golden shard cut in two (rank 1 gets fewer rows AND its own block count)
    lines = open(TRAIN + "-00000", "rb").read().splitlines(keepends=True)
    prefix = str(tmp_path / "train")
    open(prefix + "-00000", "wb").write(b"".join(lines[:120]))
    open(prefix + "-00001", "wb").write(b"".join(lines[120:]))
    epochs = 5
    procs = []
    for rank in range(2):
        cwd = tmp_path / ("rank%d" % rank)
        cwd.mkdir()
        env = dict(os.environ, XFLOW_OPTIMIZER="ftrl", XFLOW_RANK=str(rank), XFLOW_WORLD="2", XFLOW_DEVICE=str(rank),
                   XFLOW_COMM_FILE=str(tmp_path / "comm.id"), XFLOW_MG_TIMEOUT_S="60", XFLOW_SEED="0")
        procs.append((cwd, subprocess.Popen([exe, prefix, TEST, model, str(epochs)], cwd=str(cwd), env=env,
                                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    outs = []
    for cwd, p in procs:
        out, _ = p.communicate(timeout=600)
        assert p.returncode == 0, out
        outs.append(out)
    assert "my rank is = 0" in outs[0] and "my rank is = 1" in outs[1]
    assert "logloss" in outs[0] and "logloss" not in outs[1]          # only rank 0 predicts and prints
    assert not os.path.exists(str(tmp_path / "rank1" / "pred_1_0.txt")) or os.path.getsize(str(tmp_path / "rank1" / "pred_1_0.txt")) == 0
    ll, auc, tp, fp = _parse(outs[0])
    # the oracle: one table, every worker computes from the same state, pushes land in rank order
    t = O.Table(K=K, opt=O.OPT_FTRL, init_mode=O.INIT_COUNTER, seed=0)
    for _ in range(2):
        t.init_push()
    shards = [next(iter(O.load_blocks(prefix + "-%05d" % r, 2 << 20))) for r in range(2)]
    for _ in range(epochs):
        pend = [t.worker_compute(rp, keys, lab) for rp, keys, lab in shards]
        for uk, gw, gv, _ in pend:
            t.push(uk, gw, gv if K else None)
    lab, p = O.predict_file(t, TEST + "-00000", (4 << 20) if K == 0 else (2 << 20))
    m = O.auc_logloss(lab, p)
    assert tp + fp == lab.size
    assert abs(ll - m["logloss"]) <= 2e-5 * abs(m["logloss"]) + 1e-6
    assert abs(auc - m["auc"]) <= 2e-5
    pred = np.loadtxt(str(tmp_path / "rank0" / "pred_0_0.txt"), ndmin=2)
    assert np.all(np.abs(pred[:, 0] - p) <= 2e-5 * np.abs(p) + 1.1e-6)


def test_rank_without_world_is_refused(tmp_path):
    exe = os.path.join(ROOT, "xflow_b200", "bin", "xflow_lr")
    env = dict(os.environ, XFLOW_RANK="1")
    env.pop("XFLOW_WORLD", None); env.pop("WORLD_SIZE", None)
    r = subprocess.run([exe, TRAIN, TEST, "0", "1"], cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "XFLOW_WORLD" in (r.stdout + r.stderr)


def test_cli_device_metric_prints_the_same_numbers(tmp_path):
    """XFLOW_DEVICE_METRIC=1: logloss / AUC computed on the device (metric.cu) — the same quantities as the reference's
    Base::calculate_auc without its float accumulators; on the bundled shard they agree with the reference's printout
    to the tolerance its own float arithmetic leaves (ties between rows of different labels are the only other
    difference: the reference orders them by std::sort, the device keeps input order)."""
    exe = os.path.join(ROOT, "xflow_b200", "bin", "xflow_lr")
    outs = {}
    for dev in ("0", "1"):
        d = tmp_path / ("m" + dev)
        d.mkdir()
        env = dict(os.environ, XFLOW_OPTIMIZER="ftrl", XFLOW_DEVICE_METRIC=dev, XFLOW_EXACT_METRIC="1")
        r = subprocess.run([exe, TRAIN, TEST, "0", "10"], cwd=str(d), env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        outs[dev] = (_parse(r.stdout), re.search(r"exact: logloss\(ln\) = (\S+)\s+auc = (\S+)", r.stdout))
    (ll0, auc0, tp0, fp0), ex0 = outs["0"]
    (ll1, auc1, tp1, fp1), ex1 = outs["1"]
    assert (tp0, fp0) == (tp1, fp1)
    assert abs(ll0 - ll1) <= 2e-5 * abs(ll0) + 1e-6
    assert abs(auc0 - auc1) <= 5e-3            # tie order only
    assert abs(float(ex0.group(1)) - float(ex1.group(1))) <= 1e-6 and abs(float(ex0.group(2)) - float(ex1.group(2))) <= 1e-9
