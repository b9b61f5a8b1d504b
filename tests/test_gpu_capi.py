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
