"""Helpers shared by the test modules."""
import os

import numpy as np

from cases import CASES, SYN, SYN_TEST  # tests/golden/cases.py
from xflow_b200 import datagen

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")

REL_TOL = 1e-5   # north_star: "within 1e-5 relative on float logloss and learned weights"
ABS_FLOOR = 1e-7  # abs floor near 0 (FTRL produces exact zeros)


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def materialise_syn(tmp):
    tr, te = os.path.join(tmp, "syn_train"), os.path.join(tmp, "syn_test")
    datagen.write_text(tr + "-00000", *datagen.make_ids(**SYN))
    datagen.write_text(te + "-00000", *datagen.make_ids(**SYN_TEST))
    return tr, te


def data_prefixes(case, syn_data):
    if CASES[case]["data"] == "small":
        d = os.path.join(GOLDEN, "data")
        return os.path.join(d, "small_train"), os.path.join(d, "small_test")
    return syn_data


def close(a, b, rel=REL_TOL, abs_floor=ABS_FLOOR):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b) <= rel * np.abs(b) + abs_floor


def assert_close(a, b, what, rel=REL_TOL, abs_floor=ABS_FLOOR, max_bad_frac=0.0):
    """All (or all but max_bad_frac) elements within rel*|b| + abs_floor."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    ok = close(a, b, rel, abs_floor)
    bad = int((~ok).sum())
    if bad > max_bad_frac * max(ok.size, 1):
        idx = np.argwhere(~ok)[:5]
        detail = ", ".join("%s: got %.9g want %.9g" % (tuple(i), a[tuple(i)], b[tuple(i)]) for i in idx)
        raise AssertionError("%s: %d/%d outside tolerance (%s)" % (what, bad, ok.size, detail))


def bits_equal(a, b):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))
