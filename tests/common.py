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


def assert_close_noise_aware(got, ref, exact, what, rel=REL_TOL, abs_floor=ABS_FLOOR, max_noisy_frac=2e-3,
                             noise_factor=8.0):
    """Parity against the reference where the reference's own float32 summation order is part of its
    result.  A key that occurs thousands of times in one batch gets its gradient summed sequentially
    in float by the reference (lr_worker.cc:108-113), in the order its unstable std::sort left the
    occurrences; that sum carries rounding noise far above 1e-5 that no other summation order can
    reproduce.  `exact` is the same algorithm with the per-key sums accumulated in double
    (oracle.exact_sums), so |ref - exact| measures that noise element by element.  Required:
      * every element within rel*|ref| + abs_floor of the reference, OR within noise_factor times
        the reference's own measured noise;
      * the second clause is needed by at most max_noisy_frac of the elements."""
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    exact = np.asarray(exact, np.float64)
    assert got.shape == ref.shape == exact.shape, what
    strict = close(got, ref, rel, abs_floor)
    noise = np.abs(ref - exact)
    loose = np.abs(got - ref) <= noise_factor * noise + rel * np.abs(ref) + abs_floor
    bad = ~(strict | loose)
    if bad.any():
        idx = np.argwhere(bad)[:5]
        detail = ", ".join("%s: got %.9g ref %.9g exact %.9g" % (tuple(i), got[tuple(i)], ref[tuple(i)],
                                                                exact[tuple(i)]) for i in idx)
        raise AssertionError("%s: %d/%d outside tolerance AND outside the reference's own noise (%s)"
                             % (what, int(bad.sum()), bad.size, detail))
    noisy = int((~strict).sum())
    assert noisy <= max_noisy_frac * strict.size, "%s: %d/%d elements needed the noise clause" % (
        what, noisy, strict.size)
    return noisy


def oracle_case_run(case, syn_data, exact=False):
    """Run a golden case through the CPU restatement.  exact=True accumulates gradient sums in double
    (noise yardstick, see assert_close_noise_aware).  Returns (export dict on the golden keys, labels, pctr)."""
    from oracle import oracle as O
    c = CASES[case]
    g = golden(case)
    train, test = data_prefixes(case, syn_data)
    opt = O.OPT_FTRL if c["opt"] == "ftrl" else O.OPT_SGD
    if c.get("preinit"):
        t = O.Table(K=c["K"], opt=opt, init_mode=O.INIT_ZERO)
        t.import_(g["keys"], w=g["init_w"], v=g["init_v"])
    else:
        t = O.Table(K=c["K"], opt=opt)
    block = c.get("block_mb", 2) << 20
    if exact:
        with O.exact_sums():
            O.train_file(t, train + "-00000", block, c["epochs"])
    else:
        O.train_file(t, train + "-00000", block, c["epochs"])
    lab, p = O.predict_file(t, test + "-00000", (4 << 20) if c["model"] == "lr" else (2 << 20))
    return t.export(g["keys"]), lab, p
