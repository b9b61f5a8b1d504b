"""CPU model of the lazy LR table's protocol (xflow_b200/csrc/table.cuh: xf_lazy_fold / xf_lazy_deposit,
step_lazy.cu phases A and B) under arbitrary interleavings of the warps of a batch.

The CUDA kernels themselves are checked on the GPU (tests/test_gpu_parity.py, incl. a contention stress test);
this file checks the DESIGN they implement, where every interleaving can be enumerated cheaply:

  * a row is  {n, z | tag, g}:  FTRL state and, if tag != 0, the integer residual sum g of batch `tag` whose
    optimizer step has not been applied yet; the weight is the closed form f(z, n) (ftrl.h:66-74);
  * phase A of a data row snapshots the rows of its tokens and computes, without writing, the weight the batch
    pulls (state with the pending step applied);
  * phase B makes one compare-and-swap per token  (snapshot) -> (folded state, tag = this batch, own residual);
    a token whose CAS fails finds the row open for this batch and adds its residual to g (integer add).

Claims checked: (1) every token of a batch pulls the same weight for a key whenever its snapshot is taken;
(2) after any interleaving the table equals the eager semantics (one FTRL step per touched key and batch with
gradient float32(sum of residuals) / rows, lr_worker.cc:116-118) bit for bit; (3) therefore two different
interleavings give bit-identical tables."""
import numpy as np
import pytest

F = np.float32
ALPHA, BETA, L1, L2 = F(5e-2), F(1.0), F(5e-5), F(10.0)
FIX = 134217728.0  # 2^27


def ftrl_w(z, n):
    if abs(z) <= L1:
        return F(0.0)
    tmpr = F(z - L1) if z > 0 else F(z + L1)
    tmpl = F(-F(F(F(BETA + np.sqrt(n, dtype=F)) / ALPHA) + L2))
    return F(tmpr / tmpl)


def ftrl_coord(g, w, n, z):
    nn = F(n + F(g * g))
    sig = F(F(np.sqrt(nn, dtype=F) - np.sqrt(n, dtype=F)) / ALPHA)
    z = F(z + F(g - F(sig * w)))
    return ftrl_w(z, nn), nn, z


def grad_of(gfix, rows):
    return F(np.float64(F(gfix / FIX)) / rows)      # the sum is rounded to float once, then divided in double


class LazyTable:
    def __init__(self):
        self.rows = {}            # key -> [n, z, tag, gfix]
        self.rows_by_seq = {}

    def snapshot(self, key):
        return tuple(self.rows.setdefault(key, [F(0), F(0), 0, 0]))

    @staticmethod
    def fold(snap, seq, rows_by_seq):
        n, z, tag, gfix = snap
        w = ftrl_w(z, n)
        if tag != 0 and tag != seq:
            w, n, z = ftrl_coord(grad_of(gfix, rows_by_seq[tag]), w, n, z)
        return w, n, z

    def deposit(self, key, snap, seq, fix):
        cur = self.rows[key]
        if snap[2] == seq:                       # the snapshot already saw the row open for this batch
            assert cur[2] == seq
            cur[3] += fix
            return
        if tuple(cur) == snap:                   # CAS succeeds: fold + open + deposit in one transition
            _, n, z = self.fold(snap, seq, self.rows_by_seq)
            self.rows[key] = [n, z, seq, fix]
            return
        assert cur[2] == seq, "a failed CAS must find the row open for this batch"
        cur[3] += fix

    def flushed(self):
        out = {}
        for k, snap in self.rows.items():
            w, n, z = self.fold(tuple(snap), -1, self.rows_by_seq)
            out[k] = (w, n, z)
        return out


def run_lazy(batches, rng):
    t = LazyTable()
    pulled_log = []
    for seq, (rows_tokens, labels) in enumerate(batches, start=1):
        B = len(rows_tokens)
        t.rows_by_seq[seq] = float(B)
        # every data row: phase A, later phase B; the two lists are interleaved at random across data rows
        events = [("A", r) for r in range(B)] + [("B", r) for r in range(B)]
        order = rng.permutation(len(events))
        pos = {e: i for i, e in zip(np.argsort(order), events)}
        sched = sorted(events, key=lambda e: (pos[e] if e[0] == "A" else max(pos[e], pos[("A", e[1])] + 0.5)))
        snaps, resid = {}, {}
        pulled = {}
        for kind, r in sched:
            if kind == "A":
                snaps[r] = [t.snapshot(k) for k in rows_tokens[r]]
                ws = [t.fold(s, seq, t.rows_by_seq)[0] for s in snaps[r]]
                for k, w in zip(rows_tokens[r], ws):
                    assert pulled.setdefault(k, w) == w, "a key must pull one weight per batch"   # claim (1)
                wx = F(0)
                for w in ws:
                    wx = F(wx + w)
                p = 1.0 / (1.0 + np.exp(-np.float64(wx)))
                resid[r] = F(F(p) - F(labels[r]))
            else:
                fix = int(np.rint(np.float64(resid[r]) * FIX))
                for k, s in zip(rows_tokens[r], snaps[r]):
                    t.deposit(k, s, seq, fix)
        pulled_log.append((pulled, dict(resid)))
    return t.flushed(), pulled_log


def run_eager(batches, pulled_log):
    state = {}
    for (rows_tokens, labels), (pulled, resid) in zip(batches, pulled_log):
        B = len(rows_tokens)
        gsum = {}
        for r, toks in enumerate(rows_tokens):
            fix = int(np.rint(np.float64(resid[r]) * FIX))
            for k in toks:
                gsum[k] = gsum.get(k, 0) + fix
        for k, gfix in gsum.items():
            w, n, z = state.get(k, (F(0), F(0), F(0)))
            assert w == pulled[k], "the weight a batch pulls is the eager table's weight"
            state[k] = ftrl_coord(grad_of(gfix, float(B)), w, n, z)
    return state


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_lazy_protocol_equals_eager_semantics_under_any_interleaving(seed):
    rng = np.random.default_rng(seed)
    batches = []
    for _ in range(6):
        B = int(rng.integers(3, 24))
        rows_tokens = [list(rng.zipf(1.6, int(rng.integers(0, 9))) % 37) for _ in range(B)]   # hot keys, duplicates, empty rows
        labels = rng.integers(0, 2, B)
        batches.append((rows_tokens, labels))
    a, log_a = run_lazy(batches, np.random.default_rng(100 + seed))
    b, _ = run_lazy(batches, np.random.default_rng(200 + seed))
    eager = run_eager(batches, log_a)
    assert set(a) == set(eager)
    for k in a:
        assert tuple(x.tobytes() for x in a[k]) == tuple(x.tobytes() for x in eager[k]), k      # claim (2)
        assert tuple(x.tobytes() for x in a[k]) == tuple(x.tobytes() for x in b[k]), k          # claim (3)
