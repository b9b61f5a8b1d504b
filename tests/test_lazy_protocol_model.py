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


# ---------------------------------------------------------------------------------------------------
# the sharded Push (mg_kernels.cu xf_k_push_tokens_lr): sources in rank order, every (round, source) pair its
# own batch number, deposits made from the look at the row that the round's Pull stashed; a token whose CAS
# fails either finds the row open for its own (round, source) -> integer add, or opened by an EARLIER source of
# the round -> it continues from the words the failed CAS returned.
# ---------------------------------------------------------------------------------------------------
def run_sharded(rounds, rng):
    t = LazyTable()
    seq = 0
    log = []
    for sources in rounds:                                   # sources[s] = (rows_tokens, labels) of rank s
        # Pull of the round, all sources: the stash is the raw row, the answer has the pending step applied
        stash, resid = [], []
        for rows_tokens, labels in sources:
            st = [[t.snapshot(k) for k in toks] for toks in rows_tokens]
            rs = []
            for toks, snaps, y in zip(rows_tokens, st, labels):
                wx = F(0)
                for s_ in snaps:
                    wx = F(wx + t.fold(s_, -1, t.rows_by_seq)[0])
                rs.append(F(F(1.0 / (1.0 + np.exp(-np.float64(wx)))) - F(y)))
            stash.append(st)
            resid.append(rs)
        log.append(resid)
        # Push, sources in rank order; inside a source the tokens' atomic operations interleave at random
        for s, (rows_tokens, labels) in enumerate(sources):
            seq += 1
            t.rows_by_seq[seq] = float(len(rows_tokens))
            work = [(k, snap, int(np.rint(np.float64(resid[s][r]) * FIX)))
                    for r, toks in enumerate(rows_tokens) for k, snap in zip(toks, stash[s][r])]
            pend = [work[i] for i in rng.permutation(len(work))]
            while pend:
                i = int(rng.integers(0, len(pend)))
                k, snap, fix = pend.pop(i)
                cur = t.rows[k]
                if tuple(cur) == snap:                       # CAS from the (possibly refreshed) look succeeds
                    _, n, z = t.fold(snap, seq, t.rows_by_seq)
                    t.rows[k] = [n, z, seq, fix]
                elif cur[2] == seq:                          # open for this (round, source): integer add
                    cur[3] += fix
                else:                                        # an earlier source got there: retry from what came back
                    assert cur[2] > snap[2] or snap[2] == 0
                    pend.append((k, tuple(cur), fix))
    return t.flushed(), log


def run_sharded_eager(rounds, log):
    state = {}
    for sources, resid in zip(rounds, log):
        for s, (rows_tokens, labels) in enumerate(sources):   # one step per (source, key), rank order
            gsum = {}
            for r, toks in enumerate(rows_tokens):
                fix = int(np.rint(np.float64(resid[s][r]) * FIX))
                for k in toks:
                    gsum[k] = gsum.get(k, 0) + fix
            for k, gfix in gsum.items():
                w, n, z = state.get(k, (F(0), F(0), F(0)))
                state[k] = ftrl_coord(grad_of(gfix, float(len(rows_tokens))), w, n, z)
    return state


@pytest.mark.parametrize("seed,S", [(1, 2), (2, 4), (3, 8)])
def test_sharded_push_from_stale_looks_equals_rank_ordered_steps(seed, S):
    rng = np.random.default_rng(seed)
    rounds = []
    for _ in range(4):
        sources = []
        for _s in range(S):
            B = int(rng.integers(2, 10))
            sources.append(([list(rng.zipf(1.5, int(rng.integers(0, 7))) % 23) for _ in range(B)], rng.integers(0, 2, B)))
        rounds.append(sources)
    a, log = run_sharded(rounds, np.random.default_rng(10 + seed))
    b, _ = run_sharded(rounds, np.random.default_rng(20 + seed))
    eager = run_sharded_eager(rounds, log)
    touched = {k for sources in rounds for toks_l, _ in sources for toks in toks_l for k in toks}
    assert set(a) == touched == set(eager)
    for k in a:
        assert tuple(x.tobytes() for x in a[k]) == tuple(x.tobytes() for x in eager[k]), k
        assert tuple(x.tobytes() for x in a[k]) == tuple(x.tobytes() for x in b[k]), k
