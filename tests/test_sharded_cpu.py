"""CPU (gloo, world_size 2) test of the multi-GPU path's HOST LOGIC: the sharding rule and the
exchange protocol of xflow_b200/csrc/comm.cu (dedup -> bucket by owner -> all-to-all keys -> owner
pull -> all-to-all values -> worker forward/gradient -> all-to-all gradients -> owners apply the
pushes in rank order), mirrored in Python over torch.distributed with the oracle as the arithmetic.
It must reproduce, bit for bit, the single-table lock-step schedule of the oracle (every worker pulls
before any push of the round; pushes applied in rank order) — the schedule DESIGN.md defines for N>1."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as O
from xflow_b200 import api, datagen

WORLD = 2
ROUNDS = 3
B, D, SPACE = 256, 12, 3000


def _batch(rank, rnd):
    return datagen.make_csr_keys(50 + 10 * rnd + rank, B, D, SPACE, api.hash_decimal_ids, dist="zipf", zipf_s=1.2)


def _all_to_all_var(arrs, dtype, width=1):
    """arrs[q] goes to rank q; returns list of arrays received from each rank."""
    world = dist.get_world_size()
    counts = torch.tensor([a.shape[0] for a in arrs], dtype=torch.int64)
    rcounts = torch.zeros(world, dtype=torch.int64)
    dist.all_to_all_single(rcounts, counts)
    # gloo has no all_to_all for tensors of different sizes on every version: use pairwise send/recv
    out = []
    reqs = []
    for q in range(world):
        shape = (int(rcounts[q]),) if width == 1 else (int(rcounts[q]), width)
        out.append(torch.zeros(shape, dtype=dtype))
    me = dist.get_rank()
    for q in range(world):
        if q == me:
            out[q].copy_(torch.from_numpy(np.ascontiguousarray(arrs[q])).view(dtype).reshape(out[q].shape))
            continue
        reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(arrs[q])).view(dtype).reshape(
            (arrs[q].shape[0],) if width == 1 else (arrs[q].shape[0], width)), q))
        reqs.append(dist.irecv(out[q], q))
    for r in reqs:
        r.wait()
    return [o.numpy() for o in out]


def _worker(rank, port, K, opt, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    shard = O.Table(K=K, opt=opt, init_mode=O.INIT_COUNTER, seed=9)   # this rank's key range
    if api.shard_of(0, WORLD) == rank:
        for _ in range(WORLD):                                        # every worker's init push of key 0
            shard.init_push()
    losses = []
    for rnd in range(ROUNDS):
        rp, keys, lab = _batch(rank, rnd)
        uk = np.unique(keys)                                          # sorted unique keys of the slice
        owner = np.array([api.shard_of(int(k), WORLD) for k in uk], np.int64)
        buckets = [uk[owner == q] for q in range(WORLD)]
        # keys are sorted, so buckets are contiguous ranges exactly like DefaultSlicer's lower_bound cut
        assert np.array_equal(np.concatenate(buckets), uk)
        # all-to-all #1: Pull requests
        req = _all_to_all_var(buckets, torch.int64)
        req = [r.view(np.uint64) for r in req]
        # owner: pull (insert on pull)
        resp_w, resp_v = [], []
        for r in req:
            w, v = shard.pull(r)
            resp_w.append(w)
            resp_v.append(v if K else np.zeros((r.size, 0), np.float32))
        # all-to-all #2: values back
        got_w = _all_to_all_var(resp_w, torch.float32)
        w_u = np.concatenate(got_w)
        v_u = None
        if K:
            got_v = _all_to_all_var(resp_v, torch.float32, width=K)
            v_u = np.concatenate(got_v)
        gw, gv, loss = O.worker_compute_given(K, rp.astype(np.int64), keys, lab.astype(np.int32), w_u, v_u)
        losses.append(loss)
        # all-to-all #3: gradients to the owners
        offs = np.cumsum([0] + [b.size for b in buckets])
        g_w = _all_to_all_var([gw[offs[q]:offs[q + 1]] for q in range(WORLD)], torch.float32)
        g_v = _all_to_all_var([gv[offs[q]:offs[q + 1]] for q in range(WORLD)], torch.float32, width=K) if K else None
        # owner: pushes applied in source-rank order
        for q in range(WORLD):
            if req[q].size:
                shard.push(req[q], g_w[q], g_v[q] if K else None)
        dist.barrier()
    # collect: every rank reports its shard contents for all keys ever seen
    allk = np.unique(np.concatenate([_batch(r, rnd)[1] for r in range(WORLD) for rnd in range(ROUNDS)] +
                                    [np.zeros(1, np.uint64)]))
    mine = np.array([api.shard_of(int(k), WORLD) == rank for k in allk])
    e = shard.export(allk[mine])
    ret[rank] = dict(keys=allk[mine], e=e, losses=losses)
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("K,opt", [(0, O.OPT_FTRL), (4, O.OPT_FTRL), (3, O.OPT_SGD)])
def test_sharded_protocol_equals_lockstep_oracle(K, opt):
    O.lib()
    api.lib()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(_free_port(), K, opt, ret), nprocs=WORLD, join=True)

    # single-table lock-step schedule
    t = O.Table(K=K, opt=opt, init_mode=O.INIT_COUNTER, seed=9)
    for _ in range(WORLD):
        t.init_push()
    ref_losses = {r: [] for r in range(WORLD)}
    for rnd in range(ROUNDS):
        pend = []
        for r in range(WORLD):                       # every worker pulls + computes first ...
            rp, keys, lab = _batch(r, rnd)
            uk, gw, gv, loss = t.worker_compute(rp.astype(np.int64), keys, lab.astype(np.int32))
            pend.append((uk, gw, gv))
            ref_losses[r].append(loss)
        for uk, gw, gv in pend:                      # ... then the pushes land in rank order
            t.push(uk, gw, gv if K else None)
    for r in range(WORLD):
        got = ret[r]
        for a, b in zip(got["losses"], ref_losses[r]):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
        ref = t.export(got["keys"])
        assert np.array_equal(got["e"]["present"], ref["present"])
        for k in ("w", "nw", "zw") + (("v", "nv", "zv") if K else ()):
            assert np.array_equal(got["e"][k].view(np.uint32), ref[k].view(np.uint32)), (r, k)
    # every key lives on exactly the shard the bucketing rule names
    assert sum(ret[r]["keys"].size for r in range(WORLD)) == np.unique(np.concatenate(
        [_batch(r, rnd)[1] for r in range(WORLD) for rnd in range(ROUNDS)] + [np.zeros(1, np.uint64)])).size
