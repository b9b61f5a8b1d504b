"""Multi-GPU parity (pytest -m gpu, needs >= 2 GPUs on the box; skipped otherwise): the sharded step of
xflow_b200/csrc/comm.cu (all-to-all pull / push over peer memory or NCCL against a range-sharded device table), driven
through the C ABI from one process per GPU, against the oracle's single-table lock-step schedule."""
import os
import socket

import numpy as np
import pytest

from common import assert_close
from oracle import oracle as O
from xflow_b200 import api, datagen

pytestmark = pytest.mark.gpu

ROUNDS = 3
B, D, SPACE = 4096, 32, 200000


def _batch(rank, rnd):
    return datagen.make_csr_keys(500 + 10 * rnd + rank, B, D, SPACE, api.hash_decimal_ids)


def _all_keys(world):
    return np.unique(np.concatenate([_batch(r, rnd)[1] for r in range(world) for rnd in range(ROUNDS)] +
                                    [np.zeros(1, np.uint64)]))


def _worker(rank, world, id_path, model, opt, K, ret):
    from xflow_b200 import api as A
    if rank == 0:
        cid = A.Comm.new_id()
        np.save(id_path + ".tmp.npy", cid)
        os.replace(id_path + ".tmp.npy", id_path)
    else:
        import time
        while not os.path.exists(id_path):
            time.sleep(0.05)
        cid = np.load(id_path)
    comm = A.Comm(cid, rank, world, rank)
    table = A.Table(latent_dim=K, optimizer=opt, device=rank, v_init=A.VINIT_COUNTER, seed=9, shard_index=rank,
                    num_shards=world, capacity=1 << 16)
    tr = A.Trainer(table, model=model, max_rows=B, max_nnz=B * D, keep_loss=True, comm=comm)
    tr.init_push()
    comm.barrier()
    losses = []
    for rnd in range(ROUNDS):
        rp, keys, lab = _batch(rank, rnd)
        tr.step_host(rp, keys, lab)
        losses.append(tr.get_loss(B))
    # forward only through the sharded path (same collective schedule on every rank), on a batch whose keys
    # all exist already, so nothing changes in the table
    rp, keys, lab = _batch(rank, ROUNDS - 1)
    pctr = tr.predict_host(rp, keys)
    comm.barrier()
    allk = _all_keys(world)
    mine = np.array([A.shard_of(int(k), world) == rank for k in allk])
    ret[rank] = dict(keys=allk[mine], e=table.export(allk[mine]), losses=losses, size=table.size(),
                     uniq=tr.stats()["unique_keys"], pctr=pctr)
    # foreign keys must be absent from this shard
    other = table.export(allk[~mine][:1000])
    assert not other["present"].any()
    tr.close()
    table.close()
    comm.close()


@pytest.mark.parametrize("exchange", ["peer", "nccl"])
@pytest.mark.parametrize("model,opt,K", [("lr", "ftrl", 0), ("fm", "ftrl", 8), ("fm", "sgd", 4)])
def test_two_gpu_sharded_step_matches_lockstep_oracle(model, opt, K, exchange, tmp_path, monkeypatch):
    """exchange = "peer": cudaIpc-mapped buffers read over NVLink (default); "nccl": grouped send/recv."""
    if api.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    if exchange == "nccl" and (model, opt) != ("fm", "ftrl"):
        pytest.skip("the fallback exchange is checked on one configuration")
    monkeypatch.setenv("XFLOW_P2P", "1" if exchange == "peer" else "0")  # inherited by the spawned ranks
    import torch.multiprocessing as mp
    world = 2
    gopt, oopt = (api.OPT_FTRL, O.OPT_FTRL) if opt == "ftrl" else (api.OPT_SGD, O.OPT_SGD)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, str(tmp_path / "ncclid.npy"), api.MODEL_LR if model == "lr" else api.MODEL_FM,
                            gopt, K, ret), nprocs=world, join=True)
    t = O.Table(K=K, opt=oopt, init_mode=O.INIT_COUNTER, seed=9)
    for _ in range(world):
        t.init_push()
    ref_losses = {r: [] for r in range(world)}
    uniq = {r: 0 for r in range(world)}
    for rnd in range(ROUNDS):
        pend = []
        for r in range(world):
            rp, keys, lab = _batch(r, rnd)
            uk, gw, gv, loss = t.worker_compute(rp.astype(np.int64), keys, lab.astype(np.int32))
            pend.append((uk, gw, gv))
            ref_losses[r].append(loss)
            uniq[r] += uk.size
        for uk, gw, gv in pend:
            t.push(uk, gw, gv if K else None)
    total = 0
    for r in range(world):
        got = ret[r]
        assert got["uniq"] == uniq[r]                                   # dedup is exact
        for a, b in zip(got["losses"], ref_losses[r]):
            assert_close(a, b, "loss rank %d" % r, abs_floor=1e-6)
        ref = t.export(got["keys"])
        assert np.array_equal(got["e"]["present"], ref["present"])      # bucketing bit-exact
        for k in ("w", "nw", "zw") + (("v", "nv", "zv") if K else ()):
            assert_close(got["e"][k], ref[k], "rank %d %s" % (r, k))
        total += got["size"]
    assert total == t.size()
    # predictions after training: the oracle's forward pass on the final table
    for r in range(world):
        rp, keys, lab = _batch(r, ROUNDS - 1)
        _, _, _, loss = t.worker_compute(rp.astype(np.int64), keys, lab.astype(np.int32))
        assert_close(ret[r]["pctr"], loss.astype(np.float64) + lab, "pctr rank %d" % r, abs_floor=1e-6)
