#!/usr/bin/env python
"""bench.py — training examples/sec of the xflow hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload lr_ftrl|fm_ftrl] [--impl reference]

A "step" is one pass of the hot path over one batch: LRWorker::update / FMWorker::update (pull,
forward, gradient, push) plus the server-side FTRL step it triggers, on one CSR batch per GPU.
N=1 workload = BASELINE.json configs[1]: LR + FTRL, synthetic ids uniform in a 1e7-feature space,
64 nnz/row, batch 65536 (keys = std::hash of the decimal id string, as the reference's loader makes
them).  For N>1 (torchrun, one rank per GPU) every rank trains its own batch against the table
sharded by key range over the N GPUs (weak scaling; id space 1e7 per GPU).

  value     whole-job examples/s, batches already resident in HBM (device-timed, max over ranks)
  e2e       same metric through the C ABI with page-locked HOST batches: H2D copies and the
            per-step result read-back inside the timed region
  roofline  dominant kernel: SURVEY §8d algorithmic bytes / its CUDA-event time, vs the measured HBM peak
  cpu_baseline  the reference's own CPU implementation (oracle/_ref, compiled from the reference's
            sources) on the box's host cores, on a bounded sample of the same workload

--impl reference times only that CPU implementation and prints the same line with "impl": "reference".
"""
import argparse
import json
import os
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]
    "lr_ftrl": dict(model="lr", opt="ftrl", K=0, id_space=10 ** 7, nnz=64, batch=65536, dist="uniform",
                    name="LR+FTRL, synthetic libffm ids uniform in 1e7-feature space, 64 nnz/row, batch 65536"),
    # BASELINE.json configs[4] shape on the GPUs available
    "fm_ftrl": dict(model="fm", opt="ftrl", K=16, id_space=10 ** 8, nnz=64, batch=65536, dist="zipf",
                    name="FM k=16+FTRL, synthetic libffm ids Zipf(1.05) in 1e8-feature space, 64 nnz/row, batch 65536"),
}
RING = 8  # distinct batches cycled through (8 x 33.6 MB of keys > 126 MB L2; the table is ~1 GB)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(wl, B, nnz, U):
    """keys + row_ptr + labels + pull + optimizer state read + write, per batch."""
    D = 1 + wl["K"]
    R = 3 if wl["opt"] == "ftrl" else 1
    step = nnz * 8 + (B + 1) * 4 + B * 4 + U * D * 4          # fused step kernel: CSR in, w/v rows pulled
    update = U * D * 4 * R * 2                                 # optimizer kernel: state read + written
    return step, update


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the benchmark runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.max_mhz = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                self.samples.append((time.monotonic(), mhz, reasons))
            except Exception:
                pass
            time.sleep(0.005)

    def summary(self, t0, t1):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        nv = self.nv
        win = [s for s in self.samples if t0 <= s[0] <= t1] or self.samples[-3:]
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        seen = set()
        for _, _, r in win:
            for bit, name in names.items():
                if r & bit:
                    seen.add(name)
        mhz = sorted(s[1] for s in win)
        return {"sm_mhz": float(mhz[len(mhz) // 2]) if mhz else None, "sm_max_mhz": float(self.max_mhz),
                "reasons": sorted(seen), "samples": len(win)}


def make_ring(wl, rank, n_batches):
    from xflow_b200 import api, datagen
    ring = []
    for i in range(n_batches):
        rp, ids, lab = datagen.make_ids(seed=1 + 1000 * rank + i, rows=wl["batch"], nnz_per_row=wl["nnz"],
                                        id_space=wl["id_space"], dist=wl["dist"], zipf_s=1.05)
        ring.append((rp, api.hash_decimal_ids(ids), lab, ids.astype(np.uint32)))
    return ring


# --------------------------------------------------------------------------------------------------
# reference arm / cpu baseline
# --------------------------------------------------------------------------------------------------
def cpu_reference_rate(wl, rows, steps, warmup):
    """Times the reference's CPU implementation of the same step on `rows` rows per step.
    Returns (examples/s, mean seconds per step, descriptor dict)."""
    from oracle import oracle as O
    from xflow_b200 import datagen
    tmp = tempfile.mkdtemp(prefix="xfbench_")
    rp, ids, lab = datagen.make_ids(seed=4242, rows=rows, nnz_per_row=wl["nnz"], id_space=wl["id_space"],
                                    dist=wl["dist"], zipf_s=1.05)
    train = os.path.join(tmp, "train")
    datagen.write_text(train + "-00000", rp, ids, lab)
    open(os.path.join(tmp, "empty-00000"), "w").close()
    size_mb = os.path.getsize(train + "-00000") // (1 << 20) + 2
    cores = os.cpu_count() or 1
    times = []
    if O.have_ref():
        kind = "reference"
        for i in range(warmup + steps):
            r = O.run_ref(wl["model"], wl["opt"], train, os.path.join(tmp, "empty"), 1, tmp, core=cores,
                          block_mb=size_mb, vdim=wl["K"] or 10, no_predict=True)
            if i >= warmup:
                times.append(r["train_seconds"])
        used = cores
        how = "oracle/_ref/xflow_ref = the reference's src/ compiled unmodified (-O2) + in-process ps shim " \
              "(zero transport cost), core_num=%d" % cores
        # SURVEY 8d also asks for the single-thread figure: same binary, core_num=1, a quarter of the rows
        single = None
        try:
            rows1 = max(1024, rows // 4)
            train1 = os.path.join(tmp, "train1")
            datagen.write_text(train1 + "-00000", rp[:rows1 + 1], ids[:int(rp[rows1])], lab[:rows1])
            r1 = O.run_ref(wl["model"], wl["opt"], train1, os.path.join(tmp, "empty"), 1, tmp, core=1,
                           block_mb=size_mb, vdim=wl["K"] or 10, no_predict=True)
            single = {"value": rows1 / r1["train_seconds"], "unit": "examples/s", "cores": 1, "rows": rows1}
        except Exception as ex:  # informational only
            single = {"value": None, "error": repr(ex)}
    else:
        kind = "port"
        O.build()
        for i in range(warmup + steps):
            t = O.Table(K=wl["K"], opt=O.OPT_FTRL if wl["opt"] == "ftrl" else O.OPT_SGD)
            t0 = time.perf_counter()
            O.train_file(t, train + "-00000", size_mb << 20, 1)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
        used = 1
        how = "oracle/xflow_oracle.cc (scalar port), 1 thread"
    mean_t = float(np.mean(times))
    desc = {"kind": kind, "cores": used, "single_core": single if kind == "reference" else None,
            "sample": "%d rows x %d nnz of the same workload per step (text parse + update(), 1 epoch, empty "
                      "table); %s" % (rows, wl["nnz"], how)}
    return rows / mean_t, mean_t, desc


def run_reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    budget_s = 150.0
    rows = int(min(wl["batch"], max(1024, budget_s * 4000.0 / max(1, args.steps + args.warmup))))
    rate, mean_t, desc = cpu_reference_rate(wl, rows, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "training examples/sec", "value": rate, "unit": "examples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": mean_t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "rows_per_step": rows},
        "cpu_baseline": dict(desc, value=rate, unit="examples/s"),
        "e2e": {"value": rate, "unit": "examples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def text_leg(api, tr, wl, B, passes=12, warm=2):
    """File -> model throughput: a text shard of one batch (page-cache warm), every pass = block read +
    H2D of the raw text + device parse/hash + one training step.  Host wall clock around synced passes."""
    import ctypes as C
    from xflow_b200 import datagen
    tmp = tempfile.mkdtemp(prefix="xftext_")
    path = os.path.join(tmp, "shard-00000")
    rp, ids, lab = datagen.make_ids(seed=777, rows=B, nnz_per_row=wl["nnz"], id_space=wl["id_space"],
                                    dist=wl["dist"], zipf_s=1.05)
    datagen.write_text(path, rp, ids, lab)
    size = os.path.getsize(path)
    lib = api.lib()
    text, ln, r, z = C.c_void_p(), C.c_uint64(), C.c_uint32(), C.c_uint32()
    times = []
    for p in range(warm + passes):
        ld = api.Loader(path, size + (1 << 20))
        tr.sync()
        t0 = time.perf_counter()
        rows = 0
        while True:
            assert lib.xf_loader_next_raw(ld.h, C.byref(text), C.byref(ln)) == 0
            if not ln.value:
                break
            assert lib.xf_trainer_ingest_text(tr.h, text, ln.value, C.byref(r), C.byref(z)) == 0, lib.xf_last_error()
            assert lib.xf_trainer_step_ingested(tr.h, 0, r.value) == 0, lib.xf_last_error()
            rows += r.value
        tr.sync()
        if p >= warm:
            times.append(time.perf_counter() - t0)
        assert rows == B
        ld.close()
    os.remove(path)
    t = float(np.mean(times))
    return {"value": B / t, "unit": "examples/s", "ms_per_step": t * 1e3, "text_bytes_per_step": size,
            "text_gbs": size / t / 1e9, "passes": passes,
            "api": "xf_loader_next_raw + xf_trainer_ingest_text + xf_trainer_step_ingested (C ABI): block read "
                   "from the page cache, H2D of raw text, parse + hash + step on the device"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="lr_ftrl", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-text-e2e", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        run_reference_arm(args, wl)
        return

    import torch
    from xflow_b200 import api

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if api.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    comm = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        cid = torch.from_numpy(api.Comm.new_id() if rank == 0 else np.zeros(api.COMM_ID_BYTES, np.uint8)).cuda()
        dist.broadcast(cid, 0)
        comm = api.Comm(cid.cpu().numpy(), rank, world, local)

    B, nnz = wl["batch"], wl["batch"] * wl["nnz"]
    model = api.MODEL_LR if wl["model"] == "lr" else api.MODEL_FM
    keys_per_shard = wl["id_space"]  # weak scaling: the id space grows with N
    wl = dict(wl, id_space=wl["id_space"] * world)
    stride_guess = 32 if wl["K"] == 0 else 32 + 16 * wl["K"]
    cap = 1
    while cap < 2.5 * keys_per_shard:
        cap <<= 1
    table = api.Table(latent_dim=wl["K"], optimizer=api.OPT_FTRL if wl["opt"] == "ftrl" else api.OPT_SGD,
                      device=local, capacity=cap, seed=1, shard_index=rank, num_shards=world)
    stream = torch.cuda.Stream()
    table.set_stream(stream.cuda_stream)
    tr = api.Trainer(table, model=model, max_rows=B, max_nnz=nnz, comm=comm)

    ring = make_ring(wl, rank, RING)
    # device-resident copies (raw bytes; torch is only the allocator here)
    dev = []
    for rp, keys, lab, ids in ring:
        dev.append(tuple(torch.from_numpy(a.view(np.uint8)).cuda() for a in (rp, keys, lab)))
    # page-locked host copies for the end-to-end legs: CSR of u32 feature ids (hashed to keys on the
    # device, the loader's hashing moved to the GPU) and, for comparison, CSR of ready-made u64 keys
    pin = []
    for rp, keys, lab, ids in ring:
        pin.append(tuple(torch.from_numpy(a.view(np.uint8)).pin_memory() for a in (rp, keys, lab, ids)))
    results = torch.zeros(max(args.steps, 1), dtype=torch.float32).pin_memory()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def run_device(k0, k):
        for i in range(k0, k0 + k):
            d = dev[i % RING]
            tr.step_device(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), B, nnz)

    sampler = ClockSampler(local)
    sampler.start()
    with torch.cuda.stream(stream):
        tr.init_push()
        run_device(0, RING)              # population pass: every key of the ring gets inserted
        run_device(0, args.warmup)       # warm-up steps (untimed)
        tr.sync()
        # ---------------- device-resident timed region
        st0 = tr.stats()
        l0 = tr.launches()
        tr.set_profile(True)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_w0 = time.monotonic()
        e0.record(stream)
        run_device(args.warmup, args.steps)
        e1.record(stream)
        barrier()
        t_w1 = time.monotonic()
        ms = e0.elapsed_time(e1)
        prof = tr.profile()
        tr.set_profile(False)
        st1 = tr.stats()
        launches = tr.launches() - l0
        # ---------------- end-to-end timed regions (host batches, H2D + result D2H every step)
        def e2e_leg(use_ids):
            def one(i, out):
                p = pin[i % RING]
                if use_ids:
                    tr.step_host_ids_async(p[0].data_ptr(), p[3].data_ptr(), p[2].data_ptr(), B, nnz, out)
                else:
                    tr.step_host_async(p[0].data_ptr(), p[1].data_ptr(), p[2].data_ptr(), B, nnz, out)
            for i in range(args.warmup):
                one(i, results.data_ptr())
            tr.sync()
            barrier()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record(stream)
            for i in range(args.steps):
                one(i, results.data_ptr() + 4 * i)
            f1.record(stream)
            tr.sync()
            barrier()
            return f0.elapsed_time(f1)
        ms_e2e_keys = e2e_leg(False)
        ms_e2e = e2e_leg(True)
        # ---------------- file -> model (reported separately): one batch as a text shard in the
        # reference's format, read by xf_loader_next_raw, parsed + hashed + trained on the device
        text_e2e = None
        if world == 1 and not args.no_text_e2e:
            text_e2e = text_leg(api, tr, wl, B)
    sampler.stop_flag = True
    sampler.join(timeout=1.0)

    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms, ms_e2e, ms_e2e_keys], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e, ms_e2e_keys = float(t[0]), float(t[1]), float(t[2])
    assert np.isfinite(results[: args.steps].numpy()).all()

    if rank == 0:
        steps = args.steps
        value = world * B * steps / (ms * 1e-3)
        e2e_value = world * B * steps / (ms_e2e * 1e-3)
        U = (st1["unique_keys"] - st0["unique_keys"]) / max(steps, 1)
        peak, peak_src = load_peaks()
        b_step, b_update = algorithmic_bytes(wl, B, nnz, U)
        t_step = prof["step_ms"] / max(prof["steps"], 1) * 1e-3
        t_upd = prof["update_ms"] / max(prof["steps"], 1) * 1e-3
        if t_upd > 0.05 * t_step:
            kern = [("xf_k_step (fused pull+forward+gradient)", b_step, t_step),
                    ("xf_k_update (FTRL over touched rows)", b_update, t_upd)]
        else:
            # LR tables fold the optimizer step into the next touch of a row: one kernel does all of it
            kern = [("xf_k_step_lr_lazy (pull+forward+gradient+optimizer in one kernel)", b_step + b_update, t_step)]
        dom = max(kern, key=lambda k: k[2])
        traffic = sectors = ceiling = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                tj = json.load(open(tpath))
                traffic = tj.get(args.workload, {}).get(dom[0].split(" ")[0])
                sectors = tj.get(args.workload, {}).get(dom[0].split(" ")[0] + "_dram_sectors")
                ceiling = tj.get("random_sector_ceiling_gsectors_per_s")
            except Exception:
                traffic = None
        roofline = {
            "bound": "hbm", "kernel": dom[0], "achieved": dom[1] / dom[2] / 1e9, "peak": peak, "unit": "GB/s",
            "frac": dom[1] / dom[2] / 1e9 / peak, "traffic": traffic, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": dom[1], "avg_launch_ms": dom[2] * 1e3,
            "kernels": [{"kernel": n, "algorithmic_bytes": b, "avg_ms": t * 1e3, "gbs": b / t / 1e9 if t else None}
                        for n, b, t in kern],
            "step_algorithmic_bytes": b_step + b_update, "step_gbs": (b_step + b_update) / (ms * 1e-3 / steps) / 1e9,
            "step_frac": (b_step + b_update) / (ms * 1e-3 / steps) / 1e9 / peak,
            "unique_keys_per_batch": U,
            # the access pattern is one random 32-byte sector per table row: the practical ceiling is the
            # measured random-sector request rate (tools/randsector.cu, profiles/r01_randsector.md), not
            # the streaming-copy bandwidth `peak` above
            "random_sector_ceiling_gsectors_per_s": ceiling,
            "dram_sector_requests_per_launch": sectors,
            "frac_of_random_sector_ceiling": (sectors / dom[2] / 1e9 / ceiling) if (sectors and ceiling) else None,
        }
        line = {
            "metric": "training examples/sec", "value": value, "unit": "examples/s", "n_gpus": world,
            "steps": steps, "warmup": args.warmup, "ms_per_step": ms / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"], "model": wl["model"], "optimizer": wl["opt"], "latent_dim": wl["K"],
                       "batch_per_gpu": B, "nnz_per_row": wl["nnz"], "id_space": wl["id_space"],
                       "table_slots_per_gpu": table.capacity(), "table_row_bytes": table.row_bytes(),
                       "ring_batches": RING,
                       "l2": "inputs larger than L2: %d distinct batches (%.0f MB of keys) cycled, table %.1f GB"
                             % (RING, RING * nnz * 8 / 1e6, table.capacity() * table.row_bytes() / 1e9),
                       "parallelism": "dp%d, table range-sharded over %d GPU(s)" % (world, world)},
            "clocks": sampler.summary(t_w0, t_w1),
            "e2e": {"value": e2e_value, "unit": "examples/s", "ms_per_step": ms_e2e / steps,
                    "h2d_bytes_per_step": (B + 1) * 4 + nnz * 4 + B, "d2h_bytes_per_step": 4,
                    "api": "xf_trainer_step_host_ids_async (C ABI): page-locked host CSR of u32 feature ids, "
                           "hashed to keys on the device",
                    "with_prehashed_u64_keys": {"value": world * B * steps / (ms_e2e_keys * 1e-3),
                                                "h2d_bytes_per_step": (B + 1) * 4 + nnz * 8 + B,
                                                "api": "xf_trainer_step_host_async"}},
            "gpu_launches": int(launches),
            "roofline": roofline,
        }
        if text_e2e:
            line["e2e_text"] = text_e2e
        if world == 1 and not args.no_cpu_baseline:
            try:
                rows = 16384
                rate, mean_t, desc = cpu_reference_rate(wl, rows, 1, 0)
                line["cpu_baseline"] = dict(desc, value=rate, unit="examples/s")
            except Exception as ex:  # the baseline is reported, never required for the GPU numbers
                line["cpu_baseline"] = {"value": None, "unit": "examples/s", "cores": 0, "kind": "port",
                                        "sample": "failed: %r" % (ex,)}
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
