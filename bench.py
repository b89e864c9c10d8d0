#!/usr/bin/env python
"""bench.py -- SpGEMM GFLOPS (2*P/t) of C = A.B on B200, one JSON line on stdout (rank 0).

A "step" is one complete multiply (analysis -> symbolic -> scan -> numeric) of the workload.
  N = 1 : BASELINE.json configs[1]: R-MAT scale 20, edge factor 16, (0.45,0.15,0.15,0.25), seed 20, fp64,
          C = A.A (P = 586 218 280, nnz(C) = 582 205 696).  The same line carries a `sweep` over the other
          BASELINE configs (SuiteSparse-shaped generators and the config-5 matrix on one GPU): ours vs the
          reference spECK CUDA build on the same GPU, with index parity flags.
  N > 1 : BASELINE.json configs[4] (strong scaling): ONE matrix, R-MAT scale 24, edge factor 4, seed 24.  Rank 0
          analyses A on its device and cuts the rows into N contiguous slabs balanced by intermediate products
          (speck_b200_partition_rows), B (= A) is NCCL-broadcast once at setup (untimed), every rank multiplies its
          slab -- no collective in the timed region -- value = 2*P*K / max_r(t_r).  Extra keys: the same matrix on
          one GPU (rank 0 alone) for the speed-up, the product imbalance, the time to concatenate C on GPU 0
          (NCCL send/recv + offset fix-up), and the round-1 weak-scaling figure (one R-MAT-20 block per rank).
Keys follow the driver contract; `roofline` is the numeric phase (all numeric kernels, bracketed by CUDA events on
the library's streams); `cpu_baseline` is the CPU oracle (a port: the reference has no CPU SpGEMM) on all host
cores.  `--impl reference` times that CPU oracle on the arm's own workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from speck_b200 import matrices as M  # noqa: E402
from speck_b200.matrices import HostCSR  # noqa: E402

CACHE = os.environ.get("SPECK_B200_CACHE", "/tmp/speck_b200_cache")

# seeds of SURVEY.md 8d
WORKLOAD_SEEDS = {"rmat20": 20, "rmat18": 18, "rmat16": 16, "rmat24": 24, "webbase_like": 3, "cant_like": 44,
                  "banded_like": 41, "econ_like": 42, "circuit_like": 43}
WORKLOAD_DESC = {
    "rmat20": "R-MAT scale 20 ef 16 (0.45,0.15,0.15,0.25) seed 20, fp64, C=A.A (BASELINE configs[1])",
    "rmat24": "R-MAT scale 24 ef 4 (0.45,0.15,0.15,0.25) seed 24, fp64, C=A.A (BASELINE configs[4]: 16.8 M rows)",
    "webbase_like": "webbase-1M-shaped generator (BASELINE configs[2]; the real file is not available offline)",
    "cant_like": "cant-shaped 3-D FEM generator (BASELINE configs[3])",
    "banded_like": "banded FEM generator, 64 nnz/row (BASELINE configs[3])",
    "econ_like": "mac_econ_fwd500-shaped generator (BASELINE configs[3])",
    "circuit_like": "scircuit-shaped generator (BASELINE configs[3])",
}
SWEEP = ["webbase_like", "econ_like", "circuit_like", "cant_like", "banded_like", "rmat24"]


def load_workload(name, seed=None):
    """Deterministic workload, cached as .npz so that tests, both bench arms and all ranks share the generation."""
    if seed is None:
        seed = WORKLOAD_SEEDS.get(name, 20)
    os.makedirs(CACHE, exist_ok=True)
    path = os.path.join(CACHE, f"{name}_s{seed}.npz")
    if os.path.exists(path):
        try:
            z = np.load(path)
            return HostCSR(int(z["rows"]), int(z["cols"]), z["rp"], z["ci"], z["v"])
        except Exception:
            pass
    gens = {"rmat20": lambda: M.rmat(20, 16, seed=seed), "rmat18": lambda: M.rmat(18, 16, seed=seed),
            "rmat16": lambda: M.rmat(16, 16, seed=seed), "rmat24": lambda: M.rmat(24, 4, seed=seed),
            "webbase_like": lambda: M.webbase_like(seed=seed), "cant_like": lambda: M.fem3d_like(seed=seed),
            "banded_like": lambda: M.banded_fem_like(seed=seed), "econ_like": lambda: M.econ_like(seed=seed),
            "circuit_like": lambda: M.circuit_like(seed=seed)}
    if name not in gens:
        raise SystemExit(f"unknown workload {name}")
    A = gens[name]()
    try:
        tmp = path + f".{os.getpid()}.tmp.npz"
        np.savez(tmp, rows=A.rows, cols=A.cols, rp=A.row_offsets, ci=A.col_ids, v=A.data)
        os.replace(tmp, path)
    except Exception:
        pass
    return A


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])), mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def host_threads():
    """All host cores for the CPU legs: torchrun exports OMP_NUM_THREADS=1 to every rank, which would time the
    CPU baseline on one core."""
    import oracle
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    oracle.set_threads(n)
    return oracle.num_threads()


def time_oracle(As: HostCSR, B: HostCSR, reps=1):
    """mean seconds per multiply of the CPU oracle over `reps` repetitions (after one warm-up)"""
    import oracle
    _, _, P, _ = oracle.row_products(As.row_offsets, As.col_ids, B.row_offsets)
    oracle.spgemm(As.row_offsets, As.col_ids, As.data, B.row_offsets, B.col_ids, B.data, B.cols)
    t0 = time.perf_counter()
    for _ in range(reps):
        oracle.spgemm(As.row_offsets, As.col_ids, As.data, B.row_offsets, B.col_ids, B.data, B.cols)
    return P, (time.perf_counter() - t0) / reps


def run_ref_gpu(workload, check=False, warmup=10, iters=10, timeout=900):
    """The reference's own CUDA build (sm_100, oracle/_ref) on the same GPU and workload, in a subprocess (a crash
    of the reference must not take the bench down).  The reference's protocol: 10 warm-up + 10 timed iterations
    (config.ini:14-17), mean and min reported.  NOT the `--impl reference` arm (that one is the CPU port)."""
    out = {}
    for variant in ("stock", "tuned"):
        so = os.path.join(ROOT, "oracle", "_ref", f"libspeck_ref_{variant}.so")
        if not os.path.exists(so):
            out[variant] = {"unavailable": "oracle/_ref not built (needs /root/reference at build time)"}
            continue
        try:
            cmd = [sys.executable, "-m", "oracle.ref_run", "--workload", workload, "--seed", str(WORKLOAD_SEEDS.get(workload, 20)),
                   "--variant", variant, "--warmup", str(warmup), "--iters", str(iters)] + (["--check"] if check else [])
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
            js = [l for l in r.stdout.splitlines() if l.startswith("{")]
            d = json.loads(js[-1]) if js else {"error": r.stderr[-200:]}
            keys = ("mean_ms", "min_ms", "gflops_mean", "gflops_best", "nnz_c", "row_ptr_equal", "col_idx_equal", "max_rel_err")
            out[variant] = {k: d[k] for k in keys if k in d} or d
        except Exception as e:  # noqa: BLE001
            out[variant] = {"error": repr(e)}
    ok = [v for v in out.values() if "mean_ms" in v]
    out["best_mean_ms"] = min((v["mean_ms"] for v in ok), default=None)
    out["best_min_ms"] = min((v["min_ms"] for v in ok), default=None)
    out["best_gflops_mean"] = max((v.get("gflops_mean", 0.0) for v in ok), default=0.0)
    out["protocol"] = f"{warmup} warm-up + {iters} timed iterations, timings.complete, C reused"
    return out


def run_reference_arm(args, rank, world):
    """CPU oracle (port; the reference has no CPU SpGEMM) on all host cores, rank 0 only, on the workload of our
    arm at this N (config #2 at N = 1, the config-5 matrix at N > 1)."""
    if rank != 0:
        return
    import oracle
    oracle.build()
    cores = host_threads()
    wl = args.workload or ("rmat20" if args.gpus == 1 else "rmat24")
    A = load_workload(wl)
    _, _, P, _ = oracle.row_products(A.row_offsets, A.col_ids, A.row_offsets)
    for _ in range(min(args.warmup, 1)):
        oracle.spgemm(A.row_offsets, A.col_ids, A.data, A.row_offsets, A.col_ids, A.data, A.cols)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.spgemm(A.row_offsets, A.col_ids, A.data, A.row_offsets, A.col_ids, A.data, A.cols)
    dt = time.perf_counter() - t0
    gflops = 2.0 * P * args.steps / dt / 1e9
    sample = f"all rows of A ({A.rows} rows, P={P}) x full B, per step"
    line = {
        "impl": "reference", "metric": "SpGEMM GFLOPS (2*P/t), C=A.A", "value": gflops, "unit": "GFLOPS",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak" if args.gpus == 1 else "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": WORKLOAD_DESC.get(wl, wl), "sample": sample},
        "cpu_baseline": {"value": gflops, "unit": "GFLOPS", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": gflops, "unit": "GFLOPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference (spECK) has no CPU SpGEMM; this arm is the CPU oracle port (oracle/spgemm_oracle.c, OpenMP)",
    }
    print(json.dumps(line), flush=True)


def pinned_csr(torch, m: HostCSR, keep):
    out = []
    for arr in (m.row_offsets, m.col_ids, m.data):
        a = np.ascontiguousarray(arr)
        t = torch.from_numpy(a.view(np.int32) if a.dtype == np.uint32 else a).pin_memory()
        keep.append(t)
        n = t.numpy()
        out.append(n.view(np.uint32) if a.dtype == np.uint32 else n)
    return HostCSR(m.rows, m.cols, out[0], out[1], out[2])


def timed_multiplies(torch, ctx, dA, dB, dC, steps, warmup, stream, barrier):
    """-> (device ms of `steps` multiplies, wall ms, per-stage means, launches, stats of the last one)"""
    for _ in range(warmup):
        ctx.multiply(dA, dB, dC)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    acc = {"analysis": [], "symbolic": [], "scan": [], "numeric": []}
    launches = 0
    t_wall = time.perf_counter()
    ev0.record(stream)
    for _ in range(steps):
        ctx.multiply(dA, dB, dC)
        s = ctx.stats()
        for k in acc:
            acc[k].append(s["ms_" + k])
        launches += s["kernel_launches"]
    ev1.record(stream)
    barrier()
    wall_ms = (time.perf_counter() - t_wall) * 1e3
    return ev0.elapsed_time(ev1), wall_ms, {k: float(np.mean(v)) for k, v in acc.items()}, launches, ctx.stats()


def sweep_one(torch, api, ctx, name, peak, ref=True):
    """ours (10 timed multiplies after 10 warm-ups, the reference's protocol) vs the reference spECK CUDA build on
    one workload; index parity of both against the CPU oracle."""
    import oracle
    A = load_workload(name)
    dA = ctx.upload(A)
    dC = api.DeviceCSR(ctx)
    for _ in range(10):
        ctx.multiply(dA, dA, dC)
    tot, wall, st4 = [], [], []
    for _ in range(10):
        t0 = time.perf_counter()
        ctx.multiply(dA, dA, dC)
        wall.append((time.perf_counter() - t0) * 1e3)
        s = ctx.stats()
        tot.append(s["ms_total"])
        st4.append([s["ms_analysis"], s["ms_symbolic"], s["ms_scan"], s["ms_numeric"]])
    st = ctx.stats()
    C = ctx.download(dC)
    dC.free()
    dA.free()
    rp, ci, v = oracle.spgemm(A.row_offsets, A.col_ids, A.data, A.row_offsets, A.col_ids, A.data, A.cols)
    idx_ok = bool(np.array_equal(rp, C.row_offsets) and np.array_equal(ci, C.col_ids))
    rel = float(np.max(np.abs(v - C.data) / np.maximum(np.abs(v), 1e-300))) if v.size else 0.0
    del C, rp, ci, v
    P, nnzC = st["products"], st["nnz_c"]
    sm = np.mean(np.array(st4), axis=0)
    nb = api.numeric_bytes(A.rows, A.nnz, P, nnzC)
    e = {"workload": name, "rows": A.rows, "nnz_a": A.nnz, "products": P, "nnz_c": nnzC,
         "ours_ms": float(np.mean(wall)), "ours_min_ms": float(np.min(wall)), "ours_device_ms": float(np.mean(tot)),
         "ours_gflops": 2.0 * P / (float(np.mean(wall)) * 1e-3) / 1e9,
         "stage_ms": {"analysis": float(sm[0]), "symbolic": float(sm[1]), "scan": float(sm[2]), "numeric": float(sm[3])},
         "wall_over_stage_sum": float(np.mean(wall)) / max(float(sm.sum()), 1e-9),
         "numeric_frac_of_hbm_peak": nb / (float(sm[3]) * 1e-3) / 1e9 / peak,
         "idx_bit_exact_vs_oracle": idx_ok, "max_rel_err_vs_oracle": rel}
    if ref:
        r = run_ref_gpu(name, check=True)
        e["ref_speck_gpu"] = {k: r[k] for k in ("best_mean_ms", "best_min_ms", "protocol")}
        for variant in ("stock", "tuned"):
            if "mean_ms" in r.get(variant, {}):
                e["ref_speck_gpu"][variant] = {k: r[variant].get(k) for k in ("mean_ms", "min_ms", "row_ptr_equal", "col_idx_equal", "max_rel_err")}
        ok = [r[v] for v in ("stock", "tuned") if "col_idx_equal" in r.get(v, {})]
        e["idx_bit_exact"] = bool(idx_ok and ok and all(x["row_ptr_equal"] and x["col_idx_equal"] for x in ok)) if ok else None
        if r["best_mean_ms"]:
            e["speedup_vs_ref_mean"] = r["best_mean_ms"] / e["ours_ms"]
            e["speedup_vs_ref_min"] = r["best_min_ms"] / e["ours_min_ms"]
    return e


def traffic_entry(workload):
    """DRAM bytes of the numeric phase from the committed ncu launch list (profiles/traffic.json, stamped with the
    commit and the capture it came from; ncu cannot run inside the bench)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return t.get(workload), {k: t.get(k) for k in ("captured_at_commit", "source", "note") if k in t}
    except Exception:
        return None, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, help="default: rmat20 at N = 1, rmat24 at N > 1")
    ap.add_argument("--cpu-reps", type=int, default=10, help="repetitions of the CPU baseline in the GPU arm (about 10-20 s in total)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip timing the reference spECK CUDA build (oracle/_ref) on the same GPU")
    ap.add_argument("--no-sweep", action="store_true", help="N = 1: skip the sweep over the other BASELINE configs")
    ap.add_argument("--no-weak", action="store_true", help="N > 1: skip the extra weak-scaling figure")
    ap.add_argument("--concat", default="push", choices=["push", "nccl", "both"],
                    help="N > 1: how the slabs of C reach GPU 0 for the extra concatenation figure: own push kernel over "
                         "IPC-opened peer memory (falls back to NCCL when IPC is unavailable), NCCL send/recv, or both")
    ap.add_argument("--part-row-cost", type=int, default=8, help="N > 1: cost of a row in products for the partition")
    ap.add_argument("--part-entry-cost", type=int, default=2, help="N > 1: cost of an entry of A in products for the partition")
    ap.add_argument("--sort-max", type=int, default=0)
    ap.add_argument("--opt", action="append", default=[], help="library option key=value (experiments)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    from speck_b200 import api

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the SpGEMM path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if args.concat != "push":
            # NCCL send / recv path of the concatenation figure: with the default of a few channels per peer one
            # send/recv pair moves ~200 GB/s of NVLink's 900 (the default path pushes with its own kernel instead)
            os.environ.setdefault("NCCL_NCHANNELS_PER_PEER", "32")
            os.environ.setdefault("NCCL_MIN_P2P_NCHANNELS", "32")
            os.environ.setdefault("NCCL_MAX_P2P_NCHANNELS", "32")
        dist.init_process_group("nccl", device_id=dev)

    ctx = api.Context(local_rank)
    if args.sort_max:
        ctx.set_option("sort_max", args.sort_max)
    for kv in args.opt:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    stream = torch.cuda.ExternalStream(ctx.lib.speck_b200_stream(ctx.h), device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def bcast_csr(B):
        """B from rank 0 to every rank over NCCL (setup, untimed) -> (DeviceCSR view, tensors kept alive, ms)"""
        meta = torch.tensor([B.rows, B.cols, B.nnz] if rank == 0 else [0, 0, 0], dtype=torch.int64, device=dev)
        dist.broadcast(meta, 0)
        rows, cols, nnz = (int(x) for x in meta.tolist())
        if rank == 0:
            ts = [torch.from_numpy(np.ascontiguousarray(B.row_offsets).view(np.int32)).to(dev),
                  torch.from_numpy(np.ascontiguousarray(B.col_ids).view(np.int32)).to(dev),
                  torch.from_numpy(np.ascontiguousarray(B.data)).to(dev)]
        else:
            ts = [torch.empty(rows + 1, dtype=torch.int32, device=dev), torch.empty(nnz, dtype=torch.int32, device=dev),
                  torch.empty(nnz, dtype=torch.float64, device=dev)]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for t in ts:
            dist.broadcast(t, 0)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        return api.DeviceCSR.from_pointers(ctx, rows, cols, nnz, ts[0].data_ptr(), ts[1].data_ptr(), ts[2].data_ptr()), ts, ms

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    keep = []
    extra = {}
    wl = args.workload or ("rmat20" if world == 1 else "rmat24")

    # ------------------------------------------------------------------ setup (untimed)
    if world == 1:
        A = load_workload(wl)
        dA = ctx.upload(A)
        dB, hostB, slab = dA, A, A
        P_total_expected = None
        bcast_ms = 0.0
        cuts = np.array([0, A.rows])
    else:
        if rank == 0:
            A = load_workload(wl)          # generate / cache once, then every rank reads the cache
        dist.barrier()
        if rank != 0:
            A = load_workload(wl)
        dB, keepB, bcast_ms = bcast_csr(A)   # B = A, rank 0's copy travels over NCCL / NVLink
        keep.append(keepB)
        hostB = A
        # product-balanced cuts from the device analysis on rank 0, then 4*(N+1) bytes to every rank
        t_cuts = torch.zeros(world + 1, dtype=torch.int64, device=dev)
        if rank == 0:
            # a row costs about as much as 7 products, an entry of A about 1.5 (measured on this matrix): balance that,
            # not the products alone (pure product cuts give the ranks with the many tiny rows 30 % more time)
            ctx.set_option("partition_row_cost", args.part_row_cost)
            ctx.set_option("partition_entry_cost", args.part_entry_cost)
            c32, part_products = ctx.partition_rows(dB, dB, world)
            t_cuts = torch.from_numpy(c32.astype(np.int64)).to(dev)
            extra["partition"] = {"cuts": [int(x) for x in c32], "balanced_cost_per_rank": [int(x) for x in part_products],
                                  "cost_model": f"products + {args.part_entry_cost} * nnz(A row) + {args.part_row_cost} per row"}
        dist.broadcast(t_cuts, 0)
        cuts = t_cuts.cpu().numpy()
        slab = A.row_slice(int(cuts[rank]), int(cuts[rank + 1]))
        dA = ctx.upload(slab)

    dC = api.DeviceCSR(ctx)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()          # nvidia-smi needs ~100 ms to start: begin before the warm-up

    # ------------------------------------------------------------------ timed region: K multiplies, device events
    dev_ms, wall_ms, stage, launches, st = timed_multiplies(torch, ctx, dA, dB, dC, args.steps, args.warmup, stream, barrier)
    P, nnzC = st["products"], st["nnz_c"]
    clocks = sampler.stop() if rank == 0 else None

    # ------------------------------------------------------------------ end to end through the host-buffer C ABI
    pA = pinned_csr(torch, slab, keep)
    pB = pA if (world == 1) else pinned_csr(torch, hostB, keep)
    ctx.multiply_host(pA, pB)  # warm-up: allocates the pinned output buffers
    barrier()
    t0 = time.perf_counter()
    up = down = 0
    for _ in range(args.e2e_steps):
        _, up, down = ctx.multiply_host(pA, pB)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    ctx.set_option("release_workspace", 1)   # the staging copies of the host path are not needed any more

    # ------------------------------------------------------------------ N > 1: concatenation on GPU 0, one-GPU run
    if world > 1:
        # per-slab nnz to everybody (8 bytes per rank), then slabs -> rank 0 with NCCL send / recv
        counts = torch.zeros(world, dtype=torch.int64, device=dev)
        counts[rank] = nnzC
        dist.all_reduce(counts)
        cnt = [int(x) for x in counts.tolist()]
        total = sum(cnt)
        concat = {"total_nnz": total}
        if total < 2 ** 32:
            class _Arr:   # library-owned device array -> torch tensor through the CUDA array interface
                def __init__(self, ptr, n, typestr):
                    self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}

            def dview(ptr, n, typestr, dtype):
                ptr = getattr(ptr, "value", ptr)
                return torch.as_tensor(_Arr(ptr, n, typestr), device=dev) if (ptr and n) else torch.zeros(n, dtype=dtype, device=dev)
            c_rp = dview(dC.s.row_offsets if nnzC else None, slab.rows + 1, "<i4", torch.int32)
            c_ci = dview(dC.s.col_ids, nnzC, "<i4", torch.int32)
            c_v = dview(dC.s.data, nnzC, "<f8", torch.float64)
            # checksums of the slabs, to verify the concatenated arrays on rank 0
            chk = torch.stack([c_ci.sum(dtype=torch.int64).to(torch.float64), c_v.sum(dtype=torch.float64)])
            dist.all_reduce(chk)

            def concat_push():
                """Own kernel over peer memory: rank 0 exports the arrays of the concatenated C through CUDA IPC, every
                rank pushes its slab into them (16-byte peer stores over NVLink, row_offsets fix-up fused)."""
                ptrs, mine, ok = None, rank == 0, 1
                try:
                    if rank == 0:
                        ptrs = [ctx._alloc((A.rows + 1) * 4), ctx._alloc(max(total, 1) * 4), ctx._alloc(max(total, 1) * 8)]
                        hb = b"".join(ctx.ipc_export(p) for p in ptrs)
                        ht = torch.tensor(list(hb), dtype=torch.uint8, device=dev)
                    else:
                        ht = torch.empty(3 * ctx.IPC_HANDLE_BYTES, dtype=torch.uint8, device=dev)
                except Exception as e:   # keep the collectives below matched on every rank
                    ok, ht = 0, torch.zeros(3 * ctx.IPC_HANDLE_BYTES, dtype=torch.uint8, device=dev)
                    sys.stderr.write(f"[bench] rank {rank}: IPC export failed: {e}\n")
                dist.broadcast(ht, 0)
                if rank != 0 and ok:
                    try:
                        hb = bytes(ht.cpu().numpy().tobytes())
                        n = ctx.IPC_HANDLE_BYTES
                        ptrs = [ctx.ipc_open(hb[i * n:(i + 1) * n]) for i in range(3)]
                    except Exception as e:
                        ok = 0
                        sys.stderr.write(f"[bench] rank {rank}: IPC open failed: {e}\n")
                flag = torch.tensor([ok], dtype=torch.int32, device=dev)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
                res = None
                if int(flag.item()):
                    base, last = sum(cnt[:rank]), rank == world - 1
                    ms, bad = 0.0, 0
                    try:
                        ctx.push_slab(dC, base, int(cuts[rank]), last, *ptrs)      # untimed: first touch of the peer mapping
                    except Exception as e:   # every rank still reaches the barriers below
                        bad = 1
                        sys.stderr.write(f"[bench] rank {rank}: push failed: {e}\n")
                    barrier()
                    try:
                        if not bad:
                            ms = ctx.push_slab(dC, base, int(cuts[rank]), last, *ptrs)  # CUDA-event time of this rank's kernel
                    except Exception as e:
                        bad = 1
                        sys.stderr.write(f"[bench] rank {rank}: push failed: {e}\n")
                    barrier()
                    t = torch.tensor([ms, float(bad)], dtype=torch.float64, device=dev)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    if t[1].item() == 0:
                        res = {"ms": float(t[0].item()), "method": "own push kernel: 16-byte peer stores over NVLink into "
                               "IPC-opened arrays on GPU 0, row_offsets fix-up fused (speck_b200_push_slab_f64)"}
                    if res is not None and rank == 0:
                        g_rp = dview(ptrs[0], A.rows + 1, "<i4", torch.int32)
                        g_ci = dview(ptrs[1], total, "<i4", torch.int32)
                        g_v = dview(ptrs[2], total, "<f8", torch.float64)
                        got = torch.stack([g_ci.sum(dtype=torch.int64).to(torch.float64), g_v.sum(dtype=torch.float64)])
                        res["row_offsets_last_equals_total"] = bool((int(g_rp[-1].item()) & 0xffffffff) == total)
                        res["row_offsets_monotone"] = bool((g_rp.to(torch.int64) & 0xffffffff).diff().ge(0).all().item())
                        res["checksums_match_slabs"] = bool(got[0].item() == chk[0].item()
                                                            and abs(got[1].item() - chk[1].item()) <= 1e-9 * abs(chk[1].item()))
                        del g_rp, g_ci, g_v
                    torch.cuda.synchronize()
                if rank != 0 and ptrs:
                    for p_ in ptrs:
                        ctx.ipc_close(p_)
                barrier()
                if rank == 0 and ptrs:
                    for p_ in ptrs:
                        ctx.lib.speck_b200_free(ctx.h, p_)
                return res

            def concat_nccl():
                if rank == 0:
                    g_rp = torch.empty(A.rows + 1, dtype=torch.int32, device=dev)
                    g_ci = torch.empty(total, dtype=torch.int32, device=dev)
                    g_v = torch.empty(total, dtype=torch.float64, device=dev)
                # NCCL point-to-point connections are set up lazily on first use: warm them up outside the timed region
                warm = torch.zeros(4, dtype=torch.int32, device=dev)
                ops = ([dist.P2POp(dist.irecv, torch.zeros(4, dtype=torch.int32, device=dev), r) for r in range(1, world)]
                       if rank == 0 else [dist.P2POp(dist.isend, warm, 0)])
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                if rank == 0:
                    # every slab lands at its final position (one grouped NCCL receive), row_offsets get the slab's base added
                    tmp_rp, ops, base, bases = {}, [], 0, []
                    for r in range(world):
                        r0, r1 = int(cuts[r]), int(cuts[r + 1])
                        bases.append(base)
                        if r > 0:
                            tmp_rp[r] = torch.empty(r1 - r0 + 1, dtype=torch.int32, device=dev)
                            ops.append(dist.P2POp(dist.irecv, tmp_rp[r], r))
                            if cnt[r]:
                                ops.append(dist.P2POp(dist.irecv, g_ci[base:base + cnt[r]], r))
                                ops.append(dist.P2POp(dist.irecv, g_v[base:base + cnt[r]], r))
                        base += cnt[r]
                    works = dist.batch_isend_irecv(ops) if ops else []
                    g_ci[:cnt[0]] = c_ci
                    g_v[:cnt[0]] = c_v
                    g_rp[int(cuts[0]):int(cuts[1]) + 1] = c_rp
                    for w in works:
                        w.wait()
                    for r in range(1, world):     # offset fix-up (values below 2^32 wrap correctly in int32)
                        g_rp[int(cuts[r]):int(cuts[r + 1]) + 1] = tmp_rp[r] + bases[r]
                else:
                    ops = [dist.P2POp(dist.isend, c_rp, 0)]
                    if nnzC:
                        ops += [dist.P2POp(dist.isend, c_ci, 0), dist.P2POp(dist.isend, c_v, 0)]
                    for w in dist.batch_isend_irecv(ops):
                        w.wait()
                e1.record()
                barrier()
                concat_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
                dist.all_reduce(concat_ms, op=dist.ReduceOp.MAX)
                res = {"ms": float(concat_ms.item()), "method": "NCCL grouped send / recv + offset add"}
                if rank == 0:
                    res["row_offsets_last_equals_total"] = bool((int(g_rp[-1].item()) & 0xffffffff) == total)
                    del g_rp, g_ci, g_v
                return res

            res = concat_push() if args.concat in ("push", "both") else None
            if res is None or args.concat == "both":
                r2 = concat_nccl()
                if res is None:
                    res = r2
                else:
                    res["nccl_send_recv_ms"] = r2["ms"]
            concat.update(res)
            concat["bytes_moved"] = int(12 * (total - cnt[0]) + 4 * (A.rows + world))
        else:
            concat["skipped"] = "total nnz(C) >= 2^32: C stays distributed (u32 row_offsets of the spECK API)"
        extra["concat_on_gpu0"] = concat
        dC.free()
        dA.free()
        # the same matrix on ONE GPU (rank 0 alone, the other ranks wait): the strong-scaling denominator
        one = torch.zeros(2, dtype=torch.float64, device=dev)
        if rank == 0:
            dA1 = api.DeviceCSR.from_pointers(ctx, dB.rows, dB.cols, dB.nnz, dB.s.row_offsets, dB.s.col_ids, dB.s.data)
            dC1 = api.DeviceCSR(ctx)
            ms1, _, stage1, _, st1 = timed_multiplies(torch, ctx, dA1, dB, dC1, max(3, args.steps // 2), 3, stream, lambda: torch.cuda.synchronize())
            one[0] = ms1 / max(3, args.steps // 2)
            extra["one_gpu_same_matrix"] = {"ms_per_step": float(one[0].item()), "gflops": 2.0 * st1["products"] / (float(one[0].item()) * 1e-3) / 1e9,
                                            "stage_ms": stage1}
            dC1.free()
            ctx.set_option("release_workspace", 1)
        dist.broadcast(one, 0)
        # round-1 figure kept as an extra: weak scaling, one R-MAT-20 block per rank against B = block 0
        if not args.no_weak:
            Aw = load_workload("rmat20", 20 + rank)
            dBw, keepBw, _ = bcast_csr(Aw)
            dAw = ctx.upload(Aw)
            dCw = api.DeviceCSR(ctx)
            msw, _, _, _, stw = timed_multiplies(torch, ctx, dAw, dBw, dCw, args.steps, 3, stream, barrier)
            red = torch.tensor([msw], dtype=torch.float64, device=dev)
            dist.all_reduce(red, op=dist.ReduceOp.MAX)
            tot = torch.tensor([float(stw["products"])], dtype=torch.float64, device=dev)
            dist.all_reduce(tot)
            extra["weak_scaling_rmat20_blocks"] = {"gflops": 2.0 * float(tot.item()) * args.steps / (float(red.item()) * 1e-3) / 1e9,
                                                   "ms_per_step": float(red.item()) / args.steps,
                                                   "workload": "rank r owns R-MAT-20 block seed 20+r, B = block 0 broadcast at setup"}
            dCw.free(), dAw.free()
            del keepBw

    # ------------------------------------------------------------------ reduce over ranks: max time, sums
    if world > 1:
        red = torch.tensor([dev_ms, e2e_s, wall_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s, wall_ms = (float(x) for x in red.tolist())
        mine = torch.zeros(world, 3, dtype=torch.float64, device=dev)
        mine[rank] = torch.tensor([float(P), stage["numeric"], stage["symbolic"]], dtype=torch.float64)
        dist.all_reduce(mine)
        tot = torch.tensor([P, nnzC, launches, up, down], dtype=torch.float64, device=dev)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        P_all, nnzC_all, launches_all, up_all, down_all = (int(x) for x in tot.tolist())
        per_rank_P = [int(x) for x in mine[:, 0].tolist()]
        extra["per_rank"] = {"products": per_rank_P, "imbalance_max_over_mean": max(per_rank_P) / (sum(per_rank_P) / world),
                             "numeric_ms": [float(x) for x in mine[:, 1].tolist()], "symbolic_ms": [float(x) for x in mine[:, 2].tolist()]}
    else:
        P_all, nnzC_all, launches_all, up_all, down_all = P, nnzC, launches, up, down

    if rank == 0:
        ms_step = dev_ms / args.steps
        gflops = 2.0 * P_all * args.steps / (dev_ms * 1e-3) / 1e9
        nb = api.numeric_bytes(slab.rows, slab.nnz, P, nnzC)   # this rank's slab (N = 1: the whole multiply)
        achieved = nb / (stage["numeric"] * 1e-3) / 1e9
        traffic, traffic_src = traffic_entry(wl) if world == 1 else (None, None)
        line = {
            "metric": "SpGEMM GFLOPS (2*P/t), C=A.A", "value": gflops, "unit": "GFLOPS",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {
                "workload": WORKLOAD_DESC.get(wl, wl) if world == 1 else
                WORKLOAD_DESC.get(wl, wl) + f"; ONE matrix, rows cut into {world} product-balanced slabs on the device, "
                f"B = A NCCL-broadcast at setup ({bcast_ms:.1f} ms, untimed), no collective in the timed region",
                "rows": int(cuts[-1]), "nnz_a": int(hostB.nnz), "products": P_all, "nnz_c": nnzC_all,
                "l2": "inputs and outputs exceed the 126 MB L2 (C alone is 7-8 GB); no explicit flush",
                "parallelism": f"row-shard x{world}, B replicated",
                "class_rows": st["class_rows"],
            },
            "stage_ms": dict(stage, wall_per_step=wall_ms / args.steps),
            "roofline": {"bound": "hbm", "kernel": "numeric phase = one launch group: k_map_rows_cta (rows of 513..16384 products) + k_map_rows (<= 512) + k_dense_rows + k_direct"
                         + ("" if world == 1 else " (rank 0's slab)"),
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes": nb, "peak_source": peak_src,
                         "whole_multiply_frac": api.total_bytes(slab.rows, slab.nnz, P, nnzC) / (ms_step * 1e-3) / 1e9 / peak},
            "e2e": {"value": 2.0 * P_all * args.e2e_steps / e2e_s / 1e9, "unit": "GFLOPS",
                    "h2d_bytes_per_step": up_all, "d2h_bytes_per_step": down_all,
                    "ms_per_step": e2e_s / args.e2e_steps * 1e3, "steps": args.e2e_steps},
            "gpu_launches": launches_all,
            "clocks": clocks,
        }
        if world > 1:
            t1 = float(one[0].item())
            extra["strong_scaling"] = {"t1_ms": t1, "tN_ms": ms_step, "speedup": t1 / ms_step, "efficiency": t1 / ms_step / world,
                                       "note": "same matrix, rank 0 alone vs N slabs; limiter: the slab with the most small-row products (per-row fixed cost) and the analysis of its rows"}
            if "ms" in extra.get("concat_on_gpu0", {}):
                extra["strong_scaling"]["gflops_with_concat"] = 2.0 * P_all / ((ms_step + extra["concat_on_gpu0"]["ms"]) * 1e-3) / 1e9
        line.update(extra)
        if not args.no_cpu_baseline and world == 1:
            import oracle
            oracle.build()
            cores = host_threads()
            Pc, tc = time_oracle(slab, hostB, args.cpu_reps)
            line["cpu_baseline"] = {"value": 2.0 * Pc / tc / 1e9, "unit": "GFLOPS", "cores": cores, "kind": "port",
                                    "sample": f"all rows of A ({slab.rows} rows, P={Pc}) x full B, mean of {args.cpu_reps} runs of {tc:.2f} s"}
        if world == 1:
            dC.free()
            dA.free()
            ctx.set_option("release_workspace", 1)
            if not args.no_ref_gpu:
                line["ref_speck_gpu"] = run_ref_gpu(wl)
                line["ref_speck_gpu"]["note"] = "reference spECK compiled for sm_100 (stock 49152/49152 and tuned dynamic smem 232448)"
            if not args.no_sweep:
                line["sweep"] = []
                for name in SWEEP:
                    try:
                        line["sweep"].append(sweep_one(torch, api, ctx, name, peak, ref=not args.no_ref_gpu))
                    except Exception as e:  # noqa: BLE001
                        line["sweep"].append({"workload": name, "error": repr(e)})
                    ctx.set_option("release_workspace", 1)
        print(json.dumps(line), flush=True)

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
