#!/usr/bin/env python
"""bench.py -- SpGEMM GFLOPS (2*P/t) of C = A.B on B200, one JSON line on stdout (rank 0).

A "step" is one complete multiply (analysis -> symbolic -> scan -> numeric) of the workload.
  N = 1 : BASELINE.json configs[1]: R-MAT scale 20, edge factor 16, (0.45,0.15,0.15,0.25),
          seed 20, fp64, C = A.A   (P = 586 218 280, nnz(C) = 582 205 696).
  N > 1 : weak scaling by rows of A: rank r owns the row block A_r = R-MAT(seed 20+r)
          of the row-stacked A, B = R-MAT(seed 20) is generated on rank 0 and NCCL-broadcast
          once at setup (not timed); every rank computes its slab C_r = A_r.B, no collective in
          the timed region; value = 2*sum_r(P_r)*K / max_r(t).
Keys follow the driver contract; `roofline` is the numeric phase (all numeric kernels run
concurrently on the library's streams and are bracketed by CUDA events on those streams);
`cpu_baseline` is the CPU oracle (a port: the reference has no CPU SpGEMM) on a bounded row
sample of the same workload.  `--impl reference` times that CPU oracle on all host cores.
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


def load_workload(name, seed):
    """Deterministic workload, cached as .npz so both arms of one run share the generation cost."""
    os.makedirs(CACHE, exist_ok=True)
    path = os.path.join(CACHE, f"{name}_s{seed}.npz")
    if os.path.exists(path):
        try:
            z = np.load(path)
            return HostCSR(int(z["rows"]), int(z["cols"]), z["rp"], z["ci"], z["v"])
        except Exception:
            pass
    if name == "rmat20":
        A = M.rmat(20, 16, seed=seed)
    elif name == "rmat18":
        A = M.rmat(18, 16, seed=seed)
    elif name == "rmat16":
        A = M.rmat(16, 16, seed=seed)
    elif name == "rmat24":
        A = M.rmat(24, 4, seed=seed)
    elif name == "webbase_like":
        A = M.webbase_like(seed=seed)
    elif name == "cant_like":
        A = M.fem3d_like(seed=seed)
    elif name == "banded_like":
        A = M.banded_fem_like(seed=seed)
    elif name == "econ_like":
        A = M.econ_like(seed=seed)
    elif name == "circuit_like":
        A = M.circuit_like(seed=seed)
    else:
        raise SystemExit(f"unknown workload {name}")
    try:
        tmp = path + f".{os.getpid()}.tmp.npz"
        np.savez(tmp, rows=A.rows, cols=A.cols, rp=A.row_offsets, ci=A.col_ids, v=A.data)
        os.replace(tmp, path)
    except Exception:
        pass
    return A


WORKLOAD_DESC = {
    "rmat20": "R-MAT scale 20 ef 16 (0.45,0.15,0.15,0.25) seed 20, fp64, C=A.A (BASELINE configs[1])",
}


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
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
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


def cpu_sample(A: HostCSR, stride):
    """Bounded CPU sample: every `stride`-th row of A against the full B = A."""
    idx = np.arange(0, A.rows, stride)
    lens = (A.row_offsets[idx + 1] - A.row_offsets[idx]).astype(np.int64)
    rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint32)
    starts = A.row_offsets[idx].astype(np.int64)
    take = np.repeat(starts - rp[:-1].astype(np.int64), lens) + np.arange(int(rp[-1]), dtype=np.int64)
    return HostCSR(idx.size, A.cols, rp, A.col_ids[take], A.data[take])


def time_oracle(As: HostCSR, B: HostCSR, reps=1):
    """mean seconds per multiply of the CPU oracle over `reps` repetitions (after one warm-up)"""
    import oracle
    _, _, P, _ = oracle.row_products(As.row_offsets, As.col_ids, B.row_offsets)
    oracle.spgemm(As.row_offsets, As.col_ids, As.data, B.row_offsets, B.col_ids, B.data, B.cols)
    t0 = time.perf_counter()
    for _ in range(reps):
        oracle.spgemm(As.row_offsets, As.col_ids, As.data, B.row_offsets, B.col_ids, B.data, B.cols)
    return P, (time.perf_counter() - t0) / reps


def time_reference_gpu(args):
    """The reference's own CUDA build (sm_100, oracle/_ref) on the same GPU and workload, in a
    subprocess (a crash of the reference must not take the bench down).  Reported next to our line;
    it is NOT the `--impl reference` arm (that one is the CPU port, per the bench contract)."""
    out = {}
    for variant in ("stock", "tuned"):
        so = os.path.join(ROOT, "oracle", "_ref", f"libspeck_ref_{variant}.so")
        if not os.path.exists(so):
            out[variant] = {"unavailable": "oracle/_ref not built (needs /root/reference at build time)"}
            continue
        try:
            r = subprocess.run([sys.executable, "-m", "oracle.ref_run", "--workload", args.workload, "--seed", str(args.seed),
                                "--variant", variant, "--warmup", "3", "--iters", "5"],
                               capture_output=True, text=True, timeout=600, cwd=ROOT)
            js = [l for l in r.stdout.splitlines() if l.startswith("{")]
            d = json.loads(js[-1]) if js else {"error": r.stderr[-200:]}
            out[variant] = {k: d[k] for k in ("mean_ms", "min_ms", "gflops_mean", "gflops_best", "nnz_c") if k in d} or d
        except Exception as e:  # noqa: BLE001
            out[variant] = {"error": repr(e)}
    best = max((v.get("gflops_mean", 0.0) for v in out.values()), default=0.0)
    out["best_gflops_mean"] = best
    out["note"] = "reference spECK compiled for sm_100 (stock 49152/49152 and tuned dynamic smem 232448), timings.complete, C reused"
    return out


def run_reference_arm(args, rank, world):
    """CPU oracle (port; the reference has no CPU SpGEMM) on all host cores, rank 0 only."""
    if rank != 0:
        return
    import oracle
    oracle.build()
    A = load_workload(args.workload, args.seed)
    stride = args.cpu_stride
    As = cpu_sample(A, stride) if stride > 1 else A
    cores = oracle.num_threads()
    _, _, P, _ = oracle.row_products(As.row_offsets, As.col_ids, A.row_offsets)
    for _ in range(min(args.warmup, 1)):
        oracle.spgemm(As.row_offsets, As.col_ids, As.data, A.row_offsets, A.col_ids, A.data, A.cols)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.spgemm(As.row_offsets, As.col_ids, As.data, A.row_offsets, A.col_ids, A.data, A.cols)
    dt = time.perf_counter() - t0
    gflops = 2.0 * P * args.steps / dt / 1e9
    sample = (f"every {stride}th row of A" if stride > 1 else "all rows of A") + f" ({As.rows} rows, P={P}) x full B, per step"
    line = {
        "impl": "reference", "metric": "SpGEMM GFLOPS (2*P/t), C=A.A", "value": gflops, "unit": "GFLOPS",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC.get(args.workload, args.workload), "sample": sample},
        "cpu_baseline": {"value": gflops, "unit": "GFLOPS", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": gflops, "unit": "GFLOPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference (spECK) has no CPU SpGEMM; this arm is the CPU oracle port (oracle/spgemm_oracle.c, OpenMP)",
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="rmat20")
    ap.add_argument("--seed", type=int, default=20)
    ap.add_argument("--cpu-stride", type=int, default=1, help="row stride of the bounded CPU sample (1 = the whole workload)")
    ap.add_argument("--cpu-reps", type=int, default=10, help="repetitions of the CPU sample in the GPU arm (about 10-20 s in total)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-gpu", action="store_true",
                    help="skip timing the reference spECK CUDA build (oracle/_ref) on the same GPU")
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
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    ctx = api.Context(local_rank)
    if args.sort_max:
        ctx.set_option("sort_max", args.sort_max)
    for kv in args.opt:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))

    # ---------------- setup (untimed): A_r on every rank, B on rank 0 -> NCCL broadcast
    A = load_workload(args.workload, args.seed + rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if rank == 0:
            B = A
            meta = torch.tensor([B.rows, B.cols, B.nnz], dtype=torch.int64, device=dev)
        else:
            B = None
            meta = torch.zeros(3, dtype=torch.int64, device=dev)
        dist.broadcast(meta, 0)
        b_rows, b_cols, b_nnz = (int(x) for x in meta.tolist())
        if rank == 0:
            t_rp = torch.from_numpy(B.row_offsets.view(np.int32)).to(dev)
            t_ci = torch.from_numpy(B.col_ids.view(np.int32)).to(dev)
            t_v = torch.from_numpy(B.data).to(dev)
        else:
            t_rp = torch.empty(b_rows + 1, dtype=torch.int32, device=dev)
            t_ci = torch.empty(b_nnz, dtype=torch.int32, device=dev)
            t_v = torch.empty(b_nnz, dtype=torch.float64, device=dev)
        t0 = time.perf_counter()
        for t in (t_rp, t_ci, t_v):
            dist.broadcast(t, 0)
        torch.cuda.synchronize()
        bcast_ms = (time.perf_counter() - t0) * 1e3
        dB = api.DeviceCSR.from_pointers(ctx, b_rows, b_cols, b_nnz, t_rp.data_ptr(), t_ci.data_ptr(), t_v.data_ptr())
        dA = ctx.upload(A)
        hostB = HostCSR(b_rows, b_cols, t_rp.cpu().numpy().view(np.uint32), t_ci.cpu().numpy().view(np.uint32),
                        t_v.cpu().numpy()) if rank != 0 else A
    else:
        bcast_ms = 0.0
        dA = ctx.upload(A)
        dB = dA
        hostB = A

    dC = api.DeviceCSR(ctx)
    stream = torch.cuda.ExternalStream(ctx.lib.speck_b200_stream(ctx.h), device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()          # nvidia-smi needs ~100 ms to start: begin before the warm-up
    for _ in range(args.warmup):
        ctx.multiply(dA, dB, dC)
    st = ctx.stats()
    P, nnzC = st["products"], st["nnz_c"]

    # ---------------- timed region: exactly K steps, CUDA events on the library's stream
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    num_ms, sym_ms, ana_ms, scan_ms, launches = [], [], [], [], 0
    t_wall = time.perf_counter()
    ev0.record(stream)
    for _ in range(args.steps):
        ctx.multiply(dA, dB, dC)
        s = ctx.stats()
        num_ms.append(s["ms_numeric"]), sym_ms.append(s["ms_symbolic"]), ana_ms.append(s["ms_analysis"])
        scan_ms.append(s["ms_scan"])
        launches += s["kernel_launches"]
    ev1.record(stream)
    barrier()
    wall_ms = (time.perf_counter() - t_wall) * 1e3
    dev_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None

    # ---------------- end to end through the host-buffer C ABI (pinned inputs, C downloaded)
    def pin(a):
        t = torch.from_numpy(a).pin_memory()
        return t, t.numpy()
    keep = []
    def pinned(m):
        out = []
        for arr in (m.row_offsets, m.col_ids, m.data):
            t, n = pin(np.ascontiguousarray(arr).view(np.int32) if arr.dtype == np.uint32 else np.ascontiguousarray(arr))
            keep.append(t)
            out.append(n.view(np.uint32) if arr.dtype == np.uint32 else n)
        return HostCSR(m.rows, m.cols, out[0], out[1], out[2])
    pA = pinned(A)
    pB = pA if hostB is A else pinned(hostB)
    ctx.multiply_host(pA, pB)  # warm-up: allocates the pinned output buffers
    barrier()
    t0 = time.perf_counter()
    up = down = 0
    for _ in range(args.e2e_steps):
        _, up, down = ctx.multiply_host(pA, pB)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0

    # ---------------- reduce over ranks: max time, sum of products
    if world > 1:
        red = torch.tensor([dev_ms, e2e_s, wall_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s, wall_ms = (float(x) for x in red.tolist())
        tot = torch.tensor([P, nnzC, launches, up, down], dtype=torch.float64, device=dev)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        P_all, nnzC_all, launches_all = int(tot[0].item()), int(tot[1].item()), int(tot[2].item())
        up_all, down_all = int(tot[3].item()), int(tot[4].item())
    else:
        P_all, nnzC_all, launches_all, up_all, down_all = P, nnzC, launches, up, down

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        nb = api.numeric_bytes(A.rows, A.nnz, P, nnzC)
        t_num = float(np.mean(num_ms)) * 1e-3
        achieved = nb / t_num / 1e9
        gflops = 2.0 * P_all * args.steps / (dev_ms * 1e-3) / 1e9
        line = {
            "metric": "SpGEMM GFLOPS (2*P/t), C=A.A", "value": gflops, "unit": "GFLOPS",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": WORKLOAD_DESC.get(args.workload, args.workload) if world == 1 else
                f"{args.workload}: rank r owns row block A_r (seed {args.seed}+r), B = block 0 NCCL-broadcast at setup ({bcast_ms:.1f} ms, untimed)",
                "rows": A.rows, "nnz_a": A.nnz, "products": P_all, "nnz_c": nnzC_all,
                "l2": "inputs (A 0.2 GB, C 7 GB) exceed the 126 MB L2; no explicit flush",
                "parallelism": f"row-shard x{world}, B replicated",
                "class_rows": st["class_rows"],
            },
            "stage_ms": {"analysis": float(np.mean(ana_ms)), "symbolic": float(np.mean(sym_ms)),
                         "scan": float(np.mean(scan_ms)), "numeric": float(np.mean(num_ms)),
                         "wall_per_step": wall_ms / args.steps},
            "roofline": {"bound": "hbm", "kernel": "numeric phase = one launch group: k_map_rows_cta (rows of 513..16384 products) + k_map_rows (<= 512) + k_dense_rows + k_direct",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "algorithmic_bytes": nb, "peak_source": peak_src},
            "e2e": {"value": 2.0 * P_all * args.e2e_steps / e2e_s / 1e9, "unit": "GFLOPS",
                    "h2d_bytes_per_step": up_all, "d2h_bytes_per_step": down_all,
                    "ms_per_step": e2e_s / args.e2e_steps * 1e3, "steps": args.e2e_steps},
            "gpu_launches": launches_all,
            "clocks": clocks,
        }
        if not args.no_cpu_baseline:
            import oracle
            oracle.build()
            As = cpu_sample(A, args.cpu_stride) if args.cpu_stride > 1 else A
            Pc, tc = time_oracle(As, hostB, args.cpu_reps)
            line["cpu_baseline"] = {"value": 2.0 * Pc / tc / 1e9, "unit": "GFLOPS", "cores": oracle.num_threads(),
                                    "kind": "port",
                                    "sample": (f"every {args.cpu_stride}th row of A" if args.cpu_stride > 1 else "all rows of A")
                                    + f" ({As.rows} rows, P={Pc}) x full B, mean of {args.cpu_reps} runs of {tc:.2f} s"}
        if not args.no_ref_gpu and world == 1:
            line["ref_speck_gpu"] = time_reference_gpu(args)
        prof = os.path.join(ROOT, "profiles", "traffic.json")
        try:
            line["roofline"]["traffic"] = json.load(open(prof)).get(args.workload)
        except Exception:
            pass
        print(json.dumps(line), flush=True)

    dC.free()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
