#!/usr/bin/env python
"""Headline benchmark: candidate architectures scored per second (WL features + GP posterior +
EI + top-5) — BASELINE.json's metric on its configs[1]: NAS-Bench-201-shaped cells, n_train=150,
WL depth chosen from h in {0..5} by marginal likelihood, a 1M-candidate EI pool per GPU.

    python bench.py --gpus 1 --steps 200 --warmup 10                     # this repo (B200)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference --gpus 1 --steps 5 --warmup 1      # CPU arm (oracle port)

A step is one pass of the hot path over one candidate pool: fused relabel->lookup->posterior->EI
kernel, device top-5, and (N>1) the top-5 all-gather.  Pools are resident in HBM before the
timed region; 16 distinct pools per GPU are rotated (256 MB > 126 MB L2) so no step finds its
input in L2.  Multi-GPU is weak scaling: every rank scores its own 1M-candidate shard.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

FAMILY = 1                       # NAS-Bench-201-shaped
N_TRAIN = 150
H_GRID = tuple(range(6))
TOP_N = 5
SURVEY_BYTES_PER_CANDIDATE = 36  # 32 B padded CSR record in + 4 B EI out
LAYOUT_BYTES_PER_CANDIDATE = 20  # this build: 16 B bit-CSR record in + 4 B EI out
ROTATING_POOLS = 16
METRIC = "candidate archs scored/sec (WL + GP posterior + EI + top-5)"
WORKLOAD_DESC = ""


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons (B200_PROFILING.md recipe).  Runs for the whole
    bench; `window()` keeps only the samples whose timestamp falls inside the timed region."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self, t_begin, t_end):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9 or f[1] != str(self.gpu_index):
                    continue
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                if ts < t_begin - 0.05 or ts > t_end + 0.05:
                    continue
                sm.append(float(f[2]))
                mx.append(float(f[3]))
                for nm, v in zip(self.NAMES, f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception as e:
            out["error"] = repr(e)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def make_training_set():
    """Seeded training cells (host generator: identical on every rank and in both arms)."""
    import torch
    from nasbowl_b200 import synth
    train = synth.generate_cells_host(FAMILY, seed=0, first_index=0, n_cells=N_TRAIN)
    y = torch.randn(N_TRAIN, generator=torch.Generator().manual_seed(0))
    return train, y


def cpu_port_rate(train, y, sample, repeats=1):
    """The oracle (CPU restatement of the reference's path) on a bounded sample: cand/s."""
    from oracle import gp_oracle
    st = gp_oracle.gp_fit(train, y.numpy().astype(np.float64), h_candidates=H_GRID, optimize_lik=True, max_lik=0.01)
    inc = gp_oracle.incumbent(st)
    t0 = time.perf_counter()
    for _ in range(repeats):
        mu, var = gp_oracle.gp_predict_diag(st, train, sample, chunk=2000)
        ei = gp_oracle.expected_improvement(mu, var, inc)
        top = gp_oracle.top_n_distinct(ei.astype(np.float32), TOP_N)
    dt = (time.perf_counter() - t0) / repeats
    return len(sample) / dt, st, ei, top


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([i.get("num_threads", 1) for i in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port — the Python reference tree
    does not travel to the GPU box) on the box's host cores, same config / metric / unit."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from nasbowl_b200 import synth
    train, y = make_training_set()
    S = args.ref_sample
    samples = [synth.generate_cells_host(FAMILY, 1234, 10_000 + s * S, S) for s in range(args.steps + args.warmup)]
    from oracle import gp_oracle
    st = gp_oracle.gp_fit(train, y.numpy().astype(np.float64), h_candidates=H_GRID, optimize_lik=True, max_lik=0.01)
    inc = gp_oracle.incumbent(st)

    def step(sample):
        mu, var = gp_oracle.gp_predict_diag(st, train, sample, chunk=2000)
        ei = gp_oracle.expected_improvement(mu, var, inc)
        return gp_oracle.top_n_distinct(ei.astype(np.float32), TOP_N)
    for s in range(args.warmup):
        step(samples[s])
    t0 = time.perf_counter()
    for s in range(args.steps):
        step(samples[args.warmup + s])
    dt = time.perf_counter() - t0
    value = S * args.steps / dt
    threads = blas_threads()
    sample_desc = "%d candidates per step (of the pool), n_train=%d, h*=%d" % (S, N_TRAIN, st.h)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "candidates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC + ", EI pool scoring",
                   "h_star": st.h, "sample_per_step": S, "host": "CPU only"},
        "cpu_baseline": {"value": value, "unit": "candidates/s", "cores": threads, "kind": "port", "sample": sample_desc},
        "e2e": {"value": value, "unit": "candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_b200(args):
    import torch
    import torch.distributed as dist
    from nasbowl_b200 import parallel
    from nasbowl_b200.bayesopt import GraphExpectedImprovement, GraphGP
    from nasbowl_b200.cells import CellBatch
    from nasbowl_b200.engine import Engine
    from nasbowl_b200.kernels import WeisfilerLehman

    rank, world, local = parallel.init_distributed()
    torch.cuda.set_device(local)
    eng = Engine.get(local)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    M = args.pool
    train, y = make_training_set()
    n_max, stride = train.n_max, (48 if train.n_max == 16 else 16)

    # ---- fit on rank 0, broadcast the pool model
    gp = GraphGP(train, y, [WeisfilerLehman(h=2)])
    fit_ms = bcast_ms = state_ms = None
    if rank == 0:
        gp.fit(wl_subtree_candidates=H_GRID, optimize_lik=True, max_lik=0.01)       # warm (first-call costs:
        gp._score_state()                                                           #  module loading, allocator)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        gp.reset_XY(train, y)
        gp.kernels[0].change_kernel_params({'h': 2})
        gp.likelihood = 1e-3
        gp.fit(wl_subtree_candidates=H_GRID, optimize_lik=True, max_lik=0.01)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        pm = gp._score_state()
        torch.cuda.synchronize()
        fit_ms, state_ms = 1e3 * (t1 - t0), 1e3 * (time.perf_counter() - t1)
    else:
        pm = None
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pm = parallel.broadcast_pool_model(pm, src=0)
        torch.cuda.synchronize()
        bcast_ms = 1e3 * (time.perf_counter() - t0)
        if rank != 0:
            gp.set_pool_model(pm)
    acq = GraphExpectedImprovement(gp)
    h_star = pm.h

    # ---- resident pools: ROTATING_POOLS distinct shards per rank, generated on device
    pools = []
    for p in range(ROTATING_POOLS):
        first = 10_000 + (p * world + rank) * M
        pools.append((eng.generate_cells(FAMILY, 1234, first, M), first))
    torch.cuda.synchronize()

    def step(i, ev=None):
        cells, first = pools[i % ROTATING_POOLS]
        acq.iters += 1
        if ev is not None:
            ev[0].record()
        scores, _, _, _ = eng.score_pool(model, cells, M, scores=score_buf)
        if ev is not None:
            ev[1].record()
        return parallel.select_topk(eng, scores, TOP_N, True, first, cand_buf, gather_buf)

    model = acq._model()
    score_buf = eng.empty((M,), torch.float32)
    cand_buf = eng.empty((TOP_N * 16,), torch.uint8)
    gather_buf = eng.empty((world * TOP_N * 16,), torch.uint8)
    for i in range(args.warmup):
        step(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = eng.launches
    t_begin = time.time()
    kernel_events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                     for _ in range(args.steps)]
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for i in range(args.steps):
        step(args.warmup + i, kernel_events[i])
    end.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    elapsed_ms = start.elapsed_time(end)
    launches = eng.launches - launches0
    clocks = sampler.stop(t_begin, time.time()) if rank == 0 else None
    kernel_ms = sum(a.elapsed_time(b) for a, b in kernel_events) / args.steps

    # ---- end to end through the public API with HOST buffers (pinned): H2D pool, score, top-5, D2H scores
    host_pool = torch.empty((M, stride), dtype=torch.uint8).pin_memory()
    host_pool.copy_(pools[0][0].cpu())

    def e2e_step():
        # host records in, top-5 indices + all EIs (host numpy) out
        xs, eis, idx = acq.propose_location(host_pool, top_n=TOP_N)
        return idx
    # (the clock sampler's shutdown above idles the GPU for ~0.15 s: warm up long enough to ramp clocks
    # and the PCIe link again before timing)
    for _ in range(5 * max(3, args.warmup)):
        e2e_step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e_steps = max(3, args.steps)
    t0 = time.perf_counter()
    for _ in range(e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_ms = 1e3 * (time.perf_counter() - t0)

    # ---- CPU baseline + parity of the timed path on a bounded sample (rank 0, N=1 only).  Runs AFTER the GPU
    # measurements: its BLAS worker threads keep spinning for a while and would slow the per-step host
    # work (launch, top-5 readback) of the timed loops — the e2e figure dropped by up to 35 % when it ran first.
    cpu_baseline, parity = None, None
    if rank == 0 and world == 1 and not args.skip_cpu:
        S = args.cpu_sample
        sample = CellBatch.from_records(pools[0][0][:S].cpu().numpy(), n_max)
        rate, st, ei_ref, top_ref = cpu_port_rate(train, y, sample)
        sc, _, _, sc64 = acq.score_pool(pools[0][0][:S].contiguous(), want64=True)
        err = float(np.max(np.abs(sc64.cpu().numpy() - ei_ref) / (np.abs(ei_ref) + 1e-12)))
        parity = {"sample": S, "max_rel_err_ei_fp64": err, "h_star_equal": bool(st.h == h_star)}
        cpu_baseline = {"value": rate, "unit": "candidates/s", "cores": blas_threads(), "kind": "port",
                        "sample": "first %d candidates of pool 0 (oracle/gp_oracle.py: joint WL + fp64 GP + EI + top-5)" % S}

    # ---- max over ranks
    if world > 1:
        t = torch.tensor([elapsed_ms, e2e_ms, kernel_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, e2e_ms, kernel_ms = [float(x) for x in t.cpu()]
        lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    peak, peak_src = load_peaks()
    traffic = None
    try:       # DRAM bytes per launch from the committed ncu capture of the same kernel / pool size
        tr = json.load(open(os.path.join(ROOT, "profiles", "score_traffic.json")))
        if tr["pool"] == M and tr["h"] == h_star and args.workload == "cfg2":
            traffic = tr["traffic_bytes_per_launch"]
    except Exception:
        pass
    value = world * M * args.steps / (elapsed_ms / 1e3)
    achieved = SURVEY_BYTES_PER_CANDIDATE * M / (kernel_ms / 1e3) / 1e9
    # SURVEY.md §8(d)'s composite roofline for the fused scoring kernel: per candidate
    # t_roof = max(bytes / BW_hbm, 2 n D_pad / P_int8 + (2 n^2 + 2 n) / P_bf16) with the DENSE algorithmic
    # op counts (this build evaluates the same posterior in feature space and does far fewer operations)
    try:
        bf16 = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"]) * 1e12
    except Exception:
        bf16 = 1590e12
    n_tr, d_pad = N_TRAIN, pm.n_features
    t_roof = max(SURVEY_BYTES_PER_CANDIDATE / (peak * 1e9), 2.0 * n_tr * d_pad / (2 * bf16) + (2.0 * n_tr ** 2 + 2 * n_tr) / bf16)
    survey_8d = {"t_roof_ns_per_candidate": t_roof * 1e9, "t_measured_ns_per_candidate": kernel_ms * 1e6 / M,
                 "frac": t_roof / (kernel_ms * 1e-3 / M), "int8_peak": "2 x measured bf16 (no int8 entry in MEASURED_PEAKS.json)"}
    line = {
        "metric": METRIC, "value": value, "unit": "candidates/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC + ", %d-candidate EI pool per GPU, top-5" % M,
                   "h_star": h_star, "n_features": pm.n_features, "pool_per_gpu": M,
                   "l2": "inputs larger than L2: %d rotating pools x %.0f MB per GPU" % (ROTATING_POOLS, M * stride / 1e6),
                   "fit_ms": fit_ms, "score_state_ms": state_ms, "broadcast_ms": bcast_ms, "model_bytes": pm.nbytes()},
        "roofline": {"bound": "hbm", "kernel": "score_pool_kernel<%d,%d,prefix>" % (n_max, h_star + 1), "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "kernel_ms": kernel_ms, "bytes_per_unit": SURVEY_BYTES_PER_CANDIDATE,
                     "layout_bytes_per_unit": LAYOUT_BYTES_PER_CANDIDATE,
                     "kernel_share_of_step": kernel_ms / (elapsed_ms / args.steps), "survey_8d": survey_8d,
                     "note": "not HBM-bound: ncu shows DRAM 0.3 %, issue slots 75 % active (64-bit integer hashing); see profiles/r01f_summary.md"},
        "cpu_baseline": cpu_baseline,
        "e2e": {"value": world * M * e_steps / (e2e_ms / 1e3), "unit": "candidates/s",
                "h2d_bytes_per_step": M * stride, "d2h_bytes_per_step": M * 4 + TOP_N * 12, "steps": e_steps,
                "api": "GraphExpectedImprovement.propose_location on pinned host records"},
        "gpu_launches": launches,
        "clocks": clocks,
        "parity": parity,
    }
    print(json.dumps(line))


WORKLOADS = {
    # BASELINE.json configs[1] (the headline), configs[2] and configs[3] per-GPU shards
    "cfg2": dict(family=1, n_train=150, h_grid=tuple(range(6)), pool=1_000_000, bytes=36,
                 desc="NB201-shaped cells (8 nodes, 5 ops), n_train=150, h in {0..5} by marginal likelihood"),
    "cfg3": dict(family=0, n_train=500, h_grid=(2,), pool=1_250_000, bytes=36,
                 desc="NB101-shaped cells (<=7 nodes, 5 ops), n_train=500, WL h=2 (1/8 of the 10M pool per GPU)"),
    "cfg4": dict(family=2, n_train=1000, h_grid=(2,), pool=1_250_000, bytes=68,
                 desc="DARTS-shaped cells (<=15 nodes, 11 ops), n_train=1000, WL h=2 (1/8 of the 10M pool per GPU)"),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pool", type=int, default=None)
    ap.add_argument("--cpu-sample", type=int, default=None)
    ap.add_argument("--ref-sample", type=int, default=1000)
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    g = globals()
    g["FAMILY"], g["N_TRAIN"], g["H_GRID"] = w["family"], w["n_train"], w["h_grid"]
    g["SURVEY_BYTES_PER_CANDIDATE"] = w["bytes"]
    g["LAYOUT_BYTES_PER_CANDIDATE"] = (48 if w["family"] == 2 else 16) + 4
    g["WORKLOAD_DESC"] = w["desc"]
    if args.pool is None:
        args.pool = w["pool"]
    if args.cpu_sample is None:
        args.cpu_sample = {"cfg2": 16000, "cfg3": 8000, "cfg4": 4000}[args.workload]
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
