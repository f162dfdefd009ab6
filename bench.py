#!/usr/bin/env python
"""Benchmark of the surrogate-scoring hot path on every configuration BASELINE.json names.

    python bench.py --gpus 1 --steps 20 --warmup 5                      # this repo (B200)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference --gpus 1 --steps 5 --warmup 1     # CPU arm (oracle port)
    python bench.py --workload cfg3 --configs ""                        # one config only (profiling)

Headline (top level of the ONE JSON line rank 0 prints): BASELINE configs[1] — NAS-Bench-201-shaped cells,
n_train = 150, WL depth chosen from h in {0..5} by marginal likelihood, a 1M-candidate EI pool per GPU
(weak scaling), top-5.  `configs` holds one sub-record per further BASELINE config:
    cfg3  configs[2]: NB101-shaped 10M-candidate pool, n_train = 500, h = 2, sharded N ways (strong scaling)
    cfg4  configs[3]: DARTS normal + reduce cells (two records per candidate, summed WL kernels), n_train = 1000,
          10M-candidate pool sharded N ways (strong scaling)
    cfg5  configs[4]: Gram-only stress, 200k x 200k WL h = 3 kernel matrix, row blocks sharded N ways
each with value / e2e / roofline / parity (never null: oracle agreement on a slice, and on N > 1 equality of the
merged top-5 with one GPU's top-5 of the whole pool).

A step is one pass of the hot path over one candidate pool: the fused relabel -> lookup -> posterior -> EI ->
local top-5 kernel, the small device merge and (N > 1) the top-5 all-gather.  Pools are resident in HBM
before the timed region; distinct pools are rotated so that no step finds its input in L2.
"""
import argparse
import json
import math
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

TOP_N = 5
METRIC = "candidate archs scored/sec (WL + GP posterior + EI + top-5)"
L2_BYTES = 126e6

WORKLOADS = {
    # BASELINE.json configs[1] (the headline), configs[2], configs[3]
    "cfg2": dict(family=1, n_train=150, h_grid=tuple(range(6)), pool=1_000_000, scaling="weak", bytes=36, slots=1,
                 cpu_sample=16000,
                 desc="NB201-shaped cells (8 nodes, 5 ops), n_train=150, h in {0..5} by marginal likelihood"),
    "cfg3": dict(family=0, n_train=500, h_grid=(2,), pool=10_000_000, scaling="strong", bytes=36, slots=1,
                 cpu_sample=6000,
                 desc="NB101-shaped cells (<=7 nodes, 5 ops), n_train=500, WL h=2, 10M-candidate pool sharded over the GPUs"),
    "cfg4": dict(family=2, n_train=1000, h_grid=(2,), pool=10_000_000, scaling="strong", bytes=136, slots=2,
                 cpu_sample=4000,
                 desc="DARTS-shaped normal + reduce cells (<=15 nodes, 11 ops; two records per candidate), summed WL "
                      "kernels h=2 with weights 0.5/0.5, n_train=1000, 10M-candidate pool sharded over the GPUs"),
}
GRAM_STRESS = dict(cells=200_000, h=3, block=4096,      # row block: 4096 measured best (profiles/r02_gram_block_sweep.json)
                   desc="Gram-only stress: 200k x 200k WL h=3 kernel matrix on NB101-shaped cells")


def pool_sizes(w, world, pool_override=None):
    """(candidates in total, candidates of rank r as a function) of a workload at `world` GPUs — shared by both arms
    so that the CPU arm describes exactly the configuration the GPU arm ran."""
    total = w["pool"] * world if w["scaling"] == "weak" else (pool_override or w["pool"])
    if w["scaling"] == "weak" and pool_override:
        total = pool_override * world
    return total


def workload_config(w, per_gpu, total, h_star):
    """The static part of `config` (identical in the b200 arm and the --impl reference arm)."""
    return {"workload": w["desc"] + ", %d-candidate EI pool per GPU (%d in total), top-5" % (per_gpu, total),
            "h_star": h_star, "pool_per_gpu": per_gpu, "pool_total": total}


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p, "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md)"


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

    def window(self, t_begin, t_end):
        """Clocks between two wall-clock stamps (the sampler keeps running)."""
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.12)
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
        except Exception as e:
            out["error"] = repr(e)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out

    def stop(self):
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            os.unlink(self.path)
        except Exception:
            pass


# ------------------------------------------------------------------------------------------ workloads
def make_training_set(w):
    """Seeded training examples (host generator: identical on every rank and in both arms): a CellBatch, or
    CellSlots when an example has several cells.  y ~ N(0, 1), seed 0."""
    import torch
    from nasbowl_b200 import synth
    from nasbowl_b200.cells import CellSlots
    n = w["n_train"]
    slots = [synth.generate_cells_host(w["family"], seed=s, first_index=0, n_cells=n) for s in range(w["slots"])]
    y = torch.randn(n, generator=torch.Generator().manual_seed(0))
    return (slots[0] if len(slots) == 1 else CellSlots(slots)), y


def make_kernels(w):
    from nasbowl_b200.kernels import WeisfilerLehman
    if w["slots"] == 1:
        return [WeisfilerLehman(h=2)], None
    return [WeisfilerLehman(h=2, slot=s) for s in range(w["slots"])], [1.0 / w["slots"]] * w["slots"]


POOL_SEEDS = (1234, 4321, 777, 999)


def pool_first_index(pool_id, total, lo):
    return 10_000 + pool_id * total + lo


def generate_pool(eng, w, pool_id, total, lo, count):
    """Device records of candidates [lo, lo + count) of pool `pool_id` (a pool of `total` candidates); slot s of a
    candidate comes from the generator stream POOL_SEEDS[s].  Shards concatenate to the single-GPU pool."""
    import torch
    first = pool_first_index(pool_id, total, lo)
    parts = [eng.generate_cells(w["family"], POOL_SEEDS[s], first, count) for s in range(w["slots"])]
    return parts[0] if len(parts) == 1 else torch.cat(parts, dim=1).contiguous()


def records_to_batches(w, rec, n_max):
    """Host numpy records -> CellBatch per slot."""
    from nasbowl_b200.cells import CellBatch
    stride = rec.shape[1] // w["slots"]
    return [CellBatch.from_records(np.ascontiguousarray(rec[:, s * stride:(s + 1) * stride]), n_max) for s in range(w["slots"])]


def oracle_fit(w, train, y):
    """The CPU oracle's fit of the same surrogate (oracle/gp_oracle.py)."""
    from oracle import gp_oracle
    yy = y.numpy().astype(np.float64)
    if w["slots"] == 1:
        st = gp_oracle.gp_fit(train, yy, h_candidates=w["h_grid"], optimize_lik=True, max_lik=0.01)
        return st, gp_oracle.incumbent(st), st.h
    trains = list(train.slots)
    st = gp_oracle.gp_fit_sum_fixed(trains, yy, [(2, False)] * w["slots"], [1.0 / w["slots"]] * w["slots"])
    return st, gp_oracle.incumbent_sum(st), 2


def oracle_score(w, st, inc, train, sample_batches, chunk=2000):
    from oracle import gp_oracle
    if w["slots"] == 1:
        mu, var = gp_oracle.gp_predict_diag(st, train, sample_batches[0], chunk=chunk)
    else:
        mu, var = gp_oracle.gp_predict_diag_sum(st, list(train.slots), sample_batches, chunk=min(chunk, 1000))
    ei = gp_oracle.expected_improvement(mu, var, inc)
    return ei, gp_oracle.top_n_distinct(ei.astype(np.float32), TOP_N)


def oracle_score_per_candidate(w, st, train, sample, count):
    """The reference's ACTUAL propose_location path (acquisitions.py:94-113): one `eval` per candidate, each a
    predict on that candidate alone plus `_get_incumbent`'s predict on the whole training set (acquisitions.py:59-92).
    Restated with the oracle's (cheaper) train-by-pool features; single-kernel workloads only.  Returns seconds."""
    from oracle import gp_oracle
    if w["slots"] != 1:
        return None
    t0 = time.perf_counter()
    for i in range(count):
        mu, var = gp_oracle.gp_predict_diag(st, train, sample[i:i + 1])
        mu_t, _ = gp_oracle.gp_predict_diag(st, train, train)
        inc = float(st.y_raw[int(np.argmax(mu_t))])
        gp_oracle.expected_improvement(mu, var, inc)
    return time.perf_counter() - t0


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([i.get("num_threads", 1) for i in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm raises the BLAS pools back to the cores this
    process may run on (the arm is timed "with all the host threads it can use")."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=n)              # not a context manager here: the limit stays
    except Exception:
        pass
    return n


def fit_gp(w, gp, train, y):
    if w["slots"] == 1:
        gp.fit(wl_subtree_candidates=w["h_grid"], optimize_lik=True, max_lik=0.01)
    else:
        gp.fit(wl_subtree_candidates=(), optimize_lik=False)


def reset_gp(w, gp, train, y):
    gp.reset_XY(train, y)
    for k in gp.kernels:
        k.change_kernel_params({'h': 2})
    gp.likelihood = 1e-3


# ------------------------------------------------------------------------------------------ CPU arm
def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port — the Python reference tree
    does not travel to the GPU box) on the box's host cores, same config / metric / unit."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from nasbowl_b200 import synth
    from nasbowl_b200.cells import CellBatch
    w = WORKLOADS[args.workload]
    use_all_host_threads()
    train, y = make_training_set(w)
    S = args.ref_sample
    samples = []
    for s in range(args.steps + args.warmup):
        samples.append([synth.generate_cells_host(w["family"], POOL_SEEDS[k], 10_000 + s * S, S) for k in range(w["slots"])])
    st, inc, h_star = oracle_fit(w, train, y)
    for s in range(args.warmup):
        oracle_score(w, st, inc, train, samples[s])
    t0 = time.perf_counter()
    for s in range(args.steps):
        oracle_score(w, st, inc, train, samples[args.warmup + s])
    dt = time.perf_counter() - t0
    value = S * args.steps / dt
    threads = blas_threads()
    P = min(20, S)
    dt_pc = oracle_score_per_candidate(w, st, train, samples[0][0], P)
    world = max(1, args.gpus)
    total = pool_sizes(w, world, args.pool)
    per_gpu = total // world
    sample_desc = ("%d candidates per step (a bounded sample of the %d-candidate pool; the step time scales linearly "
                   "in the candidates), n_train=%d, h*=%d, CPU only" % (S, total, w["n_train"], h_star))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "candidates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": w["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(workload_config(w, per_gpu, total, h_star), sample_per_step=S),
        "cpu_baseline": {"value": value, "unit": "candidates/s", "cores": threads, "kind": "port", "sample": sample_desc,
                         # SURVEY 8(d)'s second CPU number: the loop the reference's propose_location really runs
                         "propose_location_path": None if dt_pc is None else {
                             "value": P / dt_pc, "unit": "candidates/s",
                             "sample": "%d candidates, one eval each: predict(candidate) + _get_incumbent's "
                                       "predict(train) (acquisitions.py:59-92), oracle port" % P}},
        "e2e": {"value": value, "unit": "candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ scoring configs
def instruction_profile(name):
    """Warp instructions per candidate and DRAM bytes per launch of the scoring kernel, from the committed ncu
    capture of the same config (profiles/r02_score_counters.json; a profiler cannot run inside the bench)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_score_counters.json")) as f:
            return json.load(f).get(name)
    except Exception:
        return None


def run_scoring(name, args, ctx):
    """One scoring workload on the initialised process group.  Returns the record (rank 0) or None."""
    import torch
    import torch.distributed as dist
    from nasbowl_b200 import parallel
    from nasbowl_b200.bayesopt import GraphExpectedImprovement, GraphGP
    rank, world, eng, sampler, peaks, peak_src = ctx["rank"], ctx["world"], ctx["eng"], ctx["sampler"], ctx["peaks"], ctx["peak_src"]
    w = WORKLOADS[name]
    steps, warmup = args.steps, args.warmup
    total = pool_sizes(w, world, args.pool)
    lo, hi = parallel.shard_range(total, rank, world)
    M = hi - lo
    train, y = make_training_set(w)
    n_max = train.n_max
    rec_bytes = (48 if n_max == 16 else 16) * w["slots"]
    kernels, weights = make_kernels(w)

    # ---- fit on rank 0, broadcast the pool model
    gp = GraphGP(train, y, kernels, weights=weights)
    fit_ms = state_ms = bcast_cold_ms = bcast_ms = None
    if rank == 0:
        for _ in range(2):                      # warm (first-call costs: module loading, allocator growth)
            reset_gp(w, gp, train, y)
            fit_gp(w, gp, train, y)
            gp._score_state()
        torch.cuda.synchronize()
        fits, states = [], []
        for _ in range(3):                      # wall clock, best of 3 (like bo_iter_ms)
            reset_gp(w, gp, train, y)
            t0 = time.perf_counter()
            fit_gp(w, gp, train, y)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            pm = gp._score_state()
            torch.cuda.synchronize()
            fits.append(1e3 * (t1 - t0))
            states.append(1e3 * (time.perf_counter() - t1))
        fit_ms, state_ms = min(fits), min(states)
    else:
        pm = None
    if world > 1:
        for attempt in range(2):                # cold (first use of the communicator for this size) and warm
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            got = parallel.broadcast_pool_model(pm if rank == 0 else None, src=0)
            torch.cuda.synchronize()
            dt = 1e3 * (time.perf_counter() - t0)
            if attempt == 0:
                bcast_cold_ms = dt
            else:
                bcast_ms = dt
        if rank != 0:
            pm = got
            gp.set_pool_model(pm)
    acq = GraphExpectedImprovement(gp)
    h_star = pm.h
    model = acq._model()
    fused = eng.fused_topk_ok(model, TOP_N)

    # ---- resident pools, generated on device; enough distinct ones that a step never finds its input in L2
    n_pools = int(min(16, max(1, math.ceil(2 * L2_BYTES / max(M * rec_bytes, 1)))))
    pools = [generate_pool(eng, w, p, total, lo, M) for p in range(n_pools)]
    torch.cuda.synchronize()
    score_buf = eng.empty((M,), torch.float32)
    cand_buf = eng.empty((TOP_N * 16,), torch.uint8)
    gather_buf = eng.empty((world * TOP_N * 16,), torch.uint8)

    # Steps are software-pipelined one deep (parallel.PoolPipeline): pool i+1 is scored while the top-5 exchange
    # of pool i (all-gather + merge + read-back) runs on a side stream; every step's selection reaches the host
    # inside the timed region.
    pipe = parallel.PoolPipeline(eng, TOP_N, True, depth=2)

    def run_steps(first, count, events=None):
        prev, out = None, None
        for i in range(first, first + count):
            acq.iters += 1
            if events is not None:
                eng.profile_events(events[i - first][0], events[i - first][1])
            t = pipe.submit(model, pools[i % n_pools], M, lo, scores=score_buf)
            if prev is not None:
                out = pipe.collect(prev)
            prev = t
        if events is not None:
            eng.profile_events(None, None)
        return pipe.collect(prev)

    def step(i):
        acq.iters += 1
        return parallel.score_and_select(eng, model, pools[i % n_pools], M, TOP_N, True, lo, scores=score_buf,
                                         cand_buf=cand_buf, gathered=gather_buf)

    first_vals, first_idx = step(0)
    pv, pi = run_steps(0, 1)
    assert np.array_equal(pi, first_idx) and np.array_equal(pv, first_vals), "pipelined and direct selection differ"
    run_steps(1, max(warmup - 1, 1))
    # kernel duration: events recorded by the library right around the scoring kernel launch, one pair per step
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in kev:
        a.record()
        b.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = eng.launches
    t_begin = time.time()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    run_steps(warmup, steps, kev)
    end.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    elapsed_ms = start.elapsed_time(end)
    launches = eng.launches - launches0
    t_end = time.time()
    kernel_ms = sum(a.elapsed_time(b) for a, b in kev) / steps

    # ---- end to end through the public API with HOST buffers (pinned): H2D pool, score, top-5, D2H scores
    host_pool = torch.empty((M, rec_bytes), dtype=torch.uint8).pin_memory()
    host_pool.copy_(pools[0].cpu())
    # Cells travel in the wire form (cells.CompactRecords, nbw_model.compact): 8 bytes per 8-node cell, 24 per 16-node
    # cell — half the PCIe bytes.  The 16 / 48-byte records of the same pool are timed as well and reported beside it.
    from nasbowl_b200.cells import CompactRecords
    parts = [b.to_compact() for b in records_to_batches(w, host_pool.numpy(), n_max)]
    data = torch.from_numpy(np.concatenate([r for r, _ in parts], axis=1)).pin_memory()
    wire_pool, wire_bytes = CompactRecords(data, [p for _, p in parts]), (8 if n_max == 8 else 24) * w["slots"]

    def time_e2e(pool_obj):
        def e2e_step():
            xs, eis, idx = acq.propose_location(pool_obj, top_n=TOP_N)     # top-5 indices + all EIs (host numpy) out
            return idx
        for _ in range(max(3, min(3 * warmup, 15) if M <= 2_000_000 else 3)):
            e2e_step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            e2e_step()
        torch.cuda.synchronize()
        return 1e3 * (time.perf_counter() - t0)
    e_steps = max(3, steps if M <= 2_000_000 else steps // 4)
    e2e_full_ms = time_e2e(host_pool)
    e2e_ms = time_e2e(wire_pool) if wire_pool is not host_pool else e2e_full_ms

    # ---- whole BO iterations (nasbench_optimisation.py:153-216): five new evaluated cells join the training set,
    # the surrogate is refitted (reset_XY + fit: appended to the previous fit where the path allows),
    # the pool operators are rebuilt, broadcast, and one pool is scored with its top-5.  Wall clock, best of 3.
    import copy
    bo = {}
    for mode in ("default", "incremental", "scratch"):
        big, y_big = make_training_set(dict(w, n_train=w["n_train"] + 20))
        n0 = w["n_train"]
        if rank == 0:
            # default: the product's policy (append from 768 training cells on); incremental: the append path forced
            # at this size; scratch: every fit from scratch
            gp.incremental = mode != "scratch"
            gp.incremental_min_n = 768 if mode == "default" else 0
            reset_gp(w, gp, big[:n0], y_big[:n0])
            fit_gp(w, gp, big[:n0], y_big[:n0])
            gp._score_state()
        ts = []
        for it in range(1, 4):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if rank == 0:
                nn = n0 + 5 * it
                gp.reset_XY(big[:nn], y_big[:nn])
                fit_gp(w, gp, big[:nn], y_big[:nn])
                pm2 = gp._score_state()
            else:
                pm2 = None
            if world > 1:
                pm2 = parallel.broadcast_pool_model(pm2, src=0)
                if rank != 0:
                    gp.set_pool_model(pm2)
            acq2 = GraphExpectedImprovement(gp)
            parallel.score_and_select(eng, acq2._model(), pools[0], M, TOP_N, True, lo, scores=score_buf,
                                      cand_buf=cand_buf, gathered=gather_buf)
            torch.cuda.synchronize()
            ts.append(1e3 * (time.perf_counter() - t0))
        bo[mode] = min(ts)
        if rank == 0 and mode == "incremental":
            bo["appended"] = bool(gp.last_fit_incremental)
        if rank == 0 and mode == "default":
            bo["default_appended"] = bool(gp.last_fit_incremental)
    bo_iter_ms, bo_scratch_ms, bo_append_ms = bo["default"], bo["scratch"], bo["incremental"]
    # back to the benchmark's surrogate for the parity checks below
    if rank == 0:
        gp.incremental = True
        gp.incremental_min_n = 768
        reset_gp(w, gp, train, y)
        gp._prev = None
        fit_gp(w, gp, train, y)
        pm2 = gp._score_state()
    else:
        pm2 = None
    if world > 1:
        pm2 = parallel.broadcast_pool_model(pm2, src=0)
        if rank != 0:
            gp.set_pool_model(pm2)
    model = GraphExpectedImprovement(gp)._model()
    acq = GraphExpectedImprovement(gp)

    # ---- parity (never null).  N > 1: the merged top-5 of step 0 must equal one GPU's top-5 of the whole pool.
    parity = {}
    if world > 1:
        ok = None
        if rank == 0:
            scores_all = []
            for r in range(world):
                rlo, rhi = parallel.shard_range(total, r, world)
                cells_r = pools[0] if r == 0 else generate_pool(eng, w, 0, total, rlo, rhi - rlo)
                scores_all.append(eng.score_pool(model, cells_r, rhi - rlo)[0].clone())
                del cells_r
            sv, si = eng.topk(torch.cat(scores_all), TOP_N, distinct=True)
            ok = bool(np.array_equal(si, first_idx) and np.array_equal(sv, first_vals))
            parity.update(merged_top5_equals_single_gpu=ok, merged_top5=[int(i) for i in first_idx],
                          single_gpu_top5=[int(i) for i in si])
            del scores_all
        everyone = [None] * world
        dist.all_gather_object(everyone, [int(i) for i in first_idx])
        if rank == 0:
            parity["ranks_agree_on_top5"] = bool(all(e == everyone[0] for e in everyone))
    # Oracle agreement on a bounded slice + CPU baseline (rank 0).  Runs AFTER the GPU measurements: its BLAS
    # worker threads keep spinning for a while and would slow the per-step host work of the timed loops.
    cpu_baseline = None
    if rank == 0 and not args.skip_cpu:
        S = min(args.cpu_sample or w["cpu_sample"], M)
        rec = pools[0][:S].cpu().numpy()
        sample = records_to_batches(w, rec, n_max)
        t0 = time.perf_counter()
        st, inc, h_o = oracle_fit(w, train, y)
        t1 = time.perf_counter()
        ei_ref, top_ref = oracle_score(w, st, inc, train, sample)
        t2 = time.perf_counter()
        sc, _, _, sc64 = eng.score_pool(model, pools[0][:S].contiguous(), S, want64=True)
        err = float(np.max(np.abs(sc64.cpu().numpy() - ei_ref) / (np.abs(ei_ref) + 1e-12)))
        top_gpu = eng.topk(sc, TOP_N, distinct=True)[1]
        top_o32 = set(int(i) for i in top_ref)
        parity.update(sample=S, max_rel_err_ei_fp64=err, h_star_equal=bool(h_o == h_star),
                      top5_equals_oracle=bool(set(int(i) for i in top_gpu) == top_o32), tolerance=1e-6,
                      ok=bool(err < 1e-6 and h_o == h_star))
        cpu_baseline = {"value": S / (t2 - t1), "unit": "candidates/s", "cores": blas_threads(), "kind": "port",
                        "fit_s": t1 - t0,
                        "sample": "first %d candidates of pool 0 (oracle/gp_oracle.py: joint WL + fp64 GP + EI + top-5)" % S}
    elif rank == 0:
        parity.update(sample=0, note="--skip-cpu: oracle pass skipped")

    # ---- max over ranks
    if world > 1:
        t = torch.tensor([elapsed_ms, e2e_ms, kernel_ms, bo_iter_ms, bo_scratch_ms, e2e_full_ms, bo_append_ms], dtype=torch.float64,
                         device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, e2e_ms, kernel_ms, bo_iter_ms, bo_scratch_ms, e2e_full_ms, bo_append_ms = [float(x) for x in t.cpu()]
        lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    del pools, host_pool, wire_pool
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    clocks = sampler.window(t_begin, t_end) if sampler is not None else None
    hbm_peak = float(peaks["hbm_gbs"])
    value = total * steps / (elapsed_ms / 1e3)
    hbm_achieved = w["bytes"] * M / (kernel_ms / 1e3) / 1e9
    prof = instruction_profile(name)
    sm_mhz = (clocks or {}).get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
    issue_peak = 148 * 4 * sm_mhz * 1e6 / 1e9                       # G warp-instructions / s: 4 schedulers per SM
    kname = (prof or {}).get("kernel") or ("score_cell_kernel<%d>" % len(kernels) if n_max == 8 and len(kernels) <= 3
                                           else "score_packed_kernel<%d,%d>" % (n_max, len(kernels)))
    roof = {"kernel": kname, "kernel_ms": kernel_ms,
            "kernel_share_of_step": kernel_ms / (elapsed_ms / steps), "candidates_per_launch": M,
            "hbm": {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak,
                    "bytes_per_unit": w["bytes"], "layout_bytes_per_unit": rec_bytes + 4, "peak_source": peak_src + " hbm_gbs, burst copy"}}
    if prof:
        inst = float(prof["warp_inst_per_candidate"])
        achieved = inst * M / (kernel_ms / 1e3) / 1e9
        roof.update(bound="issue", achieved=achieved, peak=issue_peak, unit="Gwarp-inst/s", frac=achieved / issue_peak,
                    warp_inst_per_candidate=inst, issue_frac=achieved / issue_peak,
                    traffic=prof.get("dram_bytes_per_candidate", 0) * M if prof.get("dram_bytes_per_candidate") else None,
                    counters_source="profiles/r02_score_counters.json (ncu smsp__inst_executed.sum, dram__bytes_*.sum "
                                    "of the same kernel and config; per candidate)",
                    peak_note="148 SMs x 4 schedulers x %.0f MHz sampled under load" % sm_mhz,
                    note="integer-issue bound (relabel, class refinement, probes): DRAM moves the records once")
    else:
        roof.update(bound="hbm", achieved=hbm_achieved, peak=hbm_peak, unit="GB/s", frac=hbm_achieved / hbm_peak,
                    traffic=None, note="no instruction profile committed for this config: HBM figures only; the "
                                       "kernel is issue-bound, not HBM-bound")
    rec_out = {
        "metric": METRIC, "value": value, "unit": "candidates/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": elapsed_ms / steps, "higher_is_better": True, "scaling": w["scaling"],
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(workload_config(w, M, total, h_star), **{
                   "n_labels": int(pm.total_labels or pm.n_labels),
                   "l2": "inputs larger than L2: %d rotating pools x %.0f MB per GPU" % (n_pools, M * rec_bytes / 1e6),
                   "fused_topk": bool(fused), "pipeline": "one pool in flight: exchange of pool i under the scoring of pool i+1",
                   "fit_ms": fit_ms, "score_state_ms": state_ms, "fit_timing": "wall clock, from-scratch fit, best of 3 after 2 warm-ups",
                   "broadcast_ms": bcast_ms,
                   "broadcast_cold_ms": bcast_cold_ms, "bo_iter_ms": bo_iter_ms, "bo_iter_from_scratch_ms": bo_scratch_ms,
                   "bo_iter_append_forced_ms": bo_append_ms, "bo_iter_default_appended": bo.get("default_appended"),
                   "bo_iter_appended": bo.get("appended"),
                   "bo_iter": "5 cells appended, reset_XY + fit + pool operators + broadcast + one pool scored + top-5 (wall clock, best of 3); "
                              "bo_iter_ms: the default policy (append path, i.e. rank-k refit, from 768 training cells on; refit below: "
                              "profiles/r02_incremental_bench.json), bo_iter_append_forced_ms: the append path at this size, "
                              "bo_iter_from_scratch_ms: incremental off",
                   "model_bytes": pm.nbytes()}),
        "roofline": roof,
        "cpu_baseline": cpu_baseline,
        "e2e": {"value": total * e_steps / (e2e_ms / 1e3), "unit": "candidates/s",
                "h2d_bytes_per_step": M * wire_bytes, "d2h_bytes_per_step": M * 4 + TOP_N * 16, "steps": e_steps,
                "numa_node_rank0": ctx.get("numa"),
                "api": "GraphExpectedImprovement.propose_location on pinned host records (%d-byte %s records per candidate)"
                       % (wire_bytes, "compact wire" if wire_bytes != rec_bytes else "full"),
                "full_records": {"value": total * e_steps / (e2e_full_ms / 1e3), "h2d_bytes_per_step": M * rec_bytes}},
        "gpu_launches": launches,
        "clocks": clocks,
        "parity": parity,
    }
    return rec_out


# ------------------------------------------------------------------------------------------ Gram stress (configs[4])
def run_gram_stress(args, ctx):
    """200k x 200k WL h=3 integer Gram, row blocks dealt round-robin to the ranks (every rank regenerates the N
    cells from the seed and fits the same dictionary: no collective on the data path)."""
    import torch
    import torch.distributed as dist
    rank, world, eng, sampler, peaks, peak_src = ctx["rank"], ctx["world"], ctx["eng"], ctx["sampler"], ctx["peaks"], ctx["peak_src"]
    g = GRAM_STRESS
    N, h, B = args.gram_cells or g["cells"], g["h"], g["block"]
    cells = eng.generate_cells(0, 2024, 0, N)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fit = eng.wl_fit(cells, 8, h, False)
    torch.cuda.synchronize()
    t_fit = time.perf_counter() - t0
    D = fit.k_end(h)
    res = eng.gram_stress(fit, D, B, rank, world)
    torch.cuda.synchronize()
    total_ms, checks = res["ms"], res["checks"]
    t_begin, t_end = res["t_begin"], res["t_end"]
    # oracle agreement on a corner the CPU finishes in seconds (exact integers)
    from nasbowl_b200.cells import CellBatch
    from oracle import wl_oracle
    S = 384
    sub = CellBatch.from_records(cells[:S].cpu().numpy(), 8)
    # the dictionary of the corner alone differs from the 200k dictionary, but the Gram does not depend on it
    K_o = wl_oracle.gram_raw(sub, h)
    K_g = eng.gram(fit.phi[:S], fit.phi[:S], 0, D).cpu().numpy()
    checks["corner_equals_oracle_bit_exact"] = bool(np.array_equal(K_o, K_g))
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        fl = torch.tensor([int(all(checks.values()))], device="cuda")
        dist.all_reduce(fl, op=dist.ReduceOp.MIN)
        checks["all_ranks_ok"] = bool(fl.item())
    res.pop("result", None)          # the materialised matrix (40 GB at N = 200k) is released here
    del fit, cells
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    tiles_factor = res["computed_fraction"]
    ops_dense = 2.0 * N * N * D
    ops_done = ops_dense * tiles_factor
    tens_peak = 2.0 * float(peaks["bf16_tflops"]) * world        # TOP/s; no int8 entry in MEASURED_PEAKS.json
    clocks = sampler.window(t_begin, t_end) if sampler is not None else None
    return {
        "metric": "Gram-only stress: int8 tensor-core throughput of the exact WL kernel matrix", "unit": "Tera int8 op/s",
        "value": ops_done / total_ms / 1e9, "n_gpus": world, "steps": 1, "warmup": 1, "ms_per_step": total_ms,
        "higher_is_better": True, "scaling": "strong", "dtype": "s8 x s8 -> s32", "data": "synthetic",
        "config": {"workload": g["desc"] + " (N=%d), row blocks of %d sharded over the GPUs" % (N, res.get("row_block", B)),
                   "dims": fit_dims(res), "feature_columns_padded": D, "phi_bytes": N * D, "wl_fit_ms": 1e3 * t_fit,
                   "schedule": res["schedule"], "output": res["output"],
                   "dense_equivalent_value": ops_dense / total_ms / 1e9},
        "roofline": {"bound": "tensor", "kernel": res["kernel"], "achieved": ops_done / total_ms / 1e9, "peak": tens_peak,
                     "unit": "TOP/s", "frac": ops_done / total_ms / 1e9 / tens_peak, "traffic": None,
                     "peak_source": peak_src + ": 2 x measured bf16 (no int8 peak recorded; nominal dense int8 is 4500 TOP/s per GPU)",
                     "output_bytes": res["output_bytes"], "output_gbs": res["output_bytes"] / total_ms / 1e6},
        "e2e": None, "cpu_baseline": None, "clocks": clocks,
        "parity": dict(checks, ok=bool(all(checks.values()))),
    }


def fit_dims(res):
    return res.get("dims")


# ------------------------------------------------------------------------------------------ main
def run_b200(args):
    import torch
    import torch.distributed as dist
    from nasbowl_b200 import parallel
    from nasbowl_b200.engine import Engine
    rank, world, local = parallel.init_distributed()
    torch.cuda.set_device(local)
    # several ranks on one host: keep every rank, and the pinned pools it allocates, on its GPU's NUMA node
    numa = parallel.bind_to_gpu_numa_node(local) if world > 1 and not os.environ.get("NBW_NO_NUMA_BIND") else None
    eng = Engine.get(local)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler is not None:
        sampler.start()
    peaks, peak_src = load_peaks()
    ctx = dict(rank=rank, world=world, eng=eng, sampler=sampler, peaks=peaks, peak_src=peak_src, numa=numa)
    line = run_scoring(args.workload, args, ctx)
    subs = {}
    for name in [c for c in args.configs.split(",") if c]:
        if name == args.workload:
            continue
        try:
            if name == "cfg5":
                sub = run_gram_stress(args, ctx)
            else:
                sub = run_scoring(name, args, ctx)
        except Exception as e:            # a failing sub-config must not take the headline with it
            import traceback
            sub = {"error": repr(e), "traceback": traceback.format_exc()[-1500:]} if rank == 0 else None
            if world > 1:
                raise
        if rank == 0:
            subs[name] = sub
    if sampler is not None:
        sampler.stop()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    line["configs"] = subs
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS), help="the headline config")
    ap.add_argument("--configs", default="cfg3,cfg4,cfg5", help="further configs reported as sub-records")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pool", type=int, default=None, help="override the pool size (per GPU for cfg2, total otherwise)")
    ap.add_argument("--gram-cells", type=int, default=None)
    ap.add_argument("--cpu-sample", type=int, default=None)
    ap.add_argument("--ref-sample", type=int, default=1000)
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    # the clock sampler is a Python thread: keep its hand-over latency (default 5 ms) out of the wall-clock figures
    sys.setswitchinterval(1e-4)
    if args.pool is not None:
        args.configs = ""           # a pool override addresses one workload
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
