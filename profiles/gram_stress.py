#!/usr/bin/env python
"""BASELINE.json configs[4]: Gram-only stress — WL h=3 features of N NB101-shaped cells and the
N x N integer Gram, computed in row blocks (the int32 result of 200k x 200k would be 160 GB).
Checks size-independent properties on the fly (diagonal == self-similarity, symmetry of mirrored
blocks, SIMT cross-check of a corner) and prints one JSON object with per-stage times.

    python profiles/gram_stress.py [--cells 200000] [--h 3] [--block 16384]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node R --master-addr 127.0.0.1 profiles/gram_stress.py ...

Under torchrun (Gram-only stress): every rank regenerates all N cells from the seed and
fits the same dictionary (no collective on the data path); rank r computes the row blocks r, r + R, ...;
the reported time is the maximum over ranks, the checks are AND-ed.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=200_000)
    ap.add_argument("--h", type=int, default=3)
    ap.add_argument("--block", type=int, default=16384)
    args = ap.parse_args()
    rank, world = 0, 1
    if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        from nasbowl_b200 import parallel
        rank, world, _ = parallel.init_distributed()
    from nasbowl_b200.engine import Engine
    eng = Engine.get()
    N, h, B = args.cells, args.h, args.block
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    cells = eng.generate_cells(0, 2024, 0, N)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fit = eng.wl_fit(cells, 8, h, False)
    torch.cuda.synchronize()
    t_fit = time.perf_counter() - t0
    D = fit.k_end(h)
    out = {"cells": N, "h": h, "dims": fit.dims, "feature_columns_padded": D, "phi_bytes": N * fit.ld_phi,
           "wl_fit_s (relabel + sort/unique + histogram, all levels)": t_fit}
    buf = eng.empty((B, (N + 3) // 4 * 4), torch.int32)
    self_sim = fit.self_sim
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    total_ms, ok_diag, ok_sym = 0.0, True, True
    first_block = None
    row_blocks = list(range(0, N, B))
    mine = [r0 for i, r0 in enumerate(row_blocks) if i % world == rank]
    if rank != 0:      # the mirror check needs the first row block on every rank
        first_block = eng.gram(fit.phi[:min(B, N)], fit.phi, 0, D, out=buf[:min(B, N)])
        first_block = first_block.clone() if N <= 4 * B else first_block[:1024, :].clone()
    for r0 in mine:
        rows = min(B, N - r0)
        ev[0].record()
        K = eng.gram(fit.phi[r0:r0 + rows], fit.phi, 0, D, out=buf[:rows])
        ev[1].record()
        torch.cuda.synchronize()
        total_ms += ev[0].elapsed_time(ev[1])
        d = K[torch.arange(rows, device="cuda"), torch.arange(r0, r0 + rows, device="cuda")]
        ok_diag &= bool(torch.equal(d, self_sim[r0:r0 + rows]))
        if r0 == 0:
            first_block = K[:, :].clone() if N <= 4 * B else K[:1024, :].clone()
        else:
            # K[r0:r0+c, 0:c'] must mirror the first block's K[0:c', r0:r0+c]
            c = min(rows, first_block.shape[0])
            ok_sym &= bool(torch.equal(K[:rows, :c].t().contiguous(), first_block[:c, r0:r0 + rows]))
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        flags = torch.tensor([int(ok_diag), int(ok_sym)], device="cuda")
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        total_ms, ok_diag, ok_sym = float(t.item()), bool(flags[0].item()), bool(flags[1].item())
        out["n_gpus"] = world
    ops = 2.0 * N * N * D
    out.update({"gram_ms": total_ms, "int_ops": ops, "achieved_tops": ops / total_ms / 1e9,
                "frac_of_2x_bf16_peak": ops / total_ms / 1e9 / (2 * peaks["bf16_tflops"] * world),
                "operand_bytes_read_once": N * D, "diag_equals_self_similarity": ok_diag,
                "mirrored_blocks_equal": ok_sym})
    ref = eng.gram(fit.phi[:128], fit.phi[:512], 0, D, impl=1)
    out["simt_crosscheck_corner"] = bool(torch.equal(ref, eng.gram(fit.phi[:128], fit.phi[:512], 0, D)))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
