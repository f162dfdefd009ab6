#!/usr/bin/env python
"""Per-kernel roofline figures for the training-side kernels: relabel,
dictionary build (radix sort/unique), histogram, integer Gram.  CUDA events on the launch
stream, L2 flushed between timed launches, medians.  Prints one JSON object.

    python profiles/kernel_bench.py [--cells 2000000] [--gram-n 16384] [--gram-k 8192]
"""
import argparse
import ctypes as C
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def timed(fn, flush, reps=7):
    ts = []
    for _ in range(reps):
        flush.add_(1)                                   # 512 MB write: evicts L2
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=2_000_000)
    ap.add_argument("--family", type=int, default=0)
    ap.add_argument("--h", type=int, default=2)
    ap.add_argument("--gram-n", type=int, default=16384)
    ap.add_argument("--gram-k", type=int, default=8192)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    from nasbowl_b200.engine import Engine
    eng = Engine.get()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
    hbm = peaks["hbm_gbs"]
    i8_peak = 2.0 * peaks["bf16_tflops"]                # no int8 peak is recorded: 2x the bf16 figure
    flush = torch.zeros(128 * 1024 * 1024, dtype=torch.float32, device="cuda")
    out = {"hbm_peak_gbs": hbm, "int8_peak_tops_assumed_2x_bf16": i8_peak, "kernels": {}}
    lib, ctx, st = eng.lib, eng.ctx, eng.stream
    # context for write-dominated kernels: a pure store stream (cudaMemset of 2 GiB) on this GPU
    big = torch.empty(2 * 1024 ** 3, dtype=torch.uint8, device="cuda")
    ms = timed(lambda: big.zero_(), flush, reps=5)
    out["write_only_stream_gbs (memset 2 GiB)"] = big.numel() / ms / 1e6
    del big

    if not args.only or "wl" in args.only:
        N, n_max, h = args.cells, (16 if args.family == 2 else 8), args.h
        stride = 48 if n_max == 16 else 16
        cells = eng.generate_cells(args.family, 42, 0, N)
        nk = N * n_max
        keys = eng.empty((nk,), torch.int64)
        chks = eng.empty((nk,), torch.int64)
        # ---- relabel, level 1 (reads records + level-0 ids, writes key + chk per node)
        fit = eng.wl_fit(cells, n_max, h, False)
        ids0 = fit.ids[0]
        ms = timed(lambda: eng._check(lib.nbw_wl_signatures(ctx, eng._p(cells), N, n_max, eng._p(ids0), eng._p(fit.uniq_keys[0]), 1,
                                                            C.c_uint64(7), eng._p(keys), eng._p(chks), st)), flush)
        byt = N * stride + nk * 4 + nk * 16
        out["kernels"]["wl_signature_kernel"] = {"ms": ms, "cells": N, "algorithmic_bytes": byt,
                                                 "achieved_gbs": byt / ms / 1e6, "frac_hbm": byt / ms / 1e6 / hbm}
        # ---- dictionary build of level 1 (8 radix passes over 12-byte (key, index) pairs + unique)
        ids = eng.empty((nk,), torch.int32)
        uk = eng.empty((nk,), torch.int64)
        uc = eng.empty((nk,), torch.int64)
        d = C.c_int64(0)
        ms = timed(lambda: eng._check(lib.nbw_dict_build(ctx, eng._p(keys), eng._p(chks), nk, 64, eng._p(cells), n_max,
                                                         eng._p(ids0), eng._p(ids), eng._p(uk), eng._p(uc), nk,
                                                         C.byref(d), st)), flush)
        # hash pre-dedup path (default above 8192 keys): key + chk read, slot written and read, id written, credential
        # check of every node against the first of its class (record + previous ids of both)
        byt = nk * (8 + 4 + 4 + 8 + 4 + 4) + N * stride
        out["kernels"]["dict_build (hash pre-dedup + sort of distinct keys)"] = {
            "ms": ms, "keys": nk, "unique": int(d.value), "algorithmic_bytes": byt,
            "achieved_gbs": byt / ms / 1e6, "frac_hbm": byt / ms / 1e6 / hbm, "mkeys_per_s": nk / ms / 1e3}
        ids_hash = ids.clone()
        os.environ["NBW_DICT_RADIX"] = "1"
        ms = timed(lambda: eng._check(lib.nbw_dict_build(ctx, eng._p(keys), eng._p(chks), nk, 64, eng._p(cells), n_max,
                                                         eng._p(ids0), eng._p(ids), eng._p(uk), eng._p(uc), nk,
                                                         C.byref(d), st)), flush)
        del os.environ["NBW_DICT_RADIX"]
        assert torch.equal(ids, ids_hash)
        byt = 8 * nk * (8 + 12 + 12) + nk * (12 + 8 + 4 + 4 + 4)
        out["kernels"]["dict_build (radix sort + unique, NBW_DICT_RADIX=1)"] = {
            "ms": ms, "keys": nk, "unique": int(d.value), "algorithmic_bytes": byt,
            "achieved_gbs": byt / ms / 1e6, "frac_hbm": byt / ms / 1e6 / hbm, "mkeys_per_s": nk / ms / 1e3}
        # ---- histogram (reads ids of all levels, writes every byte of Phi once)
        ld = fit.col_off[-1]
        ms = timed(lambda: eng.build_features(fit), flush)
        byt = (h + 1) * nk * 4 + N * ld + N * 4
        out["kernels"]["hist_dense_kernel"] = {"ms": ms, "cells": N, "ld_phi": ld, "dims": fit.dims,
                                               "algorithmic_bytes": byt, "achieved_gbs": byt / ms / 1e6,
                                               "frac_hbm": byt / ms / 1e6 / hbm}
        del fit, cells, keys, chks, ids, uk, uc
        torch.cuda.empty_cache()

    if not args.only or "gram" in args.only:
        n, k = args.gram_n, args.gram_k
        g = torch.Generator(device="cuda").manual_seed(0)
        A = torch.randint(-3, 4, (n, k), dtype=torch.int8, device="cuda", generator=g)
        Cout = eng.empty((n, n), torch.int32)
        ms = timed(lambda: eng.gram(A, A, 0, k, out=Cout), flush, reps=5)
        ops = 2.0 * n * n * k
        out["kernels"]["gram_i8_pair_kernel"] = {"ms": ms, "m": n, "n": n, "k": k, "int_ops": ops,
                                                    "achieved_tops": ops / ms / 1e9,
                                                    "frac_of_assumed_int8_peak": ops / ms / 1e9 / i8_peak,
                                                    "frac_of_bf16_peak": ops / ms / 1e9 / peaks["bf16_tflops"]}
        # spot check against the SIMT kernel on a corner
        ref = eng.gram(A[:256], A[:256], 0, k, impl=1)
        assert torch.equal(ref, Cout[:256, :256])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
