#!/usr/bin/env python
"""Micro-benchmarks of individual scanerf_b200 kernels on one GPU (CUDA-event timed,
L2 flushed between iterations).  Development tool: bench.py is the contract bench.

  python tools/microbench.py encode [--log2T 24] [--rays 16384] [--samples 256] [--ref]
"""
import argparse
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "scanerf-scalable-bundle-adjusting-neural-radiance-fields-for-large-scale-scene-rendering_b200"


def ray_points(n_rays, n_samples, gen, dev):
    """Contracted sample positions of synthetic rays: first half of every ray's
    samples in the foreground cube [-1,1]^3, second half contracted background."""
    o = (torch.rand(n_rays, 3, generator=gen, device=dev) - 0.5)
    d = torch.nn.functional.normalize(torch.randn(n_rays, 3, generator=gen, device=dev), dim=-1)
    half = n_samples // 2
    z_f = torch.linspace(0.0, 1.0, half, device=dev)[None, :] * 0.8
    t = torch.linspace(0.0, 1.0, n_samples - half, device=dev)[None, :]
    z_b = 1.0 / ((1 - t) / 0.9 + t / 1e3)
    z = torch.cat([z_f.expand(n_rays, -1), z_b.expand(n_rays, -1)], 1)
    x = o[:, None, :] + z[..., None] * d[:, None, :]
    n = x.abs().amax(-1, keepdim=True)
    xc = torch.where(n > 1, x * ((2 - 1 / n) / n), x)
    return xc.reshape(-1, 3).contiguous()


def timeit(fn, iters, flush):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.add_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def bench_encode(a):
    pkg = importlib.import_module(PKG); pkg.install()
    from hashgrid.lib import HASHGRID as ops
    dev = "cuda:0"
    gen = torch.Generator(device=dev).manual_seed(0)
    L, T = 16, 2 ** a.log2T
    table = torch.randn(L, T, 2, device=dev, generator=gen) * 0.1
    base = torch.tensor([49., 32., 73.]); fin = torch.tensor([12603., 8192., 18904.])
    if a.log2T <= 19:
        base = torch.tensor([16., 16., 16.]); fin = torch.tensor([512., 512., 512.])
    b = torch.exp((torch.log(fin) - torch.log(base)) / (L - 1))
    res = torch.stack([(base * b ** i).int() for i in range(L)], 0).to(dev)
    pts = ray_points(a.rays, a.samples, gen, dev)
    if a.uniform:
        pts = torch.rand(pts.shape, device=dev, generator=gen) * 4 - 2
    B = pts.shape[0]
    out = torch.zeros(B, L, 2, device=dev)
    gin = torch.randn(B, L, 2, device=dev, generator=gen)
    gp = torch.zeros(B, 3, device=dev)
    gt = torch.zeros_like(table)
    flush = torch.zeros(64 * 1024 * 1024, device=dev)  # 256 MB > L2
    r = {"B": B, "log2T": a.log2T, "uniform": a.uniform}
    import scanerf_b200_capi as capi
    import ctypes
    if a.sweep:
        for lpb in (1, 2, 4, 16):
            capi.lib().snrf_hash_set_levels_per_block(ctypes.c_int(lpb))
            t = timeit(lambda: ops.embedding_bg_forward_cuda(pts, out, table, res), a.iters, flush)
            r[f"lpb{lpb}_fwd_ms"] = round(t, 3)
            t = timeit(lambda: ops._encode_fwd(pts, obf0, table, None, None, res), a.iters, flush) if False else 0
            for agg in (0, 8):
                t = timeit(lambda: ops._encode_bwd(pts, gin, gp, gt, table, None, None, res, aggregate_levels=agg), a.iters, flush)
                r[f"lpb{lpb}_bwd_agg{agg}_ms"] = round(t, 3)
            t = timeit(lambda: ops._encode_bwd(pts, gin, None, gt, table, None, None, res, aggregate_levels=8), a.iters, flush)
            r[f"lpb{lpb}_bwd_nodx_ms"] = round(t, 3)
        capi.lib().snrf_hash_set_levels_per_block(ctypes.c_int(0))
        print(json.dumps(r))
        return r
    t = timeit(lambda: ops.embedding_bg_forward_cuda(pts, out, table, res), a.iters, flush)
    r["fwd_ms"] = t; r["fwd_GBs_alg"] = B * 1164 / t / 1e6
    obf = torch.zeros(B, 2 * L, device=dev, dtype=torch.bfloat16)
    t = timeit(lambda: ops._encode_fwd(pts, obf, table, None, None, res), a.iters, flush)
    r["fwd_bf16_ms"] = t
    for agg in (0, 4, 8, 16):
        t = timeit(lambda: ops._encode_bwd(pts, gin, gp, gt, table, None, None, res, aggregate_levels=agg), a.iters, flush)
        r[f"bwd_agg{agg}_ms"] = t; r[f"bwd_agg{agg}_GBs_alg"] = B * 2200 / t / 1e6
    t = timeit(lambda: ops._encode_bwd(pts, gin, None, gt, table, None, None, res, aggregate_levels=8), a.iters, flush)
    r["bwd_nodx_agg8_ms"] = t
    t = timeit(lambda: gt.zero_(), a.iters, flush)
    r["zero_table_ms"] = t
    if a.ref:
        import oracle
        ref = oracle.ref_module("HASHGRID_EMBED")
        t = timeit(lambda: ref.embedding_bg_forward_cuda(pts, out, table, res), a.iters, flush)
        r["ref_fwd_ms"] = t
        t = timeit(lambda: ref.embedding_bg_backward_cuda(pts, gin, gp, gt, table, res), a.iters, flush)
        r["ref_bwd_ms"] = t
    print(json.dumps(r))
    return r


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["encode"])
    ap.add_argument("--log2T", type=int, default=24)
    ap.add_argument("--rays", type=int, default=16384)
    ap.add_argument("--samples", type=int, default=256)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--uniform", action="store_true")
    ap.add_argument("--ref", action="store_true")
    ap.add_argument("--sweep", action="store_true")
    a = ap.parse_args()
    {"encode": bench_encode}[a.what](a)
