import sys, torch, numpy as np
sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo")
import test_field_encode_gpu as T
from conftest import load_pkg
load_pkg()
import scanerf_b200_capi as capi
from scanerf_b200_capi import ptr, c_int, c_void_p
from hashgrid.PyHashGridBG import HashEmbeddingBG
from oracle import torch_ref as tr
DEV = "cuda:0"
R, S, log2T = 512, 64, 19
table, res, bmin, bsize, o, d, z_fg, z_bg, g = T._case(R, S, log2T, R + S)
for mode, z in ((1, z_fg), (2, z_bg)):
    N = R * S
    od, dd, zd = o.to(DEV), d.to(DEV), z.to(DEV)
    x = (od[:, None] + zd[..., None] * dd[:, None]).reshape(-1, 3)
    cx = (tr.contract_fore if mode == 1 else tr.contract_bg)(x, bmin.to(DEV), bsize.to(DEV)).contiguous()
    ref = HashEmbeddingBG(cx, table.to(DEV), res.to(DEV))
    L, Tn = 16, 2 ** log2T
    out0 = torch.empty(L, N, 2, device=DEV)
    capi.check(capi.lib().snrf_field_encode_fwd(c_void_p(0), c_void_p(0), c_void_p(0), ptr(cx), c_void_p(0), c_void_p(0), c_int(0), ptr(table.to(DEV)),
               ptr(res.to(DEV)), ptr(out0), c_void_p(0), c_int(N), c_int(S), c_int(L), c_int(Tn), capi.stream()), "f")
    out1 = torch.empty(L, N, 2, device=DEV)
    bm, bs = bmin.to(DEV), bsize.to(DEV)
    capi.check(capi.lib().snrf_field_encode_fwd(ptr(od), ptr(dd), ptr(zd), c_void_p(0), ptr(bm), ptr(bs), c_int(mode), ptr(table.to(DEV)),
               ptr(res.to(DEV)), ptr(out1), c_void_p(0), c_int(N), c_int(S), c_int(L), c_int(Tn), capi.stream()), "f")
    torch.cuda.synchronize()
    a, b = out0.permute(1, 0, 2), out1.permute(1, 0, 2)
    print("mode", mode, "points-mode vs op:", float((a - ref).abs().max()), "fused vs op:", float((b - ref).abs().max()),
          "n diff samples", int(((b - ref).abs().amax((1, 2)) > 0).sum()))
    bad = ((b - ref).abs().amax((1, 2)) > 0).nonzero()[:3, 0]
    for i in bad.tolist():
        r = i // S
        print("  sample", i, "x", x[i].tolist(), "cx", cx[i].tolist(), "levels differing", ((b[i] - ref[i]).abs().amax(-1) > 0).nonzero()[:, 0].tolist())
