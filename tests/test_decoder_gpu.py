"""GPU: decoder MLP on the tensor cores (tcgen05 / TMEM) -- conventions self-test and parity of
the fused decoder forward/backward against the torch restatement of network.ShallowMLP
(oracle/torch_ref.py, pinned by tests/golden/py_golden_mlp.npz)."""
import numpy as np
import pytest
import torch

from conftest import load_pkg

pytestmark = pytest.mark.gpu


def _bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


def test_umma_selftest_three_gemm_shapes():
    load_pkg()
    import scanerf_b200_capi as capi
    from scanerf_b200_capi import ptr
    g = torch.Generator().manual_seed(0)
    X, W, G = torch.randn(128, 64, generator=g), torch.randn(64, 64, generator=g), torch.randn(128, 64, generator=g)
    dev = "cuda:0"
    Xd, Wd, Gd = X.to(dev), W.to(dev), G.to(dev)
    Y, DX, DW = torch.zeros(128, 64, device=dev), torch.zeros(128, 64, device=dev), torch.zeros(64, 64, device=dev)
    DWo, YS = torch.zeros(64, 16, device=dev), torch.zeros(128, 16, device=dev)
    capi.check(capi.lib().snrf_umma_selftest(ptr(Xd), ptr(Wd), ptr(Gd), ptr(Y), ptr(DX), ptr(DW), ptr(DWo), ptr(YS),
                                             capi.stream()), "snrf_umma_selftest")
    torch.cuda.synchronize()
    Xb, Wb, Gb = _bf16(X).double(), _bf16(W).double(), _bf16(G).double()
    for name, got, want in (("Y", Y, Xb @ Wb.T), ("DX", DX, Gb @ Wb), ("DW", DW, 2 * Gb.T @ Xb),
                            ("DWo", DWo, Gb.T @ Xb[:, 32:48]), ("YS", YS, Xb[:, 32:64] @ Wb[0:16, 0:32].T)):
        err = float((got.cpu().double() - want).abs().max())
        assert err < 1e-3, f"{name}: max abs err {err}"


def _decoder_and_inputs(N, S, seed):
    from hashgrid._decoder import ShallowMLP, decoder_params
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    dec = ShallowMLP(32)
    for p in dec.parameters():                      # non-zero biases so that every term is exercised
        if p.dim() == 1:
            p.data = torch.randn(p.shape, generator=g) * 0.05
    feats = torch.randn(N, 32, generator=g) * 0.3
    mask = torch.rand(32, generator=g)
    rays_d = torch.randn((N + S - 1) // S, 3, generator=g) * 1.7
    return dec, decoder_params(dec), feats, mask, rays_d


@pytest.mark.parametrize("split", [True, False])
@pytest.mark.parametrize("N,S", [(128, 128), (1000, 8), (128 * 37 + 5, 64)])
def test_decoder_forward_matches_torch(N, S, split):
    load_pkg()
    from hashgrid import _field
    dec, params, feats, mask, rays_d = _decoder_and_inputs(N, S, N)
    dirs = rays_d.repeat_interleave(S, 0)[:N]
    with torch.no_grad():
        ref = dec(torch.cat([feats, dirs], -1), weight_feature=mask)
    dev = "cuda:0"
    _field.set_precision(split)
    try:
        out = _field.decoder_forward(feats.to(dev), mask.to(dev), rays_d.to(dev), S, [p.to(dev) for p in params]).cpu()
    finally:
        _field.set_precision(True)
    got = {"sigma": out[:, 0:1], "tint": out[:, 1:4], "diffuse": out[:, 4:7], "specular": out[:, 7:10]}
    # split (bf16x3) operands: fp32-grade, bar 1e-4 absolute on every head (so composited RGB stays within 1e-4);
    # plain bf16 operands: the Gaussian activations amplify the 2^-9 operand rounding to ~1 %
    max_tol, rel_tol = (1e-4, 5e-5) if split else (4e-2, 2e-2)
    for k in got:
        err = float((got[k] - ref[k]).abs().max())
        mean_rel = float((got[k] - ref[k]).abs().mean() / ref[k].abs().mean())
        print(f"decoder fwd split={split} {k}: max abs err {err:.2e}, mean rel err {mean_rel:.2e}")
        assert err < max_tol, f"{k}: max abs err {err}"
        assert mean_rel < rel_tol, f"{k}: mean relative err {mean_rel}"


@pytest.mark.parametrize("split", [True, False])
@pytest.mark.parametrize("N,S", [(128, 128), (128 * 5 + 17, 32), (4096 * 3, 128)])
def test_decoder_backward_matches_torch(N, S, split):
    load_pkg()
    from hashgrid import _field
    dec, params, feats, mask, rays_d = _decoder_and_inputs(N, S, N + 1)
    g = torch.Generator().manual_seed(3)
    cot = torch.randn(N, 10, generator=g)
    # torch reference (fp32, CPU)
    f_ref = feats.clone().requires_grad_(True)
    d_ref = rays_d.clone().requires_grad_(True)
    o = dec(torch.cat([f_ref, d_ref.repeat_interleave(S, 0)[:N]], -1), weight_feature=mask)
    ref_heads = torch.cat([o["sigma"], o["tint"], o["diffuse"], o["specular"]], -1)
    (ref_heads * cot).sum().backward()
    # tensor-core path
    dev = "cuda:0"
    f_gpu = feats.to(dev).requires_grad_(True)
    d_gpu = rays_d.to(dev).requires_grad_(True)
    p_gpu = [p.detach().to(dev).requires_grad_(True) for p in params]
    _field.set_precision(split)
    try:
        heads = _field.decoder_apply(f_gpu, d_gpu, mask.to(dev), S, p_gpu)
        (heads * cot.to(dev)).sum().backward()
        torch.cuda.synchronize()
    finally:
        _field.set_precision(True)

    def rel(a, b):      # relative error in the L2 sense
        return float((a.cpu() - b).norm() / b.norm().clamp_min(1e-20))
    errs = {"feats": rel(f_gpu.grad, f_ref.grad), "rays_d": rel(d_gpu.grad, d_ref.grad)}
    for q, p in zip(p_gpu, params):
        errs[f"param{tuple(p.shape)}"] = rel(q.grad, p.grad)
    print(f"decoder bwd split={split} rel L2 errs:", {k: f"{v:.2e}" for k, v in errs.items()})
    # split: the input-gradient chain is error-compensated (bar 2e-3, the north star's bf16-path tolerance;
    # measured ~4e-4, limited by the fp16 activation derivatives); weight gradients keep bf16 activations
    # (2^-9 operand rounding, independent per sample: bar 5e-3 on this incoherent random cotangent).
    # plain bf16: the Gaussian activations amplify operand rounding to ~1e-1.
    for k, v in errs.items():
        tol = (2e-3 if k in ("feats", "rays_d") else 5e-3) if split else 2e-1
        assert v < tol, f"{k}: {v}"
