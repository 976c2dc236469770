"""GPU parity of the fused compositing kernels (snrf_composite_fwd / _bwd) against the torch
restatement of HashGrid.cal_integrate_weight / accumulate (oracle/torch_ref.py, pinned by
tests/golden/py_golden_render_*.npz).  Bar (north_star): composited RGB within 1e-4."""
import numpy as np
import pytest
import torch

from conftest import load_pkg
from oracle import torch_ref as tr

pytestmark = pytest.mark.gpu


def _inputs(R, S, seed, big_sigma=False):
    g = torch.Generator().manual_seed(seed)
    heads = {"sigma": torch.rand(R, S, 1, generator=g) * (30.0 if big_sigma else 3.0),
             "tint": torch.rand(R, S, 3, generator=g), "diffuse": torch.rand(R, S, 3, generator=g),
             "specular": torch.rand(R, S, 3, generator=g)}
    z = torch.cumsum(torch.rand(R, S, generator=g) * 0.1 + 0.01, -1)
    dists = torch.cat([z[:, 1:] - z[:, :-1], torch.full((R, 1), 1e-6)], -1)
    d = torch.randn(R, 3, generator=g) * 1.3
    return heads, z, dists, d


@pytest.mark.parametrize("R,S,infinity", [(1, 1, False), (7, 31, True), (64, 128, False), (64, 128, True), (33, 200, True)])
def test_composite_forward_backward(R, S, infinity):
    load_pkg()
    from hashgrid import _render
    heads, z, dists, d = _inputs(R, S, R * 1000 + S)
    # oracle (CPU, torch autograd)
    hc = {k: v.clone().requires_grad_(True) for k, v in heads.items()}
    dc = d.clone().requires_grad_(True)
    ref = tr.composite(hc, z, dists, dc, infinity, train=True)
    wts = torch.Generator().manual_seed(5)
    cot = {k: torch.randn(ref[k].shape, generator=wts) for k in ("rgb", "depth", "T_left", "tint")}
    loss_ref = sum((ref[k] * cot[k]).sum() for k in cot) + 0.3 * ref["l2_reg_specular"]
    loss_ref.backward()
    # kernels
    dev = "cuda:0"
    hg = {k: v.to(dev).requires_grad_(True) for k, v in heads.items()}
    dg = d.to(dev).requires_grad_(True)
    out = _render.composite(hg, z.to(dev), dists.to(dev), dg, infinity, train=True)
    loss = sum((out[k] * cot[k].to(dev)).sum() for k in cot) + 0.3 * out["l2_reg_specular"]
    loss.backward()
    for k in ("rgb", "depth", "T_left", "tint", "diffuse", "specular"):
        assert torch.allclose(out[k].cpu(), ref[k], atol=1e-4, rtol=1e-4), k
    assert torch.allclose(out["weights"].cpu(), ref["weights"], atol=1e-5, rtol=1e-4)
    assert abs(float(out["l2_reg_specular"]) - float(ref["l2_reg_specular"])) < 1e-4
    for k in heads:
        a, b = hg[k].grad.cpu(), hc[k].grad
        scale = max(float(b.abs().max()), 1e-6)
        # 1 - exp(-x) cancels catastrophically in fp32 for x ~ 1e-6 (the reference's own formula): absolute floor
        assert float((a - b).abs().max()) < 2e-4 * scale + 3e-7, f"grad {k}: {float((a - b).abs().max())} vs scale {scale}"
    scale = max(float(dc.grad.abs().max()), 1e-6)
    assert float((dg.grad.cpu() - dc.grad).abs().max()) < 2e-4 * scale + 3e-7


def test_composite_saturating_density_is_finite():
    load_pkg()
    from hashgrid import _render
    heads, z, dists, d = _inputs(16, 128, 3, big_sigma=True)
    dev = "cuda:0"
    hg = {k: v.to(dev).requires_grad_(True) for k, v in heads.items()}
    out = _render.composite(hg, z.to(dev), dists.to(dev), d.to(dev), True, train=True)
    (out["rgb"].sum() + out["l2_reg_specular"]).backward()
    ref = tr.composite(heads, z, dists, d, True, train=True)
    assert torch.allclose(out["rgb"].cpu(), ref["rgb"], atol=1e-4)
    for k in heads:
        assert torch.isfinite(hg[k].grad).all()


@pytest.mark.parametrize("R,S", [(64, 128), (33, 200), (7, 31)])
def test_packed_rows_staged_through_shared_memory_equal_strided_access(R, S):
    """Packed [R*S,10] head rows: the backward always stages a warp's 32 rows through shared memory (coalesced 8-byte copies),
    the forward can (snrf_composite_set_fwd_packed); same values as the four separate tensors through the strided kernels."""
    import ctypes
    load_pkg()
    import scanerf_b200_capi as capi
    from hashgrid import _render
    heads, z, dists, d = _inputs(R, S, 7 * R + S, big_sigma=True)
    dev = "cuda:0"
    packed = torch.cat([heads["sigma"], heads["tint"], heads["diffuse"], heads["specular"]], -1).reshape(R * S, 10)
    hs = {k: v.to(dev).requires_grad_(True) for k, v in heads.items()}
    ref = _render.composite(hs, z.to(dev), dists.to(dev), d.to(dev), True, train=True)
    (ref["rgb"].sum() + 0.3 * ref["depth"].sum() + ref["T_left"].sum() + ref["l2_reg_specular"]).backward()
    want_g = torch.cat([hs["sigma"].grad, hs["tint"].grad, hs["diffuse"].grad, hs["specular"].grad], -1).reshape(R * S, 10)
    for fwd_packed in (0, 1):
        capi.lib().snrf_composite_set_fwd_packed(ctypes.c_int(fwd_packed))
        try:
            hp = packed.to(dev).requires_grad_(True)
            out = _render.composite_packed(hp, z.to(dev), dists.to(dev), d.to(dev), True, train=True)
            (out["rgb"].sum() + 0.3 * out["depth"].sum() + out["T_left"].sum() + out["l2_reg_specular"]).backward()
        finally:
            capi.lib().snrf_composite_set_fwd_packed(ctypes.c_int(0))
        for k in ("rgb", "depth", "T_left", "diffuse", "specular", "tint", "weights"):
            assert torch.equal(out[k], ref[k]), (fwd_packed, k)
        assert torch.equal(hp.grad, want_g), fwd_packed
