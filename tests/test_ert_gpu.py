"""GPU: early ray termination in training (opt-in, HashGrid.ert_eps / TileStep(ert_eps=...); north star subsystem 5).
The reference evaluates and back-propagates every sample (hashgrid/__init__.py:512-596), so the bar is relative to the
un-terminated path: eps -> 0 reproduces it to the bit; at eps = 1e-4 the composited colour stays within 1e-4 (the north
star's RGB bar) and the gradients within the weight of what was cut; the kernels agree with a torch restatement of the
truncated compositing (dead samples: weight 0, no gradient)."""
import pytest
import torch

from conftest import load_pkg

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _packed_inputs(R, S, seed, sigma_scale):
    g = torch.Generator().manual_seed(seed)
    heads = torch.rand(R * S, 10, generator=g)
    heads[:, 0] *= sigma_scale
    z = torch.cumsum(torch.rand(R, S, generator=g) * 0.1 + 0.01, -1)
    dists = torch.cat([z[:, 1:] - z[:, :-1], torch.full((R, 1), 1e-6)], -1)
    d = torch.randn(R, 3, generator=g) * 1.3
    return heads, z, dists, d


def _truncated_oracle(heads, z, dists, d, infinity, eps, front=None):
    """torch restatement: weights of samples behind transmittance < eps are zero and their heads carry no gradient."""
    R, S = z.shape
    h = heads.reshape(R, S, 10)
    delta = dists * d.norm(dim=-1, keepdim=True)
    if infinity:
        delta = torch.cat([delta[:, :-1], torch.full_like(delta[:, :1], 1e10)], -1)
    alpha = 1.0 - torch.exp(-h[..., 0] * delta)
    T = torch.cumprod(torch.cat([torch.ones(R, 1), 1.0 - alpha + 1e-6], -1), -1)[:, :-1]
    f = torch.ones(R, 1) if front is None else front
    dead = (f * T).detach() < eps
    # dead samples: constants (no gradient through their alpha), weight 0
    alpha_e = torch.where(dead, alpha.detach(), alpha)
    T_e = torch.cumprod(torch.cat([torch.ones(R, 1), 1.0 - alpha_e + 1e-6], -1), -1)[:, :-1]
    w = torch.where(dead, torch.zeros_like(T_e), alpha_e * T_e)
    hh = torch.where(dead[..., None], h.detach(), h)
    tint, dif, spe = hh[..., 1:4], hh[..., 4:7], hh[..., 7:10]
    out = {"depth": (w * z).sum(-1, keepdim=True), "tint": (w[..., None] * tint).sum(1), "diffuse": (w[..., None] * dif).sum(1),
           "specular": (w[..., None] * tint * spe).sum(1), "T_left": T_e[:, -1], "weights": w, "live": ~dead}
    return out


@pytest.mark.parametrize("R,S,infinity", [(64, 128, True), (33, 200, False), (7, 31, True)])
def test_tiny_eps_reproduces_the_unterminated_compositing(R, S, infinity):
    load_pkg()
    from hashgrid import _render
    heads, z, dists, d = _packed_inputs(R, S, R + S, 30.0)
    outs = []
    for ert in (None, _render.ErtState(1e-37)):
        hg = heads.to(DEV).requires_grad_(True)
        o = _render.composite_packed(hg, z.to(DEV), dists.to(DEV), d.to(DEV), infinity, train=True, ert=ert)
        (o["rgb"].sum() + o["depth"].sum() + o["T_left"].sum() + o["l2_reg_specular"]).backward()
        outs.append((o, hg.grad.clone(), ert))
    (a, ga, _), (b, gb, ert) = outs
    for k in ("rgb", "depth", "T_left", "diffuse", "specular", "tint"):
        assert torch.equal(a[k], b[k]), k
    # (only samples whose transmittance underflowed below 1e-37 -- subnormal weights -- can be flagged at this eps)
    assert float((a["weights"] - b["weights"]).abs().max()) < 1e-36
    assert float((ga - gb).abs().max()) < 1e-30
    assert float(ert.sample_live.float().mean()) > 0.2


@pytest.mark.parametrize("eps", [1e-4, 1e-2])
@pytest.mark.parametrize("R,S,infinity", [(64, 128, True), (33, 200, False)])
def test_terminated_compositing_matches_the_truncated_restatement(R, S, infinity, eps):
    load_pkg()
    from hashgrid import _render
    heads, z, dists, d = _packed_inputs(R, S, 3 * R + S, 12.0)
    hc = heads.clone().requires_grad_(True)
    ref = _truncated_oracle(hc, z, dists, d, infinity, eps)
    g = torch.Generator().manual_seed(9)
    cot = {k: torch.randn(ref[k].shape, generator=g) for k in ("depth", "tint", "diffuse", "specular", "T_left")}
    sum((ref[k] * cot[k]).sum() for k in cot).backward()
    hg = heads.to(DEV).requires_grad_(True)
    ert = _render.ErtState(eps)
    out = _render.composite_packed(hg, z.to(DEV), dists.to(DEV), d.to(DEV), infinity, train=True, ert=ert)
    sum((out[k] * cot[k].to(DEV)).sum() for k in cot).backward()
    live = ert.sample_live.reshape(R, S).bool().cpu()
    assert torch.equal(live, ref["live"]), "termination flags"
    assert 0.02 < float(live.float().mean()) < 0.98, "the case must terminate some samples, not all"
    for k in ("depth", "tint", "diffuse", "specular", "T_left"):
        assert torch.allclose(out[k].cpu(), ref[k].detach(), atol=2e-5, rtol=1e-4), k
    assert torch.allclose(out["weights"][..., 0].cpu(), ref["weights"].detach(), atol=1e-6, rtol=1e-4)
    a, b = hg.grad.cpu(), hc.grad
    assert float(a.reshape(R, S, 10)[~live].abs().max()) == 0.0, "dead samples get no gradient"
    scale = float(b.abs().max())
    assert float((a - b).abs().max()) < 2e-4 * scale + 3e-7
    # against the un-terminated compositing: colour within eps
    full = _render.composite_packed(heads.to(DEV), z.to(DEV), dists.to(DEV), d.to(DEV), infinity, train=True)
    assert float((full["rgb"] - out["rgb"]).abs().max()) < 2.0 * eps + 1e-6


def _dense_tile(dev, ert_eps):
    import test_tile_step_gpu as tts
    from tile_step import TileStep  # noqa: F401
    step, locs, gt = tts._tile(dev, log2T=15, S=32)
    step.featureGrid.ert_eps = ert_eps
    with torch.no_grad():           # an opaque field: rays saturate after a few samples
        step.decoder.sigma_layer.mlp[0].bias += 6.0
    return step, locs, gt


def test_terminated_step_stays_within_the_colour_bar_and_skips_work():
    """Whole step on an opaque field: colour within 1e-4 of the un-terminated step, gradients of the colour loss within the
    cut weight, and the joint batch terminates the background chain behind an opaque foreground.  (The specular L2
    regulariser, tile.py:999, is a sum over every sample's OWN-chain weight: terminating the occluded background samples drops
    their share of it -- the one term of the loss that sees samples the pixel does not; compared separately.)"""
    load_pkg()
    dev = torch.device(DEV)
    res = []
    for eps in (0.0, 1e-4):
        step, locs, gt = _dense_tile(dev, eps)
        loss, out = step.loss(locs.to(dev), gt.to(dev))
        loss = loss - 0.01 * out["l2_reg_specular"]          # the colour loss alone
        loss.backward()
        torch.cuda.synchronize()
        res.append((float(loss), out["pred_color"].detach().clone(), step.featureGrid.HE.features.grad.clone(),
                    [p.grad.clone() for p in step.decoder.parameters()], step.poses.se3_refine.grad.clone()))
    (l0, c0, t0, d0, p0), (l1, c1, t1, d1, p1) = res
    assert float((c0 - c1).abs().max()) < 1e-4, float((c0 - c1).abs().max())
    assert abs(l0 - l1) < 1e-5

    def rel(a, b):
        return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-20)
    assert rel(t1, t0) < 2e-3, rel(t1, t0)
    assert rel(p1, p0) < 2e-3, rel(p1, p0)
    # decoder parameters: against the largest parameter gradient of the step (on an opaque field the density head's own
    # gradient is orders of magnitude below the colour heads', and what termination cuts is measured on that common scale)
    scale = max(float(b.abs().max()) for b in d0)
    for a, b in zip(d1, d0):
        assert float((a - b).abs().max()) < 2e-3 * scale, (tuple(a.shape), float((a - b).abs().max()), scale)
    # the terminated step touched fewer table entries
    assert int((t1 != 0).sum()) < 0.8 * int((t0 != 0).sum()), (int((t1 != 0).sum()), int((t0 != 0).sum()))


def test_terminated_fused_update_equals_terminated_scatter_then_adam():
    """The scatter + Adam fusion honours the sample flags like the unfused pair: same tables after a step."""
    load_pkg()
    dev = torch.device(DEV)
    tabs = []
    for fused in (True, False):
        step, locs, gt = _dense_tile(dev, 1e-3)
        step.fused_table_update = fused
        step.step_device(locs.to(dev), gt.to(dev))
        step.step_device(locs.to(dev), gt.to(dev))
        torch.cuda.synchronize()
        tabs.append(step.featureGrid.HE.features.detach().clone())
    a, b = tabs
    # (two Adam steps of lr 1e-3: an entry the two paths treated differently would differ by ~1e-3; summation order of the
    # atomics alone moves a normalised update by a fraction of a percent)
    assert float((a - b).abs().max()) < 5e-5, float((a - b).abs().max())
