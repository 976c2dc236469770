"""GPU: the fused training-path field encode (csrc/field_encode.cu: sample position -> contraction ->
hash encode, level-major, stored Jacobians, range-partitioned gradient scatter) against the unfused
chain it replaces: torch `o + z d`, HashGrid.contract_fore / contract_bg (golden-pinned restatements of
hashgrid/__init__.py:394-411) and the reference-shaped encode operator (itself parity-checked against
the C oracle and the rebuilt reference kernels in test_hash_encode_gpu.py).
Bars: forward bit-identical (same contracted points, same blend order => same hash cells and sums);
gradients 1e-5 relative (fp32 atomics: summation order differs)."""
import numpy as np
import pytest
import torch

from conftest import load_pkg
from oracle import native as on
from oracle import torch_ref as tr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _case(R, S, log2T, seed):
    g = torch.Generator().manual_seed(seed)
    L, T = 16, 2 ** log2T
    table = torch.randn(L, T, 2, generator=g) * 0.1
    res = tr.resolution_ladder(torch.tensor([24, 16, 36]), torch.tensor([3000, 2048, 4500])).int()
    bmin, bsize = torch.tensor([-10.0, -6.5, -15.0]), torch.tensor([40.0, 26.0, 60.0])
    o = torch.tensor([10.0, 6.5, 15.0]) + (torch.rand(R, 3, generator=g) - 0.5) * torch.tensor([16.0, 10.0, 24.0])
    d = torch.randn(R, 3, generator=g) * (0.5 + torch.rand(R, 1, generator=g))
    z_fg = (torch.rand(R, S, generator=g) * 4.0).sort(-1)[0]
    z_bg = 12.0 + (torch.rand(R, S, generator=g) * 300.0).sort(-1)[0]
    return table, res, bmin, bsize, o, d, z_fg.contiguous(), z_bg.contiguous(), g


def _rel(a, b):
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-20)


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("R,S,log2T", [(1, 1, 10), (37, 5, 12), (512, 64, 19)])
def test_fused_encode_matches_unfused_chain(R, S, log2T, mode):
    load_pkg()
    from hashgrid import _field
    from hashgrid.PyHashGridBG import HashEmbeddingBG
    table, res, bmin, bsize, o, d, z_fg, z_bg, g = _case(R, S, log2T, R + S)
    z = z_fg if mode == 1 else z_bg
    N = R * S
    # ---- unfused chain (torch position + contraction, reference-shaped encode op)
    o1, d1 = o.to(DEV).requires_grad_(True), d.to(DEV).requires_grad_(True)
    t1 = table.to(DEV).requires_grad_(True)
    x = (o1[:, None] + z.to(DEV)[..., None] * d1[:, None]).reshape(-1, 3)
    cx = (tr.contract_fore if mode == 1 else tr.contract_bg)(x, bmin.to(DEV), bsize.to(DEV))
    ref = HashEmbeddingBG(cx, t1, res.to(DEV))                                   # [N,16,2]
    cot = torch.randn(N, 16, 2, generator=g).to(DEV)
    (ref * cot).sum().backward()
    # ---- fused
    o2, d2 = o.to(DEV).requires_grad_(True), d.to(DEV).requires_grad_(True)
    t2 = torch.nn.Parameter(table.to(DEV).clone())
    out = _field.field_encode(o2, d2, z.to(DEV), t2, res.to(DEV), bmin.to(DEV), bsize.to(DEV), mode)       # [16,N,2]
    assert out.shape == (16, N, 2)
    assert torch.equal(out.permute(1, 0, 2), ref.detach()), f"forward differs: {float((out.permute(1, 0, 2) - ref).abs().max())}"
    (out * cot.permute(1, 0, 2)).sum().backward()
    torch.cuda.synchronize()
    assert _rel(t2.grad, t1.grad) < 1e-5, "table gradient"
    assert _rel(o2.grad, o1.grad) < 2e-5, f"d/d rays_o: {_rel(o2.grad, o1.grad)}"
    assert _rel(d2.grad, d1.grad) < 2e-5, f"d/d rays_d: {_rel(d2.grad, d1.grad)}"
    # ---- and against the C oracle on the same contracted points (hash_bg_kernel restatement)
    want = on.hash_encode_fwd(cx.detach().cpu().numpy(), table.numpy(), res.numpy())
    assert np.allclose(out.detach().permute(1, 0, 2).cpu().numpy(), want, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("pair,first_level,l2", [(1, 0, 0), (2, 0, 0), (2, 9, 0), (1, 0, 1), (2, 0, 1), (0, 0, 1)])
@pytest.mark.parametrize("R,S,log2T,mode", [(37, 5, 12, 1), (512, 64, 19, 2), (256, 32, 21, 3)])
def test_forward_load_variants_are_bit_identical(R, S, log2T, mode, pair, first_level, l2):
    """snrf_field_set_fwd_pair_loads / snrf_field_set_fwd_l2_policy change how the table is fetched (16-byte x-pairs, whole
    32-byte sectors, L2 eviction policy), never what is computed: features and ray gradients (through the stored
    Jacobians) equal to the bit.  log2T = 21 is a table large enough for one level per CTA row (where the policy applies)."""
    import ctypes
    load_pkg()
    import scanerf_b200_capi as capi
    from hashgrid import _field
    table, res, bmin, bsize, o, d, z_fg, z_bg, g = _case(R, S, log2T, R + S + 1)
    z = torch.cat([z_fg[: R // 2], z_bg[R // 2:]]) if mode == 3 else (z_fg if mode == 1 else z_bg)
    cot = torch.randn(16, R * S, 2, generator=g).to(DEV)
    lib = capi.lib()

    def run():
        o2, d2 = o.to(DEV).requires_grad_(True), d.to(DEV).requires_grad_(True)
        t2 = torch.nn.Parameter(table.to(DEV).clone())
        out = _field.field_encode(o2, d2, z.to(DEV), t2, res.to(DEV), bmin.to(DEV), bsize.to(DEV), mode, split=R // 2)
        (out * cot).sum().backward()
        torch.cuda.synchronize()
        return out.detach().clone(), o2.grad.clone(), d2.grad.clone()

    try:
        lib.snrf_field_set_fwd_pair_loads(ctypes.c_int(0), ctypes.c_int(0))
        lib.snrf_field_set_fwd_l2_policy(ctypes.c_int(0), ctypes.c_int(0))
        want = run()
        lib.snrf_field_set_fwd_pair_loads(ctypes.c_int(pair), ctypes.c_int(first_level))
        lib.snrf_field_set_fwd_l2_policy(ctypes.c_int(l2), ctypes.c_int(0))
        got = run()
    finally:
        lib.snrf_field_set_fwd_pair_loads(ctypes.c_int(0), ctypes.c_int(0))
        lib.snrf_field_set_fwd_l2_policy(ctypes.c_int(0), ctypes.c_int(0))        # the library's default
    assert torch.equal(got[0], want[0]), "features"
    # (the ray gradients are sums of atomics: same values, run-to-run order)
    assert _rel(got[1], want[1]) < 1e-5 and _rel(got[2], want[2]) < 1e-5


@pytest.mark.parametrize("bits", [0, 1, 3])
def test_scatter_passes_are_equivalent(bits):
    """The range-partitioned scatter is a cache optimisation only: any number of passes gives the same table gradient."""
    load_pkg()
    import scanerf_b200_capi as capi
    from hashgrid import _field
    table, res, bmin, bsize, o, d, z_fg, _, g = _case(256, 32, 14, 3)
    cot = torch.randn(16, 256 * 32, 2, generator=g).to(DEV)
    grads = []
    for b in (-1, bits):
        capi.lib().snrf_field_set_passes_log2(capi.c_int(b))
        t = torch.nn.Parameter(table.to(DEV).clone())
        out = _field.field_encode(o.to(DEV), d.to(DEV), z_fg.to(DEV), t, res.to(DEV), bmin.to(DEV), bsize.to(DEV), 1)
        (out * cot).sum().backward()
        grads.append(t.grad.clone())
    capi.lib().snrf_field_set_passes_log2(capi.c_int(-1))
    assert _rel(grads[1], grads[0]) < 1e-6


@pytest.mark.parametrize("S", [32, 7, 1])
@pytest.mark.parametrize("run", [2, 4, 8])
def test_run_merging_scatter_equals_cross_lane_scatter(run, S):
    """The scatter kernel that merges runs of samples in one cell in registers (R samples per thread) against the
    cross-lane kernel: same table gradient and same d/d rays_o, d/d rays_d, for sample counts that are and are not
    multiples of R (a thread's samples then straddle two rays), with masked rays and a fore/background split."""
    load_pkg()
    import scanerf_b200_capi as capi
    from hashgrid import _field
    Rn = 301
    table, res, bmin, bsize, o, d, z_fg, _, g = _case(Rn, S, 14, 5 + S)
    cot = torch.randn(16, Rn * S, 2, generator=g).to(DEV)
    valid = (torch.rand(Rn, generator=g) < 0.8).to(DEV)
    out = []
    for r in (0, run):
        capi.lib().snrf_field_set_run_length(capi.c_int(r))
        t = torch.nn.Parameter(table.to(DEV).clone())
        oo, dd = o.to(DEV).clone().requires_grad_(True), d.to(DEV).clone().requires_grad_(True)
        enc = _field.field_encode(oo, dd, z_fg.to(DEV), t, res.to(DEV), bmin.to(DEV), bsize.to(DEV), 3, valid, 120)
        (enc * cot * valid.repeat_interleave(S)[None, :, None]).sum().backward()
        out.append((t.grad.clone(), oo.grad.clone(), dd.grad.clone()))
    capi.lib().snrf_field_set_run_length(capi.c_int(0))
    for name, a, b in zip(("table", "rays_o", "rays_d"), out[1], out[0]):
        assert _rel(a, b) < 2e-6, name
    assert float(out[1][0].abs().max()) > 0 and float(out[1][1].abs().max()) > 0


@pytest.mark.parametrize("mode,S,log2T,bits,lpg,slice_log2", [(1, 32, 14, -1, 0, 23), (3, 7, 12, 2, 0, 23), (2, 16, 10, -1, 3, 23),
                                                              (3, 33, 16, 1, 0, 23), (3, 32, 14, -1, 0, 12), (1, 9, 15, -1, 0, 13)])
def test_scatter_update_fusion_equals_scatter_then_adam(mode, S, log2T, bits, lpg, slice_log2):
    """snrf_field_encode_bwd_adam (gradient slices scattered into the L2-resident scratch and consumed by the sparse Adam
    on the spot) against snrf_field_encode_bwd into a gradient table followed by snrf_adam_step, over two steps (the second
    with non-trivial moments): same parameters, moments, ray gradients; untouched entries bit-identical; the scratch is
    left all-zero; index ranges (bits), ragged level groups (lpg) and -- with a slice smaller than a level (slice_log2) --
    the single-pass coarse levels on the side stream next to the range-partitioned fine levels included."""
    load_pkg()
    import scanerf_b200_capi as capi
    from hashgrid import _field, _gradmode
    from vdbAdam import vdbAdam
    Rn = 301
    table, res, bmin, bsize, o, d, z_fg, z_bg, g = _case(Rn, S, log2T, 11 + S)
    z = z_bg if mode == 2 else z_fg
    cots = [torch.randn(16, Rn * S, 2, generator=g).to(DEV) for _ in range(2)]
    valid = (torch.rand(Rn, generator=g) < 0.8).to(DEV)
    vmask = valid.repeat_interleave(S)[None, :, None]
    results = []
    for fused in (False, True):
        t = torch.nn.Parameter(table.to(DEV).clone())
        opt = vdbAdam([t], lr=1e-2, betas=(0.9, 0.99), eps=1e-15, bias_correction="standard", fused_zero_grad=True)
        capi.lib().snrf_field_set_passes_log2(capi.c_int(bits if fused else -1))
        capi.lib().snrf_field_set_levels_per_group(capi.c_int(lpg if fused else 0))
        capi.lib().snrf_field_set_slice_log2(capi.c_int(slice_log2))
        opt.scratch_log2 = slice_log2
        grads = []
        for cot in cots:
            oo, dd = o.to(DEV).clone().requires_grad_(True), d.to(DEV).clone().requires_grad_(True)
            enc = _field.field_encode(oo, dd, z.to(DEV), t, res.to(DEV), bmin.to(DEV), bsize.to(DEV), mode, valid, 120)
            opt.zero_grad()
            with opt.table_backward(fused=fused):
                (enc * cot * vmask).sum().backward()
            opt.step()
            grads.append((oo.grad.clone(), dd.grad.clone()))
        capi.lib().snrf_field_set_passes_log2(capi.c_int(-1))
        capi.lib().snrf_field_set_levels_per_group(capi.c_int(0))
        capi.lib().snrf_field_set_slice_log2(capi.c_int(23))
        if fused:
            if slice_log2 < log2T:
                assert opt._scratch.shape[0] == 2 ** log2T + 2 ** slice_log2, "coarse single-pass region + one fine slice"
            assert t.grad is None, "the fused path must not materialise a gradient table"
            assert float(opt._scratch.abs().max()) == 0.0, "scratch must be left all-zero"
        results.append((t.detach().clone(), opt.params[0][1].clone(), opt.params[0][2].clone(), grads, opt.t))
    (p0, m0, v0, g0, t0), (p1, m1, v1, g1, t1) = results
    assert t0 == t1 == 2
    orig = table.to(DEV)
    assert torch.equal(p0 == orig, p1 == orig), "the same entries are touched"
    assert 0 < int((p1 != orig).sum()) < p1.numel()
    assert _rel(m1, m0) < 1e-5 and _rel(v1, v0) < 1e-5
    # the update is lr * m / sqrt(v): compare where the gradient is not a cancellation residue
    assert float(((p1 - p0).abs() > 1e-5).float().mean()) < 1e-4, float(((p1 - p0).abs() > 1e-5).float().mean())
    for (a_o, a_d), (b_o, b_d) in zip(g1, g0):
        assert _rel(a_o, b_o) < 2e-5 and _rel(a_d, b_d) < 2e-5


def test_second_encode_in_fused_backward_raises():
    load_pkg()
    from hashgrid import _field
    from vdbAdam import vdbAdam
    table, res, bmin, bsize, o, d, z_fg, _, g = _case(8, 4, 10, 2)
    t = torch.nn.Parameter(table.to(DEV).clone())
    opt = vdbAdam([t], bias_correction="standard", fused_zero_grad=True)
    args = (o.to(DEV), d.to(DEV), z_fg.to(DEV), t, res.to(DEV), bmin.to(DEV), bsize.to(DEV), 1)
    loss = _field.field_encode(*args).sum() + _field.field_encode(*args).sum()
    with pytest.raises(RuntimeError, match="more than once"):
        with opt.table_backward(fused=True):
            loss.backward()


def test_encode_backward_has_no_side_effect_outside_the_training_context():
    """a backward outside `table_backward` (normal / validation pass, torch.autograd.grad) returns the dense
    table gradient through autograd and leaves `.grad` alone; one that lands in `.grad` between step() and zero_grad() is
    cleared by zero_grad() even though the fused update had left the tensor clean."""
    load_pkg()
    from hashgrid import _field
    from vdbAdam import vdbAdam
    table, res, bmin, bsize, o, d, z_fg, _, g = _case(16, 8, 10, 4)
    t = torch.nn.Parameter(table.to(DEV).clone())
    args = (o.to(DEV), d.to(DEV), z_fg.to(DEV), t, res.to(DEV), bmin.to(DEV), bsize.to(DEV), 1)
    (gt,) = torch.autograd.grad(_field.field_encode(*args).sum(), t)
    assert t.grad is None and float(gt.abs().max()) > 0
    opt = vdbAdam([t], bias_correction="standard", fused_zero_grad=True)
    with opt.table_backward(fused=False):
        _field.field_encode(*args).sum().backward()
    opt.step()
    assert float(t.grad.abs().max()) == 0.0
    _field.field_encode(*args).sum().backward()           # e.g. a validation pass: autograd accumulates into .grad
    assert float(t.grad.abs().max()) > 0
    opt.zero_grad()
    assert float(t.grad.abs().max()) == 0.0, "zero_grad must not trust the clean flag after a foreign backward"


def test_encodes_of_one_backward_pass_share_the_dense_gradient_table():
    """The reference's drivers encode the table twice per step (foreground + background chain, tile.py:661-681).  Outside any
    training context the encode Functions return a dense grad_features like the reference's (PyHashGridBG.py:20-30), but the
    encodes of ONE backward pass scatter into one shared tensor (_gradmode.shared_table_grad): same `.grad` as two
    independent tensors summed by autograd, for the fused Functions and for the reference-shaped PyHashGridBG Function."""
    load_pkg()
    from hashgrid import _embedding, _field, _gradmode
    table, res, bmin, bsize, o, d, z_fg, z_bg, g = _case(16, 10, 19, 8)
    t = torch.nn.Parameter(table.to(DEV).clone())
    common = (res.to(DEV), bmin.to(DEV), bsize.to(DEV))
    w = torch.randn(16, z_fg.numel(), 2, generator=torch.Generator().manual_seed(3)).to(DEV)
    pts = (torch.rand(512, 3, generator=torch.Generator().manual_seed(4)) * 4 - 2).to(DEV)
    w2 = torch.randn(512, 16, 2, generator=torch.Generator().manual_seed(5)).to(DEV)
    got = {}
    for share in (True, False):
        _gradmode.share_enabled = share
        try:
            t.grad = None
            for _ in range(2):           # second pass: accumulates onto .grad like any autograd gradient
                loss = (_field.field_encode(o.to(DEV), d.to(DEV), z_fg.to(DEV), t, *common, 1) * w).sum() \
                    + (_field.field_encode(o.to(DEV), d.to(DEV), z_bg.to(DEV), t, *common, 2) * w).sum() \
                    + (_embedding._EncodeFn.apply(pts, t, None, None, res.to(DEV), False) * w2).sum()
                loss.backward()
            got[share] = t.grad.detach().clone()
        finally:
            _gradmode.share_enabled = True
    assert float(got[False].abs().max()) > 0
    assert torch.allclose(got[True], got[False], rtol=1e-5, atol=1e-6)


def test_hashgrid_fused_and_unfused_render_agree():
    """HashGrid.render_batch_rays with and without the fused encode gives the same composited colours and gradients."""
    load_pkg()
    from hashgrid import HashGrid, TRAIN
    from hashgrid._decoder import ShallowMLP
    dev = torch.device(DEV)
    torch.manual_seed(0)
    hg = HashGrid(dev, torch.tensor([0.0, 0.0, 0.0], device=dev), torch.tensor([20.0, 13.0, 30.0], device=dev), 15, [16, 256], 4, False, "")
    dec = ShallowMLP(32).to(dev)
    g = torch.Generator().manual_seed(1)
    R, S = 96, 32
    o = (torch.tensor([10.0, 6.5, 15.0]) + torch.randn(R, 3, generator=g)).to(dev)
    d = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1).to(dev)
    z = (torch.rand(R, S, generator=g) * 6).sort(-1)[0].to(dev)
    dist = torch.cat([z[:, 1:] - z[:, :-1], torch.full((R, 1), 1e-3, device=dev)], -1)
    res = {}
    for fused in (True, False):
        hg.fused_encode = fused
        hg.HE.features.grad = None
        oo, dd = o.clone().requires_grad_(True), d.clone().requires_grad_(True)
        for contract, inf in ((hg.contract_fore, False), (hg.contract_bg, True)):
            out, ok = hg.render_batch_rays(oo, dd, z if not inf else z + 20.0, dist, dec, TRAIN, contract, infinity=inf, global_step=9000)
            assert ok
            (out["rgb"].sum() + out["depth"].sum() * 0.01 + out["l2_reg_specular"]).backward()
            res[(fused, inf)] = (out["rgb"].detach().clone(), oo.grad.clone(), dd.grad.clone(), hg.HE.features.grad.clone())
    for inf in (False, True):
        a, b = res[(True, inf)], res[(False, inf)]
        assert torch.allclose(a[0], b[0], atol=1e-6), "rgb"
        for k, name in ((1, "d/d rays_o"), (2, "d/d rays_d"), (3, "table gradient")):
            assert _rel(a[k], b[k]) < 1e-4, f"{name} (infinity={inf}): {_rel(a[k], b[k])}"


def test_out_normal_takes_the_torch_decoder_branch():
    """render_batch_rays(out_normal=True) with the default (tensor-core) decoder must not run the fused chain
    and raise afterwards: it is routed to the torch decoder branch up front and returns the reference's `normal`."""
    load_pkg()
    from hashgrid import HashGrid, INFERENCE
    from hashgrid._decoder import ShallowMLP
    dev = torch.device(DEV)
    torch.manual_seed(0)
    hg = HashGrid(dev, torch.tensor([0.0, 0.0, 0.0], device=dev), torch.tensor([20.0, 13.0, 30.0], device=dev), 13, [16, 128], 4, False, "")
    dec = ShallowMLP(32).to(dev)
    g = torch.Generator().manual_seed(2)
    R, S = 33, 16
    o = (torch.tensor([10.0, 6.5, 15.0]) + torch.randn(R, 3, generator=g)).to(dev)
    d = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1).to(dev)
    z = (torch.rand(R, S, generator=g) * 6).sort(-1)[0].to(dev)
    dist = torch.cat([z[:, 1:] - z[:, :-1], torch.full((R, 1), 1e-3, device=dev)], -1)
    assert hg.fused_decoder
    # (as in the reference, the sample positions must be part of the autograd graph: pose-refined rays are)
    out, ok = hg.render_batch_rays(o.clone().requires_grad_(True), d, z, dist, dec, INFERENCE, hg.contract_fore, out_normal=True, global_step=9000)
    assert ok and out["normal"].shape == (R, 3) and bool(torch.isfinite(out["normal"]).all())
    plain, _ = hg.render_batch_rays(o, d, z, dist, dec, INFERENCE, hg.contract_fore, global_step=9000)
    assert torch.allclose(plain["rgb"], out["rgb"], atol=1e-4)


def test_masked_render_equals_compacted_render():
    """render_fore_rays / render_bg_rays: the sync-free masked path (every kernel skips masked-out rays)
    against the reference-style boolean compaction + scatter-back, values and gradients."""
    load_pkg()
    from hashgrid import HashGrid, TRAIN
    from hashgrid._decoder import ShallowMLP
    dev = torch.device(DEV)
    torch.manual_seed(0)
    hg = HashGrid(dev, torch.tensor([0.0, 0.0, 0.0], device=dev), torch.tensor([20.0, 13.0, 30.0], device=dev), 15, [16, 256], 4, False, "")
    g = torch.Generator().manual_seed(2)
    hg.occupied_grid = (torch.rand(hg.occupied_grid.shape, generator=g) < 0.5).to(dev)
    dec = ShallowMLP(32).to(dev)
    R, S = 300, 32
    o = torch.tensor([10.0, 6.5, 15.0]) + torch.randn(R, 3, generator=g) * torch.tensor([14.0, 6.0, 20.0])     # many origins outside the tile
    d = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1)
    occl = (torch.rand(R, 1, generator=g) < 0.8).to(dev)
    res = {}
    for fused in (True, False):
        hg.fused_encode = fused
        hg.HE.features.grad = None
        dec.zero_grad()
        oo, dd = o.to(dev).requires_grad_(True), d.to(dev).requires_grad_(True)
        fg, ok1 = hg.render_fore_rays(oo, dd, S, dec, TRAIN, occlusion_mask=occl, global_step=7000)
        bg, ok2 = hg.render_bg_rays(oo, dd, S, dec, TRAIN, occlusion_mask=occl, global_step=7000, bg_mode="IZ", invalid_underground=True)
        assert ok1 and ok2
        assert 0 < int(fg["fore_valid"].sum()) < R and 0 < int(bg["valid"].sum()) <= R
        col = fg["pred_color"] + fg["T_left"] * bg["rgb"]
        loss = (col ** 2).sum() + 0.1 * fg["pred_depth"].sum() + fg["l2_reg_specular"] + bg["l2_reg_specular"]
        loss.backward()
        res[fused] = dict(col=col.detach().clone(), T=fg["T_left"].detach().clone(), depth=bg["depth"].detach().clone(),
                          go=oo.grad.clone(), gd=dd.grad.clone(), gt=hg.HE.features.grad.clone(),
                          gw=dec.Spatial_MLP.mlp[0].weight.grad.clone(), l2=float(fg["l2_reg_specular"]) + float(bg["l2_reg_specular"]))
    a, b = res[True], res[False]
    assert a["T"].shape == b["T"].shape == (R, 1)
    for k in ("col", "T", "depth"):
        assert torch.allclose(a[k], b[k], atol=1e-5), k
    assert abs(a["l2"] - b["l2"]) < 1e-6
    for k in ("go", "gd", "gt", "gw"):
        assert _rel(a[k], b[k]) < 2e-4, f"{k}: {_rel(a[k], b[k])}"


def test_joint_fore_background_batch_equals_separate_chains():
    """HashGrid.render_fore_bg_rays (one 2R-ray batch through every kernel) against render_fore_rays + render_bg_rays."""
    load_pkg()
    from hashgrid import HashGrid, TRAIN
    from hashgrid._decoder import ShallowMLP
    dev = torch.device(DEV)
    torch.manual_seed(0)
    hg = HashGrid(dev, torch.tensor([0.0, 0.0, 0.0], device=dev), torch.tensor([20.0, 13.0, 30.0], device=dev), 15, [16, 256], 4, False, "")
    g = torch.Generator().manual_seed(4)
    hg.occupied_grid = (torch.rand(hg.occupied_grid.shape, generator=g) < 0.6).to(dev)
    dec = ShallowMLP(32).to(dev)
    R, S = 257, 32
    o = torch.tensor([10.0, 6.5, 15.0]) + torch.randn(R, 3, generator=g) * torch.tensor([12.0, 5.0, 18.0])
    d = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1)
    res = {}
    for joint in (True, False):
        hg.HE.features.grad = None
        dec.zero_grad()
        oo, dd = o.to(dev).requires_grad_(True), d.to(dev).requires_grad_(True)
        if joint:
            fg, bg = hg.render_fore_bg_rays(oo, dd, S, dec, TRAIN, global_step=8000, invalid_underground=True)
        else:
            fg, _ = hg.render_fore_rays(oo, dd, S, dec, TRAIN, global_step=8000)
            bg, _ = hg.render_bg_rays(oo, dd, S, dec, TRAIN, global_step=8000, bg_mode="IZ", invalid_underground=True)
        col = fg["pred_color"] + fg["T_left"] * bg["rgb"]
        ((col ** 2).sum() + 0.1 * (fg["pred_depth"] + fg["T_left"] * bg["depth"]).sum() + fg["l2_reg_specular"] + bg["l2_reg_specular"]).backward()
        res[joint] = dict(col=col.detach().clone(), fv=fg["fore_valid"].clone(), bv=bg["valid"].clone(), go=oo.grad.clone(), gd=dd.grad.clone(),
                          gt=hg.HE.features.grad.clone(), gw=dec.Directional_MLP.mlp[2].weight.grad.clone())
    a, b = res[True], res[False]
    assert torch.equal(a["fv"], b["fv"]) and torch.equal(a["bv"], b["bv"])
    assert torch.allclose(a["col"], b["col"], atol=1e-6)
    for k in ("go", "gd", "gt", "gw"):
        assert _rel(a[k], b[k]) < 1e-4, f"{k}: {_rel(a[k], b[k])}"
