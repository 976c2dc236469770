"""GPU parity of the hash-grid encode (C-ABI snrf_hash_fwd / snrf_hash_bwd through
the reference-named Python ops) against the CPU oracle and, when oracle/_ref was
built, against the reference's own CUDA kernels on the same inputs.

Bars (BASELINE.json north_star): hash indices bit-exact; features and gradients
within 1e-5 relative (fp32).
"""
import numpy as np
import pytest
import torch

from conftest import load_pkg, ref_module
from oracle import native as on

pytestmark = pytest.mark.gpu
REL = 1e-5


def _ladder(base, fin, L=16):
    b = torch.exp((torch.log(torch.as_tensor(fin)) - torch.log(torch.as_tensor(base))) / (L - 1))
    return torch.stack([(torch.as_tensor(base) * b ** i).int() for i in range(L)], 0)


def _case(B, log2T, seed, bbox):
    g = torch.Generator().manual_seed(seed)
    L, T = 16, 2 ** log2T
    table = torch.randn(L, T, 2, generator=g) * 0.1
    res = _ladder(torch.tensor([49, 32, 73]), torch.tensor([1260, 819, 1890]), L)
    if bbox:
        corner = torch.tensor([-3.0, 1.0, 2.0])
        size = torch.tensor([20.0, 13.0, 30.0])
        pts = corner + size * (torch.rand(B, 3, generator=g) * 1.2 - 0.1)   # some outside -> clamp
    else:
        corner = size = None
        pts = torch.rand(B, 3, generator=g) * 4 - 2
    gin = torch.randn(B, L, 2, generator=g)
    return pts, table, res, corner, size, gin


def _relerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.mark.parametrize("bbox", [False, True])
@pytest.mark.parametrize("B,log2T", [(1, 12), (33, 12), (4097, 15), (20000, 19)])
def test_forward_and_indices_vs_oracle(bbox, B, log2T):
    load_pkg()
    from hashgrid.lib import HASHGRID as ops
    pts, table, res, corner, size, _ = _case(B, log2T, 1, bbox)
    dev = "cuda:0"
    c = corner.to(dev) if bbox else None
    s = size.to(dev) if bbox else None
    idx, out = ops.hash_indices(pts.to(dev), table.to(dev), res.to(dev), c, s)
    ref_out, ref_idx = on.hash_encode_fwd(pts.numpy(), table.numpy(), res.numpy(),
                                          corner.numpy() if bbox else None,
                                          size.numpy() if bbox else None, want_idx=True)
    assert np.array_equal(idx.cpu().numpy().astype(np.uint32), ref_idx), "hash indices must be bit-exact"
    assert _relerr(out.cpu().numpy(), ref_out) <= REL
    # the reference-named op writes the same thing in place
    out2 = torch.zeros(B, 16, 2, device=dev)
    if bbox:
        ops.embedding_forward_cuda(pts.to(dev), out2, table.to(dev), c, s, res.to(dev))
    else:
        ops.embedding_bg_forward_cuda(pts.to(dev), out2, table.to(dev), res.to(dev))
    assert torch.equal(out2, out)


@pytest.mark.parametrize("bbox", [False, True])
@pytest.mark.parametrize("agg", [0, 16])
@pytest.mark.parametrize("B,log2T", [(1, 12), (1000, 12), (30011, 16)])
def test_backward_vs_oracle(bbox, agg, B, log2T):
    load_pkg()
    from hashgrid.lib import HASHGRID as ops
    pts, table, res, corner, size, gin = _case(B, log2T, 2, bbox)
    # make runs of consecutive points share coarse cells, like samples along a ray
    pts = pts[:1] + (pts - pts[:1]) * torch.linspace(0, 1, B)[:, None] ** 2
    dev = "cuda:0"
    c = corner.to(dev) if bbox else None
    s = size.to(dev) if bbox else None
    gp = torch.zeros(B, 3, device=dev)
    gt = torch.zeros_like(table, device=dev)
    ops._encode_bwd(pts.to(dev), gin.to(dev), gp, gt, table.to(dev), c, s, res.to(dev), aggregate_levels=agg)
    rgp, rgt = on.hash_encode_bwd(pts.numpy(), gin.numpy(), table.numpy(), res.numpy(),
                                  corner.numpy() if bbox else None, size.numpy() if bbox else None)
    assert _relerr(gt.cpu().numpy(), rgt) <= REL
    assert _relerr(gp.cpu().numpy(), rgp) <= 5 * REL   # 16-term fp32 sum, order differs


def test_empty_and_errors():
    load_pkg()
    from hashgrid.lib import HASHGRID as ops
    dev = "cuda:0"
    table = torch.zeros(16, 1024, 2, device=dev)
    res = _ladder(16, 512).to(dev)[:, None].repeat(1, 3).contiguous()
    out = torch.zeros(0, 16, 2, device=dev)
    ops.embedding_bg_forward_cuda(torch.zeros(0, 3, device=dev), out, table, res)   # no-op
    with pytest.raises(RuntimeError):
        ops.embedding_bg_forward_cuda(torch.zeros(4, 3), torch.zeros(4, 16, 2), table.cpu(), res.cpu())  # CPU tensors
    bad = torch.zeros(16, 1000, 2, device=dev)   # T not a power of two
    with pytest.raises(RuntimeError):
        ops.embedding_bg_forward_cuda(torch.zeros(4, 3, device=dev), torch.zeros(4, 16, 2, device=dev), bad, res)


def test_autograd_module_matches_oracle():
    load_pkg()
    from hashgrid import PyHashGridBG
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    hg = PyHashGridBG(dev, None, None, log2_hashmap_size=14,
                      base_resolution=torch.tensor([16, 16, 16]), finest_resolution=torch.tensor([512, 512, 512]))
    x = (torch.rand(5, 7, 3, device=dev) * 4 - 2).requires_grad_(True)
    y = hg(x)
    assert y.shape == (5, 7, 32)
    w = torch.randn_like(y)
    (y * w).sum().backward()
    ro = on.hash_encode_fwd(x.detach().cpu().reshape(-1, 3).numpy(), hg.features.detach().cpu().numpy(),
                            hg.resolution.cpu().numpy())
    assert _relerr(y.detach().cpu().numpy().reshape(-1, 16, 2), ro) <= REL
    rgp, rgt = on.hash_encode_bwd(x.detach().cpu().reshape(-1, 3).numpy(), w.cpu().reshape(-1, 16, 2).numpy(),
                                  hg.features.detach().cpu().numpy(), hg.resolution.cpu().numpy())
    assert _relerr(hg.features.grad.cpu().numpy(), rgt) <= REL
    assert _relerr(x.grad.cpu().reshape(-1, 3).numpy(), rgp) <= 5 * REL


@pytest.mark.parametrize("bbox", [False, True])
def test_against_reference_cuda_kernels(bbox):
    """The reference's own kernels (unmodified sources rebuilt for sm_100a) on the same inputs."""
    ref = ref_module("HASHGRID_EMBED")
    if ref is None:
        pytest.skip("oracle/_ref/HASHGRID_EMBED.so not built")
    load_pkg()
    from hashgrid.lib import HASHGRID as ops
    B = 50000
    pts, table, res, corner, size, gin = _case(B, 19, 3, bbox)
    dev = "cuda:0"
    pts, table, res, gin = pts.to(dev), table.to(dev), res.to(dev), gin.to(dev)
    o_ref = torch.zeros(B, 16, 2, device=dev)
    o_new = torch.zeros(B, 16, 2, device=dev)
    gp_ref, gp_new = torch.zeros(B, 3, device=dev), torch.zeros(B, 3, device=dev)
    gt_ref, gt_new = torch.zeros_like(table), torch.zeros_like(table)
    if bbox:
        c, s = corner.to(dev), size.to(dev)
        ref.embedding_forward_cuda(pts, o_ref, table, c, s, res)
        ops.embedding_forward_cuda(pts, o_new, table, c, s, res)
        ref.embedding_backward_cuda(pts, gin, gp_ref, gt_ref, table, c, s, res)
        ops.embedding_backward_cuda(pts, gin, gp_new, gt_new, table, c, s, res)
    else:
        ref.embedding_bg_forward_cuda(pts, o_ref, table, res)
        ops.embedding_bg_forward_cuda(pts, o_new, table, res)
        ref.embedding_bg_backward_cuda(pts, gin, gp_ref, gt_ref, table, res)
        ops.embedding_bg_backward_cuda(pts, gin, gp_new, gt_new, table, res)
    torch.cuda.synchronize()
    assert _relerr(o_new.cpu().numpy(), o_ref.cpu().numpy()) <= REL
    assert _relerr(gt_new.cpu().numpy(), gt_ref.cpu().numpy()) <= REL
    assert _relerr(gp_new.cpu().numpy(), gp_ref.cpu().numpy()) <= 5 * REL
