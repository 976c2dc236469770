"""GPU parity of the sparse Adam step (snrf_adam_step via cuda.adam_step_cuda{,_fp16} and
vdbAdam) against the C oracle and, when built, the reference's own kernel."""
import numpy as np
import pytest
import torch

from conftest import load_pkg, ref_module
from oracle import native as on

pytestmark = pytest.mark.gpu


def _case(K, D, seed, sparsity=0.5):
    g = torch.Generator().manual_seed(seed)
    p = torch.randn(K, D, generator=g)
    grad = torch.randn(K, D, generator=g) * 1e-2
    grad[torch.rand(K, D, generator=g) < sparsity] = 0.0
    m = torch.randn(K, D, generator=g) * 1e-3
    v = torch.rand(K, D, generator=g) * 1e-5
    return p, grad, m, v


@pytest.mark.parametrize("K", [1, 1000, 100003])
def test_adam_step_matches_oracle(K):
    load_pkg()
    import cuda
    p, g, m, v = _case(K, 8, K)
    dev = "cuda:0"
    pg, gg, mg, vg = p.to(dev), g.to(dev), m.to(dev), v.to(dev)
    cuda.adam_step_cuda(pg, gg, mg, vg, 1e-3, 0.9, 0.999, 1e-6, 4)          # kernel sees step 5
    rp, rm, rv = on.adam_step(p.numpy(), g.numpy(), m.numpy(), v.numpy(), 1e-3, 0.9, 0.999, 1e-6, 5)
    assert np.allclose(pg.cpu().numpy(), rp, rtol=1e-6, atol=1e-7)
    assert np.allclose(mg.cpu().numpy(), rm, rtol=1e-6, atol=1e-9)
    assert np.allclose(vg.cpu().numpy(), rv, rtol=1e-6, atol=1e-12)
    untouched = (g == 0)
    assert torch.equal(pg.cpu()[untouched], p[untouched]) and torch.equal(mg.cpu()[untouched], m[untouched])


def test_adam_fp16_state_matches_oracle():
    load_pkg()
    import cuda
    p, g, m, v = _case(5000, 8, 2)
    m, v = (m * 100).half(), (v * 1e4).half()
    dev = "cuda:0"
    pg, gg, mg, vg = p.to(dev), g.to(dev), m.to(dev), v.to(dev)
    cuda.adam_step_cuda_fp16(pg, gg, mg, vg, 1e-3, 0.9, 0.999, 1e-6, 0)
    rp, rm, rv = on.adam_step(p.numpy(), g.numpy(), m.float().numpy(), v.float().numpy(), 1e-3, 0.9, 0.999, 1e-6, 1,
                              half_state=True)
    assert np.allclose(pg.cpu().numpy(), rp, rtol=1e-5, atol=1e-6)
    assert np.allclose(mg.float().cpu().numpy(), rm, rtol=2e-3, atol=1e-6)
    assert np.allclose(vg.float().cpu().numpy(), rv, rtol=2e-3, atol=1e-8)


def test_dense_layout_and_fused_zero_grad():
    load_pkg()
    import cuda
    L, T = 3, 4099
    g0 = torch.Generator().manual_seed(7)
    p = torch.randn(L, T, 2, generator=g0)
    grad = torch.randn(L, T, 2, generator=g0)
    grad[torch.rand(L, T, 2, generator=g0) < 0.7] = 0
    dev = "cuda:0"
    pg, gg = p.to(dev), grad.to(dev)
    mg, vg = torch.zeros_like(pg), torch.zeros_like(pg)
    cuda.adam_step_sparse(pg, gg, mg, vg, 1e-2, 0.9, 0.99, 1e-15, 1, zero_grad=True)
    rp, rm, rv = on.adam_step(p.reshape(-1, 1).numpy(), grad.reshape(-1, 1).numpy(), np.zeros((L * T * 2, 1), np.float32),
                              np.zeros((L * T * 2, 1), np.float32), 1e-2, 0.9, 0.99, 1e-15, 1)
    assert np.allclose(pg.cpu().numpy().reshape(-1, 1), rp, rtol=1e-6, atol=1e-7)
    assert float(gg.abs().max()) == 0.0, "consumed gradients must be cleared"


def test_against_reference_adam_kernel():
    ref = ref_module("CUDA_EXT")
    if ref is None:
        pytest.skip("oracle/_ref/CUDA_EXT.so not built")
    load_pkg()
    import cuda
    p, g, m, v = _case(20000, 8, 11)
    dev = "cuda:0"
    a = [t.to(dev).clone() for t in (p, g, m, v)]
    b = [t.to(dev).clone() for t in (p, g, m, v)]
    ref.adam_step_cuda(a[0], a[1], a[2], a[3], 1e-3, 0.9, 0.999, 1e-6, 0)
    cuda.adam_step_cuda(b[0], b[1], b[2], b[3], 1e-3, 0.9, 0.999, 1e-6, 0)
    torch.cuda.synchronize()
    for x, y in zip(a, b):
        assert torch.allclose(x, y, rtol=1e-6, atol=1e-9)


def test_vdbadam_wrapper():
    load_pkg()
    from vdbAdam import vdbAdam
    dev = "cuda:0"
    w = torch.nn.Parameter(torch.ones(64, 8, device=dev))
    opt = vdbAdam([w], lr=1e-2)
    (w[:10] ** 2).sum().backward()
    before = w.detach().clone()
    opt.step()
    assert not torch.equal(before[:10], w.detach()[:10]) and torch.equal(before[10:], w.detach()[10:])
    opt.zero_grad()
    assert float(w.grad.abs().max()) == 0.0
