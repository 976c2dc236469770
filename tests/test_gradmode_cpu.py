"""CPU: hashgrid._gradmode.shared_table_grad -- the dense gradient table that the encode backwards of ONE backward pass share
(the reference's drivers encode the same table for the foreground and the background chain, tile.py:661-681; its own Functions
return one dense grad_features each, hashgrid/PyHashGridBG.py:20-30).  A stand-in Function with the same backward protocol
(accumulate into the tensor, return it only when it is new) is checked against plain autograd in every situation the sharing
has to survive: two producers in a pass, passes without zero_grad, retain_graph, autograd.grad, hooks, a user holding .grad,
a non-leaf table, create_graph."""
import torch

from conftest import load_pkg


def _fn():
    load_pkg()
    from hashgrid import _gradmode

    class Scale(torch.autograd.Function):
        """y = k * table; the backward ACCUMULATES k * g into the shared tensor the way the scatter kernels do (raw writes:
        the tensor's version counter does not move)."""
        calls = []

        @staticmethod
        def forward(ctx, table, k):
            ctx.k, ctx.table = k, table
            return table.detach() * k

        @staticmethod
        def backward(ctx, g):
            gt, first = _gradmode.shared_table_grad(ctx.table)
            Scale.calls.append(first)
            gt.data.add_(ctx.k * g)
            return (gt if first else None), None
    return _gradmode, Scale


def _plain(table, ks, w):
    return sum(((table * k) * w).sum() for k in ks)


def test_two_producers_share_one_table_and_sum_exactly():
    gm, Scale = _fn()
    torch.manual_seed(0)
    w = torch.randn(4, 5)
    t = torch.randn(4, 5, requires_grad=True)
    ref = torch.randn(4, 5, requires_grad=True)
    with torch.no_grad():
        ref.copy_(t)
    Scale.calls.clear()
    ((Scale.apply(t, 2.0) * w).sum() + (Scale.apply(t, 3.0) * w).sum() + (Scale.apply(t, -1.0) * w).sum()).backward()
    _plain(ref, (2.0, 3.0, -1.0), w).backward()
    assert Scale.calls == [True, False, False]
    assert torch.allclose(t.grad, ref.grad, rtol=0, atol=1e-6)


def test_passes_without_zero_grad_accumulate_and_do_not_share_across_passes():
    gm, Scale = _fn()
    w = torch.ones(3)
    t = torch.zeros(3, requires_grad=True)
    Scale.calls.clear()
    for _ in range(3):
        ((Scale.apply(t, 2.0) * w).sum() + (Scale.apply(t, 5.0) * w).sum()).backward()
    assert Scale.calls == [True, False] * 3          # a new table per backward pass (graph task)
    assert torch.equal(t.grad, torch.full((3,), 21.0))


def test_user_reference_to_grad_is_never_written_by_a_later_pass():
    gm, Scale = _fn()
    t = torch.zeros(3, requires_grad=True)
    (Scale.apply(t, 2.0).sum() + Scale.apply(t, 3.0).sum()).backward()
    kept = t.grad
    snapshot = kept.clone()
    t.grad = None
    (Scale.apply(t, 7.0).sum() + Scale.apply(t, 1.0).sum()).backward()
    assert torch.equal(kept, snapshot)
    assert torch.equal(t.grad, torch.full((3,), 8.0))


def test_retain_graph_autograd_grad_and_hooks():
    gm, Scale = _fn()
    t = torch.zeros(3, requires_grad=True)
    seen = []
    t.register_hook(lambda g: seen.append(g.clone()))
    loss = Scale.apply(t, 2.0).sum() + Scale.apply(t, 3.0).sum()
    (g,) = torch.autograd.grad(loss, t, retain_graph=True)
    assert torch.equal(g, torch.full((3,), 5.0)) and t.grad is None
    loss.backward()
    assert torch.equal(t.grad, torch.full((3,), 5.0))
    assert len(seen) == 2 and all(torch.equal(s, torch.full((3,), 5.0)) for s in seen)      # the hook saw the complete sum
    assert torch.equal(g, torch.full((3,), 5.0))                                             # the first result was not written again


def test_non_leaf_table_and_disabled_switch_fall_back_to_independent_tables():
    gm, Scale = _fn()
    base = torch.ones(3, requires_grad=True)
    t = base * 1.5                                    # not a leaf: no sharing
    Scale.calls.clear()
    (Scale.apply(t, 2.0).sum() + Scale.apply(t, 3.0).sum()).backward()
    assert Scale.calls == [True, True]
    assert torch.equal(base.grad, torch.full((3,), 7.5))
    leaf = torch.zeros(3, requires_grad=True)
    gm.share_enabled = False
    try:
        Scale.calls.clear()
        (Scale.apply(leaf, 2.0).sum() + Scale.apply(leaf, 3.0).sum()).backward()
        assert Scale.calls == [True, True]
        assert torch.equal(leaf.grad, torch.full((3,), 5.0))
    finally:
        gm.share_enabled = True


def test_two_tables_in_one_pass_and_outside_of_a_backward():
    gm, Scale = _fn()
    a = torch.zeros(3, requires_grad=True)
    b = torch.zeros(3, requires_grad=True)
    Scale.calls.clear()
    (Scale.apply(a, 2.0).sum() + Scale.apply(b, 3.0).sum() + Scale.apply(a, 4.0).sum() + Scale.apply(b, 5.0).sum()).backward()
    assert torch.equal(a.grad, torch.full((3,), 6.0)) and torch.equal(b.grad, torch.full((3,), 8.0))
    # called outside of a backward pass (no graph task): always a fresh table
    g1, f1 = gm.shared_table_grad(a)
    g2, f2 = gm.shared_table_grad(a)
    assert f1 and f2 and g1 is not g2


def test_create_graph_does_not_share():
    gm, Scale = _fn()
    t = torch.zeros(3, requires_grad=True)
    Scale.calls.clear()
    (g,) = torch.autograd.grad(Scale.apply(t, 2.0).sum() + Scale.apply(t, 3.0).sum(), t, create_graph=True)
    assert Scale.calls == [True, True] and torch.equal(g.detach(), torch.full((3,), 5.0))
