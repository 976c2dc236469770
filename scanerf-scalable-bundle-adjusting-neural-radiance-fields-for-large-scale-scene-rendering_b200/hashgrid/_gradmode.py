"""Where the hash-table gradient of an encode backward goes.

Default (no context active): the encode Functions behave like any autograd Function and RETURN a dense
`grad_features` tensor, exactly as the reference's do (hashgrid/PyHashGridBG.py:20-30) -- `.grad` is only ever
touched by autograd itself, so `torch.autograd.grad(...)`, normal / validation passes etc. see no side effect.

The training step opts in explicitly, around its `loss.backward()`:

    with table_backward("direct"):            # scatter straight into features.grad (no 2 GiB dense temporary)
    with table_backward("fused", optimizer):  # scatter + sparse Adam in one pass (snrf_field_encode_bwd_adam);
                                              # the table gradient never exists in HBM

`vdbAdam.table_backward(fused=...)` is the usual way in.
"""
import contextlib

_mode = None        # None | "direct" | "fused"
_optimizer = None   # the vdbAdam of the "fused" mode


def mode():
    return _mode


def optimizer():
    return _optimizer


@contextlib.contextmanager
def table_backward(new_mode, opt=None):
    global _mode, _optimizer
    if new_mode not in ("direct", "fused"):
        raise ValueError(f"table_backward: unknown mode {new_mode!r}")
    if new_mode == "fused" and opt is None:
        raise ValueError("table_backward('fused') needs the vdbAdam that owns the table")
    prev = (_mode, _optimizer)
    _mode, _optimizer = new_mode, opt
    try:
        yield
    finally:
        _mode, _optimizer = prev
