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


# ---- default mode: one dense gradient table per backward pass, shared by the encodes of that pass ------------------------
#
# The reference's drivers encode the same table twice per step (foreground and background chain, tile.py:661-681).  With two
# independent dense `grad_features` tensors autograd pays, per step at T = 2^24: two 2 GiB zero fills and one 2 GiB + 2 GiB
# -> 2 GiB add (1.7 ms of 21.7 on a B200, profiles/r5a_tile_py_c2_dropin_kernels.json).  A consumer of a tensor's gradient
# (AccumulateGrad, a captured `autograd.grad` input, a tensor hook) runs only after ALL producers of that gradient in the
# running graph task have returned, and the first contribution is held by reference, not copied.  So within ONE graph task
# the later encodes of the SAME leaf table may scatter into the tensor the first one returned and contribute `None`
# themselves: the sum autograd forms is the same, the extra fill and the add are gone.  Every condition that makes this exact is
# checked; anything else falls back to a fresh zero-filled tensor.

import weakref

import torch

share_enabled = True
_shared = None      # (graph task id, weakref(features), weakref(grad table), its _version)


def shared_table_grad(features):
    """-> (g_table, first).  `first`: g_table is a new zero-filled tensor to RETURN as grad_features; otherwise g_table is the
    tensor an earlier encode backward of this graph task returned for the same leaf: accumulate into it and return None."""
    global _shared
    tid = torch._C._current_graph_task_id() if hasattr(torch._C, "_current_graph_task_id") else -1
    ok = share_enabled and tid != -1 and features.is_leaf and not torch.is_grad_enabled()
    if ok and _shared is not None and _shared[0] == tid and _shared[1]() is features:
        g = _shared[2]()
        if g is not None and g._version == _shared[3] and g.shape == features.shape and g.device == features.device:
            return g, False
    g = torch.zeros_like(features)
    _shared = (tid, weakref.ref(features), weakref.ref(g), g._version) if ok else None
    return g, True
