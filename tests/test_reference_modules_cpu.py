"""CPU, build container only (skipped where /root/reference is absent, e.g. on the GPU box): the drop-in's host
logic against objects of the REFERENCE's own Python classes.
  * hashgrid._decoder.decoder_params() must read the 16 parameter tensors off the reference's network.ShallowMLP
    (that is the decoder tile.py passes into HashGrid.render_*_rays), in state_dict order;
  * hashgrid._decoder.flatten_for_inference() must equal the flat vector rendering.py:101-113 builds;
  * the mirror ShallowMLP has the reference's state_dict keys and shapes (decoder.pth interchange)."""
import importlib.util
import os
import sys
import types

import pytest
import torch

from conftest import load_pkg

REF = os.environ.get("SCANERF_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "network.py")), reason="reference checkout not present")


def _reference_network():
    class _EasyDict(dict):
        __getattr__ = dict.__getitem__
        __setattr__ = dict.__setitem__
    if "easydict" not in sys.modules:
        m = types.ModuleType("easydict")
        m.EasyDict = _EasyDict
        sys.modules["easydict"] = m
    spec = importlib.util.spec_from_file_location("_ref_network", os.path.join(REF, "network.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_decoder_params_reads_the_reference_module():
    load_pkg()
    from hashgrid import _decoder
    net = _reference_network()
    torch.manual_seed(0)
    ref = net.ShallowMLP(32)
    ps = _decoder.decoder_params(ref)
    assert ps is not None and len(ps) == 16
    sd = ref.state_dict()
    keys = list(sd.keys())
    assert len(keys) == 16
    for p, k in zip(ps, keys):                      # state_dict order = (weight, bias) per Linear, LAYERS order
        assert p.data_ptr() == sd[k].data_ptr(), k
    mine = _decoder.ShallowMLP(32)
    assert list(mine.state_dict().keys()) == keys
    assert [tuple(v.shape) for v in mine.state_dict().values()] == [tuple(v.shape) for v in sd.values()]
    mine.load_state_dict(sd)                        # decoder.pth written by either side loads in the other
    x = torch.randn(5, 7, 35)
    w = torch.rand(32)
    a, b = ref(x, weight_feature=w), mine(x, weight_feature=w)
    for k in ("sigma", "tint", "diffuse", "specular"):
        assert torch.allclose(a[k], b[k], atol=1e-6), k
    # the flat inference layout of rendering.py:101-113: per Linear, bias then W^T flattened
    flat = torch.cat([t for w_, b_ in zip([sd[k] for k in keys[0::2]], [sd[k] for k in keys[1::2]]) for t in (b_, w_.transpose(1, 0).flatten())])
    assert flat.numel() == 13994
    assert torch.equal(_decoder.flatten_for_inference(ref), flat)
    assert _decoder.decoder_params(torch.nn.Linear(3, 3)) is None
