"""Time the decoder kernels alone (2.1 M samples) in split / plain precision."""
import sys, os, importlib, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
pkg = importlib.import_module(bench.PKG); pkg.install()
from hashgrid import _field
from hashgrid._decoder import ShallowMLP, decoder_params
dev = "cuda:0"
R, S = 16384, 128
N = R * S
torch.manual_seed(0)
dec = ShallowMLP(32).to(dev)
params = [p.detach().requires_grad_(True) for p in decoder_params(dec)]
feats = (torch.randn(16, N, 2, device=dev) * 0.3).requires_grad_(True)
rays_d = torch.randn(R, 3, device=dev).requires_grad_(True)
mask = torch.ones(32, device=dev)
cot = torch.randn(N, 10, device=dev)
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
import scanerf_b200_capi as capi
for split in (True, False, True, False):
    _field.set_precision(split)
    fwd = t(lambda: _field.decoder_forward(feats.detach(), mask, rays_d.detach(), S, params))
    def fb():
        h = _field.decoder_apply(feats, rays_d, mask, S, params)
        h.backward(cot)
    both = t(fb)
    capi.time_calls("snrf_decoder_bwd"); fb(); fb(); ms, _ = capi.timed_results(); capi.time_calls(None)
    print("   bwd kernel ms (events around the C call):", [round(m, 3) for m in ms])
    print(f"split={split}: fwd {fwd:.3f} ms, fwd+bwd {both:.3f} ms -> bwd ~{both - fwd:.3f} ms  ({N} samples)")
_field.set_precision(True)
