import sys, os, importlib, importlib.util, math, tempfile, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
import bench, scenes
pkg = importlib.import_module(bench.PKG); pkg.install()
from hashgrid import INFERENCE
from tile_step import TileStep
import test_psnr_gpu as T
DEV = torch.device("cuda:0")
H, W, n_cam, S, log2T, steps = 48, 64, 8, 32, 15, 300
gen = torch.Generator().manual_seed(0)
Ks, c2w = scenes.camera_rig(n_cam, H, W, gen, center=(10.37, 3.21, 15.53), radius=4.83, fx=60.0)
ply = os.path.join(tempfile.mkdtemp(), "mesh.ply")
scenes.write_proxy_mesh_ply(ply, (0, 0, 0), (20, 13, 30), seed=0, ground_res=16, n_boxes=6)
batches = []
for _ in range(steps + 4):
    locs = torch.stack([torch.randint(0, n_cam, (512,), generator=gen), torch.randint(0, W, (512,), generator=gen), torch.randint(0, H, (512,), generator=gen)], -1).int()
    batches.append((locs.to(DEV), T._target(locs, H, W).to(DEV)))
held, train = batches[-4:], batches[:-4]
spec = importlib.util.spec_from_file_location("ref_cuda_step", os.path.join(ROOT, "tools", "ref_cuda_step.py"))
rcs = importlib.util.module_from_spec(spec); spec.loader.exec_module(rcs)
def build(variant):
    torch.manual_seed(0)
    if variant.startswith("ref"):
        st = rcs.build_reference_step(DEV, (0.0, 0.0, 0.0), (20.0, 13.0, 30.0), Ks, c2w, log2T, (16, 512), 4, S, S, ply, global_step=6000)
        if "nopose" in variant:
            st.optimizer = torch.optim.Adam([{"params": st.decoder.parameters(), "lr": 1e-3, "weight_decay": 1e-6}])
        if "ourposes" in variant:
            import tile_step as ts
            st.poses = ts.Poses(Ks, c2w, DEV, None)
            st.optimizer = torch.optim.Adam([{"params": st.decoder.parameters(), "lr": 1e-3, "weight_decay": 1e-6}, {"params": st.poses.se3_refine, "lr": 1e-4}])
        return st
    st = TileStep(DEV, (0.0, 0.0, 0.0), (20.0, 13.0, 30.0), Ks, c2w, log2_hashmap_size=log2T, grid_resolution=(16, 512), num_sample=S, num_bg_sample=S,
                  mesh_path=ply, global_step=6000, dense_table_adam=True)
    if "nodec" in variant: st.featureGrid.fused_decoder = False
    if "noenc" in variant: st.featureGrid.fused_encode = False; st.featureGrid.fused_decoder = False
    if "foreach" in variant:
        st.optimizer = torch.optim.Adam([{"params": st.decoder.parameters(), "lr": 1e-3, "weight_decay": 1e-6}, {"params": st.poses.se3_refine, "lr": 1e-4}])
    if "nopose" in variant:
        st.optimizer = torch.optim.Adam([{"params": st.decoder.parameters(), "lr": 1e-3, "weight_decay": 1e-6}], fused=True)
    if "round" in variant:
        # emulate, on the fp32 torch decoder, the one uncompensated piece of the tensor-core backward: the weight-gradient
        # GEMM reads its activation operand as a single bf16 value (layers selected by digits after "round": 1..8 in
        # module order; "roundall" = every Linear).  "affine": inputs that are Gaussian activations in (0, 1] are stored
        # as bf16(a - 0.5) instead (the 0.5 folds into the bias / bias gradient exactly).
        st.featureGrid.fused_decoder = False
        class RoundedLinearFn(torch.autograd.Function):
            @staticmethod
            def forward(ctx, x, w, b, shift):
                ctx.save_for_backward(x, w)
                ctx.shift = shift
                return torch.nn.functional.linear(x, w, b)
            @staticmethod
            def backward(ctx, gy):
                x, w = ctx.saved_tensors
                xr = (x - ctx.shift).to(torch.bfloat16).float() + ctx.shift
                g2, x2 = gy.reshape(-1, gy.shape[-1]), xr.reshape(-1, xr.shape[-1])
                return gy @ w, g2.t() @ x2, g2.sum(0), None
        lins = [m for m in st.decoder.modules() if isinstance(m, torch.nn.Linear)]
        sel = range(len(lins)) if "roundall" in variant else [int(c) - 1 for c in variant.split("round")[1] if c.isdigit()]
        for i in sel:
            lin = lins[i]
            shift = 0.5 if ("affine" in variant and i in (1, 6, 7)) else 0.0
            lin.forward = (lambda x, lin=lin, shift=shift: RoundedLinearFn.apply(x, lin.weight, lin.bias, shift))
    if "plainbf16" in variant:
        import scanerf_b200_capi as capi
        capi.lib().snrf_decoder_set_precision(capi.c_int(0))
    return st
def psnr(st):
    se = 0.0
    with torch.no_grad():
        for locs, gt in held:
            o, d = st.poses.rays(locs)
            out, _ = st.render_rays(o, d, None, INFERENCE)
            se += float(torch.mean((out["pred_color"] - gt) ** 2))
    return -10.0 * math.log10(se / len(held))
for variant in sys.argv[1:]:
    vals = []
    for rep in range(2):
        st = build(variant)
        for l, g in train:
            st.step_device(l, g)
        vals.append(round(psnr(st), 3))
        del st
    import scanerf_b200_capi as capi
    capi.lib().snrf_decoder_set_precision(capi.c_int(1))
    print(variant, vals, flush=True)
