#!/usr/bin/env python
"""Build the UNMODIFIED reference CUDA extensions for sm_100a into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported by the product
package; only tests/, __graft_entry__.smoke() and bench.py's baseline legs may
load what this script produces.

The reference sources are compiled *where they lie* under /root/reference (no
copy into this repo).  The reference's own setup.py files cannot be used: they
import matplotlib, link unused OpenCV libraries and hard-code cpython-38 names
(SURVEY.md section 8c).  This recipe calls nvcc / g++ directly with nvcc's
defaults for numerics (no fast-math, -fmad=true), exactly like the reference's
CUDAExtension build would.

Targets (python oracle/build_ref.py [target ...] [-j N]):
  embed     HASHGRID_EMBED  hashgrid/src/hashgrid{,_bg}_kernel.cu + our 20-line
                            pybind glue (oracle/ref_glue/hashgrid_embed_binding.cpp)
  cuda      CUDA_EXT        cuda/*.cu + cnpy.cpp + the reference's own binding.cpp
  fastmesh  fastMesh        fastMesh/src/fastMesh_kernel.cu + reference binding.cpp
  hashgrid  HASHGRID        the full reference hashgrid module incl. rendering_kernel.cu
                            (about 40 minutes of single-threaded ptxas time)
  drivers   ref_drivers.zip the reference's Python side (tile.py, admm_trainer.py, rendering.py, camera*.py, network.py,
                            losses, tools/, config/*.yaml and the __init__.py / PyHashGrid*.py wrappers of its three
                            extension packages), zipped UNMODIFIED where it lies into one build artefact that
                            tests/ref_driver_harness.py puts on sys.path (zipimport), so that the reference's own drivers
                            can run on the GPU box -- on top of the drop-in, or on top of the rebuilt reference modules
Outputs: oracle/_ref/<MODULE>.so (+ objects in oracle/_ref/obj/), oracle/_ref/ref_drivers.zip.  oracle/_ref/ is
git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import argparse
import concurrent.futures as cf
import os
import subprocess
import sys
import sysconfig
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SCANERF_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
OBJ = os.path.join(OUT, "obj")


def _torch_paths():
    import torch  # noqa: F401
    from torch.utils import cpp_extension as ce
    inc = ce.include_paths()
    lib = ce.library_paths()
    return inc, lib


def _common_flags(module, incs):
    inc, _ = _torch_paths()
    flags = []
    for i in incs + inc + [sysconfig.get_paths()["include"]]:
        flags += ["-I", i]
    flags += [
        "-DTORCH_API_INCLUDE_EXTENSION_H",
        f"-DTORCH_EXTENSION_NAME={module}",
        "-D_GLIBCXX_USE_CXX11_ABI=1",
    ]
    return flags


def _compile(src, obj, module, incs):
    if os.path.exists(obj) and os.path.getmtime(obj) > os.path.getmtime(src):
        return obj, 0.0, "cached"
    t0 = time.time()
    if src.endswith(".cu"):
        cmd = ["nvcc", "-c", src, "-o", obj, "-std=c++17",
               "-gencode", "arch=compute_100a,code=sm_100a",
               "--expt-relaxed-constexpr", "-w",
               "-Xcompiler", "-fPIC"] + _common_flags(module, incs)
    else:
        cmd = ["g++", "-c", src, "-o", obj, "-std=c++17", "-fPIC", "-O1", "-w"] + \
              _common_flags(module, incs) + ["-I", "/usr/local/cuda/include"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(f"FAILED {src}\n{r.stderr[-4000:]}\n")
        raise SystemExit(1)
    return obj, time.time() - t0, "built"


def _link(objs, module):
    _, lib = _torch_paths()
    so = os.path.join(OUT, f"{module}.so")
    cmd = ["g++", "-shared", "-o", so] + objs
    for l in lib + ["/usr/local/cuda/lib64"]:
        cmd += ["-L", l, f"-Wl,-rpath,{l}"]
    cmd += ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch",
            "-ltorch_python", "-lcudart", "-lz"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stderr[-4000:])
        raise SystemExit(1)
    return so


TARGETS = {
    "embed": dict(
        module="HASHGRID_EMBED",
        incs=[f"{REF}/hashgrid/include"],
        srcs=[f"{REF}/hashgrid/src/hashgrid_kernel.cu",
              f"{REF}/hashgrid/src/hashgrid_bg_kernel.cu",
              f"{HERE}/ref_glue/hashgrid_embed_binding.cpp"]),
    "cuda": dict(
        module="CUDA_EXT",
        incs=[f"{REF}/cuda/include"],
        srcs=[f"{REF}/cuda/{n}" for n in (
            "adam_kernel.cu", "sample_kernel.cu", "helper_kernel.cu",
            "compute_ray_kernel.cu", "grid_sample_kernel.cu",
            "view_selection_kernel.cu", "build_blocks_kernel.cu",
            "cnpy.cpp", "binding.cpp")]),
    "fastmesh": dict(
        module="fastMesh",
        incs=[f"{REF}/fastMesh/include"],
        srcs=[f"{REF}/fastMesh/src/fastMesh_kernel.cu",
              f"{REF}/fastMesh/binding.cpp"]),
    "hashgrid": dict(
        module="HASHGRID",
        incs=[f"{REF}/hashgrid/include"],
        srcs=[f"{REF}/hashgrid/src/rendering_kernel.cu",
              f"{REF}/hashgrid/src/hashgrid_kernel.cu",
              f"{REF}/hashgrid/src/hashgrid_bg_kernel.cu",
              f"{REF}/hashgrid/src/sampler_kernel.cu",
              f"{REF}/hashgrid/src/rendering/renderbase_kernel.cu",
              f"{REF}/hashgrid/binding.cpp"]),
}


def build_drivers():
    """oracle/_ref/ref_drivers.zip: the reference's Python files, byte for byte, plus three empty generated
    `<pkg>/lib/__init__.py` so that `hashgrid.lib` / `cuda.lib` / `fastMesh.lib` exist as packages (the reference's make.sh
    copies the built .so files there; the harness maps those module names to oracle/_ref/*.so)."""
    import glob
    import zipfile
    out = os.path.join(OUT, "ref_drivers.zip")
    files = sorted(glob.glob(os.path.join(REF, "*.py")) + glob.glob(os.path.join(REF, "tools", "*.py")) +
                   [os.path.join(REF, "preprocess", "build_tiles.py")] +
                   glob.glob(os.path.join(REF, "config", "*.yaml")) +
                   [os.path.join(REF, "hashgrid", n) for n in ("__init__.py", "PyHashGrid.py", "PyHashGridBG.py")] +
                   [os.path.join(REF, "cuda", "__init__.py"), os.path.join(REF, "fastMesh", "__init__.py")])
    with zipfile.ZipFile(out, "w", zipfile.ZIP_DEFLATED) as z:
        for f in files:
            z.write(f, os.path.relpath(f, REF))
        for pkg in ("hashgrid", "cuda", "fastMesh"):
            z.writestr(f"{pkg}/lib/__init__.py", "")
    print(f"[oracle/build_ref] wrote {out} ({len(files)} reference files)", flush=True)


def build(names, jobs):
    if not os.path.isdir(REF):
        print(f"[oracle/build_ref] {REF} not present - keeping prebuilt oracle/_ref as is")
        return
    os.makedirs(OBJ, exist_ok=True)
    if "drivers" in names:
        build_drivers()
        names = [n for n in names if n != "drivers"]
    work = []
    for n in names:
        t = TARGETS[n]
        for s in t["srcs"]:
            o = os.path.join(OBJ, f"{t['module']}__{os.path.basename(s)}.o")
            work.append((n, s, o))
    done = {}
    with cf.ThreadPoolExecutor(max_workers=jobs) as ex:
        futs = {ex.submit(_compile, s, o, TARGETS[n]["module"], TARGETS[n]["incs"]): (n, s)
                for n, s, o in work}
        for f in cf.as_completed(futs):
            n, s = futs[f]
            obj, dt, how = f.result()
            done.setdefault(n, []).append(obj)
            print(f"[oracle/build_ref] {n}: {os.path.basename(s)} {how} {dt:.0f}s", flush=True)
    for n in names:
        so = _link(sorted(done[n]), TARGETS[n]["module"])
        print(f"[oracle/build_ref] linked {so}", flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("targets", nargs="*", default=["embed", "cuda", "fastmesh"])
    ap.add_argument("-j", type=int, default=4)
    a = ap.parse_args()
    build(a.targets, a.j)
