"""Python face of the fastMesh extension module (reference: fastMesh/binding.cpp:7-18): class
`fastMesh` with build / getSceneBound / destroy / sample_points / fisrtHit / firstEnter (the
misspelling is API).  Every query forwards raw device pointers to libscanerf_b200.so."""
import ctypes

import torch

import scanerf_b200_capi as capi
from scanerf_b200_capi import c_int, c_void_p, inp, Out, ptr

f32, i32 = torch.float32, torch.int32


class fastMesh:
    def __init__(self):
        self._h = None

    def build(self, model_path):
        """fastMesh.h:22-26 -- read the PLY, build the 64^3 grid, upload it to the current device."""
        self.destroy()
        h = c_void_p(0)
        capi.check(capi.lib().snrf_mesh_create(ctypes.c_char_p(str(model_path).encode()), ctypes.byref(h)), "snrf_mesh_create")
        self._h = h

    def build_from_arrays(self, verts, faces):
        """Extension: same from CPU tensors verts [V,3] f32, faces [F,3] i32."""
        self.destroy()
        v = verts.detach().cpu().to(f32).contiguous()
        f = faces.detach().cpu().to(i32).contiguous()
        h = c_void_p(0)
        capi.check(capi.lib().snrf_mesh_create_from_arrays(ptr(v), c_int(v.shape[0]), ptr(f), c_int(f.shape[0]), ctypes.byref(h)),
                   "snrf_mesh_create_from_arrays")
        self._h = h

    def _handle(self):
        if self._h is None:
            raise RuntimeError("fastMesh: build() has not been called")
        return self._h

    def getSceneBound(self):
        """fastMesh.h:28-38 -- CPU float32 [6] = min xyz, max xyz of the vertices."""
        out = (ctypes.c_float * 6)()
        capi.check(capi.lib().snrf_mesh_bounds(self._handle(), out), "snrf_mesh_bounds")
        return torch.tensor(list(out), dtype=f32)

    def stats(self):
        out = (ctypes.c_longlong * 3)()
        capi.check(capi.lib().snrf_mesh_stats(self._handle(), out), "snrf_mesh_stats")
        return {"occupied_cells": out[0], "list_entries": out[1], "faces": out[2]}

    def destroy(self):
        if self._h is not None:
            capi.lib().snrf_mesh_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def fisrtHit(self, rays_o, rays_d, z_depth, hit_face=None):
        B = int(rays_o.shape[0])
        o, d = inp(rays_o, f32, "rays_o"), inp(rays_d, f32, "rays_d")
        z = Out(z_depth, f32, "z_depth")
        hf = Out(hit_face, i32, "hit_face") if hit_face is not None else None
        capi.check(capi.lib().snrf_mesh_first_hit(self._handle(), ptr(o), ptr(d), z.ptr, hf.ptr if hf else c_void_p(0),
                                                  c_int(B), capi.stream()), "snrf_mesh_first_hit")
        z.done()
        if hf:
            hf.done()

    def firstEnter(self, rays_o, rays_d, z_depth):
        B = int(rays_o.shape[0])
        o, d = inp(rays_o, f32, "rays_o"), inp(rays_d, f32, "rays_d")
        z = Out(z_depth, f32, "z_depth")
        capi.check(capi.lib().snrf_mesh_first_enter(self._handle(), ptr(o), ptr(d), z.ptr, c_int(B), capi.stream()),
                   "snrf_mesh_first_enter")
        z.done()

    def sample_points(self, rays_o, rays_d, t_start, z_vals):
        B, S = int(rays_o.shape[0]), int(z_vals.shape[1])
        o, d, t = inp(rays_o, f32, "rays_o"), inp(rays_d, f32, "rays_d"), inp(t_start, f32, "t_start")
        z = Out(z_vals, f32, "z_vals")
        capi.check(capi.lib().snrf_mesh_sample(self._handle(), ptr(o), ptr(d), ptr(t), z.ptr, c_int(B), c_int(S), capi.stream()),
                   "snrf_mesh_sample")
        z.done()
