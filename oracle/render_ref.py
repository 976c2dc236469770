"""TEST INFRASTRUCTURE ONLY (imported by tests/, never by the product path).

numpy restatement of the reference's multi-tile render stages and of the fastMesh queries -- small,
loop-based (seconds for a few hundred rays), each function citing the reference it follows:
  ray_block_intersection   hashgrid/src/rendering_kernel.cu:126-153  (RayAABBIntersection: hashgrid/include/cuda_utils.h:564-613)
  prepare_points           :391-431
  accumulate_color         :623-674
  inverse_z_sampling       :815-851  (inverse_z_sample_bound: cuda_utils.h:61-75)
  update_outgoing_bidx     :1262-1373
  get_last_block           :1211-1241
  fastMesh build           fastMesh/include/tile_structure.h:13-106
  fastMesh fisrtHit        fastMesh/src/fastMesh_kernel.cu:230-329 (+ dda.h:7-131, cuda_utils.h:403-436)
Parity: pinned on the GPU box against the rebuilt reference extensions (oracle/_ref/HASHGRID.so,
fastMesh.so) by tests/test_render_gpu.py / tests/test_mesh_gpu.py; the reference holds no golden
vectors for these ops.  The sampler / fused evaluation stages are covered by oracle/native.py
(occupancy walk, hash encode) and oracle/torch_ref.py (decoder).
"""
import numpy as np

f32 = np.float32
MISS = f32(1e7)


def _safe_div(a, b):
    return f32(a) / f32(b) if b != 0 else f32(1e8)


def ray_aabb(o, d, center, half):
    """cuda_utils.h:564-613 -> (near, far) or (-1, -1); near clamps at 0, far starts at 1e5."""
    lo_acc, hi_acc = f32(0), f32(100000.0)
    for a in range(3):
        inv = _safe_div(1.0, d[a])
        lo, hi = f32(f32(center[a] - half[a] - o[a]) * inv), f32(f32(center[a] + half[a] - o[a]) * inv)
        if hi < lo:
            lo, hi = hi, lo
        if hi < lo_acc or lo > hi_acc:
            return f32(-1), f32(-1)
        lo_acc, hi_acc = max(lo_acc, lo), min(hi_acc, hi)
        if lo_acc > hi_acc:
            return f32(-1), f32(-1)
    return lo_acc, hi_acc


def ray_block_intersection(rays_o, rays_d, corners, sizes):
    B, nb = rays_o.shape[0], corners.shape[0]
    out = np.full((B, nb, 2), MISS, f32)
    for b in range(nb):
        half = (sizes[b] * f32(0.5)).astype(f32)
        center = (corners[b] + half).astype(f32)
        for i in range(B):
            n, fr = ray_aabb(rays_o[i], rays_d[i], center, half)
            if n != -1:
                out[i, b] = (n, fr)
    return out


def prepare_points(z_vals, running, intersections):
    B, S = z_vals.shape
    nb = intersections.shape[1]
    out = np.full((B, S, 4), -1, np.int16)
    for i in range(B):
        if not running[i]:
            continue
        for k in range(S):
            z = z_vals[i, k]
            if z == -1:
                continue
            idx = 0
            for b in range(nb):
                if intersections[i, b, 0] <= z <= intersections[i, b, 1] and idx < 4:
                    out[i, k, idx] = b
                    idx += 1
    return out


def accumulate_color(pts_diffuse, pts_specular, pts_alpha, transparency, z_vals, diffuse, specular, depth):
    T, dif, spe, dep = transparency.copy(), diffuse.copy(), specular.copy(), depth.copy()
    B, S = z_vals.shape
    for i in range(B):
        t = f32(T[i, 0])
        if t < 1e-5:
            continue
        for k in range(S):
            a = f32(pts_alpha[i, k, 0])
            dif[i] += t * pts_diffuse[i, k]
            spe[i] += t * pts_specular[i, k]
            dep[i, 0] += t * a * z_vals[i, k]
            t = f32(t * (1 - a))
        T[i, 0] = t
    return T, dif, spe, dep


def inverse_z_sampling(intersections, related_bidx, S, sample_range, fill=-1.0):
    B = intersections.shape[0]
    z = np.full((B, S), fill, f32)
    for i in range(B):
        b = int(related_bidx[i])
        if b == -1 or intersections[i, b, 0] == MISS:
            continue
        near = f32(intersections[i, b, 1])
        far = f32(near + f32(sample_range))
        inv_near, inv_far = f32(1) / near, f32(1) / far
        inv_bound, step = f32(inv_far - inv_near), f32(1) / f32(S - 1)
        for k in range(S):
            z[i, k] = f32(1) / f32(f32(f32(step * f32(k)) * inv_bound) + inv_near)
    return z


def get_last_block(tracing_blocks, intersections):
    B, nb = tracing_blocks.shape
    out = np.full(B, -1, np.int32)
    for i in range(B):
        for s in range(nb):
            b = tracing_blocks[i, s]
            if intersections[i, b, 0] == MISS:
                break
            out[i] = b
    return out


def update_outgoing_bidx(rays_o, rays_d, corners, sizes, tracing_blocks, intersections, skip=False):
    B, nb = tracing_blocks.shape
    ids, w = np.full((B, 4), -1, np.int16), np.zeros((B, 4), f32)
    for i in range(B):
        far, cur = f32(-1), []
        for s in range(nb):
            b = int(tracing_blocks[i, s])
            n, fr = intersections[i, b]
            if n == MISS:
                break
            if not skip and (n > far and far != -1):
                break
            if fr > far:
                far, cur = fr, [b]
            elif fr == far and len(cur) < 4:
                cur.append(b)
        if far == -1:
            continue
        if len(cur) == 1:
            ids[i, 0], w[i, 0] = cur[0], 1.0
            continue
        p = (rays_o[i] + far * rays_d[i]).astype(f32)
        for k, b in enumerate(cur):
            q = np.clip((p - corners[b]) / sizes[b], 0, 1).astype(f32)
            dis = ((f32(0.5) - np.abs(q - f32(0.5))) * sizes[b]).astype(f32)
            if dis[0] != 0 and dis[2] != 0:
                w[i, k] = dis[0] * dis[2]
            elif dis[0] != 0:
                w[i, k] = dis[0]
            elif dis[2] != 0:
                w[i, k] = dis[2]
            ids[i, k] = b
    return ids, w


# ------------------------------------------------------------------------------------------------ fastMesh
def mesh_build(verts, faces, side=64):
    """tile_structure.h:13-106: cubic grid anchored at the vertex AABB minimum, cell = max extent / 64; a face is
    listed in every cell its AABB overlaps.  Returns (origin, cell, {cell index -> [face ids]})."""
    verts = np.asarray(verts, f32)
    mn, mx = verts.min(0), verts.max(0)
    cell = f32((mx - mn).max() / f32(side))
    inv = f32(1) / cell
    lists = {}
    smax = mn + f32(side) * cell
    for fi, f in enumerate(np.asarray(faces)):
        tri = verts[f]
        lo, hi = tri.min(0), tri.max(0)
        ctr, half = (lo + hi) * f32(0.5), (hi - lo) * f32(0.5)
        lo, hi = ctr - half, ctr + half
        if (hi <= mn).any() or (lo >= smax).any():
            continue
        i0 = np.clip(((lo - mn) * inv).astype(np.int64), 0, side - 1)
        i1 = np.clip(((hi - mn) * inv).astype(np.int64), 0, side - 1)
        for x in range(i0[0], i1[0] + 1):
            for y in range(i0[1], i1[1] + 1):
                for z in range(i0[2], i1[2] + 1):
                    lists.setdefault((x << 12) | (y << 6) | z, []).append(fi)
    return mn.astype(f32), cell, lists


def _ray_triangle(o, d, v0, v1, v2):
    """cuda_utils.h:403-436: Moeller-Trumbore with back-face rejection (det < 1e-8)."""
    e1, e2, s = v1 - v0, v2 - v0, o - v0
    p = np.cross(d, e2)
    det = f32(np.dot(e1, p))
    if det < 1e-8:
        return f32(-1e8)
    det = f32(1) / det
    u = f32(np.dot(s, p)) * det
    if u < 0 or u > 1:
        return f32(-1e8)
    q = np.cross(s, e1)
    v = f32(np.dot(d, q)) * det
    if v < 0 or v > 1 or u + v < 0 or u + v > 1:
        return f32(-1e8)
    return f32(np.dot(e2, q)) * det


def mesh_first_hit(mesh, verts, faces, o, d, side=64, max_steps=4096):
    """fastMesh_kernel.cu:230-295: walk the cells from the origin (index space, cell size 1); in each listed cell test
    all its faces two-sided; return the nearest t > 0 of the FIRST cell that has any hit, 0 if none."""
    origin, cell, lists = mesh
    verts = np.asarray(verts, f32)
    inv = f32(1) / cell
    p = ((np.asarray(o, f32) - origin) * inv).astype(f32)
    d = np.asarray(d, f32)
    c = np.clip(p.astype(np.int64), 0, side - 1)
    step = np.where(d >= 0, 1, -1)
    nxt = (c + step).astype(f32) + (step < 0)
    tmax = np.array([max(_safe_div(nxt[a] - p[a], d[a]), f32(0)) for a in range(3)], f32)
    tdel = np.array([abs(_safe_div(1.0, d[a])) for a in range(3)], f32)
    for _ in range(max_steps):
        if (c < 0).any() or (c >= side).any() or (tmax <= 0).all():
            break
        mx = tmax[0] < tmax[1] and tmax[0] <= tmax[2]
        my = tmax[1] < tmax[2] and tmax[1] <= tmax[0]
        mz = tmax[2] < tmax[0] and tmax[2] <= tmax[1]
        if not (mx or my or mz):
            mz = True
        key = (int(c[0]) << 12) | (int(c[1]) << 6) | int(c[2])
        if key in lists:
            best = f32(1e8)
            for fi in lists[key]:
                A, B_, C = verts[faces[fi][0]], verts[faces[fi][1]], verts[faces[fi][2]]
                n = np.cross(B_ - A, C - B_)
                t = _ray_triangle(o, d, A, C, B_) if np.dot(d, n) > 0 else _ray_triangle(o, d, A, B_, C)
                if 0 < t < best:
                    best = t
            if best != f32(1e8):
                return best
        a = 0 if mx else (1 if my else 2)
        tmax[a] += tdel[a]
        c[a] += step[a]
    return f32(0)
