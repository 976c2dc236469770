"""Seeded synthetic inputs shared by the tests (and by bench.py)."""
import numpy as np
import torch


def camera_rig(n_cam, H, W, gen, center=(10.0, 3.0, 15.0), radius=6.0, fx=600.0):
    """Cameras on a ring looking at `center`.  Returns Ks [N,3,3], c2w [N,3,4] (float32)."""
    ang = torch.linspace(0, 2 * np.pi, n_cam + 1)[:-1] + 0.1 * torch.rand(n_cam, generator=gen)
    c = torch.tensor(center)
    pos = torch.stack([c[0] + radius * torch.cos(ang), c[1] + 0.5 * torch.rand(n_cam, generator=gen),
                       c[2] + radius * torch.sin(ang)], -1)
    fwd = torch.nn.functional.normalize(c[None] - pos + 0.3 * torch.randn(n_cam, 3, generator=gen), dim=-1)
    up = torch.tensor([0.0, 1.0, 0.0]).expand_as(fwd)
    right = torch.nn.functional.normalize(torch.cross(fwd, up, dim=-1), dim=-1)
    down = torch.cross(fwd, right, dim=-1)
    R = torch.stack([right, down, fwd], -1)          # columns = camera axes in world
    c2w = torch.cat([R, pos[..., None]], -1).float()
    K = torch.tensor([[fx, 0, W / 2.0], [0, fx, H / 2.0], [0, 0, 1.0]]).expand(n_cam, 3, 3).contiguous()
    return K.float(), c2w


def random_rays(B, gen, corner, size, inside=True):
    """Ray origins in/around the box and random directions (some axis-parallel, some zero components)."""
    corner, size = torch.as_tensor(corner, dtype=torch.float32), torch.as_tensor(size, dtype=torch.float32)
    u = torch.rand(B, 3, generator=gen)
    o = corner + size * (u if inside else (u * 2.0 - 0.5))
    d = torch.randn(B, 3, generator=gen)
    k = max(B // 16, 1)
    d[:k, 0] = 0.0                      # exercise safe_divide
    d[k:2 * k, 1] = 0.0
    d[2 * k:3 * k] = torch.tensor([0.0, 0.0, 1.0])
    d = d * (0.5 + torch.rand(B, 1, generator=gen))      # not normalised
    return o.contiguous(), d.contiguous()


def occupancy(log2dim, gen, p=0.3):
    shape = [2 ** int(v) for v in log2dim]
    return torch.rand(shape, generator=gen) < p
