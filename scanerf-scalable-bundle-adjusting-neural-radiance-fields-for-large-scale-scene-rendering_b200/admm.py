"""ADMM pose consensus across tiles as ONE collective per synchronisation.

The reference exchanges poses through a CPU master process: every tile pickles
{"pose": se3_refine [n,6], "idx": global camera ids, "confidence": [n]} into a
multiprocessing.Manager slot (tile.py:477-495), the master busy-waits for all tiles, forms the
confidence-weighted mean per camera, the overlap set and the residuals
(admm_trainer.py:124-170) and writes per-tile slices back (:173-179); tiles busy-wait, then
ConsensusManager.update applies the dual step (consensus.py:40-50).

Here one process owns one GPU (rank) and any number of tiles; the exchange is a single
`all_reduce(SUM)` of a dense [N_cam_global, 8] fp32 buffer (6 x conf*se3, sum conf, count) over
NCCL / NVLink, after which every rank derives z, the overlap flags and the residuals locally with
the master's arithmetic.  The payload is <= 32 KB per 1000 cameras: latency-bound, nothing to fuse
with a kernel.  (`torch.distributed` with the `gloo` backend runs the same code on CPU tensors --
that is what the CPU tests use.)
"""
import torch
import torch.distributed as dist


class PoseConsensus:
    """Replacement of ADMM_TRAINER.master_process's consensus step (admm_trainer.py:124-179)."""

    def __init__(self, num_camera, device, group=None):
        self.num_camera = int(num_camera)
        self.device = torch.device(device)
        self.group = group
        self.shared_poses = torch.zeros(self.num_camera, 6, dtype=torch.float32, device=self.device)   # z
        self.buf = torch.zeros(self.num_camera, 8, dtype=torch.float32, device=self.device)
        self.primal_residual = self.dual_residual = None

    def _reduce(self, t):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    @torch.no_grad()
    def exchange(self, tiles):
        """tiles: list of (pose [n,6] f32, idx int64 [n] global camera ids, confidence [n] f32) for
        the tiles THIS rank owns.  Returns one dict per tile: {"shared_poses": z[idx] [n,6],
        "overlap_idxs": int64 positions inside the tile whose camera is seen by >= 2 tiles}."""
        buf = self.buf.zero_()
        for pose, idx, conf in tiles:
            idx = idx.to(self.device).long()
            pose, conf = pose.detach().to(self.device, torch.float32), conf.to(self.device, torch.float32)
            row = torch.cat([conf[:, None] * pose, conf[:, None], torch.ones_like(conf)[:, None]], -1)
            buf.index_add_(0, idx, row)                       # a tile lists a camera at most once
        self._reduce(buf)
        weight = buf[:, 6].clone()
        weight[weight == 0] = 1                               # admm_trainer.py:151
        z = buf[:, :6] / weight[:, None]
        overlap = buf[:, 7] >= 2                              # :150 (count >= 2)
        self.dual_residual = torch.mean(torch.abs(self.shared_poses - z))          # :154
        self.shared_poses = z
        # primal residual: mean over ALL tiles of mean |x_t - z[idx_t]| (:158-165)
        acc = torch.zeros(2, dtype=torch.float32, device=self.device)
        out = []
        for pose, idx, conf in tiles:
            idx = idx.to(self.device).long()
            acc[0] += torch.mean(torch.abs(pose.detach().to(self.device, torch.float32) - z[idx]))
            acc[1] += 1
            out.append({"shared_poses": z[idx], "overlap_idxs": torch.nonzero(overlap[idx])[:, 0]})
        self._reduce(acc)
        self.primal_residual = acc[0] / acc[1].clamp_min(1)
        return out

    def residual_line(self):
        """The line the master appends to admm_error.txt (admm_trainer.py:169-170)."""
        return f"primal_residual: {float(self.primal_residual):.8f}\tdual_residual: {float(self.dual_residual):.8f}\n"


class ConsensusManager:
    """Per-tile ADMM state (consensus.py:4-82): consensus copy z, scaled dual u (`delta_se3`),
    overlap flags, penalty rho.  Same attribute names and checkpoint keys; no CPU round trip in
    update()."""

    def __init__(self, se3_refine, rho, device=None):
        self.se3_refine = se3_refine                          # nn.Parameter [n,6] (CAM.se3_refine)
        self.device = device if device is not None else se3_refine.device
        n = se3_refine.shape[0]
        self.shared_se3 = se3_refine.detach().clone()
        self.delta_se3 = torch.zeros(n, 6, dtype=torch.float32, device=self.device)
        self.overlap_flags = torch.zeros(n, dtype=torch.bool, device=self.device)
        self.rho = torch.ones(6, dtype=torch.float32, device=self.device) * float(rho)
        self.has_overlap = False          # host-side copy of overlap_flags.any(), updated without a device sync

    def export_check_point(self):
        f = lambda t: t.detach().cpu().numpy()
        return {"shared_se3": f(self.shared_se3), "delta_se3": f(self.delta_se3), "overlap_flags": f(self.overlap_flags),
                "rho": f(self.rho)}

    def load_check_point(self, ckp):
        f = lambda a: torch.from_numpy(a).to(self.device)
        self.shared_se3, self.delta_se3 = f(ckp["shared_se3"]), f(ckp["delta_se3"])
        self.overlap_flags, self.rho = f(ckp["overlap_flags"]), f(ckp["rho"])
        self.has_overlap = bool(ckp["overlap_flags"].any())

    @torch.no_grad()
    def update(self, shared_se3, overlap_idxs):
        """consensus.py:40-50: z <- shared, u <- u + 1.5 (x - z) (over-relaxed dual step), mark overlap cameras."""
        self.shared_se3 = shared_se3.to(self.device)
        self.delta_se3 = self.delta_se3 + (1 + 0.5) * (self.se3_refine.detach() - self.shared_se3)
        if overlap_idxs.shape[0] > 0:        # (the shape is known on the host: nonzero() inside exchange() already synchronised)
            self.overlap_flags[overlap_idxs.to(self.device)] = True
            self.has_overlap = True

    def camera_loss(self):
        """consensus.py:70-76: mean over overlap cameras and the 6 generators of rho (x - z + u)^2."""
        c = (self.se3_refine - self.shared_se3 + self.delta_se3) ** 2
        # masked mean instead of boolean indexing: same value, no device->host synchronisation per step
        f = self.overlap_flags[:, None].to(c.dtype)
        return torch.sum(self.rho[None, :] * c * f) / (6.0 * f.sum().clamp_min(1.0))

    def __call__(self):
        return self.camera_loss() if self.has_overlap else None


class DepthExchange:
    """The `shared_depth` exchange (tile.py:432-475 render_shared_depth, :366-430 update_occlusion_mask,
    admm_trainer.py:31-32, 117-118).  In the reference every tile renders a half-resolution depth map for each of its
    overlap cameras whose centre lies inside the tile and stores it in a Manager().list indexed by the global camera id;
    afterwards every tile reads the maps of ITS cameras to rebuild its occlusion masks.  Here the tiles live on different
    ranks: one padded all_gather of (camera id, writer tile, map) over NCCL.  Where two tiles wrote the same camera the
    reference keeps whichever process stored last; here the higher tile index wins (deterministic)."""

    def __init__(self, h, w, device, group=None):
        self.h, self.w = int(h), int(w)
        self.device = torch.device(device)
        self.group = group

    @torch.no_grad()
    def exchange(self, entries, want_ids):
        """entries: list of (tile index, global camera ids int64 [k], depth maps [k, h, w] f32) written by the tiles THIS
        rank owns (k may be 0); want_ids: int64 [m] global camera ids this rank wants back.
        Returns (have bool [m], maps f32 [m, h, w]): the map of every wanted camera somebody rendered."""
        dev = self.device
        ids = [e[1].to(dev).long() for e in entries] or [torch.zeros(0, dtype=torch.long, device=dev)]
        pri = [torch.full_like(e[1].to(dev).long(), int(e[0])) for e in entries] or [torch.zeros(0, dtype=torch.long, device=dev)]
        maps = [e[2].to(dev, torch.float32).reshape(-1, self.h, self.w) for e in entries] or [torch.zeros(0, self.h, self.w, device=dev)]
        ids, pri, maps = torch.cat(ids), torch.cat(pri), torch.cat(maps)
        world = dist.get_world_size(self.group) if (dist.is_available() and dist.is_initialized()) else 1
        if world > 1:
            n = torch.tensor([ids.shape[0]], dtype=torch.long, device=dev)
            counts = [torch.zeros_like(n) for _ in range(world)]
            dist.all_gather(counts, n, group=self.group)
            counts = [int(c.item()) for c in counts]
            cap = max(max(counts), 1)
            pad = cap - ids.shape[0]
            meta = torch.cat([torch.stack([ids, pri], -1), torch.full((pad, 2), -1, dtype=torch.long, device=dev)])
            body = torch.cat([maps, torch.zeros(pad, self.h, self.w, device=dev)])
            metas = [torch.empty_like(meta) for _ in range(world)]
            bodies = [torch.empty_like(body) for _ in range(world)]
            dist.all_gather(metas, meta, group=self.group)
            dist.all_gather(bodies, body, group=self.group)
            ids = torch.cat([m[:c, 0] for m, c in zip(metas, counts)])
            pri = torch.cat([m[:c, 1] for m, c in zip(metas, counts)])
            maps = torch.cat([b[:c] for b, c in zip(bodies, counts)])
        want = want_ids.to(dev).long()
        have = torch.zeros(want.shape[0], dtype=torch.bool, device=dev)
        out = torch.zeros(want.shape[0], self.h, self.w, dtype=torch.float32, device=dev)
        if ids.numel() == 0 or want.numel() == 0:
            return have, out
        # for every wanted camera: the entry with the highest writer tile
        order = torch.argsort(pri, stable=True)                     # ascending: later (higher-tile) entries overwrite
        ids, maps = ids[order], maps[order]
        hit = ids[None, :] == want[:, None]                         # [m, n]
        have = hit.any(1)
        last = (hit.long() * torch.arange(1, ids.shape[0] + 1, device=dev)[None, :]).max(1)[0] - 1
        out[have] = maps[last[have]]
        return have, out


def synchronize_tiles(exchange, steps):
    """TILE.commit + the master's consensus + TILE.synchronize (tile.py:477-508, admm_trainer.py:124-179, 241-262) for ALL
    tiles resident on this rank at once: one all-reduce, then every tile's dual update.  `steps`: TileStep objects with
    enable_consensus() done."""
    outs = exchange.exchange([(st.poses.se3_refine, st.camera_ids, st.confidence) for st in steps])
    for st, o in zip(steps, outs):
        st.consensus.update(o["shared_poses"], o["overlap_idxs"])
    return outs
