"""Seeded synthetic inputs for the rasterizer path.

Everything is generated on the CPU with an explicit torch.Generator so the same bytes are
produced in this container and on the GPU box.
"""
import math

import numpy as np
import torch

BOX = np.array([3.0, 1.5, 2.0], dtype=np.float64)  # half extents of the 6 x 3 x 4 m room


def synth_scene(P, seed, sh_degree_max=3, device="cpu"):
    """70 % of the means on the six faces of the room (+N(0,0.02) jitter), 30 % uniform inside.
    Scales mimic the reference init (scene/gaussian_model.py:155-156: log(sqrt(mean 3-NN dist^2)))
    via the expected sample spacing, times exp(N(0,0.3)) per axis."""
    g = torch.Generator().manual_seed(int(seed))
    n_surf = int(round(0.7 * P))
    n_vol = P - n_surf
    box = torch.tensor(BOX, dtype=torch.float32)
    u = torch.rand(n_surf, 3, generator=g) * 2 - 1
    face = torch.randint(0, 6, (n_surf,), generator=g)
    axis, sign = face // 2, (face % 2).float() * 2 - 1
    u[torch.arange(n_surf), axis] = sign
    surf = u * box + torch.randn(n_surf, 3, generator=g) * 0.02
    vol = (torch.rand(n_vol, 3, generator=g) * 2 - 1) * box
    means = torch.cat([surf, vol], 0)
    area = 8.0 * (BOX[0] * BOX[1] + BOX[0] * BOX[2] + BOX[1] * BOX[2])
    volume = 8.0 * BOX.prod()
    s_surf = math.sqrt(area / max(n_surf, 1))
    s_vol = (volume / max(n_vol, 1)) ** (1.0 / 3.0)
    base = torch.cat([torch.full((n_surf,), s_surf), torch.full((n_vol,), s_vol)])
    log_scales = torch.log(base)[:, None] + torch.randn(P, 3, generator=g) * 0.3
    rot = torch.randn(P, 4, generator=g)
    rot = rot / rot.norm(dim=1, keepdim=True)
    opacity = torch.sigmoid(torch.randn(P, 1, generator=g) * 2.0)
    M = (sh_degree_max + 1) ** 2
    sh = torch.randn(P, M, 3, generator=g) * 0.05
    sh[:, 0, :] = (torch.rand(P, 3, generator=g) - 0.5) / 0.28209479177387814
    perm = torch.randperm(P, generator=g)  # surface / volume points interleaved like a real cloud
    out = dict(means3D=means[perm], scales=torch.exp(log_scales)[perm], rotations=rot[perm],
               opacities=opacity[perm], shs=sh[perm], confidence=torch.rand(P, 1, generator=g))
    return {k: v.contiguous().to(device) for k, v in out.items()}


def projection_matrix(fovx, fovy):
    """utils/graphics_utils.py:51-75 (the reference's non-standard P: P[2,2] = P[3,2] = 1)."""
    P = torch.zeros(4, 4)
    P[0, 0] = 1.0 / math.tan(fovx / 2)
    P[1, 1] = 1.0 / math.tan(fovy / 2)
    P[2, 2] = 1.0
    P[3, 2] = 1.0
    return P


def synth_camera(seed, width, height, fovx_deg=90.0, device="cpu", pos=None, yaw=None, pitch=None):
    """Camera built like PseudoCamera (scene/cameras.py:67-93). fy = fx (square pixels)."""
    rng = np.random.default_rng(int(seed))
    c = (rng.uniform(-1, 1, 3) * 0.5) if pos is None else np.asarray(pos, dtype=np.float64)
    yaw = rng.uniform(0, 2 * math.pi) if yaw is None else yaw
    pitch = rng.normal(0, math.radians(10.0)) if pitch is None else pitch
    f = np.array([math.cos(pitch) * math.sin(yaw), math.sin(pitch), math.cos(pitch) * math.cos(yaw)])
    r = np.cross([0.0, 1.0, 0.0], f)
    r /= np.linalg.norm(r)
    d = np.cross(f, r)
    Rc = np.stack([r, d, f], 1)  # camera-to-world rotation (columns = camera axes)
    Rt = np.eye(4)
    Rt[:3, :3] = Rc.T
    Rt[:3, 3] = -Rc.T @ c
    fovx = math.radians(fovx_deg)
    fx = width / (2 * math.tan(fovx / 2))
    fovy = 2 * math.atan(height / (2 * fx))
    wvt = torch.tensor(np.float32(Rt)).transpose(0, 1).contiguous()
    proj = projection_matrix(fovx, fovy).transpose(0, 1)
    full = (wvt.unsqueeze(0).bmm(proj.unsqueeze(0))).squeeze(0).contiguous()
    campos = wvt.inverse()[3, :3].contiguous()
    return dict(width=width, height=height, tanfovx=math.tan(fovx * 0.5), tanfovy=math.tan(fovy * 0.5),
                viewmatrix=wvt.to(device), projmatrix=full.to(device), campos=campos.to(device))


CONFIGS = {
    # name: (P, W, H, seed)  -- SURVEY.md section 8
    "tiny": (2_000, 128, 128, 20260000),
    "small": (50_000, 320, 240, 20260001),
    "C2": (500_000, 640, 480, 20260002),
    "C4": (800_000, 1600, 1066, 20260004),
    "C5": (2_000_000, 640, 480, 20260005),
}
