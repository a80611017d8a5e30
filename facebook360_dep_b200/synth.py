"""Deterministic synthetic rigs and rig frames.

Rigs follow the reference's RigSimulator ring recipe (source/rig/RigSimulator.cpp:360-387:
camera i at angle -2*pi*i/S on a circle of radius 0.218 m, forward = radial, up = +z) and are
emitted in the reference's rig-JSON format (docs/rig.md), so the same file drives DerpCLI.
Frames are u16 BGR renders of a procedural 3-D texture ray-cast onto a sphere scene, so every
camera sees the same world and matching costs have real minima.

Pure torch/numpy; runs on CPU or GPU (device argument).  Input generation only — nothing here is
on the timed path.
"""
import math

import numpy as np
import torch


def ring_rig(num_cams, width, height, kind="FTHETA", radius=0.218, hfov_deg=None, fov=None,
             distortion=None):
    """Rig JSON dict for a horizontal ring of identical cameras."""
    cams = []
    for i in range(num_cams):
        th = -2.0 * math.pi * i / num_cams
        fwd = [math.cos(th), math.sin(th), 0.0]
        up = [0.0, 0.0, 1.0]
        # right = forward x up  (Camera.cpp:89-91)
        right = [fwd[1] * up[2] - fwd[2] * up[1], fwd[2] * up[0] - fwd[0] * up[2], fwd[0] * up[1] - fwd[1] * up[0]]
        cam = {
            "version": 1,
            "type": kind,
            "origin": [radius * fwd[0], radius * fwd[1], 0.0],
            "forward": fwd,
            "up": up,
            "right": right,
            "resolution": [width, height],
            "id": "cam%d" % i,
        }
        if kind == "RECTILINEAR":
            h = math.radians(hfov_deg if hfov_deg else 77.7)
            f = (width / 2.0) / math.tan(h / 2.0)  # RigSimulator.cpp:381-384
            cam["focal"] = [f, -f]
        else:
            f = width / math.pi  # image circle of half-angle pi/2 inscribed in the sensor width
            cam["focal"] = [f, -f]
            cam["fov"] = fov if fov is not None else 1.5707963
        if distortion:
            cam["distortion"] = list(distortion)
        cams.append(cam)
    return {"cameras": cams}


def wall_rig(num_cams, width, height, kind="RECTILINEAR", spacing=0.06, hfov_deg=90.0):
    """Planar array: all cameras look along +x from positions spread along y (and a little z), so every scene
    point is seen by every other camera — exercises cost evaluations with many (> 8) contributing sources."""
    cams = []
    for i in range(num_cams):
        fwd, up = [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]
        right = [fwd[1] * up[2] - fwd[2] * up[1], fwd[2] * up[0] - fwd[0] * up[2], fwd[0] * up[1] - fwd[1] * up[0]]
        cam = {"version": 1, "type": kind, "origin": [0.0, spacing * (i - (num_cams - 1) / 2.0), 0.02 * ((i * 7) % 3 - 1)],
               "forward": fwd, "up": up, "right": right, "resolution": [width, height], "id": "cam%d" % i}
        if kind == "RECTILINEAR":
            f = (width / 2.0) / math.tan(math.radians(hfov_deg) / 2.0)
        else:
            f = width / math.pi
            cam["fov"] = 1.5707963
        cam["focal"] = [f, -f]
        cams.append(cam)
    return {"cameras": cams}


def sphere_rig(num_cams, width, height, radius=0.33, seed=3, fov=1.57079632679):
    """Cameras on a sphere looking outward, like the reference's 16-camera test rig (res/test/rigs/rig.json: FTHETA,
    3360 x 2160, fov pi/2, shared distortion, individually calibrated focal / principal / roll): Fibonacci-sphere
    positions, random roll about the optical axis, focal and principal jittered per camera, non-square sensor.
    Exercises what the ring rigs do not: arbitrary orientations, off-centre principals, per-camera intrinsics."""
    rng = np.random.RandomState(seed)
    cams = []
    golden = math.pi * (3.0 - math.sqrt(5.0))
    for i in range(num_cams):
        z = 1.0 - 2.0 * (i + 0.5) / num_cams
        r = math.sqrt(max(0.0, 1.0 - z * z))
        th = golden * i
        fwd = np.array([r * math.cos(th), r * math.sin(th), z])
        helper = np.array([0.0, 0.0, 1.0]) if abs(z) < 0.9 else np.array([1.0, 0.0, 0.0])
        right0 = np.cross(fwd, helper)
        right0 /= np.linalg.norm(right0)
        up0 = np.cross(right0, fwd)
        roll = rng.uniform(0, 2 * math.pi)
        up = math.cos(roll) * up0 + math.sin(roll) * right0
        right = np.cross(fwd, up)  # right = forward x up (Camera.cpp:89-91)
        f = (width / 3.0) * (1.0 + 0.004 * rng.standard_normal())
        cams.append({
            "version": 1, "type": "FTHETA", "id": "cam%d" % i, "fov": fov,
            "origin": list(radius * fwd * (1.0 + 0.02 * rng.standard_normal())),
            "forward": list(fwd), "up": list(up), "right": list(right),
            "resolution": [width, height],
            "focal": [f, -f],
            "principal": [width / 2.0 + 0.01 * width * rng.standard_normal(),
                          height / 2.0 + 0.01 * height * rng.standard_normal()],
            "distortion": [-0.03413328161902581, 0.0004374554953464843, -0.0018843963208481174],
        })
    return {"cameras": cams}


# ---- minimal camera unprojection (pixel -> unit ray in rig space), fp64 torch -------------------
def _undistort(y, d):
    if not any(d):
        return y
    x = y.clone()
    for _ in range(20):  # Newton on distort(x) = y
        x2 = x * x
        f = x * (1 + x2 * (d[0] + x2 * (d[1] + x2 * d[2]))) - y
        df = 1 + x2 * (3 * d[0] + x2 * (5 * d[1] + x2 * 7 * d[2]))
        x = x - f / df
    return x


def pixel_rays(cam, width, height, device="cpu"):
    """Unit ray directions (H, W, 3) and origin (3,) for pixel centres of `cam` rescaled to width x height."""
    dd = dict(dtype=torch.float64, device=device)
    res = cam["resolution"]
    sx, sy = width / res[0], height / res[1]
    principal = cam.get("principal", [res[0] / 2.0, res[1] / 2.0])
    px, py = principal[0] * sx, principal[1] * sy
    fx, fy = cam["focal"][0] * sx, cam["focal"][1] * sy
    ys, xs = torch.meshgrid(torch.arange(height, **dd) + 0.5, torch.arange(width, **dd) + 0.5, indexing="ij")
    sxn = (xs - px) / fx
    syn = (ys - py) / fy
    norm = torch.sqrt(sxn * sxn + syn * syn).clamp_min(1e-300)
    dist = list(cam.get("distortion", [])) + [0.0] * 3
    r = _undistort(norm, dist[:3])
    kind = cam["type"]
    if kind == "FTHETA":
        theta = r
    elif kind == "RECTILINEAR":
        theta = torch.atan(r)
    elif kind == "EQUISOLID":
        theta = 2 * torch.asin((r / 2).clamp(max=1.0))
    else:
        theta = torch.asin(r.clamp(max=1.0))
    s = torch.sin(theta) / norm
    unit = torch.stack([s * sxn, s * syn, -torch.cos(theta)], dim=-1)
    right = torch.tensor(cam["right"], **dd)
    up = torch.tensor(cam["up"], **dd)
    back = -torch.tensor(cam["forward"], **dd)
    rot = torch.stack([right, up, back], dim=0)  # rows
    dirs = unit @ rot  # rotation^T * unit
    origin = torch.tensor(cam["origin"], **dd)
    return dirs, origin, theta


# ---- procedural texture ---------------------------------------------------------------------
def _hash3(ix, iy, iz, seed):
    h = (ix * 73856093) ^ (iy * 19349663) ^ (iz * 83492791) ^ (seed * 2654435761)
    h = h & 0xFFFFFFFF
    h = ((h ^ (h >> 16)) * 0x45D9F3B) & 0xFFFFFFFF
    h = ((h ^ (h >> 16)) * 0x45D9F3B) & 0xFFFFFFFF
    h = h ^ (h >> 16)
    return h.to(torch.float64) / 4294967296.0


def value_noise3(p, seed):
    """Trilinear value noise in [0,1) at points p (..., 3)."""
    pf = torch.floor(p)
    f = p - pf
    f = f * f * (3 - 2 * f)
    i = pf.to(torch.int64)
    out = 0
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                w = (f[..., 0] if dx else 1 - f[..., 0]) * (f[..., 1] if dy else 1 - f[..., 1]) * \
                    (f[..., 2] if dz else 1 - f[..., 2])
                out = out + w * _hash3(i[..., 0] + dx, i[..., 1] + dy, i[..., 2] + dz, seed)
    return out


def texture_bgr(points, seed=1234, base_freq=3.0, octaves=4):
    """Procedural colour in [0,1]^3 at world points (..., 3)."""
    chans = []
    for c in range(3):
        v = 0
        amp, freq, tot = 1.0, base_freq, 0.0
        for o in range(octaves):
            v = v + amp * value_noise3(points * freq + 17.0 * c, seed + 101 * c + o)
            tot += amp
            amp *= 0.6
            freq *= 2.3
        chans.append(v / tot)
    return torch.stack(chans, dim=-1)


class Scene:
    """Concentric shell at `shell_radius` plus random spheres between 1 and 10 m."""

    def __init__(self, seed=42, num_spheres=24, shell_radius=3.0, shift=(0.0, 0.0, 0.0)):
        rng = np.random.RandomState(seed)
        centers, radii = [], []
        for _ in range(num_spheres):
            d = rng.uniform(1.0, 10.0)
            az = rng.uniform(0, 2 * math.pi)
            el = rng.uniform(-0.6, 0.6)
            centers.append([d * math.cos(el) * math.cos(az) + shift[0], d * math.cos(el) * math.sin(az) + shift[1],
                            d * math.sin(el) + shift[2]])
            radii.append(rng.uniform(0.08, 0.3) * d)
        # keep spheres inside the shell only if closer than it; others are occluded by the shell anyway
        self.centers = np.array(centers)
        self.radii = np.array(radii)
        self.shell_radius = shell_radius

    def intersect(self, dirs, origin):
        """Distance along each ray to the first hit (H, W)."""
        dd = dict(dtype=torch.float64, device=dirs.device)
        o = origin
        # shell: |o + t d| = R, camera inside the shell
        b = (dirs * o).sum(-1)
        c = (o * o).sum() - self.shell_radius ** 2
        t = -b + torch.sqrt((b * b - c).clamp_min(0))
        for ctr, rad in zip(self.centers, self.radii):
            oc = o - torch.tensor(ctr, **dd)
            b = (dirs * oc).sum(-1)
            c = (oc * oc).sum() - rad * rad
            disc = b * b - c
            ts = -b - torch.sqrt(disc.clamp_min(0))
            hit = (disc > 0) & (ts > 0.05) & (ts < t)
            t = torch.where(hit, ts, t)
        return t


def render_camera(cam, width, height, scene, tex_seed=1234, noise_seed=None, device="cpu"):
    """u16 BGR image (H, W, 3) numpy + true disparity (H, W) numpy float32 (1/distance from the camera)."""
    dirs, origin, _ = pixel_rays(cam, width, height, device)
    t = scene.intersect(dirs, origin)
    pts = origin + dirs * t[..., None]
    col = texture_bgr(pts, seed=tex_seed)
    img = 2000.0 + col * 61000.0
    if noise_seed is not None:
        g = torch.Generator(device="cpu").manual_seed(noise_seed)
        img = img + (torch.randn(img.shape, generator=g, dtype=torch.float64) * (0.5 / 255 * 65535)).to(img.device)
    img = img.round().clamp(0, 65535).to(torch.int32).cpu().numpy().astype(np.uint16)
    return img, (1.0 / t).to(torch.float32).cpu().numpy()


def render_rig(rig, width, height, scene=None, tex_seed=1234, noise=True, device="cpu"):
    scene = scene or Scene()
    colors, disps = [], []
    for i, cam in enumerate(rig["cameras"]):
        img, d = render_camera(cam, width, height, scene, tex_seed, (7 + i) if noise else None, device)
        colors.append(img)
        disps.append(d)
    return colors, disps


def downscale_area(img, factor):
    """cv2.INTER_AREA-style integer-factor box downscale of a u16 HxWx3 image (resize.py:79)."""
    h, w, c = img.shape
    a = img.astype(np.float64).reshape(h // factor, factor, w // factor, factor, c).mean(axis=(1, 3))
    return np.clip(np.rint(a), 0, 65535).astype(np.uint16)


def random_colors(num_cams, width, height, seed=0):
    """Worst-case gather locality micro-benchmark input: i.i.d. uniform u16."""
    rng = np.random.RandomState(seed)
    return [rng.randint(0, 65536, size=(height, width, 3)).astype(np.uint16) for _ in range(num_cams)]
