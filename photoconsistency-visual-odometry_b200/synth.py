"""Deterministic synthetic RGB-D scene.

A tilted textured plane rendered in closed form for any camera pose -- no resampling, so the
ground-truth motion is exact.  Motion convention is the reference's: X1 = Rt(xi) * X0 with
xi = (x, y, z, yaw, pitch, roll) and Rt = eigenPose (CPhotoconsistencyOdometry.h:47-71).

numpy implementation for tests / small cases; `render_batch_torch` builds large batches on the
GPU for bench.py (data synthesis only -- never part of a timed region).
"""
import numpy as np

K_FRAME_ALIGNMENT = np.array([[525., 0., 319.5], [0., 525., 239.5], [0., 0., 1.]])   # FrameAlignment.cpp:69-71
K_VISUAL_ODOMETRY = np.array([[517.3, 0., 318.6], [0., 516.5, 255.3], [0., 0., 1.]])  # VisualOdometry.cpp:171-173
K_8K = np.array([[6300., 0., 3839.5], [0., 6300., 2159.5], [0., 0., 1.]])

PLANE_N = np.array([0.15, -0.10, 1.0]) / np.linalg.norm([0.15, -0.10, 1.0])
PLANE_D = 2.0
XI_CONFIG1 = np.array([0.02, -0.01, 0.015, 0.01, -0.008, 0.006])


def state_to_rt(xi):
    x, y, z, yaw, pitch, roll = [float(v) for v in xi]
    cy, sy, cp, sp, cr, sr = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch), np.cos(roll), np.sin(roll)
    return np.array([
        [cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr, x],
        [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr, y],
        [-sp, cp * sr, cp * cr, z],
        [0., 0., 0., 1.]])


def texture(x, y, xp=np):
    t = (0.5 + 0.18 * xp.sin(7.1 * x + 0.3) * xp.cos(5.3 * y - 0.2) + 0.12 * xp.sin(17. * x - 11. * y)
         + 0.10 * xp.cos(29. * y + 13. * x + 1.) + 0.08 * xp.sin(41. * x) * xp.sin(37. * y))
    return xp.clip(t, 0., 1.)


def render(K, rows, cols, xi=None):
    """Noise-free intensity in [0,1] and depth (z, metres) seen by the camera whose frame is
    X_cam = Rt(xi) * X_0 (xi=None: camera 0)."""
    fx, fy, ox, oy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    c, r = np.meshgrid(np.arange(cols, dtype=np.float64), np.arange(rows, dtype=np.float64))
    dx, dy = (c - ox) / fx, (r - oy) / fy
    if xi is None:
        R, t = np.eye(3), np.zeros(3)
    else:
        Rt = state_to_rt(xi)
        R, t = Rt[:3, :3], Rt[:3, 3]
    n1 = R @ PLANE_N                       # plane normal in the camera frame
    s = (PLANE_D + n1 @ t) / (n1[0] * dx + n1[1] * dy + n1[2])   # depth along z of the hit
    X1 = np.stack([s * dx, s * dy, s], axis=-1)
    X0 = (X1 - t) @ R                      # R^T (X1 - t)
    return texture(X0[..., 0], X0[..., 1]), s


def make_pair(rows=480, cols=640, K=K_FRAME_ALIGNMENT, xi=XI_CONFIG1, seed=0, holes=True,
              depth_f32=True):
    """One RGB-D pair: (gray0 u8, depth0 f64, gray1 u8, depth1 f64).

    Intensity = round(255*T + N(0,1)) -> u8; depth = exact z, optionally rounded to an
    fp32-representable double (so the device's fp32 upload is lossless at level 0);
    3% Bernoulli holes (0) + a 16-px column band at 6.0 m exercise both validity bounds."""
    rng = np.random.default_rng(seed)
    out = []
    for pose in (None, xi):
        T, z = render(K, rows, cols, pose)
        g = np.clip(np.rint(255. * T + rng.standard_normal(T.shape)), 0, 255).astype(np.uint8)
        d = z.copy()
        if holes:
            d[rng.random(T.shape) < 0.03] = 0.
            b0 = (cols * 5) // 8
            d[:, b0:b0 + max(1, cols // 40)] = 6.0
        if depth_f32:
            d = d.astype(np.float32).astype(np.float64)
        out += [g, d]
    return tuple(out)


def random_motion(seed):
    rng = np.random.default_rng(10_000 + seed)
    return np.concatenate([rng.uniform(-0.02, 0.02, 3), rng.uniform(-0.01, 0.01, 3)])


def make_batch(num_pairs, rows=480, cols=640, K=K_FRAME_ALIGNMENT, seed0=0, depth_f32=True):
    """BASELINE config 4 inputs on the host: pair p uses seed seed0+p and motion U(+-0.02 m, +-0.01 rad)."""
    g0 = np.empty((num_pairs, rows, cols), np.uint8)
    g1 = np.empty((num_pairs, rows, cols), np.uint8)
    d0 = np.empty((num_pairs, rows, cols), np.float64)
    xis = np.empty((num_pairs, 6))
    for p in range(num_pairs):
        xis[p] = random_motion(seed0 + p)
        g0[p], d0[p], g1[p], _ = make_pair(rows, cols, K, xis[p], seed0 + p, True, depth_f32)
    return g0, d0, g1, xis


def lissajous_pose(k, n=1000):
    """Camera pose of frame k on a smooth path, <= ~1 cm / 0.3 deg per frame (config 2)."""
    t = 2. * np.pi * k / n
    return np.array([0.8 * np.sin(3 * t) * 0.3, 0.5 * np.sin(2 * t + 0.5) * 0.3, 0.4 * np.sin(t) * 0.3,
                     0.25 * np.sin(2 * t) * 0.3, 0.2 * np.sin(3 * t + 1.) * 0.3, 0.2 * np.sin(t + 2.) * 0.3])


def make_sequence_frame(k, rows=480, cols=640, K=K_VISUAL_ODOMETRY, n=1000, depth_f32=True):
    """Frame k of the config-2 sequence: (gray u8, depth f64), camera at lissajous_pose(k)."""
    rng = np.random.default_rng(50_000 + k)
    T, z = render(K, rows, cols, lissajous_pose(k, n))
    g = np.clip(np.rint(255. * T + rng.standard_normal(T.shape)), 0, 255).astype(np.uint8)
    d = z.copy()
    d[rng.random(T.shape) < 0.03] = 0.
    if depth_f32:
        d = d.astype(np.float32).astype(np.float64)
    return g, d


def render_sequence_torch(num_frames, rows, cols, K, device, seed=0, chunk=32, depth_dtype=None):
    """GPU synthesis of the config-2 sequence (same scene and Lissajous path as make_sequence_frame,
    noise and holes from torch's generator): gray u8 [N,R,C] and depth [N,R,C] on the device."""
    import torch
    depth_dtype = depth_dtype or torch.float32
    gen = torch.Generator(device=device)
    gen.manual_seed(4321 + seed)
    fx, fy, ox, oy = float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2])
    gray = torch.empty((num_frames, rows, cols), dtype=torch.uint8, device=device)
    depth = torch.empty((num_frames, rows, cols), dtype=depth_dtype, device=device)
    c = torch.arange(cols, dtype=torch.float64, device=device)[None, None, :]
    r = torch.arange(rows, dtype=torch.float64, device=device)[None, :, None]
    dx, dy = (c - ox) / fx, (r - oy) / fy
    n0 = torch.tensor(PLANE_N, dtype=torch.float64, device=device)
    for a in range(0, num_frames, chunk):
        b = min(num_frames, a + chunk)
        Rt = torch.tensor(np.stack([state_to_rt(lissajous_pose(k, num_frames)) for k in range(a, b)]), dtype=torch.float64, device=device)
        R, t = Rt[:, :3, :3], Rt[:, :3, 3]
        n1 = torch.einsum("bij,j->bi", R, n0)
        num = (PLANE_D + (n1 * t).sum(-1))[:, None, None]
        s = num / (n1[:, 0, None, None] * dx + n1[:, 1, None, None] * dy + n1[:, 2, None, None])
        X1 = torch.stack([s * dx, s * dy, s], dim=-1) - t[:, None, None, :]
        X0 = torch.einsum("brck,bkj->brcj", X1, R)
        T = texture(X0[..., 0], X0[..., 1], xp=torch)
        noise = torch.randn(T.shape, generator=gen, device=device, dtype=torch.float64)
        gray[a:b] = torch.clamp(torch.round(255. * T + noise), 0, 255).to(torch.uint8)
        d = s.clone()
        d[torch.rand(T.shape, generator=gen, device=device) < 0.03] = 0.
        depth[a:b] = d.to(depth_dtype)
    return gray, depth


def render_batch_torch(num_pairs, rows, cols, K, device, seed0=0, chunk=64, depth_dtype=None, xis=None):
    """GPU synthesis of `num_pairs` pairs (same scene/motion law as make_batch; noise and holes
    come from torch's generator, so values differ from the numpy version).  Returns device
    tensors gray0 u8 [P,R,C], depth0 f32 [P,R,C], gray1 u8 [P,R,C] and the motions [P,6] (cpu);
    `xis` overrides the random motions."""
    import torch
    depth_dtype = depth_dtype or torch.float32
    gen = torch.Generator(device=device)
    gen.manual_seed(1234 + seed0)
    fx, fy, ox, oy = float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2])
    g0 = torch.empty((num_pairs, rows, cols), dtype=torch.uint8, device=device)
    g1 = torch.empty_like(g0)
    d0 = torch.empty((num_pairs, rows, cols), dtype=depth_dtype, device=device)
    xis = np.stack([random_motion(seed0 + p) for p in range(num_pairs)]) if xis is None else np.asarray(xis, dtype=np.float64).reshape(num_pairs, 6)
    c = torch.arange(cols, dtype=torch.float64, device=device)[None, None, :]
    r = torch.arange(rows, dtype=torch.float64, device=device)[None, :, None]
    dx, dy = (c - ox) / fx, (r - oy) / fy
    n0 = torch.tensor(PLANE_N, dtype=torch.float64, device=device)
    for a in range(0, num_pairs, chunk):
        b = min(num_pairs, a + chunk)
        Rt = torch.tensor(np.stack([state_to_rt(x) for x in xis[a:b]]), dtype=torch.float64, device=device)
        for which in (0, 1):
            if which == 0:
                R = torch.eye(3, dtype=torch.float64, device=device).expand(b - a, 3, 3)
                t = torch.zeros((b - a, 3), dtype=torch.float64, device=device)
            else:
                R, t = Rt[:, :3, :3], Rt[:, :3, 3]
            n1 = torch.einsum("bij,j->bi", R, n0)
            num = (PLANE_D + (n1 * t).sum(-1))[:, None, None]
            s = num / (n1[:, 0, None, None] * dx + n1[:, 1, None, None] * dy + n1[:, 2, None, None])
            X1 = torch.stack([s * dx, s * dy, s], dim=-1) - t[:, None, None, :]
            X0 = torch.einsum("brck,bkj->brcj", X1, R)
            T = texture(X0[..., 0], X0[..., 1], xp=torch)
            noise = torch.randn(T.shape, generator=gen, device=device, dtype=torch.float64)
            g = torch.clamp(torch.round(255. * T + noise), 0, 255).to(torch.uint8)
            if which == 0:
                g0[a:b] = g
                d = s.clone()
                d[torch.rand(T.shape, generator=gen, device=device) < 0.03] = 0.
                b0 = (cols * 5) // 8
                d[:, :, b0:b0 + max(1, cols // 40)] = 6.0
                d0[a:b] = d.to(depth_dtype)
            else:
                g1[a:b] = g
    return g0, d0, g1, xis
