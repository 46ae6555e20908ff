"""Independent numpy + OpenCV(cv2) restatement of the reference analytic path.

TEST INFRASTRUCTURE.  Written separately from oracle/phovo_oracle.c (vectorised, compact closed
form of the Jacobian, the REAL OpenCV for resize/Scharr/GaussianBlur/convertTo, numpy's LAPACK
inverse) so that agreement between the two is evidence that both restate the reference:
  AN = phovo/include/CPhotoconsistencyOdometryAnalytic.h  (AN:115-189, 191-367, 376-426, 500-563)
  CE = phovo/include/CPhotoconsistencyOdometryCeres.h     (CE:156-269), third_party/sample.h
Used by tests/ and by tests/golden/make_golden.py only.
"""
import numpy as np

try:
    import cv2
except Exception:  # pragma: no cover - cv2 exists in the build container; the GPU box may lack it
    cv2 = None


def build_pyramids(gray0, depth0, gray1, num_levels, blur, grad_scale):
    """AN:466-491 replayed through cv2 (same calls as the reference makes in C++), on OpenCV's portable
    C++ code path: the IPP / SIMD dispatch of this particular cv2 build rounds differently in the last bit
    (and does not go through single precision on border cells of the half-size resize)."""
    assert cv2 is not None
    was_optimized = cv2.useOptimized()
    cv2.setUseOptimized(False)
    try:
        return _build_pyramids(gray0, depth0, gray1, num_levels, blur, grad_scale)
    finally:
        cv2.setUseOptimized(was_optimized)


def _build_pyramids(gray0, depth0, gray1, num_levels, blur, grad_scale):
    a0 = gray0.astype(np.float64) * (1. / 255)        # convertTo(CV_64F, 1./255)
    a1 = gray1.astype(np.float64) * (1. / 255)
    I0, D0, I1, Gx, Gy = [], [], [], [], []
    factor = 1.
    for lvl in range(num_levels):
        def rs(img):
            return img if lvl == 0 else cv2.resize(img, (0, 0), fx=factor, fy=factor)

        def bl(img):
            k = int(blur[lvl])
            if k > 0:
                img = cv2.GaussianBlur(img, (k, k), 3)
                img = cv2.GaussianBlur(img, (k, k), 3)
            return img
        i0, i1, d0 = bl(rs(a0)), bl(rs(a1)), rs(depth0.astype(np.float64))
        I0.append(i0); I1.append(i1); D0.append(d0)
        Gx.append(cv2.Scharr(i1, cv2.CV_64F, 1, 0, scale=float(grad_scale[lvl]), delta=0, borderType=cv2.BORDER_DEFAULT))
        Gy.append(cv2.Scharr(i1, cv2.CV_64F, 0, 1, scale=float(grad_scale[lvl]), delta=0, borderType=cv2.BORDER_DEFAULT))
        factor = factor / 2
    return I0, D0, I1, Gx, Gy


def _rot(state):
    x, y, z, yaw, pitch, roll = state
    cy, sy, cp, sp, cr, sr = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch), np.cos(roll), np.sin(roll)
    R = np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                  [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                  [-sp, cp * sr, cp * cr]])
    return R, (cy, sy, cp, sp, cr, sr)


def round_half_away(v):
    return np.sign(v) * np.floor(np.abs(v) + 0.5)


def analytic_eval(I0, D0, I1, Gx, Gy, K, level, state, min_depth=0.3, max_depth=5.0, fixed=False):
    """One pass of AN:191-367 + the products of AN:538-539, gather form.
    Returns H (6x6), g (6), count, residual vector, winner map."""
    rows, cols = I0.shape
    N = rows * cols
    sf = 1.0 / 2 ** level
    fx, fy, ox, oy = K[0, 0] * sf, K[1, 1] * sf, K[0, 2] * sf, K[1, 2] * sf
    x, y, z = state[:3]
    R, (cy, sy, cp, sp, cr, sr) = _rot(state)
    c, r = np.meshgrid(np.arange(cols, dtype=np.float64), np.arange(rows, dtype=np.float64))
    d = D0
    valid = (min_depth < d) & (d < max_depth)
    px = (c - ox) * d * (1. / fx)
    py = (r - oy) * d * (1. / fy)
    pz = d
    q0 = R[0, 0] * px + R[0, 1] * py + R[0, 2] * pz
    q1 = R[1, 0] * px + R[1, 1] * py + R[1, 2] * pz
    q2 = R[2, 0] * px + R[2, 1] * py + R[2, 2] * pz
    X, Y, Z = q0 + x, q1 + y, q2 + z
    with np.errstate(all="ignore"):
        iz = 1.0 / Z
        tc = (X * fx) * iz + ox
        tr = (Y * fy) * iz + oy
        ti, tj = round_half_away(tr), round_half_away(tc)
        inb = valid & (ti >= 0) & (ti < rows) & (tj >= 0) & (tj < cols)
    idx = np.flatnonzero(inb.ravel())
    tgt = (cols * ti.ravel()[idx] + tj.ravel()[idx]).astype(np.int64)
    winner = np.full(N, -1, dtype=np.int64)
    np.maximum.at(winner, tgt, idx)               # raster order: the largest source index wrote last
    res = np.zeros(N)
    has = winner >= 0
    res[has] = I1.ravel()[has] - I0.ravel()[winner[has]]
    # Jacobian (SURVEY appendix C closed form of AN:243-342)
    A = q0 + (x if fixed else px * x)
    B = q1 + y
    Zp = -(sp * sr * py + sp * cr * pz + cp * px)
    Zr = R[2, 2] * py - R[2, 1] * pz
    iz2 = iz * iz
    Ju = [fx * iz, 0 * iz, -fx * A * iz2, -fx * q1 * iz,
          fx * (cy * q2 * iz - Zp * A * iz2), fx * ((R[0, 2] * py - R[0, 1] * pz) * iz - Zr * A * iz2)]
    Jv = [0 * iz, fy * iz, -fy * B * iz2, fy * q0 * iz,
          fy * (sy * q2 * iz - Zp * B * iz2), fy * ((R[1, 2] * py - R[1, 1] * pz) * iz - Zr * B * iz2)]
    J = np.zeros((N, 6))
    m = inb.ravel()
    for k in range(6):
        J[m, k] = (Gx * Ju[k] + Gy * Jv[k]).ravel()[m]
    H = J.T @ J
    g = J.T @ res
    return H, g, int(m.sum()), res, winner


def analytic_optimize(I0, D0, I1, Gx, Gy, K, cfg_levels, max_iters, lam, min_grad, state0,
                      min_depth=0.3, max_depth=5.0, fixed=False):
    """AN:500-563 + AN:376-392.  Returns final state and per-iteration log."""
    state = np.array(state0, dtype=np.float64)
    log = []
    for level in range(cfg_levels - 1, -1, -1):
        it = 0
        g = np.zeros(6)
        while True:
            if max_iters[level] > 0:
                H, g, cnt, _, _ = analytic_eval(I0[level], D0[level], I1[level], Gx[level], Gy[level], K, level,
                                                state, min_depth, max_depth, fixed)
                s_in = state.copy()
                state = state - lam[level] * (np.linalg.inv(H) @ g)
                log.append(dict(level=level, iteration=it, num_valid=cnt, H=H, g=g, state_in=s_in,
                                state_out=state.copy(), grad_norm=float(np.linalg.norm(g))))
            it += 1
            if it >= max_iters[level]:
                break
            if np.linalg.norm(g) < min_grad[level]:
                break
    return state, log


def ceres_eval(I0, D0, I1, Gx, Gy, K, level, state, min_depth=0.3, max_depth=5.0):
    """CE:156-269 residuals + autodiff Jacobian (loop form; use on small images)."""
    rows, cols = I0.shape
    N = rows * cols
    fx, fy = K[0, 0] / 2.0 ** level, K[1, 1] / 2.0 ** level
    ox, oy = K[0, 2] / 2.0 ** level, K[1, 2] / 2.0 ** level
    x, y, z = state[:3]
    R, (cy, sy, cp, sp, cr, sr) = _rot(state)
    res = np.zeros(N)
    J = np.zeros((N, 6))

    def axis(v, size):
        iv = int(v)
        if iv < 0:
            return 0, 0, 1.0
        if iv > size - 2:
            return size - 1, size - 1, 1.0
        return iv, iv + 1, (iv + 1) - v

    for r in range(rows):
        for c in range(cols):
            d = D0[r, c]
            if not (min_depth < d < max_depth):
                continue
            p = np.array([(c - ox) * d / fx, (r - oy) * d / fy, d])
            q = R @ p
            X, Y, Z = q[0] + x, q[1] + y, q[2] + z
            if Z == 0:
                continue
            tc, tr = X * fx / Z + ox, Y * fy / Z + oy
            if not (0 <= tr < rows and 0 <= tc < cols):
                continue
            y1, y2, dy = axis(tr - 0.5, rows)
            x1, x2, dx = axis(tc - 0.5, cols)

            def bil(img):
                return dy * (dx * img[y1, x1] + (1 - dx) * img[y1, x2]) + (1 - dy) * (dx * img[y2, x1] + (1 - dx) * img[y2, x2])
            t = cols * int(tr) + int(tc)
            res[t] = bil(I1) - I0[r, c]
            iz = 1 / Z
            Zp = -(sp * sr * p[1] + sp * cr * p[2] + cp * p[0])
            Zr = R[2, 2] * p[1] - R[2, 1] * p[2]
            Ju = np.array([fx * iz, 0, -fx * X * iz * iz, -fx * q[1] * iz, fx * (cy * q[2] * iz - Zp * X * iz * iz),
                           fx * ((R[0, 2] * p[1] - R[0, 1] * p[2]) * iz - Zr * X * iz * iz)])
            Jv = np.array([0, fy * iz, -fy * Y * iz * iz, fy * q[0] * iz, fy * (sy * q[2] * iz - Zp * Y * iz * iz),
                           fy * ((R[1, 2] * p[1] - R[1, 1] * p[2]) * iz - Zr * Y * iz * iz)])
            J[t] = bil(Gx) * Ju + bil(Gy) * Jv
    return res, J


def pack_upper(H):
    return np.array([H[a, b] for a in range(6) for b in range(a, 6)])
