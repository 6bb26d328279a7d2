"""apod_np.py — CPU restatement of the reference's apodization generators (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product path
(qups_b200/) never does.

Restates src/UltrasoundSystem.m:4892-5429 (apScanline, apMultiline, apTranslatingAperture, apApertureGrowth,
apTxParallelogram, apAcceptanceAngle, apCosineAngle) for explicit geometry arrays:

    Pi   3 x I1 x I2 x I3 pixel positions (Scan.positions(), src/Scan.m:194)
    Pn   3 x N element positions, nn 3 x N element normals, ae N element angles in degrees
         (Transducer.positions / orientations, src/TransducerArray.m:95-108)
    Pf   3 x M foci (Sequence.focus)

Two modes per generator:
  * literal=True  — the reference's expressions as written, in float64 (MATLAB's default numeric type);
  * literal=False — the canonical fp32 sequence the device functions in qups_b200/csrc/apod_fused.cuh evaluate
                    (one individually rounded float32 operation per step), which the GPU must match bit for bit.
The two differ only for pixels within rounding distance of a mask boundary (tests bound that set).
Outputs are MATLAB-shaped: receive weights I1 x I2 x I3 x N, transmit weights I1 x I2 x I3 x 1 x M.
Parity unpinned: MATLAB is absent, so these follow the reference's source text, not its executed output.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def _sind(x):
    return np.sin(np.deg2rad(x))


def _cosd(x):
    return np.cos(np.deg2rad(x))


def _norm3_f32(x, y, z):
    q = x * x
    q = q + y * y
    q = q + z * z
    return np.sqrt(q)


def _dircos(Pi, Pn, nn, literal):
    """r = Pi - Pn; r = r ./ vecnorm(r,2,1); pagemtimes(n','r)  (src/UltrasoundSystem.m:5356-5363, :5414-5417)."""
    dt = np.float64 if literal else f32
    P = np.asarray(Pi, dt)[..., None]                       # 3 x I1 x I2 x I3 x 1
    E = np.asarray(Pn, dt).reshape(3, 1, 1, 1, -1)          # 3 x 1 x 1 x 1 x N
    n = np.asarray(nn, dt).reshape(3, 1, 1, 1, -1)
    r = P - E
    with np.errstate(invalid="ignore", divide="ignore"):
        if literal:
            r = r / np.sqrt((r * r).sum(0, keepdims=True))
            return (n * r).sum(0)
        d = _norm3_f32(r[0], r[1], r[2])
        q = n[0] * (r[0] / d)
        q = q + n[1] * (r[1] / d)
        q = q + n[2] * (r[2] / d)
        return q


def apAcceptanceAngle(Pi, Pn, nn, theta=45.0, literal=True):
    """apod = r >= cosd(theta)   (:5303-5375)."""
    c = _dircos(Pi, Pn, nn, literal)
    thr = _cosd(float(theta)) if literal else f32(_cosd(float(theta)))
    with np.errstate(invalid="ignore"):
        return (c >= thr).astype(np.float64 if literal else f32)


def apCosineAngle(Pi, Pn, nn, theta=45.0, literal=True):
    """apod = cosd(min(90, (90/theta) * acosd(max(-1, min(1, r)))))   (:5377-5429)."""
    c = _dircos(Pi, Pn, nn, literal)
    if literal:
        c = np.fmax(-1.0, np.fmin(1.0, c))  # MATLAB min/max ignore NaN
        return _cosd(np.minimum(90.0, (90.0 / float(theta)) * np.rad2deg(np.arccos(c))))
    c = np.fmax(f32(-1), np.fmin(f32(1), c)).astype(f32)
    t = f32(90.0 / float(theta)) * np.arccos(c).astype(f32)
    return np.where(t >= f32(np.pi / 2), f32(0), np.cos(t).astype(f32)).astype(f32)


def apApertureGrowth(Pi, Pn, ae=None, f=1.5, Dmax=np.inf, literal=True):
    """apod = (z > f*abs(2*d)) .* (abs(2*d) < Dmax)   (:5165-5267); non-planar when any(ae)."""
    dt = np.float64 if literal else f32
    P = np.asarray(Pi, dt)[..., None]
    E = np.asarray(Pn, dt).reshape(3, 1, 1, 1, -1)
    Xi, Zi, Xn, Zn = P[0], P[2], E[0], E[2]
    nonplanar = ae is not None and np.any(np.asarray(ae) != 0)
    if nonplanar:
        a = np.asarray(ae, np.float64).reshape(1, 1, 1, -1)
        if literal:
            rp = np.hypot(Xi - Xn, Zi - Zn)
            ap = np.rad2deg(np.arctan2(Xi - Xn, Zi - Zn))
            d = rp * _sind(ap - a)
            z = np.abs(rp * _cosd(ap - a))
        else:  # the same rotation written algebraically: rp*sind(ap-ae) = dx*cosd(ae) - dz*sind(ae), ...
            ca, sa = _cosd(a).astype(f32), _sind(a).astype(f32)
            dx, dz = Xi - Xn, Zi - Zn
            d = dx * ca - dz * sa
            z = np.abs(dz * ca + dx * sa)
    else:
        d = Xn - Xi
        z = Zi + 0 * d
    a2d = np.abs(dt(2) * d)
    return ((z > dt(f) * a2d) & (a2d < dt(Dmax))).astype(dt)


def _lateral(Pi, lat, lat_dim, dt):
    """xi = swapdim(us.scan.x, 2, xdim) (ScanCartesian) or scan.a along adim (ScanPolar)."""
    Isz = np.asarray(Pi).shape[1:]
    if lat is None:
        return np.asarray(Pi, dt)[0]
    shp = [1, 1, 1]
    shp[lat_dim - 1] = -1
    return np.broadcast_to(np.asarray(lat, dt).reshape(shp), Isz)


def apScanline(Pi, xv, tol, lat=None, lat_dim=2, literal=True):
    """apod = abs(xi - xv) < tol   (:4892-4968) -> I1 x I2 x I3 x 1 x M."""
    dt = np.float64 if literal else f32
    xi = _lateral(Pi, lat, lat_dim, dt)[..., None, None]
    return (np.abs(xi - np.asarray(xv, dt).reshape(1, 1, 1, 1, -1)) < dt(tol)).astype(dt)


def apTranslatingAperture(Pi, xv, xn, tol, lat=None, lat_dim=2, literal=True):
    """apod = abs(xi - xv) <= tol(1) & abs(xi - xn) <= tol(end)   (:5074-5163) -> I1 x I2 x I3 x N x M."""
    dt = np.float64 if literal else f32
    tol = np.atleast_1d(np.asarray(tol, np.float64))
    xi = _lateral(Pi, lat, lat_dim, dt)[..., None, None]
    a = np.abs(xi - np.asarray(xv, dt).reshape(1, 1, 1, 1, -1)) <= dt(tol[0])
    b = np.abs(xi - np.asarray(xn, dt).reshape(1, 1, 1, -1, 1)) <= dt(tol[-1])
    return (a & b).astype(dt)


def apTxParallelogram(Pi, theta, phi, bounds, literal=True):
    """pg - nv .* (pg_z ./ nv_z) projected to z = 0, inside the transducer bounds for any of the two tilts (:5269-5301)."""
    dt = np.float64 if literal else f32
    P = np.asarray(Pi, dt)[..., None]                        # 3 x I x 1(M)
    th = np.asarray(theta, np.float64).reshape(1, 1, 1, -1)
    phi = np.atleast_1d(np.asarray(phi, np.float64))
    phi = np.array([phi[0], phi[-1]])
    lo, hi = dt(bounds[0]), dt(bounds[1])
    xs = []
    for p in phi:
        nx, nz = _sind(p + th).astype(dt), _cosd(p + th).astype(dt)
        xs.append(P[0] - nx * (P[2] / nz))
    a = (lo < xs[0]) | (lo < xs[1])
    b = (xs[0] <= hi) | (xs[1] <= hi)
    return (a & b).astype(dt)[:, :, :, None, :]


def apMultiline(x, xv):
    """Linear weights between the two transmits straddling each scan line (:4970-5072); returns I_lat x M."""
    x, xv = np.asarray(x, np.float64).reshape(-1), np.asarray(xv, np.float64).reshape(-1)
    A = np.zeros((x.size, xv.size))
    for i, xi in enumerate(x):
        da = xi - xv
        l = np.nonzero(da >= 0)[0]
        r = np.nonzero(da <= 0)[0]
        if l.size == 0 or r.size == 0:
            continue
        li, ri = l[-1], r[0]                                # find(.,1,'last') / find(.,1,'first')
        dlr = abs(xv[li] - xv[ri])
        if dlr == 0:
            al, ar = 1.0, 0.0
        else:
            al, ar = 1 - abs(xv[li] - xi) / dlr, 1 - abs(xv[ri] - xi) / dlr
        A[i, li] += al
        A[i, ri] += ar
    return A
