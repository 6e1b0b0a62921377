"""TEST INFRASTRUCTURE -- a second, independently written CPU restatement of the pose solve, in numpy.

Purpose: pin the C++ oracle's Levenberg-Marquardt loop (oracle/ea_oracle.cpp `lm_solve`) with a separate code path, since
real Ceres cannot be built in this image (parity stays "partial").  Nothing here shares code with the C++
oracle or with the CUDA product:

  residual    vectorised Catmull-Rom tensor-product weights (not the nested CubicHermiteSpline of Ceres' header), which is
              the same polynomial:  ceres/cubic_interpolation.h as used at standalone/utils.h:77
  Jacobian    the collapsed closed form  J = [ 2 (R X) x g | g ]   instead of Jet<7> autodiff
              (standalone/utils.h:87 AutoDiffCostFunction<EAResidue,1,4,3> + QuaternionParameterization, SEA:277)
  loss        Cauchy / Huber / Trivial with the rho'' <= 0 corrector (standalone_edge_align.cpp:272)
  LM          TrustRegionMinimizer + LevenbergMarquardtStrategy per SURVEY.md A.4, with the step taken from
              numpy.linalg.lstsq (SVD) on the stacked [J_s; sqrt(diag / radius)] -- neither the oracle's Householder QR
              nor the product's 6x6 LDL^T on the normal equations
  termination order as in trust_region_minimizer.cc >= 1.13: parameter tolerance, function tolerance, then acceptance

Only tests/ import this module.
"""
import numpy as np

LOSS_TRIVIAL, LOSS_CAUCHY, LOSS_HUBER = 0, 1, 2


def rotation(q):
    """Eigen::Quaternion::toRotationMatrix (no normalisation), standalone/utils.h:51-53."""
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _weights(s):
    w = np.stack([0.5 * (-s**3 + 2 * s**2 - s), 0.5 * (3 * s**3 - 5 * s**2 + 2), 0.5 * (-3 * s**3 + 4 * s**2 + s), 0.5 * (s**3 - s**2)])
    d = np.stack([0.5 * (-3 * s**2 + 4 * s - 1), 0.5 * (9 * s**2 - 10 * s), 0.5 * (-9 * s**2 + 8 * s + 1), 0.5 * (3 * s**2 - 2 * s)])
    return w, d


def residual_jacobian(xyz, dt, K, pose7, want_J=True):
    """raw residuals r_i = DT(u_i, v_i) and the ambient-free local Jacobian (n x 6); ok=False if any |z'| < 0.01."""
    fx, fy, cx, cy = K
    R = rotation(pose7[:4])
    Y = xyz @ R.T
    p = Y + pose7[4:]
    ok = not np.any(np.abs(p[:, 2]) < 0.01)
    u = fx * p[:, 0] / p[:, 2] + cx
    v = fy * p[:, 1] / p[:, 2] + cy
    H, W = dt.shape
    iu = np.floor(u).astype(np.int64); iv = np.floor(v).astype(np.int64)
    wu, du = _weights(u - iu); wv, dv = _weights(v - iv)
    D = dt.astype(np.float64)
    f = np.zeros(len(u)); fdu = np.zeros(len(u)); fdv = np.zeros(len(u))
    for a in range(4):
        xi = np.clip(iu - 1 + a, 0, W - 1)
        for b in range(4):
            pix = D[np.clip(iv - 1 + b, 0, H - 1), xi]
            f += wu[a] * wv[b] * pix; fdu += du[a] * wv[b] * pix; fdv += wu[a] * dv[b] * pix
    if not want_J:
        return f, None, ok
    iz = 1.0 / p[:, 2]
    g = np.stack([fdu * fx * iz, fdv * fy * iz, -(fdu * fx * p[:, 0] + fdv * fy * p[:, 1]) * iz * iz], 1)
    return f, np.concatenate([2.0 * np.cross(Y, g), g], 1), ok


def robustify(r, J, loss, a):
    """cost = 1/2 sum rho(r^2);  r, J scaled by sqrt(rho') (Corrector with rho'' <= 0)."""
    s = r * r
    if loss == LOSS_CAUCHY:
        b = a * a
        rho = b * np.log1p(s / b); rp = np.maximum(np.finfo(float).tiny, 1.0 / (1.0 + s / b))
    elif loss == LOSS_HUBER:
        big = s > a * a
        ar = np.sqrt(np.where(big, s, 1.0))
        rho = np.where(big, 2 * a * ar - a * a, s); rp = np.where(big, np.maximum(np.finfo(float).tiny, a / ar), 1.0)
    else:
        rho = s; rp = np.ones_like(s)
    w = np.sqrt(rp)
    return 0.5 * rho.sum(), r * w, (J * w[:, None] if J is not None else None)


def plus(x, d):
    """QuaternionParameterization::Plus on q (delta = half-angle vector, left multiplication), t += d[3:]."""
    n = np.linalg.norm(d[:3])
    q = x[:4]
    if n > 0:
        s = np.sin(n) / n
        a0, a1, a2, a3 = np.cos(n), s * d[0], s * d[1], s * d[2]
        b0, b1, b2, b3 = q
        q = np.array([a0 * b0 - a1 * b1 - a2 * b2 - a3 * b3, a0 * b1 + a1 * b0 + a2 * b3 - a3 * b2,
                      a0 * b2 - a1 * b3 + a2 * b0 + a3 * b1, a0 * b3 + a1 * b2 - a2 * b1 + a3 * b0])
    return np.concatenate([q, x[4:] + d[3:]])


def solve(xyz, dt, K, x0, stride=30, loss=LOSS_CAUCHY, loss_scale=1.0, max_iterations=50, ftol=1e-6, gtol=1e-10, ptol=1e-8,
          radius0=1e4, max_radius=1e16, min_radius=1e-32, min_relative_decrease=1e-3, min_diag=1e-6, max_diag=1e32):
    """Returns (pose7, summary dict, trace rows [cost of the accepted iterate, cost change, radius after the update, accepted])."""
    pts = np.asarray(xyz, np.float64)[::stride]
    x = np.array(x0, np.float64)

    def evaluate(xx, want_J):
        r, J, ok = residual_jacobian(pts, dt, K, xx, want_J)
        c, rw, Jw = robustify(r, J, loss, loss_scale)
        return c, rw, Jw, ok

    cost, r, J, ok = evaluate(x, True)
    out = dict(initial_cost=cost, iterations=0, accepted=0, rejected=0, termination="", n_residuals=len(pts))
    trace = [(cost, 0.0, radius0, 1)]
    if not ok:
        out.update(termination="FAILURE_EVAL_X0", final_cost=cost)
        return x, out, np.array(trace)
    scale = 1.0 / (1.0 + np.sqrt((J * J).sum(0)))
    g = J.T @ r

    def gmax():
        return np.abs(x - plus(x, -g)).max()

    def done(term):
        out.update(termination=term, final_cost=cost)
        return x, out, np.array(trace)

    if gmax() <= gtol:
        return done("CONVERGENCE_GRADIENT")
    radius, decrease, diag, reuse, invalid = radius0, 2.0, None, False, 0
    while True:
        if out["iterations"] >= max_iterations:
            return done("NO_CONVERGENCE")
        if radius < min_radius:
            return done("CONVERGENCE_MIN_RADIUS")
        out["iterations"] += 1
        Js = J * scale
        if not reuse:
            diag = np.clip((Js * Js).sum(0), min_diag, max_diag)
        reuse = True
        A = np.vstack([Js, np.diag(np.sqrt(diag / radius))])
        y = np.linalg.lstsq(A, np.concatenate([r, np.zeros(6)]), rcond=None)[0]
        step = -y
        m = Js @ step
        model_change = -(m * (r + 0.5 * m)).sum()
        if not (np.all(np.isfinite(step)) and model_change > 0):
            invalid += 1
            if invalid >= 5:
                return done("FAILURE_INVALID_STEPS")
            radius *= 0.5; reuse = False; out["rejected"] += 1
            trace.append((cost, 0.0, radius, 0))
            continue
        invalid = 0
        xc = plus(x, step * scale)
        cc, _, _, okc = evaluate(xc, False)
        if not okc:
            cc = np.finfo(float).max
        if np.linalg.norm(x - xc) <= ptol * (np.linalg.norm(x) + ptol):
            return done("CONVERGENCE_PARAMETER")
        change = cost - cc
        if abs(change) <= ftol * cost:
            trace.append((cost, change, radius, 0))
            return done("CONVERGENCE_FUNCTION")
        rel = change / model_change if okc else -np.finfo(float).max
        if rel > min_relative_decrease:
            x = xc
            cost, r, J, _ = evaluate(x, True)
            g = J.T @ r
            radius = min(max_radius, radius / max(1.0 / 3.0, 1.0 - (2.0 * rel - 1.0) ** 3))
            decrease = 2.0; reuse = False; out["accepted"] += 1
            trace.append((cost, change, radius, 1))
            if gmax() <= gtol:
                return done("CONVERGENCE_GRADIENT")
        else:
            radius /= decrease; decrease *= 2.0; reuse = True; out["rejected"] += 1
            trace.append((cost, change, radius, 0))
