"""Independent pins for the solver side of the oracle (the Ceres boundary itself is
UNPINNED: the reference ships no tests and Ceres cannot be built here -- DESIGN.md).

  * Jet<7> autodiff Jacobian == closed form == central differences
  * bicubic: node interpolation, quadratic reproduction, clamp-to-edge
  * an independent numpy restatement (tests/golden/make_golden.py) agrees per point
  * SURVEY.md A.6 anchors (a second, independently written restatement) reproduce
  * LM optimum agrees with scipy.optimize.least_squares(loss='cauchy')
  * README loose check (standalone/README.md:49-70)
"""
import numpy as np
import pytest

from conftest import IDENTITY, rot_angle_between


@pytest.fixture(scope="module")
def pair13(oracle, frames):
    xyz, _ = oracle.get_aX(frames["bgr"][0], frames["depth"][0], frames["K"], frames["zscale"])
    dt, _ = oracle.get_distance_transform(frames["bgr"][2])
    return xyz, dt


def test_bicubic_properties(oracle):
    O = oracle
    K = (1.0, 1.0, 0.0, 0.0)  # u = X/Z, v = Y/Z  -> sample positions are directly controllable
    rng = np.random.default_rng(1)
    H, W = 20, 30
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    # (a) nodes are interpolated exactly
    dt = rng.random((H, W)).astype(np.float32)
    pts = np.array([[x, y, 1.0] for y in range(H) for x in range(W)], float)
    raw, uv, ok = O.evaluate_raw(pts, dt, K, IDENTITY)
    np.testing.assert_array_equal(raw.reshape(H, W), dt.astype(np.float64))
    # (b) quadratics are reproduced (value and gradient) away from the borders
    a = (0.5 + 0.25 * xx - 0.125 * yy + 0.0625 * xx * yy + 0.03125 * xx * xx - 0.015625 * yy * yy).astype(np.float32)
    p = np.stack([rng.uniform(2, W - 3, 200), rng.uniform(2, H - 3, 200), np.ones(200)], 1)
    e = O.evaluate(p, a, K, IDENTITY, options=O.default_options(loss_type=O.LOSS_TRIVIAL))
    x, y = p[:, 0], p[:, 1]
    f = 0.5 + 0.25 * x - 0.125 * y + 0.0625 * x * y + 0.03125 * x * x - 0.015625 * y * y
    np.testing.assert_allclose(e["raw"], f, rtol=0, atol=1e-12)
    np.testing.assert_allclose(e["J"][:, 3], 0.25 + 0.0625 * y + 0.0625 * x, atol=1e-12)   # d/dtx == df/du (Z=1, fx=1)
    np.testing.assert_allclose(e["J"][:, 4], -0.125 + 0.0625 * x - 0.03125 * y, atol=1e-12)
    # (c) clamp to edge: far outside is a constant field with zero gradient (no bounds test in utils.h:77)
    far = np.array([[-50.0, -70.0, 1.0], [500.0, 400.0, 1.0], [-3.0, 5.0, 1.0]])
    e = O.evaluate(far, dt, K, IDENTITY, options=O.default_options(loss_type=O.LOSS_TRIVIAL))
    assert e["raw"][0] == dt[0, 0] and e["raw"][1] == dt[H - 1, W - 1] and e["raw"][2] == dt[5, 0]
    assert np.all(e["J"][:2] == 0)


def test_z_guard_fails_evaluation(oracle):
    # utils.h:70-73: |z'| < 0.01 makes the functor return false => evaluation failure
    dt = np.zeros((8, 8), np.float32)
    pts = np.array([[0.1, 0.1, 1.0], [0.0, 0.0, 0.005]])
    assert not oracle.evaluate(pts, dt, (5, 5, 4, 4), IDENTITY)["ok"]
    assert oracle.evaluate(pts[:1], dt, (5, 5, 4, 4), IDENTITY)["ok"]
    pose, s, _ = oracle.solve(pts, dt, (5, 5, 4, 4), IDENTITY, stride=1)
    assert s["termination_name"] == "FAILURE_EVAL_X0"


def test_jet_jacobian_closed_form_and_finite_differences(oracle, pair13, frames, numpy_pins):
    O = oracle
    xyz, dt = pair13
    K = frames["K"]
    x = numpy_pins["xpert"]
    opt = O.default_options(loss_type=O.LOSS_TRIVIAL)
    e = O.evaluate(xyz, dt, K, x, stride=30, options=opt)
    # independent numpy restatement with the collapsed closed form  J = [2 (R X) x g | g]
    np.testing.assert_allclose(e["raw"], numpy_pins["np_r_pert_1_3"], rtol=0, atol=2e-6)  # pins used the IPP DT (<=4.6e-6)
    np.testing.assert_allclose(e["J"], numpy_pins["np_J_pert_1_3"], rtol=0, atol=2e-3)
    # central differences through Plus() on smooth points (bicubic is only C1: skip cells crossed by the probe)
    h = 1e-6
    Jfd = np.zeros_like(e["J"])
    for k in range(6):
        d = np.zeros(6); d[k] = h
        rp, uvp, _ = O.evaluate_raw(xyz, dt, K, O.quat_plus(x, d), stride=30)
        rm, uvm, _ = O.evaluate_raw(xyz, dt, K, O.quat_plus(x, -d), stride=30)
        Jfd[:, k] = (rp - rm) / (2 * h)
        same = (np.floor(uvp) == np.floor(uvm)).all(1)
        np.testing.assert_allclose(e["J"][same, k], Jfd[same, k], rtol=1e-5, atol=5e-5)
    # robustified: Cauchy corrector scales r and J by sqrt(rho')
    ec = O.evaluate(xyz, dt, K, x, stride=30, options=O.default_options(loss_type=O.LOSS_CAUCHY, loss_scale=1.0))
    s1 = np.sqrt(1.0 / (1.0 + e["raw"] ** 2))
    np.testing.assert_allclose(ec["residuals"], e["raw"] * s1, rtol=1e-14)
    np.testing.assert_allclose(ec["J"], e["J"] * s1[:, None], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(ec["cost"], 0.5 * np.log1p(e["raw"] ** 2).sum(), rtol=1e-13)
    eh = O.evaluate(xyz, dt, K, x, stride=30, options=O.default_options(loss_type=O.LOSS_HUBER, loss_scale=0.1))
    a = 0.1; big = np.abs(e["raw"]) > a
    rho = np.where(big, 2 * a * np.abs(e["raw"]) - a * a, e["raw"] ** 2)
    np.testing.assert_allclose(eh["cost"], 0.5 * rho.sum(), rtol=1e-13)
    # sums are consistent with J, r
    np.testing.assert_allclose(ec["b"], ec["J"].T @ ec["residuals"], rtol=1e-12)
    np.testing.assert_allclose(ec["H"], ec["J"].T @ ec["J"], rtol=1e-12)


def test_survey_anchors(oracle, pair13, frames, numpy_pins):
    """values produced by the survey's own (discarded) numpy probe with cv2's IPP DT.
    Our DT is OpenCV's portable path (<=4.6e-6 normalised deviation), so digits agree to ~1e-5."""
    O = oracle
    xyz, dt = pair13
    K = frames["K"]
    e0 = O.evaluate(xyz, dt, K, IDENTITY, stride=30)
    np.testing.assert_allclose(e0["raw"][[0, 1, 2, 3, 4]], [0.151461631, 0.0, 0.025166525, 0.025166525, 0.036084317], atol=5e-6)
    assert abs(e0["cost"] - 4.303494) < 2e-5
    e = O.evaluate(xyz, dt, K, numpy_pins["xpert"], stride=30, options=O.default_options(loss_type=O.LOSS_TRIVIAL))
    assert abs(e["cost"] - 3.201015808297) < 2e-5
    np.testing.assert_allclose(e["J"][0], [-6.443053883, -13.75115282, -8.88704505, -2.135347303, 1.330478039, -0.510568932], rtol=1e-4)
    np.testing.assert_allclose(np.diag(e["H"]), [118594.438808709, 111991.583127897, 17457.560618616, 14138.913582004, 15781.791112345, 2326.285761664], rtol=1e-5)
    np.testing.assert_allclose(e["b"], [58.349804682, -282.85393069, 40.580400829, -88.92447264, -35.040528553, 1.91997738], rtol=1e-4)
    e1 = O.evaluate(xyz, dt, K, numpy_pins["xpert"], stride=1, options=O.default_options(loss_type=O.LOSS_TRIVIAL))
    assert abs(e1["cost"] - 100.5295511068) < 5e-4
    # LM trajectory anchor: 4.303494 -> 1.425450, 31 iterations, no rejections, function tolerance
    pose, s, tr = O.solve(xyz, dt, K, IDENTITY, stride=30)
    assert s["termination_name"] == "CONVERGENCE_FUNCTION" and s["rejected"] == 0 and abs(s["iterations"] - 31) <= 1
    assert abs(s["final_cost"] - 1.425450) < 5e-6
    pt, st, _ = O.solve(xyz, dt, K, IDENTITY, stride=30, options=O.default_options(loss_type=O.LOSS_TRIVIAL))
    assert abs(st["initial_cost"] - 4.422140) < 2e-5 and abs(st["final_cost"] - 1.491656) < 5e-6 and abs(st["iterations"] - 38) <= 1


def test_optimum_matches_scipy(oracle, frames, numpy_pins):
    O = oracle
    K = frames["K"]
    xyz, _ = O.get_aX(frames["bgr"][0], frames["depth"][0], K, frames["zscale"])
    for j in (3, 5):
        dt, _ = O.get_distance_transform(frames["bgr"][j - 1])
        pose, s, _ = O.solve(xyz, dt, K, IDENTITY, stride=30)
        sp = numpy_pins["scipy_pose_1_%d" % j]
        # Ceres' function-tolerance stop leaves the iterate a few 1e-5 short of the true optimum
        assert rot_angle_between(pose[:4], sp[:4]) < 1e-4
        assert np.abs(pose[4:] - sp[4:]).max() < 1e-4
        assert abs(s["final_cost"] - float(numpy_pins["scipy_cost_1_%d" % j])) < 5e-6
        # tighter tolerances converge onto scipy's optimum
        tight = O.default_options(function_tolerance=1e-14, parameter_tolerance=1e-14, max_num_iterations=500)
        pose2, s2, _ = O.solve(xyz, dt, K, IDENTITY, stride=30, options=tight)
        assert rot_angle_between(pose2[:4], sp[:4]) < 2e-5 and np.abs(pose2[4:] - sp[4:]).max() < 2e-5


def test_readme_loose_check(oracle, frames):
    """standalone/README.md:49-70 (Ceres 1.12 + OpenCV 3, pair 1->5): 8.743 -> 0.5418, 30 its,
    YPR (-0.32, 1.52, 2.50) deg, t (-0.01, 0.00, -0.05).  Loose: different library versions."""
    O = oracle
    K = frames["K"]
    xyz, _ = O.get_aX(frames["bgr"][0], frames["depth"][0], K, frames["zscale"])
    dt, _ = O.get_distance_transform(frames["bgr"][4])
    pose, s, _ = O.solve(xyz, dt, K, IDENTITY, stride=30)
    assert s["n_residuals"] == 1482 and s["termination_name"] == "CONVERGENCE_FUNCTION" and s["rejected"] == 0
    assert 20 <= s["iterations"] <= 35
    assert abs(s["initial_cost"] - 8.743202) / 8.743202 < 0.10 and abs(s["final_cost"] - 0.5418352) / 0.5418352 < 0.10
    T = O.quat_to_matrix(pose)
    R = T[:3, :3]
    yaw = np.degrees(np.arctan2(R[1, 0], R[0, 0])); pitch = np.degrees(np.arcsin(-R[2, 0])); roll = np.degrees(np.arctan2(R[2, 1], R[2, 2]))
    np.testing.assert_allclose([yaw, pitch, roll], [-0.32, 1.52, 2.50], atol=0.15)
    np.testing.assert_allclose(T[:3, 3], [-0.01, 0.00, -0.05], atol=0.006)


def test_pyramid_and_batch_entry_points(oracle, frames):
    """eo_align_pair / eo_align_batch (CPU baseline legs) agree with the staged calls."""
    O = oracle
    K = frames["K"]
    cfg = O.pair_cfg(640, 480, K, n_levels=1, stride=30)
    pose, S = O.align_pair(frames["bgr"][0], frames["depth"][0], frames["bgr"][2], cfg, IDENTITY)
    xyz, _ = O.get_aX(frames["bgr"][0], frames["depth"][0], K, frames["zscale"])
    dt, _ = O.get_distance_transform(frames["bgr"][2])
    ref, s, _ = O.solve(xyz, dt, K, IDENTITY, stride=30)
    np.testing.assert_array_equal(pose, ref)
    assert S[0]["iterations"] == s["iterations"]
    cfg3 = O.pair_cfg(640, 480, K, n_levels=3, stride=4)
    opts = O.default_options(loss_type=O.LOSS_HUBER, loss_scale=0.1)
    p3, S3 = O.align_pair(frames["bgr"][0], frames["depth"][0], frames["bgr"][2], cfg3, IDENTITY, opts)
    assert all(s["termination"] in (1, 2, 3, 4, 5) for s in S3)
    poses, sums, sec = O.align_batch(frames["bgr"], frames["depth"], [0, 0, 1], [2, 2, 3], cfg3, np.tile(IDENTITY, (3, 1)), opts, n_threads=2)
    np.testing.assert_array_equal(poses[0], p3); np.testing.assert_array_equal(poses[1], p3)
    assert sec > 0


def test_residual_variants_jet_vs_finite_differences(oracle, pair13, frames, numpy_pins):
    """EAResidueEx / SecondCam / SecondCamEx restatements (standalone/utils.h:101-421): Jet Jacobians agree with central
    differences through Plus(); the plain view equals the single-camera entry point; identity rig / zero distortion
    reduce to EAResidue."""
    O = oracle
    xyz, dt = pair13
    K = frames["K"]
    x = numpy_pins["xpert"]
    opt = O.default_options(loss_type=O.LOSS_TRIVIAL)
    base = O.evaluate(xyz, dt, K, x, stride=30, options=opt)
    ang = 0.05
    T21 = np.eye(4); T21[:3, :3] = [[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]]; T21[:3, 3] = [-0.1, 0.01, 0.02]
    dist = (0.1, -0.2, 0.001, -0.002, 0.05)
    for kw in (dict(), dict(dist=(0, 0, 0, 0, 0)), dict(T21=np.eye(4))):
        e = O.evaluate_views([dict(xyz=xyz, dt=dt, K=K, stride=30, **kw)], x, opt)
        np.testing.assert_allclose(e["raw"], base["raw"], rtol=0, atol=1e-13)
        np.testing.assert_allclose(e["J"], base["J"], rtol=0, atol=1e-9)
    for kw in (dict(dist=dist), dict(T21=T21), dict(dist=dist, T21=T21)):
        views = [dict(xyz=xyz, dt=dt, K=K, stride=30, **kw)]
        e = O.evaluate_views(views, x, opt)
        assert e["ok"] and np.abs(e["raw"] - base["raw"]).max() > 1e-3          # the variant really changes the projection
        h = 1e-6
        for k in range(6):
            d = np.zeros(6); d[k] = h
            rp = O.evaluate_views(views, O.quat_plus(x, d), opt)["raw"]; rm = O.evaluate_views(views, O.quat_plus(x, -d), opt)["raw"]
            fd = (rp - rm) / (2 * h)
            ok = np.abs(e["J"][:, k] - fd) <= 1e-5 * np.abs(fd) + 5e-5
            assert ok.mean() > 0.99          # bicubic is only C1: a few probes straddle a cell boundary
    # two cameras constrain one pose: sums add up
    v2 = [dict(xyz=xyz, dt=dt, K=K, stride=30), dict(xyz=xyz[5:], dt=dt, K=K, stride=30, T21=T21, dist=dist)]
    e2 = O.evaluate_views(v2, x, opt)
    ea_, eb_ = O.evaluate_views(v2[:1], x, opt), O.evaluate_views(v2[1:], x, opt)
    np.testing.assert_allclose(e2["sums"], ea_["sums"] + eb_["sums"], rtol=1e-12)
    pose, s = O.solve_views(v2, IDENTITY, O.default_options())
    assert s["termination"] in (1, 2, 3) and s["n_residuals"] == len(e2["raw"])


def test_dogleg_strategy(oracle, frames, numpy_pins):
    """TRUST_REGION + traditional DOGLEG (src/SolveEA.cpp:191-192).  At the Ceres default radius (1e4, in Jacobi-scaled
    units) both strategies take near Gauss-Newton steps, so they trace the same path; a small starting radius drives the
    Cauchy-point and interpolation branches, and the solve still lands on scipy's optimum."""
    O = oracle
    K = frames["K"]
    xyz, _ = O.get_aX(frames["bgr"][0], frames["depth"][0], K, frames["zscale"])
    dt, _ = O.get_distance_transform(frames["bgr"][2])
    lm, slm, _ = O.solve(xyz, dt, K, IDENTITY, stride=30)
    dl, sdl, _ = O.solve(xyz, dt, K, IDENTITY, stride=30, options=O.default_options(strategy=O.STRATEGY_DOGLEG))
    assert sdl["iterations"] == slm["iterations"] and abs(sdl["final_cost"] - slm["final_cost"]) < 1e-7
    assert rot_angle_between(lm[:4], dl[:4]) < 1e-6 and np.abs(lm[4:] - dl[4:]).max() < 1e-6
    sp = numpy_pins["scipy_pose_1_3"]
    for r in (1e-2, 1e-3):
        p, s, trace = O.solve(xyz, dt, K, IDENTITY, stride=30,
                              options=O.default_options(strategy=O.STRATEGY_DOGLEG, initial_radius=r, max_num_iterations=200))
        assert s["termination_name"] == "CONVERGENCE_FUNCTION" and s["iterations"] > slm["iterations"]
        assert rot_angle_between(p[:4], sp[:4]) < 2e-4 and np.abs(p[4:] - sp[4:]).max() < 2e-4
        costs = trace[trace[:, 6] > 0, 0]
        assert np.all(np.diff(costs) <= 0)            # accepted steps only ever decrease the cost
        assert trace[1, 3] < 0.2 * trace[-1, 5]       # first step was clipped well inside the final region
    with pytest.raises(TypeError):
        O.default_options(initial_trust_region_radius=1.0)
