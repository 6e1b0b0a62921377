"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI, against the
CPU oracle on the same inputs and against the committed golden fixtures.

Tolerances (BASELINE.json north_star):
  * integer / byte / index work (masks, point lists, order, distance transform): bit exact
  * per-point residuals: |gpu - oracle| <= 1e-5 * max(|r|, R_FLOOR)  (R_FLOOR: "relative" is meaningless as r -> 0
    on an edge; 0.05 of the normalised DT range ~ 4 px)
  * converged poses: 1e-4 rad, 1e-4 m
"""
import numpy as np
import pytest

from conftest import IDENTITY, rot_angle_between, sha

pytestmark = pytest.mark.gpu
R_FLOOR = 0.05


@pytest.fixture(scope="module")
def ea():
    import edge_alignment_b200 as ea
    return ea


@pytest.fixture(scope="module")
def ctx(ea):
    c = ea.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def fs5(ea, ctx, frames):
    """The five bundled frames preprocessed on the GPU for both roles (single level, reference defaults)."""
    fs = ea.FrameSet(ctx, ea.frame_params(), 5)
    fs.preprocess_host(np.arange(5), frames["bgr"], frames["depth"], ea.ROLE_BOTH)
    yield fs
    fs.close()


# ---------------------------------------------------------------------------------------------- preprocessing
@pytest.mark.parametrize("i", range(5))
def test_preprocess_bit_exact(ea, fs5, frames, cv2_stages, oracle, i):
    g = cv2_stages["frames"][i]
    assert sha(fs5.edge_mask(i, 0, median=False)) == g["mask"]       # Laplacian>35 mask == genuine OpenCV
    assert sha(fs5.edge_mask(i, 0, median=True)) == g["median"]      # medianBlur 3
    pts = fs5.points(i)
    assert len(pts) == g["n_points"]
    assert sha(pts[:, :3].astype(np.int32)) == g["uvd"]              # ordered compaction == get_aX order
    assert np.all(pts[:, 3] == 1.0)
    dt = fs5.dt(i)
    assert sha(dt) == g["dt_norm"]                                   # chamfer DT + normalize, bit exact
    odt, _ = oracle.get_distance_transform(frames["bgr"][i])
    np.testing.assert_array_equal(dt, odt)


def test_preprocess_pyramid_and_variants(ea, ctx, frames, oracle):
    O = oracle
    fp = ea.frame_params(n_levels=3)
    fs = ea.FrameSet(ctx, fp, 2)
    try:
        fs.preprocess_host([1, 0], frames["bgr"][[0, 2]], frames["depth"][[0, 2]], ea.ROLE_BOTH)   # scrambled slots
        for slot, fi in ((1, 0), (0, 2)):
            bgr, dep = frames["bgr"][fi], frames["depth"][fi]
            for l in range(3):
                w, h, K = fs.level_geometry(l)
                assert (w, h) == (640 >> l, 480 >> l)
                s = 1.0 / (1 << l)
                np.testing.assert_allclose(K, (525 * s, 525 * s, (319.5 + .5) * s - .5, (239.5 + .5) * s - .5), rtol=0, atol=1e-12)
                xyz, uvd = O.get_aX(bgr, dep, K)
                np.testing.assert_array_equal(fs.points(slot, l)[:, :3].astype(np.int32), uvd)
                odt, omask = O.get_distance_transform(bgr)
                np.testing.assert_array_equal(fs.edge_mask(slot, l, median=True), omask)
                np.testing.assert_array_equal(fs.dt(slot, l), odt)
                bgr, dep = O.half_linear(bgr), O.half_nearest(dep)
    finally:
        fs.close()
    # no median, [0,255] normalisation, other threshold
    fs = ea.FrameSet(ctx, ea.frame_params(use_median=0, dt_normalize=ea.NORM_255, grad_threshold=60), 1)
    try:
        fs.preprocess_host([0], frames["bgr"][3:4], frames["depth"][3:4], ea.ROLE_BOTH)
        odt, omask = O.get_distance_transform(frames["bgr"][3], thresh=60, use_median=0, norm_mode=2)
        np.testing.assert_array_equal(fs.dt(0), odt)
        _, uvd = O.get_aX(frames["bgr"][3], frames["depth"][3], frames["K"], thresh=60)
        np.testing.assert_array_equal(fs.points(0)[:, :3].astype(np.int32), uvd)
    finally:
        fs.close()


def test_distance_transform_frames_side_by_side(ea, ctx, frames, oracle):
    """The coarse pyramid levels of a 640-pixel frame (320 / 160 / 80 pixels) run as 2 / 4 / 8 frames per warp in the chamfer kernel
    (segmented shuffles): an odd number of frames in scrambled slots leaves the last warp's trailing segments empty, and every
    frame, level and segment position must still match the oracle bit for bit."""
    O = oracle
    order = [4, 0, 3, 1, 2, 0, 2]                          # 7 frames (with repeats) -> segments 0..6 of the level-3 warp
    slots = [3, 6, 0, 5, 1, 4, 2]
    fs = ea.FrameSet(ctx, ea.frame_params(n_levels=4), len(order))
    try:
        fs.preprocess_host(slots, frames["bgr"][order], frames["depth"][order], ea.ROLE_NOW)
        for slot, fi in zip(slots, order):
            bgr = frames["bgr"][fi]
            for l in range(4):
                odt, omask = O.get_distance_transform(bgr)
                np.testing.assert_array_equal(fs.edge_mask(slot, l, median=True), omask, err_msg="mask frame %d level %d" % (fi, l))
                np.testing.assert_array_equal(fs.dt(slot, l), odt, err_msg="dt frame %d level %d" % (fi, l))
                bgr = O.half_linear(bgr)
    finally:
        fs.close()


def test_preprocess_two_lane_pipeline_on_big_batches(ea, ctx, oracle):
    """From 8 frames per SM the distance transforms of the first half of a batch run on an auxiliary stream beside the edge
    kernels of the second half (ea_launch_preprocess).  An odd, scrambled batch of small frames: the frames on either side of
    the split, and the first and last ones, must match the oracle bit for bit -- masks, distance transforms, point lists."""
    O = oracle
    import torch
    n_sm = torch.cuda.get_device_properties(0).multi_processor_count
    n = 8 * n_sm + 3
    w, h = 64, 48
    rng = np.random.default_rng(11)
    K = (50.0, 50.0, w / 2.0, h / 2.0)
    fp = ea.frame_params(width=w, height=h, fx=K[0], fy=K[1], cx=K[2], cy=K[3], max_points=w * h, n_levels=2)
    fs = ea.FrameSet(ctx, fp, n)
    try:
        base = (rng.integers(0, 4, (16, h, w, 3)) * 64).astype(np.uint8)             # 16 distinct blocky images ...
        shift = rng.integers(0, 16, n)
        bgr = np.stack([np.roll(base[i % 16], int(shift[i]), axis=1) for i in range(n)])   # ... each frame its own shift
        dep = (rng.integers(0, 3, (n, h, w)) * 700).astype(np.uint16)
        slots = rng.permutation(n).astype(np.int32)
        fs.preprocess_host(slots, bgr, dep, ea.ROLE_BOTH)
        half = n // 2
        for i in (0, 1, half - 1, half, half + 1, n - 2, n - 1):
            s = int(slots[i])
            b, d = bgr[i], dep[i]
            for l in range(2):
                Kl = fs.level_geometry(l)[2]
                odt, omask = O.get_distance_transform(b)
                np.testing.assert_array_equal(fs.edge_mask(s, l, median=True), omask, err_msg="mask frame %d level %d" % (i, l))
                np.testing.assert_array_equal(fs.dt(s, l), odt, err_msg="dt frame %d level %d" % (i, l))
                _, uvd = O.get_aX(b, d, Kl)
                assert fs.num_points(s, l) == len(uvd)
                np.testing.assert_array_equal(fs.points(s, l)[:, :3].astype(np.int32).reshape(-1, 3), uvd.reshape(-1, 3))
                b, d = O.half_linear(b), O.half_nearest(d)
    finally:
        fs.close()


def test_preprocess_edge_cases(ea, ctx, oracle):
    O = oracle
    rng = np.random.default_rng(3)
    for (w, h) in ((64, 48), (40, 24), (352, 16), (1280, 32)):   # ragged widths (not multiples of 32), multi-chunk rows
        fp = ea.frame_params(width=w, height=h, fx=50.0, fy=50.0, cx=w / 2.0, cy=h / 2.0, max_points=w * h)
        fs = ea.FrameSet(ctx, fp, 3)
        try:
            bgr = np.stack([rng.integers(0, 256, (h, w, 3)), np.full((h, w, 3), 90), rng.integers(0, 256, (h, w, 3))]).astype(np.uint8)
            bgr[2] = (bgr[2] // 64) * 64
            dep = rng.integers(0, 3, (3, h, w)).astype(np.uint16) * 700   # many invalid (zero) depths
            fs.preprocess_host([0, 1, 2], bgr, dep, ea.ROLE_BOTH)
            for s in range(3):
                odt, omask = O.get_distance_transform(bgr[s])
                np.testing.assert_array_equal(fs.edge_mask(s, 0, median=True), omask)
                np.testing.assert_array_equal(fs.dt(s), odt)
                _, uvd = O.get_aX(bgr[s], dep[s], (50.0, 50.0, w / 2.0, h / 2.0))
                assert fs.num_points(s) == len(uvd)
                np.testing.assert_array_equal(fs.points(s)[:, :3].astype(np.int32).reshape(-1, 3), uvd.reshape(-1, 3))
            assert fs.num_points(1) == 0 and np.all(fs.dt(1) == 0)     # flat image: no edges, DT saturates then normalises to 0
        finally:
            fs.close()
    # capacity overflow is reported, not silent -- per (slot, level), and only while the truncated list is the current one
    fp = ea.frame_params(width=64, height=48, max_points=100)
    fs = ea.FrameSet(ctx, fp, 2)
    try:
        bgr = rng.integers(0, 256, (1, 48, 64, 3)).astype(np.uint8)
        flat = np.full((1, 48, 64, 3), 90, np.uint8)
        dep = np.full((1, 48, 64), 1000, np.uint16)
        fs.preprocess_host([0], bgr, dep, ea.ROLE_BOTH)
        fs.preprocess_host([1], flat, dep, ea.ROLE_BOTH)
        with pytest.raises(ea.EaError) as e:
            fs.num_points(0)
        assert e.value.code == 4
        assert fs.num_points(1) == 0                       # the other slot fits: not poisoned by slot 0
        sp = ea.solve_params(point_stride=1)
        with pytest.raises(ea.EaError) as e:               # a solve that uses the truncated list says so ...
            ctx.solve_batch(fs, [0], fs, [0], None, sp)
        assert e.value.code == 4
        ctx.solve_batch(fs, [1], fs, [0], None, sp)        # ... one that does not is unaffected
        fs.preprocess_host([0], flat, dep, ea.ROLE_BOTH)   # the slot is re-preprocessed with a frame that fits: flag cleared
        assert fs.num_points(0) == 0
        ctx.solve_batch(fs, [0], fs, [1], None, sp)
    finally:
        fs.close()


@pytest.mark.parametrize("w", [8, 33, 100, 255, 256, 257, 383, 384, 385, 510, 639, 640, 641, 700])
def test_chamfer_dt_widths(ea, ctx, oracle, w):
    """Every strip-length switch of the warp-per-frame DT kernel (8 / 12 / 20 pixels per lane), partial strips, widths that
    are not multiples of 4 (scalar row I/O) and the hand-over to the block kernel above 640 px -- sparse and dense edges,
    two pyramid levels, raw values (no normalisation) bit for bit against the oracle."""
    O = oracle
    rng = np.random.default_rng(w)
    h = 38
    fp = ea.frame_params(width=w, height=h, fx=50.0, fy=50.0, cx=w / 2.0, cy=h / 2.0, dt_normalize=ea.NORM_NONE, use_median=0,
                         n_levels=2 if (w >= 16 and w % 2 == 0) else 1)   # pyramid levels need even sizes
    fs = ea.FrameSet(ctx, fp, 3)
    try:
        sparse = np.full((h, w, 3), 70, np.uint8)
        for _ in range(3):                                       # a few isolated bright blocks: long propagation distances
            y, x = rng.integers(0, h - 2), rng.integers(0, w - 2)
            sparse[y:y + 2, x:x + 2] = 255
        dense = ((rng.integers(0, 4, (h, w, 3)) * 80).astype(np.uint8))
        one = np.full((h, w, 3), 70, np.uint8); one[h - 1, w - 1] = 255     # a single seed in the last corner
        bgr = np.stack([sparse, dense, one])
        fs.preprocess_host([0, 1, 2], bgr, None, ea.ROLE_NOW)
        for s_ in range(3):
            odt, _ = O.get_distance_transform(bgr[s_], use_median=False, norm_mode=0)
            np.testing.assert_array_equal(fs.dt(s_), odt)
            if fp.n_levels == 2:
                odt1, _ = O.get_distance_transform(O.half_linear(bgr[s_]), use_median=False, norm_mode=0)
                np.testing.assert_array_equal(fs.dt(s_, 1), odt1)
    finally:
        fs.close()


# ------------------------------------------------------------------------------------------------- evaluation
def _assert_residual_parity(gpu_raw, ref_raw):
    tol = 1e-5 * np.maximum(np.abs(ref_raw), R_FLOOR)
    bad = np.abs(gpu_raw - ref_raw) > tol
    assert not bad.any(), "max excess %g" % (np.abs(gpu_raw - ref_raw) / tol).max()


@pytest.mark.parametrize("a,b", [(1, 3), (1, 5), (3, 1), (4, 2)])
@pytest.mark.parametrize("stride", [30, 1])
def test_eval_parity(ea, ctx, fs5, frames, oracle, numpy_pins, a, b, stride):
    O = oracle
    K = frames["K"]
    xyz, _ = O.get_aX(frames["bgr"][a - 1], frames["depth"][a - 1], K)
    dt, _ = O.get_distance_transform(frames["bgr"][b - 1])
    for pose in (IDENTITY, numpy_pins["xpert"]):
        for loss, scale in ((ea.LOSS_TRIVIAL, 1.0), (ea.LOSS_CAUCHY, 1.0), (ea.LOSS_HUBER, 0.1)):
            sp = ea.solve_params(point_stride=stride, loss_type=loss, loss_scale=scale)
            g = ctx.eval(fs5, a - 1, fs5, b - 1, pose, sp)
            o = O.evaluate(xyz, dt, K, pose, stride=stride, options=O.default_options(loss_type=loss, loss_scale=scale))
            assert g["failed"] == 0 and g["n_residuals"] == len(o["raw"])
            _assert_residual_parity(g["raw"], o["raw"])
            _assert_residual_parity(g["residuals"], o["residuals"])
            jscale = np.abs(o["J"]).max(0)
            assert (np.abs(g["J"] - o["J"]) <= 2e-5 * jscale + 2e-5 * np.abs(o["J"])).all()
            # 27 normal-equation sums + cost through the production reduction tree
            np.testing.assert_allclose(g["cost"], o["cost"], rtol=2e-6)
            np.testing.assert_allclose(g["H"], o["H"], rtol=0, atol=2e-6 * np.abs(np.diag(o["H"])).max())
            np.testing.assert_allclose(g["b"], o["b"], rtol=0, atol=2e-6 * (np.abs(o["J"]) * np.abs(o["residuals"])[:, None]).sum(0).max())


def test_unfloored_residual_error_distribution(ea, ctx, fs5, frames, oracle, numpy_pins, capsys):
    """north_star: per-point residuals within 1e-5 RELATIVE of the fp64 evaluation.  The other parity tests use a floor
    (|delta| <= 1e-5 max(|r|, 0.05)); this one reports and bounds the UNFLOORED distribution on all 20 ordered pairs of the bundled
    frames, every edge point, at a perturbed pose: the fraction of points within 1e-5 relative, the largest absolute error
    among the small residuals (|r| < 0.05) and the largest relative error among the rest."""
    O = oracle
    K = frames["K"]
    pose = numpy_pins["xpert"]
    sp = ea.solve_params(point_stride=1, loss_type=ea.LOSS_TRIVIAL)
    n_all = 0; n_in = 0; worst_abs_small = 0.0; worst_rel_big = 0.0; worst_abs = 0.0; n_zero = 0
    edges = [0.0, 1e-7, 1e-6, 1e-5, 1e-4, 1e-3, np.inf]
    hist = np.zeros(len(edges) - 1, np.int64)
    for a in range(5):
        xyz, _ = O.get_aX(frames["bgr"][a], frames["depth"][a], K)
        for b in range(5):
            if a == b:
                continue
            dt, _ = O.get_distance_transform(frames["bgr"][b])
            g = ctx.eval(fs5, a, fs5, b, pose, sp, want_jac=False)
            o = O.evaluate(xyz, dt, K, pose, stride=1, options=O.default_options(loss_type=O.LOSS_TRIVIAL))
            ref = o["raw"]; err = np.abs(g["raw"] - ref)
            nz = np.abs(ref) > 0
            rel = err[nz] / np.abs(ref[nz])
            hist += np.histogram(rel, bins=edges)[0]
            n_zero += int((~nz).sum()); worst_abs = max(worst_abs, float(err.max()))
            n_all += int(nz.sum()); n_in += int((rel <= 1e-5).sum())
            small = np.abs(ref) < 0.05
            worst_abs_small = max(worst_abs_small, float(err[small].max()))
            big = nz & ~small
            worst_rel_big = max(worst_rel_big, float((err[big] / np.abs(ref[big])).max()))
            assert np.all(err[~nz] <= 2e-9)          # exact zeros of the reference (on-edge texels): within fp32 noise of the field
    frac = n_in / n_all
    with capsys.disabled():
        print("\n[unfloored residual parity, 20 pairs, %d points (+%d exact zeros)] within 1e-5 relative: %.4f %%; relative-error histogram %s over bins %s; "
              "max |err| for |r| < 0.05: %.3g; max relative error for |r| >= 0.05: %.3g; max |err| overall: %.3g"
              % (n_all, n_zero, 100 * frac, hist.tolist(), edges, worst_abs_small, worst_rel_big, worst_abs))
    # fp32 interpolation of a field normalised to [0, 1]: the absolute error is bounded by a few fp32 ulps of the local texel
    # magnitude, so relative error only grows where the residual itself is tiny
    assert frac > 0.99
    assert worst_rel_big < 1e-5
    assert worst_abs_small < 5e-8


def test_eval_bypass_hooks_and_xyz_mode(ea, ctx, frames, oracle, numpy_pins):
    """Oracle-made inputs through ea_frameset_set_points / set_dt: pixel stream and fp32 XYZ stream."""
    O = oracle
    K = frames["K"]
    xyz, uvd = O.get_aX(frames["bgr"][0], frames["depth"][0], K)
    dt, _ = O.get_distance_transform(frames["bgr"][2])
    fs = ea.FrameSet(ctx, ea.frame_params(), 2)
    try:
        fs.set_dt(0, dt); fs.set_dt(1, dt)
        fs.set_points(0, ea.pixel_points(uvd), mode=ea.POINTS_PIXEL)
        p4 = np.ones((len(xyz), 4), np.float32); p4[:, :3] = xyz
        fs.set_points(1, p4, mode=ea.POINTS_XYZ)
        sp = ea.solve_params(point_stride=7, loss_type=ea.LOSS_TRIVIAL)
        o = O.evaluate(xyz, dt, K, numpy_pins["xpert"], stride=7, options=O.default_options(loss_type=0))
        g = ctx.eval(fs, 0, fs, 0, numpy_pins["xpert"], sp)
        _assert_residual_parity(g["raw"], o["raw"])
        # XYZ points are rounded to fp32 (6e-8 relative): compare against the oracle on the same rounded points
        o32 = O.evaluate(p4[:, :3].astype(np.float64), dt, K, numpy_pins["xpert"], stride=7, options=O.default_options(loss_type=0))
        g = ctx.eval(fs, 1, fs, 0, numpy_pins["xpert"], sp)
        _assert_residual_parity(g["raw"], o32["raw"])
    finally:
        fs.close()


def test_points_mode_returns_to_pixel_after_preprocess(ea, ctx, frames, oracle, numpy_pins):
    """A slot that held caller-supplied XYZ points is a pixel-point slot again once it is preprocessed as a reference frame."""
    O = oracle
    K = frames["K"]
    fs = ea.FrameSet(ctx, ea.frame_params(), 2)
    try:
        fs.preprocess_host([0, 1], frames["bgr"][[0, 2]], frames["depth"][[0, 2]], ea.ROLE_BOTH)
        want = ctx.eval(fs, 0, fs, 1, numpy_pins["xpert"], ea.solve_params(point_stride=30))["raw"].copy()
        xyz, _ = O.get_aX(frames["bgr"][0], frames["depth"][0], K)
        fs.set_points(0, np.concatenate([xyz, np.ones((len(xyz), 1))], 1), mode=ea.POINTS_XYZ)
        fs.preprocess_host([0], frames["bgr"][:1], frames["depth"][:1], ea.ROLE_REF)
        got = ctx.eval(fs, 0, fs, 1, numpy_pins["xpert"], ea.solve_params(point_stride=30))["raw"]
        np.testing.assert_array_equal(got, want)
    finally:
        fs.close()


def test_eval_out_of_image_and_z_guard(ea, ctx, oracle):
    O = oracle
    rng = np.random.default_rng(5)
    w, h = 64, 48
    K = (40.0, 40.0, 31.5, 23.5)
    fs = ea.FrameSet(ctx, ea.frame_params(width=w, height=h, fx=K[0], fy=K[1], cx=K[2], cy=K[3], depth_scale=1000.0, max_points=w * h), 1)
    try:
        dt = rng.random((h, w)).astype(np.float32)
        fs.set_dt(0, dt)
        # points that project far outside, onto the border band, and behind the camera (no bounds test in utils.h:77)
        p4 = np.ones((400, 4), np.float32)
        p4[:, 0] = rng.uniform(-3, 3, 400); p4[:, 1] = rng.uniform(-3, 3, 400); p4[:, 2] = rng.uniform(0.5, 2.0, 400)
        p4[::7, 2] *= -1
        fs.set_points(0, p4, mode=ea.POINTS_XYZ)
        sp = ea.solve_params(point_stride=1, loss_type=ea.LOSS_TRIVIAL)
        pose = np.array([0.9987502603949663, 0.03, -0.02, 0.03, 0.1, -0.1, 0.2]); pose[:4] /= np.linalg.norm(pose[:4])
        g = ctx.eval(fs, 0, fs, 0, pose, sp)
        o = O.evaluate(p4[:, :3].astype(np.float64), dt, K, pose, stride=1, options=O.default_options(loss_type=0))
        assert g["failed"] == 0 and o["ok"]
        np.testing.assert_allclose(g["raw"], o["raw"], rtol=0, atol=2e-5)   # random field: slope ~1/px, fp64 frac keeps it tight
        # |z'| < 0.01 => evaluation failure (utils.h:70-73)
        p4[5, :3] = (0.1, 0.1, 0.005 - pose[6])
        fs.set_points(0, p4, mode=ea.POINTS_XYZ)
        g = ctx.eval(fs, 0, fs, 0, np.array([1, 0, 0, 0, 0.1, -0.1, pose[6]]), sp)
        assert g["failed"] == 1
        poses, S = ctx.solve_batch(fs, [0], fs, [0], [np.array([1, 0, 0, 0, 0.1, -0.1, pose[6]])], sp)
        assert S[0][0]["termination_name"] == "FAILURE_EVAL_X0"
    finally:
        fs.close()


# ------------------------------------------------------------------------------------------------------ solve
def _check_solution(pose, summ, gpose, gsum, it_slack=2):
    assert rot_angle_between(pose[:4], gpose[:4]) < 1e-4, "rotation off by %g rad" % rot_angle_between(pose[:4], gpose[:4])
    assert np.abs(pose[4:] - gpose[4:]).max() < 1e-4
    assert abs(summ["final_cost"] - gsum[1]) <= 1e-5 * gsum[1]
    assert abs(summ["initial_cost"] - gsum[0]) <= 1e-5 * gsum[0]
    assert abs(summ["iterations"] - int(gsum[2])) <= it_slack
    assert summ["termination"] == int(gsum[5])


@pytest.mark.parametrize("loss", ["cauchy", "trivial"])
def test_solve_all_bundled_pairs_vs_golden(ea, ctx, fs5, solver_golden, loss):
    """All 20 ordered pairs of the bundled frames in ONE batched launch (stride 30, Ceres defaults;
    edge_align_test1 is the pair 1->3) against the oracle's frozen results."""
    pairs = [(a, b) for a in range(5) for b in range(5) if a != b]
    sp = ea.solve_params(point_stride=30, loss_type=ea.LOSS_CAUCHY if loss == "cauchy" else ea.LOSS_TRIVIAL)
    poses, S = ctx.solve_batch(fs5, [p[0] for p in pairs], fs5, [p[1] for p in pairs], None, sp)
    for i, (a, b) in enumerate(pairs):
        key = "%d_%d_%s" % (a + 1, b + 1, loss)
        _check_solution(poses[i], S[i][0], solver_golden["pose_" + key], solver_golden["summary_" + key])
        assert S[i][0]["n_residuals"] == (fs5.num_points(a) + 29) // 30


@pytest.mark.parametrize("cluster", [1, 2, 4, 8])
def test_solve_cluster_sizes_and_stride1(ea, ctx, fs5, solver_golden, cluster):
    sp = ea.solve_params(point_stride=1, cluster_size=cluster)
    poses, S = ctx.solve_batch(fs5, [0], fs5, [2], None, sp)
    _check_solution(poses[0], S[0][0], solver_golden["pose_1_3_cauchy_stride1"], solver_golden["summary_1_3_cauchy_stride1"])
    assert S[0][0]["n_residuals"] == 44458
    sp30 = ea.solve_params(point_stride=30, cluster_size=cluster)
    poses, S = ctx.solve_batch(fs5, [0, 0, 3], fs5, [2, 4, 1], None, sp30)
    for i, key in enumerate(["1_3_cauchy", "1_5_cauchy", "4_2_cauchy"]):
        _check_solution(poses[i], S[i][0], solver_golden["pose_" + key], solver_golden["summary_" + key])


def test_solve_matches_live_oracle_pyramid_huber(ea, ctx, frames, oracle):
    """3-level coarse-to-fine with Huber loss (the bench configuration) against the oracle run live."""
    O = oracle
    K = frames["K"]
    fs = ea.FrameSet(ctx, ea.frame_params(n_levels=3), 5)
    try:
        fs.preprocess_host(np.arange(5), frames["bgr"], frames["depth"], ea.ROLE_BOTH)
        sp = ea.solve_params(point_stride=1, loss_type=ea.LOSS_HUBER, loss_scale=0.1)
        pairs = [(0, 2), (0, 4), (2, 0), (3, 4)]
        poses, S = ctx.solve_batch(fs, [p[0] for p in pairs], fs, [p[1] for p in pairs], None, sp)
        cfg = O.pair_cfg(640, 480, K, n_levels=3, stride=1)
        opts = O.default_options(loss_type=O.LOSS_HUBER, loss_scale=0.1)
        for i, (a, b) in enumerate(pairs):
            op, oS = O.align_pair(frames["bgr"][a], frames["depth"][a], frames["bgr"][b], cfg, IDENTITY, opts)
            assert rot_angle_between(poses[i][:4], op[:4]) < 1e-4 and np.abs(poses[i][4:] - op[4:]).max() < 1e-4
            for l in range(3):
                assert S[i][l]["n_residuals"] == oS[l]["n_residuals"]
                assert abs(S[i][l]["iterations"] - oS[l]["iterations"]) <= 2
                assert abs(S[i][l]["final_cost"] - oS[l]["final_cost"]) <= 2e-5 * oS[l]["final_cost"]
    finally:
        fs.close()


def test_solve_edge_cases(ea, ctx, frames):
    fs = ea.FrameSet(ctx, ea.frame_params(), 2)
    try:
        flat = np.full((1, 480, 640, 3), 128, np.uint8)
        fs.preprocess_host([0], flat, np.full((1, 480, 640), 5000, np.uint16), ea.ROLE_BOTH)
        fs.preprocess_host([1], frames["bgr"][:1], frames["depth"][:1], ea.ROLE_BOTH)
        # no reference points => level skipped, pose untouched
        start = np.array([1.0, 0, 0, 0, 0.01, 0.02, 0.03])
        poses, S = ctx.solve_batch(fs, [0], fs, [1], [start], ea.solve_params())
        assert S[0][0]["termination_name"] == "SKIPPED_NO_POINTS" and np.array_equal(poses[0], start)
        # a now-frame without edges: constant DT, zero gradient => immediate gradient convergence
        poses, S = ctx.solve_batch(fs, [1], fs, [0], None, ea.solve_params())
        assert S[0][0]["termination_name"] == "CONVERGENCE_GRADIENT" and np.array_equal(poses[0], IDENTITY)
        # max_num_iterations is honoured
        poses, S = ctx.solve_batch(fs, [1], fs, [1], [np.array([1.0, 0, 0, 0, 0.02, 0, 0])], ea.solve_params(max_num_iterations=3))
        assert S[0][0]["iterations"] <= 3
        # argument errors
        with pytest.raises(ea.EaError):
            ctx.solve_batch(fs, [0], fs, [5], None, ea.solve_params())
        with pytest.raises(ea.EaError):
            ctx.solve_batch(fs, [0], fs, [1], [np.array([2.0, 0, 0, 0, 0, 0, 0])], ea.solve_params())
        with pytest.raises(ea.EaError):
            ctx.solve_batch(fs, [0], fs, [1], None, ea.solve_params(point_stride=0))
    finally:
        fs.close()


@pytest.mark.parametrize("radius", [1e4, 1e-2, 1e-3])
@pytest.mark.parametrize("cluster", [1, 4])
def test_solve_dogleg_matches_oracle(ea, ctx, fs5, frames, oracle, radius, cluster):
    """trust_region_strategy = DOGLEG (src/SolveEA.cpp:192) against the oracle's DoglegStrategy restatement; the small
    radii force the Cauchy-point and interpolated steps, the default one is the Gauss-Newton branch."""
    O = oracle
    K = frames["K"]
    pairs = [(0, 2), (0, 4), (3, 1)]
    sp = ea.solve_params(point_stride=30, trust_region_strategy=ea.STRATEGY_DOGLEG, initial_trust_region_radius=radius,
                         max_num_iterations=200, cluster_size=cluster)
    poses, S = ctx.solve_batch(fs5, [p[0] for p in pairs], fs5, [p[1] for p in pairs], None, sp)
    opts = O.default_options(strategy=O.STRATEGY_DOGLEG, initial_radius=radius, max_num_iterations=200)
    for i, (a, b) in enumerate(pairs):
        xyz, _ = O.get_aX(frames["bgr"][a], frames["depth"][a], K, frames["zscale"])
        dt, _ = O.get_distance_transform(frames["bgr"][b])
        op, oS, _ = O.solve(xyz, dt, K, IDENTITY, stride=30, options=opts)
        assert rot_angle_between(poses[i][:4], op[:4]) < 1e-4 and np.abs(poses[i][4:] - op[4:]).max() < 1e-4
        assert abs(S[i][0]["iterations"] - oS["iterations"]) <= 2 and S[i][0]["termination"] == oS["termination"]
        assert abs(S[i][0]["final_cost"] - oS["final_cost"]) <= 1e-5 * oS["final_cost"]
        assert S[i][0]["rejected"] == oS["rejected"]
    with pytest.raises(ea.EaError):
        ctx.solve_batch(fs5, [0], fs5, [1], None, ea.solve_params(trust_region_strategy=2))


@pytest.mark.parametrize("strategy,radius", [(0, 1e4), (0, 1e-2), (1, 1e-2)])
def test_iteration_log_matches_oracle(ea, ctx, fs5, frames, oracle, strategy, radius):
    """ea_solve_traced: the per-iteration log (cost of the accepted iterate, trust-region radius, accept / reject) follows the
    oracle's iteration by iteration -- a much stronger statement than agreeing at the optimum.  Small starting radii make
    the radius sequence non-trivial (LM damping, dogleg Cauchy / interpolated steps)."""
    O = oracle
    K = frames["K"]
    for a, b in ((0, 2), (0, 4)):
        xyz, _ = O.get_aX(frames["bgr"][a], frames["depth"][a], K, frames["zscale"])
        dt, _ = O.get_distance_transform(frames["bgr"][b])
        opts = O.default_options(strategy=strategy, initial_radius=radius, max_num_iterations=200)
        op, oS, otr = O.solve(xyz, dt, K, IDENTITY, stride=30, options=opts)
        sp = ea.solve_params(point_stride=30, trust_region_strategy=strategy, initial_trust_region_radius=radius, max_num_iterations=200)
        pose, S, tr = ctx.solve_traced(fs5, a, fs5, b, None, sp)
        assert rot_angle_between(pose[:4], op[:4]) < 1e-4 and np.abs(pose[4:] - op[4:]).max() < 1e-4
        n = min(len(tr), len(otr)) - 1                      # the last record is the terminating evaluation on both sides
        assert n >= 10 and abs(len(tr) - len(otr)) <= 2
        assert tr[0][5] == -1 and abs(tr[0][2] - otr[0][0]) <= 1e-6 * otr[0][0] and tr[0][4] == radius
        np.testing.assert_allclose(tr[1:n, 2], otr[1:n, 0], rtol=2e-6)          # cost after every iteration
        np.testing.assert_array_equal(tr[1:n, 5], otr[1:n, 6])                   # same accept / reject decisions
        np.testing.assert_allclose(tr[1:n, 4], otr[1:n, 5], rtol=1e-3)          # same trust-region radius sequence
        assert np.all(tr[:, 0] == 0) and np.array_equal(tr[1:n, 1], np.arange(1, n))


# ----------------------------------------------------------------------------------------- kernels / tracker
@pytest.mark.parametrize("kernel", [1, 2, 4, 8])
def test_solve_kernel_variants_agree(ea, ctx, fs5, solver_golden, kernel):
    """CTA-per-pair (1) and cluster (2, 4, 8) launches are the same solver: same poses, costs, iterations."""
    pairs = [(a, b) for a in range(5) for b in range(5) if a != b]
    sp = ea.solve_params(point_stride=3, cluster_size=kernel)
    poses, S = ctx.solve_batch(fs5, [p[0] for p in pairs], fs5, [p[1] for p in pairs], None, sp)
    sp1 = ea.solve_params(point_stride=3, cluster_size=1)
    ref, R = ctx.solve_batch(fs5, [p[0] for p in pairs], fs5, [p[1] for p in pairs], None, sp1)
    for i in range(len(pairs)):
        assert rot_angle_between(poses[i][:4], ref[i][:4]) < 2e-5 and np.abs(poses[i][4:] - ref[i][4:]).max() < 2e-5
        assert abs(S[i][0]["iterations"] - R[i][0]["iterations"]) <= 2
        assert abs(S[i][0]["final_cost"] - R[i][0]["final_cost"]) <= 1e-5 * R[i][0]["final_cost"]
    poses2, _ = ctx.solve_batch(fs5, [p[0] for p in pairs], fs5, [p[1] for p in pairs], None, sp)
    assert np.array_equal(poses, poses2)      # fixed reduction order: bitwise reproducible


def test_tracker_matches_pairwise_solves_and_pipelined_host_path(ea, ctx, frames, oracle):
    """Frame-to-keyframe tracker (2 streams, key frame every 2 frames) == the same solves issued pair by pair with the
    oracle, and the pipelined host path (submit t, wait t-1) == the synchronous one."""
    import torch
    O = oracle
    K = frames["K"]
    order = [[0, 1, 2, 3, 4], [4, 3, 2, 1, 0]]                      # two streams walking the bundled frames
    fp = ea.frame_params(n_levels=2)
    sp = ea.solve_params(point_stride=4, loss_type=ea.LOSS_HUBER, loss_scale=0.1)
    seq_b = np.stack([np.stack([frames["bgr"][order[s][t]] for s in range(2)]) for t in range(5)])     # [T,S,h,w,3]
    seq_d = np.stack([np.stack([frames["depth"][order[s][t]] for s in range(2)]) for t in range(5)])
    hb = torch.from_numpy(seq_b).pin_memory(); hd = torch.from_numpy(seq_d).pin_memory()
    fb, fd = seq_b[0].nbytes, seq_d[0].nbytes
    tr = ea.Tracker(ctx, fp, sp, 2, keyframe_interval=2)
    try:
        sync_poses = []
        for t in range(5):
            poses, _ = tr.step_host(hb.data_ptr() + t * fb, hd.data_ptr() + t * fd, fetch=True)
            sync_poses.append(poses.copy())
        # oracle replay of the same schedule
        cfg = O.pair_cfg(640, 480, K, n_levels=2, stride=4)
        opts = O.default_options(loss_type=O.LOSS_HUBER, loss_scale=0.1)
        for s in range(2):
            key, pose = 0, IDENTITY.copy()
            for t in range(1, 5):
                a, b = order[s][key], order[s][t]
                pose, _ = O.align_pair(frames["bgr"][a], frames["depth"][a], frames["bgr"][b], cfg, pose, opts)
                g = sync_poses[t][s]
                assert rot_angle_between(g[:4], pose[:4]) < 1e-4 and np.abs(g[4:] - pose[4:]).max() < 1e-4, (s, t)
                if t % 2 == 0:
                    key, pose = t, IDENTITY.copy()
        # pipelined submission gives identical results
        tr.reset()
        got = {}
        for t in range(5):
            tr.step_host(hb.data_ptr() + t * fb, hd.data_ptr() + t * fd, fetch=False)
            if t >= 1:
                got[t - 1] = tr.wait(t - 1)[0].copy()
        got[4] = tr.wait(4)[0].copy()
        for t in range(1, 5):
            np.testing.assert_array_equal(got[t], sync_poses[t])
        with pytest.raises(ea.EaError):
            tr.wait(1)          # fell out of the 2-deep ring
    finally:
        tr.close()


def test_tracker_overlap_modes_agree(ea, ctx):
    """Device frames declared complete (frame t+1 preprocessed on its own stream while frame t is aligned, three slot sets
    per camera rotating through key-frame switches) give bit-identical poses to the stream-ordered default, over several
    key-frame intervals and with every frame a key frame."""
    import sys, os
    import torch
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import synth
    bgr, depth, _ = synth.make_sequences(6, 14, seed=5, device="cuda")
    torch.cuda.synchronize()
    fb, fd = bgr[0].numel(), depth[0].numel() * 2
    for interval in (1, 2, 5):
        runs = []
        for ready in (False, True):
            tr = ea.Tracker(ctx, ea.frame_params(n_levels=2), ea.solve_params(point_stride=2, loss_type=ea.LOSS_HUBER, loss_scale=0.1), 6, interval)
            try:
                tr.set_inputs_ready(ready)
                out = []
                for t in range(14):
                    tr.step_device(bgr.data_ptr() + t * fb, depth.data_ptr() + t * fd)
                    if ready and t % 3 != 0:
                        continue                      # let several steps queue up behind each other
                    out.append((t, tr.poses()[0].copy()))
                runs.append(dict(out))
            finally:
                tr.close()
        for t, p in runs[1].items():
            np.testing.assert_array_equal(p, runs[0][t])


def test_tracker_on_synthetic_sequence_recovers_ground_truth(ea, ctx):
    """Synthetic ray-cast sequence with known camera motion: the tracked keyframe_T_frame poses follow the ground truth
    (edge alignment is a pixel-level method: a few mm / tenths of a degree on this scene)."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import synth
    S, T = 3, 7
    bgr, depth, traj = synth.make_sequences(S, T, seed=7, device="cuda")
    fp = ea.frame_params(n_levels=3)
    sp = ea.solve_params(point_stride=1, loss_type=ea.LOSS_HUBER, loss_scale=0.1)
    tr = ea.Tracker(ctx, fp, sp, S, keyframe_interval=10)
    try:
        fb, fd = 640 * 480 * 3 * S, 640 * 480 * 2 * S
        for t in range(T):
            tr.step_device(bgr.data_ptr() + t * fb, depth.data_ptr() + t * fd)
            if t == 0:
                continue
            poses, Ss = tr.poses()
            for s in range(S):
                gt = synth.relative_pose(traj[s][0], traj[s][1], 0, t)
                assert np.degrees(rot_angle_between(poses[s][:4], gt[:4])) < 1.0, (s, t)
                assert np.abs(poses[s][4:] - gt[4:]).max() < 0.03, (s, t)
    finally:
        tr.close()


# --------------------------------------------------------------------- residual variants (standalone/utils.h:101-421)
def test_residual_variants_and_two_camera_solve(ea, ctx, fs5, frames, oracle, numpy_pins):
    """EAResidueEx (distortion), EAResidueSecondCam (rig), EAResidueSecondCamEx (both) against the oracle's Jet
    restatement: per-point residuals, Jacobians, normal equations, and a two-camera solve of one pose."""
    O = oracle
    K = frames["K"]
    x = numpy_pins["xpert"]
    ang = 0.05
    T21 = np.eye(4); T21[:3, :3] = [[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]]; T21[:3, 3] = [-0.1, 0.01, 0.02]
    dist = (0.1, -0.2, 0.001, -0.002, 0.05)
    pts = {i: O.get_aX(frames["bgr"][i], frames["depth"][i], K)[0] for i in (0, 1)}
    dts = {i: O.get_distance_transform(frames["bgr"][i])[0] for i in (2, 3)}
    for stride in (30, 1):
        for kw in (dict(), dict(dist=dist), dict(T21=T21), dict(dist=dist, T21=T21)):
            for loss, scale in ((ea.LOSS_TRIVIAL, 1.0), (ea.LOSS_HUBER, 0.1)):
                sp = ea.solve_params(point_stride=stride, loss_type=loss, loss_scale=scale)
                g = ctx.eval_views([dict(ref=fs5, ref_slot=0, now=fs5, now_slot=2, **kw)], x, sp)
                o = O.evaluate_views([dict(xyz=pts[0], dt=dts[2], K=K, stride=stride, **kw)], x, O.default_options(loss_type=loss, loss_scale=scale))
                assert g["failed"] == 0 and o["ok"] and g["n_residuals"] == len(o["raw"])
                _assert_residual_parity(g["raw"], o["raw"])
                jscale = np.abs(o["J"]).max(0)
                assert (np.abs(g["J"] - o["J"]) <= 3e-5 * jscale + 3e-5 * np.abs(o["J"])).all()
                np.testing.assert_allclose(g["cost"], o["cost"], rtol=3e-6)
                np.testing.assert_allclose(g["H"], o["H"], rtol=0, atol=3e-6 * np.abs(np.diag(o["H"])).max())
    # two cameras, one pose (SEA:791-803): camera 1 plain, camera 2 through the rig with distortion
    sp = ea.solve_params(point_stride=10)
    gv = [dict(ref=fs5, ref_slot=0, now=fs5, now_slot=2), dict(ref=fs5, ref_slot=1, now=fs5, now_slot=3, T21=T21, dist=dist)]
    ov = [dict(xyz=pts[0], dt=dts[2], K=K, stride=10), dict(xyz=pts[1], dt=dts[3], K=K, stride=10, T21=T21, dist=dist)]
    g = ctx.eval_views(gv, x, sp); o = O.evaluate_views(ov, x, O.default_options())
    _assert_residual_parity(g["raw"], o["raw"])
    np.testing.assert_allclose(g["cost"], o["cost"], rtol=3e-6)
    pose, s = ctx.solve_views(gv, None, sp)
    op, os_ = O.solve_views(ov, IDENTITY, O.default_options())
    assert rot_angle_between(pose[:4], op[:4]) < 1e-4 and np.abs(pose[4:] - op[4:]).max() < 1e-4
    assert abs(s["iterations"] - os_["iterations"]) <= 2 and s["n_residuals"] == os_["n_residuals"]
    assert abs(s["final_cost"] - os_["final_cost"]) <= 1e-5 * os_["final_cost"]
    # a single plain view through the views API == the batched solver
    p1, s1 = ctx.solve_views(gv[:1], None, ea.solve_params(point_stride=30))
    p2, s2 = ctx.solve_batch(fs5, [0], fs5, [2], None, ea.solve_params(point_stride=30))
    assert rot_angle_between(p1[:4], p2[0][:4]) < 2e-6 and np.abs(p1[4:] - p2[0][4:]).max() < 2e-6


# ----------------------------------------------------- Canny + exact Euclidean DT (src/SolveEA.cpp:46,102-109 flavour)
@pytest.mark.parametrize("i", [0, 2, 4])
def test_canny_and_exact_edt_bit_exact(ea, ctx, frames, cv2_stages, oracle, i):
    O = oracle
    g = cv2_stages["frames"][i]
    # SolveEA flavour: Canny(colour, 150, 100, 3, L2), exact DT of the inverted map, normalised to [0, 255]
    fp = ea.frame_params(edge_detector=ea.EDGE_CANNY_COLOR, canny_low=150.0, canny_high=100.0, canny_l2=1,
                         dt_kind=ea.DT_EXACT, dt_normalize=ea.NORM_255, max_points=640 * 480)
    fs = ea.FrameSet(ctx, fp, 1)
    try:
        fs.preprocess_host([0], frames["bgr"][i:i + 1], frames["depth"][i:i + 1], ea.ROLE_BOTH)
        mask = fs.edge_mask(0, 0, median=False)                         # 0 on edges (reference convention)
        assert sha((255 - mask).astype(np.uint8)) == g["canny_color_l2"]
        assert sha(fs.dt(0)) == g["edt_precise_norm255"]
        cc = O.canny(frames["bgr"][i], 150, 100, l2=True)
        vs, us = np.nonzero((cc > 0) & (frames["depth"][i] > 0))
        np.testing.assert_array_equal(fs.points(0)[:, :3].astype(np.int32), np.stack([us, vs, frames["depth"][i][vs, us]], 1))
    finally:
        fs.close()
    # standalone Canny variant: blur 3x3 -> gray -> Canny(30, 90) L1, chamfer DT normalised to [0, 1]  (utils.cpp:85-106, 371-462)
    fp = ea.frame_params(edge_detector=ea.EDGE_CANNY_GRAY, canny_low=30.0, canny_high=90.0, canny_l2=0, max_points=640 * 480)
    fs = ea.FrameSet(ctx, fp, 1)
    try:
        fs.preprocess_host([0], frames["bgr"][i:i + 1], frames["depth"][i:i + 1], ea.ROLE_BOTH)
        assert sha((255 - fs.edge_mask(0, 0)).astype(np.uint8)) == g["canny_gray_l1"]
        assert sha(fs.dt(0)) == g["dt2_norm"]
        assert fs.num_points(0) == g["n_points_canny"] and sha(fs.points(0)[:, :3].astype(np.int32)) == g["uvd_canny"]
    finally:
        fs.close()


def test_solveea_flavour_pipeline_matches_oracle(ea, ctx, frames, oracle):
    """Canny(colour) + exact DT [0,255] + no loss + every point + 25 iterations (src/SolveEA.cpp:124-198, with LM instead of
    DOGLEG) on a 2-level pyramid and ragged sizes, against the oracle fed with the same stages."""
    O = oracle
    K = frames["K"]
    fp = ea.frame_params(edge_detector=ea.EDGE_CANNY_COLOR, canny_low=150.0, canny_high=100.0, canny_l2=1,
                         dt_kind=ea.DT_EXACT, dt_normalize=ea.NORM_255, max_points=640 * 480, n_levels=2)
    fs = ea.FrameSet(ctx, fp, 2)
    try:
        fs.preprocess_host([0, 1], frames["bgr"][[0, 2]], frames["depth"][[0, 2]], ea.ROLE_BOTH)
        # level 1 of the pyramid through the same kernels
        hb, hd = O.half_linear(frames["bgr"][0]), O.half_nearest(frames["depth"][0])
        c1 = O.canny(hb, 150, 100, l2=True)
        vs, us = np.nonzero((c1 > 0) & (hd > 0))
        np.testing.assert_array_equal(fs.points(0, 1)[:, :3].astype(np.int32), np.stack([us, vs, hd[vs, us]], 1))
        e1 = O.normalize_minmax(O.exact_edt(255 - O.canny(O.half_linear(frames["bgr"][2]), 150, 100, l2=True)), 0, 255)
        np.testing.assert_array_equal(fs.dt(1, 1), e1)
        # level-0 solve, SolveEA settings
        cc = O.canny(frames["bgr"][0], 150, 100, l2=True)
        vs, us = np.nonzero((cc > 0) & (frames["depth"][0] > 0))
        Z = frames["depth"][0][vs, us] / 5000.0
        xyz = np.stack([(us - K[2]) * Z / K[0], (vs - K[3]) * Z / K[1], Z], 1)
        dt = O.normalize_minmax(O.exact_edt(255 - O.canny(frames["bgr"][2], 150, 100, l2=True)), 0, 255)
        opts = O.default_options(loss_type=O.LOSS_TRIVIAL, max_num_iterations=25)
        op, os_, _ = O.solve(xyz, dt, K, IDENTITY, stride=1, options=opts)
        sp = ea.solve_params(point_stride=1, loss_type=ea.LOSS_TRIVIAL, max_num_iterations=25, coarsest_level=0)
        poses, S = ctx.solve_batch(fs, [0], fs, [1], None, sp)
        assert rot_angle_between(poses[0][:4], op[:4]) < 1e-4 and np.abs(poses[0][4:] - op[4:]).max() < 1e-4
        assert S[0][0]["n_residuals"] == len(xyz) and abs(S[0][0]["iterations"] - os_["iterations"]) <= 2
    finally:
        fs.close()
    rng = np.random.default_rng(2)
    for (w, h) in ((72, 40), (200, 24)):
        fp = ea.frame_params(width=w, height=h, fx=50.0, fy=50.0, cx=w / 2.0, cy=h / 2.0, edge_detector=ea.EDGE_CANNY_GRAY,
                             canny_low=30.0, canny_high=90.0, dt_kind=ea.DT_EXACT, dt_normalize=ea.NORM_NONE, max_points=w * h)
        fs = ea.FrameSet(ctx, fp, 2)
        try:
            bgr = (rng.integers(0, 4, (2, h, w, 3)) * 80).astype(np.uint8); bgr[1] = 60       # random blocks / flat image
            dep = rng.integers(0, 2, (2, h, w)).astype(np.uint16) * 900
            fs.preprocess_host([0, 1], bgr, dep, ea.ROLE_BOTH)
            for s_ in range(2):
                cg = O.canny(O.rgb2gray(O.box3(bgr[s_])), 30, 90, l2=False)
                np.testing.assert_array_equal(255 - fs.edge_mask(s_, 0), cg)
                np.testing.assert_array_equal(fs.dt(s_), O.exact_edt(255 - cg))
        finally:
            fs.close()


def test_masked_reference_points(ea, ctx, frames, oracle):
    """get_aX_mask (utils.cpp:283-369): point list restricted to mask > 0, same order; pyramid levels sample the mask."""
    O = oracle
    rng = np.random.default_rng(4)
    mask = np.zeros((480, 640), np.uint8); mask[100:400, 150:500] = 7; mask[rng.integers(0, 480, 5000), rng.integers(0, 640, 5000)] = 0
    fs = ea.FrameSet(ctx, ea.frame_params(n_levels=2), 1)
    try:
        fs.preprocess_masked([0], frames["bgr"][:1], frames["depth"][:1], mask[None], ea.ROLE_BOTH)
        _, uvd = O.get_aX(frames["bgr"][0], frames["depth"][0], frames["K"])
        keep = mask[uvd[:, 1], uvd[:, 0]] > 0
        np.testing.assert_array_equal(fs.points(0)[:, :3].astype(np.int32), uvd[keep])
        hb, hd = O.half_linear(frames["bgr"][0]), O.half_nearest(frames["depth"][0])
        _, uvd1 = O.get_aX(hb, hd, fs.level_geometry(1)[2])
        keep1 = mask[uvd1[:, 1] * 2, uvd1[:, 0] * 2] > 0
        np.testing.assert_array_equal(fs.points(0, 1)[:, :3].astype(np.int32), uvd1[keep1])
        odt, _ = O.get_distance_transform(frames["bgr"][0])
        np.testing.assert_array_equal(fs.dt(0), odt)              # the mask does not touch the now role
    finally:
        fs.close()


def test_masked_distance_transform(ea, ctx, frames, cv2_stages, oracle):
    """get_distance_transform2_masked / _masked_NoNormalize (utils.cpp:108-141,166-199): blur -> gray -> Canny(30,90) ->
    edges * (mask > 1) -> chamfer DT -> [0,255] or raw.  Bit-exact against the oracle stages and the cv2 golden hashes."""
    O = oracle
    mask = np.zeros((480, 640), np.uint8); mask[60:420, 80:600] = 255; mask[200:260, 300:380] = 1   # 1 is NOT > 1: masked out
    for norm, key in ((ea.NORM_255, "dt2_masked_norm255"), (ea.NORM_NONE, "dt2_masked_raw")):
        fp = ea.frame_params(edge_detector=ea.EDGE_CANNY_GRAY, canny_low=30.0, canny_high=90.0, canny_l2=0, dt_normalize=norm, n_levels=2)
        fs = ea.FrameSet(ctx, fp, 2)
        try:
            fs.preprocess_now_masked([1, 0], frames["bgr"][[0, 2]], np.stack([mask, mask]))
            for slot, i in ((1, 0), (0, 2)):
                cg = O.canny(O.rgb2gray(O.box3(frames["bgr"][i])), 30, 90, l2=False)
                e = np.where(mask > 1, cg, 0).astype(np.uint8)
                d = O.chamfer3_dt((255 - e).astype(np.uint8))
                want = O.normalize_minmax(d, 0, 255) if norm == ea.NORM_255 else d
                np.testing.assert_array_equal(fs.dt(slot), want)
                assert sha(fs.dt(slot)) == cv2_stages["frames"][i][key]
                hb = O.half_linear(frames["bgr"][i])
                c1 = O.canny(O.rgb2gray(O.box3(hb)), 30, 90, l2=False)
                d1 = O.chamfer3_dt((255 - np.where(mask[::2, ::2] > 1, c1, 0)).astype(np.uint8))
                np.testing.assert_array_equal(fs.dt(slot, 1), O.normalize_minmax(d1, 0, 255) if norm == ea.NORM_255 else d1)
        finally:
            fs.close()


def test_float_depth_and_zero_depth_policy(ea, ctx, frames, oracle):
    """SolveEA's depth convention (src/SolveEA.cpp:27,68-69): CV_32F metres, edge pixels without depth kept at Z = 1."""
    O = oracle
    K = frames["K"]
    depth_m = (frames["depth"][0].astype(np.float32) / np.float32(5000.0)).astype(np.float32)
    fp = ea.frame_params(depth_type=ea.DEPTH_F32, zero_depth_to_one=1, max_points=640 * 480, n_levels=2)
    fs = ea.FrameSet(ctx, fp, 2)
    try:
        fs.preprocess_host([0], frames["bgr"][:1], depth_m[None], ea.ROLE_BOTH)
        fs.preprocess_host([1], frames["bgr"][2:3], None, ea.ROLE_NOW)
        lap = O.laplacian3_abs(O.rgb2gray(O.gaussian3(frames["bgr"][0])))
        vs, us = np.nonzero(lap > 35)                                   # every edge pixel is kept
        z = depth_m[vs, us].copy(); z[z == 0] = 1.0
        pts = fs.points(0)
        np.testing.assert_array_equal(pts[:, 0], us.astype(np.float32)); np.testing.assert_array_equal(pts[:, 1], vs.astype(np.float32))
        np.testing.assert_array_equal(pts[:, 2], z)                      # metres, bit for bit
        # residuals against the oracle on X = Z (u - cx) / fx built from the same float depths
        Z = z.astype(np.float64)
        xyz = np.stack([(us - K[2]) * Z / K[0], (vs - K[3]) * Z / K[1], Z], 1)
        dt, _ = O.get_distance_transform(frames["bgr"][2])
        pose = np.array([0.99995, 0.006, -0.004, 0.005, 0.01, -0.01, 0.02]); pose[:4] /= np.linalg.norm(pose[:4])
        sp = ea.solve_params(point_stride=3, loss_type=ea.LOSS_TRIVIAL)
        g = ctx.eval(fs, 0, fs, 1, pose, sp)
        o = O.evaluate(xyz, dt, K, pose, stride=3, options=O.default_options(loss_type=0))
        _assert_residual_parity(g["raw"], o["raw"])
        # level 1: nearest-decimated float depth
        hb = O.half_linear(frames["bgr"][0]); hd = depth_m[::2, ::2]
        lap1 = O.laplacian3_abs(O.rgb2gray(O.gaussian3(hb)))
        vs1, us1 = np.nonzero(lap1 > 35)
        z1 = hd[vs1, us1].copy(); z1[z1 == 0] = 1.0
        np.testing.assert_array_equal(fs.points(0, 1)[:, 2], z1)
    finally:
        fs.close()
