"""Second pin for the LM loop: an independent numpy restatement of SURVEY.md A.1-A.4 (oracle/numpy_lm.py: tensor-product
Catmull-Rom weights, closed-form Jacobian, SVD least squares on the stacked [J; D]) must trace the same path as the C++
oracle (Jet autodiff, Householder QR) -- iteration by iteration.  Real Ceres is not installable here, so the Ceres boundary
itself stays unpinned (parity "partial"); this closes the gap between two separately written restatements."""
import numpy as np
import pytest

from conftest import IDENTITY, rot_angle_between


@pytest.mark.parametrize("now,loss,anchor", [
    (3, "cauchy", (4.303494, 1.425450, 31)),      # SURVEY.md A.6 anchors (a third, discarded, restatement)
    (5, "cauchy", (9.398242, 0.5893809, 25)),
    (3, "trivial", (4.422140, 1.491656, 38)),
])
def test_numpy_lm_traces_the_oracle_path(oracle, frames, now, loss, anchor):
    from oracle import numpy_lm as N
    O = oracle
    K = frames["K"]
    xyz, _ = O.get_aX(frames["bgr"][0], frames["depth"][0], K, frames["zscale"])
    dt, _ = O.get_distance_transform(frames["bgr"][now - 1])
    lt = O.LOSS_CAUCHY if loss == "cauchy" else O.LOSS_TRIVIAL
    op, os_, otr = O.solve(xyz, dt, K, IDENTITY, stride=30, options=O.default_options(loss_type=lt, loss_scale=1.0))
    npose, ns, ntr = N.solve(xyz, dt, K, IDENTITY, stride=30, loss=N.LOSS_CAUCHY if loss == "cauchy" else N.LOSS_TRIVIAL)
    # same trajectory: iteration count, every accept / reject decision, cost and radius after every iteration
    assert ns["iterations"] == os_["iterations"] and ns["accepted"] == os_["accepted"] and ns["rejected"] == os_["rejected"]
    assert ns["termination"] == os_["termination_name"]
    assert len(ntr) == len(otr)
    np.testing.assert_array_equal(ntr[:, 3], otr[:, 6])
    np.testing.assert_allclose(ntr[:, 0], otr[:, 0], rtol=1e-9)
    np.testing.assert_allclose(ntr[:, 2], otr[:, 5], rtol=1e-6)
    assert rot_angle_between(npose[:4], op[:4]) < 1e-9 and np.abs(npose[4:] - op[4:]).max() < 1e-9
    # and both sit on the survey's anchors
    c0, c1, it = anchor
    # (the anchors were made with cv2's IPP distance transform, <= 4.6e-6 off the portable one used here: DESIGN.md 2)
    assert abs(ns["initial_cost"] - c0) < 5e-5 and abs(ns["final_cost"] - c1) < 1e-5 and abs(ns["iterations"] - it) <= 1


def test_numpy_lm_huber_with_rejected_steps(oracle, frames):
    """Huber(0.1) from a perturbed start; a high min_relative_decrease makes both restatements reject the steps whose
    actual / predicted decrease falls short: exercises the reject branch (radius / 2, / 4, ..., diagonal reuse)."""
    from oracle import numpy_lm as N
    O = oracle
    K = frames["K"]
    xyz, _ = O.get_aX(frames["bgr"][1], frames["depth"][1], K, frames["zscale"])
    dt, _ = O.get_distance_transform(frames["bgr"][3])
    x0 = O.quat_plus(np.array(IDENTITY), 2.0 * np.array([0.01, -0.005, 0.008, 0.03, -0.02, 0.01]))
    rejected = 0
    for mrd in (1e-3, 1.9, 1.99):      # (outliers of a Huber loss make the Gauss-Newton model under-predict: ratios sit near 2)
        opts = O.default_options(loss_type=O.LOSS_HUBER, loss_scale=0.1, min_relative_decrease=mrd, max_num_iterations=60)
        op, os_, otr = O.solve(xyz, dt, K, x0, stride=10, options=opts)
        npose, ns, ntr = N.solve(xyz, dt, K, x0, stride=10, loss=N.LOSS_HUBER, loss_scale=0.1, min_relative_decrease=mrd, max_iterations=60)
        assert (ns["iterations"], ns["accepted"], ns["rejected"]) == (os_["iterations"], os_["accepted"], os_["rejected"])
        assert ns["termination"] == os_["termination_name"]
        np.testing.assert_array_equal(ntr[:, 3], otr[:, 6])
        np.testing.assert_allclose(ntr[:, 0], otr[:, 0], rtol=1e-8)
        np.testing.assert_allclose(ntr[:, 2], otr[:, 5], rtol=1e-5)
        rejected += os_["rejected"]
    assert rejected > 0       # otherwise this pins nothing beyond the previous test
