"""Pins the oracle's OpenCV restatements bit-exactly against genuine OpenCV outputs
(cv2 4.13, portable code path) recorded in tests/golden/cv2_stages.json for the
reference's bundled frames (standalone/utils.cpp:38-83, 201-281)."""
import numpy as np
import pytest

from conftest import sha


@pytest.mark.parametrize("i", range(5))
def test_stages_bit_exact_vs_cv2(oracle, frames, cv2_stages, i):
    O = oracle
    g = cv2_stages["frames"][i]
    bgr, depth = frames["bgr"][i], frames["depth"][i]
    blur = O.gaussian3(bgr); assert sha(blur) == g["blur"]
    gray = O.rgb2gray(blur); assert sha(gray) == g["gray"]
    lap8 = O.laplacian3_abs(gray); assert sha(lap8) == g["lap8"]
    B = np.where(lap8 > 35, 0, 255).astype(np.uint8); assert sha(B) == g["mask"]
    Bf = O.median3(B); assert sha(Bf) == g["median"]
    d = O.chamfer3_dt(Bf); assert sha(d) == g["dt_raw"]
    assert sha(O.normalize_minmax(d, 0, 1)) == g["dt_norm"]
    assert sha(O.normalize_minmax(d, 0, 255)) == g["dt_norm255"]
    assert sha(O.half_linear(bgr)) == g["half_bgr"]
    assert sha(O.half_nearest(depth)) == g["half_depth"]
    assert sha(O.box3(bgr)) == g["box3"]
    # documented: an IPP-enabled OpenCV build differs from the portable path by < 4e-4 px
    assert g["dt_raw_ipp_max_abs_dev"] < 4e-4


@pytest.mark.parametrize("i", range(5))
def test_get_aX_and_dt_entry_points(oracle, frames, cv2_stages, i):
    O = oracle
    g = cv2_stages["frames"][i]
    xyz, uvd = O.get_aX(frames["bgr"][i], frames["depth"][i], frames["K"], frames["zscale"])
    assert len(xyz) == g["n_points"] and sha(uvd.astype(np.int32)) == g["uvd"]
    fx, fy, cx, cy = frames["K"]
    Z = uvd[:, 2] / frames["zscale"]
    np.testing.assert_array_equal(xyz[:, 2], Z)
    np.testing.assert_array_equal(xyz[:, 0], (uvd[:, 0] - cx) * Z / fx)   # utils.cpp:236
    np.testing.assert_array_equal(xyz[:, 1], (uvd[:, 1] - cy) * Z / fy)   # utils.cpp:237
    dt, mask = O.get_distance_transform(frames["bgr"][i])
    assert sha(dt) == g["dt_norm"] and sha(mask) == g["median"] and int((mask == 0).sum()) == g["n_edge_now"]


def test_readme_residual_block_count(oracle, frames):
    # standalone/README.md:34-35 -- 1482 residual blocks = ceil(44458/30) on frame 1 (SEA:267)
    xyz, _ = oracle.get_aX(frames["bgr"][0], frames["depth"][0], frames["K"], frames["zscale"])
    assert len(xyz) == 44458 and (len(xyz) + 29) // 30 == 1482


def test_edge_cases_tiny_and_empty(oracle):
    O = oracle
    rng = np.random.default_rng(0)
    # no edges at all: DT saturates, normalise must not divide by zero
    flat = np.full((8, 12, 3), 77, np.uint8)
    dt, mask = O.get_distance_transform(flat)
    assert (mask == 255).all() and np.isfinite(dt).all()
    xyz, uvd = O.get_aX(flat, np.full((8, 12), 1000, np.uint16), (10, 10, 6, 4))
    assert len(xyz) == 0
    # all depth invalid -> no points even with edges
    noisy = rng.integers(0, 255, (16, 20, 3)).astype(np.uint8)
    xyz, _ = O.get_aX(noisy, np.zeros((16, 20), np.uint16), (10, 10, 10, 8))
    assert len(xyz) == 0
    # chamfer DT closed form on a single seed
    m = np.full((9, 11), 255, np.uint8); m[4, 5] = 0
    d = O.chamfer3_dt(m)
    HV, DG = round(0.955 * 65536), round(1.3693 * 65536)
    yy, xx = np.mgrid[0:9, 0:11]
    adx, ady = abs(xx - 5), abs(yy - 4)
    expect = ((HV * abs(adx - ady) + DG * np.minimum(adx, ady)).astype(np.float32) / np.float32(65536)).astype(np.float32)
    np.testing.assert_array_equal(d, expect)


@pytest.mark.parametrize("i", range(5))
def test_canny_and_exact_edt_bit_exact_vs_cv2(oracle, frames, cv2_stages, i):
    """SolveEA flavour (src/SolveEA.cpp:46,102-109) and the standalone Canny variants (utils.cpp:85-106): Canny and the
    exact Euclidean DT restatements reproduce genuine OpenCV bit for bit."""
    O = oracle
    g = cv2_stages["frames"][i]
    bgr = frames["bgr"][i]
    cc = O.canny(bgr, 150, 100, l2=True)                       # thresholds swapped inside, as OpenCV does
    assert sha(cc) == g["canny_color_l2"] and int((cc > 0).sum()) == g["n_canny_color"]
    edt = O.exact_edt(255 - cc)
    assert sha(edt) == g["edt_precise"]
    assert sha(O.normalize_minmax(edt, 0, 255)) == g["edt_precise_norm255"]
    cg = O.canny(O.rgb2gray(O.box3(bgr)), 30, 90, l2=False)
    assert sha(cg) == g["canny_gray_l1"]
    inv = np.where(cg > 127, 0, 255).astype(np.uint8)
    assert sha(O.normalize_minmax(O.chamfer3_dt(inv), 0, 1)) == g["dt2_norm"]
    # masked variants (utils.cpp:108-141,166-199): only edges under mask > 1 seed the distance transform
    mask = np.zeros((480, 640), np.uint8); mask[60:420, 80:600] = 255; mask[200:260, 300:380] = 1
    dm = O.chamfer3_dt((255 - np.where(mask > 1, cg, 0)).astype(np.uint8))
    assert sha(dm) == g["dt2_masked_raw"] and sha(O.normalize_minmax(dm, 0, 255)) == g["dt2_masked_norm255"]


def test_exact_edt_edge_cases(oracle):
    O = oracle
    m = np.full((20, 300), 255, np.uint8); m[0, 0] = 0           # columns without any zero pixel are "infinite", not capped
    yy, xx = np.mgrid[0:20, 0:300]
    np.testing.assert_array_equal(O.exact_edt(m), np.sqrt((xx ** 2 + yy ** 2).astype(np.float32)))
    assert (O.exact_edt(np.full((12, 16), 255, np.uint8)) == 65536.0).all()   # OpenCV's value for an image without zeros
    assert (O.exact_edt(np.zeros((12, 16), np.uint8)) == 0).all()
