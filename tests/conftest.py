import hashlib
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="session")
def frames():
    """The reference's five bundled TUM frames (standalone/rgb-d), committed as data."""
    z = np.load(os.path.join(GOLDEN, "frames.npz"))
    return dict(bgr=z["bgr"], depth=z["depth"], K=tuple(float(v) for v in z["K"]), zscale=float(z["zscale"]))


@pytest.fixture(scope="session")
def cv2_stages():
    with open(os.path.join(GOLDEN, "cv2_stages.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def numpy_pins():
    return dict(np.load(os.path.join(GOLDEN, "numpy_pins.npz")))


@pytest.fixture(scope="session")
def solver_golden():
    return dict(np.load(os.path.join(GOLDEN, "solver_golden.npz")))


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


IDENTITY = np.array([1.0, 0, 0, 0, 0, 0, 0])


def rot_angle_between(qa, qb):
    """Rotation angle (rad) between two unit quaternions (wxyz)."""
    d = abs(float(np.dot(qa / np.linalg.norm(qa), qb / np.linalg.norm(qb))))
    return 2.0 * np.arccos(min(1.0, d))
