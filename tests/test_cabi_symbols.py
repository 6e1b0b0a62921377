"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/ea_cabi.h
declares, the Python binding covers them all, and compute entry points fail loudly without a GPU."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "ea_cabi.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ea_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    from edge_alignment_b200 import _lib as L
    from edge_alignment_b200 import build
    build.build()
    lib = ctypes.CDLL(L.SO_PATH)
    syms = declared_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(lib, s), "libea_b200.so does not export %s" % s
        assert s in L.PROTOTYPES, "python binding lacks %s" % s
    assert sorted(L.PROTOTYPES) == syms
    assert L.lib().ea_abi_version() == 2


def test_struct_layouts_match_header():
    from edge_alignment_b200 import _lib as L
    assert ctypes.sizeof(L.FrameParams) == 8 * 4 + 5 * 8 + 2 * 4 + 2 * 8 + 2 * 4
    assert ctypes.sizeof(L.SolveParams) == 8 * 4 + 10 * 8 + 2 * 4
    assert ctypes.sizeof(L.Summary) == 6 * 4 + 2 * 8 + 2 * 4
    import edge_alignment_b200 as ea
    fp, sp = ea.frame_params(), ea.solve_params()
    # defaults reproduce edge_align_test1 (standalone_edge_align.cpp:152,160,267,272; utils.cpp:65,75,81)
    assert (fp.width, fp.height, fp.grad_threshold, fp.use_median, fp.dt_normalize) == (640, 480, 35, 1, 1)
    assert (fp.fx, fp.fy, fp.cx, fp.cy, fp.depth_scale) == (525.0, 525.0, 319.5, 239.5, 5000.0)
    assert (sp.point_stride, sp.loss_type, sp.loss_scale, sp.max_num_iterations) == (30, 1, 1.0, 50)
    assert (sp.function_tolerance, sp.gradient_tolerance, sp.parameter_tolerance) == (1e-6, 1e-10, 1e-8)
    assert sp.initial_trust_region_radius == 1e4 and sp.min_relative_decrease == 1e-3


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import edge_alignment_b200 as ea
    with pytest.raises(ea.EaError) as e:
        ea.Context(0)
    assert e.value.code == 3 and "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    """The product path must not route through the oracle (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "edge_alignment_b200")
    pat = re.compile(r"^\s*(import|from)\s+oracle|libea_oracle|ea_oracle\.|eo_[a-z_]+\(", re.M)
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                assert not pat.search(open(os.path.join(dp, f)).read()), f
    for f in os.listdir(os.path.join(ROOT, "include")):
        p = os.path.join(ROOT, "include", f)
        if os.path.isfile(p):
            assert not pat.search(open(p).read()), f
