"""CPU-side checks of the drop-in boundary: libd3dsweep.so loads, exports every symbol that
include/d3d_sweep.h declares, the ctypes structs have the compiled layout, and argument validation
reports errors without touching a GPU.  No compute is launched here."""
import ctypes as C
import os
import re

import pytest
import torch

from deep3d_aerial_b200 import _lib, sweep

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "d3d_sweep.h")


@pytest.fixture(scope="module")
def lib():
    if not os.path.isfile(_lib.LIB_PATH):
        from deep3d_aerial_b200.build import build
        build()
    return _lib.load()


def _declared():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(d3d_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    names = _declared()
    assert "d3d_cost_volume" in names and "d3d_depth_regress" in names and len(names) >= 8
    for n in names:
        assert hasattr(lib, n), "libd3dsweep.so does not export " + n
    assert sorted(_lib.EXPORTS) == names, "binding and header disagree"


def test_struct_layouts_match(lib):
    for which, struct in enumerate((_lib.CostVolumeArgs, _lib.RegressArgs, _lib.SamplesArgs, _lib.FuseArgs)):
        assert lib.d3d_abi_sizeof(which) == C.sizeof(struct)
    assert lib.d3d_abi_sizeof(99) == -1
    assert lib.d3d_version() == 100


def test_header_enums_match_binding():
    text = open(HEADER).read()
    for name, val in re.findall(r"\b(D3D_[A-Z0-9_]+)\s*=\s*(\d+)", text):
        py = name[4:]
        if hasattr(_lib, py):
            assert getattr(_lib, py) == int(val), name


def test_validation_errors_need_no_gpu(lib):
    assert lib.d3d_cost_volume(None, None) == _lib.ERR_BAD_ARGUMENT
    assert b"NULL" in lib.d3d_last_error()
    a = _lib.CostVolumeArgs()
    a.struct_size = 8
    assert lib.d3d_cost_volume(C.byref(a), None) == _lib.ERR_BAD_ARGUMENT
    assert b"struct_size" in lib.d3d_last_error()
    a.struct_size = C.sizeof(a)
    a.mode = 17
    assert lib.d3d_cost_volume(C.byref(a), None) == _lib.ERR_BAD_ARGUMENT
    a.mode = _lib.AGG_VARIANCE
    a.num_views, a.channels, a.height, a.width, a.num_depth = 3, 6, 8, 8, 4
    a.feats = a.pose = a.hyps = a.out = 256        # never dereferenced: validation fails first
    assert lib.d3d_cost_volume(C.byref(a), None) == _lib.ERR_UNSUPPORTED     # C not a multiple of 4
    a.channels = 8
    a.num_views = 11
    assert lib.d3d_cost_volume(C.byref(a), None) == _lib.ERR_UNSUPPORTED     # too many views
    a.num_views = 3
    a.d_begin, a.d_count = 3, 5
    assert lib.d3d_cost_volume(C.byref(a), None) == _lib.ERR_BAD_ARGUMENT    # slice outside the sweep
    a.d_begin, a.d_count = 0, 0
    a.texel_slots = 4                              # views named as slots of a texel pool: every slot must be inside it
    a.view_slot[0], a.view_slot[1], a.view_slot[2] = 3, 0, 4
    assert lib.d3d_cost_volume(C.byref(a), None) == _lib.ERR_BAD_ARGUMENT
    assert b"view_slot[2]" in lib.d3d_last_error()
    a.texel_slots = -1
    assert lib.d3d_cost_volume(C.byref(a), None) == _lib.ERR_BAD_ARGUMENT
    r = _lib.RegressArgs()
    r.struct_size = C.sizeof(r)
    r.num_depth, r.height, r.width = 4, 4, 4
    assert lib.d3d_depth_regress(C.byref(r), None) == _lib.ERR_BAD_ARGUMENT  # NULL logits
    r.logits = r.hyps = r.depth = r.conf = 256
    r.d_begin, r.d_count = 1, 2
    assert lib.d3d_depth_regress(C.byref(r), None) == _lib.ERR_BAD_ARGUMENT  # slices need RAW_EXP
    assert b"slice" in lib.d3d_last_error()
    s = _lib.SamplesArgs()
    s.struct_size = C.sizeof(s)
    s.mode, s.num_depth, s.height, s.width = _lib.SAMPLES_AROUND, 1, 4, 4
    assert lib.d3d_depth_samples(C.byref(s), None) == _lib.ERR_BAD_ARGUMENT  # D < 2
    assert lib.d3d_nchw_to_nhwc(None, None, 4, 4, 4, None) == _lib.ERR_BAD_ARGUMENT
    assert lib.d3d_resize_bilinear(None, None, 1, 4, 4, 8, 8, None) == _lib.ERR_BAD_ARGUMENT
    assert lib.d3d_resize_bilinear(256, 256, 0, 4, 4, 8, 8, None) == _lib.ERR_BAD_ARGUMENT      # no maps


def test_cpu_tensors_are_refused(lib):
    """The product path has no CPU fallback: CPU tensors raise instead of being computed elsewhere."""
    tex = torch.zeros(3, 8, 8, 4)
    pose = torch.eye(4).repeat(2, 1, 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sweep.cost_volume(tex, pose, torch.ones(4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sweep.depth_regress(torch.zeros(4, 8, 8), torch.ones(4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sweep.to_texels([torch.zeros(4, 8, 8)])


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libd3dsweep.so"))
    with pytest.raises(ImportError, match="no CPU or PyTorch fallback"):
        _lib.load()
