"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/legoloam_b200.h declares; parameter defaults equal the reference's constants; without a GPU the
product path fails loudly instead of computing anything on the CPU; host-side prior matches the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, has_gpu


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as ge
    ge.build()
    from lego_loam_sr_b200 import abi
    return abi


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "legoloam_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ll_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built):
    lib = built.load()
    syms = _declared_symbols()
    assert len(syms) >= 25, syms
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_default_params_are_the_reference_constants(built, orc):
    p = built.default_params()
    o = orc.default_params()
    device_only = {"lm_solver", "persistent_lm", "lm_max_blocks"}  # execution knobs of the CUDA library; the oracle has no such choice
    for name, _ in built.Params._fields_:
        if name not in device_only:
            assert getattr(p, name) == getattr(o, name), name
    assert p.lm_solver == 0 and p.persistent_lm == 1 and p.lm_max_blocks == 0
    # mapOptmization.cpp:92-94, :1279, :1325, :1416, :1363, :1370, :1453, :1519; loam_config.yaml:32-35
    assert (p.corner_leaf, p.surf_leaf) == (np.float32(0.2), np.float32(0.4))
    assert p.max_sq_dist == 1.0 and p.eig_ratio == 3.0 and p.plane_tol == 0.2 and p.weight_k == 0.9
    assert p.min_weight == 0.1 and p.min_residuals == 50 and p.eig_thresh == 100.0
    assert p.iter_count_thres == 200 and p.step_size == 1.0 and p.stop_thres == np.float32(0.05)
    assert p.apply_update == 0 and p.eigen_mode == 0 and p.map_downsample == 1
    assert C.sizeof(built.Params) == C.sizeof(orc.Params) + 16  # three ints, padded to the struct's 8-byte alignment


def test_invalid_arguments_are_rejected_without_a_device(built):
    lib = built.load()
    assert lib.ll_default_params(None) == -1
    h = C.c_void_p()
    bad = built.default_params(max_sq_dist=4.0)  # the 1 m grid is exact only up to 1 m
    assert lib.ll_create(C.byref(bad), 0, C.byref(h)) == -1
    assert lib.ll_create(C.byref(built.default_params()), 0, None) == -1
    assert lib.ll_destroy(None) == 0
    assert lib.ll_last_error_string(None) == b"invalid handle"


@pytest.mark.skipif(has_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback(built):
    """Without a CUDA device the product path must fail loudly (LL_ERR_CUDA), never compute on the CPU."""
    with pytest.raises(built.LegoLoamError) as e:
        built.Handle()
    assert e.value.code == -2
    from lego_loam_sr_b200 import host
    with pytest.raises(RuntimeError):
        host.MapOptimization()


def test_host_prior_matches_oracle(built, orc):
    """host/map_optimization.cpp composes transformAssociateToMap with 3x3 matrices; the oracle follows the
    reference's expanded trigonometry (:458-581). Same transform, independent derivations."""
    from lego_loam_sr_b200 import host
    rng = np.random.default_rng(2)
    for _ in range(20):
        s, b, a = (np.concatenate([rng.uniform(-0.6, 0.6, 3), rng.uniform(-50, 50, 3)]).astype(np.float32) for _ in range(3))
        t_host = host.transform_associate_to_map(s, b, a)
        t_orc = orc.transform_associate_to_map(s, b, a)
        assert np.allclose(t_host[:3], t_orc[:3], atol=5e-6), (t_host, t_orc)
        assert np.allclose(t_host[3:], t_orc[3:], atol=2e-4), (t_host, t_orc)


def test_python_binding_layouts(built):
    assert C.sizeof(built.Result) == 4 * 10 + 4 * 12
    with pytest.raises(ValueError):
        built._cloud(np.zeros((3, 5), np.float32))
    a, n, stride = built._cloud(np.zeros((3, 8)))
    assert (n, stride, a.dtype) == (3, 32, np.float32)
