"""Test configuration.

`-m "not gpu"` (CPU container): the oracle against its golden vectors and cross-checks, host logic,
and that the C-ABI library loads and exports every symbol include/legoloam_b200.h declares.
`-m gpu` (B200 box): the parity tests proper — the CUDA path, called through the C ABI, against the
oracle on identical seeded inputs. Nothing here reads /root/reference at run time.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


# Small world so that the CPU suite stays in the minutes range; same generator as bench.py.
SMALL_WORLD = dict(n_blocks=24, n_poles=160, half_x=50, half_z=70)


@pytest.fixture(scope="session")
def orc():
    import orc as _orc
    _orc.build()
    return _orc


@pytest.fixture(scope="session")
def synth():
    from lego_loam_sr_b200 import synth as _s
    return _s


@pytest.fixture(scope="session")
def single_problem(synth):
    """BASELINE config-1 shape at reduced size: 30 map scans, one query scan, perturbed prior."""
    return synth.make_single(seed=1, n_map_scans=30, step=0.7, world_kw=SMALL_WORLD)


@pytest.fixture(scope="session")
def single_maps(single_problem, orc):
    """Oracle-built downsampled local map of single_problem."""
    d = single_problem
    kf = d["keyframes"]
    prm = orc.default_params()
    cds, sds, ncr, nsr = orc.build_local_map(d["key_poses"], [f["corner_scan"] for f in kf],
                                             [f["surf_last_ds"] for f in kf], [f["outlier_last_ds"] for f in kf], prm)
    return dict(corner_ds=cds, surf_ds=sds, n_corner_raw=ncr, n_surf_raw=nsr)


@pytest.fixture(scope="session")
def sequence(synth):
    return synth.make_sequence(seed=2, n_scans=24, step=0.7, world_kw=SMALL_WORLD)


def upstream_params(mod, **kw):
    """Headline mode: pose update applied, upstream eigen semantics, exactly 10 iterations."""
    base = dict(apply_update=1, eigen_mode=1, iter_count_thres=10, stop_thres=0.0)
    base.update(kw)
    return mod.default_params(**base)


def rot_err(a, b):
    d = np.abs(np.asarray(a[:3], np.float64) - np.asarray(b[:3], np.float64))
    d = np.minimum(d, np.abs(d - 2 * np.pi))
    return float(d.max())


def trans_err(a, b):
    return float(np.abs(np.asarray(a[3:], np.float64) - np.asarray(b[3:], np.float64)).max())
