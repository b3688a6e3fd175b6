import os
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))

GOLD = REPO / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """(Re)builds the in-tree libraries when missing or older than their sources
    (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build_host()
    g.build_cuda()
    g.build_cuda_debug()
    g.build_cli()
    g.build_oracle()


@pytest.fixture(scope="session")
def workdir(tmp_path_factory):
    from whittedstyle_raytracer_b200 import fixtures
    wd = tmp_path_factory.mktemp("scenes")
    fixtures.ensure_assets(wd)
    return wd


def load_golden_scene(workdir, name, kind="img"):
    """Scene rebuilt from the config text stored inside a golden file."""
    from whittedstyle_raytracer_b200 import Scene, fixtures
    g = np.load(GOLD / f"{kind}_{name}.npz", allow_pickle=False)
    fixtures.write_config(workdir, name, str(g["config"]))
    return Scene.from_workdir(workdir, name, bunny=bool(g["bunny"])), g


def image_diff(a, b):
    d = np.abs(a.astype(np.int64) - b.astype(np.int64)).max(axis=2)
    return dict(n=d.size, exact=int((d == 0).sum()), within1=int((d <= 1).sum()), max=int(d.max()))


def psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


def ulp_diff(a, b):
    """Distance in float32 representable steps (same-sign finite values)."""
    ai = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64)
    bi = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, -(ai & 0x7FFFFFFF), ai)
    bi = np.where(bi < 0, -(bi & 0x7FFFFFFF), bi)
    return np.abs(ai - bi)


IMG_SCENES = ["config_small", "water_small", "spheres", "parallel", "bump", "directional", "smooth"]
RAY_SCENES = ["config_small", "water_small", "spheres", "bump", "directional", "smooth"]
SOFT_SCENES = ["water_soft", "spheres_soft"]
