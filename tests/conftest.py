import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # a clean checkout has no built extension (the .so files are git-ignored): build it once (nvcc cross-compiles
    # sm_100a without a GPU; about a minute), through build.py loaded by path - the package cannot be imported yet
    import importlib.util
    import sysconfig

    pkg = ROOT / "loco_hd_b200"
    if not (pkg / "liblocohd_b200.so").exists() or not (ROOT / "loco_hd" / ("loco_hd" + sysconfig.get_config_var("EXT_SUFFIX"))).exists():
        spec = importlib.util.spec_from_file_location("_locohd_build", pkg / "build.py")
        b = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(b)
        b.build_all()


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def gpu_ctx():
    """One C-ABI context on cuda:0; fails (does not skip) when the CUDA library or device is missing."""
    from loco_hd_b200 import _capi
    ctx = _capi.Context(0)
    yield ctx
    ctx.close()
