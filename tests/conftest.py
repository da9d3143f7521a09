import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def rt():
    import __graft_entry__ as ge
    return ge.load_package()


@pytest.fixture(scope="session")
def scene_dir(tmp_path_factory):
    return str(tmp_path_factory.mktemp("scenes"))


@pytest.fixture(scope="session")
def have_ref():
    """True when the unmodified reference was compiled into oracle/_ref (always here; prebuilt on the GPU box)."""
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_raytracer"))
