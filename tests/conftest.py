import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "faster-voxelpose_b200")
for p in (os.path.join(ROOT, "tests"), os.path.join(PKG, "lib"), PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (B200) device; run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def built_library():
    """libfvp_b200.so (built in-tree by __graft_entry__.build(); nvcc cross-compiles without a GPU)."""
    so = os.path.join(PKG, "libfvp_b200.so")
    if not os.path.isfile(so):
        subprocess.run(["bash", os.path.join(PKG, "csrc", "build.sh"), PKG], check=True)
    return so


_GOLDEN = {}


@pytest.fixture(scope="session")
def golden():
    from golden_util import Golden

    def get(name):
        if name not in _GOLDEN:
            _GOLDEN[name] = Golden(name)
        return _GOLDEN[name]

    return get
