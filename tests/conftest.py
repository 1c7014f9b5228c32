import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the in-tree shared libraries exist (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as ge
    lib = os.path.join(ROOT, "oscar_mpc_planner_mr_modification_b200", "lib", "libmpcgpu.so")
    # In the build container rebuild on source changes; on the GPU box the prebuilt .so files travel with
    # the snapshot (file times are not meaningful there), so only build what is missing.
    if os.path.isdir("/root/reference") or not os.path.exists(lib):
        ge.build_cuda()
    ge.build_oracle()
