import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


def _native_built() -> bool:
    return all(os.path.exists(os.path.join(REPO, p)) for p in (
        "portrayer_b200/lib/libportrayer_gpu.so", "portrayer_b200/lib/libportrayer_host.so", "oracle/liboracle.so"))


@pytest.fixture(scope="session", autouse=True)
def native_libraries():
    """Build the in-tree libraries once if they are missing (nvcc cross-compiles without a GPU)."""
    if not _native_built():
        import __graft_entry__

        __graft_entry__.build()
    assert _native_built(), "native libraries failed to build"


@pytest.fixture(scope="session")
def gpu_ready(native_libraries):
    """Initialise the CUDA library; GPU tests FAIL (not skip) if the device path cannot run."""
    import portrayer_b200._ffi as ffi

    rc = ffi.gpu.pt_init(-1)
    assert rc == 0, f"pt_init failed on a GPU test run: {ffi.gpu.pt_last_error().decode()}"
    return True


def has_reference_assets() -> bool:
    return os.path.exists(os.path.join(REPO, "assets", "earth.jpg"))
