import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_lib():
    """The plain-C restatement, built on demand (gcc only)."""
    import subprocess
    from oracle import cpudrv
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True, capture_output=True)
    return cpudrv.OracleTransport()


@pytest.fixture(scope="session")
def ref_lib():
    """The unmodified reference + harness; only where it was built (needs /root/reference at build time)."""
    from oracle import cpudrv
    if not cpudrv.have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    return cpudrv.RefTransport()


@pytest.fixture(scope="session")
def gpu():
    from ompmc_b200 import build
    build.build()
    from ompmc_b200.api import GpuTransport
    return GpuTransport(0)
