import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200 box)")


@pytest.fixture(scope="session", autouse=True)
def built():
    """Build what is missing.  The oracle and host tools need only g++; the CUDA library is
    cross-compiled by nvcc (no GPU needed) and normally prebuilt by __graft_entry__.build()."""
    need = []
    if not os.path.exists(os.path.join(ROOT, "build", "grb-synth")):
        need.append("host-tools")
    if not os.path.exists(os.path.join(ROOT, "oracle", "_build", "libgrb_oracle.so")):
        need.append("oracle")
    if not os.path.exists(os.path.join(ROOT, "goldrush_b200", "_lib", "libgoldrush_b200.so")):
        need.append("lib")
    if not os.path.exists(os.path.join(ROOT, "build", "goldrush-path")):
        need.append("goldrush-path")
    if not os.path.exists(os.path.join(ROOT, "build", "goldpolish-index")):
        need.append("goldpolish-index")
    if need:
        subprocess.check_call(["make", "-s", "-C", ROOT] + need)
    return True


@pytest.fixture(scope="session")
def workdir(tmp_path_factory):
    return str(tmp_path_factory.mktemp("grb"))
