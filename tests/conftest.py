import os
import subprocess
import sys
import tarfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def refdata(tmp_path_factory):
    """The reference's own test data (inputs + golden outputs), unpacked from
    tests/golden/reference_test_data.tar.xz (see tests/golden/make_fixtures.sh)."""
    d = tmp_path_factory.mktemp("refdata")
    with tarfile.open(os.path.join(ROOT, "tests", "golden", "reference_test_data.tar.xz")) as tf:
        tf.extractall(d)
    return str(d)


@pytest.fixture(scope="session", autouse=True)
def _build_oracle():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
