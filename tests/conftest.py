import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "mod-interleaveavx_multithreads-faid_b200"))
sys.path.insert(0, str(ROOT / "oracle"))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "ref: needs the compiled reference under oracle/_ref (built in the dev container)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    import pyoracle
    pyoracle.build(ref=False)
    return pyoracle.Oracle()


@pytest.fixture(scope="session")
def engine_lib():
    import subprocess

    import ldpc_b200
    # Always go through build.py: it rebuilds only the objects that are older than a source (or whose flags changed), so a
    # stale prebuilt .so can never be what the tests exercise.  On the GPU box the snapshot's objects are up to date and
    # this is a no-op of a few milliseconds.
    if not os.environ.get("LDPC_B200_LIB"):
        subprocess.run([sys.executable, str(ldpc_b200.PKG_DIR / "build.py")], check=True)
    return ldpc_b200.load_library()
