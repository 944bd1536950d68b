import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "dg-solver-for-pbte-with-mfem_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs oracle/_ref (the reference's own sources compiled against the stand-in)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The libraries are built in-tree by __graft_entry__.build(); build them on first use if they are missing."""
    lib = os.path.join(ROOT, "dg-solver-for-pbte-with-mfem_b200", "lib", "libpbte_cuda.so")
    ora = os.path.join(ROOT, "oracle", "_build", "libpbte_oracle.so")
    if not (os.path.exists(lib) and os.path.exists(ora)):
        import __graft_entry__
        __graft_entry__.build()


def golden(name):
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
