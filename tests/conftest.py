import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Make sure the oracle, the host math harness and libtmkcl.so exist (cheap if up to date)."""
    import __graft_entry__ as g
    from tmkit_b200 import amino as A
    harness = os.path.join(ROOT, "tests", "harness", "libdevmath_host.so")
    oracle = os.path.join(ROOT, "oracle", "libtmkoracle.so")
    if not (os.path.exists(A.LIB_PATH) and os.path.exists(harness) and os.path.exists(oracle)):
        g.build()
    return True


@pytest.fixture(scope="session")
def sussman(built):
    """Baxter + Sussman scene: neutral description, oracle scene with the q0 allowance applied."""
    from oracle import oracle as O
    from tmkit_b200 import scenes as S
    sc = S.baxter_sussman()
    flat = S.flatten(sc)
    orc = O.OracleScene(flat)
    q0 = S.q0_vector(sc)
    orc.allow_config(q0)
    return sc, flat, orc, q0


def have_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
