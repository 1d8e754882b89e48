import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built():
    """The in-tree libraries; (re)built on demand so a fresh checkout can run the CPU suite."""
    import __graft_entry__ as g
    from satdump_b200 import capi
    if not os.path.exists(capi.LIB_PATH) or not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        g.build()
    return True


@pytest.fixture(scope="session", autouse=True)
def _floors_for_this_machines_signals(request):
    """GPU sessions: the parity floors (tests/floors.py) belong to one realisation of the synthetic signals; on a machine whose torch
    generator produces another one (CUDA) they are re-measured from the compiled reference before the tests run, in parallel."""
    if not any(item.get_closest_marker("gpu") for item in request.session.items):
        return
    try:
        import torch
        from oracle import ref
        if torch.cuda.is_available() and ref.available():
            from tests import floors
            n = floors.warm()
            if n:
                print(f"\n[floors] re-measured {n} entries for this machine's signals", flush=True)
    except Exception as e:  # the tests then measure what they need one by one
        print(f"\n[floors] warm-up skipped: {e!r}", flush=True)
