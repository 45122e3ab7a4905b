import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.build()
    return o


@pytest.fixture(scope="session")
def ctx():
    import gl_slam_b200 as g
    c = g.Context(device=0)
    yield c
    c.close()


def pytest_terminal_summary(terminalreporter):
    """A skipped fuzz draw is an untested draw: count and print them so the number is in the log."""
    skipped = terminalreporter.stats.get("skipped", [])
    fuzz = [r for r in skipped if "test_gpu_fuzz" in r.nodeid]
    ran = sum(1 for k in ("passed", "failed") for r in terminalreporter.stats.get(k, []) if "test_gpu_fuzz" in r.nodeid)
    if fuzz or ran:
        terminalreporter.write_line(f"fuzz draws: {ran} gated, {len(fuzz)} skipped as degenerate "
                                    f"({', '.join(r.nodeid.split('[')[-1].rstrip(']') for r in fuzz) or 'none'})")
