import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG_NAME = "opendlv-logic-cfsd18-sensation-slam_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_pkg():
    pkg = importlib.import_module(PKG_NAME)
    sys.modules.setdefault("slam_b200", pkg)
    return pkg


@pytest.fixture(scope="session")
def pkg():
    return load_pkg()


@pytest.fixture(scope="session")
def synth(pkg):
    return pkg.synth


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (checker).  'best' = the build against the reference's vendored Eigen if present."""
    from oracle import oracle
    return oracle.load("best")


@pytest.fixture(scope="session")
def orc_port():
    from oracle import oracle
    return oracle.load("port")


@pytest.fixture(scope="session")
def ctx(pkg):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    c = pkg.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def c1_drive(synth):
    return synth.trackdrive(1)


@pytest.fixture(scope="session")
def c1_graph(synth, c1_drive):
    return synth.graph_from_drive(c1_drive)


def small_graph(synth, n_poses=120, seed=7):
    """A short open drive on the ellipse track: a few hundred unknowns, seconds on the oracle."""
    trk = synth.ellipse_track()
    d = synth.simulate_drive(trk, n_poses, s_step=trk.length / 1000, seed=seed, closed=True)
    return synth.graph_from_drive(d)


def skip_if_sanitizer_runtime_unusable(output):
    """Sanitizer runtimes refuse to start on some kernels / container settings (ASLR entropy, ptrace limits);
    that is an environment limit, not a finding."""
    for marker in ("unexpected memory mapping", "Shadow memory range interleaves", "ReserveShadowMemoryRange failed",
                   "LeakSanitizer has encountered a fatal error", "failed to intercept"):
        if marker in output:
            pytest.skip("sanitizer runtime cannot start here: " + marker)
