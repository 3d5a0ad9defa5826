import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SCENES = os.path.join(ROOT, "tests", "golden", "scenes")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_scene(name):
    from openrave_b200.environment import Environment
    env = Environment()
    env.Load(os.path.join(SCENES, name + ".npz"))
    return env


@pytest.fixture()
def lab1():
    return load_scene("lab1")


def sample_configs(robot, n, seed, active=None):
    """RobotConfigurationSampler semantics (plugins/basesamplers/robotconfiguration.cpp): lower + u*(upper-lower)."""
    lo, hi = robot.GetDOFLimits()
    idx = list(range(robot.GetDOF())) if active is None else list(active)
    lo = np.array([lo[i] for i in idx])
    hi = np.array([hi[i] for i in idx])
    rng = np.random.default_rng(seed)
    return lo + rng.random((n, len(idx))) * (hi - lo)


def have_ref_oracle():
    import oracle
    return oracle.available("reference")
