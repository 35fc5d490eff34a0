import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "path-tracing__ray-tracer_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (container only)")


@pytest.fixture(scope="session")
def cornell():
    """Cornell scene (seed 0, synthetic textures) built with the b200rt mirror of the reference API."""
    import random
    from b200rt.cornell import CustomSceneBuilder
    random.seed(0)
    b = CustomSceneBuilder(texture_dir=False)
    scene = b.build_scene()
    return scene, b


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
