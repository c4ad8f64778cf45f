import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "depth-aware-shader-effects-for-nerf_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (PKG, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100 (B200) device; run with -m gpu on the GPU box")
    # The tests load the C-ABI library.  In a fresh checkout (the .so is git-ignored) build it first, exactly like
    # __graft_entry__.build(); where nvcc is missing too the ABI tests fail loudly with the ImportError of nerfw._lib.
    lib_path = os.path.join(PKG, "nerfw", "libnerfw_sm100.so")
    if not os.path.exists(lib_path):
        import shutil
        if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
            sys.path.insert(0, ROOT)
            import __graft_entry__
            __graft_entry__.build()


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"))
    return load


@pytest.fixture(scope="session")
def manifest():
    with open(os.path.join(GOLDEN, "manifest.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle():
    import nerfw_oracle
    return nerfw_oracle


@pytest.fixture(scope="session")
def state_dict(oracle):
    """seed-0 reference-init weights + the golden embedding (randn(32) drawn right after the model)."""
    sd = oracle.make_state_dict(0)
    emb = torch.randn(32)
    return sd, emb


@pytest.fixture(scope="session")
def cuda_model(state_dict):
    import nerfw
    from config import Config
    sd, emb = state_dict
    m = nerfw.NeRF(Config())
    m.load_state_dict(sd, strict=True)
    return m.cuda(), emb.cuda()
