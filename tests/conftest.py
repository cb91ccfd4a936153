import os
import sys

import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle

    return Oracle()


@pytest.fixture(scope="session")
def scene(oracle):
    from common import Scene

    return Scene(oracle, frames=(0, 7))


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stage1_ref.npz"))


def _make_ctx(scene, **kw):
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from insta_b200 import capi

    capi.load_library(build_if_missing=False)
    ctx = capi.Context(device=0, **kw)
    ctx.set_canonical(scene.verts, scene.faces, scene.masks)
    for v in scene.frame_verts:
        ctx.add_frame(v)
    return ctx


@pytest.fixture(scope="session")
def gpu_ctx(scene):
    """A context on cuda:0 holding the canonical mesh and the scene's frames, default arithmetic (= the reference's CUDA
    build).  Fails loudly when the CUDA library or the device is missing -- GPU tests must never pass on a fallback."""
    ctx = _make_ctx(scene)
    yield ctx
    ctx.close()


@pytest.fixture(scope="session")
def gpu_ctx_host(scene):
    """Same, with arith = 1: the rounding of the reference's HOST build (what the committed golden vectors were made with)."""
    ctx = _make_ctx(scene, arith=1)
    yield ctx
    ctx.close()


@pytest.fixture(params=["device", "host"])
def arith_pair(request, oracle, gpu_ctx, gpu_ctx_host):
    """(context, oracle) set to the same arithmetic flavour."""
    oracle.set_arith(request.param)
    yield (gpu_ctx if request.param == "device" else gpu_ctx_host), oracle
    oracle.set_arith("host")
