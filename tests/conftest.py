"""Shared fixtures.  GPU tests are marked ``@pytest.mark.gpu`` and skipped without CUDA.

``ref`` exposes the upstream reference (read-only at /root/reference, or the offline install at
baseline/_ref) as a *test oracle only*: its pure-PyTorch ops run on CPU once matplotlib/kornia/
lpips are stubbed and ``torch.cuda.synchronize`` is a no-op.  Tests that need
it skip when it is absent; nothing from it is ever copied into the package.
"""
import importlib
import os
import sys

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "multigpu: needs >= 2 CUDA devices")


def pytest_collection_modifyitems(config, items):
    has_cuda = torch.cuda.is_available()
    ngpu = torch.cuda.device_count() if has_cuda else 0
    for item in items:
        if "gpu" in item.keywords and not has_cuda:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))
        if "multigpu" in item.keywords and ngpu < 2:
            item.add_marker(pytest.mark.skip(reason="needs >= 2 GPUs"))


def _reference_root():
    for cand in ("/root/reference", os.path.join(REPO, "baseline", "_ref")):
        if os.path.exists(os.path.join(cand, "operations", "mpi_rendering.py")):
            return cand
    return None


class _RefModules:
    """Imports reference modules under a private name prefix so they never shadow ours."""

    def __init__(self, root):
        self.root = root
        self._cache = {}

    def load(self, dotted):
        if dotted in self._cache:
            return self._cache[dotted]
        from mine_b200.bench.ref_shims import install_import_shims, isolated_reference_imports
        install_import_shims()
        with isolated_reference_imports(self.root):
            mod = importlib.import_module(dotted)
        self._cache[dotted] = mod
        return mod


@pytest.fixture(scope="session")
def ref():
    root = _reference_root()
    if root is None:
        pytest.skip("upstream reference not available")
    return _RefModules(root)


@pytest.fixture
def rng():
    g = torch.Generator().manual_seed(1234)
    return g
