import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run by the driver on the GPU box with -m gpu)")


def _gpu_present() -> bool:
    try:
        import bayesnets_jl_b200 as bn
        return bn.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # GPU tests only run where a GPU exists; there is no CPU fallback to hide behind.
    if _gpu_present():
        return
    skip = pytest.mark.skip(reason="no CUDA device visible (GPU tests run under gpurun)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle():
    import oracle as orc
    orc.capi.build()
    return orc


@pytest.fixture(scope="session")
def bn():
    import bayesnets_jl_b200
    return bayesnets_jl_b200


def make_mirror_net(bnmod, nodes):
    """nodes = [[name, parents, rows], ...] (the golden JSON layout) -> bayesnets_jl_b200.BayesNet"""
    net = bnmod.BayesNet()
    for name, parents, rows in nodes:
        pn = [net.ncategories(p) for p in parents]
        net.push(bnmod.CategoricalCPD(name, parents, pn, rows))
    return net


def oracle_net_from_layout(lay):
    """DeviceLayout -> the duck-typed flat net oracle.capi expects (same arrays, independent consumer)."""
    from oracle.factors_np import FlatNet
    return FlatNet(list(lay.names), lay.card, lay.par_ptr, lay.par_idx, lay.cpt_off, lay.cpt, dict(lay.index))


def sigma_bound(p, n_eff):
    return np.sqrt(np.maximum(p * (1 - p), 1e-12) / n_eff)
