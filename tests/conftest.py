import json, os, sys
import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "visual-context-attentional-gan_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    return dict(np.load(os.path.join(GOLD, "golden_small.npz")))


@pytest.fixture(scope="session")
def golden_tail():
    """Waveform tail / mel front vectors from the reference's dataset methods (tests/golden/make_golden_tail.py)."""
    return dict(np.load(os.path.join(GOLD, "golden_tail.npz")))


@pytest.fixture(scope="session")
def state_spec():
    return json.load(open(os.path.join(GOLD, "state_spec.json")))


def make_state(spec, module, requires_grad=False):
    """Deterministic name-keyed weights for `module` ('v_front', 'gen', ...) from the committed spec."""
    from oracle import vca_oracle as O
    sd = {}
    for k, (shape, dt) in spec[module].items():
        t = O.det_tensor(module + "." + k, shape, getattr(torch, dt))
        if requires_grad and t.is_floating_point() and not k.endswith(("running_mean", "running_var")):
            t.requires_grad_(True)
        sd[k] = t
    return sd


def golden_inputs(B=2, T=20):
    g = torch.Generator().manual_seed(1234)
    vid = torch.randn(B, 1, T, 112, 112, generator=g)
    mel = torch.rand(B, 1, 80, 4 * T, generator=g) * 2 - 1
    spec = torch.rand(B, 1, 321, 4 * T, generator=g)
    noise = torch.randn(B, 128, 20, T, generator=g)
    return vid, mel, spec, noise


def rel_l2(a, b):
    a = torch.as_tensor(a).double().flatten(); b = torch.as_tensor(b).double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))
