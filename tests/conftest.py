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
def golden_preproc():
    """Clip preprocessing vectors from the reference's build_tensor (tests/golden/make_golden_preproc.py)."""
    return dict(np.load(os.path.join(GOLD, "golden_preproc.npz")))


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


def synthetic_frames(seed, n, H, W):
    """Seeded uint8 RGB frames (n,H,W,3): a smooth moving pattern plus noise (so both interpolation and rounding
    matter).  Shared by tests/golden/make_golden_preproc.py and the preprocessing tests; numpy's PCG64 stream is
    stable across versions, so the frames themselves need not be committed."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    out = np.empty((n, H, W, 3), np.uint8)
    for i in range(n):
        base = np.stack([127 + 120 * np.sin(xx / (9.0 + c) + 0.3 * i) * np.cos(yy / (7.0 + 2 * c) - 0.2 * i) for c in range(3)], -1)
        out[i] = np.clip(base + rng.integers(-40, 41, (H, W, 3)), 0, 255).astype(np.uint8)
    return out


def sample_index(numel: int, ns: int = 512) -> torch.Tensor:
    """Up to `ns` fixed positions spread over a flattened parameter (all of it when it is small): the sampling of the
    gradient fixtures (tests/golden/make_golden_bf16.py) and of the tests that read them."""
    if numel <= ns:
        return torch.arange(numel)
    return torch.linspace(0, numel - 1, ns, dtype=torch.float64).round().long()
