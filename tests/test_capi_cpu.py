"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/vcagan.h
declares; the Python facade keeps the reference's import paths, class names and state_dict keys; the product path
refuses to run without CUDA (no CPU fallback)."""
import ctypes
import json
import os
import pytest
import torch

from conftest import GOLD, ROOT


def test_library_exports_every_declared_symbol():
    import vcagan_b200
    from vcagan_b200._lib import parse_header, LIB_PATH
    protos = parse_header()
    assert len(protos) >= 40
    cdll = ctypes.CDLL(LIB_PATH)
    missing = [n for n in protos if not hasattr(cdll, n)]
    assert not missing, missing
    assert vcagan_b200.lib().cdll.vca_abi_version() == 1


def test_header_and_sources_agree():
    """every extern "C" vca_* definition in csrc is declared in the header and vice versa"""
    import re
    from vcagan_b200._lib import parse_header
    declared = set(parse_header())
    defined = set()
    csrc = os.path.join(ROOT, "visual-context-attentional-gan_b200", "csrc")
    for f in os.listdir(csrc):
        if f.endswith(".cu"):
            defined |= set(re.findall(r"^(?:int|const char\*)\s+(vca_\w+)\s*\(", open(os.path.join(csrc, f)).read(), flags=re.M))
    assert declared == defined - {"vca_set_error"}, (declared ^ defined)


def test_facade_import_paths_and_state_dict_keys():
    from src.models.visual_front import Visual_front
    from src.models.generator import (Decoder, Discriminator, gan_loss, sync_Discriminator, Postnet, ResBlk1D, ResBlk,  # noqa
                                      GenResBlk, Flatten, Avgpool, AVAttention, final_length)
    from src.models.resnet import conv3x3, downsample_basic_block, downsample_basic_block_v2, BasicBlock, ResNet  # noqa
    spec = json.load(open(os.path.join(GOLD, "state_spec.json")))
    mods = dict(v_front=Visual_front(in_channels=1), gen=Decoder(), post=Postnet(), dis1=Discriminator(phase='1'),
                dis2=Discriminator(phase='2'), dis3=Discriminator(num_class=1, max_conv_dim=512, phase='3'),
                s_dis=sync_Discriminator(temp=1.0))
    for k, m in mods.items():
        got = {n: [list(t.shape), str(t.dtype).replace("torch.", "")] for n, t in m.state_dict().items()}
        assert got == spec[k], k
    assert [final_length(t) for t in (40, 50, 75, 160, 250)] == [10, 12, 18, 40, 62]
    n_params = {k: sum(p.numel() for p in m.parameters()) for k, m in mods.items()}
    assert n_params == dict(v_front=19588096, gen=30477283, post=1744896, dis1=3263202, dis2=9850594, dis3=32921314,
                            s_dis=4100224)


def test_no_cpu_fallback():
    from src.models.generator import Postnet
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError):
        Postnet()(torch.zeros(1, 1, 80, 16))


def test_conv_geometry_and_tc_support_predicate():
    import vcagan_b200
    from vcagan_b200.ops import _geom
    g, oshape = _geom((2, 20, 75, 640), (512, 640, 5, 5), (1, 1), (2, 2))
    assert oshape == (2, 20, 75, 512)
    assert vcagan_b200.lib().query("vca_conv_tc_supported", g, 0) == 1
    g, oshape = _geom((2, 5, 112, 112, 1), (64, 1, 5, 7, 7), (1, 2, 2), (2, 3, 3))
    assert oshape == (2, 5, 56, 56, 64)
    assert vcagan_b200.lib().query("vca_conv_tc_supported", g, 0) == 0   # Cin=1, strided: SIMT path
    g, oshape = _geom((2, 1, 300, 80), (128, 80, 7), (1,), (3,))
    assert oshape == (2, 1, 300, 128)
