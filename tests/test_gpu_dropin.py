"""The drop-in claim of SURVEY 8(b), executed: the reference's own call site -- the loop body of train.py:166-237,
statement by statement (baseline/stock_step.py) -- runs UNCHANGED on the facade modules imported through the
reference's import paths (`from src.models.visual_front import Visual_front`, `from src.models.generator import ...`):
CPU-leaf mels with `.cuda()` copies, `torch.autograd.grad(..., create_graph=True)` w.r.t. those CPU leaves,
`torch.optim.Adam(amsgrad=True)` on `.parameters()`, `backward(retain_graph=True)` then a second backward over the shared
v_front / gen graph, the `zero_grad` pattern of train.py:235, optionally every module wrapped in nn.DataParallel
(train.py:112-119).  Results are held to the golden step of the unmodified reference (tests/golden/make_golden.py).
Also: the checkpoint layout of train.py:303-309 round-trips through torch.save / load_state_dict, and the test.py:153-155
`.npz` files are written."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import make_state, golden_inputs, rel_l2, GOLD
from oracle import vca_oracle as O

pytestmark = pytest.mark.gpu


def _facade_modules(spec):
    import vcagan_b200 as V  # noqa: F401  (loads the library)
    from src.models.visual_front import Visual_front                                   # train.py:7
    from src.models.generator import Decoder, Discriminator, gan_loss, sync_Discriminator, Postnet   # train.py:8
    mods = dict(v_front=Visual_front(in_channels=1), gen=Decoder(), post=Postnet(), dis1=Discriminator(phase='1'),
                dis2=Discriminator(phase='2'), dis3=Discriminator(phase='3'), s_dis=sync_Discriminator(temp=1.0))
    for k, m in mods.items():
        m.load_state_dict(make_state(spec, k))         # strict: the key set is the reference's
        m.cuda()                                        # train.py:104-110
        m.train()
    mods["v_front"].dropout.p = 0.0                     # the golden step was generated with dropout off ...
    mods["v_front"].sentence_encoder.dropout = 0.0
    return mods, gan_loss


def _stock_step_on_facade(golden, dataparallel):
    import vcagan_b200 as V
    from baseline import stock_step as S
    spec = json.load(open(os.path.join(GOLD, "state_spec.json")))
    V.set_precision("fp32")
    mods, gan_loss = _facade_modules(spec)
    vid, mel, sp, noise = golden_inputs()
    mods["gen"].fixed_noise = noise                     # ... and the noise of generator.py:248 injected
    named = {f"{k}.{n}": p for k, m in mods.items() for n, p in m.named_parameters()}
    bufs = {f"{k}.{n}": b for k, m in mods.items() for n, b in m.named_buffers()}
    g_opt, d_opt = S.build_optimizers(mods)             # torch.optim.Adam(amsgrad=True) on .parameters(), train.py:78-83
    if dataparallel:
        mods = {k: torch.nn.DataParallel(m) for k, m in mods.items()}    # train.py:112-119
    norms = {}

    def hook(name):
        for n, p in named.items():
            if p.grad is not None:
                norms[(name, n)] = float(p.grad.norm())
    out = S.stock_train_step(mods, g_opt, d_opt, (mel, sp, vid, torch.tensor([20, 13])), gan_loss, hook=hook)
    torch.cuda.synchronize()
    for k in ("dis_loss", "sync_loss", "real_loss", "fake_loss", "gen_loss", "g_sync", "recon"):
        ref, got = float(golden["step_" + k]), float(out[k])
        assert abs(got - ref) <= 1e-4 * max(1.0, abs(ref)), (k, got, ref)
    assert rel_l2(out["grad_pen"], golden["step_grad_pen"]) < 2e-4
    for k in ("g1", "g2", "g3", "gs"):
        assert rel_l2(out[k].cpu(), golden["step_" + k]) < 1e-4, k
    names = json.load(open(os.path.join(GOLD, "grad_norm_names.json")))
    # gradient norms after each of the two backward passes; truth = the reference's fp64 run, the bar = what the
    # reference's own fp32 arithmetic achieves against it (train-mode BN at B = 2 is ill-conditioned, see test_gpu_step.py)
    for key, tag, when in (("d", "d_grad_norms", "d_backward"), ("vf_d", "vf_d_grad_norms", "d_backward"), ("g", "g_grad_norms", "g_backward")):
        mine = torch.tensor([norms[(when, n)] for n in names[key]])
        t64, r32 = golden["step64_" + tag], golden["step_" + tag]
        e_mine, e_ref = rel_l2(mine, t64), rel_l2(r32, t64)
        print(f"{tag}: facade under the stock loop vs fp64 {e_mine:.2e}; reference fp32 vs fp64 {e_ref:.2e}")
        assert e_mine <= max(1e-4, 3 * e_ref), (tag, e_mine, e_ref)
    # the D-phase backward must have left its sync gradient in the v_front CNN and none in GRU / fc
    assert norms[("d_backward", "v_front.frontend.0.weight")] > 0
    assert ("d_backward", "v_front.fc.weight") not in norms or norms[("d_backward", "v_front.fc.weight")] == 0
    cn = json.load(open(os.path.join(GOLD, "checksum_names.json")))
    chk = torch.tensor([float(named[n].detach().double().abs().sum()) for n in cn["params"]], dtype=torch.float64)
    ref = torch.from_numpy(golden["step_param_checksums"])[:, 1]
    numel = torch.tensor([float(named[n].numel()) for n in cn["params"]], dtype=torch.float64)
    err = (chk - ref).abs()
    assert bool((err <= 1e-4 * numel * 1.01 + 1e-9).all())          # Adam moves nothing by more than lr (first step)
    rel = err / ref.abs().clamp_min(1e-9)
    assert float((rel < 1e-5).double().mean()) > 0.8, float((rel < 1e-5).double().mean())
    bs = torch.tensor([float(bufs[n].double().sum()) for n in cn["buffers"]], dtype=torch.float64)
    assert rel_l2(bs, golden["step_buffer_sums"]) < 1e-5


def test_stock_train_loop_on_facade_fp32(golden):
    _stock_step_on_facade(golden, dataparallel=False)


def test_stock_train_loop_on_facade_dataparallel(golden):
    _stock_step_on_facade(golden, dataparallel=True)


def test_stock_train_loop_on_facade_bf16_second_iteration():
    """Two iterations of the stock loop in bf16 (the tcgen05 path) with dropout and device noise on: finite, weights
    move, and the v_front gradient left by iteration 1 is cleared by the `v_front.zero_grad()` at the top of iteration 2."""
    import vcagan_b200 as V
    from baseline import stock_step as S
    spec = json.load(open(os.path.join(GOLD, "state_spec.json")))
    try:
        V.set_precision("bf16")
        mods, gan_loss = _facade_modules(spec)
        mods["v_front"].dropout.p = 0.3
        mods["v_front"].sentence_encoder.dropout = 0.3
        g_opt, d_opt = S.build_optimizers(mods)
        vid, mel, sp, _ = golden_inputs()
        w0 = mods["gen"].decode[0].conv1.weight.detach().clone()
        for _ in range(2):
            out = S.stock_train_step(mods, g_opt, d_opt, (mel.clone(), sp, vid, torch.tensor([20, 13])), gan_loss)
        torch.cuda.synchronize()
        assert np.isfinite(out["gen_loss"]) and bool(torch.isfinite(out["dis_loss"]))
        assert float((mods["gen"].decode[0].conv1.weight.detach() - w0).abs().max()) > 0
    finally:
        V.set_precision("fp32")


def test_checkpoint_roundtrip_and_eval_npz(tmp_path, state_spec):
    """train.py:303-309 saves {'<name>_state_dict': module.state_dict()}; train.py:91-102 / test.py:70-75 load it back
    with load_state_dict.  Save from a Trainer that has taken a step (so the parameters are views into its flat buffers
    and the BN buffers have moved), load into FRESH facade modules, require the same eval outputs (fp32 mode; the bf16
    split-K kernels accumulate with atomics and are not bit-reproducible run to run); then write the test.py:153-155
    .npz files."""
    import vcagan_b200 as V
    from vcagan_b200 import infer
    from vcagan_b200.trainer import Trainer
    try:
        state = {m: make_state(state_spec, m) for m in O.MODULES}
        tr = Trainer(precision="fp32", state=state, dropout=True)     # fp32: the forward kernels are deterministic
        vid, mel, sp, noise = golden_inputs()
        tr.step(vid.cuda(), mel.cuda(), sp.cuda(), [20, 13])
        ck = tr.state_dicts()
        assert sorted(ck) == sorted(f"{k}_state_dict" for k in O.MODULES)
        for k in O.MODULES:        # exact key set and shapes of the reference's checkpoints
            sd = ck[f"{k}_state_dict"]
            assert sorted(sd) == sorted(state_spec[k]), k
            assert all(list(sd[n].shape) == state_spec[k][n][0] for n in sd), k
        path = str(tmp_path / "ckpt.ckpt")
        torch.save(ck, path)
        loaded = torch.load(path, map_location=lambda storage, loc: storage.cuda())      # train.py:93
        from src.models.visual_front import Visual_front
        from src.models.generator import Decoder, Postnet
        v2, g2, p2 = Visual_front(in_channels=1), Decoder(), Postnet()
        v2.load_state_dict(loaded['v_front_state_dict']); g2.load_state_dict(loaded['gen_state_dict'])
        p2.load_state_dict(loaded['post_state_dict'])
        for m in (v2, g2, p2):
            m.cuda().eval()
        v1, g1, p1 = tr.mods["v_front"].eval(), tr.mods["gen"].eval(), tr.mods["post"].eval()
        g1.fixed_noise = g2.fixed_noise = noise
        with torch.no_grad():
            a = p1(g1(*reversed(v1(vid.cuda())), [20, 13])[2])
            b = p2(g2(*reversed(v2(vid.cuda())), [20, 13])[2])
        assert rel_l2(a.cpu(), b.cpu()) < 1e-6, rel_l2(a.cpu(), b.cpu())
        # the .npz / .wav files test.py:145-159 writes
        g2.fixed_noise = None            # flip TTA runs the clip and its mirror image as one batch of 2B: device noise
        o = infer.synthesize(v2, g2, p2, vid.cuda(), torch.tensor([20, 13]), n_iters=4, tta=True)
        names = ["s1/video/bbaf2n", "s2/video/lgwm5a"]
        paths = infer.save_eval_outputs(str(tmp_path / "test"), names, o["mel"], o["spec"], [80, 52], wav=o["wav"])
        z = np.load(paths[1])
        assert z["mel"].shape == (1, 80, 52) and z["spec"].shape == (1, 321, 52)
        assert np.array_equal(z["mel"], o["mel"][1, :, :, :52].float().cpu().numpy())
        import wave
        with wave.open(str(tmp_path / "test" / "wav" / "s1" / "bbaf2n.wav")) as f:
            assert f.getframerate() == 16000 and f.getsampwidth() == 2 and f.getnframes() == o["wav"].shape[1]
    finally:
        V.set_precision("fp32")
