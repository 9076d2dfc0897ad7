"""The reference's own call site for the hot path, as a callable: the loop body of train.py:166-237 (GRID) and
train_LRS.py:179-243 (LRS), statement by statement, over whatever seven modules it is handed.

It is measurement / test infrastructure, not product code.  Users:
  * bench.py `--impl reference` (the unmodified reference modules of baseline/_ref on the host cores),
  * bench.py `--impl reference-gpu` and the `gpu_eager_baseline` key (the same modules, PyTorch eager on the B200:
    the denominator of the north star's >= 10x target, train.py:53-54 cudnn.benchmark=True),
  * tests/test_gpu_dropin.py (the facade modules of visual-context-attentional-gan_b200/src/models driven by exactly
    this body: CPU-leaf mels, `.cuda()` copies, torch.optim.Adam, retain_graph + second backward, zero_grad pattern).

Nothing is changed against the reference body except: the batch is an argument instead of a DataLoader item, losses are
returned instead of logged, and `train_data.denormalize` is the two-line function of src/data/vid_aud_grid.py:238-240
(the dataset module itself needs librosa / matplotlib / torchaudio at import time).
"""
import math
import os
import sys
import types

import torch
import torch.nn.functional as F
from torch.autograd import grad

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
log1e5 = math.log(1e-5)


def denormalize(melspec):
    """src/data/vid_aud_grid.py:238-240"""
    return ((melspec + 1) * (-log1e5 / 2)) + log1e5


def reference_available():
    return os.path.isfile(os.path.join(REF_DIR, "src", "models", "generator.py"))


def import_reference(cpu_shim=False):
    """Import the unmodified reference modules from baseline/_ref.  cpu_shim=True makes `Tensor.cuda()` the identity,
    which is all the reference needs to run on the host (generator.py:248 hard-codes `.cuda()` on the noise)."""
    if not reference_available():
        raise RuntimeError("baseline/_ref is missing: run `python baseline/install_reference.py` in the build container")
    if cpu_shim:
        torch.Tensor.cuda = lambda self, *a, **k: self
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        del sys.modules[k]           # the facade package exports the same `src.models` import path
    # the reference's `src` has no __init__.py (a namespace package), and a regular package of the same name anywhere on
    # sys.path (the facade) would win over it regardless of order: import with only the reference tree visible
    saved = list(sys.path)
    sys.path[:] = [REF_DIR] + [q for q in saved if not os.path.isfile(os.path.join(q or ".", "src", "__init__.py"))]
    try:
        from src.models.visual_front import Visual_front
        from src.models.generator import Decoder, Discriminator, gan_loss, sync_Discriminator, Postnet
    finally:
        sys.path[:] = saved
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        sys.modules["_vca_ref_" + k] = sys.modules.pop(k)
    return types.SimpleNamespace(Visual_front=Visual_front, Decoder=Decoder, Discriminator=Discriminator, gan_loss=gan_loss,
                                 sync_Discriminator=sync_Discriminator, Postnet=Postnet)


def build_modules(ns, temp=1.0):
    """train.py:70-76"""
    return dict(v_front=ns.Visual_front(in_channels=1), gen=ns.Decoder(), post=ns.Postnet(), dis1=ns.Discriminator(phase='1'),
                dis2=ns.Discriminator(phase='2'), dis3=ns.Discriminator(phase='3'), s_dis=ns.sync_Discriminator(temp=temp))


def build_optimizers(mods, lr=1e-4, weight_decay=1e-5, lrs=False):
    """train.py:78-83 (amsgrad=True) / train_LRS.py:92-98 (plain Adam)"""
    g_params = [{'params': mods["v_front"].parameters()}, {'params': mods["gen"].parameters()}, {'params': mods["post"].parameters()}]
    d_params = [{'params': mods["dis1"].parameters()}, {'params': mods["dis2"].parameters()}, {'params': mods["dis3"].parameters()},
                {'params': mods["s_dis"].parameters()}]
    if lrs:
        return (torch.optim.Adam(g_params, lr=lr, weight_decay=weight_decay), torch.optim.Adam(d_params, lr=lr, weight_decay=weight_decay))
    return (torch.optim.Adam(g_params, lr=lr, weight_decay=weight_decay, amsgrad=True),
            torch.optim.Adam(d_params, lr=lr, weight_decay=weight_decay, amsgrad=True))


def stock_train_step(mods, g_optimizer, d_optimzier, batch, gan_loss, lrs=False, hook=None):
    """One iteration of train.py:166-237 (lrs=False) or train_LRS.py:179-243 (lrs=True).  batch = (mel, spec, vid, vid_len)
    as the DataLoader yields them: CPU tensors.  `hook(name)` (optional) is called at the two points where the reference
    reads gradients out of the modules: after dis_loss.backward() ('d_backward') and after gen_loss.backward()
    ('g_backward'), before the respective optimizer step."""
    v_front, gen, post = mods["v_front"], mods["gen"], mods["post"]
    dis1, dis2, dis3, s_dis = mods["dis1"], mods["dis2"], mods["dis3"], mods["s_dis"]
    criterion = torch.nn.L1Loss()
    mel, spec, vid, vid_len = batch

    v_front.zero_grad(), gen.zero_grad(), post.zero_grad()

    mel1 = F.interpolate(mel, scale_factor=0.25, mode='bilinear')
    mel2 = F.interpolate(mel, scale_factor=0.5, mode='bilinear')

    phon, sent = v_front(vid.cuda())  # B,S,512, B,512
    g1, g2, g3 = gen(sent, phon, vid_len)

    mel.requires_grad = True
    mel1.requires_grad = True
    mel2.requires_grad = True

    ################################### DIS ########################################

    ur_lo1, cr_lo1 = dis1(mel1.cuda(), sent.detach(), phon.size(1))
    ur_lo2, cr_lo2 = dis2(mel2.cuda(), sent.detach(), phon.size(1))
    ur_lo3, cr_lo3 = dis3(mel.cuda(), sent.detach(), phon.size(1))

    sync_loss = s_dis(phon, mel.cuda()).mean()  # B*S, 1

    grad_r1 = grad(outputs=ur_lo1.sum(), inputs=mel1, create_graph=True)[0]
    grad_r2 = grad(outputs=ur_lo2.sum(), inputs=mel2, create_graph=True)[0]
    grad_r3 = grad(outputs=ur_lo3.sum(), inputs=mel, create_graph=True)[0]

    grad_p1 = (grad_r1.view(grad_r1.size(0), -1).norm(2, dim=1) ** 2).mean()
    grad_p2 = (grad_r2.view(grad_r2.size(0), -1).norm(2, dim=1) ** 2).mean()
    grad_p3 = (grad_r3.view(grad_r3.size(0), -1).norm(2, dim=1) ** 2).mean()

    uf_lo1, cf_lo1 = dis1(g1.detach(), sent.detach(), phon.size(1))
    uf_lo2, cf_lo2 = dis2(g2.detach(), sent.detach(), phon.size(1))
    uf_lo3, cf_lo3 = dis3(g3.detach(), sent.detach(), phon.size(1))

    real_loss = 1 / 3 * (gan_loss(ur_lo1, True) + gan_loss(ur_lo2, True) + gan_loss(ur_lo3, True)
                         + gan_loss(cr_lo1, True) + gan_loss(cr_lo2, True) + gan_loss(cr_lo3, True)) \
        + 1 / 3 * (grad_p1 + grad_p2 + grad_p3)

    fake_loss = 1 / 3 * (gan_loss(uf_lo1, False) + gan_loss(uf_lo2, False) + gan_loss(uf_lo3, False)
                         + gan_loss(cf_lo1, False) + gan_loss(cf_lo2, False) + gan_loss(cf_lo3, False))

    dis_loss = real_loss + fake_loss + (0.5 * sync_loss if lrs else sync_loss)

    d_optimzier.zero_grad()
    dis_loss.backward(retain_graph=True)    # accumulate v_front grad
    if hook is not None:
        hook("d_backward")
    d_optimzier.step()

    ################################### GEN ########################################

    gs = post(g3)

    ug_lo1, cg_lo1 = dis1(g1, sent.detach(), phon.size(1))
    ug_lo2, cg_lo2 = dis2(g2, sent.detach(), phon.size(1))
    ug_lo3, cg_lo3 = dis3(g3, sent.detach(), phon.size(1))

    g_sync_loss = s_dis(phon.detach(), g3, True).mean()  # B*S, 1

    g_loss = 1 / 3 * (gan_loss(ug_lo1, True) + gan_loss(ug_lo2, True) + gan_loss(ug_lo3, True)
                      + gan_loss(cg_lo1, True) + gan_loss(cg_lo2, True) + gan_loss(cg_lo3, True))
    if lrs:
        recon_loss = 1 / 3 * (criterion(g1, mel1.cuda()) + criterion(g2, mel2.cuda()) + criterion(g3, mel.cuda())) + criterion(gs, spec.cuda())
        gen_loss = g_loss + recon_loss * 50.0 + g_sync_loss
    else:
        g_loss = g_loss + g_sync_loss
        recon_loss = (criterion(denormalize(g1), denormalize(mel1.cuda())) +
                      criterion(denormalize(g2), denormalize(mel2.cuda())) +
                      criterion(denormalize(g3), denormalize(mel.cuda()))) / 3. \
            + criterion(gs, spec.cuda())
        gen_loss = g_loss + recon_loss * 50.0

    loss_value = gen_loss.cpu().item()      # train.py:233 (a device sync, kept)

    dis1.zero_grad(), dis2.zero_grad(), dis3.zero_grad(), s_dis.zero_grad(), gen.zero_grad(), post.zero_grad()
    gen_loss.backward()
    if hook is not None:
        hook("g_backward")
    g_optimizer.step()

    return dict(gen_loss=loss_value, dis_loss=dis_loss.detach(), sync_loss=sync_loss.detach(), g_sync=g_sync_loss.detach(),
                recon=recon_loss.detach(), real_loss=real_loss.detach(), fake_loss=fake_loss.detach(),
                grad_pen=torch.stack([grad_p1.detach().cpu(), grad_p2.detach().cpu(), grad_p3.detach().cpu()]),
                g1=g1.detach(), g2=g2.detach(), g3=g3.detach(), gs=gs.detach())
