"""Generate tests/golden/*.npz + state_spec.json from the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference, which does not travel to the GPU box):
    python tests/golden/make_golden.py
The reference hard-codes .cuda() (src/models/generator.py:248) and imports librosa
(src/data/stft.py:32); both are shimmed exactly as SURVEY.md section 8(c) describes.  Weights are the
name-keyed deterministic tensors of oracle.vca_oracle.det_tensor so that no 400 MB checkpoint has to
be committed; inputs come from seeded CPU generators.
"""
import json, os, sys, types
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("VCA_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

torch.Tensor.cuda = lambda self, *a, **k: self  # CPU shim for generator.py:248
lib = types.ModuleType("librosa"); util = types.ModuleType("librosa.util"); filt = types.ModuleType("librosa.filters")
util.pad_center = lambda data, size, **k: data
util.tiny = lambda x: np.finfo(np.float32).tiny
util.normalize = lambda x, norm=None, **k: x
lib.util = util; lib.filters = filt
sys.modules.update({"librosa": lib, "librosa.util": util, "librosa.filters": filt})

from src.models.visual_front import Visual_front  # noqa: E402
from src.models.generator import Decoder, Discriminator, sync_Discriminator, Postnet, gan_loss, final_length  # noqa: E402
from src.data.stft import STFT  # noqa: E402
from src.data.audio_processing import griffin_lim  # noqa: E402
import src.data.audio_processing as ap  # noqa: E402
from oracle import vca_oracle as O  # noqa: E402

torch.set_num_threads(8)
B, T = 2, 20
LENS = [20, 13]


def gen_inputs():
    g = torch.Generator().manual_seed(1234)
    vid = torch.randn(B, 1, T, 112, 112, generator=g)
    mel = torch.rand(B, 1, 80, 4 * T, generator=g) * 2 - 1
    spec = torch.rand(B, 1, 321, 4 * T, generator=g)
    noise = torch.randn(B, 128, 20, T, generator=g)
    return vid, mel, spec, noise


def load_det(mod, prefix):
    sd = O.fill_deterministic(mod.state_dict(), prefix)
    mod.load_state_dict(sd)
    return mod


def build():
    mods = dict(v_front=Visual_front(), gen=Decoder(), post=Postnet(), dis1=Discriminator(phase='1'),
                dis2=Discriminator(phase='2'), dis3=Discriminator(phase='3'), s_dis=sync_Discriminator())
    for k, m in mods.items():
        load_det(m, k)
    mods["v_front"].dropout.p = 0.0
    mods["v_front"].sentence_encoder.dropout = 0.0
    return mods


class FixedNoise:
    """Replace torch.randn inside Decoder.forward by the injected noise tensor."""
    def __init__(self, noise): self.noise = noise
    def __enter__(self):
        self.orig = torch.randn
        torch.randn = lambda *a, **k: self.noise.clone()
    def __exit__(self, *a): torch.randn = self.orig


def ref_step(out, vid, mel, spec, noise, dtype, pre):
    """train.py:166-237 verbatim on the reference modules (dropout off, noise injected), in `dtype`."""
    mods = build()
    for m in mods.values():
        m.train()
        m.to(dtype)
    vid, mel, spec, noise = (t.to(dtype) for t in (vid, mel, spec, noise))
    v_front, gen, post = mods["v_front"], mods["gen"], mods["post"]
    dis1, dis2, dis3, s_dis = mods["dis1"], mods["dis2"], mods["dis3"], mods["s_dis"]
    g_params = [{'params': v_front.parameters()}, {'params': gen.parameters()}, {'params': post.parameters()}]
    d_params = [{'params': dis1.parameters()}, {'params': dis2.parameters()}, {'params': dis3.parameters()},
                {'params': s_dis.parameters()}]
    g_opt = torch.optim.Adam(g_params, lr=1e-4, weight_decay=1e-5, amsgrad=True)
    d_opt = torch.optim.Adam(d_params, lr=1e-4, weight_decay=1e-5, amsgrad=True)
    F = torch.nn.functional
    grad = torch.autograd.grad
    criterion = torch.nn.L1Loss()
    denorm = O.denormalize
    melr = mel.clone()
    v_front.zero_grad(), gen.zero_grad(), post.zero_grad()
    mel1 = F.interpolate(melr, scale_factor=0.25, mode='bilinear')
    mel2 = F.interpolate(melr, scale_factor=0.5, mode='bilinear')
    phon, sent = v_front(vid)
    with FixedNoise(noise):
        g1, g2, g3 = gen(sent, phon, torch.tensor(LENS))
    melr.requires_grad = True; mel1.requires_grad = True; mel2.requires_grad = True
    ur1, cr1 = dis1(mel1, sent.detach(), phon.size(1))
    ur2, cr2 = dis2(mel2, sent.detach(), phon.size(1))
    ur3, cr3 = dis3(melr, sent.detach(), phon.size(1))
    sync_loss = s_dis(phon, melr).mean()
    gr1 = grad(outputs=ur1.sum(), inputs=mel1, create_graph=True)[0]
    gr2 = grad(outputs=ur2.sum(), inputs=mel2, create_graph=True)[0]
    gr3 = grad(outputs=ur3.sum(), inputs=melr, create_graph=True)[0]
    gp1 = (gr1.view(gr1.size(0), -1).norm(2, dim=1) ** 2).mean()
    gp2 = (gr2.view(gr2.size(0), -1).norm(2, dim=1) ** 2).mean()
    gp3 = (gr3.view(gr3.size(0), -1).norm(2, dim=1) ** 2).mean()
    uf1, cf1 = dis1(g1.detach(), sent.detach(), phon.size(1))
    uf2, cf2 = dis2(g2.detach(), sent.detach(), phon.size(1))
    uf3, cf3 = dis3(g3.detach(), sent.detach(), phon.size(1))
    real_loss = 1 / 3 * (gan_loss(ur1, True) + gan_loss(ur2, True) + gan_loss(ur3, True)
                         + gan_loss(cr1, True) + gan_loss(cr2, True) + gan_loss(cr3, True)) + 1 / 3 * (gp1 + gp2 + gp3)
    fake_loss = 1 / 3 * (gan_loss(uf1, False) + gan_loss(uf2, False) + gan_loss(uf3, False)
                         + gan_loss(cf1, False) + gan_loss(cf2, False) + gan_loss(cf3, False))
    dis_loss = real_loss + fake_loss + sync_loss
    d_opt.zero_grad()
    dis_loss.backward(retain_graph=True)
    def gnorms(ms):
        return {f"{k}.{n}": float(p.grad.norm()) for k in ms for n, p in mods[k].named_parameters() if p.grad is not None}
    d_gn = gnorms(("dis1", "dis2", "dis3", "s_dis"))
    vf_d_gn = gnorms(("v_front",))
    d_opt.step()
    gs = post(g3)
    ug1, cg1 = dis1(g1, sent.detach(), phon.size(1))
    ug2, cg2 = dis2(g2, sent.detach(), phon.size(1))
    ug3, cg3 = dis3(g3, sent.detach(), phon.size(1))
    g_sync = s_dis(phon.detach(), g3, True).mean()
    g_loss = 1 / 3 * (gan_loss(ug1, True) + gan_loss(ug2, True) + gan_loss(ug3, True)
                      + gan_loss(cg1, True) + gan_loss(cg2, True) + gan_loss(cg3, True)) + g_sync
    recon = (criterion(denorm(g1), denorm(mel1)) + criterion(denorm(g2), denorm(mel2))
             + criterion(denorm(g3), denorm(melr))) / 3. + criterion(gs, spec)
    gen_loss = g_loss + recon * 50.0
    dis1.zero_grad(), dis2.zero_grad(), dis3.zero_grad(), s_dis.zero_grad(), gen.zero_grad(), post.zero_grad()
    gen_loss.backward()
    g_gn = gnorms(("v_front", "gen", "post"))
    g_opt.step()
    res = dict(dis_loss=dis_loss, sync_loss=sync_loss, real_loss=real_loss, fake_loss=fake_loss,
               grad_pen=torch.stack([gp1, gp2, gp3]), gen_loss=gen_loss, g_sync=g_sync, recon=recon,
               g1=g1, g2=g2, g3=g3, gs=gs, phon=phon, sent=sent, r1_grad3=gr3, r1_grad1=gr1)
    out.update({pre + k: v for k, v in res.items()})
    names = sorted(d_gn); out[pre + "d_grad_norms"] = torch.tensor([d_gn[n] for n in names])
    names_v = sorted(vf_d_gn); out[pre + "vf_d_grad_norms"] = torch.tensor([vf_d_gn[n] for n in names_v])
    names_g = sorted(g_gn); out[pre + "g_grad_norms"] = torch.tensor([g_gn[n] for n in names_g])
    json.dump(dict(d=names, vf_d=names_v, g=names_g), open(os.path.join(HERE, "grad_norm_names.json"), "w"))
    # post-step parameter checksums (sum and abs-sum of every parameter after both Adam steps)
    chk = {f"{k}.{n}": [float(p.double().sum()), float(p.double().abs().sum())]
           for k, m in mods.items() for n, p in m.named_parameters()}
    cn = sorted(chk); out[pre + "param_checksums"] = torch.tensor([chk[n] for n in cn], dtype=torch.float64)
    bn = {f"{k}.{n}": float(b.double().sum()) for k, m in mods.items() for n, b in m.named_buffers()}
    bnn = sorted(bn); out[pre + "buffer_sums"] = torch.tensor([bn[n] for n in bnn], dtype=torch.float64)
    json.dump(dict(params=cn, buffers=bnn), open(os.path.join(HERE, "checksum_names.json"), "w"))



def main():
    out = {}
    vid, mel, spec, noise = gen_inputs()
    mods = build()
    spec_json = {k: {n: [list(t.shape), str(t.dtype).replace("torch.", "")] for n, t in m.state_dict().items()}
                 for k, m in mods.items()}
    json.dump(spec_json, open(os.path.join(HERE, "state_spec.json"), "w"), indent=0, sort_keys=True)

    # ---- module forwards, eval mode ----
    for m in mods.values():
        m.eval()
    with torch.no_grad():
        phon, sent = mods["v_front"](vid)
        with FixedNoise(noise):
            g1, g2, g3 = mods["gen"](sent, phon, LENS)
        gs = mods["post"](g3)
        out.update(eval_phon=phon, eval_sent=sent, eval_g1=g1, eval_g2=g2, eval_g3=g3, eval_gs=gs)
        mel1 = torch.nn.functional.interpolate(mel, scale_factor=0.25, mode="bilinear")
        mel2 = torch.nn.functional.interpolate(mel, scale_factor=0.5, mode="bilinear")
        for i, (d, x) in enumerate(((mods["dis1"], mel1), (mods["dis2"], mel2), (mods["dis3"], mel)), 1):
            u, c = d(x, sent, T)
            out[f"eval_d{i}_u"], out[f"eval_d{i}_c"] = u, c
        out["eval_sync_nce"] = mods["s_dis"](phon, mel)
        out["eval_sync_cos"] = mods["s_dis"](phon, g3, True)
    out["known_gan_loss0"] = gan_loss(torch.zeros(4, 1), True)
    out["known_final_length"] = torch.tensor([final_length(t) for t in (40, 50, 75, 160, 250)])

    # ---- one full training step exactly as train.py:166-237 (fresh modules, train mode), fp32 and fp64 ----
    ref_step(out, vid, mel, spec, noise, torch.float32, "step_")
    ref_step(out, vid, mel, spec, noise, torch.float64, "step64_")

    # ---- what bf16 autocast does to the unmodified reference on the same inputs (the yardstick of the bf16 bound) ----
    mods = build()
    for m in mods.values():
        m.train()
    with torch.no_grad():
        p32, s32 = mods["v_front"](vid)
        with FixedNoise(noise):
            g32 = mods["gen"](s32, p32, torch.tensor(LENS))
        gs32 = mods["post"](g32[2])
    mods = build()
    for m in mods.values():
        m.train()
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        pa, sa = mods["v_front"](vid)
        with FixedNoise(noise):
            ga = mods["gen"](sa.float(), pa.float(), torch.tensor(LENS))
        gsa = mods["post"](ga[2].float())
    rel = lambda a, b: float((a.float() - b).norm() / b.norm())
    out["autocast_bf16_train_errs"] = torch.tensor([rel(pa, p32), rel(sa, s32), rel(ga[0], g32[0]), rel(ga[1], g32[1]),
                                                    rel(ga[2], g32[2]), rel(gsa, gs32)])
    out["autocast_bf16_train_mel_l1"] = (ga[2].float() - g32[2]).abs().mean()

    # ---- STFT / Griffin-Lim (src/data/stft.py, audio_processing.py) ----
    stft = STFT(640, 160, 640)
    g = torch.Generator().manual_seed(77)
    sig = torch.randn(2, 160 * 11, generator=g) * 0.1
    mag, ph = stft.transform(sig)
    rec = stft.inverse(mag, ph)
    out.update(stft_mag=mag, stft_phase=ph, stft_rec=rec)
    gl_mag = torch.rand(2, 321, 12, generator=g)
    init = (2 * np.pi * torch.rand(2, 321, 12, generator=g) - np.pi).float()
    orig_rand = np.random.rand
    phase01 = ((init.numpy().astype(np.float64)) / (2 * np.pi)) % 1.0
    np.random.rand = lambda *a: phase01  # audio_processing.py:59 draws the phase from numpy
    wav = griffin_lim(gl_mag, stft, 8)
    np.random.rand = orig_rand
    # what the reference actually used as angles (np.angle(exp(2j*pi*u)))
    used = np.angle(np.exp(2j * np.pi * phase01)).astype(np.float32)
    out.update(gl_mag=gl_mag, gl_init_phase=torch.from_numpy(used), gl_wav=wav)
    out["window_sumsquare_12"] = torch.from_numpy(ap.window_sumsquare('hann', 12, hop_length=160, win_length=640, n_fft=640))

    np.savez_compressed(os.path.join(HERE, "golden_small.npz"),
                        **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in out.items()})
    print("wrote", len(out), "arrays;", os.path.getsize(os.path.join(HERE, "golden_small.npz")) / 1e6, "MB")


if __name__ == "__main__":
    main()
