"""B200-native VCA-GAN modules behind the reference's module API.

Class names, constructor signatures, forward signatures and ``state_dict`` keys are those of the reference's
src/models/{resnet,visual_front,generator}.py (cited per class); torch.nn layers appear here only as *parameter
holders* (so keys, shapes and default initialisation match) -- their ``forward`` is never called.  All device
work goes through vcagan_b200.ops -> libvcagan_b200.so.  Activations are channels-last internally and are
converted at the module boundary (reference tensors are NCHW fp32).
"""
import math

import torch
import torch.nn as nn

from . import ops
from .ops import ACT_LRELU, ACT_NONE, ACT_PRELU, ACT_RELU, cfg

INV_SQRT2 = 1.0 / math.sqrt(2.0)


def _in_cl(x_nchw: torch.Tensor) -> torch.Tensor:
    """NCHW (or NCDHW) fp32 -> channels-last compute dtype."""
    nd = x_nchw.dim()
    perm = (0, 2, 3, 1) if nd == 4 else (0, 2, 3, 4, 1)
    return ops.cast(x_nchw.permute(*perm).contiguous(), cfg.dtype)


def _out_nchw(x_cl: torch.Tensor) -> torch.Tensor:
    return ops.cast(x_cl, torch.float32).permute(0, 3, 1, 2).contiguous()


def _conv(x, m: nn.Module, bn: nn.Module = None):
    """Apply the conv whose parameters live in holder `m` (nn.Conv1d/2d/3d) to channels-last x.  `bn`: the BatchNorm the
    output goes straight into -- in training mode the conv bias then has an identically zero gradient (ops.ConvFn)."""
    train_bn = bn is not None and bn.training
    return ops.conv(x, m.weight, m.bias, m.stride, m.padding, zero_bias_grad=train_bn, bn=bn if train_bn else None)


# =============================================================================================================
# resnet.py
# =============================================================================================================
def conv3x3(in_planes, out_planes, stride=1):
    """resnet.py:5-7"""
    return nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=stride, padding=1, bias=False)


def downsample_basic_block(inplanes, outplanes, stride):
    """resnet.py:10-14"""
    return nn.Sequential(nn.Conv2d(inplanes, outplanes, kernel_size=1, stride=stride, bias=False), nn.BatchNorm2d(outplanes))


def downsample_basic_block_v2(inplanes, outplanes, stride):
    """resnet.py:17-22 (parameter holder only; no reference driver enables avg_pool_downsample)."""
    return nn.Sequential(nn.AvgPool2d(kernel_size=stride, stride=stride, ceil_mode=True, count_include_pad=False),
                         nn.Conv2d(inplanes, outplanes, kernel_size=1, stride=1, bias=False), nn.BatchNorm2d(outplanes))


class BasicBlock(nn.Module):
    """resnet.py:25-66.  forward() takes/returns channels-last tensors."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, relu_type='relu'):
        super().__init__()
        assert relu_type in ['relu', 'prelu']
        self.conv1 = conv3x3(inplanes, planes, stride)
        self.bn1 = nn.BatchNorm2d(planes)
        if relu_type == 'relu':
            self.relu1, self.relu2 = nn.ReLU(inplace=True), nn.ReLU(inplace=True)
        else:
            self.relu1, self.relu2 = nn.PReLU(num_parameters=planes), nn.PReLU(num_parameters=planes)
        self.conv2 = conv3x3(planes, planes)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride
        self._prelu = relu_type == 'prelu'

    def _forward_eval_fused(self, x, act):
        """inference: each conv carries its (eval-mode) BatchNorm, the activation and the residual add in its epilogue"""
        pw1 = self.relu1.weight if self._prelu else None
        pw2 = self.relu2.weight if self._prelu else None
        c1, c2 = self.conv1, self.conv2
        out = ops.conv_epi(x, c1.weight, None, c1.stride, c1.padding, bn=self.bn1, act=act, prelu_w=pw1)
        if out is None:
            return None
        res = x
        if self.downsample is not None:
            d = self.downsample[0]
            res = ops.conv_epi(x, d.weight, None, d.stride, d.padding, bn=self.downsample[1])
            if res is None:
                return None
        return ops.conv_epi(out, c2.weight, None, c2.stride, c2.padding, bn=self.bn2, act=act, prelu_w=pw2, res=res)

    def forward(self, x):
        act = ACT_PRELU if self._prelu else ACT_RELU
        if not self.training and ops.epi_ok(x) and (self.downsample is None or len(self.downsample) == 2):
            y = self._forward_eval_fused(x, act)
            if y is not None:
                return y
        out = None
        if not self.training and ops.fold_ok(x) and tuple(self.conv1.stride) == (1, 1):
            # inference: bn1 folded into conv1's weights / bias, the activation in its epilogue (ops.conv_folded)
            out = ops.conv_folded(x, self.conv1.weight, None, self.conv1.padding, self.bn1, act, 0.0, self.relu1.weight if self._prelu else None)
        if out is None:
            out = _conv(x, self.conv1, self.bn1)
            out = ops.bn_act(out, self.bn1, act, 0.0, self.relu1.weight if self._prelu else None)
        out = _conv(out, self.conv2, self.bn2)
        res = x
        if self.downsample is not None:
            if len(self.downsample) != 2:
                raise NotImplementedError("avg_pool_downsample variant is not on the VCA-GAN path")
            res = ops.bn_act(_conv(x, self.downsample[0], self.downsample[1]), self.downsample[1], ACT_NONE)
        return ops.bn_act(out, self.bn2, act, 0.0, self.relu2.weight if self._prelu else None, res=res)


class ResNet(nn.Module):
    """resnet.py:69-123: four stages of BasicBlocks + AvgPool2d(4) + flatten.  Channels-last in, (N,512) out."""

    STAGES = ((64, 1), (128, 2), (256, 2), (512, 2))      # (planes, stride of the stage's first block): layer1 .. layer4

    def __init__(self, block, layers, num_classes=1000, relu_type='relu', gamma_zero=False, avg_pool_downsample=False):
        super().__init__()
        self.relu_type, self.gamma_zero = relu_type, gamma_zero
        shortcut = downsample_basic_block_v2 if avg_pool_downsample else downsample_basic_block
        width = 64
        # the state_dict keys (`layer{i}.{j}.conv1.weight`, `layer{i}.0.downsample.0.weight`, ...) are the interface: built
        # from the stage table, one nn.Sequential per stage
        for i, ((planes, stride), depth) in enumerate(zip(self.STAGES, layers), start=1):
            out = planes * block.expansion
            proj = shortcut(inplanes=width, outplanes=out, stride=stride) if (stride != 1 or width != out) else None
            stage = [block(width, planes, stride, proj, relu_type=relu_type)]
            stage += [block(out, planes, relu_type=relu_type) for _ in range(depth - 1)]
            setattr(self, f"layer{i}", nn.Sequential(*stage))
            width = out
        self.avgpool = nn.AvgPool2d(4)
        self._init_parameters()

    def _init_parameters(self):
        """default initialisation of resnet.py:85-97 (He-normal conv weights by fan-out, unit BatchNorm, optional zero gamma)"""
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                fan_out = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                nn.init.normal_(m.weight, 0.0, math.sqrt(2.0 / fan_out))
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
        if self.gamma_zero:
            for m in self.modules():
                if isinstance(m, BasicBlock):
                    nn.init.zeros_(m.bn2.weight)

    def forward(self, x):
        for layer in (self.layer1, self.layer2, self.layer3, self.layer4):
            for blk in layer:
                x = blk(x)
        if x.shape[1] != 4 or x.shape[2] != 4:
            raise ValueError(f"ResNet trunk expects a 4x4 final map (112x112 lip crops); got {tuple(x.shape[1:3])}")
        return ops.spatial_mean(x)  # AvgPool2d(4) on a 4x4 map + flatten


# =============================================================================================================
# visual_front.py
# =============================================================================================================
class Visual_front(nn.Module):
    """visual_front.py:4-37.  forward(x (B,1,T,112,112)) -> (phons (B,T,512), sentence (B,512,T)), both fp32."""

    def __init__(self, in_channels=1):
        super().__init__()
        self.in_channels = in_channels
        self.frontend = nn.Sequential(
            nn.Conv3d(self.in_channels, 64, kernel_size=(5, 7, 7), stride=(1, 2, 2), padding=(2, 3, 3), bias=False),
            nn.BatchNorm3d(64), nn.PReLU(64), nn.MaxPool3d(kernel_size=(1, 3, 3), stride=(1, 2, 2), padding=(0, 1, 1)))
        self.resnet = ResNet(BasicBlock, [2, 2, 2, 2], relu_type='prelu')
        self.dropout = nn.Dropout(0.3)
        self.sentence_encoder = nn.GRU(512, 512, 2, bidirectional=True, dropout=0.3)
        self.fc = nn.Linear(1024, 512)
        self.drop_masks = None  # parity tests: (feat_mask (B*T,512), gru_mask (T,B,1024)), pre-scaled

    def _gru_params(self, layer):
        g = self.sentence_encoder
        names = [f"{p}_l{layer}{s}" for s in ("", "_reverse") for p in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
        return [getattr(g, n) for n in names]

    def forward(self, x):
        phons = self.features(x)
        return phons, self.sentence(phons)

    # The two halves of forward(), callable separately: the generator's first six blocks need only `phons`, so the trainer
    # runs the (latency-bound, sequential) GRU of sentence() on a side stream underneath them.
    def features(self, x):
        """frontend + ResNet + dropout: x (B,1,T,112,112) -> phons (B,T,512) fp32 (visual_front.py:24-29)"""
        B, _, T = x.shape[:3]
        c0 = self.frontend[0]
        if (cfg.dtype == torch.bfloat16 and cfg.use_tc and self.in_channels == 1 and c0.kernel_size == (5, 7, 7)
                and c0.stride == (1, 2, 2) and c0.padding == (2, 3, 3)):
            x = ops.stem_conv(x, c0.weight, self.frontend[1] if self.frontend[1].training else None,
                              fold=None if self.training else (self.frontend[1], self.frontend[2].weight))   # im2col(7x7) + (5,1) conv on tcgen05
        else:
            x = _conv(_in_cl(x), c0)                    # generic exact path: (B,T,112,112,1) -> (B,T,56,56,64)
        if getattr(x, "_vca_activated", False):           # inference: BN + PReLU already applied in the conv's epilogue
            x = ops.maxpool3x3s2(x.view(B * T, x.shape[2], x.shape[3], x.shape[4]))
        elif ops.bn_prelu_maxpool_supported(x, x.shape[-1]):
            stats = getattr(x, "_vca_bn_sums", None)
            x = x.view(B * T, x.shape[2], x.shape[3], x.shape[4])
            if stats is not None:
                x._vca_bn_sums = stats
            x = ops.bn_prelu_maxpool(x, self.frontend[1], self.frontend[2].weight)   # BN3d + PReLU + MaxPool3d in one pass
        else:
            x = ops.bn_act(x, self.frontend[1], ACT_PRELU, 0.0, self.frontend[2].weight)
            x = ops.maxpool3x3s2(x.view(B * T, x.shape[2], x.shape[3], x.shape[4]))   # (B*T,28,28,64)
        x = self.resnet(x)                              # (B*T,512)
        fm, gm = self.drop_masks if self.drop_masks is not None else (None, None)
        x = ops.dropout(x, self.dropout.p, self.training, fm)
        return ops.cast(x, torch.float32).view(B, T, -1)

    def sentence(self, x):
        """2-layer bi-GRU + fc: phons (B,T,512) -> sentence (B,512,T) fp32 (visual_front.py:30-36)"""
        gm = self.drop_masks[1] if self.drop_masks is not None else None
        phons_tb = x.permute(1, 0, 2).contiguous()      # (T,B,512)
        h = ops.gru_layer(phons_tb, self._gru_params(0))
        h = ops.dropout(h, float(self.sentence_encoder.dropout), self.training, gm)
        h = ops.gru_layer(h, self._gru_params(1))       # (T,B,1024)
        s = ops.cast(ops.linear(ops.cast(h, cfg.dtype), self.fc.weight, self.fc.bias), torch.float32)  # (T,B,512)
        return s.permute(1, 2, 0).contiguous()


# =============================================================================================================
# generator.py -- generator half
# =============================================================================================================
class ResBlk1D(nn.Module):
    """generator.py:8-49.  Channels-last (B,1,L,C) in/out."""

    def __init__(self, dim_in, dim_out, actv=None, normalize=False, downsample=False):
        super().__init__()
        if normalize or downsample:
            raise NotImplementedError("the reference only instantiates ResBlk1D(normalize=False, downsample=False)")
        self.normalize, self.downsample = normalize, downsample
        self.learned_sc = dim_in != dim_out
        self.conv1 = nn.Conv1d(dim_in, dim_in, 5, 1, 2)
        self.conv2 = nn.Conv1d(dim_in, dim_out, 5, 1, 2)
        if self.learned_sc:
            self.conv1x1 = nn.Conv1d(dim_in, dim_out, 1, 1, 0, bias=False)

    def forward(self, x):
        r = _conv(ops.lrelu(x), self.conv1)
        r = _conv(ops.lrelu(r), self.conv2)
        s = _conv(x, self.conv1x1) if self.learned_sc else x
        return ops.add_scale(s, r, INV_SQRT2)


class ResBlk(nn.Module):
    """generator.py:51-92 (discriminator block, no normalisation).  Channels-last in/out."""

    def __init__(self, dim_in, dim_out, actv=None, normalize=False, downsample=False):
        super().__init__()
        if normalize:
            raise NotImplementedError("the reference only instantiates ResBlk(normalize=False)")
        self.normalize, self.downsample = normalize, downsample
        self.learned_sc = dim_in != dim_out
        self.conv1 = nn.Conv2d(dim_in, dim_in, 5, 1, 2)
        self.conv2 = nn.Conv2d(dim_in, dim_out, 5, 1, 2)
        if self.learned_sc:
            self.conv1x1 = nn.Conv2d(dim_in, dim_out, 1, 1, 0, bias=False)

    def forward(self, x):
        r = _conv(ops.lrelu(x), self.conv1)
        if self.downsample:
            r = ops.avg_pool2(r)
        r = _conv(ops.lrelu(r), self.conv2)
        s = _conv(x, self.conv1x1) if self.learned_sc else x
        if self.downsample:
            s = ops.avg_pool2(s)
        return ops.add_scale(s, r, INV_SQRT2)


class GenResBlk(nn.Module):
    """generator.py:94-131: BN -> LReLU -> [up] -> 5x5 -> BN -> LReLU -> 5x5, + ([up] -> 1x1), / sqrt(2)."""

    def __init__(self, dim_in, dim_out, actv=None, upsample=False):
        super().__init__()
        self.upsample = upsample
        self.learned_sc = dim_in != dim_out
        self.conv1 = nn.Conv2d(dim_in, dim_out, 5, 1, 2)
        self.conv2 = nn.Conv2d(dim_out, dim_out, 5, 1, 2)
        self.norm1 = nn.BatchNorm2d(dim_in)
        self.norm2 = nn.BatchNorm2d(dim_out)
        if self.learned_sc:
            self.conv1x1 = nn.Conv2d(dim_in, dim_out, 1, 1, 0, bias=False)

    def forward(self, x, const_channels=0):
        """const_channels > 0: the caller guarantees that the first `const_channels` channels of x are constant along
        dim 1 (the tiled phoneme features of generator.py:249-250).  BN + LeakyReLU are per-channel maps, so that still
        holds at conv1's input and the constant part of conv1 collapses to one row (ops.conv_rowconst): an exact
        saving the reference does not take (SURVEY appendix A #14 i)."""
        r = ops.bn_act(x, self.norm1, ACT_LRELU, 0.2)
        if self.upsample:
            r = ops.upsample2(r)
        if not self.training and ops.epi_ok(r) and not (const_channels and not self.upsample and cfg.rowconst):
            # inference: norm2 + LeakyReLU ride in conv1's epilogue, the shortcut add and 1/sqrt(2) in conv2's
            c1, c2 = self.conv1, self.conv2
            r1 = ops.conv_epi(r, c1.weight, c1.bias, c1.stride, c1.padding, bn=self.norm2, act=ACT_LRELU, slope=0.2)
            if r1 is not None:
                s = ops.upsample2(x) if self.upsample else x
                if self.learned_sc:
                    s = _conv(s, self.conv1x1)
                y = ops.conv_epi(r1, c2.weight, c2.bias, c2.stride, c2.padding, res=s, res_scale=INV_SQRT2, out_scale=INV_SQRT2)
                if y is not None:
                    return y
                r = r1          # conv2 has no fused route: finish on the separate kernels
                r = _conv(r, self.conv2)
                return ops.add_scale(r, s, INV_SQRT2)
        rf = None
        if not self.training and ops.fold_ok(r) and not (const_channels and not self.upsample and cfg.rowconst):
            # inference: norm2 folded into conv1's weights / bias, LeakyReLU in its epilogue (ops.conv_folded)
            rf = ops.conv_folded(r, self.conv1.weight, self.conv1.bias, self.conv1.padding, self.norm2, ACT_LRELU, 0.2)
        if rf is not None:
            r = rf
        else:
            if const_channels and not self.upsample and cfg.rowconst:
                r = ops.conv_rowconst(r, const_channels, self.conv1.weight, self.conv1.bias, tuple(self.conv1.padding),
                                      zero_bias_grad=self.norm2.training)
            else:
                r = _conv(r, self.conv1, self.norm2)
            r = ops.bn_act(r, self.norm2, ACT_LRELU, 0.2)
        r = _conv(r, self.conv2)
        s = ops.upsample2(x) if self.upsample else x
        if self.learned_sc:
            s = _conv(s, self.conv1x1)
        return ops.add_scale(r, s, INV_SQRT2)


class Flatten(nn.Module):
    """generator.py:133-135"""

    def forward(self, input):
        return input.reshape(input.size(0), -1)


class Avgpool(nn.Module):
    """generator.py:137-140 -- mean over the spatial dims of a channels-last tensor."""

    def forward(self, input):
        return ops.spatial_mean(input)


def _lens_tensor(lens, device):
    if torch.is_tensor(lens):
        return lens.to(device=device, dtype=torch.int32).contiguous()
    return torch.tensor([int(v) for v in lens], dtype=torch.int32, device=device)


class AVAttention(nn.Module):
    """generator.py:142-171.  forward(ph (B,S,512) fp32, g channels-last (B,F,T,C), len) -> channels-last (B,F,T,C')."""

    def __init__(self, out_dim):
        super().__init__()
        self.softmax = nn.Softmax(2)
        self.k = nn.Linear(512, out_dim)
        self.v = nn.Linear(512, out_dim)
        self.q = nn.Linear(2560, out_dim)
        self.out_dim = out_dim
        self.mel = nn.Linear(out_dim, 20 * 64)

    def forward(self, ph, g, len):
        B, Fq, T, C = g.shape
        lens = _lens_tensor(len, g.device)
        f32 = torch.float32
        ph = ops.cast(ph, cfg.dtype)                                         # projections run in the compute dtype
        k = ops.linear(ph, self.k.weight, self.k.bias)                       # (B,S,256)
        v = ops.linear(ph, self.v.weight, self.v.bias)
        gq = g.permute(0, 2, 3, 1).reshape(B, T, C * Fq)                     # index c*F+f as in g.view(B,C*F,T)
        q = ops.linear(gq, self.q.weight, self.q.bias)                       # (B,T,256)
        if ops.attention_supported(q, k):
            # bf16 mode: QK^T -> key mask -> softmax -> PV in one tcgen05 kernel (scores / softmax in fp32 on chip)
            val = ops.attention(q, k, v, lens, 1.0 / math.sqrt(self.out_dim))
            out = ops.linear(val, self.mel.weight, self.mel.bias)            # (B,T,1280)
            return out.view(B, T, Fq, -1).permute(0, 2, 1, 3).contiguous()   # channels-last (B,F,T,C')
        q, k, v = ops.cast(q, f32), ops.cast(k, f32), ops.cast(v, f32)
        att = ops.bmm(q, k.transpose(1, 2), 1.0 / math.sqrt(self.out_dim))   # (B,T,S)  scores/softmax in fp32
        att = ops.masked_softmax(att, lens)
        val = ops.bmm(att, v)                                                # (B,T,256)
        out = ops.linear(ops.cast(val, cfg.dtype), self.mel.weight, self.mel.bias)   # (B,T,1280)
        return out.view(B, T, Fq, -1).permute(0, 2, 1, 3).contiguous()       # channels-last (B,F,T,C')


class Postnet(nn.Module):
    """generator.py:173-192.  forward((B,1,80,L)) -> (B,1,321,L)."""

    def __init__(self):
        super().__init__()
        self.postnet = nn.Sequential(nn.Conv1d(80, 128, 7, 1, 3), nn.BatchNorm1d(128), nn.LeakyReLU(0.2), ResBlk1D(128, 256),
                                     ResBlk1D(256, 256), ResBlk1D(256, 256), nn.Conv1d(256, 321, 1, 1, 0, bias=False))

    def forward(self, x):
        p = self.postnet
        B, _, Fm, L = x.shape
        h = ops.cast(x.reshape(B, Fm, L).permute(0, 2, 1).contiguous().view(B, 1, L, Fm), cfg.dtype)
        h = _conv(h, p[0], p[1])
        h = ops.bn_act(h, p[1], ACT_LRELU, 0.2)
        for i in (3, 4, 5):
            h = p[i](h)
        h = _conv(h, p[6])                                                   # (B,1,L,321)
        return ops.cast(h, torch.float32).view(B, L, -1).permute(0, 2, 1).contiguous().unsqueeze(1)


class _ToMel(nn.Sequential):
    """to_mel{1,2,3} of generator.py:208-225: BN -> LReLU -> 1x1 conv (C -> 1) -> tanh."""

    def __init__(self, c):
        super().__init__(nn.BatchNorm2d(c), nn.LeakyReLU(0.2), nn.Conv2d(c, 1, 1, 1, 0), nn.Tanh())

    def forward(self, x):
        h = ops.bn_act(x, self[0], ACT_LRELU, 0.2)
        h = ops.tanh(_conv(h, self[2]))                                      # (B,F,T,1)
        return _out_nchw(h)


class Decoder(nn.Module):
    """generator.py:194-265.  forward(s (B,512,T), x (B,T,512), len) -> g1 (B,1,20,T), g2 (B,1,40,2T), g3 (B,1,80,4T)."""

    def __init__(self):
        super().__init__()
        self.decode = nn.ModuleList()
        self.g1 = nn.ModuleList()
        self.g2 = nn.ModuleList()
        self.g3 = nn.ModuleList()
        self.att1 = AVAttention(256)
        self.attconv1 = nn.Conv2d(128 + 64, 128, 5, 1, 2)
        self.att2 = AVAttention(256)
        self.attconv2 = nn.Conv2d(64 + 32, 64, 5, 1, 2)
        self.to_mel1 = _ToMel(128)
        self.to_mel2 = _ToMel(64)
        self.to_mel3 = _ToMel(32)
        self.decode.append(GenResBlk(512 + 128, 512))
        self.decode.append(GenResBlk(512, 256))
        self.decode.append(GenResBlk(256, 256))
        self.g1.append(GenResBlk(256, 128))
        self.g1.append(GenResBlk(128, 128))
        self.g1.append(GenResBlk(128, 128))
        self.g2.append(GenResBlk(128, 64, upsample=True))
        self.g2.append(GenResBlk(64, 64))
        self.g2.append(GenResBlk(64, 64))
        self.g3.append(GenResBlk(64, 32, upsample=True))
        self.g3.append(GenResBlk(32, 32))
        self.g3.append(GenResBlk(32, 32))
        self.fixed_noise = None     # parity tests: (B,128,20,T) replaces the N(0,1) draw of generator.py:248
        self.noise_source = "device"  # "device": Philox kernel; "host": torch.randn on the CPU like the reference

    def _noise(self, B, T, device):
        if self.fixed_noise is not None:
            n = self.fixed_noise.to(device)
        elif self.noise_source == "host":
            n = torch.randn([B, 128, 20, T]).to(device)
        else:
            return ops.randn((B, 20, T, 128), cfg.dtype, device)
        return ops.cast(n.permute(0, 2, 3, 1).contiguous(), cfg.dtype)

    def forward(self, s, x, len):
        return self.tail(s, self.stem(x), len)

    # forward() in two stages: stem() needs only the phoneme features, tail() is everything from the first attention on.
    def stem(self, x):
        """noise + tiling + decode x3 + g1 x3 (generator.py:246-255): x (B,T,512) -> channels-last (B,20,T,128)"""
        B, T = x.size(0), x.size(1)
        n = self._noise(B, T, x.device)                                      # (B,20,T,128)
        xt = ops.spatial_tile(ops.cast(x.contiguous(), cfg.dtype).view(B, T * x.size(2)), 20).view(B, 20, T, x.size(2))
        h = torch.cat([xt, n], 3)                                            # (B,20,T,640)
        for i, blk in enumerate(self.decode):
            h = blk(h, const_channels=x.size(2)) if i == 0 else blk(h)       # xt is constant along the 20 mel rows
        for blk in self.g1:
            h = blk(h)
        return h

    def tail(self, s, h, len):
        """att1 ... to_mel3 (generator.py:256-265); s (B,512,T) sentence embedding, h = stem() output"""
        s = s.transpose(1, 2).contiguous()                                   # (B,T,512)
        f1 = h
        c1 = self.att1(s, f1, len)
        h = _conv(torch.cat([h, c1], 3), self.attconv1)
        for blk in self.g2:
            h = blk(h)
        f2 = h
        c2 = self.att2(s, f2, len)
        h = _conv(torch.cat([h, c2], 3), self.attconv2)
        for blk in self.g3:
            h = blk(h)
        return self.to_mel1(f1), self.to_mel2(f2), self.to_mel3(h)


# =============================================================================================================
# generator.py -- discriminator half
# =============================================================================================================
class Discriminator(nn.Module):
    """generator.py:267-317.  forward(x (B,1,H,W), c (B,512,T), vid_max_length) -> (uout (B,1), cout (B,1))."""

    def __init__(self, num_class=1, max_conv_dim=512, phase='1'):
        super().__init__()
        dim_in = 32
        blocks = [nn.Conv2d(1, dim_in, 5, 1, 2)]
        repeat_num = 2 if phase == '1' else (3 if phase == '2' else 4)
        for _ in range(repeat_num):
            dim_out = min(dim_in * 2, max_conv_dim)
            blocks += [ResBlk(dim_in, dim_out, downsample=True)]
            dim_in = dim_out
        self.main = nn.Sequential(*blocks)
        self.uncond = nn.Sequential(nn.LeakyReLU(0.2), nn.Conv2d(dim_out, dim_out, 5, 1, 0), nn.LeakyReLU(0.2), Avgpool(),
                                    nn.Linear(dim_out, num_class))
        self.cond = nn.Sequential(nn.LeakyReLU(0.2), nn.Conv2d(dim_out + 512, dim_out, 5, 1, 2), nn.LeakyReLU(0.2),
                                  nn.Conv2d(dim_out, dim_out, 5, 1, 0), nn.LeakyReLU(0.2), Avgpool(), nn.Linear(dim_out, num_class))

    def forward(self, x, c, vid_max_length):
        h = self.features(x, vid_max_length)
        return self.uncond_head(h), self.cond_head(h, c, vid_max_length)

    # The three stages of forward(), callable separately: only cond_head needs the sentence embedding, so the trainer
    # can start the trunk, the unconditional head and the R1 penalty of the real pass before the visual front-end ends.
    def features(self, x, vid_max_length):
        f_len = final_length(vid_max_length)
        h = _in_cl(x)
        h = _conv(h, self.main[0])
        for blk in list(self.main)[1:]:
            h = blk(h)                                                       # (B,5,f_len,C)
        if h.shape[1] != 5 or h.shape[2] != f_len:
            raise ValueError(f"discriminator map {tuple(h.shape[1:3])} does not match (5, final_length={f_len})")
        return h

    def uncond_head(self, h):
        u = _conv(ops.lrelu(h), self.uncond[1])
        u = ops.cast(ops.spatial_mean(ops.lrelu(u)), torch.float32)
        u = ops.linear(u, self.uncond[4].weight, self.uncond[4].bias)
        return u.view(h.size(0), -1)

    def cond_head(self, h, c, vid_max_length):
        f_len = final_length(vid_max_length)
        B = h.size(0)
        cm = ops.spatial_mean(ops.cast(c, torch.float32).permute(0, 2, 1).contiguous().unsqueeze(1))   # (B,512)
        ct = ops.cast(ops.spatial_tile(cm, 5 * f_len), cfg.dtype).view(B, 5, f_len, -1)
        k = _conv(ops.lrelu(torch.cat([h, ct], 3)), self.cond[1])
        k = _conv(ops.lrelu(k), self.cond[3])
        k = ops.cast(ops.spatial_mean(ops.lrelu(k)), torch.float32)
        k = ops.linear(k, self.cond[6].weight, self.cond[6].bias)
        return k.view(B, -1)


class sync_Discriminator(nn.Module):
    """generator.py:319-361.  forward(v_feat (B,S,512), aud (B,1,80,4S), gen=False) -> (B,)."""

    def __init__(self, temp=1.0):
        super().__init__()
        self.frontend = nn.Sequential(
            nn.Conv2d(1, 128, kernel_size=(3, 3), stride=(2, 2), padding=(1, 1)), nn.BatchNorm2d(128), nn.PReLU(128),
            nn.Conv2d(128, 256, kernel_size=(3, 3), stride=(2, 2), padding=(1, 1)), nn.BatchNorm2d(256), nn.PReLU(256))
        self.Res_block = nn.Sequential(BasicBlock(256, 256))
        self.Linear = nn.Linear(256 * 20, 512)
        self.temp = temp

    def forward(self, v_feat, aud, gen=False):
        f = self.frontend
        a = _in_cl(aud)
        a = ops.bn_act(_conv(a, f[0], f[1]), f[1], ACT_PRELU, 0.0, f[2].weight)
        a = ops.bn_act(_conv(a, f[3], f[4]), f[4], ACT_PRELU, 0.0, f[5].weight)
        a = self.Res_block[0](a)                                             # (B,20,S,256)
        B, Fq, S, C = a.shape
        a = a.permute(0, 2, 3, 1).reshape(B, S, C * Fq)                      # index c*F+f (generator.py:344)
        a = ops.cast(ops.linear(a, self.Linear.weight, self.Linear.bias), torch.float32)   # (B,S,512)
        v = ops.cast(v_feat, torch.float32)
        if gen:
            return ops.CosAbsMeanFn.apply(v, a)
        vn, an = ops.L2NormFn.apply(v), ops.L2NormFn.apply(a)
        sim = ops.bmm(vn, an.transpose(1, 2), 1.0 / self.temp)
        return ops.NceDiagFn.apply(sim)


def gan_loss(inputs, label=None):
    """generator.py:363-366"""
    return ops.SoftplusMeanFn.apply(ops.cast(inputs, torch.float32), -1.0 if label else 1.0)


def final_length(vid_length):
    """generator.py:368-371"""
    return (vid_length // 2) // 2
