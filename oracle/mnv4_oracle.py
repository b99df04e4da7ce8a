"""TEST INFRASTRUCTURE ONLY -- plain PyTorch restatement of the MobileNetV4-conv-small ``features_only`` encoder that the reference
obtains from timm (estimator/models/blocks/lightweight_refiner.py:259-262, stem surgery estimator/models/patchrefinerplus.py:159-165).

PARITY UNPINNED: timm is neither vendored under /root/reference nor installable offline (the reference pins ``timm==0.9.2`` in
environment.yml:27, which predates MobileNetV4; the configs need a timm >= 1.0 release), so this module restates the PUBLISHED
architecture (MobileNetV4 paper, MNv4-Conv-S; timm's ``_gen_mobilenet_v4`` 'small' block strings, quoted in
patchrefinerv2_b200/mnv4.py) with timm's module / state-dict names.  It is the checker for the CUDA encoder, not a timm oracle."""
import torch
import torch.nn as nn
import torch.nn.functional as F


def make_divisible(v, divisor=8, round_limit=0.9):
    new_v = max(divisor, int(v + divisor / 2) // divisor * divisor)
    if new_v < round_limit * v:
        new_v += divisor
    return new_v


class ConvNormAct(nn.Module):                      # timm.layers.ConvNormAct: .conv, .bn (BatchNormAct2d: BN then the activation)
    def __init__(self, cin, cout, k, stride=1, groups=1, act=True):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, stride, k // 2, groups=groups, bias=False)
        self.bn = nn.BatchNorm2d(cout, eps=1e-5)
        self.act = act

    def forward(self, x):
        x = self.bn(self.conv(x))
        return F.relu(x) if self.act else x


class ConvBnAct(nn.Module):                        # timm _efficientnet_blocks.ConvBnAct: .conv, .bn1; no skip for these blocks (in != out or stride 2)
    def __init__(self, cin, cout, k, stride):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, stride, k // 2, bias=False)
        self.bn1 = nn.BatchNorm2d(cout, eps=1e-5)
        self.has_skip = False

    def forward(self, x):
        return F.relu(self.bn1(self.conv(x)))


class UniversalInvertedResidual(nn.Module):        # timm _efficientnet_blocks.UniversalInvertedResidual (no SE, no layer scale, no dw_end)
    def __init__(self, cin, cout, k_start, k_mid, stride, exp_ratio):
        super().__init__()
        self.has_skip = cin == cout and stride == 1
        if k_start:
            self.dw_start = ConvNormAct(cin, cin, k_start, stride if not k_mid else 1, groups=cin, act=False)
        else:
            self.dw_start = nn.Identity()
        mid = make_divisible(cin * exp_ratio)
        self.pw_exp = ConvNormAct(cin, mid, 1)
        if k_mid:
            self.dw_mid = ConvNormAct(mid, mid, k_mid, stride, groups=mid)
        else:
            self.dw_mid = nn.Identity()
        self.pw_proj = ConvNormAct(mid, cout, 1, act=False)

    def forward(self, x):
        y = self.pw_proj(self.dw_mid(self.pw_exp(self.dw_start(x))))
        return y + x if self.has_skip else y


ARCH = [
    [("cn", 3, 2, 32), ("cn", 1, 1, 32)],
    [("cn", 3, 2, 96), ("cn", 1, 1, 64)],
    [("uir", 5, 5, 2, 3.0, 96)] + [("uir", 0, 3, 1, 2.0, 96)] * 4 + [("uir", 3, 0, 1, 4.0, 96)],
    [("uir", 3, 3, 2, 6.0, 128), ("uir", 5, 5, 1, 4.0, 128), ("uir", 0, 5, 1, 4.0, 128), ("uir", 0, 5, 1, 3.0, 128),
     ("uir", 0, 3, 1, 4.0, 128), ("uir", 0, 3, 1, 4.0, 128)],
    [("cn", 1, 1, 960)],
]


class MobileNetV4ConvSmallFeatures(nn.Module):
    """features_only=True, out_indices=(0..4): [stem (32, /2), stage 0 (32, /4), stage 1 (64, /8), stage 2 (96, /16), stage 4 (960, /32)]."""
    default_cfg = dict(mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225))

    def __init__(self, in_chans=4):
        super().__init__()
        self.conv_stem = nn.Conv2d(in_chans, 32, 3, 2, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(32, eps=1e-5)
        stages, cin = [], 32
        for stage in ARCH:
            blocks = []
            for b in stage:
                if b[0] == "cn":
                    blocks.append(ConvBnAct(cin, b[3], b[1], b[2]))
                    cin = b[3]
                else:
                    blocks.append(UniversalInvertedResidual(cin, b[5], b[1], b[2], b[3], b[4]))
                    cin = b[5]
            stages.append(nn.Sequential(*blocks))
        self.blocks = nn.Sequential(*stages)

    def forward(self, x):
        x = F.relu(self.bn1(self.conv_stem(x)))
        feats = [x]
        for i, st in enumerate(self.blocks):
            x = st(x)
            if i in (0, 1, 2, 4):
                feats.append(x)
        return feats


def init_healthy(m: nn.Module, seed: int = 0) -> nn.Module:
    """Random weights whose BatchNorm statistics are CALIBRATED on one random batch (train-mode pass with momentum 1), so the
    activations stay O(1) through the 17 blocks and every feature level carries test signal."""
    g = torch.Generator().manual_seed(seed)
    for mod in m.modules():
        if isinstance(mod, nn.Conv2d):
            fan = mod.weight.shape[1] * mod.weight.shape[2] * mod.weight.shape[3]
            mod.weight.data = torch.randn(mod.weight.shape, generator=g) * (2.0 / fan) ** 0.5
        elif isinstance(mod, nn.BatchNorm2d):
            c = mod.num_features
            mod.weight.data = 0.75 + 0.5 * torch.rand(c, generator=g)
            mod.bias.data = 0.2 * torch.randn(c, generator=g)
            mod.momentum = 1.0
    m.train()
    with torch.no_grad():
        m(torch.rand(2, m.conv_stem.in_channels, 128, 128, generator=g) * 2 - 0.5)
    for mod in m.modules():
        if isinstance(mod, nn.BatchNorm2d):
            mod.momentum = 0.1
    return m.eval()
