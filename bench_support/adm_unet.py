"""A guided-diffusion (ADM) style UNet written for the benchmark: the PyTorch *caller* of the
hot path.  Architecture and hyper-parameters are those the reference builds through
``create_model`` (reference src/utils/exp_utils.py:40-97) for its AAPM / ellipses configs
(configs/aapm_configs.py:32-50): 256 base channels, channel multipliers (1,1,2,2,4,4) at
256x256, one residual block per level with scale-shift time conditioning and residual
up/down-sampling, self-attention at 16x16 with 64-channel heads, 1 input channel,
2 output channels of which the first is the noise prediction (reference unet.py:668-669).
Weights are random (no checkpoints offline); the output convolutions the original
zero-initialises get a small normal init so that the score is not identically zero.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def sinusoidal_embedding(t, dim, max_period=10000.0):
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def _norm(ch):
    return nn.GroupNorm(32, ch)


def _small_init(conv, std=0.02):
    nn.init.normal_(conv.weight, std=std)
    nn.init.zeros_(conv.bias)
    return conv


class Res(nn.Module):
    def __init__(self, cin, cout, emb_ch, resample=None):
        super().__init__()
        self.resample = resample                       # None | 'up' | 'down'
        self.n1 = _norm(cin)
        self.c1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.emb = nn.Linear(emb_ch, 2 * cout)
        self.n2 = _norm(cout)
        self.c2 = _small_init(nn.Conv2d(cout, cout, 3, padding=1))
        self.skip = nn.Identity() if cin == cout else nn.Conv2d(cin, cout, 1)

    def _rs(self, v):
        if self.resample == 'up':
            return F.interpolate(v, scale_factor=2, mode='nearest')
        if self.resample == 'down':
            return F.avg_pool2d(v, 2)
        return v

    def forward(self, x, emb):
        h = F.silu(self.n1(x))
        if self.resample is not None:
            h, x = self._rs(h), self._rs(x)
        h = self.c1(h)
        scale, shift = self.emb(F.silu(emb))[:, :, None, None].chunk(2, dim=1)
        h = self.n2(h) * (1 + scale) + shift
        h = self.c2(F.silu(h))
        return self.skip(x) + h


class Attn(nn.Module):
    def __init__(self, ch, head_ch):
        super().__init__()
        self.heads = ch // head_ch
        self.norm = _norm(ch)
        self.qkv = nn.Conv1d(ch, 3 * ch, 1)
        self.proj = _small_init(nn.Conv1d(ch, ch, 1))

    def forward(self, x, emb=None):
        b, c, hh, ww = x.shape
        q, k, v = self.qkv(self.norm(x).reshape(b, c, -1)).reshape(b * self.heads, 3 * c // self.heads, -1).chunk(3, dim=1)
        a = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))
        a = a.transpose(1, 2).reshape(b, c, -1)
        return x + self.proj(a).reshape(b, c, hh, ww)


class Stage(nn.ModuleList):
    def forward(self, x, emb):
        for m in self:
            x = m(x, emb)
        return x


class AdmUNet(nn.Module):
    def __init__(self, image_size=256, in_channels=1, model_channels=256, out_channels=2,
                 num_res_blocks=1, channel_mult=(1, 1, 2, 2, 4, 4), attention_resolutions=(16,),
                 num_head_channels=64, max_period=1e4):
        super().__init__()
        ch = model_channels
        emb_ch = 4 * ch
        self.model_channels = ch
        self.max_period = max_period
        self.out_channels = out_channels
        self.time = nn.Sequential(nn.Linear(ch, emb_ch), nn.SiLU(), nn.Linear(emb_ch, emb_ch))
        attn_ds = {image_size // r for r in attention_resolutions}
        self.stem = nn.Conv2d(in_channels, ch, 3, padding=1)
        self.down = nn.ModuleList()
        skips = [ch]
        cur, ds = ch, 1
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                blk = [Res(cur, mult * ch, emb_ch)]
                cur = mult * ch
                if ds in attn_ds:
                    blk.append(Attn(cur, num_head_channels))
                self.down.append(Stage(blk)); skips.append(cur)
            if level != len(channel_mult) - 1:
                self.down.append(Stage([Res(cur, cur, emb_ch, 'down')])); skips.append(cur)
                ds *= 2
        self.mid = Stage([Res(cur, cur, emb_ch), Attn(cur, num_head_channels), Res(cur, cur, emb_ch)])
        self.up = nn.ModuleList()
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                blk = [Res(cur + skips.pop(), mult * ch, emb_ch)]
                cur = mult * ch
                if ds in attn_ds:
                    blk.append(Attn(cur, num_head_channels))
                if level and i == num_res_blocks:
                    blk.append(Res(cur, cur, emb_ch, 'up'))
                    ds //= 2
                self.up.append(Stage(blk))
        self.head = nn.Sequential(_norm(cur), nn.SiLU(), _small_init(nn.Conv2d(cur, out_channels, 3, padding=1)))

    def forward(self, x, timesteps):
        emb = self.time(sinusoidal_embedding(timesteps, self.model_channels, self.max_period))
        h = self.stem(x)
        hs = [h]
        for blk in self.down:
            h = blk(h, emb)
            hs.append(h)
        h = self.mid(h, emb)
        for blk in self.up:
            h = blk(torch.cat([h, hs.pop()], dim=1), emb)
        out = self.head(h)
        return out[:, :1] if self.out_channels == 2 else out


def small_unet():
    """The CPU-runnable score model of BASELINE config 1."""
    return AdmUNet(image_size=256, model_channels=32, out_channels=1, channel_mult=(1, 2, 4),
                   attention_resolutions=(16,), num_head_channels=32)


def aapm_unet():
    """Full-size model of the AAPM 256x256 config (BASELINE config 2)."""
    return AdmUNet()
