"""TEST INFRASTRUCTURE ONLY -- restatement of the published SwinV2 algorithm of
``timm==0.6.12`` (``timm/models/swin_transformer_v2.py``), the pinned third-party
dependency the reference takes its encoder from (reference requirements.txt:12,
call sites SOccDPT/model/backbones/swin2.py:7-27, hook points
SOccDPT/model/backbones/swin_common.py:16-27, driver
SOccDPT/model/backbones/utils.py:64-81).

Module tree / parameter names follow SURVEY.md Appendix A.1 so that MiDaS /
SOccDPT checkpoints keyed ``pretrained.model.*`` load.  Parity of this file is
"unpinned" by the reference (no tests, no weights); it is cross-checked against
``transformers.Swinv2Model`` by tests/test_oracle_swinv2_vs_hf.py.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def _pair(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


def window_partition(x, window_size):
    B, H, W, C = x.shape
    x = x.view(B, H // window_size[0], window_size[0], W // window_size[1], window_size[1], C)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, window_size[0], window_size[1], C)


def window_reverse(windows, window_size, img_size):
    H, W = img_size
    B = int(windows.shape[0] / (H * W / window_size[0] / window_size[1]))
    x = windows.view(B, H // window_size[0], W // window_size[1], window_size[0], window_size[1], -1)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(B, H, W, -1)


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = nn.GELU()
        self.drop1 = nn.Dropout(0.0)
        self.fc2 = nn.Linear(hidden_features, in_features)
        self.drop2 = nn.Dropout(0.0)

    def forward(self, x):
        return self.drop2(self.fc2(self.drop1(self.act(self.fc1(x)))))


class WindowAttention(nn.Module):
    """Cosine window attention with log-spaced continuous position bias."""

    def __init__(self, dim, window_size, num_heads, pretrained_window_size=(0, 0)):
        super().__init__()
        self.dim = dim
        self.window_size = window_size
        self.pretrained_window_size = pretrained_window_size
        self.num_heads = num_heads

        self.logit_scale = nn.Parameter(torch.log(10 * torch.ones((num_heads, 1, 1))))
        self.cpb_mlp = nn.Sequential(
            nn.Linear(2, 512, bias=True), nn.ReLU(inplace=True), nn.Linear(512, num_heads, bias=False))

        rh = torch.arange(-(window_size[0] - 1), window_size[0], dtype=torch.float32)
        rw = torch.arange(-(window_size[1] - 1), window_size[1], dtype=torch.float32)
        table = torch.stack(torch.meshgrid([rh, rw], indexing="ij")).permute(1, 2, 0).contiguous().unsqueeze(0)
        if pretrained_window_size[0] > 0:
            table[:, :, :, 0] /= pretrained_window_size[0] - 1
            table[:, :, :, 1] /= pretrained_window_size[1] - 1
        else:
            table[:, :, :, 0] /= window_size[0] - 1
            table[:, :, :, 1] /= window_size[1] - 1
        table *= 8
        table = torch.sign(table) * torch.log2(torch.abs(table) + 1.0) / math.log2(8)
        self.register_buffer("relative_coords_table", table, persistent=False)

        ch = torch.arange(window_size[0])
        cw = torch.arange(window_size[1])
        coords = torch.flatten(torch.stack(torch.meshgrid([ch, cw], indexing="ij")), 1)
        rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
        rel[:, :, 0] += window_size[0] - 1
        rel[:, :, 1] += window_size[1] - 1
        rel[:, :, 0] *= 2 * window_size[1] - 1
        self.register_buffer("relative_position_index", rel.sum(-1), persistent=False)

        self.qkv = nn.Linear(dim, dim * 3, bias=False)
        self.q_bias = nn.Parameter(torch.zeros(dim))
        self.register_buffer("k_bias", torch.zeros(dim), persistent=False)
        self.v_bias = nn.Parameter(torch.zeros(dim))
        self.attn_drop = nn.Dropout(0.0)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(0.0)
        self.softmax = nn.Softmax(dim=-1)

    def relative_position_bias(self):
        n = self.window_size[0] * self.window_size[1]
        tab = self.cpb_mlp(self.relative_coords_table).view(-1, self.num_heads)
        bias = tab[self.relative_position_index.view(-1)].view(n, n, -1).permute(2, 0, 1).contiguous()
        return 16 * torch.sigmoid(bias)

    def forward(self, x, mask=None):
        B_, N, C = x.shape
        qkv_bias = torch.cat((self.q_bias, self.k_bias, self.v_bias))
        qkv = F.linear(input=x, weight=self.qkv.weight, bias=qkv_bias)
        qkv = qkv.reshape(B_, N, 3, self.num_heads, -1).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)

        attn = F.normalize(q, dim=-1) @ F.normalize(k, dim=-1).transpose(-2, -1)
        logit_scale = torch.clamp(self.logit_scale, max=math.log(1.0 / 0.01)).exp()
        attn = attn * logit_scale
        attn = attn + self.relative_position_bias().unsqueeze(0)

        if mask is not None:
            nW = mask.shape[0]
            attn = attn.view(B_ // nW, nW, self.num_heads, N, N) + mask.unsqueeze(1).unsqueeze(0)
            attn = attn.view(-1, self.num_heads, N, N)
        attn = self.attn_drop(self.softmax(attn))
        x = (attn @ v).transpose(1, 2).reshape(B_, N, C)
        return self.proj_drop(self.proj(x))


class SwinTransformerBlock(nn.Module):
    def __init__(self, dim, input_resolution, num_heads, window_size, shift_size, mlp_ratio,
                 pretrained_window_size):
        super().__init__()
        self.dim = dim
        self.input_resolution = _pair(input_resolution)
        self.num_heads = num_heads
        tw, ts = _pair(window_size), _pair(shift_size)
        ws = [r if r <= w else w for r, w in zip(self.input_resolution, tw)]
        ss = [0 if r <= w else s for r, w, s in zip(self.input_resolution, ws, ts)]
        self.window_size, self.shift_size = tuple(ws), tuple(ss)
        self.window_area = self.window_size[0] * self.window_size[1]
        self.mlp_ratio = mlp_ratio

        self.attn = WindowAttention(dim, self.window_size, num_heads, _pair(pretrained_window_size))
        self.norm1 = nn.LayerNorm(dim)
        self.drop_path1 = nn.Identity()
        self.mlp = Mlp(dim, int(dim * mlp_ratio))
        self.norm2 = nn.LayerNorm(dim)
        self.drop_path2 = nn.Identity()

        if any(self.shift_size):
            H, W = self.input_resolution
            img_mask = torch.zeros((1, H, W, 1))
            cnt = 0
            for h in (slice(0, -self.window_size[0]), slice(-self.window_size[0], -self.shift_size[0]),
                      slice(-self.shift_size[0], None)):
                for w in (slice(0, -self.window_size[1]), slice(-self.window_size[1], -self.shift_size[1]),
                          slice(-self.shift_size[1], None)):
                    img_mask[:, h, w, :] = cnt
                    cnt += 1
            mw = window_partition(img_mask, self.window_size).view(-1, self.window_area)
            attn_mask = mw.unsqueeze(1) - mw.unsqueeze(2)
            attn_mask = attn_mask.masked_fill(attn_mask != 0, float(-100.0)).masked_fill(attn_mask == 0, float(0.0))
        else:
            attn_mask = None
        self.register_buffer("attn_mask", attn_mask)

    def _attn(self, x):
        H, W = self.input_resolution
        B, L, C = x.shape
        x = x.view(B, H, W, C)
        has_shift = any(self.shift_size)
        if has_shift:
            x = torch.roll(x, shifts=(-self.shift_size[0], -self.shift_size[1]), dims=(1, 2))
        xw = window_partition(x, self.window_size).view(-1, self.window_area, C)
        aw = self.attn(xw, mask=self.attn_mask)
        aw = aw.view(-1, self.window_size[0], self.window_size[1], C)
        x = window_reverse(aw, self.window_size, self.input_resolution)
        if has_shift:
            x = torch.roll(x, shifts=self.shift_size, dims=(1, 2))
        return x.view(B, H * W, C)

    def forward(self, x):
        x = x + self.drop_path1(self.norm1(self._attn(x)))
        x = x + self.drop_path2(self.norm2(self.mlp(x)))
        return x


class PatchMerging(nn.Module):
    def __init__(self, input_resolution, dim):
        super().__init__()
        self.input_resolution = input_resolution
        self.dim = dim
        self.reduction = nn.Linear(4 * dim, 2 * dim, bias=False)
        self.norm = nn.LayerNorm(2 * dim)

    def forward(self, x):
        H, W = self.input_resolution
        B, L, C = x.shape
        x = x.view(B, H, W, C)
        x = torch.cat([x[:, 0::2, 0::2, :], x[:, 1::2, 0::2, :], x[:, 0::2, 1::2, :], x[:, 1::2, 1::2, :]], -1)
        x = x.view(B, -1, 4 * C)
        return self.norm(self.reduction(x))


class BasicLayer(nn.Module):
    def __init__(self, dim, input_resolution, depth, num_heads, window_size, mlp_ratio, downsample,
                 pretrained_window_size):
        super().__init__()
        self.dim, self.input_resolution, self.depth = dim, input_resolution, depth
        self.blocks = nn.ModuleList([
            SwinTransformerBlock(dim, input_resolution, num_heads, window_size,
                                 0 if (i % 2 == 0) else window_size // 2, mlp_ratio, pretrained_window_size)
            for i in range(depth)])
        self.downsample = PatchMerging(input_resolution, dim) if downsample else nn.Identity()

    def forward(self, x):
        for blk in self.blocks:
            x = blk(x)
        return self.downsample(x)

    def _init_respostnorm(self):
        for blk in self.blocks:
            for n in (blk.norm1, blk.norm2):
                nn.init.constant_(n.bias, 0)
                nn.init.constant_(n.weight, 0)


class PatchEmbed(nn.Module):
    def __init__(self, img_size, patch_size, in_chans, embed_dim):
        super().__init__()
        self.img_size, self.patch_size = _pair(img_size), _pair(patch_size)
        self.grid_size = (self.img_size[0] // self.patch_size[0], self.img_size[1] // self.patch_size[1])
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=self.patch_size, stride=self.patch_size)
        self.norm = nn.LayerNorm(embed_dim)

    def forward(self, x):
        B, C, H, W = x.shape
        assert H == self.img_size[0] and W == self.img_size[1], "Input image size doesn't match model"
        return self.norm(self.proj(x).flatten(2).transpose(1, 2))


class SwinTransformerV2(nn.Module):
    def __init__(self, img_size=224, patch_size=4, in_chans=3, num_classes=1000, embed_dim=96,
                 depths=(2, 2, 6, 2), num_heads=(3, 6, 12, 24), window_size=7, mlp_ratio=4.0,
                 pretrained_window_sizes=(0, 0, 0, 0), **kwargs):
        super().__init__()
        self.num_classes = num_classes
        self.num_layers = len(depths)
        self.embed_dim = embed_dim
        self.num_features = int(embed_dim * 2 ** (self.num_layers - 1))
        self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        self.absolute_pos_embed = None
        self.pos_drop = nn.Dropout(0.0)
        grid = self.patch_embed.grid_size
        self.layers = nn.ModuleList([
            BasicLayer(int(embed_dim * 2 ** i), (grid[0] // (2 ** i), grid[1] // (2 ** i)), depths[i],
                       num_heads[i], window_size, mlp_ratio, i < self.num_layers - 1,
                       pretrained_window_sizes[i])
            for i in range(self.num_layers)])
        self.norm = nn.LayerNorm(self.num_features)
        self.head = nn.Linear(self.num_features, num_classes) if num_classes > 0 else nn.Identity()
        self.apply(self._init_weights)
        for bly in self.layers:
            bly._init_respostnorm()

    @staticmethod
    def _init_weights(m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)

    def forward_features(self, x):
        x = self.pos_drop(self.patch_embed(x))
        for layer in self.layers:
            x = layer(x)
        return self.norm(x)

    def forward(self, x):
        return self.head(self.forward_features(x).mean(dim=1))
