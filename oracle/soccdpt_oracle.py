"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the SOccDPT-V3 inference hot path.

A restatement of the reference's algorithm that can travel to the GPU box (where
/root/reference does not exist).  Only tests/, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this module;
the product package ``soccdpt_b200`` must never do so.

What follows what (paths under /root/reference/):
  * encoder ............ timm==0.6.12 SwinV2 (third party, not vendored; restated in
                         oracle/timm_shim) driven as SOccDPT/model/backbones/utils.py:64-81
                         with hooks [1,1,5,1] / [1,1,17,1] (SOccDPT/model/dpt.py:61-72)
  * decoder ............ SOccDPT/model/dpt.py:142-182, SOccDPT/model/blocks.py:391-414,466-497
  * depth head ......... SOccDPT/model/dpt.py:199-219,226-232
  * seg head ........... SOccDPT/model/SOccDPT.py:660-674, scaled_tanh.py:8-10
  * post-processing .... SOccDPT/model/SOccDPT.py:264-372 (resize in torch, the bit-exact
                         unproject/quirk/rotate/voxelise stage in oracle/voxel_oracle.c)

Pinning: ``tests/test_oracle_vs_reference.py`` runs this file against the UNMODIFIED
reference (imported through oracle/ref_env.py) in the build container and against
the committed fixtures ``tests/golden/*.npz`` (made by oracle/make_golden.py).
The encoder part is "parity unpinned" by the reference itself (no timm here).
"""
import ctypes
import math
import os
import subprocess

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# The synthetic pinhole camera every config uses (values from the
# reference's media/manydepth/intrinsics.json:1-5, frame size from
# SOccDPT/datasets/bengaluru_driving_dataset.py:118-121).
SYNTHETIC_CALIB = {
    "Camera.fx": 1250.6, "Camera.fy": 1254.8, "Camera.cx": 978.4, "Camera.cy": 562.1,
    "Camera.k1": 0.0, "Camera.k2": 0.0, "Camera.p1": 0.0, "Camera.p2": 0.0,
    "Camera.width": 1920, "Camera.height": 1080,
}

ENCODERS = {
    # model_type: (timm name, hooks, stage channels)
    "dpt_swin2_tiny_256": ("swinv2_tiny_window16_256", (1, 1, 5, 1), (96, 192, 384, 768)),
    "dpt_swin2_base_384": ("swinv2_base_window12to24_192to384_22kft1k", (1, 1, 17, 1), (128, 256, 512, 1024)),
    # ViT-B/16 + ResNetV2-50 stem; hooks: stages[0], stages[1], blocks[8], blocks[11] (dpt.py:86, vit.py:164-171)
    "dpt_hybrid_384": ("vit_base_resnet50_384", (0, 1, 8, 11), (256, 512, 768, 768)),
}


def build_lib():
    """Compile oracle/voxel_oracle.c (gcc) if the shared object is missing or stale."""
    so = os.path.join(HERE, "_build", "libvoxel_oracle.so")
    src = os.path.join(HERE, "voxel_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_lib())
        _LIB.soccdpt_oracle_voxelize.restype = ctypes.c_longlong
    return _LIB


class Geometry:
    """Constructor constants of the reference base class (SOccDPT.py:134-228)."""

    def __init__(self, calib=None, num_classes=3, grid_size=(256, 256, 32), scale=(2.0, 2.0, 0.666),
                 pc_scale=(10000.0, 50000.0, 800.0), pc_shift=(55.0, -20.0, 15.0),
                 correction_angle=(7.0, 0, 0)):
        calib = dict(SYNTHETIC_CALIB if calib is None else calib)
        self.fx, self.fy = calib["Camera.fx"], calib["Camera.fy"]
        self.cx, self.cy = calib["Camera.cx"], calib["Camera.cy"]
        self.width, self.height = int(calib["Camera.width"]), int(calib["Camera.height"])
        self.num_classes = num_classes
        self.grid_size = tuple(int(g) for g in grid_size)
        self.pc_scale, self.pc_shift = tuple(pc_scale), tuple(pc_shift)
        self.correction_angle = tuple(correction_angle)
        # SOccDPT.py:175-181
        self.occupancy_shape = np.array([float(grid_size[i] / scale[i]) for i in range(3)], dtype=np.float32)

    def rotation_matrices(self):
        """SOccDPT.py:74-111 on CPU tensors: fp32 deg2rad/cos/sin, three 3x3 fp32 matrices."""
        a, b, c = torch.tensor(self.correction_angle).to(dtype=torch.float32)
        a, b, c = torch.deg2rad(a), torch.deg2rad(b), torch.deg2rad(c)
        Ra = torch.tensor([[1, 0, 0], [0, torch.cos(a), -torch.sin(a)], [0, torch.sin(a), torch.cos(a)]])
        Rb = torch.tensor([[torch.cos(b), 0, torch.sin(b)], [0, 1, 0], [-torch.sin(b), 0, torch.cos(b)]])
        Rc = torch.tensor([[torch.cos(c), -torch.sin(c), 0], [torch.sin(c), torch.cos(c), 0], [0, 0, 1]])
        return torch.stack([Ra, Rb, Rc]).to(torch.float32).numpy().copy()


def voxelize(inv_depth_up, seg_up, geom, compute_occ=True, per_frame=False, threads=0):
    """Bit-exact stage. inv_depth_up (B,H,W) f32, seg_up (B,C,H,W) f32 (numpy or torch).
    Returns (inv_depth_clamped, points (B,H,W,3), grid (B,G0,G1,G2,C) | None) as numpy."""
    inv = np.ascontiguousarray(np.array(inv_depth_up, dtype=np.float32, copy=True))
    seg = np.ascontiguousarray(np.asarray(seg_up, dtype=np.float32))
    B, H, W = inv.shape
    C = seg.shape[1]
    assert seg.shape == (B, C, H, W)
    pts = np.empty((B, H, W, 3), np.float32)
    G = geom.grid_size
    grid = np.empty((B, G[0], G[1], G[2], C), np.float32) if compute_occ else None
    f = ctypes.c_float
    fp = ctypes.POINTER(ctypes.c_float)

    def P(a):
        return a.ctypes.data_as(fp) if a is not None else None

    pcs = np.array(geom.pc_scale, np.float32)
    pct = np.array(geom.pc_shift, np.float32)
    rot = np.ascontiguousarray(geom.rotation_matrices().reshape(-1))
    occ = np.ascontiguousarray(geom.occupancy_shape)
    gsz = np.array(G, np.int32)
    _lib().soccdpt_oracle_voxelize(
        P(inv), P(seg), B, H, W, C,
        f(np.float32(geom.fx)), f(np.float32(geom.fy)), f(np.float32(geom.cx)), f(np.float32(geom.cy)),
        P(pcs), P(pct), P(rot), P(occ), gsz.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
        P(pts), P(grid), int(per_frame), int(threads))
    return inv, pts, grid


def get_semantic_occupancy(inv_depth, segmentation, geom, compute_occ=True, per_frame=False, threads=0):
    """SOccDPT.py:264-372 (torch CPU tensors in, the reference's 4-tuple out, squeeze quirks kept)."""
    if inv_depth.dim() == 3:
        inv_depth = inv_depth.unsqueeze(1)
    inv_up = F.interpolate(inv_depth, size=(geom.height, geom.width), mode="bicubic", align_corners=False).squeeze()
    seg_up = F.interpolate(segmentation, size=(geom.height, geom.width), mode="nearest").squeeze()
    if inv_up.dim() == 2:
        inv_up = inv_up.unsqueeze(0)
    seg_b = seg_up.reshape(-1, geom.num_classes, geom.height, geom.width)
    inv_c, pts, grid = voxelize(inv_up.numpy(), seg_b.numpy(), geom, compute_occ, per_frame, threads)
    return (torch.from_numpy(inv_c), seg_up, torch.from_numpy(pts),
            torch.from_numpy(grid) if grid is not None else None)


# ----------------------------------------------------------------------------- network
def _conv(sd, name, x, padding):
    return F.conv2d(x, sd[name + ".weight"], sd.get(name + ".bias"), stride=1, padding=padding)


def _bn(sd, name, x):
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"], sd[name + ".bias"],
                        False, 0.1, 1e-5)


def _rcu(sd, p, x):
    """ResidualConvUnit_custom (blocks.py:391-414); bn1 / bn2 exist with use_bn (DPTSegmentationModel, dpt.py:240)."""
    out = _conv(sd, p + ".conv1", F.relu(x), 1)
    if p + ".bn1.weight" in sd:
        out = _bn(sd, p + ".bn1", out)
    out = _conv(sd, p + ".conv2", F.relu(out), 1)
    if p + ".bn2.weight" in sd:
        out = _bn(sd, p + ".bn2", out)
    return out + x


def _fusion(sd, p, xs, size):
    """FeatureFusionBlock_custom.forward (blocks.py:466-497)."""
    out = xs[0]
    if len(xs) == 2:
        out = out + _rcu(sd, p + ".resConfUnit1", xs[1])
    out = _rcu(sd, p + ".resConfUnit2", out)
    kw = {"scale_factor": 2} if size is None else {"size": size}
    out = F.interpolate(out, **kw, mode="bilinear", align_corners=True)
    return _conv(sd, p + ".out_conv", out, 0)


class OracleV3:
    """SOccDPT_V3 (SOccDPT.py:626-685) evaluated functionally from a reference-keyed state_dict."""

    PREFIX = "depth_net."       # the DPT whose encoder / decoder this instance evaluates

    def __init__(self, state_dict, model_type="dpt_swin2_tiny_256", sigmoid=True, geom=None,
                 compute_occ=True):
        import sys
        shim = os.path.join(HERE, "timm_shim")
        if shim not in sys.path:
            sys.path.insert(0, shim)
        import timm  # the shim
        name, self.hooks, self.chans = ENCODERS[model_type]
        self.sd = {k: v.detach().to(torch.float32).cpu() for k, v in state_dict.items()}
        self.encoder = timm.create_model(name, pretrained=False).eval()
        pfx = self.PREFIX + "pretrained.model."
        enc_sd = {k[len(pfx):]: v for k, v in self.sd.items() if k.startswith(pfx)}
        missing, unexpected = self.encoder.load_state_dict(enc_sd, strict=False)
        assert not unexpected and not missing, (missing, unexpected)
        self.sigmoid = sigmoid
        self.geom = geom if geom is not None else Geometry()
        self.compute_occ = compute_occ
        self.hybrid = model_type == "dpt_hybrid_384"

    @torch.no_grad()
    def _hybrid_taps(self, x):
        """forward_vit -> forward_adapted_unflatten(pretrained, x, "forward_flex") (vit.py:19-85, utils.py:84-133)
        with the post-processing of _make_vit_b_rn50_backbone (vit.py:179-219, readout = "project")."""
        m, sd, pp = self.encoder, self.sd, self.PREFIX + "pretrained.act_postprocess"
        b, c, h, w = x.shape
        gh, gw = h // 16, w // 16
        # _resize_pos_embed (vit.py:23-41)
        tok_pe, grid_pe = m.pos_embed[:, :1], m.pos_embed[0, 1:]
        gs = int(math.sqrt(len(grid_pe)))
        grid_pe = F.interpolate(grid_pe.reshape(1, gs, gs, -1).permute(0, 3, 1, 2), size=(gh, gw), mode="bilinear")
        pos = torch.cat([tok_pe, grid_pe.permute(0, 2, 3, 1).reshape(1, gh * gw, -1)], dim=1)
        bb = m.patch_embed.backbone
        s0 = bb.stages[0](bb.stem(x))          # hook "1": (B,256,h/4,w/4)
        s1 = bb.stages[1](s0)                  # hook "2": (B,512,h/8,w/8)
        s2 = bb.stages[2](s1)
        t = m.patch_embed.proj(s2).flatten(2).transpose(1, 2)
        t = torch.cat((m.cls_token.expand(b, -1, -1), t), dim=1) + pos
        hooked = {}
        for i, blk in enumerate(m.blocks):
            t = blk(t)
            if i in self.hooks[2:]:
                hooked[i] = t

        def readout(tk, idx):                  # ProjectReadout (utils.py:27-40) + Transpose + Unflatten + 1x1 conv
            feats = torch.cat((tk[:, 1:], tk[:, 0].unsqueeze(1).expand_as(tk[:, 1:])), -1)
            y = F.gelu(F.linear(feats, sd[f"{pp}{idx}.0.project.0.weight"], sd[f"{pp}{idx}.0.project.0.bias"]))
            y = y.transpose(1, 2).unflatten(2, (gh, gw))
            return F.conv2d(y, sd[f"{pp}{idx}.3.weight"], sd[f"{pp}{idx}.3.bias"])

        l3 = readout(hooked[self.hooks[2]], 3)
        l4 = F.conv2d(readout(hooked[self.hooks[3]], 4), sd[f"{pp}4.4.weight"], sd[f"{pp}4.4.bias"], stride=2, padding=1)
        return [s0, s1, l3, l4]

    @torch.no_grad()
    def encoder_taps(self, x):
        """forward_default (utils.py:64-81) + act_postprocess (swin_common.py:38-52): 4 NCHW maps."""
        if self.hybrid:
            return self._hybrid_taps(x)
        m = self.encoder
        t = m.pos_drop(m.patch_embed(x))
        taps = []
        for s, layer in enumerate(m.layers):
            for j, blk in enumerate(layer.blocks):
                t = blk(t)
                if j == self.hooks[s]:
                    h, w = layer.input_resolution
                    taps.append(t.transpose(1, 2).unflatten(2, (h, w)))
            t = layer.downsample(t)
        return taps

    @torch.no_grad()
    def decoder(self, taps):
        sd, s = self.sd, self.PREFIX + "scratch."
        l1, l2, l3, l4 = [_conv(sd, f"{s}layer{i + 1}_rn", taps[i], 1) for i in range(4)]
        p4 = _fusion(sd, s + "refinenet4", [l4], l3.shape[2:])
        p3 = _fusion(sd, s + "refinenet3", [p4, l3], l2.shape[2:])
        p2 = _fusion(sd, s + "refinenet2", [p3, l2], l1.shape[2:])
        return _fusion(sd, s + "refinenet1", [p2, l1], None)

    @torch.no_grad()
    def depth_head(self, path_1):
        """DPTDepthModel head (dpt.py:199-219) + squeeze (dpt.py:226-232)."""
        sd, h = self.sd, self.PREFIX + "scratch.output_conv."
        d = _conv(sd, h + "0", path_1, 1)
        d = F.interpolate(d, scale_factor=2, mode="bilinear", align_corners=True)
        d = F.relu(_conv(sd, h + "2", d, 1))
        return F.relu(_conv(sd, h + "4", d, 0)).squeeze(1)

    @torch.no_grad()
    def seg_head(self, path_1, h="seg_head."):
        """conv3x3 (no bias) -> BN -> ReLU -> conv1x1 -> x2 bilinear -> sigmoid | scaled tanh (SOccDPT.py:660-674, dpt.py:242-252)."""
        sd = self.sd
        g = _bn(sd, h + "1", _conv(sd, h + "0", path_1, 1))
        g = _conv(sd, h + "4", F.relu(g), 0)
        g = F.interpolate(g, scale_factor=2, mode="bilinear", align_corners=True)
        return torch.sigmoid(g) if self.sigmoid else 0.5 * torch.tanh(g) + 0.5

    @torch.no_grad()
    def seg_logits(self, path_1, h="seg_head."):
        """the head's output before Interpolate and the activation (SOccDPT.py:660-670): (B, C, h/2, w/2)."""
        sd = self.sd
        g = _bn(sd, h + "1", _conv(sd, h + "0", path_1, 1))
        return _conv(sd, h + "4", F.relu(g), 0)

    @torch.no_grad()
    def heads(self, path_1):
        return self.depth_head(path_1), self.seg_head(path_1)

    @torch.no_grad()
    def network(self, x):
        """image -> (inv_depth (B,h,w), seg (B,C,h,w), path_1, taps)"""
        taps = self.encoder_taps(x)
        path_1 = self.decoder(taps)
        d, g = self.heads(path_1)
        return d, g, path_1, taps

    @torch.no_grad()
    def __call__(self, x, threads=0):
        d, g, _, _ = self.network(x)
        return get_semantic_occupancy(d, g, self.geom, self.compute_occ, threads=threads)


class _OracleSegDPT(OracleV3):
    """the DPTSegmentationModel half of SOccDPT_V1: keys under ``seg_net.``, head = its own scratch.output_conv."""
    PREFIX = "seg_net."


class OracleV1:
    """SOccDPT_V1 (SOccDPT.py:470-523): depth_net(x) and seg_net(x) are two complete DPTs, then get_semantic_occupancy."""

    def __init__(self, state_dict, model_type="dpt_swin2_tiny_256", geom=None, compute_occ=True):
        self.depth = OracleV3(state_dict, model_type, True, geom, compute_occ)
        self.seg = _OracleSegDPT(state_dict, model_type, True, geom, compute_occ)
        self.geom, self.compute_occ = self.depth.geom, compute_occ

    @torch.no_grad()
    def network(self, x):
        """image -> (inv_depth (B,h,w), seg (B,C,h,w), depth path_1, seg path_1)"""
        p_d = self.depth.decoder(self.depth.encoder_taps(x))
        p_s = self.seg.decoder(self.seg.encoder_taps(x))
        return self.depth.depth_head(p_d), self.seg.seg_head(p_s, "seg_net.scratch.output_conv."), p_d, p_s

    @torch.no_grad()
    def __call__(self, x, threads=0):
        d, g, _, _ = self.network(x)
        return get_semantic_occupancy(d, g, self.geom, self.compute_occ, threads=threads)


def config5_maps(B, H, W, C=3, seed=0, scaled_tanh=False):
    """The synthetic voxeliser inputs of SURVEY.md 8(d) config 5."""
    g = torch.Generator().manual_seed(seed)
    inv = torch.rand(B, H, W, generator=g) * 0.2 + 0.02
    bad = torch.rand(B, H, W, generator=g) < 0.01
    kind = torch.randint(0, 4, (B, H, W), generator=g)
    vals = torch.tensor([0.0, -1.0, float("nan"), float("inf")])[kind]
    inv = torch.where(bad, vals, inv)
    seg = torch.sigmoid(torch.randn(B, C, H, W, generator=g))
    if scaled_tanh:
        seg = torch.where(torch.rand(B, C, H, W, generator=g) < 0.3, torch.zeros(()), seg)
    return inv.contiguous(), seg.contiguous()


def occupancy_indices(grid):
    """Sorted (b,i,j,k,c) index list of a dense grid (for index-level comparisons)."""
    g = torch.as_tensor(grid)
    return g.nonzero()


def occupancy_grid_to_points(occupancy_grid, grid_size=(256, 256, 32), scale=(2.0, 2.0, 0.666)):
    """Restatement of SOccDPT/utils/__init__.py:532-568: cells >= 0.5 as float64 rows (x, y, z, class); per class the
    np.argwhere order, coordinates = float32(index / grid_size * float32(grid_size / scale))."""
    g = np.asarray(occupancy_grid)
    assert g.ndim == 4
    occ = np.array([float(grid_size[i] / scale[i]) for i in range(3)], dtype=np.float32)
    idx = np.argwhere(g >= 0.5)
    rows = []
    for c in range(g.shape[3]):
        ijk = idx[idx[:, 3] == c][:, :3]
        xyz = (ijk / np.array(grid_size[:3]) * occ).astype(np.float32)
        rows.append(np.concatenate([xyz, np.full((xyz.shape[0], 1), c)], axis=1))
    return np.concatenate(rows, axis=0)
