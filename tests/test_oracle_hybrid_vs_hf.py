"""Second opinion on the oracle's restatement of timm 0.6.12 ``vit_base_resnet50_384`` (oracle/timm_shim, the encoder of
dpt_hybrid_384): the same weights pushed through HuggingFace's independent implementations of the same published networks
  * ``transformers.BitModel`` with ``layer_type="bottleneck"`` (BiT / ResNetV2 with non-pre-activation bottlenecks, weight-
    standardised convs with TF "SAME" padding, GroupNorm: the backbone HF's own ViT-hybrid uses) for the ResNetV2-50 trunk,
  * ``transformers`` ViT encoder layers (pre-norm attention / MLP) for the 12 transformer blocks
must give the same features.  timm itself is not installed and not vendored."""
import pytest
import torch

import ref_env

ref_env.enable_shim()
import timm  # noqa: E402  (the shim)

transformers = pytest.importorskip("transformers")


def _seed(m, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "norm" in n and n.endswith("weight"):
                p.copy_(torch.rand(p.shape, generator=g) + 0.5)
            elif n.endswith("norm3.weight"):
                p.copy_((torch.rand(p.shape, generator=g) + 0.5) * 0.2)
            elif p.dim() >= 2:
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)
            else:
                p.copy_(torch.rand(p.shape, generator=g) * 0.2 - 0.1)
    return g


def test_shim_resnetv2_trunk_matches_hf_bit():
    m = timm.create_model("vit_base_resnet50_384", pretrained=False).eval()
    g = _seed(m, 3)
    bb = m.patch_embed.backbone
    cfg = transformers.BitConfig(global_padding="same", layer_type="bottleneck", depths=[3, 4, 9], hidden_sizes=[256, 512, 1024],
                                 embedding_size=64, embedding_dynamic_padding=True, num_groups=32, drop_path_rate=0.0)
    hf = transformers.BitModel(cfg).eval()
    sd = {}
    for k, v in bb.state_dict().items():
        k2 = k.replace("stem.conv.", "embedder.convolution.").replace("stem.norm.", "embedder.norm.")
        k2 = k2.replace("stages.", "encoder.stages.").replace(".blocks.", ".layers.")
        sd[k2] = v
    missing, unexpected = hf.load_state_dict(sd, strict=False)
    missing = [k for k in missing if not k.startswith("norm.")]          # HF's final norm is not part of the trunk's taps
    assert not missing and not unexpected, (missing, unexpected)
    x = torch.randn(1, 3, 128, 128, generator=g)
    with torch.no_grad():
        s0 = bb.stages[0](bb.stem(x))
        s1 = bb.stages[1](s0)
        s2 = bb.stages[2](s1)
        out = hf(pixel_values=x, output_hidden_states=True).hidden_states
    # hidden_states: embedder output, then one entry per stage
    for ours, theirs in zip((s0, s1, s2), out[1:4]):
        assert ours.shape == theirs.shape
        err = (ours - theirs).abs().max().item()
        assert err <= 2e-4 * max(1.0, theirs.abs().max().item()), err


def test_shim_vit_blocks_match_hf_vit_layers():
    m = timm.create_model("vit_base_resnet50_384", pretrained=False).eval()
    g = _seed(m, 4)
    cfg = transformers.ViTConfig(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072,
                                 hidden_act="gelu", hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
                                 layer_norm_eps=1e-6, qkv_bias=True)
    from transformers.models.vit.modeling_vit import ViTEncoder
    enc = ViTEncoder(cfg).eval()
    sd = {}
    for i, blk in enumerate(m.blocks):
        b = blk.state_dict()
        t = f"layer.{i}."
        wq, wk, wv = b["attn.qkv.weight"].split(768, dim=0)
        bq, bk, bv = b["attn.qkv.bias"].split(768, dim=0)
        sd.update({t + "attention.attention.query.weight": wq, t + "attention.attention.key.weight": wk,
                   t + "attention.attention.value.weight": wv, t + "attention.attention.query.bias": bq,
                   t + "attention.attention.key.bias": bk, t + "attention.attention.value.bias": bv,
                   t + "attention.output.dense.weight": b["attn.proj.weight"], t + "attention.output.dense.bias": b["attn.proj.bias"],
                   t + "layernorm_before.weight": b["norm1.weight"], t + "layernorm_before.bias": b["norm1.bias"],
                   t + "layernorm_after.weight": b["norm2.weight"], t + "layernorm_after.bias": b["norm2.bias"],
                   t + "intermediate.dense.weight": b["mlp.fc1.weight"], t + "intermediate.dense.bias": b["mlp.fc1.bias"],
                   t + "output.dense.weight": b["mlp.fc2.weight"], t + "output.dense.bias": b["mlp.fc2.bias"]})
    missing, unexpected = enc.load_state_dict(sd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    tok = torch.randn(2, 65, 768, generator=g)
    with torch.no_grad():
        a = tok
        for blk in m.blocks:
            a = blk(a)
        out = enc(tok)
        b = out.last_hidden_state if hasattr(out, "last_hidden_state") else out[0]
    err = (a - b).abs().max().item()
    assert err <= 2e-4 * max(1.0, b.abs().max().item()), err
