"""Second opinion on the oracle's SwinV2 restatement (oracle/timm_shim): the same
weights pushed through HuggingFace ``transformers.Swinv2Model`` (an independent
implementation of the same published algorithm) must give the same features.
timm 0.6.12 itself is not installed and not vendored, so this is the strongest
pin available for the encoder ("parity unpinned")."""
import pytest
import torch

import ref_env

ref_env.enable_shim()
import timm  # noqa: E402  (the shim)


def _to_hf_state(sd, depths):
    out = {}
    out["embeddings.patch_embeddings.projection.weight"] = sd["patch_embed.proj.weight"]
    out["embeddings.patch_embeddings.projection.bias"] = sd["patch_embed.proj.bias"]
    out["embeddings.norm.weight"] = sd["patch_embed.norm.weight"]
    out["embeddings.norm.bias"] = sd["patch_embed.norm.bias"]
    out["layernorm.weight"] = sd["norm.weight"]
    out["layernorm.bias"] = sd["norm.bias"]
    for i, d in enumerate(depths):
        for j in range(d):
            s = f"layers.{i}.blocks.{j}."
            t = f"encoder.layers.{i}.blocks.{j}."
            C = sd[s + "attn.qkv.weight"].shape[1]
            wq, wk, wv = sd[s + "attn.qkv.weight"].split(C, dim=0)
            out[t + "attention.self.query.weight"] = wq
            out[t + "attention.self.key.weight"] = wk
            out[t + "attention.self.value.weight"] = wv
            out[t + "attention.self.query.bias"] = sd[s + "attn.q_bias"]
            out[t + "attention.self.value.bias"] = sd[s + "attn.v_bias"]
            out[t + "attention.self.logit_scale"] = sd[s + "attn.logit_scale"]
            out[t + "attention.self.continuous_position_bias_mlp.0.weight"] = sd[s + "attn.cpb_mlp.0.weight"]
            out[t + "attention.self.continuous_position_bias_mlp.0.bias"] = sd[s + "attn.cpb_mlp.0.bias"]
            out[t + "attention.self.continuous_position_bias_mlp.2.weight"] = sd[s + "attn.cpb_mlp.2.weight"]
            out[t + "attention.output.dense.weight"] = sd[s + "attn.proj.weight"]
            out[t + "attention.output.dense.bias"] = sd[s + "attn.proj.bias"]
            out[t + "layernorm_before.weight"] = sd[s + "norm1.weight"]
            out[t + "layernorm_before.bias"] = sd[s + "norm1.bias"]
            out[t + "layernorm_after.weight"] = sd[s + "norm2.weight"]
            out[t + "layernorm_after.bias"] = sd[s + "norm2.bias"]
            out[t + "intermediate.dense.weight"] = sd[s + "mlp.fc1.weight"]
            out[t + "intermediate.dense.bias"] = sd[s + "mlp.fc1.bias"]
            out[t + "output.dense.weight"] = sd[s + "mlp.fc2.weight"]
            out[t + "output.dense.bias"] = sd[s + "mlp.fc2.bias"]
        if i < len(depths) - 1:
            out[f"encoder.layers.{i}.downsample.reduction.weight"] = sd[f"layers.{i}.downsample.reduction.weight"]
            out[f"encoder.layers.{i}.downsample.norm.weight"] = sd[f"layers.{i}.downsample.norm.weight"]
            out[f"encoder.layers.{i}.downsample.norm.bias"] = sd[f"layers.{i}.downsample.norm.bias"]
    return out


@pytest.mark.parametrize("name,img,ws,dim,depths,heads,pws", [
    ("swinv2_tiny_window16_256", 256, 16, 96, (2, 2, 6, 2), (3, 6, 12, 24), (0, 0, 0, 0)),
    # dpt_swin2_base_384: windows 24 / 24 / 24 / 12 with PRETRAINED windows 12 / 12 / 12 / 6 (the log-spaced coordinate table
    # of the continuous position bias is normalised by the pretrained window: the branch the tiny model never takes)
    ("swinv2_base_window12to24_192to384_22kft1k", 384, 24, 128, (2, 2, 18, 2), (4, 8, 16, 32), (12, 12, 12, 6)),
])
def test_shim_swinv2_matches_hf(name, img, ws, dim, depths, heads, pws):
    transformers = pytest.importorskip("transformers")
    torch.manual_seed(0)
    m = timm.create_model(name, pretrained=False).eval()
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for n, p in m.named_parameters():  # default init makes every block an identity
            if "norm" in n and n.endswith("weight"):
                p.copy_(torch.rand(p.shape, generator=g) + 0.5)
            elif n.endswith("logit_scale"):
                p.copy_(torch.rand(p.shape, generator=g) * 1.4 + 1.6)
            elif p.dim() >= 2:
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)
            else:
                p.copy_(torch.rand(p.shape, generator=g) * 0.2 - 0.1)
    cfg = transformers.Swinv2Config(
        image_size=img, patch_size=4, num_channels=3, embed_dim=dim, depths=list(depths),
        num_heads=list(heads), window_size=ws, pretrained_window_sizes=list(pws), mlp_ratio=4.0,
        qkv_bias=True, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, drop_path_rate=0.0,
        hidden_act="gelu", layer_norm_eps=1e-5)
    hf = transformers.Swinv2Model(cfg, add_pooling_layer=False).eval()
    missing, unexpected = hf.load_state_dict(_to_hf_state(m.state_dict(), depths), strict=False)
    missing = [k for k in missing if "relative_" not in k and "key.bias" not in k]
    assert not missing and not unexpected, (missing, unexpected)
    x = torch.randn(2 if img <= 256 else 1, 3, img, img, generator=g)
    with torch.no_grad():
        a = m.forward_features(x)
        b = hf(pixel_values=x).last_hidden_state
    err = (a - b).abs().max().item()
    assert err <= 2e-4 * max(1.0, b.abs().max().item()), err
