"""Second, independent pin of the ViT restatement: transformers' Dinov2WithRegistersModel (same architecture family as
timm's vit_giant_patch14_reg4_dinov2) with weights mapped through the table the reference itself ships
(src/generators/foundation_models.py:230-318)."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import model as om  # noqa: E402

transformers = pytest.importorskip("transformers")


def _hf_from_oracle(sd, cfg):
    from transformers import Dinov2WithRegistersConfig, Dinov2WithRegistersModel

    D = cfg.embed_dim
    assert cfg.hidden == (int(D * 4) * 2 // 3 + 7) // 8 * 8  # HF Dinov2SwiGLUFFN hidden width
    hc = Dinov2WithRegistersConfig(
        hidden_size=D, num_hidden_layers=cfg.depth, num_attention_heads=cfg.num_heads, mlp_ratio=4,
        image_size=cfg.img_size, patch_size=14, num_register_tokens=4, use_swiglu_ffn=True,
        layerscale_value=1e-5, layer_norm_eps=1e-6, qkv_bias=True)
    hf = Dinov2WithRegistersModel(hc).eval()
    hsd = hf.state_dict()
    v = "encoder.vit."
    new = {}
    new["embeddings.cls_token"] = sd[v + "cls_token"]
    new["embeddings.register_tokens"] = sd[v + "reg_token"]
    new["embeddings.mask_token"] = hsd["embeddings.mask_token"]
    pos = torch.cat([torch.zeros(1, 1, D), sd[v + "pos_embed"]], dim=1)  # timm no_embed_class: cls gets no pos
    new["embeddings.position_embeddings"] = pos
    new["embeddings.patch_embeddings.projection.weight"] = sd[v + "patch_embed.proj.weight"]
    new["embeddings.patch_embeddings.projection.bias"] = sd[v + "patch_embed.proj.bias"]
    for i in range(cfg.depth):
        b = v + "blocks.%d." % i
        h = "encoder.layer.%d." % i
        new[h + "norm1.weight"] = sd[b + "norm1.weight"]
        new[h + "norm1.bias"] = sd[b + "norm1.bias"]
        w, bias = sd[b + "attn.qkv.qkv.weight"], sd[b + "attn.qkv.qkv.bias"]
        for j, nm in enumerate(("query", "key", "value")):
            new[h + "attention.attention.%s.weight" % nm] = w[j * D:(j + 1) * D]
            new[h + "attention.attention.%s.bias" % nm] = bias[j * D:(j + 1) * D]
        new[h + "attention.output.dense.weight"] = sd[b + "attn.proj.weight"]
        new[h + "attention.output.dense.bias"] = sd[b + "attn.proj.bias"]
        new[h + "layer_scale1.lambda1"] = sd[b + "ls1.gamma"]
        new[h + "norm2.weight"] = sd[b + "norm2.weight"]
        new[h + "norm2.bias"] = sd[b + "norm2.bias"]
        new[h + "mlp.weights_in.weight"] = sd[b + "mlp.fc1.weight"]
        new[h + "mlp.weights_in.bias"] = sd[b + "mlp.fc1.bias"]
        new[h + "mlp.weights_out.weight"] = sd[b + "mlp.fc2.weight"]
        new[h + "mlp.weights_out.bias"] = sd[b + "mlp.fc2.bias"]
        new[h + "layer_scale2.lambda1"] = sd[b + "ls2.gamma"]
    new["layernorm.weight"] = sd[v + "norm.weight"]
    new["layernorm.bias"] = sd[v + "norm.bias"]
    missing = set(hsd) - set(new)
    assert not missing, missing
    for k in new:
        assert tuple(new[k].shape) == tuple(hsd[k].shape), (k, new[k].shape, hsd[k].shape)
    hf.load_state_dict(new, strict=True)
    return hf


def test_vit_restatement_matches_hf_dinov2_with_registers():
    # HF SwiGLU hidden = (int(D*4) * 2 // 3 + 7) // 8 * 8 ; D=96 -> 256
    cfg = om.Config(img_size=112, embed_dim=96, depth=3, num_heads=2, hidden=256, out_chans=1)
    sd = om.init_state_dict(cfg, seed=7, perturb=True)
    # LoRA must be neutral for this comparison (HF has no LoRA)
    for k in sd:
        if k.endswith(("lora_q.B", "lora_v.B")):
            sd[k] = torch.zeros_like(sd[k])
    hf = _hf_from_oracle(sd, cfg)
    x = om.normalize_tiles(om.synthetic_tiles_u8(2, cfg.img_size, seed=5))
    with torch.no_grad():
        want = hf(pixel_values=x).last_hidden_state
        got = om.vit_forward(sd, x, cfg)
    assert got.shape == want.shape == (2, 64 + 5, 96)
    assert (got - want).abs().max().item() < 1e-4, (got - want).abs().max().item()
    assert want.abs().mean().item() > 0.1
