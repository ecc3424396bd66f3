"""Imports the reference's OWN hot-path modules from /root/reference for oracle validation and golden generation.

TEST INFRASTRUCTURE ONLY. The reference needs timm / segmentation_models_pytorch, neither installable here, so
those modules are stubbed: the decoder, heads, LoRA wrapper and losses then import and run
verbatim, wrapped around a ViT that subclasses the stub `VisionTransformer` and evaluates oracle.model.vit_forward's
arithmetic through timm's parameter names.  /root/reference does not exist on the GPU box: nothing under tests -m gpu,
smoke() or bench.py reaches this file.
"""
import importlib
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

REF_ROOT = os.environ.get("MIPHEI_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "src", "generators"))


class _VisionTransformer(nn.Module):
    pass


class _SwinTransformer(nn.Module):
    pass


class _ResNet(nn.Module):
    pass


def _resample_abs_pos_embed(posemb, new_size, num_prefix_tokens=1, **kw):
    n = posemb.shape[1] - num_prefix_tokens
    if n == new_size[0] * new_size[1]:
        return posemb
    old = int(n ** 0.5)
    pre, grid = posemb[:, :num_prefix_tokens], posemb[:, num_prefix_tokens:]
    grid = grid.reshape(1, old, old, -1).permute(0, 3, 1, 2).float()
    grid = F.interpolate(grid, size=new_size, mode="bicubic", antialias=True)
    grid = grid.permute(0, 2, 3, 1).reshape(1, new_size[0] * new_size[1], -1)
    return torch.cat([pre, grid], dim=1)


_mods = None


def load():
    """Returns dict(mipheivit=..., lora=..., unet=..., loss=...) of verbatim reference modules."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)

    def stub(name, **kw):
        m = types.ModuleType(name)
        m.__dict__.update(kw)
        sys.modules.setdefault(name, m)
        return sys.modules[name]

    stub("timm")
    stub("timm.layers", resample_abs_pos_embed=_resample_abs_pos_embed)
    stub("timm.layers.helpers", to_2tuple=lambda x: (x, x))
    stub("timm.models", VisionTransformer=_VisionTransformer, SwinTransformer=_SwinTransformer, ResNet=_ResNet,
         load_state_dict_from_hf=None, parse_model_name=None)
    stub("segmentation_models_pytorch")
    pkg = types.ModuleType("refgen")
    pkg.__path__ = [os.path.join(REF_ROOT, "src", "generators")]
    sys.modules["refgen"] = pkg
    out = {}
    out["mipheivit"] = importlib.import_module("refgen.mipheivit")
    out["lora"] = importlib.import_module("refgen.lora")
    out["unet"] = importlib.import_module("refgen.unet")
    spec = importlib.util.spec_from_file_location("refloss", os.path.join(REF_ROOT, "src", "loss.py"))
    loss = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(loss)
    out["loss"] = loss
    _mods = out
    return out


# ---------------------------------------------------------------------------------------------------------------
# A timm-shaped ViT (module tree and parameter names of timm 1.0.15 VisionTransformer) so that the reference's
# apply_lora / Encoder / ViTMatte wrap it unchanged.
class _PatchEmbed(nn.Module):
    def __init__(self, img_size, dim):
        super().__init__()
        self.img_size = (img_size, img_size)
        self.patch_size = (14, 14)
        self.grid_size = (img_size // 14, img_size // 14)
        self.proj = nn.Conv2d(3, dim, 14, 14)

    def forward(self, x):
        assert x.shape[-2:] == self.img_size
        return self.proj(x).flatten(2).transpose(1, 2)


class _Attention(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.num_heads = heads
        self.qkv = nn.Linear(dim, 3 * dim)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        x = F.scaled_dot_product_attention(q, k, v)
        return self.proj(x.transpose(1, 2).reshape(B, N, C))


class _LayerScale(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.gamma = nn.Parameter(torch.full((dim,), 1e-5))

    def forward(self, x):
        return x * self.gamma


class _GluMlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, 2 * hidden)
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        x1, x2 = self.fc1(x).chunk(2, dim=-1)
        return self.fc2(F.silu(x1) * x2)


class _Block(nn.Module):
    def __init__(self, dim, heads, hidden):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _Attention(dim, heads)
        self.ls1 = _LayerScale(dim)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _GluMlp(dim, hidden)
        self.ls2 = _LayerScale(dim)

    def forward(self, x):
        x = x + self.ls1(self.attn(self.norm1(x)))
        return x + self.ls2(self.mlp(self.norm2(x)))


class TimmLikeViT(_VisionTransformer):
    def __init__(self, img_size, embed_dim, depth, num_heads, hidden):
        super().__init__()
        self.embed_dim = embed_dim
        self.num_prefix_tokens = 5
        self.no_embed_class = True
        self.patch_embed = _PatchEmbed(img_size, embed_dim)
        g = img_size // 14
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.reg_token = nn.Parameter(torch.zeros(1, 4, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, g * g, embed_dim))
        self.blocks = nn.Sequential(*[_Block(embed_dim, num_heads, hidden) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)

    def forward(self, x):
        x = self.patch_embed(x) + self.pos_embed
        B = x.shape[0]
        x = torch.cat([self.cls_token.expand(B, -1, -1), self.reg_token.expand(B, -1, -1), x], dim=1)
        return self.norm(self.blocks(x))


def build_reference_model(cfg, state_dict=None):
    """get_vitmatte (mipheivit.py:224-233) with the timm model replaced by TimmLikeViT; reference code otherwise."""
    m = load()
    vit = TimmLikeViT(cfg.img_size, cfg.embed_dim, cfg.depth, cfg.num_heads, cfg.hidden)
    m["lora"].apply_lora(vit, rank=8, alpha=1.)
    enc = m["mipheivit"].Encoder(vit)
    dec = m["mipheivit"].Detail_Capture(emb_chans=enc.embed_dim, out_chans=cfg.out_chans, use_attention=True,
                                        activation=nn.Tanh())
    model = m["mipheivit"].ViTMatte(encoder=enc, decoder=dec)
    if state_dict is not None:
        model.load_state_dict(state_dict, strict=True)
    return model
