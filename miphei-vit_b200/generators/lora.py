"""LoRA containers mirroring the reference's src/generators/lora.py (LoRALayer 8-18, QkvWithLoRA 21-33,
apply_lora 48-83): same attribute names (`qkv`, `lora_q`, `lora_v`, `A`, `B`), same init, same freezing rule.
The rank-8 updates are evaluated inside the QKV tensor-core GEMM through a K-extended weight (engine.py)."""
import math

import torch
import torch.nn as nn

from ..vit import VisionTransformer, _EngineOnly


class LoRALayer(_EngineOnly):
    def __init__(self, in_dim, out_dim, rank, alpha):
        super().__init__()
        self.A = nn.Parameter(torch.randn(in_dim, rank) / math.sqrt(rank))
        self.B = nn.Parameter(torch.zeros(rank, out_dim))
        self.alpha = alpha
        self.rank = rank


class QkvWithLoRA(_EngineOnly):
    def __init__(self, qkv, rank, alpha):
        super().__init__()
        self.qkv = qkv
        self.dim = qkv.in_features
        self.lora_q = LoRALayer(self.dim, self.dim, rank, alpha)
        self.lora_v = LoRALayer(self.dim, self.dim, rank, alpha)


def apply_lora(model, rank, alpha):
    """Wrap every block's qkv, freeze everything, un-freeze the LoRA matrices."""
    if not isinstance(model, VisionTransformer):
        raise NotImplementedError("LoRA is implemented for the ViT encoder only, got %s" % type(model))
    for block in model.blocks:
        block.attn.qkv = QkvWithLoRA(block.attn.qkv, rank=rank, alpha=alpha)
    for p in model.parameters():
        p.requires_grad = False
    for block in model.blocks:
        for p in block.attn.qkv.lora_q.parameters():
            p.requires_grad = True
        for p in block.attn.qkv.lora_v.parameters():
            p.requires_grad = True
