"""Parameter container with the module tree and parameter names of timm 1.0.15's `vit_giant_patch14_reg4_dinov2`
VisionTransformer (the model the reference builds at src/generators/foundation_models.py:53-57), so that reference
checkpoints load unchanged.  It holds weights only: the arithmetic runs in the CUDA engine (engine.py)."""
import math

import torch
import torch.nn as nn

PATCH = 14


class _EngineOnly(nn.Module):
    """Leaf containers are not executed one by one: the fused engine runs the whole path from their parameters."""

    def forward(self, *a, **k):  # pragma: no cover - guard
        raise RuntimeError(
            "%s is a parameter container of the B200 engine; call the generator (ViTMatte) / Encoder / Detail_Capture "
            "instead of individual layers" % type(self).__name__)


class PatchEmbed(_EngineOnly):
    def __init__(self, img_size, embed_dim):
        super().__init__()
        self.img_size = (img_size, img_size)
        self.patch_size = (PATCH, PATCH)
        self.grid_size = (img_size // PATCH, img_size // PATCH)
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = nn.Conv2d(3, embed_dim, PATCH, PATCH)

    def set_input_size(self, img_size):
        self.img_size = tuple(img_size)
        self.grid_size = (img_size[0] // PATCH, img_size[1] // PATCH)
        self.num_patches = self.grid_size[0] * self.grid_size[1]


class Attention(_EngineOnly):
    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.qkv = nn.Linear(dim, 3 * dim, bias=True)
        self.proj = nn.Linear(dim, dim, bias=True)


class LayerScale(_EngineOnly):
    def __init__(self, dim, init_values=1e-5):
        super().__init__()
        self.gamma = nn.Parameter(init_values * torch.ones(dim))


class GluMlp(_EngineOnly):
    """timm SwiGLUPacked: fc1 -> chunk(2) -> silu(x1) * x2 -> fc2."""

    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, 2 * hidden, bias=True)
        self.fc2 = nn.Linear(hidden, dim, bias=True)


class Block(_EngineOnly):
    def __init__(self, dim, num_heads, hidden, init_values):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = Attention(dim, num_heads)
        self.ls1 = LayerScale(dim, init_values)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = GluMlp(dim, hidden)
        self.ls2 = LayerScale(dim, init_values)


def _trunc_normal_(t, std):
    with torch.no_grad():
        return t.normal_(0.0, std).clamp_(-2.0, 2.0)


def resample_abs_pos_embed(posemb, new_size, num_prefix_tokens=1):
    """timm.layers.resample_abs_pos_embed semantics: bicubic + antialias in fp32 on the [h, w] grid, prefix kept."""
    n = posemb.shape[1] - num_prefix_tokens
    if n == new_size[0] * new_size[1]:
        return posemb
    old = int(math.sqrt(n))
    pre, grid = posemb[:, :num_prefix_tokens], posemb[:, num_prefix_tokens:]
    dt = grid.dtype
    grid = grid.float().reshape(1, old, old, -1).permute(0, 3, 1, 2)
    grid = torch.nn.functional.interpolate(grid, size=tuple(new_size), mode="bicubic", antialias=True)
    grid = grid.permute(0, 2, 3, 1).reshape(1, new_size[0] * new_size[1], -1).to(dt)
    return torch.cat([pre, grid], dim=1)


class VisionTransformer(_EngineOnly):
    """ViT-g/14 with 4 register tokens (H-Optimus-0 geometry by default)."""

    def __init__(self, img_size=224, embed_dim=1536, depth=40, num_heads=24, hidden=4096, init_values=1e-5,
                 reg_tokens=4):
        super().__init__()
        self.embed_dim = self.num_features = embed_dim
        self.num_prefix_tokens = 1 + reg_tokens
        self.num_reg_tokens = reg_tokens
        self.no_embed_class = True
        self.depth = depth
        self.num_heads = num_heads
        self.hidden = hidden
        self.patch_embed = PatchEmbed(img_size, embed_dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.reg_token = nn.Parameter(torch.zeros(1, reg_tokens, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches, embed_dim))
        self.blocks = nn.Sequential(*[Block(embed_dim, num_heads, hidden, init_values) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)
        self.init_weights()

    def init_weights(self):
        _trunc_normal_(self.pos_embed, 0.02)
        nn.init.normal_(self.cls_token, std=1e-6)
        nn.init.normal_(self.reg_token, std=1e-6)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                _trunc_normal_(m.weight, 0.02)
                nn.init.zeros_(m.bias)

    def set_input_size(self, img_size):
        img_size = tuple(img_size)
        new_grid = (img_size[0] // PATCH, img_size[1] // PATCH)
        if new_grid != self.patch_embed.grid_size:
            with torch.no_grad():
                pe = resample_abs_pos_embed(self.pos_embed.data, new_grid, num_prefix_tokens=0)
            self.pos_embed = nn.Parameter(pe, requires_grad=self.pos_embed.requires_grad)
        self.patch_embed.set_input_size(img_size)
