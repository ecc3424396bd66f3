"""Output-head containers mirroring src/generators/unet.py:407-438 (AttentionBlock, SegmentationHead) and the decoder
initialiser 522-531. Parameter names follow the reference (`psi.0`, `psi.1`, `psi.3`, and Sequential slots 0/1/2)."""
import torch.nn as nn

from ..vit import _EngineOnly


class AttentionBlock(_EngineOnly):
    def __init__(self, in_chns):
        super().__init__()
        self.psi = nn.Sequential(
            nn.Conv2d(in_chns, in_chns // 2, kernel_size=1, bias=True),
            nn.BatchNorm2d(in_chns // 2),
            nn.ReLU(),
            nn.Conv2d(in_chns // 2, 1, kernel_size=1, bias=True),
            nn.Sigmoid(),
        )


class SegmentationHead(nn.Sequential):
    """[attention gate, 3x3 conv to one channel, activation]; executed fused for all heads by the engine."""

    def __init__(self, in_channels, out_channels, kernel_size=3, activation=None, use_attention=False):
        if not use_attention or out_channels != 1 or kernel_size != 3 or not isinstance(activation, nn.Tanh):
            raise NotImplementedError("the B200 engine implements the MIPHEI-ViT head: gated, 3x3, one channel, Tanh")
        super().__init__(AttentionBlock(in_channels), nn.Conv2d(in_channels, out_channels, 3, padding=1), activation)
        initialize_decoder_head(self)

    def forward(self, x):  # pragma: no cover - guard
        raise RuntimeError("SegmentationHead is executed by the fused heads kernel; call the generator instead")


def initialize_decoder_head(module):
    for m in module.modules():
        if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
            nn.init.normal_(m.weight, 0.0, 0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.BatchNorm2d):
            nn.init.normal_(m.weight, 1.0, 0.02)
            nn.init.constant_(m.bias, 0)
