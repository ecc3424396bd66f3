"""MIPHEI-ViT generator (ViTMatte-style) — drop-in for src/generators/mipheivit.py of the reference.

Same classes, constructor arguments, attribute surface and state-dict keys as the reference (Basic_Conv3x3 20-41,
ConvStream 44-73, Fusion_Block 76-93, ViTMatte 96-121, Encoder 124-163, Detail_Capture 166-220, get_vitmatte 224-233);
the modules hold the parameters, the arithmetic runs in hand-written sm_100a kernels through engine.MipheiEngine.
"""
import torch
import torch.nn as nn

from ..engine import MipheiEngine
from ..vit import VisionTransformer, _EngineOnly
from .foundation_models import FOUNDATION_MODEL_REGISTRY
from .lora import apply_lora
from .unet import SegmentationHead, initialize_decoder_head


class Basic_Conv3x3(_EngineOnly):
    """Conv3x3 (no bias) + BatchNorm2d + ReLU; stride 2 in ConvStream, 1 in the fusion blocks."""

    def __init__(self, in_chans, out_chans, stride=2, padding=1):
        super().__init__()
        self.conv = nn.Conv2d(in_chans, out_chans, 3, stride, padding, bias=False)
        self.bn = nn.BatchNorm2d(out_chans)
        self.relu = nn.ReLU(inplace=False)


class ConvStream(_EngineOnly):
    def __init__(self, in_chans=4, out_chans=[48, 96, 192]):
        super().__init__()
        self.convs = nn.ModuleList()
        self.conv_chans = [in_chans] + list(out_chans)
        for i in range(len(self.conv_chans) - 1):
            self.convs.append(Basic_Conv3x3(self.conv_chans[i], self.conv_chans[i + 1]))


class Fusion_Block(_EngineOnly):
    def __init__(self, in_chans, out_chans):
        super().__init__()
        self.conv = Basic_Conv3x3(in_chans, out_chans, stride=1, padding=1)


class Detail_Capture(nn.Module):
    def __init__(self, emb_chans, in_chans=3, out_chans=1, convstream_out=[48, 96, 192],
                 fusion_out=[256, 128, 64, 32], use_attention=True, activation=torch.nn.Identity()):
        super().__init__()
        assert len(fusion_out) == len(convstream_out) + 1
        if in_chans != 3 or list(convstream_out) != [48, 96, 192] or list(fusion_out) != [256, 128, 64, 32]:
            raise NotImplementedError("the B200 engine implements the MIPHEI-ViT decoder geometry only")
        self.convstream = ConvStream(in_chans=in_chans, out_chans=list(convstream_out))
        self.conv_chans = self.convstream.conv_chans
        self.num_heads = out_chans
        self.fusion_blks = nn.ModuleList()
        self.fus_channs = [emb_chans] + list(fusion_out)
        for i in range(len(self.fus_channs) - 1):
            self.fusion_blks.append(Fusion_Block(in_chans=self.fus_channs[i] + self.conv_chans[-(i + 1)],
                                                 out_chans=self.fus_channs[i + 1]))
        for idx in range(self.num_heads):
            setattr(self, "segmentation_head_%d" % idx,
                    SegmentationHead(in_channels=fusion_out[-1], out_channels=1, activation=activation, kernel_size=3,
                                     use_attention=use_attention))
        self._engine_ref = None

    def forward(self, features, images):
        """features [B, D, S/16, S/16] (NCHW, as Encoder.forward returns), images [B,3,S,S] -> [B, heads, S, S]."""
        if self._engine_ref is None:
            raise RuntimeError("Detail_Capture must be part of a ViTMatte generator to run")
        return self._engine_ref().decode(features, images)


class Encoder(nn.Module):
    def __init__(self, vit):
        super().__init__()
        if not isinstance(vit, VisionTransformer):
            raise ValueError("Model should be the B200 VisionTransformer container, got %s" % type(vit))
        self.vit = vit
        self.is_swint = False
        self.grid_size = self.vit.patch_embed.grid_size
        self.num_prefix_tokens = self.vit.num_prefix_tokens
        self.embed_dim = self.vit.embed_dim
        img_size = self.vit.patch_embed.img_size
        assert img_size[0] % 16 == 0
        assert img_size[1] % 16 == 0
        tgt = (img_size[0] / 16, img_size[1] / 16)
        self.scale_factor = (tgt[0] / self.grid_size[0], tgt[1] / self.grid_size[1])
        self._engine_ref = None

    def forward(self, x):
        """x [B,3,S,S] -> features [B, D, S/16, S/16] (channel-last memory, like the reference's strided view)."""
        if self._engine_ref is None:
            raise RuntimeError("Encoder must be part of a ViTMatte generator to run")
        return self._engine_ref().encode(x)


class ViTMatte(nn.Module):
    def __init__(self, encoder, decoder):
        super().__init__()
        self.encoder = encoder
        self.decoder = decoder
        self.initialize()
        object.__setattr__(self, "_engine", MipheiEngine(self))
        import weakref
        ref = weakref.ref(self._engine)
        encoder._engine_ref = ref
        decoder._engine_ref = ref

    @property
    def engine(self):
        return self._engine

    def forward(self, x):
        return self._engine.forward(x)

    def initialize(self):
        initialize_decoder_head(self.decoder)

    def set_input_size(self, img_size):
        if any((s & (s - 1)) != 0 or s == 0 for s in img_size):
            raise ValueError("Both height and width in img_size must be powers of 2")
        if any(s < 128 for s in img_size):
            raise ValueError("Height and width must be greater or equal to 128")
        if img_size[0] != img_size[1]:
            raise NotImplementedError("square tiles only")
        self.encoder.vit.set_input_size(img_size=img_size)
        self.encoder.grid_size = self.encoder.vit.patch_embed.grid_size
        tgt = (img_size[0] / 16, img_size[1] / 16)
        self.encoder.scale_factor = (tgt[0] / self.encoder.grid_size[0], tgt[1] / self.encoder.grid_size[1])
        self._engine.invalidate()

    # weights changed behind the engine's back -> re-pack lazily
    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        if "_engine" in self.__dict__:
            self._engine.invalidate()
        return out

    def load_state_dict(self, *a, **k):
        out = super().load_state_dict(*a, **k)
        self._engine.invalidate()
        return out


def get_vitmatte(encoder_name, img_size, num_classes, use_lora=False, ckpt_path=None, drop_path_rate=0, **geometry):
    vit = FOUNDATION_MODEL_REGISTRY[encoder_name](
        img_size, ckpt_path=ckpt_path, drop_path_rate=drop_path_rate, global_pool="", **geometry)
    if not use_lora:
        raise NotImplementedError("the reference always builds myvitmatte with use_lora=True (generators/__init__.py:42-45)")
    apply_lora(vit, rank=8, alpha=1.)
    encoder = Encoder(vit)
    decoder = Detail_Capture(emb_chans=encoder.embed_dim, out_chans=num_classes, use_attention=True, activation=nn.Tanh())
    return ViTMatte(encoder=encoder, decoder=decoder)
