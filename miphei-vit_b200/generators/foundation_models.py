"""Foundation-model registry: only H-Optimus-0 (`hoptimus0`, src/generators/foundation_models.py:50-69) is on the
accelerated path; the reference's other backbones are paper ablations and out of scope."""
import torch

from ..vit import VisionTransformer, resample_abs_pos_embed


def resize_pos_embed_statedict(state_dict, model, img_size):
    """src/generators/foundation_models.py:198-208."""
    if img_size != 224:
        state_dict["pos_embed"] = resample_abs_pos_embed(
            state_dict["pos_embed"], new_size=model.patch_embed.grid_size,
            num_prefix_tokens=0 if model.no_embed_class else model.num_prefix_tokens)
    return state_dict


def hoptimus0(img_size, pretrained=True, ckpt_path=None, drop_path_rate=0., global_pool="", **geometry):
    """ViT-g/14-reg4 at `img_size`. `ckpt_path` is a timm-format state_dict (.bin); without one the weights stay at
    their random initialisation (there is no hub access here, unlike the reference's HF download branch).
    `geometry` (embed_dim, depth, num_heads, hidden) exists for reduced-size test models only."""
    if drop_path_rate not in (0, 0.0):
        raise NotImplementedError("drop_path_rate must be 0 (the reference passes 0, mipheivit.py:224-226)")
    if global_pool != "":
        raise NotImplementedError('global_pool must be "" (token sequence output)')
    model = VisionTransformer(img_size=img_size, init_values=1e-5, **geometry)
    if ckpt_path:
        state_dict = torch.load(ckpt_path, map_location="cpu")
        state_dict = resize_pos_embed_statedict(state_dict, model, img_size)
        model.load_state_dict(state_dict)
    elif pretrained:
        print("Warning: no encoder_weights given and no hub access: random initialization for hoptimus0")
    return model


FOUNDATION_MODEL_REGISTRY = {"hoptimus0": hoptimus0}
