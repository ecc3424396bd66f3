"""Model factory — drop-in for src/generators/__init__.py:9-56 of the reference (the `myvitmatte` branch)."""
from .mipheivit import get_vitmatte


def get_generator(model_name, img_size, nc_in, nc_out, cfg):
    if model_name.startswith("myvitmatte"):
        if nc_in != 3:
            raise NotImplementedError("MIPHEI-ViT takes 3-channel H&E tiles")
        ckpt_path = cfg.model.encoder.encoder_weights
        geometry = {}
        enc = cfg.model.encoder
        test_geometry = enc.get("test_geometry", None) if hasattr(enc, "get") else getattr(enc, "test_geometry", None)
        if test_geometry:
            geometry = dict(test_geometry)
        return get_vitmatte(cfg.model.encoder.encoder_name, img_size, nc_out, use_lora=True, ckpt_path=ckpt_path,
                            **geometry)
    if model_name.startswith(("smp_unet", "unet", "hemit")):
        raise NotImplementedError(
            "%s is one of the reference's comparison baselines and is out of scope of the B200 hot path "
            "(SURVEY.md section 2, rows 6-7)" % model_name)
    raise NotImplementedError(model_name)
