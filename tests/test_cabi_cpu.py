"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports exactly what include/miphei_b200.h
declares, the Python binding covers it, the module tree / state-dict layout matches the oracle's (== the reference's),
and nothing computes without a GPU."""
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import model as om  # noqa: E402


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "miphei_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(mv_[a-z0-9_]+)\s*\(", src))


@pytest.fixture(scope="module")
def built():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    ge.build()
    from miphei_vit_b200 import lib
    return lib


def test_library_exports_every_declared_symbol(built):
    lib = built
    declared = _header_symbols()
    assert declared, "no symbols parsed from the header"
    bound = set(lib.exported_symbols())
    assert declared == bound, (declared - bound, bound - declared)
    out = subprocess.run(["nm", "-D", "--defined-only", lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (mv_[a-z0-9_]+)", out))
    assert declared <= exported, declared - exported
    handle = lib.load()
    for s in declared:
        assert hasattr(handle, s)


def test_library_targets_sm100a_tensor_core_path(built):
    out = subprocess.run(["cuobjdump", "-sass", built.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out or "SM100a" in out or "sm_100" in out
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in out, mnemonic + " missing: tcgen05 / TMA path not compiled in"


def test_no_compute_without_gpu(built):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(built.MipheiB200Error):
        built.init(0)
    from miphei_vit_b200 import ops
    with pytest.raises(built.MipheiB200Error):
        ops.layernorm_fwd(torch.zeros(4, 128), torch.ones(128), torch.zeros(128))


def test_generator_factory_and_state_dict_layout():
    from types import SimpleNamespace as NS
    from miphei_vit_b200.generators import get_generator

    geo = dict(embed_dim=128, depth=2, num_heads=2, hidden=256)
    cfg = NS(model=NS(model_name="myvitmatte", encoder=NS(encoder_name="hoptimus0", encoder_weights=None, test_geometry=geo)),
             train=NS(foreground_head=False))
    g = get_generator("myvitmatte", 128, 3, 16, cfg)
    ocfg = om.Config(img_size=128, out_chans=16, **geo)
    sd = om.init_state_dict(ocfg, seed=1)
    assert set(g.state_dict().keys()) == set(sd.keys())
    g.load_state_dict(sd, strict=True)
    # attribute surface probed by the reference's callers
    assert not hasattr(g, "foreground_head") and not hasattr(g, "swinT")
    assert hasattr(g, "encoder") and hasattr(g.encoder, "vit") and hasattr(g, "decoder")
    assert g.encoder.vit.patch_embed.grid_size == (9, 9) and g.encoder.vit.num_prefix_tokens == 5
    assert g.encoder.embed_dim == 128 and g.encoder.scale_factor == (8 / 9, 8 / 9)
    train = {n for n, p in g.named_parameters() if p.requires_grad}
    assert train == set(om.trainable_keys(sd))
    with pytest.raises(ValueError):
        g.set_input_size((200, 200))
    with pytest.raises(ValueError):
        g.set_input_size((64, 64))
    g.set_input_size((256, 256))
    assert g.encoder.vit.pos_embed.shape == (1, 18 * 18, 128) and g.encoder.grid_size == (18, 18)
    with pytest.raises(NotImplementedError):
        get_generator("smp_unet_convnext", 128, 3, 16, cfg)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            g(torch.zeros(1, 3, 256, 256))


def test_safetensors_style_partial_load_contract():
    """inference.py:135-153: LoRA + decoder keys only, strict=False, missing keys must all be frozen encoder weights."""
    from miphei_vit_b200.generators.mipheivit import get_vitmatte

    g = get_vitmatte("hoptimus0", 128, 3, use_lora=True, embed_dim=128, depth=2, num_heads=2, hidden=256)
    sd = g.state_dict()
    part = {k: v for k, v in sd.items() if ".lora" in k or k.startswith("decoder.")}
    info = g.load_state_dict(part, strict=False)
    assert not info.unexpected_keys
    assert all(("encoder.vit." in k) and (".lora" not in k) for k in info.missing_keys)
