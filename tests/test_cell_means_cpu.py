"""Oracle of the per-nucleus mean extraction (oracle/cell_means.py) pinned to the reference's MeanCellExtrator
(src/utils.py:17-121, imported with its absent third-party modules stubbed) and to the committed golden vectors."""
import importlib.util
import os
import sys
import types

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cell_means as oc  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cell_means.pt")
REF_UTILS = os.path.join(os.environ.get("MIPHEI_REFERENCE", "/root/reference"), "src", "utils.py")


def _case(seed, B=3, C=5, S=64, n_cells=12, empty=()):
    g = torch.Generator().manual_seed(seed)
    pred = torch.rand((B, C, S, S), generator=g)
    target = torch.rand((B, C, S, S), generator=g)
    nuclei = oc.synthetic_nuclei(B, S, n_cells, seed=seed, id_offset=3, empty=empty)
    return pred, target, nuclei


def _reference_class():
    for name in ("pytorch_lightning", "hydra", "hydra.core", "hydra.core.hydra_config", "wandb"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["pytorch_lightning"].LightningModule = object
    sys.modules["hydra.core.hydra_config"].HydraConfig = object
    sys.modules["wandb"].Artifact = object
    spec = importlib.util.spec_from_file_location("ref_utils", REF_UTILS)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.MeanCellExtrator


@pytest.mark.skipif(not os.path.exists(REF_UTILS), reason="/root/reference not present")
@pytest.mark.parametrize("seed,empty", [(1, ()), (2, (1,)), (3, (0, 1, 2))])
def test_oracle_matches_reference_extract_mean(seed, empty):
    pred, target, nuclei = _case(seed, empty=empty)
    ref = _reference_class()(scale_factor=1.0)
    rp, rt, rid = ref.extract_mean(pred, target, nuclei.unsqueeze(1))
    op, ot, oid, _ = oc.extract_mean(pred, target, nuclei)
    assert torch.equal(oid, rid.long())
    assert rp.shape == op.shape
    assert torch.allclose(op, rp, rtol=1e-5, atol=1e-6) and torch.allclose(ot, rt, rtol=1e-5, atol=1e-6)


def test_oracle_matches_golden():
    g = torch.load(GOLDEN, map_location="cpu", weights_only=False)
    for case in g["cases"]:
        pred, target, nuclei = _case(case["seed"], B=case["B"], C=case["C"], S=case["S"], n_cells=case["n_cells"],
                                     empty=tuple(case["empty"]))
        op, ot, oid, cnt = oc.extract_mean(pred, target, nuclei)
        assert torch.equal(oid, case["ids"]) and torch.equal(cnt, case["counts"])
        assert torch.allclose(op, case["pred_means"], rtol=1e-5, atol=1e-6)
        assert torch.allclose(ot, case["target_means"], rtol=1e-5, atol=1e-6)
