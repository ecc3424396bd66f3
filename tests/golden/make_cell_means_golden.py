"""Generates tests/golden/cell_means.pt from the REFERENCE's MeanCellExtrator.extract_mean (src/utils.py:49-121), run in
this container (python tests/golden/make_cell_means_golden.py). Counts come from the oracle (the reference does not
return them)."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import cell_means as oc  # noqa: E402
from test_cell_means_cpu import _case, _reference_class  # noqa: E402

cases = []
ref = _reference_class()(scale_factor=1.0)
for seed, B, C, S, n_cells, empty in [(11, 2, 4, 64, 10, ()), (12, 3, 16, 128, 40, (1,)), (13, 2, 3, 64, 0, ())]:
    pred, target, nuclei = _case(seed, B=B, C=C, S=S, n_cells=n_cells, empty=empty)
    rp, rt, rid = ref.extract_mean(pred, target, nuclei.unsqueeze(1))
    _, _, _, cnt = oc.extract_mean(pred, target, nuclei)
    cases.append(dict(seed=seed, B=B, C=C, S=S, n_cells=n_cells, empty=list(empty), pred_means=rp, target_means=rt,
                      ids=rid.long(), counts=cnt))
torch.save(dict(cases=cases, source="reference src/utils.py MeanCellExtrator.extract_mean"), os.path.join(HERE, "cell_means.pt"))
print("wrote", len(cases), "cases")
