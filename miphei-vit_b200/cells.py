"""Host-side mirror of the reference's MeanCellExtrator (src/utils.py:17-121) on the B200 per-nucleus reduction kernel.

Same constructor, same `forward(pred, target, nuclei) -> (pred_means, target_means, cell_ids)` contract (rows of image 0
first, ids ascending per image). The scale_factor < 1 path keeps the reference's F.interpolate glue; the reduction itself
(torch.unique + scatter_add_ per image in the reference) is one kernel launch for the whole batch."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops


class MeanCellExtrator(nn.Module):
    def __init__(self, scale_factor=1.):
        super().__init__()
        if not (0. < scale_factor <= 1):
            raise ValueError("scale_factor should be between 0 and 1")
        self.scale_factor = scale_factor

    def forward(self, pred, target, nuclei):
        if target is None:
            target = torch.zeros_like(pred)
        if nuclei.ndim == 3:
            nuclei = torch.unsqueeze(nuclei, dim=1).long()
        if self.scale_factor < 1.:
            pred = F.interpolate(pred, scale_factor=self.scale_factor, mode='area')
            target = F.interpolate(target, scale_factor=self.scale_factor, mode='area')
            nuclei = F.interpolate(nuclei.float(), scale_factor=self.scale_factor, mode='nearest-exact').long()
        return self.extract_mean(pred, target, nuclei)

    def extract_mean(self, pred, target, nuclei):
        dt = pred.dtype
        pm, tm, ids, _ = ops.cell_means(pred.float().contiguous(), target.float().contiguous(), nuclei)
        return pm.to(dt), tm.to(dt), ids.to(nuclei.dtype)
