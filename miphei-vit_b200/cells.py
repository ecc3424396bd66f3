"""Host-side mirror of the reference's MeanCellExtrator (src/utils.py:17-121) on the B200 per-nucleus reduction kernel.

Same constructor, same `forward(pred, target, nuclei) -> (pred_means, target_means, cell_ids)` contract (rows of image 0
first, ids ascending per image). The scale_factor < 1 path keeps the reference's F.interpolate glue; the reduction itself
(torch.unique + scatter_add_ per image in the reference) is one kernel launch for the whole batch."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops


class _CellMeans(torch.autograd.Function):
    """(pred, target, nuclei) -> (pred_means, target_means, ids); differentiable in pred and target like the reference's
    scatter_add_ formulation (the cell loss of training_step back-propagates through it, src/models.py:120-131)."""

    @staticmethod
    def forward(ctx, pred, target, nuclei):
        pm, tm, ids, nu, cnt = ops.cell_means(pred.float().contiguous(), target.float().contiguous(), nuclei,
                                              return_counts=True)
        ctx.save_for_backward(ids, cnt, nu, nuclei)
        ctx.shape, ctx.dtypes = pred.shape, (pred.dtype, target.dtype)
        ids_out = ids.to(nuclei.dtype)
        ctx.mark_non_differentiable(ids_out)
        return pm.to(pred.dtype), tm.to(pred.dtype), ids_out

    @staticmethod
    def backward(ctx, dpm, dtm, _dids):
        ids, cnt, nu, nuclei = ctx.saved_tensors
        B, C, H, W = ctx.shape
        gp = gt = None
        if ctx.needs_input_grad[0] and dpm is not None:
            gp = (ops.cell_means_bwd(dpm, ids, cnt, nu, nuclei, B, C, H, W) if ids.numel() else
                  torch.zeros(ctx.shape, device=dpm.device)).to(ctx.dtypes[0])
        if ctx.needs_input_grad[1] and dtm is not None:
            gt = (ops.cell_means_bwd(dtm, ids, cnt, nu, nuclei, B, C, H, W) if ids.numel() else
                  torch.zeros(ctx.shape, device=dtm.device)).to(ctx.dtypes[1])
        return gp, gt, None


class MeanCellExtrator(nn.Module):
    def __init__(self, scale_factor=1.):
        super().__init__()
        if not (0. < scale_factor <= 1):
            raise ValueError("scale_factor should be between 0 and 1")
        self.scale_factor = scale_factor

    def forward(self, pred, target, nuclei):
        """pred / target [B, C, H, W], nuclei [B, H, W] or [B, 1, H, W] integer labels (0 = background)."""
        labels = nuclei[:, None].long() if nuclei.ndim == 3 else nuclei
        target = pred.new_zeros(pred.shape) if target is None else target
        sf = self.scale_factor
        if sf < 1.:  # the reference shrinks maps by area averaging and the label map by exact nearest sampling first
            pred, target = (F.interpolate(t, scale_factor=sf, mode="area") for t in (pred, target))
            labels = F.interpolate(labels.float(), scale_factor=sf, mode="nearest-exact").long()
        return self.extract_mean(pred, target, labels)

    def extract_mean(self, pred, target, nuclei):
        return _CellMeans.apply(pred, target, nuclei)
