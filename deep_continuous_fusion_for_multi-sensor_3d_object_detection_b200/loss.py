"""LossTotal with its target assignment on the device: drop-in for loss.py:33-72 of the reference.

The reference builds the positive / negative cell lists of every frame in Python (5x5 windows around the ground-truth
centres, np.random.shuffle, a rejection loop over np.random.randint; loss.py:74-127), copies them to the GPU with
torch.tensor(...).cuda(), and evaluates the class / regression terms frame by frame and box by box (loss.py:129-189).
Here `cf_loss_targets` (csrc/cf_loss.cu) assigns the targets of the whole batch in one launch from random draws made on
the device, and the two loss terms are evaluated for all frames and boxes at once with gathers -- no Python loop over
frames or boxes, no host list, no H2D copy, nothing that synchronises with the host, so the step can be captured in a
CUDA graph.

Reference behaviours kept, and flagged:
  * loss.py:71 assigns `total_loss = ...` inside the frame loop instead of accumulating, so the value returned (and
    back-propagated, train.py:33-35) is the LAST frame's loss only.  `batch_reduction="last"` (default) reproduces that;
    "sum" / "mean" are what was presumably meant.
  * loss.py:125 tests `sample > sample_threshold` after counting, so neg_sample_threshold + 1 = 129 negatives are drawn.
  * overlapping windows put the same cell into the positive list more than once (loss.py:96); duplicates are kept.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import geometry, ops
from .model import AnchorBoundingBoxFeature


class LossTotal(nn.Module):
    """forward(reference_bboxes_batch (B,M,8), num_ref_bbox_batch (B), pred_cls (B,4,H,W), pred_reg (B,14,H,W)) -> loss (1,)

    `draws=(shuffle_keys, candidates)` fixes the random numbers (RNG contract of cf_loss_targets, include/cf_b200.h);
    by default they come from `generator` (a torch.Generator on the prediction's device) or the global CUDA generator."""

    def __init__(self, config, batch_reduction: str = "last", candidates_per_negative: int = 4, generator=None):
        super().__init__()
        if batch_reduction not in ("last", "sum", "mean"):
            raise ValueError("batch_reduction must be 'last' (the reference, loss.py:71), 'sum' or 'mean'")
        self.config = config
        self.batch_reduction = batch_reduction
        self.candidates_per_negative = int(candidates_per_negative)
        self.generator = generator
        anchors = AnchorBoundingBoxFeature(config)()          # (14, H, W): anchor 0 | anchor 1, (x, y, z, l, w, h, yaw)
        self.register_buffer("anchor_set", anchors, persistent=False)
        self.regress_type = int(config["regress_type"])
        self.positive_range = int(config["positive_range"])
        self.pos_threshold = int(config["pos_sample_threshold"])
        self.neg_threshold = int(config["neg_sample_threshold"])
        self.gain = float(config["regress_loss_gain"])
        self.scales = geometry.voxel_scales(config)           # loss.py:80-83 == data_import_carla.py:35-39
        self.reduced_scale = float(config["anchor_bbox_feature"]["reduced_scale"])

    # ---- random draws ------------------------------------------------------------------------------------------------
    def draw(self, B: int, M: int, H: int, W: int, device):
        """(shuffle_keys (B, M*R*R) fp32 in [0,1), candidates (B, L, 2) int32 uniform over the map): the draws one
        forward consumes.  L = candidates_per_negative * (neg_threshold + 1); a frame whose rejection loop would need
        more than L draws reports neg_count < neg_threshold + 1 (at most pos_threshold of the H*W cells reject)."""
        R, L = self.positive_range, self.candidates_per_negative * (self.neg_threshold + 1)
        keys = torch.rand((B, M * R * R), device=device, generator=self.generator)
        cx = torch.randint(0, H, (B, L), device=device, generator=self.generator, dtype=torch.int32)
        cy = torch.randint(0, W, (B, L), device=device, generator=self.generator, dtype=torch.int32)
        return keys, torch.stack((cx, cy), dim=2).contiguous()

    def targets(self, reference_bboxes_batch, num_ref_bbox_batch, H, W, draws=None):
        B, M, _ = reference_bboxes_batch.shape
        keys, cand = draws if draws is not None else self.draw(B, M, H, W, reference_bboxes_batch.device)
        return ops.loss_targets(reference_bboxes_batch, num_ref_bbox_batch, H, W, self.scales, self.reduced_scale,
                                self.positive_range, self.regress_type, self.pos_threshold, self.neg_threshold, keys, cand)

    # ---- the two terms, all frames and boxes at once -----------------------------------------------------------------
    def class_term(self, pred_cls, cells, count, label: int):
        """sum over the two anchors of CrossEntropy(mean) of the gathered 2-class logits against `label`
        (getClassSum, loss.py:129-144, called for channels [:2] and [2:4], loss.py:62-63); 0 for an empty list."""
        B, _, H, W = pred_cls.shape
        P = cells.shape[1]
        logits = pred_cls.reshape(B, 2, 2, H * W)                                     # (frame, anchor, class, cell)
        idx = cells.clamp(min=0).long()[:, None, None, :].expand(B, 2, 2, P)
        logp = F.log_softmax(torch.gather(logits, 3, idx), dim=2)[:, :, label, :]     # (B, anchor, P)
        mask = torch.arange(P, device=cells.device)[None, :] < count[:, None]
        s = -(logp * mask[:, None, :]).sum(dim=(1, 2))
        return torch.where(count > 0, s / count.clamp(min=1), torch.zeros_like(s))

    def regress_term(self, pred_reg, ref, reg_cells):
        """getRegSum / LossReg (loss.py:146-189): per ground-truth box the SmoothL1 between the predicted offsets of both
        anchors at the box's regression cells and the box encoded against those anchors, mean over (cells, 2, 7); summed
        over the boxes that have a cell."""
        B, _, H, W = pred_reg.shape
        _, M, RR = reg_cells.shape
        mask = reg_cells >= 0                                                         # (B, M, RR)
        idx = reg_cells.clamp(min=0).long().reshape(B, 1, M * RR)
        pred = torch.gather(pred_reg.reshape(B, 14, H * W), 2, idx.expand(B, 14, M * RR))
        pred = pred.reshape(B, 2, 7, M, RR).permute(0, 3, 4, 1, 2)                    # (B, M, RR, anchor, 7)
        anc = self.anchor_set.reshape(1, 14, H * W).expand(B, 14, H * W)
        anc = torch.gather(anc, 2, idx.expand(B, 14, M * RR)).reshape(B, 2, 7, M, RR).permute(0, 3, 4, 1, 2)
        has = mask.any(dim=2)                                                         # boxes with a regression cell
        safe = torch.where(has[:, :, None], ref[:, :, :7], torch.ones_like(ref[:, :, :7]))   # rows past num_ref are padding
        r = safe[:, :, None, None, :]
        xy = (r[..., :2] - anc[..., :2]) / torch.sqrt(anc[..., 3:4] ** 2 + anc[..., 4:5] ** 2)
        z = (r[..., 2:3] - anc[..., 2:3]) / anc[..., 5:6]
        whd = torch.log(r[..., 3:6] / anc[..., 3:6])
        d = r[..., 6] - anc[..., 6]
        ori = torch.atan2(torch.sin(d), torch.cos(d))
        target = torch.cat((xy, z, whd, ori[..., None]), dim=-1).expand_as(pred)
        l1 = F.smooth_l1_loss(pred, target, reduction="none")
        l1 = torch.where(mask[:, :, :, None, None], l1, torch.zeros_like(l1)).sum(dim=(2, 3, 4))    # (B, M)
        n = mask.sum(dim=2)
        return torch.where(n > 0, l1 / (n.clamp(min=1) * 14).to(l1.dtype), torch.zeros_like(l1)).sum(dim=1)

    def per_frame(self, reference_bboxes_batch, num_ref_bbox_batch, pred_cls, pred_reg, draws=None):
        """(B,) loss of every frame: class terms of both lists + regress_loss_gain * regression term (loss.py:62-71)."""
        _, _, H, W = pred_cls.shape
        # (the CARLA reader collates object_data as it finds it in the HDF5 file; the device path works on fp32 on the predictions' device)
        reference_bboxes_batch = reference_bboxes_batch.to(device=pred_cls.device, dtype=torch.float32)
        pos, npos, neg, nneg, reg = self.targets(reference_bboxes_batch, num_ref_bbox_batch, H, W, draws)
        cls = self.class_term(pred_cls, pos, npos, 1) + self.class_term(pred_cls, neg, nneg, 0)
        return cls + self.gain * self.regress_term(pred_reg, reference_bboxes_batch.to(pred_reg.dtype), reg)

    def forward(self, reference_bboxes_batch, num_ref_bbox_batch, predicted_class_feature_batch, predicted_regress_feature_batch,
                draws=None):
        per = self.per_frame(reference_bboxes_batch, num_ref_bbox_batch, predicted_class_feature_batch,
                             predicted_regress_feature_batch, draws)
        if self.batch_reduction == "last":
            return per[-1:]
        return per.sum(dim=0, keepdim=True) if self.batch_reduction == "sum" else per.mean(dim=0, keepdim=True)
