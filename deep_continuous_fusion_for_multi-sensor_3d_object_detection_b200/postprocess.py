"""Rotated-box post-process on device, behind the method names the reference's evaluator uses.

`PostProcess.get_bboxes / NMS_SAT / NMS_IOU` take and return what `Test.get_bboxes` (test.py:88-108),
`Test.NMS_SAT` (test.py:142-175) and `Test.NMS_IOU` (test.py:110-140) take and return, so the caller at
test.py:83-85 can switch by replacing `self.` with a PostProcess instance.  Internally the boxes stay on the
GPU in a padded (B, cap, 7) layout; the list-of-tensors views are built only at the boundary.
"""
from __future__ import annotations

import torch

from . import ops


def pad_boxes(boxes_list, cap=None, device="cuda"):
    """list of (n_b,7) tensors -> padded (B,cap,7) f32 + counts (B,) i32 on `device`."""
    B = len(boxes_list)
    nmax = max([int(b.shape[0]) for b in boxes_list] + [1])
    cap = max(cap or 0, nmax)
    cap = (cap + 63) // 64 * 64
    out = torch.zeros((B, cap, 7), dtype=torch.float32, device=device)
    counts = torch.zeros((B,), dtype=torch.int32)
    for i, b in enumerate(boxes_list):
        n = int(b.shape[0])
        if n:
            out[i, :n] = b.to(device=device, dtype=torch.float32)
        counts[i] = n
    return out, counts.to(device)


class PostProcess:
    def __init__(self, config=None, cap: int = 4096, device="cuda"):
        self.config = config or {}
        self.cap = int(cap)
        self.device = device

    # -- P-1 -------------------------------------------------------------------------------------------
    def get_bboxes(self, pred_cls, pred_reg, score_threshold=0.8):
        """-> list (len B) of (N_b,7) tensors, anchor 0 then anchor 1, row-major cells (test.py:88-108)."""
        boxes, counts, raw = self.get_bboxes_padded(pred_cls, pred_reg, score_threshold)
        counts_h, raw_h = counts.tolist(), raw.tolist()
        if any(r > self.cap for r in raw_h):
            raise RuntimeError(f"get_bboxes: {max(raw_h)} boxes exceed cap={self.cap}; raise PostProcess(cap=...)")
        return [boxes[b, :counts_h[b]] for b in range(boxes.shape[0])]

    def get_bboxes_padded(self, pred_cls, pred_reg, score_threshold=0.8):
        pred_cls = pred_cls.to(self.device, torch.float32)
        pred_reg = pred_reg.to(self.device, torch.float32)
        return ops.get_bboxes(pred_cls, pred_reg, score_threshold, self.cap)

    # -- P-2..P-4 --------------------------------------------------------------------------------------
    def NMS_SAT(self, pred_bboxes):
        """list of (N_b,7) -> list of lists of kept (7,) rows, input order (test.py:142-175)."""
        return self._rows(pred_bboxes, self.nms_sat_indices(pred_bboxes))

    def nms_sat_indices(self, pred_bboxes):
        """list of (N_b,7) -> list of int64 index tensors (ascending kept input indices)."""
        boxes, counts = pad_boxes(pred_bboxes, device=self.device)
        keep, kcnt = ops.nms_sat(boxes, counts)
        return self._unpad(keep, kcnt)

    # -- P-7 -------------------------------------------------------------------------------------------
    def NMS_IOU(self, pred_bboxes, nms_iou_score_theshold=0.01):
        boxes, counts = pad_boxes(pred_bboxes, device=self.device)
        keep, kcnt = ops.nms_iou(boxes, counts, nms_iou_score_theshold)
        return self._rows(pred_bboxes, self._unpad(keep, kcnt))

    # -- P-5/P-6/P-8 -----------------------------------------------------------------------------------
    def box3d_iou(self, boxes_a, boxes_b):
        """(na,7), (nb,7) [x,y,z,l,w,h,yaw] -> (iou3d, iou2d) float64 (na,nb); IOU.py:91-155 conventions."""
        a = boxes_a.to(self.device, torch.float32).reshape(-1, 7)
        b = boxes_b.to(self.device, torch.float32).reshape(-1, 7)
        return ops.box_iou(a, b)

    @staticmethod
    def _unpad(keep, kcnt):
        kc = kcnt.tolist()
        return [keep[b, :kc[b]].to(torch.int64) for b in range(keep.shape[0])]

    @staticmethod
    def _rows(pred_bboxes, idx_list):
        out = []
        for boxes, idx in zip(pred_bboxes, idx_list):
            sel = boxes[idx.to(boxes.device)] if idx.numel() else boxes[:0]
            out.append([sel[i] for i in range(sel.shape[0])])
        return out
