"""Visualisation helper for ``Poser._vis`` (ref:cs_vit/net/ti_poser.py:780-813, ref:cs_vit/utils/img.py:393-456).

Host-side cv2 drawing, outside the GPU hot path (SURVEY.md §2.1 marks image utils out of scope); kept minimal so
that ``Poser.forward`` can return the reference's ``logs["image"]["img_reproj"]`` entry when frames are on disk.
"""
from typing import List, Tuple, Union

import torch
import torch.nn.functional as F


def expand_bbox_square(bboxes: torch.Tensor, expansion_ratio: float = 1.0) -> torch.Tensor:
    """``[...,4]`` xyxy boxes -> squares on the longer side, scaled about the centre   (ref:cs_vit/utils/img.py:25-52)."""
    x1, y1, x2, y2 = bboxes.unbind(-1)
    half = torch.max(x2 - x1, y2 - y1) * 0.5 * expansion_ratio
    cx, cy = (x1 + x2) * 0.5, (y1 + y2) * 0.5
    return torch.stack([cx - half, cy - half, cx + half, cy + half], dim=-1)


def square_boxes_from_tight(tight_bbox: torch.Tensor, expansion_ratio: float) -> torch.Tensor:
    """The square crop boxes of ``crop_tensor_with_square_box`` (ref:cs_vit/utils/img.py:358-370)."""
    centers = (tight_bbox[:, :2] + tight_bbox[:, 2:]) / 2
    side = (tight_bbox[:, 2:] - tight_bbox[:, :2]).max(dim=1).values * expansion_ratio
    half = torch.stack([side, side], dim=1) / 2
    return torch.cat([centers - half, centers + half], dim=1)


def crop_and_resize_host(img: torch.Tensor, box: torch.Tensor, size: int) -> torch.Tensor:
    """One ``[C,H,W]`` image, one xyxy box -> ``[C,size,size]``: output pixel (u, v) samples the source bilinearly at
    ``(x1 + u (x2-x1)/(size-1), y1 + v (y2-y1)/(size-1))``, zeros outside the image.  This is what
    ``kornia.geometry.transform.crop_and_resize(mode='bilinear', padding_mode='zeros', align_corners=True)`` computes for an
    axis-aligned box (perspective transform box -> [0, size-1]^2, then ``grid_sample`` with ``align_corners=True``): the data
    sets' evaluation-time crop (ref:cs_vit/utils/img.py:372-388).  Host-side (data-loader workers); the device-side
    equivalent for frames already in HBM is ``cs_vit.ops.crop_resize``."""
    C, H, W = img.shape
    x1, y1, x2, y2 = [float(v) for v in box]
    t = torch.arange(size, dtype=torch.float32, device=img.device) / max(size - 1, 1)
    xs, ys = x1 + t * (x2 - x1), y1 + t * (y2 - y1)
    gx = 2.0 * xs / max(W - 1, 1) - 1.0
    gy = 2.0 * ys / max(H - 1, 1) - 1.0
    grid = torch.stack([gx[None, :].expand(size, size), gy[:, None].expand(size, size)], dim=-1)[None]
    return F.grid_sample(img[None].float(), grid, mode="bilinear", padding_mode="zeros", align_corners=True)[0]


def crop_tensor_with_square_box(img_list: Union[List[torch.Tensor], torch.Tensor], tight_bbox: torch.Tensor,
                                expansion_ratio: float = 2.0, output_size: int = 224) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Tight xyxy boxes ``[N,4]`` -> (``[N,C,S,S]`` crops, ``[N]`` source-pixels-per-output-pixel, ``[N,4]`` square boxes)
    (ref:cs_vit/utils/img.py:339-390)."""
    square = square_boxes_from_tight(tight_bbox, expansion_ratio)
    crops = [crop_and_resize_host(img, box, output_size) for img, box in zip(img_list, square)]
    scales = (square[:, 2] - square[:, 0]) / output_size
    return torch.stack(crops), scales, square



def reprojection_overlay(predict, batch, connection):
    import cv2
    import numpy as np

    jc = predict["joint_cam"][0].detach().float().cpu()               # [T,21,3]
    f, c = batch["focal"][0].float().cpu(), batch["princpt"][0].float().cpu()
    uv = (jc[..., :2] * f[:, None] + c[:, None] * jc[..., 2:]) / jc[..., 2:]
    frames = []
    for t, path in enumerate(batch["imgs_path"][0]):
        img = cv2.imread(path)
        if img is None:
            raise FileNotFoundError(path)
        img = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
        if batch["flip"][0]:
            img = np.ascontiguousarray(img[:, ::-1])
        for pts, colour in ((batch["joint_img"][0][t].cpu(), (0, 255, 0)), (uv[t], (255, 0, 0))):
            for a, b in connection:
                pa, pb = pts[a].tolist(), pts[b].tolist()
                cv2.line(img, (int(pa[0]), int(pa[1])), (int(pb[0]), int(pb[1])), colour, 1)
        frames.append(torch.from_numpy(img).permute(2, 0, 1).float() / 255)
    return torch.stack(frames)
