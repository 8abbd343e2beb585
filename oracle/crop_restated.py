"""CPU restatement of the reference's evaluation-time hand crop.  TEST INFRASTRUCTURE: only tests/, smoke() and bench.py's CPU
baseline may import this.

    crop_tensor_with_square_box(img_list, tight_bbox, expansion_ratio, output_size)      ref:cs_vit/utils/img.py:339-390

PARITY UNPINNED for the resampling step: the reference delegates it to ``kornia.geometry.transform.crop_and_resize(...,
mode='bilinear', padding_mode='zeros', align_corners=True)`` (ref :372-383); kornia is not in this image and is not vendored by
the reference, so the reference's function cannot be executed here and no golden vector exists.  What is restated is kornia's
published algorithm (kornia 0.7/0.8 ``crop_by_boxes``): the perspective transform that maps the source box corners onto
``[0, 0] .. [S-1, S-1]`` followed by ``warp_perspective(align_corners=True)`` = ``grid_sample`` of the inverse map; for an
axis-aligned box the inverse map is ``x = x1 + u (x2 - x1) / (S - 1)``, ``y = y1 + v (y2 - y1) / (S - 1)``.  The box arithmetic
(ref :358-370) is plain torch code in the reference and is restated exactly.

Written with explicit neighbour loops in numpy float64 so that it shares nothing with the CUDA kernel or with
``torch.nn.functional.grid_sample`` (which cs_vit.utils.img.crop_and_resize_host uses and the tests cross-check against)."""
import numpy as np


def square_boxes(tight: np.ndarray, expansion_ratio: float) -> np.ndarray:
    """ref:cs_vit/utils/img.py:358-370."""
    tight = np.asarray(tight, dtype=np.float64)
    centers = (tight[:, :2] + tight[:, 2:]) / 2
    side = (tight[:, 2:] - tight[:, :2]).max(axis=1) * expansion_ratio
    half = np.stack([side, side], axis=1) / 2
    return np.concatenate([centers - half, centers + half], axis=1)


def crop_and_resize(img: np.ndarray, box, size: int) -> np.ndarray:
    """img [C,H,W] -> [C,size,size]; bilinear, zeros outside the frame, box endpoints land on the first / last output pixel."""
    img = np.asarray(img, dtype=np.float64)
    C, H, W = img.shape
    x1, y1, x2, y2 = [float(v) for v in box]
    out = np.zeros((C, size, size), dtype=np.float64)
    den = max(size - 1, 1)
    for v in range(size):
        sy = y1 + v * (y2 - y1) / den
        iy = int(np.floor(sy)); ay = sy - iy
        for u in range(size):
            sx = x1 + u * (x2 - x1) / den
            ix = int(np.floor(sx)); ax = sx - ix
            for dy, wy in ((0, 1 - ay), (1, ay)):
                for dx, wx in ((0, 1 - ax), (1, ax)):
                    yy, xx = iy + dy, ix + dx
                    if 0 <= yy < H and 0 <= xx < W:
                        out[:, v, u] += wy * wx * img[:, yy, xx]
    return out


def crop_tensor_with_square_box(imgs, tight_bbox, expansion_ratio: float = 2.0, output_size: int = 224):
    """-> (crops [N,C,S,S], scales [N], square boxes [N,4])   ref:cs_vit/utils/img.py:339-390."""
    sq = square_boxes(tight_bbox, expansion_ratio)
    crops = np.stack([crop_and_resize(img, box, output_size) for img, box in zip(imgs, sq)])
    return crops, (sq[:, 2] - sq[:, 0]) / output_size, sq
