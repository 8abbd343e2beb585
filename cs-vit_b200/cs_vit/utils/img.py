"""Visualisation helper for ``Poser._vis`` (ref:cs_vit/net/ti_poser.py:780-813, ref:cs_vit/utils/img.py:393-456).

Host-side cv2 drawing, outside the GPU hot path (SURVEY.md §2.1 marks image utils out of scope); kept minimal so
that ``Poser.forward`` can return the reference's ``logs["image"]["img_reproj"]`` entry when frames are on disk.
"""
import torch


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
