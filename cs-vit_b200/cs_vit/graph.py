"""CUDA-graph capture of ``Poser.predict_batch`` for fixed input shapes.

A Swin-B step is ~190 kernel launches of 20-400 us each; launched one by one from Python the GPU idles a few
per cent of the step between them.  ``GraphedPredict`` records the whole step once (all launches go through the C
ABI on the capture stream, the torch glue ops of the fp32 tail are captured with them) and replays it with one
``cudaGraphLaunch``.  Inputs are copied into static buffers, outputs are returned as views of static buffers
(clone them if they must outlive the next call).
"""
from __future__ import annotations

from typing import Dict

import torch

KEYS = ("patches", "square_bboxes", "timestamp", "focal", "princpt")


class GraphedPredict:
    def __init__(self, model, example: Dict[str, torch.Tensor], warmup: int = 2):
        self.model = model
        self.static_in = {k: example[k].clone() for k in KEYS}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):       # populates the packed-weight caches and the allocator pool
                self._run()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.static_out = self._run()

    def _run(self):
        i = self.static_in
        return self.model.predict_batch(i["patches"], i["square_bboxes"], i["timestamp"], i["focal"], i["princpt"])

    def __call__(self, img_tensor, square_bboxes, timestamp, focal, princpt) -> Dict[str, torch.Tensor]:
        for k, v in zip(KEYS, (img_tensor, square_bboxes, timestamp, focal, princpt)):
            if v.data_ptr() != self.static_in[k].data_ptr():
                self.static_in[k].copy_(v, non_blocking=True)
        self.graph.replay()
        return self.static_out
