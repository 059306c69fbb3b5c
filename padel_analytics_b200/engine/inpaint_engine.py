"""InpaintNet (trajectory-repair 1-D U-Net) on the single fused CUDA kernel csrc/inpaintnet.cu.

Replaces `self.inpaintnet` of the reference BallTracker (/root/reference/trackers/ball_tracker/ball_tracker.py:268-272,
called at :573-576 with hard-coded .cuda()); network definition models.py:101-130."""
from __future__ import annotations

import torch

from .. import _lib as L

_LAYERS = ["down_1.conv", "down_2.conv", "down_3.conv", "buttleneck.conv_1.conv", "buttleneck.conv_2.conv",
           "up_1.conv", "up_2.conv", "up_3.conv", "predictor"]


class InpaintNetEngine:
    """nn.Module-like: __call__(coor (N,L,2) f32, mask (N,L,1) f32) -> (N,L,2) f32; .to(), .eval(), .load_state_dict()."""

    def __init__(self, state_dict: dict | None = None, device: str = "cuda"):
        if not torch.cuda.is_available():
            raise L.PbError("InpaintNetEngine needs a CUDA device (no CPU fallback)")
        L.lib()
        self.device = torch.device(device)
        self.blob = None
        if state_dict is not None:
            self.load_state_dict(state_dict)

    def to(self, device):
        return self

    def eval(self):
        return self

    def load_state_dict(self, sd: dict):
        parts = []
        for name in _LAYERS:
            parts += [sd[f"{name}.weight"].float().reshape(-1), sd[f"{name}.bias"].float().reshape(-1)]
        self.blob = torch.cat(parts).contiguous().to(self.device)
        return self

    @torch.no_grad()
    def __call__(self, coor: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        if self.blob is None:
            raise L.PbError("InpaintNetEngine: no weights loaded")
        N, Lseq, _ = coor.shape
        c = coor.to(self.device, torch.float32).contiguous()
        m = mask.to(self.device, torch.float32).reshape(N, Lseq).contiguous()
        out = torch.empty((N, Lseq, 2), dtype=torch.float32, device=self.device)
        L.check(L.lib().pb_inpaintnet_forward(c.data_ptr(), m.data_ptr(), N, Lseq, self.blob.data_ptr(),
                                              out.data_ptr(), L.stream_ptr()))
        return out
