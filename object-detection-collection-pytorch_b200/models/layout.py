"""Head-tensor layout: making `permute(0, 2, 3, 1) + reshape` free.

The reference's head() turns the last conv's NCHW output into [N,S_h,S_w,A,5+C] with a permute and a
reshape (reference models/yolov2.py:338-362); on a contiguous NCHW tensor that reshape is a full
copy of the head tensor (one extra read + write of y, and the same again for its gradient),
immediately upstream of the kernels of this package.  If the head convolution runs in
channels_last memory format, cuDNN already writes NHWC: the permute is then a contiguous view,
the reshape free, and autograd hands the kernel's dL/dy back to the conv without a copy either.
Nothing here changes values or state_dict keys.
"""
from __future__ import annotations

import torch


def use_channels_last_head(module: torch.nn.Module) -> torch.nn.Module:
    """Put every Conv2d of `module` (the detection head, or the whole model) into channels_last."""
    for m in module.modules():
        if isinstance(m, torch.nn.Conv2d):
            m.to(memory_format=torch.channels_last)
    return module


def head_tensor(conv_out: torch.Tensor, num_anchor_box: int) -> torch.Tensor:
    """NCHW-shaped conv output [N, A*(5+C), S_h, S_w] -> [N, S_h, S_w, A, 5+C], the reference's
    permute + reshape; a view (no copy) when conv_out is channels_last."""
    n, ch, s_h, s_w = conv_out.shape
    return conv_out.permute(0, 2, 3, 1).reshape(n, s_h, s_w, num_anchor_box, ch // num_anchor_box)


def is_free_view(conv_out: torch.Tensor, y: torch.Tensor) -> bool:
    """True if y aliases conv_out's storage (the layout change cost nothing)."""
    return y.is_contiguous() and y.data_ptr() == conv_out.data_ptr()
