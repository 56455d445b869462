"""QConv2D: the reference's drop-in nn.Conv2d replacement (python/qtorch/nn/QConv2d.py:6-22) on liblowbit-cnn."""
from __future__ import annotations

from torch import nn

from .functional import qconv2d


class QConv2D(nn.Conv2d):
    """Same constructor as the reference: bias is disabled (QConv2d.py:9-10), weights stay fp32 OIHW and are quantised
    per call.  (The reference stores the weight as a VECT_C=4 view, QConv2d.py:14; that is a cuDNN layout detail - this
    module keeps plain OIHW and lets the library's prepack choose the kernel layout.)"""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, num_bits=8,
                 min_value=None, max_value=None, stochastic=True):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, False)
        self.num_bits = num_bits
        self.min_value, self.max_value = min_value, max_value
        self.stochastic = stochastic

    def forward(self, input):
        return qconv2d(input, self.weight, self.stride, self.padding, self.dilation, self.groups)
