"""PyramidKVPress: SnapKV scores with a per-layer budget that shrinks linearly with depth.

API mirror of `/root/reference/kvpress/presses/pyramidkv_press.py:17-112`. Only the number of kept
positions differs from SnapKVPress, so the fused sm_100a SnapKV path (window attention, pooling, top-k,
compaction) is used unchanged with this layer's n_kept.
"""
from __future__ import annotations

from dataclasses import dataclass

from torch import nn

from kvpress_b200.presses.snapkv_press import SnapKVPress


def pyramid_layer_budget(q_len: int, compression_ratio: float, window_size: int, beta: int, num_layers: int,
                         layer_idx: int) -> int:
    """pyramidkv_press.py:47-86. The budgets of all layers form an arithmetic sequence from `hi` (layer 0)
    down to `lo` (last layer) whose mean is q_len*(1-ratio); lo = mean/beta. When the sequence would leave
    [window_size, q_len] the plain ScorerPress budget (rounded, not truncated) is used for every layer."""
    assert beta >= 1, "Beta should >= 1"
    mean = q_len * (1 - compression_ratio)  # = max_capacity_prompt - window_size
    lo = mean / beta
    hi = 2 * mean - lo
    if hi >= q_len - window_size:
        hi = q_len - window_size
        lo = 2 * mean - hi
    if not (q_len >= hi >= lo >= window_size):
        return round(q_len * (1 - compression_ratio))
    step = (hi - lo) / (num_layers - 1)
    return round(hi - layer_idx * step)


@dataclass
class PyramidKVPress(SnapKVPress):
    compression_ratio: float = 0.0
    window_size: int = 64
    kernel_size: int = 5
    beta: int = 20

    def get_layer_budget(self, module: nn.Module, q_len: int) -> int:
        return pyramid_layer_budget(q_len, self.compression_ratio, self.window_size, self.beta,
                                    module.config.num_hidden_layers, module.layer_idx)

    def _n_kept(self, module: nn.Module, k_len: int) -> int:
        return self.get_layer_budget(module, k_len)
