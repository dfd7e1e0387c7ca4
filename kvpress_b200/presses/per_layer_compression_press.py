"""PerLayerCompressionPress: one compression ratio per layer around a ScorerPress.
API mirror of `/root/reference/kvpress/presses/per_layer_compression_press.py:17-72`. (Caches whose
layers have different lengths need an attention backend that takes per-layer key lengths from the cache,
as in the reference.)"""
from __future__ import annotations

import logging
from dataclasses import dataclass
from typing import List

from torch import nn

from kvpress_b200.presses.base_press import BasePress
from kvpress_b200.presses.scorer_press import ScorerPress

logger = logging.getLogger(__name__)


@dataclass
class PerLayerCompressionPress(BasePress):
    press: ScorerPress
    compression_ratios: List[float]

    def __post_init__(self):
        assert isinstance(self.press, ScorerPress), "PerLayerCompressionPress requires a ScorerPress as input"
        logger.warning("Per layer compression is experimental: the attention backend must accept caches whose "
                       "layers have different lengths (flash attention does).")

    def post_init_from_model(self, model):
        self.press.post_init_from_model(model)

    def forward_hook(self, module: nn.Module, input, kwargs: dict, output: list):
        saved = self.press.compression_ratio
        self.press.compression_ratio = self.compression_ratios[module.layer_idx]
        try:
            return self.press.forward_hook(module, input, kwargs, output)
        finally:
            self.press.compression_ratio = saved

    @property
    def compression_ratio(self):
        return sum(self.compression_ratios) / len(self.compression_ratios)

    @compression_ratio.setter
    def compression_ratio(self, value):
        raise AttributeError(f"compression ratio cannot be set for {type(self).__name__}")
