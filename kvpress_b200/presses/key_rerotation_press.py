"""KeyRerotationPress: RoPE-aware wrapper — after pruning, every kept key is re-rotated from its original
position to its new (compacted) position, so the cache looks like a contiguous prefix to RoPE.

API mirror of `/root/reference/kvpress/presses/key_rerotation_press.py:14-152` (first "next"
row). The reference sorts the top-k indices ascending before gathering — exactly the order the sm_100a
compaction emits — so here the output ROW ORDER matches the reference too. The rotation
`k * cos(delta*inv_freq) + rotate_half(k) * sin(delta*inv_freq)`, delta = new - old position, is fused
into the K half of the compaction kernel with the reference's rounding points.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
from torch import nn

from kvpress_b200 import native
from kvpress_b200.presses.base_press import BasePress
from kvpress_b200.presses.scorer_press import ScorerPress, kept_count


@dataclass
class KeyRerotationPress(BasePress):
    """Wraps a ScorerPress; keeps its selection, re-rotates the kept keys."""

    press: ScorerPress

    def __post_init__(self):
        assert isinstance(self.press, ScorerPress)

    def post_init_from_model(self, model):
        self.press.post_init_from_model(model)

    @property
    def compression_ratio(self):
        return self.press.compression_ratio

    @compression_ratio.setter
    def compression_ratio(self, value):
        self.press.compression_ratio = value

    def compress(self, module: nn.Module, hidden_states, keys: torch.Tensor, values: torch.Tensor, attentions,
                 kwargs: dict) -> tuple[torch.Tensor, torch.Tensor]:
        if self.press.compression_ratio == 0:
            return keys, values
        scores = self.press.score(module, hidden_states, keys, values, attentions, kwargs)
        n_kept = kept_count(keys.shape[2], self.press.compression_ratio)
        k_out, v_out, _ = native.scores_compress_rerotate(scores, keys, values, n_kept, module.rotary_emb.inv_freq)
        return k_out, v_out
