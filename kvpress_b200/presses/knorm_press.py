"""KnormPress: score = -||k||_2 (https://arxiv.org/pdf/2406.11430).

API mirror of `/root/reference/kvpress/presses/knorm_press.py:15-38`. `compress` runs the fused
sm_100a path: one streaming pass over K (fp32 sum of squares, one rounding to the cache dtype) that
also builds the selection histogram, then the select/compact kernels.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
from torch import nn

from kvpress_b200 import native
from kvpress_b200.presses.scorer_press import ScorerPress


@dataclass
class KnormPress(ScorerPress):
    """Keeps the keys with the smallest L2 norm."""

    needs_hidden_states = False

    def score(self, module: nn.Module, hidden_states, keys: torch.Tensor, values, attentions, kwargs) -> torch.Tensor:
        return native.knorm_score(keys)

    def _fused_compress(self, module, hidden_states, keys, values, attentions, kwargs, n_kept):
        if self._score_is_overridden(KnormPress):
            return None
        k_out, v_out, _, _ = native.knorm_compress(keys, values, n_kept)
        return k_out, v_out
