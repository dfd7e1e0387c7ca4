"""KeyDiffPress: score = -cosine similarity between a key and the mean of the normalised keys
(https://arxiv.org/abs/2504.15364): keys that look like the average key go first.

API mirror of `/root/reference/kvpress/presses/keydiff_press.py:14-46`. `compress` runs the fused sm_100a path:
two streaming passes over K (anchor, then scores + selection histogram; fp32 math, one rounding to the cache
dtype), then the select/compact kernels.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
from torch import nn

from kvpress_b200 import native
from kvpress_b200.presses.scorer_press import ScorerPress


@dataclass
class KeyDiffPress(ScorerPress):
    """Keeps the keys that differ most (in direction) from the average key."""

    needs_hidden_states = False

    def score(self, module: nn.Module, hidden_states, keys: torch.Tensor, values, attentions, kwargs) -> torch.Tensor:
        return native.keydiff_score(keys)

    def _fused_compress(self, module, hidden_states, keys, values, attentions, kwargs, n_kept):
        if self._score_is_overridden(KeyDiffPress):
            return None
        k_out, v_out, _, _ = native.keydiff_compress(keys, values, n_kept)
        return k_out, v_out
