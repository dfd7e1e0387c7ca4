"""AdaKVPress: head-wise budgets — the same total number of positions is pruned, but across all kv-heads of a
layer at once, so heads with flat scores give up more than heads with peaked scores.

API mirror of `/root/reference/kvpress/presses/adakv_press.py:16-78`. Both selections run in the sm_100a
selection kernel (`kvp_scores_select`): (1) the n_safe best positions of every head are protected, (2) the
n_pruned lowest scores of the head-flattened [B, 1, Hkv*S] row are pruned. The cache keeps its shape; the
pruned (batch, head, position) triples go to `module.masked_key_indices` for `attention_patch.py`.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
from torch import nn

from kvpress_b200 import native
from kvpress_b200.attention_patch import patch_attention_functions
from kvpress_b200.presses.base_press import BasePress
from kvpress_b200.presses.scorer_press import ScorerPress, kept_count


@dataclass
class AdaKVPress(BasePress):
    press: ScorerPress
    alpha_safeguard: float = 0.20

    def __post_init__(self):
        assert isinstance(self.press, ScorerPress), "AdaKVPress requires a ScorerPress as input"
        assert 0 <= self.alpha_safeguard <= 1, "alpha_safeguard should be in [0, 1]"
        patch_attention_functions()

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
        if self.compression_ratio == 0:
            return keys, values
        assert module.config._attn_implementation != "eager", "eager mode not supported"
        scores = self.press.score(module, hidden_states, keys, values, attentions, kwargs)
        B, H, S = scores.shape
        n_kept = kept_count(S, self.compression_ratio)
        n_safe = int(n_kept * self.alpha_safeguard)
        if n_safe > 0:  # every head keeps at least its n_safe best positions
            protected = native.scores_select(scores, n_safe)
            scores = scores.scatter(-1, protected.long(), torch.finfo(scores.dtype).max)
        n_pruned = H * (S - n_kept)
        flat = native.scores_select((-scores).reshape(B, 1, H * S), n_pruned).reshape(-1).long()
        batch = torch.arange(B, device=flat.device).repeat_interleave(n_pruned)
        module.masked_key_indices = (batch, flat // S, flat % S)
        return keys, values
