"""BlockPress: block-wise iterative compression — the context is consumed `block_size` positions at a time and the
kept set is re-selected after every block, so memory stays bounded (KeyDiff's streaming setting).

API mirror of `/root/reference/kvpress/presses/block_press.py:16-98`. Each round scores [kept so far + new block]
with the wrapped press and re-selects n_kept of them with the sm_100a selection kernel (`kvp_scores_select`); the
final compaction is the shared select+compact path driven by a 0/1 row that marks the survivors. Kept positions stay
in ascending position order through every round.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
from torch import nn

from kvpress_b200 import native
from kvpress_b200.presses.base_press import BasePress
from kvpress_b200.presses.scorer_press import ScorerPress, kept_count


def _gather_positions(x: torch.Tensor, positions: torch.Tensor) -> torch.Tensor:
    return x.gather(2, positions.unsqueeze(-1).expand(-1, -1, -1, x.shape[-1]))


@dataclass
class BlockPress(BasePress):
    press: ScorerPress
    block_size: int = 128

    def __post_init__(self):
        assert isinstance(self.press, ScorerPress), "BlockPress requires a ScorerPress"

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
        assert attentions is None, "BlockPress does not support attentions."
        B, H, S, _ = keys.shape
        block = min(self.block_size, S)
        n_kept = kept_count(S, self.compression_ratio)
        kept = torch.arange(n_kept, device=keys.device).expand(B, H, -1)
        # hidden states follow the per-head position lists: view them as [B, H, S, hidden / H]
        states = hidden_states.view(B, S, H, -1).transpose(1, 2)
        for lo in range(n_kept, S, block):
            hi = min(lo + block, S)
            current = torch.cat([kept, torch.arange(lo, hi, device=keys.device).expand(B, H, -1)], dim=-1)
            cur_states = _gather_positions(states, current).transpose(1, 2).reshape(B, -1, hidden_states.shape[-1])
            scores = self.press.score(module, cur_states, _gather_positions(keys, current),
                                      _gather_positions(values, current), attentions, kwargs)
            kept = current.gather(-1, native.scores_select(scores, n_kept).long())
        mask = torch.zeros((B, H, S), dtype=keys.dtype, device=keys.device).scatter_(-1, kept, 1)
        k_out, v_out, _ = native.scores_compress(mask, keys, values, n_kept)
        return k_out, v_out
