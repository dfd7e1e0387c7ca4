"""RandomPress: uniform random scores (a baseline). API mirror of
`/root/reference/kvpress/presses/random_press.py:17-46`; selection + compaction run in the generic
`kvp_scores_compress` path. The optional seed drives a generator on the cache's device."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
from torch import nn

from kvpress_b200.presses.scorer_press import ScorerPress


@dataclass
class RandomPress(ScorerPress):
    compression_ratio: float = 0.0
    seed: Optional[int] = None

    needs_hidden_states = False

    def score(self, module: nn.Module, hidden_states, keys: torch.Tensor, values, attentions, kwargs) -> torch.Tensor:
        generator = None
        if self.seed is not None:
            generator = torch.Generator(device=keys.device)
            generator.manual_seed(self.seed)
        return torch.rand(*keys.shape[:-1], generator=generator, device=keys.device, dtype=keys.dtype)
