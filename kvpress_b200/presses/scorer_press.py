"""ScorerPress: score every cached position, keep the n_kept best per head, compact K and V.

API mirror of `/root/reference/kvpress/presses/scorer_press.py:16-102`. The ATen sequence
`topk -> expand -> gather(K) -> gather(V)` (:95-100) is replaced by the sm_100a select+compact
kernels; the four in-scope scorers additionally fuse their `score()` into the same pass (see the
subclasses). `score()` stays a public, standalone method returning a `[B, Hkv, S]` tensor because
wrapper presses call it directly, and `compression_ratio` stays a plain mutable attribute because
wrappers (DecodingPress) set and restore it.

Output rows are in ascending position order (the reference emits score-descending order; attention
is permutation-invariant over cache rows). Ties at the threshold score go to the lowest positions.
"""
from __future__ import annotations

import logging
from dataclasses import dataclass

import torch
from torch import nn

from kvpress_b200 import native
from kvpress_b200.presses.base_press import BasePress

logger = logging.getLogger(__name__)


def kept_count(k_len: int, compression_ratio: float) -> int:
    """Same float64 arithmetic as the reference (scorer_press.py:94): e.g. int(131072*(1-0.7)) == 39321."""
    return int(k_len * (1 - compression_ratio))


@dataclass
class ScorerPress(BasePress):
    """Prunes the `compression_ratio` fraction of positions with the lowest `score`."""

    compression_ratio: float = 0.0

    # Set to False by presses whose score() ignores hidden_states: DecodingPress then skips buffering.
    needs_hidden_states = True

    def __post_init__(self):
        assert 0 <= self.compression_ratio < 1, "Compression ratio must be between 0 and 1"

    def score(
        self,
        module: nn.Module,
        hidden_states: torch.Tensor,
        keys: torch.Tensor,
        values: torch.Tensor,
        attentions: torch.Tensor,
        kwargs,
    ) -> torch.Tensor:
        """Importance of every cached position, shape [B, Hkv, S]; higher = kept longer."""
        raise NotImplementedError

    def _score_is_overridden(self, owner: type) -> bool:
        return type(self).score is not owner.score

    def _n_kept(self, module: nn.Module, k_len: int) -> int:
        """Positions to keep for this layer (scorer_press.py:93-94); PyramidKV overrides it per layer."""
        return kept_count(k_len, self.compression_ratio)

    def _fused_compress(self, module, hidden_states, keys, values, attentions, kwargs, n_kept):
        """Scorer-specific fused score+select+compact; None means 'use score() + generic select'."""
        return None

    def compress(
        self,
        module: nn.Module,
        hidden_states: torch.Tensor,
        keys: torch.Tensor,
        values: torch.Tensor,
        attentions: torch.Tensor,
        kwargs: dict,
    ) -> tuple[torch.Tensor, torch.Tensor]:
        if self.compression_ratio == 0:
            return keys, values
        n_kept = self._n_kept(module, keys.shape[2])
        fused = self._fused_compress(module, hidden_states, keys, values, attentions, kwargs, n_kept)
        if fused is not None:
            return fused
        scores = self.score(module, hidden_states, keys, values, attentions, kwargs)
        k_out, v_out, _ = native.scores_compress(scores, keys, values, n_kept)
        return k_out, v_out
