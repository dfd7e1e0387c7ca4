"""StreamingLLMPress: keep the first `n_sink` positions and the most recent ones
(https://arxiv.org/abs/2309.17453).

API mirror of `/root/reference/kvpress/presses/streaming_llm_press.py:16-54`. The reference builds
a 0/1 score tensor and runs a generic top-k over it; the kept set is two contiguous ranges, so
`compress` is a single range-copy kernel. `score()` still returns the 0/1 tensor for wrappers.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
from torch import nn

from kvpress_b200 import native
from kvpress_b200.presses.scorer_press import ScorerPress, kept_count


@dataclass
class StreamingLLMPress(ScorerPress):
    """Sliding window with attention sinks."""

    compression_ratio: float = 0.0
    n_sink: int = 4

    needs_hidden_states = False

    def _check(self, k_len: int):
        assert k_len > self.n_sink, f"Input should contain more tokens than n_sink={self.n_sink}"

    def score(self, module: nn.Module, hidden_states, keys: torch.Tensor, values, attentions, kwargs) -> torch.Tensor:
        k_len = keys.shape[2]
        self._check(k_len)
        return native.streaming_score(keys, kept_count(k_len, self.compression_ratio), self.n_sink)

    def _fused_compress(self, module, hidden_states, keys, values, attentions, kwargs, n_kept):
        if self._score_is_overridden(StreamingLLMPress):
            return None
        self._check(keys.shape[2])
        k_out, v_out, _ = native.streaming_compress(keys, values, n_kept, self.n_sink)
        return k_out, v_out
