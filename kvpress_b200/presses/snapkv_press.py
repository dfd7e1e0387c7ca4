"""SnapKVPress: importance = attention the last `window_size` queries pay to each earlier key
(https://arxiv.org/abs/2404.14469).

API mirror of `/root/reference/kvpress/presses/snapkv_press.py:14-105`. Host prologue (torch):
the 64-row `q_proj` GEMM + RoPE of the window queries (:53-58). Everything that touches the cache
— QK^T over all S keys for the Hq/Hkv query heads of each kv head, the causal mask inside the
window, the exact softmax normalisers, mean over the window, avg_pool1d, group mean, forced keep of
the window (:61-67, :95-103) — runs in the sm_100a kernel, which reads each K tile once per pass
instead of materialising repeat_kv / [B,Hq,w,S] logits / an fp32 softmax temp.

`attentions` (eager attention weights) is accepted for signature compatibility and ignored: the
kernel recomputes the same window attention from Q and K.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
from torch import nn

from kvpress_b200 import native, wide_head_scores
from kvpress_b200.presses.scorer_press import ScorerPress
from kvpress_b200.utils import apply_rope, get_prerope_query_states


@dataclass
class SnapKVPress(ScorerPress):
    compression_ratio: float = 0.0
    window_size: int = 64
    kernel_size: int = 5

    def window_queries(self, module: nn.Module, hidden_states: torch.Tensor, kwargs: dict) -> torch.Tensor:
        """RoPE'd queries of the last `window_size` positions, [B, Hq, w, D]."""
        w = self.window_size
        assert hidden_states.shape[1] > w, (
            f"Query length {hidden_states.shape[1]} should be greater than the window size {w}"
        )
        q = get_prerope_query_states(module, hidden_states[:, -w:])
        cos, sin = kwargs["position_embeddings"]
        return apply_rope(q, cos[:, -w:], sin[:, -w:]).contiguous()

    def score(self, module: nn.Module, hidden_states, keys: torch.Tensor, values, attentions, kwargs) -> torch.Tensor:
        q_window = self.window_queries(module, hidden_states, kwargs)
        if not self._on_tensor_cores(keys, q_window):
            return wide_head_scores.snapkv_scores(keys, q_window, self.window_size, self.kernel_size)
        return native.snapkv_score(keys, q_window, self.window_size, self.kernel_size)

    def _on_tensor_cores(self, keys: torch.Tensor, q_window: torch.Tensor) -> bool:
        """False for shapes the tcgen05 kernels do not instantiate (head_dim other than 64 / 128, more than 512
        window-query rows per kv head): their score stage runs on cuBLAS GEMMs (wide_head_scores.py)."""
        return wide_head_scores.snapkv_on_tensor_cores(keys.shape[-1], q_window.shape[1] // keys.shape[1],
                                                       self.window_size)

    def _fused_compress(self, module, hidden_states, keys, values, attentions, kwargs, n_kept):
        if self._score_is_overridden(SnapKVPress):
            return None
        q_window = self.window_queries(module, hidden_states, kwargs)
        if not self._on_tensor_cores(keys, q_window):
            scores = wide_head_scores.snapkv_scores(keys, q_window, self.window_size, self.kernel_size)
            return native.scores_compress(scores, keys, values, n_kept)[:2]
        k_out, v_out, _, _ = native.snapkv_compress(keys, values, q_window, self.window_size, self.kernel_size, n_kept)
        return k_out, v_out
