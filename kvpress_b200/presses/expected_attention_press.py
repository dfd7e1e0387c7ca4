"""ExpectedAttentionPress: expected attention future queries will pay to each cached key.

API mirror of `/root/reference/kvpress/presses/expected_attention_press.py:16-165`.

Host prologue (torch / cuBLAS, dense GEMMs that never touch the cache): pre-RoPE queries of the
prompt, their mean and covariance per head, and the average RoPE rotation over the next
`n_future_positions` (:62-124). Cache scan (sm_100a kernel): for every kv head one pass over K and V
computing mu.k/sqrt(d) + k^T Sigma k / (2d) for the Hq/Hkv query heads, softmax over positions,
group mean, (+epsilon) * ||v||, forced keep of the `n_sink` first positions (:138-163) — instead of
materialising repeat_kv(K)^T twice and a [B,Hq,D,S] einsum intermediate.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
from torch import nn

from kvpress_b200 import native, wide_head_scores
from kvpress_b200.presses.scorer_press import ScorerPress
from kvpress_b200.utils import get_prerope_query_states


@dataclass
class ExpectedAttentionPress(ScorerPress):
    compression_ratio: float = 0.0
    n_future_positions: int = 512
    n_sink: int = 4
    use_covariance: bool = True
    use_vnorm: bool = True
    epsilon: float = 0.0

    def get_query_statistics(self, module: nn.Module, hidden_states: torch.Tensor):
        """Mean [B,Hq,D] and covariance [B,Hq,D,D] of the pre-RoPE queries, rotated by the average RoPE."""
        q_len = hidden_states.shape[1]
        h = hidden_states[:, self.n_sink:]  # the first positions are outliers
        q = get_prerope_query_states(module, h)
        mu = q.mean(dim=2, keepdim=True)
        cov = None
        if self.use_covariance:
            centred = q - mu
            cov = torch.einsum("bnsi,bnsj->bnij", centred, centred) / h.shape[1]
        return self.apply_avg_rope(module, mu.squeeze(2), cov, q_len)

    def apply_avg_rope(self, module: nn.Module, mu: torch.Tensor, cov, q_len: int):
        """mu <- mu R^T, cov <- R cov R^T with R the mean RoPE matrix of positions q_len .. q_len+n_future-1."""
        d = module.head_dim
        positions = torch.arange(q_len, q_len + self.n_future_positions, device=mu.device).unsqueeze(0)
        cos, sin = module.rotary_emb(mu, positions)
        cos, sin = cos[0], sin[0]  # [n_future, D]
        eye = torch.eye(d, device=cos.device, dtype=cos.dtype)
        swap = torch.zeros((d, d), device=cos.device, dtype=cos.dtype)  # rotate_half as a matrix
        half = d // 2
        swap[half:, :half] = torch.eye(half, device=cos.device, dtype=cos.dtype)
        swap[:half, half:] = -torch.eye(half, device=cos.device, dtype=cos.dtype)
        R = (cos.unsqueeze(1) * eye + sin.unsqueeze(1) * swap).mean(dim=0).to(mu.device)
        mu = torch.matmul(mu, R.T)
        if cov is not None:
            cov = torch.matmul(R, torch.matmul(cov, R.T))
        return mu, cov

    def score(self, module: nn.Module, hidden_states, keys: torch.Tensor, values, attentions, kwargs) -> torch.Tensor:
        assert keys.size(2) > self.n_sink, f"Input should contain more tokens than n_sink={self.n_sink}"
        mu, cov = self.get_query_statistics(module, hidden_states)
        if not self._on_tensor_cores(keys, mu, cov):
            return wide_head_scores.expected_attention_scores(keys, values, mu, cov, self.epsilon, self.n_sink,
                                                              self.use_vnorm)
        return native.expected_attention_score(keys, values, mu, cov, self.epsilon, self.n_sink, self.use_vnorm)

    @staticmethod
    def _on_tensor_cores(keys: torch.Tensor, mu: torch.Tensor, cov) -> bool:
        """False for shapes the sm_100a scan does not instantiate (covariance with head_dim other than 64 / 128, more
        than 8 query heads per kv head): their score stage runs on cuBLAS GEMMs (wide_head_scores.py)."""
        return wide_head_scores.expected_attention_on_tensor_cores(keys.shape[-1], mu.shape[1] // keys.shape[1],
                                                                   cov is not None)

    def _fused_compress(self, module, hidden_states, keys, values, attentions, kwargs, n_kept):
        if self._score_is_overridden(ExpectedAttentionPress):
            return None
        assert keys.size(2) > self.n_sink, f"Input should contain more tokens than n_sink={self.n_sink}"
        mu, cov = self.get_query_statistics(module, hidden_states)
        if not self._on_tensor_cores(keys, mu, cov):
            scores = wide_head_scores.expected_attention_scores(keys, values, mu, cov, self.epsilon, self.n_sink,
                                                                self.use_vnorm)
            return native.scores_compress(scores, keys, values, n_kept)[:2]
        k_out, v_out, _, _ = native.expected_attention_compress(
            keys, values, mu, cov, self.epsilon, self.n_sink, self.use_vnorm, n_kept)
        return k_out, v_out
