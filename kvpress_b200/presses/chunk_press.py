"""ChunkPress: the wrapped ScorerPress is applied independently to consecutive chunks of the context, so
every chunk is compressed by the same ratio. API mirror of `/root/reference/kvpress/presses/chunk_press.py:16-87`.

Selection + compaction of ALL full-length chunks is one `kvp_scores_compress` call: a contiguous cache
[B, Hkv, n*L, D] is the same memory as [B, Hkv*n, L, D], i.e. every chunk is its own row of the per-row
top-k, and the compacted [B, Hkv*n, n_kept, D] is exactly the concatenation the reference builds. (expressed as the strided view [B*Hkv, n, L, D], so a ragged last chunk does not force a copy; it is a second call). Rows inside a chunk come out in ascending position order.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
from torch import nn

from kvpress_b200 import native
from kvpress_b200.presses.base_press import BasePress
from kvpress_b200.presses.scorer_press import ScorerPress


def _as_chunk_rows(x: torch.Tensor, n_full: int, L: int):
    """[B, H, S, D] -> the view [B*H, n_full, L, D] (chunk i of head (b, h) = row (b*H + h, i)), or None when
    batch and head cannot be flattened into a single stride."""
    B, H, _, D = x.shape
    if x.stride(3) != 1 or (B > 1 and H > 1 and x.stride(0) != H * x.stride(1)):
        return None
    row_stride = x.stride(1) if H > 1 else x.stride(0)
    return x.as_strided((B * H, n_full, L, D), (row_stride, L * x.stride(2), x.stride(2), 1))


@dataclass
class ChunkPress(BasePress):
    press: ScorerPress
    chunk_length: int = 1024

    def __post_init__(self):
        assert isinstance(self.press, ScorerPress), "ChunkPress requires a ScorerPress as input"

    def post_init_from_model(self, model):
        self.press.post_init_from_model(model)

    @property
    def compression_ratio(self):
        return self.press.compression_ratio

    @compression_ratio.setter
    def compression_ratio(self, value):
        self.press.compression_ratio = value

    def _chunk_scores(self, module, hidden_states, keys, values, kwargs, lo: int, hi: int) -> torch.Tensor:
        return self.press.score(module, hidden_states[:, lo:hi], keys[:, :, lo:hi], values[:, :, lo:hi], None, kwargs)

    def compress(self, module: nn.Module, hidden_states, keys: torch.Tensor, values: torch.Tensor, attentions,
                 kwargs: dict) -> tuple[torch.Tensor, torch.Tensor]:
        ratio = self.press.compression_ratio
        if ratio == 0:
            return keys, values
        assert attentions is None, "ChunkPress does not support attentions."
        B, H, S, D = keys.shape
        L = self.chunk_length
        n_full, tail = divmod(S, L)
        outs_k, outs_v = [], []
        if n_full > 0:
            n_kept = max(1, int(L * (1 - ratio)))
            scores = torch.stack([self._chunk_scores(module, hidden_states, keys, values, kwargs, i * L, (i + 1) * L)
                                  for i in range(n_full)], dim=2)                       # [B, H, n_full, L]
            k_rows, v_rows = _as_chunk_rows(keys, n_full, L), _as_chunk_rows(values, n_full, L)
            if k_rows is not None and v_rows is not None:
                k2, v2, _ = native.scores_compress(scores.reshape(B * H, n_full, L), k_rows, v_rows, n_kept)
                outs_k.append(k2.view(B, H, n_full * n_kept, D))
                outs_v.append(v2.view(B, H, n_full * n_kept, D))
            else:  # (b, h) not flattenable into one stride: one call per chunk
                for i in range(n_full):
                    k2, v2, _ = native.scores_compress(scores[:, :, i], keys[:, :, i * L:(i + 1) * L],
                                                       values[:, :, i * L:(i + 1) * L], n_kept)
                    outs_k.append(k2)
                    outs_v.append(v2)
        if tail > 0:
            n_kept = max(1, int(tail * (1 - ratio)))
            scores = self._chunk_scores(module, hidden_states, keys, values, kwargs, n_full * L, S)
            k2, v2, _ = native.scores_compress(scores, keys[:, :, n_full * L:], values[:, :, n_full * L:], n_kept)
            outs_k.append(k2)
            outs_v.append(v2)
        if len(outs_k) == 1:
            return outs_k[0], outs_v[0]
        return torch.cat(outs_k, dim=2), torch.cat(outs_v, dim=2)
