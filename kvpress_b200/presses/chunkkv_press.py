"""ChunkKVPress: keep or drop whole chunks of consecutive positions (semantic chunks), ranked by the wrapped
press's scores summed over heads and averaged over the chunk (https://arxiv.org/abs/2502.00299).

API mirror of `/root/reference/kvpress/presses/chunkkv_press.py:17-117`. The chunk ranking is a few tiny
device ops on the `[B, Hkv, S]` score tensor; the compaction is the shared sm_100a select+compact path, driven by a
0/1 score row that marks the kept chunks (exactly n_kept ones, so the selection has no ties to break). Like the
reference, the chunk choice of batch element 0 is applied to the whole batch, and rows come out in position order.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
from torch import nn

from kvpress_b200 import native
from kvpress_b200.presses.base_press import BasePress
from kvpress_b200.presses.scorer_press import ScorerPress


@dataclass
class ChunkKVPress(BasePress):
    press: ScorerPress
    chunk_length: int = 20

    def __post_init__(self):
        assert isinstance(self.press, ScorerPress), "ChunkKVPress requires a ScorerPress as input"

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
        ratio = self.press.compression_ratio
        if ratio == 0:
            return keys, values
        assert attentions is None, "ChunkPress does not support attentions."
        B, H, S, _ = keys.shape
        L = self.chunk_length
        n_full, tail = divmod(S, L)
        if n_full == 0:
            return self.press.compress(module, hidden_states, keys, values, attentions, kwargs)
        scores = self.press.score(module, hidden_states, keys, values, attentions, kwargs)
        per_pos = scores.sum(dim=1)                                                   # [B, S], heads summed
        chunk_scores = per_pos[:, : n_full * L].view(B, n_full, L).mean(dim=-1)
        if tail > 0:
            chunk_scores = torch.cat([chunk_scores, per_pos[:, n_full * L:].mean(dim=-1, keepdim=True)], dim=-1)
        n_chunks = n_full + (tail > 0)
        n_chunks_kept = max(1, int(n_chunks * (1 - ratio)))
        top = chunk_scores.topk(n_chunks_kept, dim=-1).indices[0]                     # batch 0 decides (as upstream)
        chunk_mask = torch.zeros(n_chunks, dtype=scores.dtype, device=scores.device)
        chunk_mask[top] = 1
        mask = chunk_mask[:n_full].repeat_interleave(L)
        tail_kept = False
        if tail > 0:
            tail_kept = bool(chunk_mask[n_full].item())                               # the one host sync of this press
            mask = torch.cat([mask, chunk_mask[n_full:].repeat_interleave(tail)])
        n_kept = n_chunks_kept * L - ((L - tail) if tail_kept else 0)
        k_out, v_out, _ = native.scores_compress(mask.expand(B, H, S), keys, values, n_kept)
        return k_out, v_out
