"""Head-wise masking for presses that prune a different number of positions per head (AdaKV).

API mirror of `/root/reference/kvpress/attention_patch.py:8-110`. A head-wise press cannot shrink the
[B, Hkv, S, D] cache (heads keep different lengths), so it records the pruned (batch, head, position) triples
in `module.masked_key_indices`, and every attention function of `transformers` is wrapped so that, while
decoding, those keys are overwritten with a "fake" key k with exp(q.k) = 0 for every current query q.
The wrapper is a no-op for modules that never set `masked_key_indices`; it also repairs `cu_seq_lens_k`
after a cache has been shortened (flash-attention varlen path).
"""
from __future__ import annotations

import torch
from transformers.modeling_utils import ALL_ATTENTION_FUNCTIONS

_PATCH_FLAG = "_kvpress_b200_patched"


def search_hyperplane(X: torch.Tensor, max_iter: int = 1000) -> torch.Tensor:
    """For X [n, m, d], a vector Y [n, d] with <X[i, j], Y[i]> <= -1e5 for all j (attention_patch.py:8-40):
    perceptron-style search for a direction with positive projection on every query, flipped and scaled."""
    Y = X.mean(1)
    for _ in range(max_iter):
        violated = torch.bmm(X, Y.unsqueeze(-1)) <= 0          # [n, m, 1]
        if not violated.any():
            return -1e5 * Y / Y.norm(dim=-1, keepdim=True) ** 2
        Y = Y + (X * violated).sum(1) / violated.sum(1).clamp(min=1)
    raise ValueError("Could not find fake keys such that for every query q, exp(<q, k>) = 0")


def attention_patch(func):
    """Wrap one attention function (attention_patch.py:43-87)."""
    if getattr(func, _PATCH_FLAG, False):
        return func

    def wrapper(module, query, key, value, attention_mask, dropout, **kwargs):
        if query.shape[2] == key.shape[2]:
            module.masked_key_indices = None                     # prefill: nothing is masked yet
        elif getattr(module, "masked_key_indices", None) is not None:
            bsz, num_heads, q_len, head_dim = query.shape
            num_kv_heads = key.shape[1]
            groups = num_heads // num_kv_heads
            q = query.view(bsz, num_kv_heads, groups, q_len, head_dim).reshape(bsz * num_kv_heads, groups * q_len, head_dim)
            fake = search_hyperplane(q).view(bsz, num_kv_heads, head_dim)
            b_idx, h_idx, s_idx = module.masked_key_indices
            key[b_idx, h_idx, s_idx] = fake[b_idx, h_idx].to(key.dtype)
        if "cu_seq_lens_k" in kwargs:
            kwargs["cu_seq_lens_k"][-1] = key.shape[-2]
        return func(module, query, key, value, attention_mask, dropout, **kwargs)

    setattr(wrapper, _PATCH_FLAG, True)
    return wrapper


def patch_attention_functions():
    """Wrap every registered attention function once (attention_patch.py:90-110); idempotent."""
    for name, func in list(ALL_ATTENTION_FUNCTIONS.items()):
        ALL_ATTENTION_FUNCTIONS[name] = attention_patch(func)
