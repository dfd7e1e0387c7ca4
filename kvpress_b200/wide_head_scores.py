"""Scores for cache shapes the two tensor-core scorers do not instantiate (GPU, torch / cuBLAS GEMMs).

The tcgen05 SnapKV and ExpectedAttention kernels keep their query block / covariance resident in shared memory
and exist for head_dim 64 and 128, at most 8 query heads per kv head and (for SnapKV) at most 512 window-query rows
per kv head. Phi-3 (head_dim 96) and Gemma-3 (head_dim 256) fall outside. For those shapes only the SCORE stage
is evaluated here with plain library GEMMs on the GPU, in fp32 with one rounding to the cache dtype — the same
definition the kernels implement ("fp32 evaluation of the reference formula, rounded once") — and the
result goes through the sm_100a selection + compaction kernels like any caller-supplied score tensor
(`native.scores_compress`). Nothing here runs on the CPU; CPU tensors are refused like everywhere else.

Formulas: SnapKV `/root/reference/kvpress/presses/snapkv_press.py:41-105`, ExpectedAttention
`/root/reference/kvpress/presses/expected_attention_press.py:136-165`.
"""
from __future__ import annotations

import logging
import math

import torch
import torch.nn.functional as F

logger = logging.getLogger(__name__)
_warned: set = set()

TENSOR_CORE_HEAD_DIMS = (64, 128)
MAX_GROUP = 8
MAX_SNAP_QUERY_ROWS = 512


def snapkv_on_tensor_cores(head_dim: int, group: int, window: int) -> bool:
    """Shapes `kvp_snapkv_*` instantiates (csrc/snapkv.cu launch_snap_d)."""
    return head_dim in TENSOR_CORE_HEAD_DIMS and 1 <= group * window <= MAX_SNAP_QUERY_ROWS


def expected_attention_on_tensor_cores(head_dim: int, group: int, has_cov: bool) -> bool:
    """Shapes `kvp_expected_attention_*` instantiates (csrc/expected_attention.cu launch_ea_t): the covariance-free
    scan takes any head_dim, the covariance path head_dim 64 / 128; both at most 8 query heads per kv head."""
    return group <= MAX_GROUP and (not has_cov or head_dim in TENSOR_CORE_HEAD_DIMS)


def _note(what: str, keys: torch.Tensor, group: int) -> None:
    key = (what, keys.shape[-1], group)
    if key not in _warned:
        _warned.add(key)
        logger.warning(f"{what}: head_dim {keys.shape[-1]} / {group} query heads per kv head is outside the tcgen05 "
                       "scorer instantiations; scoring with cuBLAS GEMMs, selection + compaction stay on the "
                       "sm_100a kernels")


def _require_cuda(keys: torch.Tensor) -> None:
    if not keys.is_cuda:
        raise RuntimeError(f"kvpress_b200 runs on CUDA tensors only; got keys on {keys.device}. There is no CPU path.")


def _pad_forced(scores: torch.Tensor, n_front: int, n_back: int) -> torch.Tensor:
    """Forced-keep positions carry max + 1 over the whole tensor (snapkv_press.py:103, expected_attention_press.py:163),
    formed on the device (no host sync) and rounded once to the score dtype."""
    sentinel = (scores.max().float() + 1).to(scores.dtype)
    B, H, _ = scores.shape
    parts = []
    if n_front:
        parts.append(sentinel.expand(B, H, n_front))
    parts.append(scores)
    if n_back:
        parts.append(sentinel.expand(B, H, n_back))
    return torch.cat(parts, dim=-1)


def snapkv_scores(keys: torch.Tensor, q_window: torch.Tensor, window: int, kernel_size: int,
                  chunk: int = 16384) -> torch.Tensor:
    """keys [B,Hkv,S,D], RoPE'd window queries [B,Hq,w,D] -> scores [B,Hkv,S] in the cache dtype."""
    _require_cuda(keys)
    B, Hkv, S, D = keys.shape
    G, w = q_window.shape[1] // Hkv, window
    _note("SnapKV", keys, G)
    q = q_window.float().reshape(B, Hkv, G * w, D) * (1.0 / math.sqrt(D))
    logits = torch.empty((B, Hkv, G * w, S), dtype=torch.float32, device=keys.device)
    for s0 in range(0, S, chunk):  # fp32 copies of K one slab at a time
        s1 = min(S, s0 + chunk)
        torch.matmul(q, keys[:, :, s0:s1].float().transpose(-1, -2), out=logits[..., s0:s1])
    # inside the window query i (position S - w + i) does not see key S - w + j for j > i
    future = torch.ones((w, w), dtype=torch.bool, device=keys.device).triu(1)
    logits.view(B, Hkv, G, w, S)[..., S - w:].masked_fill_(future, float("-inf"))
    prob = torch.softmax(logits, dim=-1)
    del logits
    col = prob[..., : S - w].mean(dim=2)                               # mean over the window and the group
    col = F.avg_pool1d(col, kernel_size=kernel_size, padding=kernel_size // 2, stride=1)  # zero pad, always / kernel
    return _pad_forced(col.to(keys.dtype), 0, w)


def expected_attention_scores(keys: torch.Tensor, values: torch.Tensor, mu: torch.Tensor, cov, epsilon: float,
                              n_sink: int, use_vnorm: bool, chunk: int = 4096) -> torch.Tensor:
    """keys/values [B,Hkv,S,D], mu [B,Hq,D], cov [B,Hq,D,D] or None -> scores [B,Hkv,S] in the cache dtype."""
    _require_cuda(keys)
    B, Hkv, S, D = keys.shape
    G = mu.shape[1] // Hkv
    _note("ExpectedAttention", keys, G)
    L = S - n_sink
    mu_f = mu.to(keys.dtype).float().reshape(B, Hkv, G, D) * (1.0 / math.sqrt(D))   # operands in the cache dtype,
    cov_f = None if cov is None else cov.to(keys.dtype).float().reshape(B, Hkv, G, D, D) * (0.5 / D)  # like the C ABI
    logits = torch.empty((B, Hkv, G, L), dtype=torch.float32, device=keys.device)
    for s0 in range(0, L, chunk):
        s1 = min(L, s0 + chunk)
        k = keys[:, :, n_sink + s0:n_sink + s1].float()               # [B,Hkv,c,D]
        part = torch.matmul(mu_f, k.transpose(-1, -2))                 # mu.k / sqrt(d)
        if cov_f is not None:                                          # + k^T Sigma k / 2d
            y = torch.matmul(k.unsqueeze(2), cov_f)                    # [B,Hkv,G,c,D]
            part = part + (y * k.unsqueeze(2)).sum(dim=-1)
        logits[..., s0:s1] = part
    score = torch.softmax(logits, dim=-1).mean(dim=2)                  # [B,Hkv,L]
    if use_vnorm:
        vn = torch.empty((B, Hkv, L), dtype=torch.float32, device=keys.device)
        for s0 in range(0, L, 4 * chunk):
            s1 = min(L, s0 + 4 * chunk)
            vn[..., s0:s1] = values[:, :, n_sink + s0:n_sink + s1].float().norm(dim=-1)
        score = (score + epsilon) * vn
    return _pad_forced(score.to(keys.dtype), n_sink, 0)
