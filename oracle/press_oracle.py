"""CPU oracle for the kvpress score -> top-k -> gather hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain torch-CPU tensor code, the algorithm of the reference's
`ScorerPress.compress` and of the four in-scope `score()` methods. It exists to CHECK the sm_100a
kernels and to serve as the timed CPU baseline; nothing under `kvpress_b200/` imports it (only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs do).

Why torch and not C/numpy: the reference's arithmetic IS a sequence of ATen calls on bf16 tensors,
and its results depend on where ATen rounds to bf16 (after the norm, after QK^T, after the softmax,
after each mean / pool ...). Restating it with the same dtype at the same points is the only way to
reproduce its score tensors; each function below says where those rounding points are.

Pinned against the reference itself: `tests/golden/make_golden.py` imports the unmodified reference
from /root/reference on CPU and stores its outputs for seeded inputs in `tests/golden/*.npz`;
`tests/test_oracle_golden.py` checks every function here against those files.

Reference citations are `/root/reference/kvpress/...` file:line.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------------
# scorer_press.py:93-100 — n_kept, top-k, gather
# --------------------------------------------------------------------------------------------------
def kept_count(k_len: int, compression_ratio: float) -> int:
    """scorer_press.py:94 — Python float64 arithmetic then truncation."""
    return int(k_len * (1 - compression_ratio))


def topk_indices(scores: torch.Tensor, n_kept: int) -> torch.Tensor:
    """scorer_press.py:95 — indices of the n_kept largest scores per (b, h) row (score-descending)."""
    return scores.topk(n_kept, dim=-1).indices


def gather_rows(x: torch.Tensor, indices: torch.Tensor) -> torch.Tensor:
    """scorer_press.py:96-100 — x[b, h, indices[b, h, i], :]."""
    idx = indices.unsqueeze(-1).expand(-1, -1, -1, x.shape[-1])
    return x.gather(2, idx).contiguous()


def compress_with_scores(scores, keys, values, n_kept):
    """scorer_press.py:93-100 with the scores already computed. Returns (K', V', indices)."""
    idx = topk_indices(scores, n_kept)
    return gather_rows(keys, idx), gather_rows(values, idx), idx


def select_lowest_index_ties(scores: torch.Tensor, n_kept: int) -> torch.Tensor:
    """The kernels' deterministic selection rule: the n_kept largest scores, ties at the threshold
    resolved towards the LOWEST position; returned ascending. (torch.topk leaves ties unspecified,
    so this is the canonical member of the set of valid answers.)  -0.0 == +0.0, NaN is largest."""
    s = scores.float()
    s = torch.where(s == 0, torch.zeros_like(s), s)  # -0.0 -> +0.0
    bits = s.contiguous().view(torch.int32).to(torch.int64)
    key = torch.where(bits < 0, -(bits & 0x7FFFFFFF), bits)  # monotone in the float value
    key = torch.where(torch.isnan(s), torch.full_like(key, 2 ** 40), key)  # NaN above +inf (torch.topk order)
    # stable descending sort keeps equal scores in ascending position order
    order = torch.sort(key, dim=-1, descending=True, stable=True).indices
    return torch.sort(order[..., :n_kept], dim=-1).values


def check_selection(ref_scores: torch.Tensor, kept: torch.Tensor, n_kept: int, ulp_slack: int = 0) -> dict:
    """Tie-aware validity of a kept-index set against reference scores:
    every position scoring strictly above the n_kept-th largest score must be kept, no position
    strictly below it may be kept, |kept| == n_kept, indices unique. `ulp_slack` widens the threshold
    by that many 16-bit ulps for scorers whose scores are only reproduced to rounding error."""
    B, H, S = ref_scores.shape
    s = ref_scores.float()
    thresh = s.topk(n_kept, dim=-1).values[..., -1:]  # [B,H,1]
    if ulp_slack:
        eps = thresh.abs() * (2.0 ** -7) * ulp_slack + 1e-30
    else:
        eps = torch.zeros_like(thresh)
    kept = kept.long()
    mask = torch.zeros((B, H, S), dtype=torch.bool)
    mask.scatter_(2, kept, True)
    unique = bool((mask.sum(-1) == n_kept).all())
    must = s > thresh + eps
    may = s >= thresh - eps
    missing = int((must & ~mask).sum())
    illegal = int((mask & ~may).sum())
    return {"ok": unique and missing == 0 and illegal == 0, "unique": unique, "missing": missing, "illegal": illegal}


# --------------------------------------------------------------------------------------------------
# knorm_press.py:38
# --------------------------------------------------------------------------------------------------
def knorm_scores(keys: torch.Tensor) -> torch.Tensor:
    """-||k||_2 per position. ATen accumulates in fp32 and rounds ONCE to the key dtype."""
    return -keys.norm(dim=-1)


# --------------------------------------------------------------------------------------------------
# streaming_llm_press.py:48-52
# --------------------------------------------------------------------------------------------------
def streaming_scores(keys: torch.Tensor, compression_ratio: float, n_sink: int) -> torch.Tensor:
    k_len = keys.shape[2]
    assert k_len > n_sink
    n_pruned = k_len - kept_count(k_len, compression_ratio)
    scores = torch.ones_like(keys[..., 0])
    scores[:, :, n_sink: n_sink + n_pruned] = 0
    return scores


def streaming_kept(k_len: int, n_kept: int, n_sink: int) -> torch.Tensor:
    """Closed form of topk over the 0/1 scores with the lowest-position tie rule (ascending)."""
    head = min(n_sink, n_kept)
    tail = n_kept - head
    return torch.cat([torch.arange(head), torch.arange(k_len - tail, k_len)])


# --------------------------------------------------------------------------------------------------
# utils.py:12-53 and snapkv_press.py:53-58 — query prologue
# --------------------------------------------------------------------------------------------------
def prerope_queries(hidden_states, q_weight, num_heads, head_dim, q_bias=None, q_norm_weight=None, eps=1e-6):
    """q_proj (+ per-head RMSNorm for Qwen3/Gemma3) -> [B, Hq, L, D]."""
    B, L, _ = hidden_states.shape
    q = F.linear(hidden_states, q_weight, q_bias)
    q = q.view(B, L, num_heads, head_dim).transpose(1, 2)
    if q_norm_weight is not None:  # Qwen3RMSNorm: fp32 variance, cast back, times weight
        qf = q.float()
        qf = qf * torch.rsqrt(qf.pow(2).mean(-1, keepdim=True) + eps)
        q = q_norm_weight * qf.to(q.dtype)
    return q


def rotate_half(x):
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def snapkv_window_queries(hidden_states, q_weight, num_heads, head_dim, cos, sin, window, **kw):
    """snapkv_press.py:53-58 — queries of the last `window` positions with RoPE applied."""
    q = prerope_queries(hidden_states[:, -window:], q_weight, num_heads, head_dim, **kw)
    cos, sin = cos[:, -window:], sin[:, -window:]
    return (q * cos.unsqueeze(1)) + (rotate_half(q) * sin.unsqueeze(1))


def repeat_kv(x: torch.Tensor, n_rep: int) -> torch.Tensor:
    B, H, S, D = x.shape
    if n_rep == 1:
        return x
    return x[:, :, None].expand(B, H, n_rep, S, D).reshape(B, H * n_rep, S, D)


# --------------------------------------------------------------------------------------------------
# snapkv_press.py:41-105
# --------------------------------------------------------------------------------------------------
def snapkv_scores(q_window: torch.Tensor, keys: torch.Tensor, window: int, kernel_size: int) -> torch.Tensor:
    """q_window [B,Hq,w,D] (RoPE'd), keys [B,Hkv,S,D] -> scores [B,Hkv,S] in the key dtype.

    Rounding points in 16-bit mode (all ATen): QK^T output; the /sqrt(D); the mask add; softmax is
    computed in fp32 and cast to 16 bit (:66); mean over the window; avg_pool1d; group mean."""
    B, Hkv, S, D = keys.shape
    Hq = q_window.shape[1]
    G = Hq // Hkv
    k_rep = repeat_kv(keys, G)                                                        # :61
    attn = torch.matmul(q_window, k_rep.transpose(2, 3)) / math.sqrt(D)               # :62
    mask = torch.ones_like(attn) * float("-inf")                                      # :63
    mask = torch.triu(mask, diagonal=S - window + 1)                                  # :64
    attn = attn + mask                                                                # :65
    attn = F.softmax(attn, dim=-1, dtype=torch.float32).to(q_window.dtype)            # :66
    attn = attn[..., :-window]                                                        # :67
    scores = attn.mean(dim=-2)                                                        # :95
    scores = F.avg_pool1d(scores, kernel_size=kernel_size, padding=kernel_size // 2, stride=1)  # :96
    scores = scores.view(B, Hkv, G, S - window).mean(2)                               # :99-100
    return F.pad(scores, (0, window), value=scores.max().item() + 1)                  # :103


def snapkv_scores_fp32(q_window, keys, window: int, kernel_size: int) -> torch.Tensor:
    """Same math evaluated in fp32 throughout (no intermediate 16-bit rounding) — what a fused kernel
    computes before its single final rounding. The forced-keep tail is +inf."""
    B, Hkv, S, D = keys.shape
    Hq = q_window.shape[1]
    G = Hq // Hkv
    q, k = q_window.float(), repeat_kv(keys, G).float()
    attn = torch.matmul(q, k.transpose(2, 3)) / math.sqrt(D)
    mask = torch.triu(torch.full_like(attn, float("-inf")), diagonal=S - window + 1)
    attn = F.softmax(attn + mask, dim=-1)[..., :-window]
    scores = attn.mean(dim=-2)
    scores = F.avg_pool1d(scores, kernel_size=kernel_size, padding=kernel_size // 2, stride=1)
    scores = scores.view(B, Hkv, G, S - window).mean(2)
    return F.pad(scores, (0, window), value=float("inf"))


# --------------------------------------------------------------------------------------------------
# tova_press.py:35-61: attention of the LAST query over all keys, mean over ALL heads
# --------------------------------------------------------------------------------------------------
def window_attention(q_window: torch.Tensor, keys: torch.Tensor, window: int) -> torch.Tensor:
    """snapkv_press.py:60-69 from the RoPE'd window queries on: [B,Hq,w,S-w] attention weights (16-bit)."""
    B, Hkv, S, D = keys.shape
    G = q_window.shape[1] // Hkv
    attn = torch.matmul(q_window, repeat_kv(keys, G).transpose(2, 3)) / math.sqrt(D)
    mask = torch.triu(torch.ones_like(attn) * float("-inf"), diagonal=S - window + 1)
    attn = attn + mask
    attn = F.softmax(attn, dim=-1, dtype=torch.float32).to(q_window.dtype)
    return attn[..., :-window]


def tova_scores(q_last: torch.Tensor, keys: torch.Tensor) -> torch.Tensor:
    """q_last [B,Hq,D] = RoPE'd query of the last position -> [B,Hkv,S] (tova_press.py:47-59): window attention
    with window 1, mean over dim 1 (all query heads), the one row repeated for every kv head, last position padded
    with max + 1."""
    attn = window_attention(q_last.unsqueeze(2), keys, 1)       # [B,Hq,1,S-1]
    scores = attn.mean(1)                                       # [B,1,S-1]
    scores = scores.repeat(1, keys.shape[1], 1)
    return F.pad(scores, (0, 1), value=scores.max().item() + 1)


def tova_scores_fp32(q_last: torch.Tensor, keys: torch.Tensor) -> torch.Tensor:
    """Same formula in fp32 throughout; the forced last position is +inf."""
    B, Hkv, S, D = keys.shape
    G = q_last.shape[1] // Hkv
    logits = torch.einsum("bhd,bhsd->bhs", q_last.float(), repeat_kv(keys, G).float()) / math.sqrt(D)
    p = F.softmax(logits, dim=-1)[..., :-1].mean(1, keepdim=True)
    return F.pad(p.repeat(1, Hkv, 1), (0, 1), value=float("inf"))


# --------------------------------------------------------------------------------------------------
# chunkkv_press.py:52-117 and block_press.py:49-98: wrappers over a score tensor / a score function
# --------------------------------------------------------------------------------------------------
def chunkkv_chunk_scores(scores: torch.Tensor, chunk_length: int) -> torch.Tensor:
    """chunkkv_press.py:78-92 — per-chunk score [B, n_chunks]: heads summed, mean over the chunk; a ragged tail is
    one more chunk."""
    S = scores.shape[-1]
    n_full, tail = divmod(S, chunk_length)
    main = scores[..., : n_full * chunk_length].sum(dim=1).view(-1, n_full, chunk_length).mean(dim=-1)
    if tail > 0:
        main = torch.cat([main, scores[..., -tail:].sum(dim=1).mean(dim=-1, keepdim=True)], dim=-1)
    return main


def chunkkv_kept_positions(chunk_scores: torch.Tensor, S: int, chunk_length: int, compression_ratio: float):
    """chunkkv_press.py:95-113 — positions of the kept chunks, ascending; batch element 0 decides for the batch."""
    n_full, tail = divmod(S, chunk_length)
    n_chunks = n_full + (tail > 0)
    n_chunks_kept = max(1, int(n_chunks * (1 - compression_ratio)))
    top = chunk_scores.topk(n_chunks_kept, dim=-1).indices[0]
    pos = []
    for c in top.tolist():
        pos.append(torch.arange(c * chunk_length, min((c + 1) * chunk_length, S)))
    return torch.cat(pos).sort().values


def block_kept_positions(score_fn, B: int, H: int, S: int, n_kept: int, block_size: int) -> torch.Tensor:
    """block_press.py:66-92 — iterative re-selection; score_fn(current_positions [B,H,n]) -> scores [B,H,n].
    The per-round top-k uses the lowest-position tie rule of the kernels (torch.topk leaves ties open)."""
    block = min(block_size, S)
    kept = torch.arange(n_kept).expand(B, H, -1)
    for lo in range(n_kept, S, block):
        hi = min(lo + block, S)
        current = torch.cat([kept, torch.arange(lo, hi).expand(B, H, -1)], dim=-1)
        sel = select_lowest_index_ties(score_fn(current), n_kept)
        kept = current.gather(-1, sel)
    return kept


# --------------------------------------------------------------------------------------------------
# expected_attention_press.py:62-165
# --------------------------------------------------------------------------------------------------
def avg_rope_matrix(cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """expected_attention_press.py:113-120 — R = mean_p (cos_p * I + sin_p * P), cos/sin [n_future, D]."""
    d = cos.shape[-1]
    eye = torch.eye(d, dtype=cos.dtype)
    P = torch.zeros((d, d), dtype=cos.dtype)
    P[d // 2:, : d // 2], P[: d // 2, d // 2:] = torch.eye(d // 2), -torch.eye(d // 2)
    R = cos.unsqueeze(1) * eye + sin.unsqueeze(1) * P
    return R.mean(dim=0)


def expected_attention_stats(hidden_states, q_weight, num_heads, head_dim, cos_future, sin_future, n_sink,
                             use_covariance=True, **kw):
    """expected_attention_press.py:62-124 — (mu [B,Hq,D], cov [B,Hq,D,D]) after the average RoPE."""
    h = hidden_states[:, n_sink:]
    q = prerope_queries(h, q_weight, num_heads, head_dim, **kw)
    mu = q.mean(dim=2, keepdim=True)
    cov = None
    if use_covariance:
        c = q - mu
        cov = torch.einsum("bnsi,bnsj->bnij", c, c) / h.shape[1]
    mu = mu.squeeze(2)
    R = avg_rope_matrix(cos_future, sin_future)
    mu = torch.matmul(mu, R.T)
    if cov is not None:
        cov = torch.matmul(R, torch.matmul(cov, R.T))
    return mu, cov


def expected_attention_scores(keys, values, mu, cov, epsilon: float, n_sink: int, use_vnorm: bool) -> torch.Tensor:
    """expected_attention_press.py:136-165 from (mu, cov) on. Every op rounds to the 16-bit dtype."""
    keys = keys[:, :, n_sink:]
    values = values[:, :, n_sink:]
    B, Hkv, L, d = keys.shape
    G = mu.shape[1] // Hkv
    kt = repeat_kv(keys, G).transpose(2, 3)                                            # :148
    scores = torch.matmul(mu.unsqueeze(2), kt).squeeze(2) / math.sqrt(d)               # :149
    if cov is not None:
        scores = scores + torch.einsum("bhin, bhij, bhjn->bhn", kt, cov, kt) / d / 2   # :151
    scores = F.softmax(scores, dim=-1)                                                 # :152
    scores = scores.view(B, Hkv, G, L).mean(dim=2)                                     # :155-156
    if use_vnorm:
        scores = (scores + epsilon) * values.norm(dim=-1)                              # :160
    return F.pad(scores, (n_sink, 0), value=scores.max().item() + 1)                   # :163


def expected_attention_scores_fp32(keys, values, mu, cov, epsilon: float, n_sink: int, use_vnorm: bool):
    """fp32 evaluation of the same formula from the (16-bit) inputs; forced-keep head is +inf."""
    k = keys[:, :, n_sink:].float()
    v = values[:, :, n_sink:].float()
    B, Hkv, L, d = k.shape
    G = mu.shape[1] // Hkv
    kr = repeat_kv(k, G)                                       # [B,Hq,L,d]
    logits = torch.einsum("bhd,bhld->bhl", mu.float(), kr) / math.sqrt(d)
    if cov is not None:
        logits = logits + torch.einsum("bhli,bhij,bhlj->bhl", kr, cov.float(), kr) / d / 2
    p = F.softmax(logits, dim=-1).view(B, Hkv, G, L).mean(dim=2)
    if use_vnorm:
        p = (p + epsilon) * v.norm(dim=-1)
    return F.pad(p, (n_sink, 0), value=float("inf"))


# --------------------------------------------------------------------------------------------------
# keydiff_press.py:36-46
# --------------------------------------------------------------------------------------------------
def keydiff_scores(keys: torch.Tensor) -> torch.Tensor:
    """keydiff_press.py:45-46 — the reference's two ATen calls; every op rounds to the key dtype."""
    anchor = F.normalize(keys, p=2, dim=-1).mean(dim=2, keepdim=True)
    return -F.cosine_similarity(keys, anchor, dim=-1)


def keydiff_scores_fp32(keys: torch.Tensor) -> torch.Tensor:
    """fp32 evaluation of the same formula from the 16-bit keys (what the kernel rounds once):
    anchor = mean_s k/max(||k||,1e-12);  score = -(k.anchor)/(max(||k||,1e-8) max(||anchor||,1e-8))."""
    k = keys.float()
    kn = k.norm(dim=-1, keepdim=True)
    anchor = (k / kn.clamp_min(1e-12)).mean(dim=2, keepdim=True)
    an = anchor.norm(dim=-1, keepdim=True).clamp_min(1e-8)
    return -((k * anchor).sum(-1) / kn.squeeze(-1).clamp_min(1e-8) / an.squeeze(-1))


# --------------------------------------------------------------------------------------------------
# key_rerotation_press.py:50-152 (SURVEY §8f "next" row 1)
# --------------------------------------------------------------------------------------------------
def rerotate_cos_sin(dtype, inv_freq: torch.Tensor, selected_positions: torch.Tensor):
    """key_rerotation_press.py:50-94 — cos/sin of (new position - old position) * inv_freq, computed in
    fp32 and rounded to the key dtype. selected_positions [B,H,n_kept] ascending; inv_freq [D/2]."""
    B, H, n_kept = selected_positions.shape
    idx = torch.arange(0, n_kept).float()[None, None, :].expand(B, H, n_kept)
    delta = (idx - selected_positions).unsqueeze(2)                        # [B,H,1,n_kept]
    freqs = (delta.float() * inv_freq[None, None, :, None].float()).transpose(2, 3)  # [B,H,n_kept,D/2]
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().contiguous().to(dtype), emb.sin().contiguous().to(dtype)


def rerotate_keys(keys: torch.Tensor, indices: torch.Tensor, inv_freq: torch.Tensor) -> torch.Tensor:
    """key_rerotation_press.py:96-131 — gather the kept keys (indices ascending) and rotate each from its
    old position to its new one. Products and the sum round to the key dtype (plain ATen ops)."""
    cos, sin = rerotate_cos_sin(keys.dtype, inv_freq, indices)
    k = gather_rows(keys, indices)
    return (k * cos) + (rotate_half(k) * sin)


def key_rerotation_compress(scores, keys, values, n_kept: int, inv_freq):
    """key_rerotation_press.py:133-152 with the scores given. Returns (K'', V', ascending indices)."""
    idx = torch.sort(topk_indices(scores, n_kept), dim=2).values
    return rerotate_keys(keys, idx, inv_freq), gather_rows(values, idx), idx


# --------------------------------------------------------------------------------------------------
# decoding_press.py:194-236
# --------------------------------------------------------------------------------------------------
def find_target_compression_ratio(q_len: int, target_tokens: int) -> float:
    if q_len <= target_tokens:
        return 0.0
    ratio = 1.0 - (target_tokens / q_len)
    low, high = 0.0, 1.0
    for _ in range(20):
        n_kept = int(q_len * (1 - ratio))
        if n_kept == target_tokens:
            break
        if n_kept > target_tokens:
            low = ratio
            ratio = (ratio + high) / 2
        else:
            high = ratio
            ratio = (low + ratio) / 2
    return ratio


# --------------------------------------------------------------------------------------------------
# whole compress() calls from raw tensors — what bench.py times as the CPU baseline
# --------------------------------------------------------------------------------------------------
def knorm_compress(keys, values, compression_ratio: float):
    n_kept = kept_count(keys.shape[2], compression_ratio)
    return compress_with_scores(knorm_scores(keys), keys, values, n_kept)


def streaming_compress(keys, values, compression_ratio: float, n_sink: int = 4):
    n_kept = kept_count(keys.shape[2], compression_ratio)
    return compress_with_scores(streaming_scores(keys, compression_ratio, n_sink), keys, values, n_kept)


def snapkv_compress(q_window, keys, values, compression_ratio: float, window: int = 64, kernel_size: int = 5):
    n_kept = kept_count(keys.shape[2], compression_ratio)
    return compress_with_scores(snapkv_scores(q_window, keys, window, kernel_size), keys, values, n_kept)


def expected_attention_compress(keys, values, mu, cov, compression_ratio: float, epsilon=0.0, n_sink=4,
                                use_vnorm=True):
    n_kept = kept_count(keys.shape[2], compression_ratio)
    scores = expected_attention_scores(keys, values, mu, cov, epsilon, n_sink, use_vnorm)
    return compress_with_scores(scores, keys, values, n_kept)
