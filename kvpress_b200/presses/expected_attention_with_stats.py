"""ExpectedAttentionStatsPress: ExpectedAttention with query statistics measured once, offline, per model.

API mirror of `/root/reference/kvpress/presses/expected_attention_with_stats.py:21-110`.
With (mu, Sigma) fixed per layer the per-prompt prologue (q_proj over the whole prompt + a [D x D] covariance per
head: 4.4 TFLOP at 128k on Llama-3.1-8B) disappears and the press is just the sm_100a cache scan; it also
stops reading `hidden_states`, so a DecodingPress around it buffers nothing.

Statistics live in a folder in the same layout `huggingface_hub.PyTorchModelHubMixin` writes (config.json +
model.safetensors with `query_mean` [L, Hq, D] and `query_cov` [L, Hq, D, D]), so folders produced by the reference's
collection script load here and the other way round. There is no network in this package: give `stats_folder`, or
assign `press.mu` / `press.cov` directly, or measure them with `collect_query_statistics`.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field
from typing import Iterable, Optional

import torch
from torch import nn

from kvpress_b200.presses.expected_attention_press import ExpectedAttentionPress
from kvpress_b200.utils import get_prerope_query_states

_WEIGHTS, _CONFIG = "model.safetensors", "config.json"


class ExpectedAttentionStats(nn.Module):
    """Container of per-layer query statistics (same fields and file layout as the reference's)."""

    def __init__(self, num_layers: int, num_heads: int, head_dim: int, dataset_name: str = "", model_name: str = "",
                 num_samples: int = 0, sample_seq_len: int = 0, n_sink: int = 4):
        super().__init__()
        self.query_mean = nn.Parameter(torch.zeros(num_layers, num_heads, head_dim), requires_grad=False)
        self.query_cov = nn.Parameter(torch.zeros(num_layers, num_heads, head_dim, head_dim), requires_grad=False)
        self.meta = dict(num_layers=num_layers, num_heads=num_heads, head_dim=head_dim, dataset_name=dataset_name,
                         model_name=model_name, num_samples=num_samples, sample_seq_len=sample_seq_len, n_sink=n_sink)

    def stats_id(self) -> str:
        m = self.meta
        return ("alessiodevoto/exp_att_stats_" + m["model_name"].replace("/", "_") + "_"
                + m["dataset_name"].replace("/", "_") + f"_{m['num_samples']}_{m['sample_seq_len']}_{m['n_sink']}")

    def save_pretrained(self, folder: str) -> None:
        from safetensors.torch import save_file

        os.makedirs(folder, exist_ok=True)
        save_file({"query_mean": self.query_mean.data.contiguous(), "query_cov": self.query_cov.data.contiguous()},
                  os.path.join(folder, _WEIGHTS))
        with open(os.path.join(folder, _CONFIG), "w") as f:
            json.dump(self.meta, f, indent=2)

    @classmethod
    def from_pretrained(cls, folder: str) -> "ExpectedAttentionStats":
        from safetensors.torch import load_file

        if not os.path.isdir(folder):
            raise ValueError(f"{folder} is not a local statistics folder (this package never downloads from the Hub)")
        with open(os.path.join(folder, _CONFIG)) as f:
            meta = json.load(f)
        stats = cls(**{k: meta[k] for k in ("num_layers", "num_heads", "head_dim", "dataset_name", "model_name",
                                            "num_samples", "sample_seq_len", "n_sink") if k in meta})
        tensors = load_file(os.path.join(folder, _WEIGHTS))
        stats.query_mean.data = tensors["query_mean"]
        stats.query_cov.data = tensors["query_cov"]
        return stats


@dataclass
class ExpectedAttentionStatsPress(ExpectedAttentionPress):
    sample_seq_len: int = 1000
    num_samples: int = 100
    dataset_name: str = "kmfoda/booksum"
    stats_folder: Optional[str] = None
    mu: torch.Tensor = field(init=False, default=None)    # [L, Hq, D], set in post_init_from_model (or by hand)
    cov: torch.Tensor = field(init=False, default=None)   # [L, Hq, D, D]

    needs_hidden_states = False      # the stored statistics replace the prompt's queries ...
    needs_hidden_states_len = True   # ... but q_len (= hidden_states.shape[1]) places the future RoPE positions

    def get_query_statistics(self, module: nn.Module, hidden_states: torch.Tensor):
        """This layer's stored statistics, rotated to the positions after the current context, for every batch row."""
        B, q_len = hidden_states.shape[0], hidden_states.shape[1]
        layer = module.layer_idx
        cov = self.cov[layer] if self.use_covariance else None
        mu, cov = self.apply_avg_rope(module, self.mu[layer], cov, q_len)
        mu = mu.unsqueeze(0).expand(B, -1, -1)
        return mu, (None if cov is None else cov.unsqueeze(0).expand(B, -1, -1, -1))

    def post_init_from_model(self, model):
        if self.mu is not None and self.cov is not None:
            return
        if self.stats_folder is None:
            wanted = ExpectedAttentionStats(
                num_layers=model.config.num_hidden_layers, num_heads=model.config.num_attention_heads,
                head_dim=model.config.head_dim, dataset_name=self.dataset_name, model_name=model.config.name_or_path,
                num_samples=self.num_samples, sample_seq_len=self.sample_seq_len, n_sink=self.n_sink).stats_id()
            raise ValueError(
                f"No statistics given for {wanted}. Pass stats_folder=<local folder> (the layout the reference's "
                "collection script writes), or assign press.mu / press.cov, or run "
                "kvpress_b200.presses.expected_attention_with_stats.collect_query_statistics(model, batches).")
        stats = ExpectedAttentionStats.from_pretrained(self.stats_folder)
        self.mu = stats.query_mean.data.to(model.device, dtype=model.dtype)
        self.cov = stats.query_cov.data.to(model.device, dtype=model.dtype)


@torch.inference_mode()
def collect_query_statistics(model, batches: Iterable[torch.Tensor], n_sink: int = 4, dataset_name: str = "",
                             sample_seq_len: int = 0) -> ExpectedAttentionStats:
    """Mean and covariance of the pre-RoPE queries of every layer over the given token batches ([1, S] id tensors),
    skipping the first `n_sink` positions of each (reference `collect_queries`, :141-186, without the dataset
    plumbing: the queries are taken from the attention input with the same projection the press uses)."""
    backbone = model.model.language_model if hasattr(model.model, "language_model") else model.model
    attns = [layer.self_attn for layer in backbone.layers]
    L = len(attns)
    cfg = model.config
    Hq, D = cfg.num_attention_heads, getattr(cfg, "head_dim", cfg.hidden_size // cfg.num_attention_heads)
    count = 0
    s1 = torch.zeros(L, Hq, D, dtype=torch.float64, device=model.device)
    s2 = torch.zeros(L, Hq, D, D, dtype=torch.float64, device=model.device)
    handles = []

    def make_hook(i, attn):
        def hook(module, args, kwargs):
            q = get_prerope_query_states(attn, kwargs["hidden_states"][:, n_sink:]).double()     # [B, Hq, S', D]
            s1[i] += q.sum(dim=(0, 2))
            s2[i] += torch.einsum("bhsi,bhsj->hij", q, q)
        return hook

    for i, attn in enumerate(attns):
        handles.append(attn.register_forward_pre_hook(make_hook(i, attn), with_kwargs=True))
    n_batches = 0
    try:
        for ids in batches:
            ids = ids.to(model.device)
            backbone(input_ids=ids)
            count += ids.shape[0] * max(0, ids.shape[1] - n_sink)
            n_batches += 1
    finally:
        for h in handles:
            h.remove()
    if count < 2:
        raise ValueError("need at least two query positions beyond n_sink to estimate a covariance")
    mean = s1 / count
    cov = (s2 - count * torch.einsum("lhi,lhj->lhij", mean, mean)) / (count - 1)
    stats = ExpectedAttentionStats(L, Hq, D, dataset_name=dataset_name, model_name=getattr(cfg, "name_or_path", ""),
                                   num_samples=n_batches, sample_seq_len=sample_seq_len, n_sink=n_sink)
    stats.query_mean.data = mean.float()
    stats.query_cov.data = cov.float()
    return stats
