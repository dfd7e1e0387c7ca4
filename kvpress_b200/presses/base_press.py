"""BasePress: hook lifecycle of a KV-cache compression method.

API mirror of `/root/reference/kvpress/presses/base_press.py` (BasePress :43-207,
is_prefilling :37-40): `with press(model): model(...)` registers one forward hook per
`layer.self_attn`; after each attention forward during prefill the hook takes that layer's K/V
out of the cache, calls `compress`, and writes the compacted tensors back.

Differences that do not change results:
  * prefill is detected from Python ints (`cache.get_seq_length(layer) == q_len`) instead of
    `cache_position[-1].item()`: no device->host sync per layer, and it works on transformers
    versions that no longer pass `cache_position` to the attention module;
  * K/V must be CUDA bf16/fp16 tensors — the compaction runs in the sm_100a library.
"""
from __future__ import annotations

import logging
from contextlib import contextmanager
from dataclasses import dataclass
from typing import Generator

import torch
from torch import nn
from transformers import PreTrainedModel

from kvpress_b200.utils import extract_keys_and_values

logger = logging.getLogger(__name__)


def _supported_models() -> tuple:
    import transformers

    names = ["LlamaForCausalLM", "MistralForCausalLM", "Phi3ForCausalLM", "Qwen2ForCausalLM", "Qwen3ForCausalLM",
             "Gemma3ForConditionalGeneration"]
    return tuple(getattr(transformers, n) for n in names if hasattr(transformers, n))


# Model classes the hooks are written for (same list as the reference, base_press.py:24-34). Shape coverage of the
# sm_100a library on top of that list: every scorer takes any head_dim that is a multiple of 8 (<= 256); the two
# tensor-core scorers — SnapKVPress / PyramidKVPress and ExpectedAttentionPress with use_covariance=True — are
# instantiated for head_dim 64 / 128, Hq/Hkv <= 8 and (Hq/Hkv) * window_size <= 512 (Llama, Mistral, Qwen2, Qwen3).
# Outside that set (Phi3: head_dim 96, Gemma3: 256) the C ABI returns "unsupported shape" and the presses evaluate the
# SCORE stage with cuBLAS GEMMs on the GPU (kvpress_b200/wide_head_scores.py, logged once) while selection and
# compaction stay on the sm_100a kernels. There is no CPU path anywhere.
SUPPORTED_MODELS = _supported_models()


def is_prefilling(cache_position, q_len: int) -> bool:
    """Reference-compatible helper (base_press.py:37-40). The hooks below do not use it (it syncs)."""
    last = cache_position[-1] + 1 == q_len
    return bool(last.item() if isinstance(last, torch.Tensor) else last)


def layer_is_prefilling(cache, layer_idx: int, q_len: int) -> bool:
    """True when the layer's cache holds exactly the tokens of this forward call (no sync)."""
    return int(cache.get_seq_length(layer_idx)) == int(q_len)


PHASE_KEY = "_kvp_is_prefilling"


def hook_is_prefilling(module: nn.Module, kwargs: dict) -> bool:
    """Phase of this forward call for a hook. A wrapper that chains several presses in one hook (ComposedPress)
    decides once, before the first press shortens the cache, and pins the answer in kwargs[PHASE_KEY]."""
    if PHASE_KEY in kwargs:
        return bool(kwargs[PHASE_KEY])
    return layer_is_prefilling(kwargs["past_key_values"], module.layer_idx, kwargs["hidden_states"].shape[1])


def write_back(cache, layer_idx: int, keys: torch.Tensor, values: torch.Tensor) -> None:
    layer = cache.layers[layer_idx]
    layer.keys = keys
    layer.values = values


@dataclass
class BasePress:
    """Base class of every press. Subclasses implement `compress`."""

    def post_init_from_model(self, model: PreTrainedModel):
        """Optional: derive press parameters from the model before hooks are registered."""

    def compress(
        self,
        module: nn.Module,
        hidden_states: torch.Tensor,
        keys: torch.Tensor,
        values: torch.Tensor,
        attentions: torch.Tensor,
        kwargs: dict,
    ) -> tuple[torch.Tensor, torch.Tensor]:
        """Return the compacted (keys, values), both [B, Hkv, n_kept, D]."""
        raise NotImplementedError("compress must be implemented by the subclass")

    def forward_hook(self, module: nn.Module, input: list[torch.Tensor], kwargs: dict, output: list):
        """Post-forward hook of an attention layer: compress this layer's cache during prefill only."""
        hidden_states = kwargs["hidden_states"]
        cache = kwargs["past_key_values"]
        layer_idx = module.layer_idx
        if not hook_is_prefilling(module, kwargs):
            return output
        keys, values = extract_keys_and_values(cache, layer_idx)
        keys, values = self.compress(module, hidden_states, keys, values, output[1], kwargs)
        write_back(cache, layer_idx, keys, values)
        return output

    @contextmanager
    def __call__(self, model: PreTrainedModel) -> Generator:
        """Context manager: hooks are live inside the `with` block and always removed on exit."""
        if SUPPORTED_MODELS and not isinstance(model, SUPPORTED_MODELS):
            logger.warning(f"Model {type(model)} not tested, supported models: {SUPPORTED_MODELS}")
        self.post_init_from_model(model)
        backbone = model.model.language_model if hasattr(model.model, "language_model") else model.model
        handles = []
        try:
            for layer in backbone.layers:
                attn = layer.self_attn
                if getattr(attn, "is_sliding", False) and type(model).__name__.startswith("Gemma3"):
                    continue  # sliding-window layers keep their full (short) cache
                attn.rotary_emb = backbone.rotary_emb  # ExpectedAttention needs future-position RoPE
                handles.append(attn.register_forward_hook(self.forward_hook, with_kwargs=True))
            yield
        finally:
            for handle in handles:
                handle.remove()
