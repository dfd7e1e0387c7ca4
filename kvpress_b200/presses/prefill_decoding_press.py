"""PrefillDecodingPress: one press for the prefill phase, a DecodingPress for generation.
API mirror of `/root/reference/kvpress/presses/prefill_decoding_press.py:18-103`; the phase is decided
from the cache length (no device sync) like everywhere else in this package."""
from __future__ import annotations

import logging
from contextlib import contextmanager
from dataclasses import dataclass
from typing import Optional

from torch import nn
from transformers import PreTrainedModel

from kvpress_b200.presses.base_press import BasePress, hook_is_prefilling
from kvpress_b200.presses.decoding_press import DecodingPress

logger = logging.getLogger(__name__)


@dataclass
class PrefillDecodingPress(BasePress):
    prefilling_press: Optional[BasePress] = None
    decoding_press: Optional[DecodingPress] = None

    def post_init_from_model(self, model):
        for press in (self.prefilling_press, self.decoding_press):
            if press is not None:
                press.post_init_from_model(model)

    def _phase_press(self, module: nn.Module, kwargs: dict):
        if hook_is_prefilling(module, kwargs):
            return self.prefilling_press
        return self.decoding_press

    def compress(self, module: nn.Module, hidden_states, keys, values, attentions, kwargs: dict):
        press = self._phase_press(module, dict(kwargs, hidden_states=hidden_states))
        if press is None:
            logger.warning("No compression applied during prefill or decoding phase")
            return keys, values
        return press.compress(module, hidden_states, keys, values, attentions, kwargs)

    def forward_hook(self, module: nn.Module, input, kwargs: dict, output: list):
        press = self._phase_press(module, kwargs)
        return output if press is None else press.forward_hook(module, input, kwargs, output)

    @contextmanager
    def __call__(self, model: PreTrainedModel):
        try:
            with super().__call__(model):
                yield
        finally:
            if self.decoding_press is not None:
                self.decoding_press.reset()
