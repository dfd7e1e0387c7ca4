"""ComposedPress: several presses applied one after the other in a single forward hook.
API mirror of `/root/reference/kvpress/presses/composed_press.py:14-62`."""
from __future__ import annotations

from dataclasses import dataclass

from kvpress_b200.presses.base_press import PHASE_KEY, BasePress, hook_is_prefilling


@dataclass
class ComposedPress(BasePress):
    presses: list[BasePress]

    def __post_init__(self):
        self.compression_ratio = None  # known after the first forward pass

    def post_init_from_model(self, model):
        for press in self.presses:
            press.post_init_from_model(model)

    def forward_hook(self, module, input, kwargs, output):
        # the first press shortens the cache: decide the phase once for the whole chain
        kwargs = dict(kwargs, **{PHASE_KEY: hook_is_prefilling(module, kwargs)})
        kept = 1.0
        for press in self.presses:
            output = press.forward_hook(module, input, kwargs, output)
            kept *= 1 - press.compression_ratio
        self.compression_ratio = 1 - kept
        return output
