"""CompressionRatioDecodingPress: DecodingPress whose target is a fraction of the tokens seen so far.
API mirror of `/root/reference/kvpress/presses/compression_ratio_decoding_press.py:13-50`."""
from __future__ import annotations

from dataclasses import dataclass, field

from kvpress_b200.presses.decoding_press import DecodingPress


@dataclass
class CompressionRatioDecodingPress(DecodingPress):
    target_compression_ratio: float = 0.5
    target_size: int = field(default=1, init=False)

    def __post_init__(self):
        super().__post_init__()
        assert 0 <= self.target_compression_ratio < 1, "target_compression_ratio must be between 0 and 1"

    def _resolve_total_tokens_seen(self, kwargs: dict) -> int:
        position_ids = kwargs.get("position_ids")
        if position_ids is None:
            raise NotImplementedError("CompressionRatioDecodingPress requires logical position_ids in kwargs")
        return int(position_ids.max().item()) + 1

    def _resolve_target_size(self, kwargs: dict) -> int:
        return max(1, int(self._resolve_total_tokens_seen(kwargs) * (1 - self.target_compression_ratio)))
