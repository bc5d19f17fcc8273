"""`params.json` schema (mirror of mistral_inference/args.py:29-59 and moe.py:10-13).

`simple_parsing` is not a dependency here: `from_dict` is a few lines (the reference only uses
`Serializable.from_dict`, transformer.py:306-307).
"""
from dataclasses import dataclass, fields
from typing import List, Optional, Union


@dataclass
class MoeArgs:
    num_experts: int
    num_experts_per_tok: int

    @classmethod
    def from_dict(cls, d: dict) -> "MoeArgs":
        return cls(**{f.name: d[f.name] for f in fields(cls) if f.name in d})


@dataclass
class LoraArgs:
    rank: int
    scaling: float

    @classmethod
    def from_dict(cls, d: dict) -> "LoraArgs":
        return cls(**{f.name: d[f.name] for f in fields(cls) if f.name in d})


@dataclass
class TransformerArgs:
    dim: int
    n_layers: int
    head_dim: int
    hidden_dim: int
    n_heads: int
    n_kv_heads: int
    norm_eps: float
    vocab_size: int

    max_batch_size: int = 0

    # For rotary embeddings. If not set, 1e6 is used (transformer.py:115).
    rope_theta: Optional[float] = None
    # If this is set, MoE layers replace the dense FeedForward.
    moe: Optional[MoeArgs] = None
    lora: Optional[LoraArgs] = None
    sliding_window: Union[None, int, List[Optional[int]]] = None
    _sliding_window: Union[None, int, List[Optional[int]]] = None
    model_type: str = "transformer"

    vision_encoder: Optional[dict] = None

    def __post_init__(self) -> None:
        assert self.model_type == "transformer", self.model_type
        assert self.sliding_window is None or self._sliding_window is None
        # same aliasing as args.py:58-59
        self.sliding_window = self.sliding_window if self.sliding_window is not None else self._sliding_window
        if self.vision_encoder is not None:
            raise NotImplementedError("vision_encoder (Pixtral) is outside the accelerated hot path (SURVEY.md 2, row 10)")
        if isinstance(self.moe, dict):
            self.moe = MoeArgs.from_dict(self.moe)
        if isinstance(self.lora, dict):
            self.lora = LoraArgs.from_dict(self.lora)

    @classmethod
    def from_dict(cls, d: dict) -> "TransformerArgs":
        known = {f.name for f in fields(cls)}
        return cls(**{k: v for k, v in d.items() if k in known})
