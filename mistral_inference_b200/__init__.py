"""B200-native drop-in for the mistral-inference transformer hot path.

Public API mirrors mistral_inference: `Transformer.from_folder / forward / forward_partial`,
`generate`, `BufferCache`, `TransformerArgs`.  All compute runs in libmb200.so (hand-written sm_100a
CUDA behind the C ABI in include/mistral_b200.h); importing this package does not load the library,
using the model does -- and fails loudly if it is missing.
"""
from .args import MoeArgs, TransformerArgs  # noqa: F401
from .cache import BufferCache  # noqa: F401
from .generate import generate  # noqa: F401  (the function; the submodule stays importable as mistral_inference_b200.generate)
from .transformer import Transformer  # noqa: F401
