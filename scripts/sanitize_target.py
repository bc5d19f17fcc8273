#!/usr/bin/env python
"""Small workload for compute-sanitizer (scripts/sanitize.sh): every kernel family on tiny shapes, few steps.
  dense tiny model: 200-token prefill (tcgen05 GEMM + tcgen05 attention), chunked prefill, batch-1 decode (megakernel),
  batch-3 decode (stream-K GEMMs, TMA decode attention, device-side step state), sampling kernels;
  tiny MoE model: grouped experts in prefill and batched decode, megakernel MoE decode."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import mistral_inference_b200 as mi  # noqa: E402
import synth  # noqa: E402
from mistral_inference_b200.transformer import Transformer  # noqa: E402

os.environ.setdefault("MB200_DECODE_GRAPH", "0")  # the sanitizer instruments kernel launches, not graph replays
which = sys.argv[1] if len(sys.argv) > 1 else "all"
for shape in ("tiny", "tiny-moe"):
    if which not in ("all", shape):
        continue
    p = synth.shape(shape, sliding_window=64)
    args = mi.TransformerArgs.from_dict(dict(p))
    args.max_batch_size = 3
    m = Transformer.empty(args, "cuda", torch.bfloat16)
    m.load_state_dict(synth.synth_state_dict(p, 1, torch.bfloat16, "cuda"))
    t1, _ = mi.generate([synth.synth_prompt(200, p["vocab_size"], 1)], m, max_tokens=3, temperature=0.0)
    t3, _ = mi.generate([synth.synth_prompt(n, p["vocab_size"], 2 + i) for i, n in enumerate((40, 33, 37))], m, max_tokens=3, temperature=0.0, chunk_size=16)
    ts, _ = mi.generate([[1, 2, 3], [4, 5, 6, 7]], m, max_tokens=2, temperature=0.7)
    torch.cuda.synchronize()
    print(shape, "ok", t1, t3, ts)
