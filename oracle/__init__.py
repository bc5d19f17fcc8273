"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's transformer hot path (mistral-inference @ 2557e12:
transformer_layers.py, cache.py, rope.py, moe.py, transformer.py, generate.py) used as the
parity checker for the sm_100a CUDA path.

Who may import this package: `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py`.  Nothing under `mistral_inference_b200/` imports it; the
product path has no CPU fallback and fails loudly when the CUDA library is missing.

Parity status: the restatement is pinned (bit-exact on CPU, see tests/test_oracle_vs_reference.py
and tests/golden/) against the reference's own modules run UNMODIFIED behind the import shims in
`oracle/ref_shims.py`.  The attention arithmetic itself lives in xformers (un-vendored dependency,
poetry.lock:1927 pins 0.0.26.post1) whose source is not available here: at that boundary parity is
UNPINNED except through the reference's own self-consistency property (decode == re-prefill,
tests/test_generate.py:36-69,199-230), which the oracle and the CUDA path are both tested for.
"""
