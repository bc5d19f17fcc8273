"""N > 1 host logic on CPU (gloo, world_size 2): the dense configs are "replicas only" (DESIGN.md section 1e), so the only
cross-rank step is the timing rule -- max over ranks -- and the whole-job aggregation; plus the reference-compatible
pipeline split of the state dict (transformer.py:94-98,244-295) seen from both ranks."""
import os
import socket
import sys
from pathlib import Path

import torch
import torch.multiprocessing as mp

REPO = Path(__file__).resolve().parents[1]


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, q):
    sys.path.insert(0, str(REPO))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    import mistral_inference_b200 as mi
    from mistral_inference_b200 import synth
    from mistral_inference_b200.transformer import Transformer

    my_ms = 100.0 + 25.0 * rank  # rank 1 is the slow replica
    job_ms = bench.max_over_ranks(my_ms, world, "cpu")
    value = bench.whole_job_tokens_per_s(world, batch=1, steps=50, elapsed_ms=job_ms)
    # pipeline split: each rank keeps only its own layers / embeddings / head
    p = synth.shape("tiny", n_layers=4)
    args = mi.TransformerArgs.from_dict(dict(p))
    m = Transformer(args, pipeline_rank=rank, num_pipeline_ranks=world).to(torch.bfloat16)
    m.load_state_dict(synth.synth_state_dict(p, 1))
    q.put((rank, job_ms, value, sorted(m.layers.keys()), m.tok_embeddings is not None, m.norm is not None))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_replicas_timing_rule_and_pipeline_split_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    (r0, ms0, v0, l0, emb0, head0), (r1, ms1, v1, l1, emb1, head1) = res
    assert ms0 == ms1 == 125.0                      # max over ranks, identical on both
    assert v0 == v1 == 2 * 1 * 50 * 1000.0 / 125.0   # whole-job tokens/s, weak scaling
    assert l0 == ["0", "1"] and l1 == ["2", "3"]
    assert emb0 and not emb1 and head1 and not head0
