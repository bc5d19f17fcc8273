"""Expert-sharded MoE with the real kernels: two processes share cuda:0 (gloo backend, so one GPU is enough),
each owns the experts e % 2 == rank of a small Mixtral-style model, every MoE layer ends in one all-reduce of [T, dim].
Prefill + decode logits must equal the unsharded model's bit for bit (top-2 routing: bf16(a + b) in any order)."""
import os
import socket
import sys
from pathlib import Path

import pytest
import torch
import torch.multiprocessing as mp

REPO = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.gpu


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, q):
    try:
        sys.path.insert(0, str(REPO))
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        torch.cuda.set_device(0)
        torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
        import mistral_inference_b200 as mi
        from mistral_inference_b200 import synth
        from mistral_inference_b200.cache import BufferCache
        from mistral_inference_b200.transformer import Transformer

        p = synth.shape("tiny-moe", sliding_window=16)
        sd = synth.synth_state_dict(p, 2, torch.bfloat16, "cuda")

        def build(expert_parallel):
            args = mi.TransformerArgs.from_dict(dict(p))
            args.max_batch_size = 2
            m = Transformer.empty(args, "cuda", torch.bfloat16, expert_parallel=expert_parallel)
            m.load_state_dict(sd)
            return m.eval()

        def run(m):
            cache = BufferCache(m.n_local_layers, 2, 64, p["n_kv_heads"], p["head_dim"], p.get("sliding_window")).to(m.device, m.dtype)
            seqlens = [12, 9]
            toks = torch.tensor(synth.synth_prompt(sum(seqlens), p["vocab_size"], 4), device="cuda")
            outs = [m.forward(toks, seqlens, cache)]
            nxt = torch.tensor([5, 7], device="cuda")
            for _ in range(3):
                lg = m.forward(nxt, [1, 1], cache)
                outs.append(lg)
                nxt = lg.argmax(-1)
            return torch.cat(outs).cpu()

        sharded = run(build((rank, world)))  # both ranks run in lock step: the all-reduces pair up
        full = run(build(None)) if rank == 0 else None
        ok = bool(torch.equal(sharded, full)) if rank == 0 else True
        worst = float((sharded - full).abs().max()) if rank == 0 else 0.0
        q.put((rank, ok, worst, ""))
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    except Exception as e:  # surface the failure in the parent instead of a queue timeout
        q.put((rank, False, -1.0, repr(e)))
        raise


def test_expert_sharded_forward_equals_unsharded():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = sorted(q.get(timeout=240) for _ in range(2))
    for pr in procs:
        pr.join(timeout=60)
    for rank, ok, worst, err in res:
        assert ok, f"rank {rank}: {err or 'sharded logits differ from the unsharded model by %g' % worst}"
    assert all(pr.exitcode == 0 for pr in procs)
