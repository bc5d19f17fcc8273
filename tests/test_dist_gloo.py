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
    import synth
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


def _moe_worker(rank: int, world: int, port: int, q):
    """Expert-parallel MoE on CPU, world 2: the exchange ALGORITHM of csrc/moe.cuh -- every rank derives the same
    deterministic row plan, computes bf16(w * expert(x)) for the rows of its own experts, the rows are gathered on every rank
    (all-gather, no reduction) and combined in ascending expert index with a bf16 rounding per step -- against the oracle's
    unsharded MoE (moe.py:24-32), for top-2 AND top-3 (any reduction-order-free design must be exact for k > 2 as well); plus the
    product's key filtering of an expert-sharded model."""
    sys.path.insert(0, str(REPO))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    import mistral_inference_b200 as mi
    import synth
    from mistral_inference_b200.transformer import Transformer
    from oracle import restatement as R
    from tests.util import moe_plan_host, moe_route_host

    p = synth.shape("tiny-moe")
    E, dim = p["moe"]["num_experts"], p["dim"]
    sd = synth.synth_state_dict(p, 3)
    experts = [tuple(sd[f"layers.0.feed_forward.experts.{e}.{n}.weight"] for n in ("w1", "w2", "w3")) for e in range(E)]
    gate_w = sd["layers.0.feed_forward.gate.weight"]
    x = synth.synth_tensor("x", (37, dim), 5, torch.bfloat16, "cpu")
    exact = []
    for k in (2, 3):
        want = R.moe_forward(x, gate_w, experts, k)
        sel, wts = moe_route_host(x, gate_w, k)
        slot, seg, tiles = moe_plan_host(sel, E, 32, (rank, world))
        yw = torch.zeros(seg[-1], dim, dtype=torch.bfloat16)  # this rank's rows (the down projection's epilogue output)
        for t in range(x.shape[0]):
            for j in range(k):
                e = int(sel[t, j])
                if e % world == rank:
                    yw[slot[t, j]] = wts[t, j] * R.feed_forward(x[t:t + 1], *experts[e])[0]
        gathered = [torch.zeros_like(yw) for _ in range(world)]
        torch.distributed.all_gather(gathered, yw)  # NVLink peer stores on the GPU; every row has exactly one writer
        rows = torch.zeros_like(yw)
        for e in range(E):
            rows[seg[e]:seg[e + 1]] = gathered[e % world][seg[e]:seg[e + 1]]
        got = torch.zeros_like(x)
        for t in range(x.shape[0]):
            r = rows[slot[t, 0]].clone()
            for j in range(1, k):
                r = r + rows[slot[t, j]]  # bf16 add, one rounding per step, ascending expert index
            got[t] = r
        exact.append(bool(torch.equal(got, want)))
        assert all(e % world == rank for e, _ in tiles)
    # key filtering of the sharded model: this rank holds the router, all attention weights and only its own experts
    args = mi.TransformerArgs.from_dict(dict(p))
    m = Transformer(args, expert_parallel=(rank, world)).to(torch.bfloat16)
    m.load_state_dict(sd)
    held = sorted({int(key.split(".")[4]) for key in m.state_dict() if ".experts." in key})
    q.put((rank, exact, held, m._megakernel_ok(1),
           m._owns_key("layers.0.feed_forward.experts.%d.w1.weight" % ((rank + 1) % world)), "layers.0.feed_forward.gate.weight" in m.state_dict()))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_expert_sharded_moe_world2_matches_unsharded_oracle():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_moe_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = sorted(q.get(timeout=180) for _ in range(2))
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    for rank, exact, held, mega, owns_other, has_gate in res:
        assert exact == [True, True], f"rank {rank}: gathered-rows MoE differs from the unsharded oracle (top-2, top-3): {exact}"
        assert held == [e for e in range(8) if e % 2 == rank]
        assert not mega and not owns_other and has_gate  # sharded decode: per-layer kernels + the fused exchange, not the megakernel
