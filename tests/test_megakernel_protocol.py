"""CPU model-check of the decode megakernel's shared-memory ring protocol (csrc/decode_megakernel.cuh).

The kernel cannot run here (no GPU), but its synchronisation skeleton can: one producer and 8 consumer warps exchange
ring slots through mbarriers whose waits are PARITY based -- `try_wait.parity(P)` succeeds iff the barrier's current
phase parity differs from P, so a waiter that is two laps ahead of a barrier aliases and passes early (reads stale
data, then corrupts the empty/full handshake: on the GPU that is a hang).  This test replays the exact stage order of
produce_matrix / consume_matrix under random interleavings and checks that every consumed slot holds the stage the
consumer expects and that nothing deadlocks.  It also shows that the naive "warp w owns pairs w, w+8, ..." order (the
first warp-per-pair version) is caught by the same model."""
import random

import pytest

WARPS = 8


class Mbar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def test(self, parity):  # try_wait.parity
        return (self.phase & 1) != parity


def stage_sequence(pairs_per_matrix, nch_per_matrix, grouped):
    """Yields per matrix the list of (it, pair, ch, owner_warp) in PRODUCTION order.  nch == 0 marks the K/V slice stages of
    the attention phase: `pairs` stages, each consumed by ALL warps (owner -1)."""
    it = 0
    mats = []
    for pairs, nch in zip(pairs_per_matrix, nch_per_matrix):
        seq = []
        if nch == 0:
            for j in range(pairs):
                seq.append((it, j, 0, -1))
                it += 1
        elif grouped:
            for g0 in range(0, pairs, WARPS):
                g = min(WARPS, pairs - g0)
                for ch in range(nch):
                    for w in range(g):
                        seq.append((it, g0 + w, ch, w))
                        it += 1
        else:  # naive: pair-major, warp = pair % 8
            for p in range(pairs):
                for ch in range(nch):
                    seq.append((it, p, ch, p % WARPS))
                    it += 1
        mats.append(seq)
    return mats


def simulate(pairs_per_matrix, nch_per_matrix, n_stages, grouped, seed, max_steps=400000, guard=True):
    """guard: the consumer first waits on empty[slot] for the PREVIOUS fill (parity of round r-1) before testing
    full[slot] for round r -- bulk copies land out of order, so without it a warp can test `full` one phase early."""
    rng = random.Random(seed)
    mats = stage_sequence(pairs_per_matrix, nch_per_matrix, grouped)
    full = [Mbar(1) for _ in range(n_stages)]
    empty = [Mbar(WARPS) for _ in range(n_stages)]  # weight stages: the owner arrives x8; K/V stages: every warp x1
    slot_data = [None] * n_stages
    inflight = []  # (slot, it): bulk copies issued, not yet landed
    prod = [s for m in mats for s in m]
    prod_i = 0
    # per-warp program: list of ("stage", it) / ("bar", k) in program order
    progs = [[] for _ in range(WARPS)]
    bar_id = 0
    for m in mats:
        if m and m[0][3] == -1:  # K/V stages: every warp waits on every stage in order (no guard needed), arrives once
            for w in range(WARPS):
                for (it, p, ch, ww) in m:
                    progs[w].append(("kvstage", it))
        elif grouped:
            groups = {}
            for (it, p, ch, w) in m:
                groups.setdefault(p // WARPS, []).append((it, p, ch, w))
            for gk in sorted(groups):
                for w in range(WARPS):
                    for (it, p, ch, ww) in sorted(x for x in groups[gk] if x[3] == w):
                        if guard:
                            progs[w].append(("prev", it))
                        progs[w].append(("stage", it))
                    progs[w].append(("bar", bar_id))
                bar_id += 1
        else:
            for w in range(WARPS):
                for (it, p, ch, ww) in m:
                    if ww == w:
                        progs[w].append(("stage", it))
        for w in range(WARPS):  # grid barrier between matrices (consumer-only sync)
            progs[w].append(("bar", bar_id))
        bar_id += 1
    pc = [0] * WARPS
    bar_arrived = {}
    for _ in range(max_steps):
        actors = []
        if inflight:
            actors.append("land")
        if prod_i < len(prod):
            it = prod[prod_i][0]
            if empty[it % n_stages].test(((it // n_stages) & 1) ^ 1):
                actors.append("prod")
        for w in range(WARPS):
            if pc[w] < len(progs[w]):
                kind, arg = progs[w][pc[w]]
                if kind == "prev":
                    if empty[arg % n_stages].test(((arg // n_stages) & 1) ^ 1):
                        actors.append(("prevpass", w))
                elif kind in ("stage", "kvstage"):
                    if full[arg % n_stages].test((arg // n_stages) & 1):
                        actors.append(("cons", w))
                else:
                    if w not in bar_arrived.setdefault(arg, set()):
                        actors.append(("bar", w))
                    elif len(bar_arrived[arg]) == WARPS:
                        actors.append(("barpass", w))
        if not actors:
            done = prod_i == len(prod) and all(pc[w] == len(progs[w]) for w in range(WARPS))
            return "ok" if done else "deadlock"
        a = rng.choice(actors)
        if a == "land":
            slot, it = inflight.pop(rng.randrange(len(inflight)))
            slot_data[slot] = it
            full[slot].arrive()
        elif a == "prod":
            it = prod[prod_i][0]
            inflight.append((it % n_stages, it))
            prod_i += 1
        elif a[0] == "cons":
            w = a[1]
            it = progs[w][pc[w]][1]
            if slot_data[it % n_stages] != it:
                return "stale"
            for _k in range(1 if progs[w][pc[w]][0] == "kvstage" else WARPS):
                empty[it % n_stages].arrive()
            pc[w] += 1
        elif a[0] == "prevpass":
            pc[a[1]] += 1
        elif a[0] == "bar":
            bar_arrived[progs[a[1]][pc[a[1]]][1]].add(a[1])
        else:
            pc[a[1]] += 1
    return "timeout"


SHAPES = [
    ([21, 14, 97, 14, 108], [1, 1, 1, 4, 1], 12),   # Mistral-7B slices of one CTA: QKV, wo, gate/up, down (4 chunks), lm head
    ([24, 17, 97, 17, 443], [2, 1, 2, 4, 2], 12),   # Nemo-like: dim 5120 -> 2 chunks
    ([3, 1, 7, 1, 2], [1, 1, 1, 1, 1], 12),         # tiny test model
    ([21, 14, 97, 14, 108], [1, 1, 1, 4, 1], 9),    # smallest legal ring
    ([21, 8, 14, 97, 14, 108], [1, 0, 1, 1, 4, 1], 12),  # with 8 K/V slice stages (all warps consume) between QKV and wo
    ([21, 60, 14, 97, 14, 108], [1, 0, 1, 1, 4, 1], 12),  # long context: the K/V slice is several laps of the ring
    ([24, 8, 17, 97, 17, 443], [2, 0, 1, 2, 4, 2], 9),
]


@pytest.mark.parametrize("pairs,nch,n_stages", SHAPES)
def test_grouped_schedule_is_safe(pairs, nch, n_stages):
    for seed in range(25):
        assert simulate(pairs, nch, n_stages, grouped=True, seed=seed) == "ok"


def test_unguarded_grouped_schedule_is_caught():
    """Without the empty[slot] guard the grouped order still aliases (out-of-order landing): the model must see it."""
    outcomes = {simulate([21, 14, 97, 14, 108], [1, 1, 1, 4, 1], 12, grouped=True, seed=s, guard=False) for s in range(10)}
    assert outcomes & {"stale", "deadlock"}


def test_naive_warp_per_pair_schedule_is_caught():
    """pair-major order with multi-chunk rows lets a warp run two laps ahead: the model must flag it."""
    outcomes = {simulate([21, 14, 97, 14, 108], [1, 1, 1, 4, 1], 12, grouped=False, seed=s) for s in range(10)}
    assert outcomes & {"stale", "deadlock"}
