"""CPU model-check of the stream-K weight-streaming GEMM's work partition and split-tile reduction (csrc/gemm_streamk.cuh).

The kernel cannot run here, but the arithmetic it relies on can: the (tile, k-block) units of a GEMM form one sequence, CTA c takes
units [first(c), first(c + 1)), a range is cut into segments at tile boundaries, and each segment plays one of three roles --
whole tile, CONTRIBUTOR (does not hold the tile's first k-block: parks its accumulator in the CTA's single workspace slot and raises
the CTA's single flag) or OWNER (holds the first k-block of a tile somebody else finishes: waits for the flags of CTAs cta+1 .. last,
sums their slots in ascending order, stores, resets the flags).  The properties checked are the ones the device code assumes:

  * the ranges tile the unit sequence exactly, and differ by at most one unit;
  * a CTA has at most ONE contributor segment and it is the FIRST segment of its range (one slot and one flag per CTA suffice,
    and a contributor that also owns a tile contributes before it waits: no cycle);
  * a CTA has at most ONE owner segment and it is the LAST segment of its range (the shared-memory ring is idle when the owner
    stages the contributors' sum in it);
  * the owner's `last` loop finds exactly the CTAs that contribute to its tile, in ascending k order, and they cover the tile;
  * a simulation with the flags (random interleaving of CTAs, owners blocking on their contributors) never deadlocks, leaves every
    flag reset, and the summation order reproduces the k order.
Shapes: every decode-sized Linear of the BASELINE configs at 148 SMs, plus random ones."""
import random

import pytest

TG_BK, SK_BN = 64, 128


def first(total, G, c):
    per, rem = divmod(total, G)
    return c * per + min(c, rem)


def segments(total, num_k, G, c):
    """[(tile, kb0, kb1)] of CTA c in processing order -- the loop of sk_gemm_body's MMA and epilogue roles."""
    u, end, out = first(total, G, c), first(total, G, c + 1), []
    while u < end:
        kb0 = u % num_k
        kb1 = min(num_k, kb0 + (end - u))
        out.append((u // num_k, kb0, kb1))
        u += kb1 - kb0
    return out


def owner_last(total, num_k, G, cta, tile):
    """The owner's scan: the highest CTA whose range starts inside `tile`."""
    tile_end = (tile + 1) * num_k
    last = cta
    while last + 1 < G and first(total, G, last + 1) < tile_end:
        last += 1
    return last


SHAPES = [  # (N, K) of the decode-sized linears: 7B, Nemo-12B, Mixtral-8x7B expert, Mixtral-8x22B, lm heads
    (6144, 4096), (4096, 4096), (28672, 4096), (4096, 14336), (6144, 5120), (5120, 4096), (28672, 5120), (5120, 14336),
    (8192, 6144), (6144, 6144), (32768, 6144), (6144, 16384), (32000, 4096), (131072, 5120), (32768, 4096), (128, 64), (256, 4096),
]
rng = random.Random(7)
SHAPES += [(SK_BN * rng.randint(1, 300), TG_BK * rng.randint(1, 300)) for _ in range(40)]


@pytest.mark.parametrize("G", [148, 132, 160, 7])
def test_partition_and_roles(G):
    for N, K in SHAPES:
        num_n, num_k = N // SK_BN, K // TG_BK
        total = num_n * num_k
        grid = min(G, total)  # launch_streamk_ta
        covered = 0
        contributors = {}  # tile -> [(kb0, cta)]
        owners = {}        # tile -> (cta, kb1)
        sizes = []
        for c in range(grid):
            segs = segments(total, num_k, grid, c)
            sizes.append(sum(b - a for _, a, b in segs))
            assert segs, "an idle CTA would never raise the counters the launch relies on"
            for i, (tile, kb0, kb1) in enumerate(segs):
                assert first(total, grid, c) + sum(b - a for _, a, b in segs[:i]) == tile * num_k + kb0
                covered += kb1 - kb0
                if kb0 != 0:
                    assert i == 0, "a contributor segment must be the first thing a CTA does (one slot, one flag, no wait cycle)"
                    contributors.setdefault(tile, []).append((kb0, c))
                elif kb1 != num_k:
                    assert i == len(segs) - 1, "an owner with contributors must be at the end of its range (the ring is staging space)"
                    owners[tile] = (c, kb1)
        assert covered == total and max(sizes) - min(sizes) <= 1
        assert set(contributors) == set(owners)
        for tile, (c, kb1) in owners.items():
            last = owner_last(total, num_k, grid, c, tile)
            parts = sorted(contributors[tile])
            assert [cta for _, cta in parts] == list(range(c + 1, last + 1)), (N, K, tile)
            k = kb1  # contributors continue where the owner stopped, without gaps, up to the end of the tile
            for kb0, cta in parts:
                assert kb0 == k
                k = [s for s in segments(total, num_k, grid, cta) if s[0] == tile][0][2]
            assert k == num_k


@pytest.mark.parametrize("seed", range(6))
def test_flag_protocol_never_deadlocks(seed):
    """CTAs advance one segment at a time in random order; an owner blocks until the flags of cta+1 .. last are up.  Every tile
    must come out as the k-ordered sum of its parts, every flag must be down at the end (the next launch starts from zero)."""
    r = random.Random(seed)
    N, K = r.choice(SHAPES[:15])
    G = r.choice([148, 132, 37])
    num_k = K // TG_BK
    total = (N // SK_BN) * num_k
    grid = min(G, total)
    segs = [segments(total, num_k, grid, c) for c in range(grid)]
    pos = [0] * grid
    flags, slots, result = [0] * grid, [None] * grid, {}
    pending = set(range(grid))
    stalled = 0
    while pending:
        c = r.choice(sorted(pending))
        tile, kb0, kb1 = segs[c][pos[c]]
        part = list(range(kb0, kb1))  # stands for the fp32 partial sum over these k-blocks
        if kb0 != 0:
            assert flags[c] == 0 and slots[c] is None
            slots[c], flags[c] = part, 1
        elif kb1 != num_k:
            last = owner_last(total, num_k, grid, c, tile)
            if not all(flags[o] for o in range(c + 1, last + 1)):
                stalled += 1
                assert stalled < 200 * grid, "owner waits forever: the contributors never come"
                continue
            acc = []
            for o in range(c + 1, last + 1):  # ascending CTA index == ascending k
                acc += slots[o]
                slots[o], flags[o] = None, 0
            result[tile] = part + acc
        else:
            result[tile] = part
        stalled = 0
        pos[c] += 1
        if pos[c] == len(segs[c]):
            pending.discard(c)
    assert not any(flags) and all(s is None for s in slots)
    assert all(result[t] == list(range(num_k)) for t in range(N // SK_BN))
