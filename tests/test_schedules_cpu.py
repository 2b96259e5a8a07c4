"""The scheduling rules of csrc/eig_cluster.cu restated in a few lines of Python and checked exhaustively on the CPU:
(1) the odd-even ordering with one column per pair in registers visits every column pair exactly once per sweep,
(2) the push rule of the cluster form (the writer stores the rotated column into the CTA of the group that reads it
next) always leaves the newest version of a column where its reader looks, and (3) the parallel fix-up rounds
(columns claimed by list index) give bit for bit the result of rotating the listed pairs one after the other."""
import itertools

import numpy as np
import pytest


def odd_even_sweep(D, nc=1, check_push_rule=True):
    """One sweep.  Position p holds column pos[p]; group g keeps position 2g+1 'in registers' and reads position
    2g (even steps) / 2g+2 (odd steps) from ITS CTA's copy of the slots; after the rotation the two columns swap."""
    ng = D // 2
    gpc = (ng + nc - 1) // nc
    cta = lambda g: g // gpc
    pos = list(range(D))                       # ground truth: column at every position
    # slots[c][p] = (column, version) as CTA c sees even position p; every CTA starts from the full matrix
    slots = [{p: (pos[p], 0) for p in range(0, D, 2)} for _ in range(nc)]
    newest = {p: 0 for p in range(0, D, 2)}
    met = []
    for s in range(D):
        odd = s & 1
        writes = []
        for g in range(ng):
            e = 2 * g + 2 * odd
            if e >= D:
                continue                       # the last group has no partner in the odd steps
            col, ver = slots[cta(g)][e]
            if check_push_rule:
                assert ver == newest[e] and col == pos[e], (s, g, e)
            reg = pos[2 * g + 1]
            met.append(frozenset((reg, col)))
            # swap: the former register column goes to position e, the partner stays in registers
            reader = g + 1 if odd else max(g - 1, 0)
            writes.append((cta(reader) if (odd or g > 0) else cta(g), e, reg, 2 * g + 1, col))
        for dst, e, reg, o, col in writes:     # all groups of a step work on the old state, then the barrier
            newest[e] += 1
            slots[dst][e] = (reg, newest[e])
            pos[e], pos[o] = reg, col
    return met, pos


@pytest.mark.parametrize("D,nc", [(4, 1), (6, 1), (20, 1), (20, 2), (20, 4), (100, 1), (100, 2), (100, 4), (164, 4), (200, 4), (16, 4)])
def test_odd_even_ordering_meets_every_pair_once_and_the_push_rule_feeds_every_reader(D, nc):
    met, pos = odd_even_sweep(D, nc)
    assert len(met) == D * (D - 1) // 2
    assert set(met) == {frozenset(p) for p in itertools.combinations(range(D), 2)}
    assert pos == list(range(D))[::-1]         # a sweep reverses the order (the next sweep starts from there)


def test_check_rows_cover_every_pair_once():
    """The post-sweep check: group t of CTA r takes rows pp = r + nc t, ... and D - 1 - pp, each against the columns to
    its right."""
    for D, nc, lgroups in [(100, 1, 52), (100, 4, 13), (20, 2, 3), (200, 4, 25)]:
        seen = []
        for rank in range(nc):
            for lg in range(lgroups):
                for pp in range(rank + nc * lg, D // 2, nc * lgroups):
                    for p in (pp, D - 1 - pp):
                        seen += [(p, q) for q in range(p + 1, D)]
        assert sorted(seen) == sorted(itertools.combinations(range(D), 2))


def rotate(M, p, q, theta):
    c, s = np.cos(theta), np.sin(theta)
    a, b = M[:, p].copy(), M[:, q].copy()
    M[:, p] = c * a - s * b
    M[:, q] = s * a + c * b


@pytest.mark.parametrize("seed", range(6))
def test_parallel_fix_up_rounds_equal_the_sequential_pass(seed):
    rng = np.random.default_rng(seed)
    D, n = 24, 40
    pairs = sorted({tuple(sorted(rng.choice(D, 2, replace=False))) for _ in range(n)})   # sorted list, as the kernel
    # the rotation of a pair depends on the CURRENT columns (as the real one does through their dot product)
    angle = lambda M, p, q: 0.3 * float(M[:, p] @ M[:, q])
    M0 = rng.standard_normal((D, D))
    seq = M0.copy()
    for p, q in pairs:
        rotate(seq, p, q, angle(seq, p, q))
    par = M0.copy()
    done = [False] * len(pairs)
    rounds = 0
    while not all(done):
        owner = {}
        for e, (p, q) in enumerate(pairs):     # atomicMin of the list index on both columns
            if not done[e]:
                owner[p] = min(owner.get(p, e), e)
                owner[q] = min(owner.get(q, e), e)
        ready = [e for e, (p, q) in enumerate(pairs) if not done[e] and owner[p] == e and owner[q] == e]
        assert ready                            # the lowest pending index always holds both of its columns
        snapshot = par.copy()
        for e in ready:                         # disjoint columns: all of a round work on the same state
            p, q = pairs[e]
            rotate(par, p, q, angle(snapshot, p, q))
            done[e] = True
        rounds += 1
    assert np.array_equal(par, seq)
    assert rounds <= len(pairs)
