"""Model check of the flag protocol behind halo_p2p_kernel / p2p_gather_kernel (csrc/comm.cu), on the CPU.

Every rank runs the same sequence of exchanges.  In exchange e it (1) stores its values into slot e mod S of each
neighbour's staging buffer, (2) raises its flag at each neighbour to e, (3) waits until every neighbour's flag
here is >= e, (4) reads slot e mod S.  The kernels use S = 2 slots and rely on SYMMETRIC neighbour lists: a rank
can only start exchange e+1 after it consumed e, and it needs this rank's e+1 signal before it can reach e+2 --
so nobody overwrites a slot that has not been read.  The scheduler below interleaves the ranks' atomic actions at
random (including adversarial bursts) and checks that every read sees exactly the value written for that epoch;
it also shows that the check has teeth: one slot, or one-directional neighbour lists, do get corrupted."""
import random

import pytest


def simulate(n_ranks, neighbours, n_slots, n_exchanges, seed, bursts=True):
    """neighbours[r] = ranks r sends to AND waits for.  Returns (violations, finished)."""
    rng = random.Random(seed)
    slot = {(dst, src, s): None for dst in range(n_ranks) for src in range(n_ranks) for s in range(n_slots)}
    flag = {(dst, src): 0 for dst in range(n_ranks) for src in range(n_ranks)}
    # program counter per rank: (epoch, phase, index); phases: 0 push, 1 signal, 2 wait, 3 read
    pc = [[1, 0, 0] for _ in range(n_ranks)]
    violations = 0
    steps = 0
    while any(p[0] <= n_exchanges for p in pc) and steps < 200000:
        steps += 1
        r = rng.randrange(n_ranks)
        burst = rng.choice([1, 1, 1, 7, 40]) if bursts else 1
        for _ in range(burst):
            e, ph, i = pc[r]
            if e > n_exchanges:
                break
            nb = neighbours[r]
            if ph == 0:
                if i < len(nb):
                    slot[(nb[i], r, e % n_slots)] = (r, e)
                    pc[r][2] += 1
                else:
                    pc[r][1:] = [1, 0]
            elif ph == 1:
                if i < len(nb):
                    flag[(nb[i], r)] = e
                    pc[r][2] += 1
                else:
                    pc[r][1:] = [2, 0]
            elif ph == 2:
                if all(flag[(r, p)] >= e for p in nb):
                    pc[r][1:] = [3, 0]
                else:
                    break                       # blocked: let somebody else run
            else:
                if i < len(nb):
                    if slot[(r, nb[i], e % n_slots)] != (nb[i], e):
                        violations += 1
                    pc[r][2] += 1
                else:
                    pc[r] = [e + 1, 0, 0]
    return violations, all(p[0] > n_exchanges for p in pc)


def ring(n):
    return {r: sorted({(r - 1) % n, (r + 1) % n} - {r}) for r in range(n)}


def all_to_all(n):
    return {r: [p for p in range(n) if p != r] for r in range(n)}


@pytest.mark.parametrize("topology", ["ring8", "all8", "pair", "star5"])
def test_two_slots_with_symmetric_neighbours_never_corrupt(topology):
    nb = {"ring8": ring(8), "all8": all_to_all(8), "pair": {0: [1], 1: [0]},
          "star5": {0: [1, 2, 3, 4], 1: [0], 2: [0], 3: [0], 4: [0]}}[topology]
    for seed in range(40):
        bad, done = simulate(len(nb), nb, 2, 30, seed)
        assert done, "deadlock"
        assert bad == 0


def test_the_model_detects_a_single_slot_and_one_directional_lists():
    """One slot is overwritten by a neighbour that is one exchange ahead; and if rank 0 sends to rank 1 without
    waiting for anything from it (a one-directional neighbour list), nothing holds the sender back and it
    overwrites slots rank 1 has not read yet -- which is why HaloPlan::ensure_p2p insists on symmetric lists."""
    assert sum(simulate(8, ring(8), 1, 30, seed)[0] for seed in range(40)) > 0
    assert sum(simulate_asymmetric(send={0: [1], 1: [0]}, wait={0: [], 1: [0]}, seed=seed) for seed in range(40)) > 0


def simulate_asymmetric(send, wait, seed):
    """Two ranks, two slots, rank r sends to send[r] but only waits for wait[r]."""
    rng = random.Random(seed)
    slot, flag = {}, {(1, 0): 0, (0, 1): 0}
    pc = {0: 1, 1: 1}
    bad, E = 0, 30
    for _ in range(5000):
        r = 0 if rng.random() < 0.8 else 1          # the unthrottled sender runs far more often
        e = pc[r]
        if e > E:
            continue
        if all(flag[(r, p)] >= e - 1 for p in wait[r]) or e == 1:
            for p in send[r]:
                slot[(p, r, e % 2)] = (r, e)
                flag[(p, r)] = e
            if all(flag[(r, p)] >= e for p in wait[r]):
                for p in wait[r]:
                    if slot.get((r, p, e % 2)) != (p, e):
                        bad += 1
                pc[r] = e + 1
    return bad
