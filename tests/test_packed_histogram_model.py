"""Model check (CPU, no GPU) of the packed 16-bit shared-memory counters of downsample_hist_kernel
(imagevis3d_b200/csrc/tvk_builder.cu, packed_hist_add): two bins share one 32-bit word, the add that sets bit 15 of
its half moves 0x8000 counts to the global bin, and the CTA adds at most 0x8000 voxels between two barriers.  The
model runs random interleavings of the threads' atomics (the move of a crossing add may be delayed until just before
the next barrier) and checks that no half ever carries into its neighbour, that no move borrows, and that
shared + global equals the true count -- the argument in the comment above packed_hist_add."""
import random

import pytest

MASK = 0xFFFFFFFF


def _run(seed, rounds, hot, eager=0.3, max_budget=0x8000):
    rng = random.Random(seed)
    word = 0                      # the packed word as the hardware sees it (mod 2^32 arithmetic)
    true = [0, 0]                 # what each half should hold if nothing carried (shared part only)
    spilled = [0, 0]
    total = [0, 0]
    for _ in range(rounds):
        # one round: the CTA's adds, at most 0x8000 voxels in all, split into chunks of 1..1024 (a warp-uniform add)
        budget = max_budget if rng.random() < 0.5 else rng.randrange(1, max_budget + 1)   # the residue wanders
        adds = []
        while budget > 0:
            cnt = min(budget, rng.choice([1, 1, 2, 3, 8, 8, 256, 1024]))
            half = 0 if rng.random() < hot else 1
            adds.append((half, cnt))
            budget -= cnt
        pending = []              # moves whose thread has not executed its atomicSub yet
        for half, cnt in adds:
            # a delayed move may run at any time before the barrier
            while pending and rng.random() < eager:
                crossed = pending.pop(rng.randrange(len(pending)))
                h = 0 if crossed == 0x8000 else 1
                assert true[h] >= 0x8000, "move would borrow"
                word = (word - crossed) & MASK
                true[h] -= 0x8000
                spilled[h] += 0x8000
            inc = cnt << (16 * half)
            old = word
            word = (word + inc) & MASK
            true[half] += cnt
            total[half] += cnt
            assert true[0] <= 0xFFFF and true[1] <= 0xFFFF, "a half overflowed"
            assert word == (true[0] | (true[1] << 16)), "carry between the halves"
            crossed = ~old & word & 0x80008000
            assert crossed in (0, 0x8000, 0x80000000)
            if crossed:
                assert not any((c == crossed) for c in pending), "two moves pending for one half"
                pending.append(crossed)
        for crossed in pending:   # every thread finishes its move before it reaches the barrier
            h = 0 if crossed == 0x8000 else 1
            assert true[h] >= 0x8000
            word = (word - crossed) & MASK
            true[h] -= 0x8000
            spilled[h] += 0x8000
        assert true[0] < 0x8000 and true[1] < 0x8000, "invariant at the barrier"
    assert [true[h] + spilled[h] for h in (0, 1)] == total
    assert word == (true[0] | (true[1] << 16))


@pytest.mark.parametrize("eager", [0.3, 0.0])          # 0.0: every move waits until just before the barrier
@pytest.mark.parametrize("hot", [1.0, 0.5, 0.9, 0.0])
def test_packed_counters_never_carry(hot, eager):
    for seed in range(40):
        _run(seed, rounds=24, hot=hot, eager=eager)


def test_model_catches_a_round_that_is_too_long():
    """The model has teeth: 0x9000 voxels between barriers do overflow a half when the moves are late."""
    caught = 0
    for seed in range(40):
        try:
            _run(seed, rounds=24, hot=1.0, eager=0.0, max_budget=0x9000)
        except AssertionError:
            caught += 1
    assert caught > 0
