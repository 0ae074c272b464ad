"""The exchange protocol of csrc/peer.cuh (k_peer_allreduce) as a state machine per rank, run under random
interleavings: slot and flag writes to each peer, the wait on the rank's own flags, the slot reads and the advance of
the exchange counter are separate steps that a scheduler orders at random (per-rank program order kept, as a CUDA
stream keeps it).  Checked: every value a rank adds in exchange s is the one its peer wrote for exchange s -- i.e. two
parities of slots and flags are enough, however far the ranks drift apart -- and nobody deadlocks.  A one-parity
variant of the same machine must fail the check (the test of the test)."""
import random


def run(world, exchanges, parities, seed, slow_rank=None):
    rng = random.Random(seed)
    slots = [[[None] * world for _ in range(parities)] for _ in range(world)]   # slots[owner][parity][src]
    flags = [[[0] * world for _ in range(parities)] for _ in range(world)]
    done = [0] * world
    sums = [[] for _ in range(world)]

    def program(r):
        for s in range(1, exchanges + 1):
            par = s % parities
            for p in range(world):                      # 1. the rank's value into its slot on every peer
                yield ("store", p, par, s)
            for p in range(world):                      # 2. ... then the flags (release after the stores)
                yield ("flag", p, par, s)
            yield ("wait", None, par, s)                 # 3. own flags of all ranks >= s
            for p in range(world):                      # 4. the slots, in rank order
                yield ("load", p, par, s)
            yield ("advance", None, par, s)

    progs = [program(r) for r in range(world)]
    pending = [next(g) for g in progs]
    acc = [0] * world
    live = set(range(world))
    idle = 0
    while live:
        weights = [0.05 if r == slow_rank else 1.0 for r in sorted(live)]
        r = rng.choices(sorted(live), weights)[0]
        op, p, par, s = pending[r]
        if op == "wait" and not all(flags[r][par][q] >= s for q in range(world)):
            idle += 1
            assert idle < 200000, "deadlock"
            continue
        idle = 0
        if op == "store":
            slots[p][par][r] = (r, s)
        elif op == "flag":
            flags[p][par][r] = s
        elif op == "load":
            got = slots[r][par][p]
            if got != (p, s):
                return False                             # a slot was overwritten (or not yet written) when read
            acc[r] += 1
        elif op == "advance":
            done[r] = s
            sums[r].append(acc[r])
        try:
            pending[r] = next(progs[r])
        except StopIteration:
            live.discard(r)
    return all(d == exchanges for d in done)


def test_two_parities_are_enough_under_any_interleaving():
    for world in (2, 3, 8):
        for seed in range(40):
            assert run(world, 12, 2, seed), (world, seed)
        for slow in range(min(world, 3)):
            for seed in range(10):
                assert run(world, 12, 2, 1000 + seed, slow_rank=slow), (world, slow, seed)


def test_one_parity_is_not():
    assert not all(run(3, 12, 1, seed, slow_rank=seed % 3) for seed in range(60))
