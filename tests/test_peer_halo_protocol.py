"""Model check of the mailbox protocol of ``utils/peer.py:PeerHalo`` (no GPU needed).

Two neighbouring ranks run the same program of halo exchanges on in-order streams:
``push(e)`` stores into the neighbour's mailbox slot ``e % SLOTS`` and raises its flag, ``wait(e)`` blocks
until the neighbour's ``push(e)`` has run, ``unpack(e)`` reads the own slot.  The only synchronisation
between the ranks is ``wait``.  Every interleaving of the two streams is explored and a push that lands
in a slot whose previous content (exchange ``e - SLOTS``) the receiver has not unpacked yet is a failure.
This is the argument of the class docstring, executed: four slots are enough when at most two exchanges
are in flight, two slots are enough for one, and two slots with two in flight is exactly the race the
2-GPU parity worker caught in round 2."""
import itertools

import pytest


def _program(pattern):
    """pattern: in-flight counts per group, e.g. (1, 2, 1): exchange 1 alone, then 2 and 3 together ..."""
    ops, e = [], 0
    for k in pattern:
        group = list(range(e + 1, e + 1 + k))
        e += k
        ops += [("push", x) for x in group]
        for x in group:
            ops += [("wait", x), ("unpack", x)]
    return ops


def _overwrite_possible(pattern, slots):
    prog = _program(pattern)
    n = len(prog)
    seen = set()
    stack = [(0, 0)]
    while stack:
        state = stack.pop()
        if state in seen:
            continue
        seen.add(state)
        pcs = list(state)
        for r in (0, 1):
            if pcs[r] == n:
                continue
            op, e = prog[pcs[r]]
            other_done = {prog[i] for i in range(pcs[1 - r])}
            if op == "wait" and ("push", e) not in other_done:
                continue  # blocked on the neighbour's flag
            if op == "push" and e > slots and ("unpack", e - slots) not in other_done:
                return True  # the neighbour still holds exchange e - slots in this slot
            nxt = pcs.copy()
            nxt[r] += 1
            stack.append(tuple(nxt))
    return False


@pytest.mark.parametrize("pattern", [p for k in (3, 4, 5, 6) for p in itertools.product((1, 2), repeat=k)][::3])
def test_four_slots_are_enough_for_at_most_two_exchanges_in_flight(pattern):
    assert not _overwrite_possible(pattern, 4)


def test_two_slots_are_enough_for_one_exchange_at_a_time_but_not_for_two():
    assert not _overwrite_possible((1,) * 8, 2)
    assert _overwrite_possible((2, 2, 2), 2)      # the round-2 race
    assert _overwrite_possible((2, 2, 2, 2), 3)   # three slots are not enough either


def test_the_implementation_uses_the_checked_parameters():
    from sopht_mpi_b200.utils.peer import PeerHalo

    assert PeerHalo.SLOTS == 4 and PeerHalo.MAX_IN_FLIGHT == 2
