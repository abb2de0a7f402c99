"""Host logic of Engine.fit: the visiting order of epoch e + 1 is drawn on a helper thread while the device runs epoch e
(deepimpute_b200.engine.PermutationPrefetcher).  It must hand out exactly what perm_fn(epoch) returns, in any access
pattern, and never ask for an epoch the run cannot reach."""
import threading

import numpy as np
import pytest

from deepimpute_b200.engine import PermutationPrefetcher, epoch_permutation


def test_prefetched_orders_are_the_orders():
    pf = PermutationPrefetcher(lambda e: epoch_permutation(7, e, 1000))
    try:
        for e in [0, 1, 2, 5, 6, 3]:                     # sequential use, a jump ahead, a jump back
            np.testing.assert_array_equal(pf.get(e), epoch_permutation(7, e, 1000))
    finally:
        pf.close()


def test_next_epoch_is_drawn_on_the_helper_thread_and_bounded_by_the_epoch_count():
    calls = []

    def perm_fn(e):
        calls.append((e, threading.current_thread().name))
        if e >= 3:
            raise IndexError("epoch beyond the run")
        return np.arange(4, dtype=np.int32) + e

    pf = PermutationPrefetcher(perm_fn, n_epochs=3)
    try:
        got = [pf.get(e).tolist() for e in range(3)]
    finally:
        pf.close()
    assert got == [[0, 1, 2, 3], [1, 2, 3, 4], [2, 3, 4, 5]]
    assert [e for e, _ in calls] == [0, 1, 2]             # nothing at or beyond n_epochs was asked for
    assert calls[0][1] == threading.current_thread().name and all(n.startswith("di-perm") for _, n in calls[1:])


def test_an_error_in_perm_fn_surfaces_when_that_epoch_is_used():
    def perm_fn(e):
        if e == 1:
            raise ValueError("bad epoch")
        return np.zeros(1, np.int32)

    pf = PermutationPrefetcher(perm_fn)
    try:
        pf.get(0)
        with pytest.raises(ValueError):
            pf.get(1)
    finally:
        pf.close()
