"""Host-side model of the fused two-stage launch's work list (dlux_b200/csrc/gemm_tc.cu: ``decode_fused``,
``gemm_tc_fused_ring``): persistent clusters walk ONE list of units in order (unit = cluster + i * n_clusters);
stage-2 units of an item wait for ``ready[item]`` (all its stage-1 units stored), stage-1 units wait for
``consumed[item - ring]`` (the ring slot's previous tenant read) before they store.  The kernel cannot be run
without a GPU, but its schedule can: these tests restate the list and check, over a grid of shapes, that every unit
appears exactly once, that every dependency points backwards in the list, and that in-order clusters always drain
it -- including the variant that was tried and rejected (deferring the ``ready`` arrival past the next unit's first
partial), which this model shows to deadlock in the tail of the list, as the GPU run did."""
import math

import pytest


def fused_ring(n_items, n1, n2, n_cl, waves=2.0):
    """lag / ring exactly as gemm_tc_fused_ring sizes them."""
    per_item = n1 + n2
    lag = int((waves * n_cl + per_item - 1) // per_item) + 1
    lag = min(lag, n_items)
    ring = min(lag + 2, n_items)
    return lag, ring


def decode_fused(unit, n_items, n1, n2, lag):
    """(stage, item, t) of list position `unit`, as the device function decodes it."""
    head = lag * n1
    if unit < head:
        return 0, unit // n1, unit % n1
    u = unit - head
    grp, n_mid = n1 + n2, n_items - lag
    if u < n_mid * grp:
        g, r = divmod(u, grp)
        return (0, lag + g, r) if r < n1 else (1, g, r - n1)
    u -= n_mid * grp
    return 1, n_mid + u // n2, u % n2


def drains(n_items, n1, n2, n_cl, deferred=False):
    """In-order clusters on the list; returns True when every cluster finishes."""
    lag, ring = fused_ring(n_items, n1, n2, n_cl)
    total = n_items * (n1 + n2)
    units = [decode_fused(u, n_items, n1, n2, lag) for u in range(total)]
    per = [units[c::n_cl] for c in range(n_cl)]
    pos, opened = [0] * n_cl, [False] * n_cl
    ready, consumed, pending = [0] * n_items, [0] * n_items, [None] * n_cl
    progress = True
    while progress:
        progress = False
        for c in range(n_cl):
            if pos[c] >= len(per[c]):
                if pending[c] is not None:
                    ready[pending[c]] += 1
                    pending[c] = None
                    progress = True
                continue
            stage, item, _ = per[c][pos[c]]
            if not opened[c]:                                   # first partial: needs the unit's operand
                if stage == 1 and ready[item] < n1:
                    continue
                opened[c] = progress = True
                if pending[c] is not None:                      # (rejected variant) the deferred arrival lands here
                    ready[pending[c]] += 1
                    pending[c] = None
            if stage == 0 and item >= ring and consumed[item - ring] < n2:
                continue                                        # epilogue: the ring slot is still being read
            if stage == 0:
                if deferred:
                    pending[c] = item
                else:
                    ready[item] += 1
            else:
                consumed[item] += 1
            pos[c] += 1
            opened[c] = False
            progress = True
    return all(pos[c] >= len(per[c]) for c in range(n_cl))


SHAPES = [(n_items, n1, n2, n_cl)
          for n_items in (1, 2, 3, 5, 8, 16, 64, 100)
          for n1 in (1, 2, 4, 16, 32)
          for n2 in (1, 2, 8, 32, 128)
          for n_cl in (74, 66, 8)]


def test_list_is_a_permutation_with_backward_dependencies():
    for n_items, n1, n2, n_cl in SHAPES:
        lag, ring = fused_ring(n_items, n1, n2, n_cl)
        assert 1 <= lag <= n_items and (lag < ring or ring == n_items)
        total = n_items * (n1 + n2)
        units = [decode_fused(u, n_items, n1, n2, lag) for u in range(total)]
        assert sorted(units) == sorted([(0, i, t) for i in range(n_items) for t in range(n1)] +
                                       [(1, i, t) for i in range(n_items) for t in range(n2)])
        last_s1 = {}
        first_s2, last_s2 = {}, {}
        for p, (s, i, _) in enumerate(units):
            if s == 0:
                last_s1[i] = p
            else:
                first_s2.setdefault(i, p)
                last_s2[i] = p
        for i in range(n_items):
            assert last_s1[i] < first_s2[i]                     # ready[i] is complete by units before S2(i)
            if i >= ring:                                       # consumed[i - ring] by units before S1(i) ends
                assert last_s2[i - ring] < last_s1[i]


def test_in_order_clusters_always_drain_the_list():
    for shape in SHAPES:
        assert drains(*shape), shape


def test_c3_and_c4_shapes():
    # C3 forward (F1 16 units, F2 8 units per item) and adjoint (16, 32); C4 (2048 -> 256, 4096 items)
    for shape in ((64, 16, 8, 74), (64, 16, 32, 74), (4096, 16, 2, 74), (4096, 16, 128, 74)):
        assert drains(*shape)


def test_deferred_ready_arrival_deadlocks_in_the_tail():
    """The rejected optimisation (DESIGN 8 / profiles/r2_summary.md): at C3's forward shape the last stage-1 unit of
    an item and its stage-2 units land on the same cluster back to back, so an arrival deferred into the next unit
    waits for itself."""
    assert drains(64, 16, 8, 74, deferred=False)
    assert not drains(64, 16, 8, 74, deferred=True)
