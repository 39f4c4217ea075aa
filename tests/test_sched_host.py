"""Scheduling arithmetic of the persistent kernel (jacobiforcing_b200/csrc/jf_sched.h, compiled for the host through
tests/host/policy_harness.cpp): the k-split ranges tile the k-chunks, and the gate/up -> down hand-over is consistent -
the number of gate/up tiles that ARRIVE on a down k-split's counter equals the number its units WAIT for (a mismatch
would hang the kernel, or let a unit read gate/up features that are not written yet)."""
import ctypes as C

import pytest

from tests.test_policy_host import _load_harness


@pytest.fixture(scope="module")
def lib():
    lib = _load_harness()
    for name in ("jf_sched_split_k0", "jf_sched_split_k1", "jf_sched_tile_feeds_split", "jf_sched_tiles_needed"):
        getattr(lib, name).restype = C.c_int
    return lib


def test_k_splits_tile_the_chunks(lib):
    for k_chunks in list(range(1, 80)) + [148, 296, 592]:
        for splits in range(1, 9):
            if splits > k_chunks:
                continue
            edges = [(lib.jf_sched_split_k0(s, splits, k_chunks), lib.jf_sched_split_k1(s, splits, k_chunks))
                     for s in range(splits)]
            assert edges[0][0] == 0 and edges[-1][1] == k_chunks
            for (a0, a1), (b0, b1) in zip(edges, edges[1:]):
                assert a1 == b0 and a0 < a1
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1                      # balanced


@pytest.mark.parametrize("inter", [512, 1152, 2048, 4864, 8960, 9472, 18944, 18944 + 64, 27648])
def test_gate_up_to_down_handover_counts_agree(lib, inter):
    k_chunks = inter // 64                     # k-chunks of the down projection
    n_tiles = (inter + 127) // 128             # gate/up tiles (128 features = two k-chunks; the last may hold one)
    for splits in range(1, 9):
        if splits * 4 > k_chunks:
            continue
        for s in range(splits):
            k0, k1 = lib.jf_sched_split_k0(s, splits, k_chunks), lib.jf_sched_split_k1(s, splits, k_chunks)
            feeders = [t for t in range(n_tiles) if lib.jf_sched_tile_feeds_split(t, s, splits, k_chunks)]
            assert len(feeders) == lib.jf_sched_tiles_needed(k0, k1)
            # exactly the tiles whose k-chunks intersect [k0, k1)
            want = [t for t in range(n_tiles) if any(k0 <= k < k1 for k in (2 * t, 2 * t + 1) if k < k_chunks)]
            assert feeders == want
        # every tile feeds at least one split and at most two
        for t in range(n_tiles):
            n = sum(lib.jf_sched_tile_feeds_split(t, s, splits, k_chunks) for s in range(splits))
            assert 1 <= n <= 2
