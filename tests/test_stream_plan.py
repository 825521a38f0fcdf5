"""Host-side planning of the bounded-footprint feed (stream.plan_windows / _chunks_of): no GPU needed."""
import numpy as np
import pytest

from aggfly_b200 import stream


@pytest.mark.parametrize("seed", range(20))
def test_windows_cover_the_axis_end_on_cuts_and_respect_the_slot_where_cuts_allow(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(50, 5000))
    cuts = np.unique(np.concatenate([rng.integers(1, n, size=int(rng.integers(0, 40))), [n]]))
    slot = int(rng.integers(5, 400))
    ws = stream.plan_windows(cuts, n, slot)
    assert ws[0][0] == 0 and ws[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(ws[:-1], ws[1:]))
    assert all(r1 in set(cuts.tolist()) and r1 > r0 for r0, r1 in ws)
    for r0, r1 in ws:
        if r1 - r0 > slot:          # only when no cut lies inside the slot: then the window is the shortest possible
            assert not np.any((cuts > r0) & (cuts < r1))
        else:                       # greedy: the next cut would not have fitted
            nxt = cuts[cuts > r1]
            assert len(nxt) == 0 or nxt[0] - r0 > slot


def test_chunks_never_straddle_a_source_break():
    breaks = np.array([17, 40, 41])
    cs = stream._chunks_of(5, 60, row_bytes=8, chunk_bytes=80, breaks=breaks)
    assert cs[0][0] == 5 and cs[-1][1] == 60 and all(a[1] == b[0] for a, b in zip(cs[:-1], cs[1:]))
    assert all(b - a <= 10 for a, b in cs)
    for br in breaks:
        assert not any(a < br < b for a, b in cs)
