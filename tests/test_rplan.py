"""Host side of the one-kernel temporal + regional path (csrc/agf_rplan.cu): the CSR lowered onto 8 x 32 cell tiles.

``agf_rplan_tables`` is the host-only twin of ``agf_rplan_create``; here its tables are walked exactly the way the kernel
walks them (per tile: slots -> entries in order -> partial sums; per region: partial sums added in ascending slot order)
in NumPy, and the result must equal the reference's scatter (aggfly/aggregate/spatial.py:181-186 as restated by the oracle)
-- bit for bit for regions inside one tile, to re-association error otherwise."""
import ctypes as C

import numpy as np
import pytest

from aggfly_b200 import _lib


def _tables(n_regions, n_lat, n_lon, row_ptr, cell_idx, w):
    L = _lib.lib()
    info = _lib.RPlanInfo()
    rp, ci, ww = (np.ascontiguousarray(row_ptr, np.int32), np.ascontiguousarray(cell_idx, np.int32),
                  np.ascontiguousarray(w, np.float64))
    args = (n_regions, n_lat, n_lon, len(ci), rp.ctypes.data, ci.ctypes.data, ww.ctypes.data)
    _lib.check(L.agf_rplan_tables(*args, C.byref(info), *([None] * 9)))
    t = dict(tile_ids=np.zeros(info.n_active_tiles, np.int32), tile_slot_ptr=np.zeros(info.n_active_tiles + 1, np.int32),
             slot_region=np.zeros(info.n_slots, np.int32), slot_ent_ptr=np.zeros(info.n_slots + 1, np.int32),
             entry_cell=np.zeros(info.n_entries, np.int32), entry_w=np.zeros(info.n_entries, np.float64),
             region_slot_ptr=np.zeros(n_regions + 1, np.int32), region_slots=np.zeros(info.n_slots, np.int32),
             slot_dst=np.zeros(info.n_slots, np.int32))
    _lib.check(L.agf_rplan_tables(*args, C.byref(info), *[t[k].ctypes.data for k in (
        "tile_ids", "tile_slot_ptr", "slot_region", "slot_ent_ptr", "entry_cell", "entry_w", "region_slot_ptr", "region_slots",
        "slot_dst")]))
    return info, t


def _random_csr(rng, n_regions, n_lat, n_lon, mean_len, empty=()):
    rows = []
    for r in range(n_regions):
        if r in empty:
            rows.append(np.zeros(0, np.int64))
            continue
        # a compact blob of cells around a random centre, in random (weights-frame) order, some cells shared with neighbours
        cy, cx = rng.integers(0, n_lat), rng.integers(0, n_lon)
        n = max(1, int(rng.poisson(mean_len)))
        yy = np.clip(cy + rng.integers(-3, 4, n), 0, n_lat - 1)
        xx = (cx + rng.integers(-6, 7, n)) % n_lon
        rows.append(np.unique(yy * n_lon + xx)[rng.permutation(len(np.unique(yy * n_lon + xx)))])
    row_ptr = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int32)
    cell_idx = np.concatenate(rows).astype(np.int32)
    w = rng.uniform(0.01, 1.0, len(cell_idx))
    return row_ptr, cell_idx, w


@pytest.mark.parametrize("shape", [(16, 64), (21, 100), (8, 32), (5, 20), (40, 236)])
def test_tables_reproduce_the_scatter(shape):
    n_lat, n_lon = shape
    rng = np.random.default_rng(n_lat * 1000 + n_lon)
    R = 37
    row_ptr, cell_idx, w = _random_csr(rng, R, n_lat, n_lon, 14, empty=(5, 36))
    info, t = _tables(R, n_lat, n_lon, row_ptr, cell_idx, w)
    tiles_x = (n_lon + 31) // 32
    assert info.n_tiles == tiles_x * ((n_lat + 7) // 8) and info.tile_lat == 8 and info.tile_lon == 32
    assert info.n_entries == len(cell_idx) and info.n_empty_regions == 2
    assert np.all(np.diff(t["tile_ids"]) > 0)                                   # time-major units walk tiles in raster order
    # every entry comes back exactly once, with its weight, inside the slot of its (tile, region)
    x = rng.normal(size=n_lat * n_lon)
    x[rng.random(x.size) < 0.1] = np.nan
    valid = ~np.isnan(x)
    part = np.zeros((info.n_slots, 2))
    seen = []
    for ti, tile in enumerate(t["tile_ids"]):
        ty, tx = divmod(int(tile), tiles_x)
        for s in range(t["tile_slot_ptr"][ti], t["tile_slot_ptr"][ti + 1]):
            r = t["slot_region"][s]
            num = den = 0.0
            for e in range(t["slot_ent_ptr"][s], t["slot_ent_ptr"][s + 1]):
                ly, lx = divmod(int(t["entry_cell"][e]), 32)
                cell = (ty * 8 + ly) * n_lon + tx * 32 + lx
                assert ty * 8 + ly < n_lat and tx * 32 + lx < n_lon
                seen.append((r, cell, t["entry_w"][e]))
                num += t["entry_w"][e] * (x[cell] if valid[cell] else 0.0)      # an invalid cell is staged as zeros
                den += t["entry_w"][e] * (1.0 if valid[cell] else 0.0)
            part[s] = num, den
    want_entries = sorted((r, int(c), float(ww)) for r in range(R) for c, ww in
                          zip(cell_idx[row_ptr[r]:row_ptr[r + 1]], w[row_ptr[r]:row_ptr[r + 1]]))
    assert sorted((int(r), int(c), float(ww)) for r, c, ww in seen) == want_entries
    # weights-frame order is kept inside a slot
    for s in range(info.n_slots):
        r = t["slot_region"][s]
        order = {int(c): k for k, c in enumerate(cell_idx[row_ptr[r]:row_ptr[r + 1]])}
        ws = t["entry_w"][t["slot_ent_ptr"][s]:t["slot_ent_ptr"][s + 1]]
        ks = [np.flatnonzero(w[row_ptr[r]:row_ptr[r + 1]] == v)[0] for v in ws]
        assert ks == sorted(ks)
    # slots of a tile come longest first; a slot that holds its whole region names the region, the others a partial row
    for ti in range(info.n_active_tiles):
        lens = np.diff(t["slot_ent_ptr"][t["tile_slot_ptr"][ti]:t["tile_slot_ptr"][ti + 1] + 1])
        assert np.all(np.diff(lens) <= 0) and np.all(lens > 0)
    n_slots_of = np.diff(t["region_slot_ptr"])
    whole = n_slots_of[t["slot_region"]] == 1
    assert np.array_equal(t["slot_dst"][whole], t["slot_region"][whole])
    assert np.array_equal(np.sort(-t["slot_dst"][~whole] - 1), np.arange(info.n_partial_rows))
    assert info.n_partial_rows == int((~whole).sum())
    # regions: partial rows in ascending slot order; one-tile regions match the sequential sum bit for bit
    n_single = 0
    for r in range(R):
        slots = t["region_slots"][t["region_slot_ptr"][r]:t["region_slot_ptr"][r + 1]]
        assert np.all(np.diff(slots) > 0) and np.all(t["slot_region"][slots] == r)
        num = den = 0.0
        for s in slots:
            num, den = num + part[s, 0], den + part[s, 1]
        wn = wd = 0.0
        for c, ww in zip(cell_idx[row_ptr[r]:row_ptr[r + 1]], w[row_ptr[r]:row_ptr[r + 1]]):
            if valid[c]:
                wn, wd = wn + ww * x[c], wd + ww
        if len(slots) == 0:
            assert row_ptr[r + 1] == row_ptr[r]
        elif len(slots) == 1:
            n_single += 1
            assert (num, den) == (wn, wd)
        else:
            assert abs(num - wn) <= 1e-13 * max(1.0, abs(wn)) and abs(den - wd) <= 1e-13 * max(1.0, wd)
    assert n_single > 0 or n_lat * n_lon > 256


def test_bad_arguments_are_rejected():
    L = _lib.lib()
    rp = np.array([0, 2], np.int32)
    ci = np.array([0, 999], np.int32)
    w = np.ones(2)
    info = _lib.RPlanInfo()
    with pytest.raises(_lib.AgfError, match="outside the grid"):
        _lib.check(L.agf_rplan_tables(1, 4, 8, 2, rp.ctypes.data, ci.ctypes.data, w.ctypes.data, C.byref(info), *([None] * 9)))
    rp2 = np.array([0, 1], np.int32)
    with pytest.raises(_lib.AgfError, match="row_ptr"):
        _lib.check(L.agf_rplan_tables(1, 4, 8, 2, rp2.ctypes.data, ci.ctypes.data, w.ctypes.data, C.byref(info), *([None] * 9)))


@pytest.mark.parametrize("lps", [4, 8, 16])
@pytest.mark.parametrize("shape,R,mean_len", [((16, 64), 37, 14), ((21, 100), 90, 6), ((8, 32), 300, 2), ((40, 236), 25, 120),
                                              ((5, 20), 3, 60)])
def test_balanced_walk_tables_visit_every_entry_once_and_in_order(shape, R, mean_len, lps):
    """agf_rplan_check_segments builds the per-lane-group segment tables of the kernel variant with ``lps`` lanes per slot
    and verifies them in the library; the loads of the lane groups must be close to the mean."""
    n_lat, n_lon = shape
    rng = np.random.default_rng(n_lat * 77 + R + lps)
    row_ptr, cell_idx, w = _random_csr(rng, R, n_lat, n_lon, mean_len, empty=(1,))
    stats = np.zeros(7, np.int64)
    rp, ci, ww = np.ascontiguousarray(row_ptr, np.int32), np.ascontiguousarray(cell_idx, np.int32), np.ascontiguousarray(w)
    _lib.check(_lib.lib().agf_rplan_check_segments(R, n_lat, n_lon, len(ci), rp.ctypes.data, ci.ctypes.data, ww.ctypes.data,
                                                   lps, stats.ctypes.data))
    n_segs, n_pent, max_load, mean_load, n_tiles, max_segs, max_pent = (int(v) for v in stats)
    assert n_pent >= len(ci) and n_pent % 4 == 0 and n_pent <= len(ci) + 3 * n_segs
    assert n_segs >= 1 and max_segs <= n_segs and max_pent <= n_pent
    # no group carries more than the mean plus one maximal segment (LPT bound), segments are at most ~ entries / groups
    info, t = _tables(R, n_lat, n_lon, row_ptr, cell_idx, w)
    per_tile = np.diff(t["slot_ent_ptr"][t["tile_slot_ptr"]])
    ng = 256 // lps
    c_max = max(4, int(-(-(-(-int(per_tile.max()) // ng)) // 4) * 4))
    assert max_load <= int(np.ceil(1.2 * per_tile.max() / ng)) + c_max + 8
