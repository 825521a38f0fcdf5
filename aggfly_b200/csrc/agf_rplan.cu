// agf_rplan.cu -- host side of K1R (agf_regional.cuh): lowers a CSR weights matrix onto the cell tiles of a
// lat x lon grid, and the C-ABI entry points agf_rplan_* / agf_temporal_regional_*.
//
// Reference semantics: _weight_triplets (aggfly/aggregate/spatial.py:157-178) fixes the order of a region's entries
// (weights-frame order); the tables below keep that order inside every (tile, region) slot, so the in-kernel sums add
// the same terms in the same order as np.add.at for every region that lies inside one tile.
#include <algorithm>
#include <mutex>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda.h>

#include "agf_host.h"
#include "agf_regional.cuh"

using namespace agf;

template <typename V>
static int upload_vec(V **dst, const std::vector<V> &src, int64_t *bytes) {
    *dst = nullptr;
    const size_t n = std::max<size_t>(src.size(), 1);
    CU(cudaMalloc((void **)dst, n * sizeof(V)));
    if (!src.empty()) CU(cudaMemcpy(*dst, src.data(), src.size() * sizeof(V), cudaMemcpyHostToDevice));
    *bytes += (int64_t)(n * sizeof(V));
    return 0;
}

extern "C" int agf_rplan_destroy(agf_rplan_t *p) {
    if (!p) return 0;
    cudaFree(p->d_tile_ids);
    cudaFree(p->d_tile_slot_ptr);
    cudaFree(p->d_slot_dst);
    cudaFree(p->d_multi_regions);
    cudaFree(p->d_slot_ent_ptr);
    cudaFree(p->d_region_slot_ptr);
    cudaFree(p->d_region_slots);
    cudaFree(p->d_entries);
    for (auto &sg : p->seg) {
        cudaFree(sg.d_tile_seg_ptr);
        cudaFree(sg.d_tile_pent_ptr);
        cudaFree(sg.d_grp);
        cudaFree(sg.d_seg);
        cudaFree(sg.d_slot_q);
        cudaFree(sg.d_pent);
    }
    delete p;
    return 0;
}

namespace {
struct HostTables {
    std::vector<int32_t> tile_ids, tile_slot_ptr, slot_region, slot_ent_ptr, region_slot_ptr, region_slots;
    std::vector<int32_t> slot_dst, multi_regions;
    std::vector<RgEntry> entries;
    int tiles_x = 0, tiles_y = 0, max_slots = 0, n_empty = 0, n_partial_rows = 0;
};
}  // namespace

static int build_tables(HostTables &t, int32_t n_regions, int32_t n_lat, int32_t n_lon, int64_t nnz,
                        const int32_t *row_ptr, const int32_t *cell_idx, const double *w) {
    if (n_regions <= 0 || n_lat <= 0 || n_lon <= 0 || nnz < 0 || nnz > 0x7fffffff)
        return agf_fail(AGF_E_INVALID, "bad sizes");
    if (!row_ptr || (nnz > 0 && (!cell_idx || !w))) return agf_fail(AGF_E_INVALID, "null CSR array");
    if (row_ptr[0] != 0 || row_ptr[n_regions] != nnz) return agf_fail(AGF_E_INVALID, "row_ptr does not span nnz");
    const int64_t n_cells = (int64_t)n_lat * n_lon;
    const int tiles_x = (n_lon + RG_TW - 1) / RG_TW, tiles_y = (n_lat + RG_TH - 1) / RG_TH;
    const int64_t n_tiles = (int64_t)tiles_x * tiles_y;
    t.tiles_x = tiles_x;
    t.tiles_y = tiles_y;

    // counting sort of the entries by tile; visiting regions in ascending order and their entries in CSR order keeps
    // (region ascending, weights-frame order) inside every tile
    std::vector<int32_t> tile_of(nnz);
    std::vector<int64_t> tile_cnt(n_tiles + 1, 0);
    for (int64_t e = 0; e < nnz; ++e) {
        const int64_t c = cell_idx[e];
        if (c < 0 || c >= n_cells) return agf_fail(AGF_E_INVALID, "cell_idx[%lld] = %lld outside the grid", (long long)e, (long long)c);
        const int lat = (int)(c / n_lon), lon = (int)(c % n_lon);
        tile_of[e] = (lat / RG_TH) * tiles_x + lon / RG_TW;
        ++tile_cnt[tile_of[e] + 1];
    }
    for (int64_t k = 0; k < n_tiles; ++k) tile_cnt[k + 1] += tile_cnt[k];
    std::vector<int64_t> pos(tile_cnt.begin(), tile_cnt.end() - 1);
    std::vector<int32_t> ent_region(nnz);
    std::vector<RgEntry> ent(nnz);
    for (int32_t r = 0; r < n_regions; ++r) {
        if (row_ptr[r] > row_ptr[r + 1]) return agf_fail(AGF_E_INVALID, "row_ptr not monotonic");
        for (int64_t e = row_ptr[r]; e < row_ptr[r + 1]; ++e) {
            const int64_t c = cell_idx[e];
            const int lat = (int)(c / n_lon), lon = (int)(c % n_lon);
            const int64_t k = pos[tile_of[e]]++;
            ent_region[k] = r;
            ent[k].w = w[e];
            ent[k].cell = (lat % RG_TH) * RG_TW + (lon % RG_TW);
            ent[k].pad = 0;
        }
    }
    // slots: runs of one region inside a tile, re-ordered LONGEST FIRST inside the tile (the kernel hands neighbouring
    // slots to the lane groups of one warp: rows of similar length keep the warp's lanes busy together)
    t.tile_slot_ptr.assign(1, 0);
    t.entries.reserve(nnz);
    struct Run { int64_t k0, k1; int32_t region; };
    std::vector<Run> runs;
    for (int64_t tl = 0; tl < n_tiles; ++tl) {
        if (tile_cnt[tl + 1] == tile_cnt[tl]) continue;
        runs.clear();
        for (int64_t k = tile_cnt[tl]; k < tile_cnt[tl + 1]; ++k) {
            if (k == tile_cnt[tl] || ent_region[k] != ent_region[k - 1]) runs.push_back(Run{k, k, ent_region[k]});
            runs.back().k1 = k + 1;
        }
        std::stable_sort(runs.begin(), runs.end(), [](const Run &a, const Run &b) { return a.k1 - a.k0 > b.k1 - b.k0; });
        for (const Run &r : runs) {
            t.slot_region.push_back(r.region);
            t.slot_ent_ptr.push_back((int32_t)t.entries.size());
            t.entries.insert(t.entries.end(), ent.begin() + r.k0, ent.begin() + r.k1);
        }
        t.tile_ids.push_back((int32_t)tl);
        t.tile_slot_ptr.push_back((int32_t)t.slot_region.size());
        t.max_slots = std::max(t.max_slots, (int)runs.size());
    }
    t.slot_ent_ptr.push_back((int32_t)nnz);
    const int n_gslots = (int)t.slot_region.size();
    // slots of every region in ascending slot order (tiles in raster order)
    t.region_slot_ptr.assign(n_regions + 1, 0);
    t.region_slots.resize(n_gslots);
    for (int s = 0; s < n_gslots; ++s) ++t.region_slot_ptr[t.slot_region[s] + 1];
    for (int r = 0; r < n_regions; ++r) t.region_slot_ptr[r + 1] += t.region_slot_ptr[r];
    {
        std::vector<int32_t> at(t.region_slot_ptr.begin(), t.region_slot_ptr.end() - 1);
        for (int s = 0; s < n_gslots; ++s) t.region_slots[at[t.slot_region[s]]++] = s;
    }
    // where a slot's sums go: the panel row of its region when the slot holds the whole region, else a partial row
    t.slot_dst.resize(n_gslots);
    for (int s = 0; s < n_gslots; ++s) {
        const int r = t.slot_region[s];
        t.slot_dst[s] = (t.region_slot_ptr[r + 1] - t.region_slot_ptr[r] == 1) ? r : -(++t.n_partial_rows);
    }
    for (int r = 0; r < n_regions; ++r)
        if (t.region_slot_ptr[r + 1] - t.region_slot_ptr[r] > 1) t.multi_regions.push_back(r);
    for (int r = 0; r < n_regions; ++r) t.n_empty += t.region_slot_ptr[r + 1] == t.region_slot_ptr[r];
    return 0;
}

static void fill_info(agf_rplan_info_t *info, const HostTables &t, int64_t table_bytes) {
    memset(info, 0, sizeof(*info));
    info->n_tiles = t.tiles_x * t.tiles_y;
    info->n_active_tiles = (int32_t)t.tile_ids.size();
    info->n_slots = (int32_t)t.slot_region.size();
    info->max_slots_per_tile = t.max_slots;
    info->n_entries = (int64_t)t.entries.size();
    info->tile_lat = RG_TH;
    info->tile_lon = RG_TW;
    info->n_empty_regions = t.n_empty;
    info->n_partial_rows = t.n_partial_rows;
    info->table_bytes = table_bytes;
}

// host-only: the tables agf_rplan_create would upload (no device needed; used by the CPU tests of the lowering)
extern "C" int agf_rplan_tables(int32_t n_regions, int32_t n_lat, int32_t n_lon, int64_t nnz, const int32_t *row_ptr,
                                const int32_t *cell_idx, const double *w, agf_rplan_info_t *info, int32_t *tile_ids,
                                int32_t *tile_slot_ptr, int32_t *slot_region, int32_t *slot_ent_ptr, int32_t *entry_cell,
                                double *entry_w, int32_t *region_slot_ptr, int32_t *region_slots, int32_t *slot_dst) {
    HostTables t;
    int rc = build_tables(t, n_regions, n_lat, n_lon, nnz, row_ptr, cell_idx, w);
    if (rc) return rc;
    if (info) fill_info(info, t, 0);
    auto put = [](int32_t *dst, const std::vector<int32_t> &src) {
        if (dst && !src.empty()) memcpy(dst, src.data(), src.size() * sizeof(int32_t));
    };
    put(tile_ids, t.tile_ids);
    put(tile_slot_ptr, t.tile_slot_ptr);
    put(slot_region, t.slot_region);
    put(slot_ent_ptr, t.slot_ent_ptr);
    put(region_slot_ptr, t.region_slot_ptr);
    put(region_slots, t.region_slots);
    put(slot_dst, t.slot_dst);
    for (size_t k = 0; k < t.entries.size(); ++k) {
        if (entry_cell) entry_cell[k] = t.entries[k].cell;
        if (entry_w) entry_w[k] = t.entries[k].w;
    }
    return 0;
}

extern "C" int agf_rplan_create(agf_rplan_t **out, int32_t n_regions, int32_t n_lat, int32_t n_lon, int64_t nnz,
                                const int32_t *row_ptr, const int32_t *cell_idx, const double *w) {
    if (!out) return agf_fail(AGF_E_INVALID, "null out");
    *out = nullptr;
    HostTables t;
    int rc = build_tables(t, n_regions, n_lat, n_lon, nnz, row_ptr, cell_idx, w);
    if (rc) return rc;
    agf_rplan *p = new agf_rplan();
    p->n_regions = n_regions;
    p->n_lat = n_lat;
    p->n_lon = n_lon;
    p->tiles_x = t.tiles_x;
    p->tiles_y = t.tiles_y;
    p->n_active = (int)t.tile_ids.size();
    p->n_gslots = (int)t.slot_region.size();
    p->max_slots = t.max_slots;
    p->n_partial_rows = t.n_partial_rows;
    p->n_multi = (int)t.multi_regions.size();
    p->n_entries = nnz;
    p->n_empty_regions = t.n_empty;
    cudaError_t e0 = cudaGetDevice(&p->device);
    if (e0 != cudaSuccess) {
        delete p;
        return agf_cuda_fail(e0, "cudaGetDevice");
    }
    if ((rc = upload_vec(&p->d_tile_ids, t.tile_ids, &p->table_bytes)) ||
        (rc = upload_vec(&p->d_tile_slot_ptr, t.tile_slot_ptr, &p->table_bytes)) ||
        (rc = upload_vec(&p->d_slot_dst, t.slot_dst, &p->table_bytes)) ||
        (rc = upload_vec(&p->d_multi_regions, t.multi_regions, &p->table_bytes)) ||
        (rc = upload_vec(&p->d_slot_ent_ptr, t.slot_ent_ptr, &p->table_bytes)) ||
        (rc = upload_vec(&p->d_region_slot_ptr, t.region_slot_ptr, &p->table_bytes)) ||
        (rc = upload_vec(&p->d_region_slots, t.region_slots, &p->table_bytes)) ||
        (rc = upload_vec((RgEntry **)&p->d_entries, t.entries, &p->table_bytes))) {
        agf_rplan_destroy(p);
        return rc;
    }
    p->h_tile_slot_ptr = t.tile_slot_ptr;
    p->h_slot_ent_ptr = t.slot_ent_ptr;
    p->h_ent_w.resize(t.entries.size());
    p->h_ent_cell.resize(t.entries.size());
    for (size_t k = 0; k < t.entries.size(); ++k) {
        p->h_ent_w[k] = t.entries[k].w;
        p->h_ent_cell[k] = t.entries[k].cell;
    }
    *out = p;
    return 0;
}

// ------------------------------------------------------------------------------------------
// Balanced walk: the entries of a tile dealt to its lane groups
// ------------------------------------------------------------------------------------------
// The kernel walks a tile's entries once per period with 256 / LPS lane groups.  Handing whole slots to the groups left
// most of them idle behind the longest slot (ncu r2m: 15 % of all stall samples on the per-period barrier).  Here every
// slot is cut into SEGMENTS of at most c entries, c ~ entries / groups, the segments are dealt to the groups longest
// first onto the least loaded group, and each group's segments are laid out back to back (each padded to a multiple of
// four entries with zero-weight pads, so the kernel's loop has no remainder).  Segment sums are added in ascending
// segment order afterwards: the association is fixed by these tables, results do not depend on scheduling.
namespace {
struct SegHost {
    std::vector<int32_t> tile_seg_ptr, tile_pent_ptr, grp, seg, slot_q;
    std::vector<RgEntry> pent;
    int max_segs = 0, max_pent = 0;
};

void build_segments(const agf_rplan *p, int ng, SegHost &o) {
    const int n_active = p->n_active;
    o.tile_seg_ptr.assign(1, 0);
    o.tile_pent_ptr.assign(1, 0);
    o.grp.reserve((size_t)n_active * (ng + 1) * 2);
    o.slot_q.resize((size_t)p->n_gslots * 2);
    struct Seg { int e0, n, sid; };
    std::vector<Seg> segs;
    std::vector<int> order, load;
    std::vector<std::vector<int>> mine(ng);
    for (int ti = 0; ti < n_active; ++ti) {
        const int s0 = p->h_tile_slot_ptr[ti], s1 = p->h_tile_slot_ptr[ti + 1];
        const int total = p->h_slot_ent_ptr[s1] - p->h_slot_ent_ptr[s0];
        const int c = std::max(4, ((total + ng - 1) / ng + 3) / 4 * 4);
        segs.clear();
        for (int s = s0; s < s1; ++s) {
            const int e0 = p->h_slot_ent_ptr[s], n = p->h_slot_ent_ptr[s + 1] - e0;
            const int k = std::max(1, (n + c - 1) / c);
            o.slot_q[2 * (size_t)s] = (int)segs.size();
            for (int i = 0; i < k; ++i) {   // even split: pieces differ by at most one entry
                const int a = (int)((int64_t)n * i / k), b = (int)((int64_t)n * (i + 1) / k);
                segs.push_back(Seg{e0 + a, b - a, (int)segs.size()});
            }
            o.slot_q[2 * (size_t)s + 1] = (int)segs.size();
        }
        order.resize(segs.size());
        for (size_t i = 0; i < segs.size(); ++i) order[i] = (int)i;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return segs[a].n > segs[b].n; });
        load.assign(ng, 0);
        for (auto &m : mine) m.clear();
        for (int i : order) {
            int g = 0;
            for (int k = 1; k < ng; ++k)
                if (load[k] < load[g]) g = k;
            load[g] += (segs[i].n + 3) / 4 * 4;
            mine[g].push_back(i);
        }
        const int seg_base = o.tile_seg_ptr.back(), pent_base = o.tile_pent_ptr.back();
        int n_seg = 0, n_pent = 0;
        for (int g = 0; g < ng; ++g) {
            o.grp.push_back(n_seg);
            o.grp.push_back(n_pent);
            std::sort(mine[g].begin(), mine[g].end());
            for (int i : mine[g]) {
                const Seg &sg = segs[i];
                for (int e = 0; e < sg.n; ++e) {
                    RgEntry en;
                    en.w = p->h_ent_w[sg.e0 + e];
                    en.cell = p->h_ent_cell[sg.e0 + e];
                    en.pad = 0;
                    o.pent.push_back(en);
                }
                for (int e = sg.n; e % 4 != 0; ++e) o.pent.push_back(RgEntry{0.0, -1, 0});
                n_pent += (sg.n + 3) / 4 * 4;
                o.seg.push_back(n_pent);
                o.seg.push_back(sg.sid);
                ++n_seg;
            }
        }
        o.grp.push_back(n_seg);
        o.grp.push_back(n_pent);
        o.tile_seg_ptr.push_back(seg_base + n_seg);
        o.tile_pent_ptr.push_back(pent_base + n_pent);
        o.max_segs = std::max(o.max_segs, n_seg);
        o.max_pent = std::max(o.max_pent, n_pent);
    }
}
}  // namespace

const agf_rplan::SegTables *agf_rplan_segments(const agf_rplan *p, int lps) {
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    const int idx = lps == 4 ? 0 : (lps == 8 ? 1 : (lps == 16 ? 2 : -1));
    if (idx < 0) {
        agf_fail(AGF_E_INVALID, "lanes per slot %d", lps);
        return nullptr;
    }
    agf_rplan::SegTables &sg = p->seg[idx];
    if (sg.ng != 0) return &sg;
    if ((int64_t)p->n_entries * 2 > 0x7fffffff) {
        agf_fail(AGF_E_UNSUPPORTED, "too many entries for the balanced walk tables");
        return nullptr;
    }
    SegHost h;
    build_segments(p, 256 / lps, h);
    int64_t bytes = 0;
    auto up = [&]() -> int {
        int rc;
        if ((rc = upload_vec(&sg.d_tile_seg_ptr, h.tile_seg_ptr, &bytes)) || (rc = upload_vec(&sg.d_tile_pent_ptr, h.tile_pent_ptr, &bytes)) ||
            (rc = upload_vec(&sg.d_grp, h.grp, &bytes)) || (rc = upload_vec(&sg.d_seg, h.seg, &bytes)) ||
            (rc = upload_vec(&sg.d_slot_q, h.slot_q, &bytes)) || (rc = upload_vec((RgEntry **)&sg.d_pent, h.pent, &bytes)))
            return rc;
        return 0;
    };
    if (up()) return nullptr;
    sg.max_segs = h.max_segs;
    sg.max_pent = h.max_pent;
    sg.n_segs = (int64_t)h.seg.size() / 2;
    sg.n_pent = (int64_t)h.pent.size();
    sg.ng = 256 / lps;
    return &sg;
}

// host-only: build the balanced walk tables for `lps` lanes per slot and check them against the slot tables --
// walking every group's segments must visit every entry of every slot exactly once, in the slot's own order when the
// segments are taken by ascending id.  stats: segments, padded entries, largest / mean group load (entries), tiles.
extern "C" int agf_rplan_check_segments(int32_t n_regions, int32_t n_lat, int32_t n_lon, int64_t nnz, const int32_t *row_ptr,
                                        const int32_t *cell_idx, const double *w, int32_t lps, int64_t *stats) {
    if (lps != 4 && lps != 8 && lps != 16) return agf_fail(AGF_E_INVALID, "lanes per slot %d", lps);
    HostTables t;
    int rc = build_tables(t, n_regions, n_lat, n_lon, nnz, row_ptr, cell_idx, w);
    if (rc) return rc;
    agf_rplan plan;
    plan.n_active = (int)t.tile_ids.size();
    plan.n_gslots = (int)t.slot_region.size();
    plan.h_tile_slot_ptr = t.tile_slot_ptr;
    plan.h_slot_ent_ptr = t.slot_ent_ptr;
    for (const RgEntry &e : t.entries) {
        plan.h_ent_w.push_back(e.w);
        plan.h_ent_cell.push_back(e.cell);
    }
    const int ng = 256 / lps;
    SegHost h;
    build_segments(&plan, ng, h);
    int64_t max_load = 0, sum_load = 0;
    for (int ti = 0; ti < plan.n_active; ++ti) {
        const int s0 = t.tile_slot_ptr[ti], s1 = t.tile_slot_ptr[ti + 1];
        const int sb = h.tile_seg_ptr[ti], nseg = h.tile_seg_ptr[ti + 1] - sb;
        const int pb = h.tile_pent_ptr[ti], npent = h.tile_pent_ptr[ti + 1] - pb;
        const int32_t *grp = &h.grp[(size_t)ti * (ng + 1) * 2];
        if (grp[0] != 0 || grp[1] != 0 || grp[2 * ng] != nseg || grp[2 * ng + 1] != npent)
            return agf_fail(AGF_E_STATE, "tile %d: group table does not span the tile", ti);
        // segment id -> (first padded entry, end)
        std::vector<int> first(nseg, -1), last(nseg, -1);
        for (int g = 0; g < ng; ++g) {
            int e = grp[2 * g + 1];
            if (grp[2 * g] > grp[2 * g + 2] || e > grp[2 * g + 3]) return agf_fail(AGF_E_STATE, "tile %d: group table not monotonic", ti);
            for (int j = grp[2 * g]; j < grp[2 * g + 2]; ++j) {
                const int end = h.seg[2 * (size_t)(sb + j)], sid = h.seg[2 * (size_t)(sb + j) + 1];
                if (sid < 0 || sid >= nseg || first[sid] != -1 || end <= e || (end - e) % 4 != 0)
                    return agf_fail(AGF_E_STATE, "tile %d: bad segment record", ti);
                first[sid] = e;
                last[sid] = end;
                e = end;
            }
            if (e != grp[2 * g + 3]) return agf_fail(AGF_E_STATE, "tile %d: group %d entries do not end at the next group", ti, g);
            max_load = std::max<int64_t>(max_load, grp[2 * g + 3] - grp[2 * g + 1]);
        }
        sum_load += npent;
        for (int s = s0; s < s1; ++s) {
            int e = t.slot_ent_ptr[s];
            for (int sid = h.slot_q[2 * (size_t)s]; sid < h.slot_q[2 * (size_t)s + 1]; ++sid) {
                if (sid < 0 || sid >= nseg || first[sid] < 0) return agf_fail(AGF_E_STATE, "tile %d: slot segment missing", ti);
                bool pad = false;
                for (int k = first[sid]; k < last[sid]; ++k) {
                    const RgEntry &pe = h.pent[(size_t)pb + k];
                    if (pe.cell < 0) {
                        if (pe.w != 0.0) return agf_fail(AGF_E_STATE, "pad with a weight");
                        pad = true;
                        continue;
                    }
                    if (pad || e >= t.slot_ent_ptr[s + 1] || pe.cell != t.entries[e].cell || pe.w != t.entries[e].w)
                        return agf_fail(AGF_E_STATE, "tile %d slot %d: entries out of order", ti, s);
                    ++e;
                }
            }
            if (e != t.slot_ent_ptr[s + 1]) return agf_fail(AGF_E_STATE, "tile %d slot %d: entries missing", ti, s);
        }
    }
    if (stats) {
        stats[0] = (int64_t)h.seg.size() / 2;
        stats[1] = (int64_t)h.pent.size();
        stats[2] = max_load;
        stats[3] = plan.n_active ? sum_load / ((int64_t)plan.n_active * ng) : 0;
        stats[4] = plan.n_active;
        stats[5] = h.max_segs;
        stats[6] = h.max_pent;
    }
    return 0;
}

extern "C" int agf_rplan_info(const agf_rplan_t *p, agf_rplan_info_t *info) {
    if (!p || !info) return agf_fail(AGF_E_INVALID, "null argument");
    memset(info, 0, sizeof(*info));
    info->n_tiles = p->tiles_x * p->tiles_y;
    info->n_active_tiles = p->n_active;
    info->n_slots = p->n_gslots;
    info->max_slots_per_tile = p->max_slots;
    info->n_entries = p->n_entries;
    info->tile_lat = RG_TH;
    info->tile_lon = RG_TW;
    info->n_empty_regions = p->n_empty_regions;
    info->n_partial_rows = p->n_partial_rows;
    info->table_bytes = p->table_bytes;
    return 0;
}

// ------------------------------------------------------------------------------------------
// 3-D tensor map over the raster [rows, n_lat, n_lon]
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn3)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int agf_make_tensor_map3(agf::TensorMap *out, const void *base, int elem_size, uint64_t n_lon, uint64_t n_lat,
                         uint64_t n_rows, uint64_t ld, int box_lon, int box_lat, int box_rows) {
    static EncodeTiledFn3 fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn3)ptr;
    }
    if (!fn) return agf_fail(AGF_E_UNSUPPORTED, "cuTensorMapEncodeTiled not available");
    cuuint64_t dims[3] = {n_lon, n_lat, n_rows};
    cuuint64_t strides[2] = {n_lon * (cuuint64_t)elem_size, ld * (cuuint64_t)elem_size};
    cuuint32_t box[3] = {(cuuint32_t)box_lon, (cuuint32_t)box_lat, (cuuint32_t)box_rows};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn((CUtensorMap *)out, elem_size == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                    3, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    // (L2 promotion 128 B = one box row: with 256 B the neighbouring tile's half line was fetched along, 32.2 GB read
    // from DRAM for 27.9 GB of tiles, ncu r2d)
    if (r != CUDA_SUCCESS) return agf_fail(AGF_E_UNSUPPORTED, "cuTensorMapEncodeTiled (3-D) failed (%d)", (int)r);
    return 0;
}

// ------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------
static int regional_args(const agf_program_t *p, const agf_rplan_t *plan, int64_t g0, int64_t g1) {
    if (!p || !plan) return agf_fail(AGF_E_INVALID, "null argument");
    int dev = -1;
    CU(cudaGetDevice(&dev));
    if (dev != p->device || dev != plan->device)
        return agf_fail(AGF_E_STATE, "handles belong to devices %d / %d, current device is %d", p->device, plan->device, dev);
    if ((int64_t)plan->n_lat * plan->n_lon != p->n_cells)
        return agf_fail(AGF_E_INVALID, "plan grid %d x %d does not match the program's %lld cells", plan->n_lat, plan->n_lon,
                        (long long)p->n_cells);
    if (g0 < 0 || g1 > p->desc.n_groups1 || g0 >= g1) return agf_fail(AGF_E_INVALID, "bad period range [%lld, %lld)", (long long)g0, (long long)g1);
    return 0;
}

extern "C" int agf_temporal_regional_plan(const agf_program_t *p, const agf_rplan_t *plan, int64_t panel_groups,
                                          agf_regional_info_t *info) {
    if (!info) return agf_fail(AGF_E_INVALID, "null info");
    memset(info, 0, sizeof(*info));
    int rc = regional_args(p, plan, 0, p ? p->desc.n_groups1 : 0);
    if (rc) return rc;
    if (panel_groups < p->desc.n_groups1) return agf_fail(AGF_E_INVALID, "panel_groups < the program's periods");
    RegionalLaunch a{};
    a.k.p = p;
    a.k.use_tma = 1;
    a.plan = plan;
    a.g_begin = 0;
    a.g_end = p->desc.n_groups1;
    a.G = panel_groups;
    RegionalChoice ch{};
    int krc = 0;
    if (p->desc.in_dtype != AGF_F32 || agf_k1_f32_regional(a, 1, &ch, &krc)) {
        agf_fail(AGF_E_UNSUPPORTED, "no regional instantiation for this program (single-level float32 programs with 24-row "
                                    "periods are covered); run agf_temporal_run + agf_spmm_run");
        return 0;  // info->supported stays 0
    }
    if (krc) return krc;
    info->supported = 1;
    info->lanes_per_slot = ch.lps;
    info->kernel_lanes = ch.lanes;
    info->smem_bytes = ch.smem_bytes;
    info->ctas_per_sm = ch.ctas_per_sm;
    info->workspace_bytes = (int64_t)plan->n_partial_rows * panel_groups * ch.lps * 16 + 256;
    return 0;
}

extern "C" int agf_temporal_regional_run(const agf_program_t *p, const agf_rplan_t *plan, const void *d_x, int64_t ld,
                                         int64_t row0, int64_t group_begin, int64_t group_end, void *d_workspace,
                                         int64_t workspace_bytes, double *d_panel, int64_t panel_groups,
                                         int32_t out_ncols, double *d_den, uintptr_t stream) {
    int rc = regional_args(p, plan, group_begin, group_end);
    if (rc) return rc;
    if (!d_x || !d_panel || (!d_workspace && plan->n_partial_rows > 0)) return agf_fail(AGF_E_INVALID, "null buffer");
    if (ld < p->n_cells) return agf_fail(AGF_E_INVALID, "ld < n_cells");
    if (row0 < 0 || row0 > p->b1[group_begin]) return agf_fail(AGF_E_INVALID, "row0 is past the first row of the period range");
    if (panel_groups < group_end) return agf_fail(AGF_E_INVALID, "panel has %lld periods, the range ends at %lld", (long long)panel_groups, (long long)group_end);
    for (int c = 0; c < p->desc.n_cols; ++c)
        if (p->desc.cols[c].dst >= out_ncols) return agf_fail(AGF_E_INVALID, "col %d: dst outside the panel (out_ncols=%d)", c, out_ncols);
    if (p->desc.in_dtype != AGF_F32) return agf_fail(AGF_E_UNSUPPORTED, "regional kernel: float32 rasters only");
    if (((uintptr_t)d_x % 16) != 0 || (ld % 4) != 0 || (plan->n_lon % 4) != 0)
        return agf_fail(AGF_E_UNSUPPORTED, "regional kernel needs a 16-byte aligned raster whose row and latitude pitches are "
                                           "multiples of 16 bytes (n_lon %% 4 == 0)");
    RegionalLaunch a{};
    a.k.p = p;
    a.k.d_x = d_x;
    a.k.ld = ld;
    a.k.row0 = row0;
    a.k.ncols = out_ncols;
    a.k.stream = (cudaStream_t)stream;
    a.k.use_tma = 1;
    a.plan = plan;
    a.g_begin = group_begin;
    a.g_end = group_end;
    a.d_workspace = d_workspace;
    a.workspace_bytes = workspace_bytes;
    a.d_panel = d_panel;
    a.d_den = d_den;
    a.G = panel_groups;
    a.out_ncols = out_ncols;
    int krc = 0;
    if (agf_k1_f32_regional(a, 0, nullptr, &krc))
        return agf_fail(AGF_E_UNSUPPORTED, "no regional instantiation for this program");
    return krc;
}
