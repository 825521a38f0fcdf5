// agf_geom.cu -- host-only geometry of the weights builder: exact area of (region polygon INTERSECT
// grid-cell rectangle) for every cell a region touches.
//
// Replaces the GEOS work of aggfly/weights/grid_weights.py:238-421 (two buffered centroid sjoins to
// find interior / border cells, shapely intersection of the border cells, area / cell_area).  The
// reference's result is "1 for cells entirely inside, area fraction for cells on the border, cells
// with zero overlap dropped"; clipping every candidate cell of the region's bounding box gives the
// same numbers directly (a cell whose clipped area equals the cell area is interior), without a
// geometry library: Sutherland-Hodgman against an axis-aligned window is exact for any simple ring,
// and rings with opposite orientation (holes) subtract through the signed shoelace sum.
//
// No CUDA here: this runs once per (grid, regions) pair and is cached; it is not on the HBM-bound
// path.  It lives in the library so that weights_from_objects(...).calculate_weights() works without
// geopandas/shapely.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "agf_host.h"

namespace {

struct Pt {
    double x, y;
};

// keep the part of ring `in` with  s * (coord - bound) >= 0  (axis 0: x, 1: y)
static void clip_half(const std::vector<Pt> &in, std::vector<Pt> &out, int axis, double bound, double s) {
    out.clear();
    const size_t n = in.size();
    if (n == 0) return;
    auto inside = [&](const Pt &p) { return s * ((axis == 0 ? p.x : p.y) - bound) >= 0.0; };
    auto cut = [&](const Pt &a, const Pt &b) {
        Pt r;
        if (axis == 0) {
            r.x = bound;
            r.y = a.y + (b.y - a.y) * (bound - a.x) / (b.x - a.x);
        } else {
            r.y = bound;
            r.x = a.x + (b.x - a.x) * (bound - a.y) / (b.y - a.y);
        }
        return r;
    };
    Pt a = in[n - 1];
    bool ia = inside(a);
    for (size_t i = 0; i < n; ++i) {
        const Pt b = in[i];
        const bool ib = inside(b);
        if (ia != ib) out.push_back(cut(a, b));
        if (ib) out.push_back(b);
        a = b;
        ia = ib;
    }
}

static double signed_area(const std::vector<Pt> &p) {
    const size_t n = p.size();
    if (n < 3) return 0.0;
    double s = 0.0;
    // shoelace relative to the first vertex: keeps the products small for far-from-origin cells
    const double ox = p[0].x, oy = p[0].y;
    for (size_t i = 1; i + 1 < n; ++i)
        s += (p[i].x - ox) * (p[i + 1].y - oy) - (p[i + 1].x - ox) * (p[i].y - oy);
    return 0.5 * s;
}

struct Axis {  // cell centres of one grid axis, any monotonic order, uniform spacing d
    std::vector<double> lo;  // lower edge of each cell, sorted ascending
    std::vector<int> idx;    // original index of the sorted cell
    double d;
    void build(const double *c, int n, double d_) {
        d = d_;
        idx.resize(n);
        for (int i = 0; i < n; ++i) idx[i] = i;
        std::sort(idx.begin(), idx.end(), [&](int a, int b) { return c[a] < c[b]; });
        lo.resize(n);
        for (int i = 0; i < n; ++i) lo[i] = c[idx[i]] - d / 2;
    }
    // sorted positions [a, b) of the cells whose interval overlaps (vmin, vmax)
    void range(double vmin, double vmax, int *a, int *b) const {
        *a = (int)(std::upper_bound(lo.begin(), lo.end(), vmin - d) - lo.begin());  // lo + d > vmin
        *b = (int)(std::lower_bound(lo.begin(), lo.end(), vmax) - lo.begin());      // lo < vmax
        if (*a > *b) *a = *b;
    }
};

}  // namespace

struct agf_overlap {
    std::vector<int32_t> region;
    std::vector<int64_t> cell;
    std::vector<double> frac;
};

extern "C" int agf_overlap_create(agf_overlap_t **out, int32_t n_regions, const int64_t *region_ring_ptr,
                                  const int64_t *ring_ptr, const double *xy, int32_t n_lon, const double *lon,
                                  double dlon, int32_t n_lat, const double *lat, double dlat, int64_t *n_pairs) {
    if (!out || !region_ring_ptr || !ring_ptr || !xy || !lon || !lat)
        return agf_fail(AGF_E_INVALID, "agf_overlap_create: null argument");
    if (n_regions < 0 || n_lon <= 0 || n_lat <= 0 || !(dlon > 0) || !(dlat > 0))
        return agf_fail(AGF_E_INVALID, "agf_overlap_create: bad sizes");
    *out = nullptr;
    Axis ax, ay;
    ax.build(lon, n_lon, dlon);
    ay.build(lat, n_lat, dlat);
    const double cell_area = dlon * dlat;
    agf_overlap *h = new agf_overlap();
    std::vector<std::vector<Pt>> rings, band;
    std::vector<Pt> t1, t2;
    struct Hit {
        int64_t cell;
        double frac;
    };
    std::vector<Hit> hits;
    for (int32_t r = 0; r < n_regions; ++r) {
        const int64_t k0 = region_ring_ptr[r], k1 = region_ring_ptr[r + 1];
        if (k1 < k0) {
            delete h;
            return agf_fail(AGF_E_INVALID, "agf_overlap_create: region_ring_ptr not monotonic");
        }
        rings.assign((size_t)(k1 - k0), {});
        double xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
        for (int64_t k = k0; k < k1; ++k) {
            int64_t v0 = ring_ptr[k], v1 = ring_ptr[k + 1];
            if (v1 - v0 >= 2 && xy[2 * v0] == xy[2 * (v1 - 1)] && xy[2 * v0 + 1] == xy[2 * (v1 - 1) + 1])
                --v1;  // closed ring: drop the repeated first vertex
            auto &ring = rings[(size_t)(k - k0)];
            for (int64_t v = v0; v < v1; ++v) {
                const Pt p{xy[2 * v], xy[2 * v + 1]};
                if (!std::isfinite(p.x) || !std::isfinite(p.y)) {
                    delete h;
                    return agf_fail(AGF_E_INVALID, "agf_overlap_create: non-finite vertex in region %d", r);
                }
                ring.push_back(p);
                xmin = std::min(xmin, p.x);
                xmax = std::max(xmax, p.x);
                ymin = std::min(ymin, p.y);
                ymax = std::max(ymax, p.y);
            }
        }
        if (!(xmin < xmax) || !(ymin < ymax)) continue;  // empty / degenerate region: no cells
        int ya, yb;
        ay.range(ymin, ymax, &ya, &yb);
        hits.clear();
        for (int yi = ya; yi < yb; ++yi) {
            const double y0 = ay.lo[yi], y1 = y0 + dlat;
            // clip every ring to the latitude band once, then to each cell of the band
            band.assign(rings.size(), {});
            double bxmin = INFINITY, bxmax = -INFINITY;
            for (size_t k = 0; k < rings.size(); ++k) {
                clip_half(rings[k], t1, 1, y0, 1.0);
                clip_half(t1, band[k], 1, y1, -1.0);
                for (const Pt &p : band[k]) {
                    bxmin = std::min(bxmin, p.x);
                    bxmax = std::max(bxmax, p.x);
                }
            }
            if (!(bxmin < bxmax)) continue;
            int xa, xb;
            ax.range(bxmin, bxmax, &xa, &xb);
            for (int xi = xa; xi < xb; ++xi) {
                const double x0 = ax.lo[xi], x1 = x0 + dlon;
                double area = 0.0;
                for (size_t k = 0; k < band.size(); ++k) {
                    if (band[k].size() < 3) continue;
                    clip_half(band[k], t1, 0, x0, 1.0);
                    clip_half(t1, t2, 0, x1, -1.0);
                    area += signed_area(t2);
                }
                double f = std::fabs(area) / cell_area;
                if (!(f > 0.0)) continue;                       // grid_weights.py:407 keeps area_weight > 0
                if (std::fabs(f - 1.0) < 1e-12) f = 1.0;        // interior cell (the reference assigns exactly 1)
                hits.push_back(Hit{(int64_t)ay.idx[yi] * n_lon + ax.idx[xi], f});
            }
        }
        std::sort(hits.begin(), hits.end(), [](const Hit &a, const Hit &b) { return a.cell < b.cell; });
        for (const Hit &e : hits) {
            h->region.push_back(r);
            h->cell.push_back(e.cell);
            h->frac.push_back(e.frac);
        }
    }
    if (n_pairs) *n_pairs = (int64_t)h->frac.size();
    *out = h;
    return 0;
}

extern "C" int agf_overlap_fetch(const agf_overlap_t *h, int32_t *region, int64_t *cell_id, double *fraction) {
    if (!h || !region || !cell_id || !fraction) return agf_fail(AGF_E_INVALID, "agf_overlap_fetch: null argument");
    std::copy(h->region.begin(), h->region.end(), region);
    std::copy(h->cell.begin(), h->cell.end(), cell_id);
    std::copy(h->frac.begin(), h->frac.end(), fraction);
    return 0;
}

extern "C" int agf_overlap_destroy(agf_overlap_t *h) {
    delete h;
    return 0;
}
