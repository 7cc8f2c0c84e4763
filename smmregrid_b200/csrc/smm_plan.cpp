// smm_plan.cpp -- CDO links -> CSR by destination row -> tile plan of the staged kernel.
//
// Restates, for the device layout, what the reference does when it wraps the CDO arrays in
// sparse.COO (smmregrid/weights.py:25-44): addresses made 0-based, column 0 of
// remap_matrix, duplicate coordinates summed.  The row-major-by-destination order keeps
// each row's links in ascending source order, i.e. the order in which the reference's
// matmul loop visits them.
#include "smm_plan.h"

#include <algorithm>
#include <numeric>

#include "../../include/smmregrid_b200.h"

namespace smm {

int build_csr(int64_t n_src, int64_t n_dst, int64_t nnz, const int32_t *src_address,
              const int32_t *dst_address, const double *remap_matrix, int32_t num_wgts,
              int32_t index_base, HostCsr &out, std::string &err)
{
    if (n_src <= 0 || n_dst <= 0 || nnz < 0 || num_wgts < 1) {
        err = "smm_create: n_src, n_dst must be > 0, nnz >= 0, num_wgts >= 1";
        return SMM_ERR_INVALID;
    }
    if (n_src > INT32_MAX || n_dst >= INT32_MAX || nnz >= INT32_MAX) {
        err = "smm_create: sizes beyond int32 addressing are not supported";
        return SMM_ERR_INVALID;
    }
    if (nnz > 0 && (!src_address || !dst_address || !remap_matrix)) {
        err = "smm_create: null link array";
        return SMM_ERR_INVALID;
    }
    out.n_src = n_src;
    out.n_dst = n_dst;

    std::vector<uint64_t> key(static_cast<size_t>(nnz));
    bool sorted = true;
    for (int64_t k = 0; k < nnz; ++k) {
        const int64_t s = static_cast<int64_t>(src_address[k]) - index_base;
        const int64_t d = static_cast<int64_t>(dst_address[k]) - index_base;
        if (s < 0 || s >= n_src || d < 0 || d >= n_dst) {
            err = "smm_create: link " + std::to_string(k) + " addresses (" +
                  std::to_string(src_address[k]) + "," + std::to_string(dst_address[k]) +
                  ") outside the grids";
            return SMM_ERR_RANGE;
        }
        key[k] = (static_cast<uint64_t>(d) << 32) | static_cast<uint64_t>(s);
        if (k > 0 && key[k] <= key[k - 1]) sorted = false;
    }
    std::vector<int64_t> order;
    if (!sorted) {
        order.resize(static_cast<size_t>(nnz));
        std::iota(order.begin(), order.end(), int64_t{0});
        std::stable_sort(order.begin(), order.end(),
                         [&](int64_t a, int64_t b) { return key[a] < key[b]; });
    }
    out.rowptr.assign(static_cast<size_t>(n_dst) + 1, 0);
    out.col.clear();
    out.val.clear();
    out.col.reserve(static_cast<size_t>(nnz));
    out.val.reserve(static_cast<size_t>(nnz));
    uint64_t prev = ~uint64_t{0};
    for (int64_t i = 0; i < nnz; ++i) {
        const int64_t k = sorted ? i : order[i];
        const double w = remap_matrix[k * static_cast<int64_t>(num_wgts)];
        if (i > 0 && key[k] == prev) {
            out.val.back() += w;   // sum_duplicates, original order
        } else {
            out.col.push_back(static_cast<int32_t>(key[k] & 0xffffffffu));
            out.val.push_back(w);
            out.rowptr[(key[k] >> 32) + 1] += 1;
            prev = key[k];
        }
    }
    out.max_row_nnz = 0;
    for (int64_t r = 0; r < n_dst; ++r) {
        out.max_row_nnz = std::max(out.max_row_nnz, out.rowptr[r + 1]);
        out.rowptr[r + 1] += out.rowptr[r];
    }
    std::vector<uint8_t> seen(static_cast<size_t>(n_src), 0);
    int64_t touched = 0;
    for (int32_t c : out.col)
        if (!seen[c]) { seen[c] = 1; ++touched; }
    out.touched_src = touched;
    return SMM_OK;
}

bool choose_lanes(int32_t m, int32_t &lpr, int32_t &kpl)
{
    if (m <= 4) { lpr = 1; kpl = 4; }
    else if (m <= 8) { lpr = 2; kpl = 4; }
    else if (m <= 16) { lpr = 1; kpl = 16; }
    else if (m <= 32) { lpr = 2; kpl = 16; }
    else if (m <= 64) { lpr = 4; kpl = 16; }
    else if (m <= 128) { lpr = 8; kpl = 16; }
    else if (m <= 256) { lpr = 16; kpl = 16; }
    else if (m <= 512) { lpr = 32; kpl = 16; }
    else return false;
    return true;
}

void build_plan(const HostCsr &csr, int32_t force_lpr, int32_t force_kpl, HostPlan &plan)
{
    plan = HostPlan{};
    int32_t lpr = force_lpr, kpl = force_kpl;
    if (lpr <= 0 || kpl <= 0) {
        if (!choose_lanes(csr.max_row_nnz, lpr, kpl)) {
            plan.why = "a destination row has more than 512 links";
            return;
        }
    } else if (csr.max_row_nnz > lpr * kpl) {
        plan.why = "row longer than the forced lane configuration";
        return;
    }
    const int64_t n_dst = csr.n_dst, n_src = csr.n_src;
    const int32_t R = kConsumerThreads / lpr;
    const int64_t ntiles = (n_dst + R - 1) / R;
    if (ntiles > INT32_MAX / 2) { plan.why = "too many tiles"; return; }
    plan.lpr = lpr; plan.kpl = kpl; plan.rows_per_tile = R;
    plan.tiles.resize(static_cast<size_t>(ntiles));
    plan.wplan.assign(static_cast<size_t>(ntiles) * kpl * kConsumerThreads, 0.0);
    plan.iplan.assign(static_cast<size_t>(ntiles) * kpl * kConsumerThreads, 0);

    std::vector<int32_t> cols;
    int64_t sum_blocks = 0;    // distinct aligned 8-element blocks touched, summed over tiles
    for (int64_t t = 0; t < ntiles; ++t) {
        const int64_t r0 = t * R, r1 = std::min(n_dst, r0 + R);
        const int32_t j0 = csr.rowptr[r0], j1 = csr.rowptr[r1];
        cols.assign(csr.col.begin() + j0, csr.col.begin() + j1);
        std::sort(cols.begin(), cols.end());
        cols.erase(std::unique(cols.begin(), cols.end()), cols.end());

        TileDesc &td = plan.tiles[t];
        td = TileDesc{};
        td.row0 = static_cast<int32_t>(r0);
        td.nrows = static_cast<int32_t>(r1 - r0);
        td.seg0 = static_cast<int32_t>(plan.segs.size());
        int64_t cs = -1, ce = -1, last_block = -1;
        uint32_t off = 0;
        auto flush = [&]() {
            if (cs < 0) return;
            Seg sg{static_cast<uint32_t>(cs), off, static_cast<uint32_t>(ce - cs), 0};
            plan.segs.push_back(sg);
            off += sg.len;
        };
        for (int32_t c : cols) {
            const int64_t blk = c / kSegAlign;
            if (blk != last_block) { ++sum_blocks; last_block = blk; }
            if (cs >= 0 && c < ce) continue;
            const int64_t s_al = blk * kSegAlign;
            const int64_t e_al = std::min<int64_t>((blk + 1) * kSegAlign, n_src);
            if (cs >= 0 && s_al - ce <= kSegGap) {
                ce = e_al;
            } else {
                flush();
                cs = s_al; ce = e_al;
            }
        }
        flush();
        td.nseg = static_cast<int32_t>(plan.segs.size()) - td.seg0;
        td.elems = static_cast<int32_t>(off);
        plan.max_tile_segments = std::max(plan.max_tile_segments, td.nseg);
        plan.max_tile_elems = std::max<int64_t>(plan.max_tile_elems, off);
        plan.sum_tile_elems += off;
        if (off > 65535u) { plan.why = "tile footprint exceeds 16-bit local indices"; return; }

        // register image: link e of a row -> lane (e % lpr) of the row's lane group, slot e / lpr
        const Seg *sb = plan.segs.data() + td.seg0;
        const size_t tbase = static_cast<size_t>(t) * kpl * kConsumerThreads;
        for (int64_t r = r0; r < r1; ++r) {
            const int32_t a = csr.rowptr[r], b = csr.rowptr[r + 1];
            for (int32_t j = a; j < b; ++j) {
                const uint32_t c = static_cast<uint32_t>(csr.col[j]);
                // last segment with src <= c
                int lo = 0, hi = td.nseg - 1;
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (sb[mid].src <= c) lo = mid; else hi = mid - 1;
                }
                const uint32_t li = sb[lo].dst + (c - sb[lo].src);
                const int32_t e = j - a;
                const int32_t thread = static_cast<int32_t>(r - r0) * lpr + (e % lpr);
                const int32_t slot = e / lpr;
                plan.wplan[tbase + static_cast<size_t>(slot) * kConsumerThreads + thread] = csr.val[j];
                plan.iplan[tbase + static_cast<size_t>(slot) * kConsumerThreads + thread] =
                    static_cast<uint16_t>(li);
            }
        }
    }
    // Staging pays when footprints are compact runs; otherwise the gather kernel is used.
    const double avg_seg = plan.segs.empty() ? 0.0
                                             : static_cast<double>(plan.sum_tile_elems) / plan.segs.size();
    if (plan.max_tile_elems > 25600) { plan.why = "tile footprint too large for two f32 stages"; return; }
    if (!plan.segs.empty() && avg_seg < 16.0) { plan.why = "source footprint is scattered (short segments)"; return; }
    if (plan.sum_tile_elems > 2 * sum_blocks * kSegAlign + 1024) {
        plan.why = "staged footprint over-reads the touched sectors";
        return;
    }
    plan.ok = true;
}

}  // namespace smm
