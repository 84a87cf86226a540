// symbolic.cpp -- host symbolic phase: CSC pattern + stamp maps (compile_structure,
// src/mna/precompile.jl:312-467) and the static-pivot LU schedule that replaces
// KLU's analyze step (call sites src/mna/solve.jl:612-613, src/sweeps.jl:600).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <numeric>
#include <set>

#include "cb200_internal.h"

namespace cb200 {

// ---------------------------------------------------------------------------
// compile_structure: unified pattern = sparse(vcat(G_I,C_I), vcat(G_J,C_J), 1)
// (precompile.jl:410-414): CSC, rows ascending inside a column, duplicates merged,
// explicit zeros kept.  Maps = compute_coo_to_nz_mapping (:253-283).
// ---------------------------------------------------------------------------
std::string build_structure(const cb200_desc &d, Structure &s)
{
    s.n_nodes = d.n_nodes; s.n_currents = d.n_currents;
    s.n_charges = d.n_charges; s.n_limits = d.n_limits;
    s.n = d.n_nodes + d.n_currents + d.n_charges + d.n_limits;
    s.nG = d.nG; s.nC = d.nC; s.nb = d.nb;
    const int n = s.n;
    if (n < 0 || d.nG < 0 || d.nC < 0 || d.nb < 0) return "negative size in description";

    auto check = [&](const int64_t *I, const int64_t *J, int64_t cnt, const char *what) -> std::string {
        for (int64_t k = 0; k < cnt; k++) {
            if (I[k] < 1 || I[k] > n || (J && (J[k] < 1 || J[k] > n)))
                return std::string(what) + ": COO index out of range (ground stamps must be "
                       "dropped before counting, context.jl:945-953)";
        }
        return "";
    };
    std::string e;
    if (!(e = check(d.G_I, d.G_J, d.nG, "G")).empty()) return e;
    if (!(e = check(d.C_I, d.C_J, d.nC, "C")).empty()) return e;
    if (!(e = check(d.b_I, nullptr, d.nb, "b")).empty()) return e;

    // unique (col,row) keys, sorted
    std::vector<int64_t> keys;
    keys.reserve(d.nG + d.nC);
    for (int64_t k = 0; k < d.nG; k++) keys.push_back((d.G_J[k] - 1) * (int64_t)n + (d.G_I[k] - 1));
    for (int64_t k = 0; k < d.nC; k++) keys.push_back((d.C_J[k] - 1) * (int64_t)n + (d.C_I[k] - 1));
    std::sort(keys.begin(), keys.end());
    keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
    s.nnz = (int64_t)keys.size();
    s.colptr.assign(n + 1, 0);
    s.rowval.resize(s.nnz);
    s.nz_row.resize(s.nnz);
    s.nz_col.resize(s.nnz);
    for (int64_t q = 0; q < s.nnz; q++) {
        int col = (int)(keys[q] / n), row = (int)(keys[q] % n);
        s.rowval[q] = row; s.nz_row[q] = row; s.nz_col[q] = col;
        s.colptr[col + 1]++;
    }
    for (int j = 0; j < n; j++) s.colptr[j + 1] += s.colptr[j];

    auto find = [&](int64_t i1, int64_t j1) -> int {
        int64_t key = (j1 - 1) * (int64_t)n + (i1 - 1);
        auto it = std::lower_bound(keys.begin(), keys.end(), key);
        return (int)(it - keys.begin());
    };
    s.G_map.resize(d.nG);
    s.C_map.resize(d.nC);
    for (int64_t k = 0; k < d.nG; k++) s.G_map[k] = find(d.G_I[k], d.G_J[k]);
    for (int64_t k = 0; k < d.nC; k++) s.C_map[k] = find(d.C_I[k], d.C_J[k]);
    s.b_rows.resize(d.nb);
    for (int64_t k = 0; k < d.nb; k++) s.b_rows[k] = (int)d.b_I[k] - 1;

    // _compute_diag_nz_indices (precompile.jl:451-467)
    s.diag_nz.assign(s.n_nodes, -1);
    s.nz_is_node_diag.assign(s.nnz, 0);
    for (int col = 0; col < s.n_nodes; col++)
        for (int q = s.colptr[col]; q < s.colptr[col + 1]; q++)
            if (s.rowval[q] == col) { s.diag_nz[col] = q; s.nz_is_node_diag[q] = 1; break; }

    // segment lists (stable: ascending COO index inside a segment)
    auto segments = [](const std::vector<int> &map, int64_t nseg, std::vector<int> &ptr,
                       std::vector<int> &idx) {
        ptr.assign(nseg + 1, 0);
        for (int m : map) if (m >= 0) ptr[m + 1]++;
        for (int64_t q = 0; q < nseg; q++) ptr[q + 1] += ptr[q];
        idx.resize(ptr[nseg]);
        std::vector<int> cur(ptr.begin(), ptr.end() - 1);
        for (size_t k = 0; k < map.size(); k++) if (map[k] >= 0) idx[cur[map[k]]++] = (int)k;
    };
    segments(s.G_map, s.nnz, s.gseg_ptr, s.gseg_idx);
    segments(s.C_map, s.nnz, s.cseg_ptr, s.cseg_idx);
    segments(s.b_rows, n, s.bseg_ptr, s.bseg_idx);
    return "";
}

// ---------------------------------------------------------------------------
// analyze_lu_sparse: the same job as the dense-bookkeeping analysis below for circuits of any
// size (c6288-class: n ~ 1e5, nnz ~ 1e6).  KLU's recipe restated for a STATIC pivot sequence:
//   1. zero-free diagonal: a row <-> column matching that prefers the (strong) diagonal and pairs
//      the structurally-zero-diagonal rows (voltage-source / branch-current rows) by augmenting
//      paths over entries that pass the magnitude threshold (MC21-style);
//   2. Markowitz minimum (r-1)(c-1) over the matched entries with a lazy heap, sparse rows kept as
//      sorted (column, magnitude) lists and merged at every elimination -- ordering and symbolic
//      fill in one pass, with the magnitude bound |a_ij| += |l| |a_pj| carried along so that a
//      matched entry that has become weak against its column (threshold 1e-3, as below) is replaced
//      by the column's largest entry (off-diagonal pivot, rows re-matched);
//   3. the factor pattern in pivot coordinates, slots, update targets, scatter map.
// Same LuSchedule as the dense path; the level schedule and every kernel are unchanged.
// ---------------------------------------------------------------------------
namespace {
const double kOffDiagThreshold = 0.1;
struct SpEntry { int col; double mag, lo; };     // upper / lower bound of |a_ij| over the probes and the elimination
typedef std::vector<SpEntry> SpRow;

static int sp_find(const SpRow &r, int col)
{
    int lo = 0, hi = (int)r.size();
    while (lo < hi) { int mid = (lo + hi) >> 1; if (r[mid].col < col) lo = mid + 1; else hi = mid; }
    return (lo < (int)r.size() && r[lo].col == col) ? lo : -1;
}
}  // namespace

static std::string analyze_lu_sparse(const Structure &s, const std::vector<double> &absJ,
                                     const std::vector<double> *absJmin, double threshold, LuSchedule &out)
{
    const int n = s.n;
    std::vector<SpRow> rows(n);
    std::vector<std::vector<int>> cols(n);          // rows with an entry in the column (may hold finished rows)
    {
        std::vector<int> cnt(n, 0);
        for (int64_t q = 0; q < s.nnz; q++) cnt[s.nz_row[q]]++;
        for (int i = 0; i < n; i++) rows[i].reserve(cnt[i] + 4);
    }
    for (int j = 0; j < n; j++)                      // CSC order: every row list ends up sorted by column
        for (int q = s.colptr[j]; q < s.colptr[j + 1]; q++) {
            const double m = std::isfinite(absJ[q]) ? absJ[q] : 0.0;
            double lo = absJmin ? (*absJmin)[q] : m;
            if (!std::isfinite(lo) || lo > m) lo = m;
            rows[s.rowval[q]].push_back({j, m, lo});
            cols[j].push_back(s.rowval[q]);
        }
    // ---- 1 + 2. threshold Markowitz over the sparse rows.  Candidates are the entries of the few
    // rows and the few columns with the smallest active counts (the MA28 / Zlatev search: the entry
    // minimising (r-1)(c-1) sits in a short row or a short column); among the cheapest, an entry is
    // admissible when its magnitude bound reaches `threshold` of its column's largest; ties go to
    // the diagonal, then to the relatively larger entry -- the rule of the dense analysis below.
    std::vector<uint8_t> row_done(n, 0), col_done(n, 0);
    std::vector<int> ccount(n, 0);
    for (int j = 0; j < n; j++) ccount[j] = (int)cols[j].size();
    std::set<std::pair<int, int>> rowset, colset;                  // (active count, index)
    for (int i = 0; i < n; i++) { rowset.insert({(int)rows[i].size(), i}); colset.insert({ccount[i], i}); }
    std::vector<double> cmax_val(n, 0.0);
    std::vector<uint8_t> cmax_ok(n, 0);
    auto colmax = [&](int j) -> double {
        if (cmax_ok[j]) return cmax_val[j];
        std::vector<int> &cl = cols[j];
        size_t w = 0;
        double m = 0.0;
        for (size_t t = 0; t < cl.size(); t++) {
            const int r = cl[t];
            if (row_done[r]) continue;
            const int q = sp_find(rows[r], j);
            if (q < 0) continue;
            cl[w++] = r;
            if (rows[r][q].mag > m) m = rows[r][q].mag;
        }
        cl.resize(w);
        cmax_val[j] = m; cmax_ok[j] = 1;
        return m;
    };
    struct Cand { int64_t cost; int row, col; double mag, lo; };
    std::vector<Cand> cand;

    out.rowperm.assign(n, -1); out.colperm.assign(n, -1);
    std::vector<std::vector<int>> Ucols(n), Lrows(n);            // per pivot, ORIGINAL indices
    SpRow merged;
    int n_offdiag = 0;
    for (int k = 0; k < n; k++) {
        int pr = -1, pc = -1;
        // admissibility, strictest first.  Levels 0-2 hold the candidate at its SMALLEST against the column
        // at its LARGEST: 0 -- diagonal entries only, `threshold` of the column (KLU's diagonal preference:
        // an MNA matrix eliminated along its diagonal stays diagonally dominant); 1 -- any entry comparable
        // with its column's largest (0.1); 2 -- any entry at `threshold`; 3 -- any entry that is non-zero at
        // EVERY probe, however small against its column (a switching circuit's off-state conductances:
        // gmin-sized but never zero -- growth is preferable to a pivot that vanishes with the state);
        // 4 -- upper bounds only (the rule of the dense analysis).  Levels 0-3 look at the 64 shortest
        // rows / columns, level 4 at everything.
        for (int level = 0; level < 5 && pr < 0; level++) {
        for (int K = 4; pr < 0; K *= 4) {
            cand.clear();
            int taken = 0;
            for (auto it = rowset.begin(); it != rowset.end() && taken < K; ++it, ++taken) {
                const int i = it->second;
                for (const SpEntry &e : rows[i])
                    cand.push_back({(int64_t)((int)rows[i].size() - 1) * (int64_t)(ccount[e.col] - 1), i, e.col, e.mag, e.lo});
            }
            taken = 0;
            for (auto it = colset.begin(); it != colset.end() && taken < K; ++it, ++taken) {
                const int j = it->second;
                colmax(j);                                        // compacts cols[j] to its active rows
                for (int r : cols[j]) {
                    const int q = sp_find(rows[r], j);
                    cand.push_back({(int64_t)((int)rows[r].size() - 1) * (int64_t)(ccount[j] - 1), r, j, rows[r][q].mag, rows[r][q].lo});
                }
            }
            std::sort(cand.begin(), cand.end(), [](const Cand &x, const Cand &y) { return x.cost < y.cost; });
            size_t g0 = 0;
            while (g0 < cand.size() && pr < 0) {
                size_t g1 = g0;
                while (g1 < cand.size() && cand[g1].cost == cand[g0].cost) g1++;
                bool best_diag = false; double best_rel = -1.0;
                for (size_t t = g0; t < g1; t++) {
                    const Cand &c = cand[t];
                    const double val = level == 4 ? c.mag : c.lo;
                    if (!(val > 0.0)) continue;
                    const bool diag = c.row == c.col;
                    if (level == 0 && !diag) continue;
                    const double cm = colmax(c.col);
                    if (level != 3 && val < ((level == 1 && !diag) ? kOffDiagThreshold : threshold) * cm) continue;
                    const double rel = val / cm;
                    if (pr < 0 || (diag && !best_diag) || (diag == best_diag && rel > best_rel * (1.0 + 1e-12))) {
                        pr = c.row; pc = c.col; best_diag = diag; best_rel = rel;
                    }
                }
                g0 = g1;
            }
            if (pr < 0 && (K >= n || (level < 4 && K >= 64))) break;
        }
        }
        if (pr < 0) return "analyze_lu: matrix is singular at the probe points (no admissible pivot at step " +
                           std::to_string(k) + ")";
        if (pr != pc) n_offdiag++;
        out.rowperm[k] = pr; out.colperm[k] = pc;
        row_done[pr] = 1; col_done[pc] = 1;
        rowset.erase({(int)rows[pr].size(), pr});
        colset.erase({ccount[pc], pc});
        const SpRow prow = rows[pr];                                 // copy: the pivot row (U row k + diagonal)
        const int qp = sp_find(prow, pc);
        const double piv_hi = prow[qp].mag, piv_lo = prow[qp].lo;
        for (const SpEntry &e : prow)
            if (e.col != pc) {
                Ucols[k].push_back(e.col);
                colset.erase({ccount[e.col], e.col});
                ccount[e.col]--;                                     // re-inserted below, after the fill is known
                cmax_ok[e.col] = 0;
            }
        // eliminate column pc from every other active row
        for (int r : cols[pc]) {
            if (row_done[r]) continue;
            SpRow &rr = rows[r];
            const int q = sp_find(rr, pc);
            if (q < 0) continue;
            Lrows[k].push_back(r);
            rowset.erase({(int)rr.size(), r});
            // multiplier at its largest / smallest; a pivot admitted on its upper bound only (level 3)
            // has no usable lower bound: the optimistic value stands in
            const double l = rr[q].mag / (piv_lo > 0.0 ? piv_lo : piv_hi), l_lo = rr[q].lo / piv_hi;
            merged.clear();
            merged.reserve(rr.size() + prow.size());
            size_t a = 0, b = 0;
            while (a < rr.size() || b < prow.size()) {
                if (b >= prow.size() || (a < rr.size() && rr[a].col < prow[b].col)) {
                    if (rr[a].col != pc) merged.push_back(rr[a]);
                    a++;
                } else if (a >= rr.size() || prow[b].col < rr[a].col) {
                    if (prow[b].col != pc) {                         // fill
                        merged.push_back({prow[b].col, l * prow[b].mag, l_lo * prow[b].lo});
                        cols[prow[b].col].push_back(r);
                        ccount[prow[b].col]++;
                    }
                    b++;
                } else {
                    // the lower bound of an updated entry is kept (the updates of an MNA matrix add to the
                    // off-diagonals and leave the diagonals dominant; cancellation is what the run-time
                    // weak-pivot check is for)
                    if (rr[a].col != pc) merged.push_back({rr[a].col, rr[a].mag + l * prow[b].mag, rr[a].lo});
                    a++; b++;
                }
            }
            rr.swap(merged);
            rowset.insert({(int)rr.size(), r});
        }
        for (int c : Ucols[k]) colset.insert({ccount[c], c});
        std::vector<int>().swap(cols[pc]);
        SpRow().swap(rows[pr]);
    }

    if (getenv("CB200_ANALYZE_DEBUG")) {
        int64_t nl = 0, nu = 0;
        for (int k = 0; k < n; k++) { nl += (int64_t)Lrows[k].size(); nu += (int64_t)Ucols[k].size(); }
        fprintf(stderr, "analyze_lu_sparse: n %d nnz %lld L %lld U %lld off-diagonal pivots %d\n", n, (long long)s.nnz,
                (long long)nl, (long long)nu, n_offdiag);
    }
    // ---- 3. factor pattern in pivot coordinates
    std::vector<int> rinv(n), cinv(n);
    for (int k = 0; k < n; k++) { rinv[out.rowperm[k]] = k; cinv[out.colperm[k]] = k; }
    out.diag_slot.resize(n);
    out.Uptr.assign(n + 1, 0);
    out.Lptr.assign(n + 1, 0);
    int64_t nlu = 0;
    for (int k = 0; k < n; k++) {
        for (int &c : Ucols[k]) c = cinv[c];
        for (int &r : Lrows[k]) r = rinv[r];
        std::sort(Ucols[k].begin(), Ucols[k].end());
        std::sort(Lrows[k].begin(), Lrows[k].end());
        out.diag_slot[k] = (int)nlu++;
        for (int b : Ucols[k]) { out.U_slot.push_back((int)nlu++); out.U_col.push_back(b); }
        out.Uptr[k + 1] = (int)out.U_slot.size();
        for (int a : Lrows[k]) { out.L_slot.push_back((int)nlu++); out.L_row.push_back(a); }
        out.Lptr[k + 1] = (int)out.L_slot.size();
        if (nlu > (int64_t)2000000000) return "analyze_lu: factor too large for 32-bit slots";
    }
    out.nlu = nlu;
    auto slot_of = [&](int a, int b) -> int {                        // (a, b) in pivot coordinates
        if (a == b) return out.diag_slot[a];
        if (a < b) {
            const int *lo = &out.U_col[0] + out.Uptr[a], *hi = &out.U_col[0] + out.Uptr[a + 1];
            const int *it = std::lower_bound(lo, hi, b);
            return (it != hi && *it == b) ? out.U_slot[it - &out.U_col[0]] : -1;
        }
        const int *lo = &out.L_row[0] + out.Lptr[b], *hi = &out.L_row[0] + out.Lptr[b + 1];
        const int *it = std::lower_bound(lo, hi, a);
        return (it != hi && *it == a) ? out.L_slot[it - &out.L_row[0]] : -1;
    };
    out.tgt_ptr.assign(n + 1, 0);
    int64_t ntgt = 0;
    for (int k = 0; k < n; k++) ntgt += (int64_t)(out.Lptr[k + 1] - out.Lptr[k]) * (out.Uptr[k + 1] - out.Uptr[k]);
    if (ntgt > (int64_t)2000000000) return "analyze_lu: update list too large";
    out.tgt.reserve((size_t)ntgt);
    for (int k = 0; k < n; k++) {
        const int nL = out.Lptr[k + 1] - out.Lptr[k], nU = out.Uptr[k + 1] - out.Uptr[k];
        for (int e = 0; e < nL; e++) {
            const int a = out.L_row[out.Lptr[k] + e];
            for (int q = 0; q < nU; q++) {
                const int t = slot_of(a, out.U_col[out.Uptr[k] + q]);
                if (t < 0) return "analyze_lu: internal error (missing fill slot)";
                out.tgt.push_back(t);
            }
        }
        out.tgt_ptr[k + 1] = (int)out.tgt.size();
        out.flops += (int64_t)nL * nU;
    }
    out.jmap.resize(s.nnz);
    std::vector<uint8_t> has_src((size_t)nlu, 0);
    for (int64_t q = 0; q < s.nnz; q++) {
        const int t = slot_of(rinv[s.nz_row[q]], cinv[s.nz_col[q]]);
        if (t < 0) return "analyze_lu: internal error (pattern entry without slot)";
        out.jmap[q] = t; has_src[t] = 1;
    }
    for (int64_t t = 0; t < nlu; t++) if (!has_src[t]) out.fill_slots.push_back((int)t);
    out.valid = true;
    return "";
}

// ---------------------------------------------------------------------------
// analyze_lu: threshold Markowitz pivoting on nominal magnitudes, symbolic fill,
// elimination schedule.  The pivot sequence is fixed for every lane and every
// Newton iteration (klu_refactor-style reuse); the numeric kernels flag lanes whose
// pivots vanish (CB200_LANE_SINGULAR).
// Dense working matrices: this path serves the small/medium-circuit regime
// (n <= kDenseLimit); the level-scheduled supernodal path for c6288-class
// circuits is a separate, later tier.
// ---------------------------------------------------------------------------
static const int kDenseLimit = 1536;

std::string analyze_lu(const Structure &s, const std::vector<double> &absJ, double threshold,
                       LuSchedule &out, const std::vector<double> *absJmin)
{
    const int n = s.n;
    out = LuSchedule();
    out.n = n;
    if (n == 0) { out.valid = true; return ""; }
    if ((int64_t)absJ.size() != s.nnz) return "analyze_lu: magnitude array has wrong length";
    if (n > kDenseLimit || getenv("CB200_SPARSE_ANALYZE")) return analyze_lu_sparse(s, absJ, absJmin, threshold, out);

    // pat: 0 absent, 1 structural; mag: nominal magnitude
    std::vector<uint8_t> pat((size_t)n * n, 0);
    std::vector<double> mag((size_t)n * n, 0.0);
    for (int j = 0; j < n; j++)
        for (int q = s.colptr[j]; q < s.colptr[j + 1]; q++) {
            pat[(size_t)s.rowval[q] * n + j] = 1;
            mag[(size_t)s.rowval[q] * n + j] = std::isfinite(absJ[q]) ? absJ[q] : 0.0;
        }
    std::vector<uint8_t> row_done(n, 0), col_done(n, 0);
    std::vector<int> rcount(n, 0), ccount(n, 0);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++)
            if (pat[(size_t)i * n + j]) { rcount[i]++; ccount[j]++; }
    out.rowperm.resize(n);
    out.colperm.resize(n);

    for (int k = 0; k < n; k++) {
        // column maxima over the active submatrix
        std::vector<double> colmax(n, 0.0);
        for (int i = 0; i < n; i++) {
            if (row_done[i]) continue;
            const double *mr = &mag[(size_t)i * n];
            for (int j = 0; j < n; j++)
                if (!col_done[j] && mr[j] > colmax[j]) colmax[j] = mr[j];
        }
        long best_cost = -1; int bi = -1, bj = -1; double best_rel = -1.0; bool best_diag = false;
        for (int i = 0; i < n; i++) {
            if (row_done[i]) continue;
            for (int j = 0; j < n; j++) {
                if (col_done[j] || !pat[(size_t)i * n + j]) continue;
                double m = mag[(size_t)i * n + j];
                if (!(m > 0.0) || m < threshold * colmax[j]) continue;
                long cost = (long)(rcount[i] - 1) * (long)(ccount[j] - 1);
                double rel = m / colmax[j];
                bool diag = (i == j);
                bool better = false;
                if (best_cost < 0 || cost < best_cost) better = true;
                else if (cost == best_cost) {
                    if (diag && !best_diag) better = true;
                    else if (diag == best_diag && rel > best_rel * (1.0 + 1e-12)) better = true;
                }
                if (better) { best_cost = cost; bi = i; bj = j; best_rel = rel; best_diag = diag; }
            }
        }
        if (bi < 0) return "analyze_lu: matrix is singular at the probe points (no admissible pivot "
                           "at step " + std::to_string(k) + ")";
        out.rowperm[k] = bi; out.colperm[k] = bj;
        row_done[bi] = 1; col_done[bj] = 1;
        const double piv = mag[(size_t)bi * n + bj];
        // eliminate: fill + magnitude bound
        for (int j = 0; j < n; j++) if (!col_done[j] && pat[(size_t)bi * n + j]) ccount[j]--;
        for (int i = 0; i < n; i++) {
            if (row_done[i] || !pat[(size_t)i * n + bj]) continue;
            rcount[i]--;
            double l = mag[(size_t)i * n + bj] / piv;
            for (int j = 0; j < n; j++) {
                if (col_done[j] || !pat[(size_t)bi * n + j]) continue;
                size_t ij = (size_t)i * n + j;
                if (!pat[ij]) { pat[ij] = 2; rcount[i]++; ccount[j]++; }
                mag[ij] += l * mag[(size_t)bi * n + j];
            }
        }
    }

    // Factor pattern in pivot coordinates: B[a][b] = A[rowperm[a]][colperm[b]].
    std::vector<int> rinv(n), cinv(n);
    for (int k = 0; k < n; k++) { rinv[out.rowperm[k]] = k; cinv[out.colperm[k]] = k; }
    std::vector<int> slot((size_t)n * n, -1);
    int64_t nlu = 0;
    // slot order: for every pivot k, its diagonal, then U row k, then L column k
    out.diag_slot.resize(n);
    out.Uptr.assign(n + 1, 0);
    out.Lptr.assign(n + 1, 0);
    for (int k = 0; k < n; k++) {
        int ri = out.rowperm[k], cj = out.colperm[k];
        slot[(size_t)k * n + k] = (int)nlu; out.diag_slot[k] = (int)nlu; nlu++;
        for (int b = k + 1; b < n; b++) {
            if (pat[(size_t)ri * n + out.colperm[b]]) {
                slot[(size_t)k * n + b] = (int)nlu;
                out.U_slot.push_back((int)nlu); out.U_col.push_back(b); nlu++;
            }
        }
        out.Uptr[k + 1] = (int)out.U_slot.size();
        for (int a = k + 1; a < n; a++) {
            if (pat[(size_t)out.rowperm[a] * n + cj]) {
                slot[(size_t)a * n + k] = (int)nlu;
                out.L_slot.push_back((int)nlu); out.L_row.push_back(a); nlu++;
            }
        }
        out.Lptr[k + 1] = (int)out.L_slot.size();
    }
    out.nlu = nlu;
    // update targets
    out.tgt_ptr.assign(n + 1, 0);
    for (int k = 0; k < n; k++) {
        int nL = out.Lptr[k + 1] - out.Lptr[k], nU = out.Uptr[k + 1] - out.Uptr[k];
        for (int e = 0; e < nL; e++) {
            int a = out.L_row[out.Lptr[k] + e];
            for (int q = 0; q < nU; q++) {
                int b = out.U_col[out.Uptr[k] + q];
                int t = slot[(size_t)a * n + b];
                if (t < 0) return "analyze_lu: internal error (missing fill slot)";
                out.tgt.push_back(t);
            }
        }
        out.tgt_ptr[k + 1] = (int)out.tgt.size();
        out.flops += (int64_t)nL * nU;
    }
    // scatter map and fill list
    out.jmap.resize(s.nnz);
    std::vector<uint8_t> has_src(nlu, 0);
    for (int64_t q = 0; q < s.nnz; q++) {
        int t = slot[(size_t)rinv[s.nz_row[q]] * n + cinv[s.nz_col[q]]];
        if (t < 0) return "analyze_lu: internal error (pattern entry without slot)";
        out.jmap[q] = t; has_src[t] = 1;
    }
    for (int64_t t = 0; t < nlu; t++) if (!has_src[t]) out.fill_slots.push_back((int)t);
    out.valid = true;
    return "";
}

// ---------------------------------------------------------------------------
// Level schedule for the lane-per-warp kernels (see LevelSchedule).
// ---------------------------------------------------------------------------
void build_level_schedule(const LuSchedule &S, LevelSchedule &out)
{
    out = LevelSchedule();
    const int n = S.n;
    // ---- factorisation levels.  Pivot k may run once every update into its row, column and
    // diagonal has been applied (lev[min(i, j)] > lev[k] for each target (i, j) of k); and the
    // updates of one target must keep ascending pivot order, so a pivot never runs at a level
    // below the last pivot that touched one of its targets.
    std::vector<int> lev(n, 0), last_lev((size_t)S.nlu, -1);
    for (int k = 0; k < n; k++) {
        const int l0 = S.Lptr[k], nL = S.Lptr[k + 1] - l0, u0 = S.Uptr[k], nU = S.Uptr[k + 1] - u0;
        const int t0 = S.tgt_ptr[k];
        for (int x = 0; x < nL * nU; x++) lev[k] = std::max(lev[k], last_lev[S.tgt[t0 + x]]);
        for (int e = 0; e < nL; e++)
            for (int q = 0; q < nU; q++) {
                last_lev[S.tgt[t0 + e * nU + q]] = lev[k];
                const int m = std::min(S.L_row[l0 + e], S.U_col[u0 + q]);
                lev[m] = std::max(lev[m], lev[k] + 1);
            }
    }
    int nlev = 0;
    for (int k = 0; k < n; k++) nlev = std::max(nlev, lev[k] + 1);
    out.n_lev = nlev;
    std::vector<std::vector<int>> by_lev(nlev);
    for (int k = 0; k < n; k++) by_lev[lev[k]].push_back(k);
    out.piv_ptr.push_back(0); out.sc_ptr.push_back(0); out.tg_ptr.push_back(0);
    std::vector<int> tg_of((size_t)S.nlu, -1);          // target slot -> index in this level's list
    for (int v = 0; v < nlev; v++) {
        std::vector<int> tslots;
        std::vector<std::vector<int>> ups;               // per target: {ls, us, k} triples, ascending k
        for (int k : by_lev[v]) {
            out.piv.push_back(k); out.piv.push_back(S.diag_slot[k]);
            const int l0 = S.Lptr[k], nL = S.Lptr[k + 1] - l0, u0 = S.Uptr[k], nU = S.Uptr[k + 1] - u0;
            for (int e = 0; e < nL; e++) { out.sc.push_back(S.L_slot[l0 + e]); out.sc.push_back(k); }
            for (int e = 0; e < nL; e++)
                for (int q = 0; q < nU; q++) {
                    const int t = S.tgt[S.tgt_ptr[k] + e * nU + q];
                    if (tg_of[t] < 0) { tg_of[t] = (int)tslots.size(); tslots.push_back(t); ups.emplace_back(); }
                    std::vector<int> &u = ups[tg_of[t]];
                    u.push_back(S.L_slot[l0 + e]); u.push_back(S.U_slot[u0 + q]); u.push_back(k);
                }
        }
        for (size_t i = 0; i < tslots.size(); i++) {
            const int ub = (int)(out.upd.size() / 4);
            for (size_t x = 0; x < ups[i].size(); x += 3) {
                out.upd.push_back(ups[i][x]); out.upd.push_back(ups[i][x + 1]);
                out.upd.push_back(ups[i][x + 2]); out.upd.push_back(0);
            }
            out.tg.push_back(tslots[i]); out.tg.push_back(ub);
            out.tg.push_back((int)(out.upd.size() / 4)); out.tg.push_back(0);
            tg_of[tslots[i]] = -1;
        }
        out.piv_ptr.push_back((int)(out.piv.size() / 2));
        out.sc_ptr.push_back((int)(out.sc.size() / 2));
        out.tg_ptr.push_back((int)(out.tg.size() / 4));
    }
    // ---- forward solve z = L^-1 P F by rows: z[i] = F[rowperm[i]] - sum_k L[i][k] z[k], k ascending
    std::vector<int> zl(n, 0);
    std::vector<std::vector<int>> lrow(n);
    for (int k = 0; k < n; k++)
        for (int e = S.Lptr[k]; e < S.Lptr[k + 1]; e++) {
            const int i = S.L_row[e];
            zl[i] = std::max(zl[i], zl[k] + 1);
            lrow[i].push_back(S.L_slot[e]); lrow[i].push_back(k);
        }
    int nf = 0;
    for (int i = 0; i < n; i++) nf = std::max(nf, zl[i] + 1);
    out.n_fwd = nf;
    out.flev_ptr.push_back(0);
    for (int v = 0; v < nf; v++) {
        for (int i = 0; i < n; i++) {
            if (zl[i] != v) continue;
            const int b = (int)(out.fent.size() / 2);
            out.fent.insert(out.fent.end(), lrow[i].begin(), lrow[i].end());
            out.frow.push_back(i); out.frow.push_back(S.rowperm[i]);
            out.frow.push_back(b); out.frow.push_back((int)(out.fent.size() / 2));
        }
        out.flev_ptr.push_back((int)(out.frow.size() / 4));
    }
    // ---- backward solve y = U^-1 z by rows, last pivot first
    std::vector<int> xl(n, 0);
    int nb = 0;
    for (int k = n - 1; k >= 0; k--) {
        for (int q = S.Uptr[k]; q < S.Uptr[k + 1]; q++) xl[k] = std::max(xl[k], xl[S.U_col[q]] + 1);
        nb = std::max(nb, xl[k] + 1);
    }
    out.n_bwd = nb;
    out.blev_ptr.push_back(0);
    for (int v = 0; v < nb; v++) {
        for (int k = n - 1; k >= 0; k--) {
            if (xl[k] != v) continue;
            const int b = (int)(out.bent.size() / 2);
            for (int q = S.Uptr[k]; q < S.Uptr[k + 1]; q++) { out.bent.push_back(S.U_slot[q]); out.bent.push_back(S.U_col[q]); }
            out.brow.push_back(k); out.brow.push_back(b);
            out.brow.push_back((int)(out.bent.size() / 2)); out.brow.push_back(0);
        }
        out.blev_ptr.push_back((int)(out.brow.size() / 4));
    }
    if (getenv("CB200_DEBUG_LU"))
        fprintf(stderr, "[cb200] level schedule: n=%d nlu=%lld factor levels=%d (targets=%zu updates=%zu) "
                        "forward levels=%d backward levels=%d\n", n, (long long)S.nlu, nlev,
                out.tg.size() / 4, out.upd.size() / 4, nf, nb);
}

// ---------------------------------------------------------------------------
// Host execution of both schedules on one matrix (test hook, cb200_host_lu_check): the serial
// loop nest is factor_and_solve of lane_kernels.cuh, the level loop nest w_factor_and_solve of
// warp_kernels.cuh, statement for statement.  x solves J x = rhs (original coordinates).
// ---------------------------------------------------------------------------
void host_lu_solve(const LuSchedule &S, const LevelSchedule &V, const std::vector<double> &J,
                   const std::vector<double> &rhs, std::vector<double> &x_serial, std::vector<double> &x_level)
{
    const int n = S.n;
    {
        std::vector<double> LU((size_t)S.nlu, 0.0), wv(n);
        for (size_t q = 0; q < J.size(); q++) LU[S.jmap[q]] = J[q];
        for (int k = 0; k < n; k++) {
            const double inv = 1.0 / LU[S.diag_slot[k]];
            LU[S.diag_slot[k]] = inv;
            const int l0 = S.Lptr[k], l1 = S.Lptr[k + 1], u0 = S.Uptr[k], u1 = S.Uptr[k + 1], t0 = S.tgt_ptr[k];
            for (int e = l0; e < l1; e++) {
                const double l = LU[S.L_slot[e]] * inv;
                LU[S.L_slot[e]] = l;
                for (int q = u0; q < u1; q++) {
                    const int ts = S.tgt[t0 + (e - l0) * (u1 - u0) + (q - u0)];
                    LU[ts] = LU[ts] - l * LU[S.U_slot[q]];
                }
            }
        }
        for (int k = 0; k < n; k++) wv[k] = rhs[S.rowperm[k]];
        for (int k = 0; k < n; k++) {
            const double zk = wv[k];
            for (int e = S.Lptr[k]; e < S.Lptr[k + 1]; e++) wv[S.L_row[e]] = wv[S.L_row[e]] - LU[S.L_slot[e]] * zk;
        }
        for (int k = n - 1; k >= 0; k--) {
            double acc = wv[k];
            for (int q = S.Uptr[k]; q < S.Uptr[k + 1]; q++) acc -= LU[S.U_slot[q]] * wv[S.U_col[q]];
            acc *= LU[S.diag_slot[k]];
            wv[k] = acc;
        }
        x_serial.assign(n, 0.0);
        for (int k = 0; k < n; k++) x_serial[S.colperm[k]] = wv[k];
    }
    {
        std::vector<double> LU((size_t)S.nlu, 0.0), DI(n), wv(n), F(rhs);
        for (size_t q = 0; q < J.size(); q++) LU[S.jmap[q]] = J[q];
        for (int v = 0; v <= V.n_lev; v++) {
            if (v < V.n_lev)
                for (int q = V.piv_ptr[v]; q < V.piv_ptr[v + 1]; q++) DI[V.piv[2 * q]] = 1.0 / LU[V.piv[2 * q + 1]];
            if (v > 0)
                for (int q = V.sc_ptr[v - 1]; q < V.sc_ptr[v]; q++) LU[V.sc[2 * q]] = LU[V.sc[2 * q]] * DI[V.sc[2 * q + 1]];
            if (v < V.n_lev)
                for (int q = V.tg_ptr[v]; q < V.tg_ptr[v + 1]; q++) {
                    double a = LU[V.tg[4 * q]];
                    for (int u = V.tg[4 * q + 1]; u < V.tg[4 * q + 2]; u++) {
                        const double lv = LU[V.upd[4 * u]] * DI[V.upd[4 * u + 2]];
                        a = a - lv * LU[V.upd[4 * u + 1]];
                    }
                    LU[V.tg[4 * q]] = a;
                }
        }
        for (int v = 0; v < V.n_fwd; v++)
            for (int q = V.flev_ptr[v]; q < V.flev_ptr[v + 1]; q++) {
                double acc = F[V.frow[4 * q + 1]];
                for (int e = V.frow[4 * q + 2]; e < V.frow[4 * q + 3]; e++) acc = acc - LU[V.fent[2 * e]] * wv[V.fent[2 * e + 1]];
                wv[V.frow[4 * q]] = acc;
            }
        std::vector<double> y(n);
        for (int v = 0; v < V.n_bwd; v++)
            for (int q = V.blev_ptr[v]; q < V.blev_ptr[v + 1]; q++) {
                const int k = V.brow[4 * q];
                double acc = wv[k];
                for (int e = V.brow[4 * q + 1]; e < V.brow[4 * q + 2]; e++) acc -= LU[V.bent[2 * e]] * y[V.bent[2 * e + 1]];
                acc *= DI[k];
                y[k] = acc;
            }
        x_level.assign(n, 0.0);
        for (int k = 0; k < n; k++) x_level[S.colperm[k]] = y[k];
    }
}

}  // namespace cb200
