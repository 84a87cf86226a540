// warp_kernels.cuh -- the batched MNA Newton / transient hot path for MEDIUM circuits:
// one sweep lane per WARP.
//
// lane_kernels.cuh gives every lane one thread.  That is the right mapping while the lane state
// fits in registers / shared memory and the sweep has >> 148 * 2048 lanes; a circuit such as
// the gf180 D flip-flop (30 Verilog-A FETs, n = 145, ~5000 workspace doubles, 16 384 lanes) then
// runs 110 threads per SM, each walking its own 40 KB column of HBM -- pure latency.  Here the 32
// threads of a warp share ONE lane:
//   * device evaluation: thread i evaluates device instance i, i + 32, ... (30 FETs = one round);
//   * assembly: thread per matrix entry sums its stamp segment in program order, then thread per
//     row forms F = C*du + G*u - b over the row's entries in column order;
//   * LU refactor on the host's static pivot order, cut into LEVELS of independent pivots: per
//     level one thread per pivot (reciprocal), one per L entry (scaling), one per target entry
//     (its updates in ascending pivot order), two __syncwarp per level;
//   * triangular solves by rows, one thread per row of a level, one __syncwarp per level;
//   * norms are summed by every thread over ALL rows in index order.
// Every floating-point operation is performed with the same operands in the same order as in
// lane_kernels.cuh (only WHICH thread performs it changes), so the two mappings agree bit for bit;
// tests/test_gpu_parity.py asserts identical waveforms and Newton iteration counts.
// The lane's workspace is a contiguous row ws[lane][slot] of a global array: consecutive slots
// are consecutive addresses, a warp's accesses coalesce, and the resident working set
// (16 warps/SM * 148 SMs * 40 KB = 95 MB) lives in the 126 MB L2.  The arrays on the
// DEPENDENT chains of the refactor and the triangular solves -- the LU factor, the residual /
// solution vector, the work vector and the reciprocal pivots (3n + nnz_LU doubles, 9 KB for the
// flip-flop) -- are staged in SHARED memory (struct Hot): a pivot step is two dependent memory
// round trips, and at L2 latency those 145 * 2 * 3 round trips were the whole iteration
// (measured: 3.75 s per C4 sweep with the factor in L2; see profiles/README.md).
#pragma once
#include "lane_kernels.cuh"

namespace cb200 {

#define CB_FULL 0xffffffffu

// latency-critical arrays of a lane: shared memory when they fit (generic pointers: the same
// code runs with them left in the lane's global row)
// The read-only tables the warp kernels walk on every Newton iteration (segment lists, CSR view,
// level schedule: 43 KB for the flip-flop).  Every data access of the assembly and of the
// refactor is index -> data, so an index that misses the L1 (the lane rows and the model body's
// spills compete for it) puts an L2 round trip on the dependent chain.  A block CAN stage the
// tables in shared memory once (stage_tables, kernels.cu; CB200_WARP_STAGE_TABLES=1) -- measured
// on C4 that loses (2.00 s vs 1.34 s): the shared memory it takes from the L1 costs more than the
// guaranteed index hits save -- so by default WTab points at the global arrays.  The struct
// itself lives in shared memory.
struct WTab {
    const int *gseg_ptr, *gseg_idx, *cseg_ptr, *cseg_idx, *bseg_ptr, *bseg_idx, *rowptr, *row_nz, *nz_col;
    const int *jmap, *fill_slots, *colperm;
    const int *piv_ptr, *sc_ptr, *tg_ptr, *flev_ptr, *blev_ptr;
    const int2 *piv, *sc, *fent, *bent;
    const int4 *tg, *upd, *frow, *brow;
};

struct Hot {
    double *F, *wv, *DI, *LU;
    double *buf;      // shared-memory gather buffer of the segmented assembly, ch doubles
    int ch;
    const WTab *tb;   // in shared memory
};

__device__ __forceinline__ int wtid() { return threadIdx.x & 31; }

template <typename PG, typename W>
__device__ __forceinline__ void w_eval_all(const PG &pg, W &w, double t, int mode, bool initjct)
{
    for (int d = wtid(); d < pg.n_dev(); d += 32) eval_device<0>(pg, w, d, t, mode, initjct);
    __syncwarp();
}

template <typename PG, typename W>
__device__ __forceinline__ void w_eval_nonlinear(const PG &pg, W &w, double t, int mode, bool initjct)
{
    for (int q = wtid(); q < pg.n_nl(); q += 32) eval_device<1>(pg, w, pg.nl_list(q), t, mode, initjct);
    __syncwarp();
}

template <typename PG, typename W>
__device__ __forceinline__ void w_eval_sources(const PG &pg, W &w, double t, int mode)
{
    for (int q = wtid(); q < pg.n_src(); q += 32) eval_device<2>(pg, w, pg.src_list(q), t, mode, false);
    __syncwarp();
}

// Segmented sums of the assembly.  A segment (all stamps of one matrix entry, in program order)
// must be added up left to right to reproduce the reference's `nzval[idx] += v`, and segment
// lengths are wildly skewed (the flip-flop's (VDD, VDD) entry collects ~900 stamps, the median
// entry 4), so one thread walking `w(off + idx[q])` serialises ~900 dependent global-memory
// round trips -- measured: 54 % of the kernel at 3.8 active threads.  Instead the warp gathers
// the stamps of `ch` consecutive segment positions into shared memory with all 32 threads
// (independent loads, one latency for the lot), and each thread then adds up the part of ITS
// segments that lies in the chunk from shared memory, carrying the running sum of a segment that
// spans chunks in a register.  Same addition order, two orders of magnitude less latency.
template <typename W, typename PtrF, typename IdxF, typename Fin>
__device__ __forceinline__ void w_segsum(W &w, const Hot &hot, int off, int nseg, int total, PtrF ptr,
                                         IdxF idx, Fin finish)
{
    const int tl = wtid();
    int s = tl, p = 0, pend = 0;
    double acc = 0.0;
    if (s < nseg) { p = ptr(s); pend = ptr(s + 1); }
    for (int c0 = 0; c0 < total; c0 += hot.ch) {
        const int cend = min(c0 + hot.ch, total);
        for (int j0 = c0 + tl; j0 < cend; j0 += 32 * 8) {     // 8 independent loads in flight per thread
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) { const int j = j0 + 32 * u; v[u] = j < cend ? w(off + idx(j)) : 0.0; }
#pragma unroll
            for (int u = 0; u < 8; u++) { const int j = j0 + 32 * u; if (j < cend) hot.buf[j - c0] = v[u]; }
        }
        __syncwarp();
        while (s < nseg) {
            const int hi = pend < cend ? pend : cend;
            // loads first, dependent adds after: the chain then runs at DADD latency per element
            const double *bp = hot.buf + (p - c0);
            int cnt = hi - p;
            for (; cnt >= 8; cnt -= 8, bp += 8) {
                const double v0 = bp[0], v1 = bp[1], v2 = bp[2], v3 = bp[3];
                const double v4 = bp[4], v5 = bp[5], v6 = bp[6], v7 = bp[7];
                acc += v0; acc += v1; acc += v2; acc += v3; acc += v4; acc += v5; acc += v6; acc += v7;
            }
            for (; cnt > 0; cnt--, bp++) acc += *bp;
            if (hi > p) p = hi;
            if (pend > cend) break;                         // the segment continues in the next chunk
            finish(s, acc);
            s += 32; acc = 0.0;
            if (s < nseg) { p = ptr(s); pend = ptr(s + 1); }
        }
        __syncwarp();
    }
    for (; s < nseg; s += 32) finish(s, 0.0);               // total == 0: every segment is empty
}

// assemble() of lane_kernels.cuh: segmented sums (above), Jacobian scatter entry-parallel,
// residual row-parallel.
template <bool TRAN, typename PG, typename LU, typename W>
__device__ __forceinline__ double w_assemble(const PG &pg, const LU &lu, W &w, const Hot &hot,
                                             double gamma, double gshunt, double srcFact, bool &bad,
                                             bool have_gc = false)
{
    const int tl = wtid();
    const int oGS = pg.off_GS(), oCS = pg.off_CS();
    const WTab &tb = *hot.tb;
    for (int q = tl; q < lu.n_fill(); q += 32) hot.LU[tb.fill_slots[q]] = 0.0;
    // have_gc: the per-entry sums in the lane row are those of the current iterate already (the
    // first residual of a time step, see `fresh` in lane_kernels.cuh: tran_fixed_body)
    if (!have_gc) {   // table pointers are copied out of the shared WTab once per phase (no reload per access)
        const int *sp = tb.gseg_ptr, *si = tb.gseg_idx;
        w_segsum(w, hot, pg.off_SG(), pg.nnz(), pg.p.nG,
                 [&](int s) { return sp[s]; }, [&](int q) { return si[q]; },
                 [&](int s, double v) {
                     if (gshunt != 0.0 && pg.nz_is_node_diag(s)) v += gshunt;
                     w(oGS + s) = v;
                 });
    }
    if (TRAN && !have_gc) {
        const int *sp = tb.cseg_ptr, *si = tb.cseg_idx;
        w_segsum(w, hot, pg.off_SC(), pg.nnz(), pg.p.nC,
                 [&](int s) { return sp[s]; }, [&](int q) { return si[q]; },
                 [&](int s, double v) { w(oCS + s) = v; });
    }
    {
        const int *sp = tb.bseg_ptr, *si = tb.bseg_idx;
        w_segsum(w, hot, pg.off_SB(), pg.n(), pg.p.nb,
                 [&](int r) { return sp[r]; }, [&](int q) { return si[q]; },
                 [&](int r, double v) { hot.wv[r] = v; });      // wv is free until the solve
    }
    const int *const jmap = tb.jmap;
    for (int s = tl; s < pg.nnz(); s += 32) {               // thread s % 32 reads back its own sums
        double jv = w(oGS + s);
        if (TRAN) jv += gamma * w(oCS + s);
        hot.LU[jmap[s]] = jv;
    }
    __syncwarp();
    // Residual by rows: one thread per row, the row's entries in column order (a cooperative
    // shared-memory staging of the operands, as for the segment sums, measured 6 % SLOWER on C4).
    const int *const rowptr = tb.rowptr, *const row_nz = tb.row_nz, *const nz_col = tb.nz_col;
    for (int r = tl; r < pg.n(); r += 32) {
        double f = 0.0;
        const int q1 = rowptr[r + 1];
        for (int q = rowptr[r]; q < q1; q++) {
            const int s = row_nz[q], j = nz_col[s];
            const double uj = w(pg.off_u() + j);
            if (TRAN) {
                const double duj = gamma * (uj - w(pg.off_un() + j)) + w(pg.off_dterm() + j);
                f += w(oCS + s) * duj;
            }
            f += w(oGS + s) * uj;
        }
        double bsum = hot.wv[r];
        if (srcFact < 1.0) bsum *= srcFact;
        hot.F[r] = f - bsum;
    }
    __syncwarp();
    double nrm2 = 0.0;
    for (int r = 0; r < pg.n(); r++) {                      // every thread, index order (broadcast loads)
        const double f = hot.F[r];
        nrm2 += f * f;
    }
    bad = !isfinite(nrm2);
    return nrm2;
}

// factor_and_solve() of lane_kernels.cuh on the LEVEL schedule (LevelSchedule, cb200_internal.h):
// the pivots of a level are independent, so a level is two warp-wide phases --
//   A: 1/pivot for the level's pivots (to its own array DI, so that no thread can see it in place
//      of the pivot) and, one level late, the scaling of the previous level's L columns;
//   B: one thread per target entry applies that entry's updates a -= (L_raw * 1/pivot) * U in
//      ascending pivot order;
// then the triangular solves by rows, one thread per row of a level.  The solution ends up in
// the F slots (pivot coordinates).  Same operands, same order per entry as the serial schedule.
template <typename PG, typename LU, typename W>
__device__ __forceinline__ bool w_factor_and_solve(const PG &pg, const LU &lu, W &w, const Hot &hot,
                                                   bool &singular)
{
    const int tl = wtid();
    const LuProgram &l = lu.l;
    const WTab &tb = *hot.tb;
    double *const sLU = hot.LU, *const sDI = hot.DI, *const sWV = hot.wv, *const sF = hot.F;
    bool sing = false;
    const int nlev = l.n_lev;
    for (int v = 0; v <= nlev; v++) {
        if (v < nlev) {
            const int p1 = tb.piv_ptr[v + 1];
            for (int q = tb.piv_ptr[v] + tl; q < p1; q += 32) {
                const int2 pk = tb.piv[q];
                const double dgl = sLU[pk.y];
                if (!(fabs(dgl) >= DBL_MIN) || !isfinite(dgl)) sing = true;
                sDI[pk.x] = 1.0 / dgl;
            }
        }
        if (v > 0) {
            const int s1 = tb.sc_ptr[v];
            for (int q = tb.sc_ptr[v - 1] + tl; q < s1; q += 32) {
                const int2 e = tb.sc[q];
                const double lv = sLU[e.x] * sDI[e.y];
                sLU[e.x] = lv;
                w.weak |= weak_pivot(lv);
            }
        }
        __syncwarp();
        if (v < nlev) {
            const int t1 = tb.tg_ptr[v + 1];
            for (int q = tb.tg_ptr[v] + tl; q < t1; q += 32) {
                const int4 t = tb.tg[q];
                double a = sLU[t.x];
                for (int u = t.y; u < t.z; u++) {
                    const int4 up = tb.upd[u];
                    const double lv = sLU[up.x] * sDI[up.z];
                    a = a - lv * sLU[up.y];
                }
                sLU[t.x] = a;
            }
            __syncwarp();
        }
    }
    // forward: z[i] = F[rowperm[i]] - sum_k L[i][k] z[k]
    for (int v = 0; v < l.n_fwd; v++) {
        const int r1 = tb.flev_ptr[v + 1];
        for (int q = tb.flev_ptr[v] + tl; q < r1; q += 32) {
            const int4 r = tb.frow[q];
            double acc = sF[r.y];
            for (int e = r.z; e < r.w; e++) {
                const int2 en = tb.fent[e];
                acc = acc - sLU[en.x] * sWV[en.y];
            }
            sWV[r.x] = acc;
        }
        __syncwarp();
    }
    // backward: y[k] = (z[k] - sum_j U[k][j] y[j]) / pivot; y in the F slots
    bool finite = true;
    for (int v = 0; v < l.n_bwd; v++) {
        const int r1 = tb.blev_ptr[v + 1];
        for (int q = tb.blev_ptr[v] + tl; q < r1; q += 32) {
            const int4 r = tb.brow[q];
            double acc = sWV[r.x];
            for (int e = r.y; e < r.z; e++) {
                const int2 en = tb.bent[e];
                acc -= sLU[en.x] * sF[en.y];
            }
            acc *= sDI[r.x];
            sF[r.x] = acc;
            finite &= isfinite(acc);
        }
        __syncwarp();
    }
    singular = __any_sync(CB_FULL, sing);
    return __all_sync(CB_FULL, finite) && !singular;
}

// u[colperm[k]] -= delta[k]   (delta in the F slots)
template <typename PG, typename LU, typename W>
__device__ __forceinline__ void w_apply_update(const PG &pg, const LU &lu, W &w, const Hot &hot)
{
    for (int k = wtid(); k < lu.n(); k += 32) {
        const int j = pg.off_u() + hot.tb->colperm[k];
        w(j) = w(j) - hot.F[k];
    }
    __syncwarp();
}

template <typename PG, typename W>
__device__ __forceinline__ void w_copy(const PG &pg, W &w, int dst, int src, int count)
{
    for (int i = wtid(); i < count; i += 32) w(dst + i) = w(src + i);
    __syncwarp();
}

template <typename PG, typename W>
__device__ __forceinline__ void w_load_lane_params(const PG &pg, W &w, const double *lanes, int64_t P,
                                                   int64_t lane)
{
    for (int c = wtid(); c < pg.n_lane_cols(); c += 32) w(pg.off_lp() + c) = lanes[(int64_t)c * P + lane];
}

// ---------------------------------------------------------------------------
// dc_body of lane_kernels.cuh (PCNR, solve.jl:599-698 / plain Newton, :542-578), one lane per warp
// ---------------------------------------------------------------------------
template <typename PG, typename LU, typename W>
__device__ __forceinline__ void w_dc_body(const PG &pg, const LU &lu, W &w, const Program &p,
                                          const SpecArgs &sp, const DcArgs &a, int64_t lane, const Hot &hot)
{
    const int tl = wtid();
    if (a.active != nullptr && !a.active[lane]) return;      // masked lane: nothing to commit
    const int n = pg.n();
    w_load_lane_params(pg, w, p.lanes, p.P, lane);
    bool cold = true;
    for (int i = tl; i < n; i += 32) {
        const double v = a.u[(int64_t)i * p.P + lane];
        w(pg.off_u() + i) = v;
        cold &= (v == 0.0);
    }
    cold = __all_sync(CB_FULL, cold);
    const double gshunt = a.gshunt_lane ? a.gshunt_lane[lane] : sp.gshunt;
    const double srcFact = a.srcfact_lane ? a.srcfact_lane[lane] : sp.srcFact;
    const int lim0 = n - pg.n_limits();
    const bool pcnr = (a.algorithm == 0);
    const double abstol2 = a.abstol * a.abstol;
    bool initjct = false;
    if (pcnr && cold) {                                       // solve.jl:622-627
        for (int k = tl; k < pg.n_limits(); k += 32) {
            const int r = pg.limit_init_ref(k);
            w(pg.off_u() + lim0 + k) = r >= 0 ? pg.uniform(r) : w(pg.off_lp() + ~r);
        }
        initjct = true;
    }
    __syncwarp();
    w_eval_all(pg, w, a.t, sp.mode, false);

    bool settling = false, conv = false;
    int status = CB200_LANE_MAXITER, solves = 0, iter = 0;
    while (true) {
        if (!settling) iter++;
        const int bound = pcnr ? a.maxiters : a.maxiters + 1;
        if (iter > bound) { status = CB200_LANE_MAXITER; break; }
        w_eval_nonlinear(pg, w, a.t, sp.mode, initjct);
        initjct = false;
        bool bad;
        const double nrm2 = w_assemble<false>(pg, lu, w, hot, 0.0, gshunt, srcFact, bad);
        if (bad) { status = CB200_LANE_NONFINITE; break; }
        if (nrm2 < abstol2) {
            if (!pcnr || settling) { conv = true; status = CB200_LANE_OK; break; }
            w_copy(pg, w, pg.off_u() + lim0, pg.off_limw(), pg.n_limits());   // settle, re-verify
            settling = true;
            continue;
        }
        settling = false;
        if (!pcnr && iter > a.maxiters) { status = CB200_LANE_MAXITER; break; }
        bool singular;
        const bool ok = w_factor_and_solve(pg, lu, w, hot, singular);
        if (!ok) { status = singular ? CB200_LANE_SINGULAR : CB200_LANE_NONFINITE; break; }
        w_apply_update(pg, lu, w, hot);
        solves++;
        if (pcnr) w_copy(pg, w, pg.off_u() + lim0, pg.off_limw(), pg.n_limits());   // CORRECT
    }
    __syncwarp();
    for (int i = tl; i < n; i += 32) a.u[(int64_t)i * p.P + lane] = w(pg.off_u() + i);
    const bool any_weak = __any_sync(CB_FULL, w.weak);
    if (tl == 0) {
        a.status[lane] = status;
        if (a.weak && any_weak) a.weak[lane] = 1;
        a.iters[lane] += solves;
        a.converged[lane] = conv ? 1 : 0;
    }
}

// dc_stepping_body of lane_kernels.cuh (gshunt stepping, then source stepping, on the device), one
// lane per warp: the controller is warp-uniform.
template <typename PG, typename LU, typename W>
__device__ __forceinline__ void w_dc_stepping_body(const PG &pg, const LU &lu, W &w, const Program &p,
                                                   const SpecArgs &sp, const DcArgs &a, int64_t lane,
                                                   const Hot &hot)
{
    const int tl = wtid();
    if (a.active != nullptr && !a.active[lane]) return;
    const int n = pg.n();
    w_load_lane_params(pg, w, p.lanes, p.P, lane);
    for (int i = tl; i < n; i += 32) { w(pg.off_u() + i) = 0.0; w(pg.off_h1() + i) = 0.0; }
    __syncwarp();
    w_eval_all(pg, w, a.t, sp.mode, false);
    StepCtl ctl;
    ctl.begin(sp.gshunt);
    if (a.algorithm == 3) ctl.start_source();
    const double abstol2 = a.abstol * a.abstol;
    int iter = 0, solves = 0;
    while (ctl.phase != 3) {
        w_eval_nonlinear(pg, w, a.t, sp.mode, false);
        bool bad;
        const double nrm2 = w_assemble<false>(pg, lu, w, hot, 0.0, ctl.gshunt(), ctl.srcfact(sp.srcFact), bad);
        int res = -1;
        if (bad) res = 0;
        else if (nrm2 < abstol2) res = 1;
        else if (iter == a.maxiters) res = 0;
        else {
            bool singular;
            if (!w_factor_and_solve(pg, lu, w, hot, singular)) res = 0;
            else { w_apply_update(pg, lu, w, hot); solves++; iter++; }
        }
        if (res < 0) continue;
        iter = 0;
        const int act_on = ctl.result(res == 1);
        if (act_on == 0) w_copy(pg, w, pg.off_h1(), pg.off_u(), n);
        else if (act_on == 1) w_copy(pg, w, pg.off_u(), pg.off_h1(), n);
        else if (act_on == 2) {
            for (int i = tl; i < n; i += 32) { w(pg.off_u() + i) = 0.0; w(pg.off_h1() + i) = 0.0; }
            __syncwarp();
        }
    }
    __syncwarp();
    for (int i = tl; i < n; i += 32) a.u[(int64_t)i * p.P + lane] = w(pg.off_u() + i);
    const bool any_weak = __any_sync(CB_FULL, w.weak);
    if (tl == 0) {
        a.status[lane] = ctl.conv ? CB200_LANE_OK : CB200_LANE_MAXITER;
        if (a.weak && any_weak) a.weak[lane] = 1;
        a.iters[lane] += solves;
        a.converged[lane] = ctl.conv ? 1 : 0;
    }
}

// One implicit step's Newton loop (shared by the fixed-step and adaptive bodies); returns the
// lane status of the step and counts the linear solves.
template <typename PG, typename LU, typename W>
__device__ __forceinline__ int w_newton_step(const PG &pg, const LU &lu, W &w, const Hot &hot,
                                             const SpecArgs &sp,
                                             double t, double gamma, double abstol2, int max_nl,
                                             int limit, int &solves, bool &fresh, int &evals)
{
    bool lim_on = false;
    int it0 = 0;
    for (int it = 0;; it++) {
        const bool have = fresh && !kNlTimeDep;            // stamps and entry sums are those of this u
        if (!have) { w_eval_nonlinear(pg, w, t, CB200_MODE_TRAN, false); evals++; }
        fresh = true;
        bool bad;
        const double nrm2 = w_assemble<true>(pg, lu, w, hot, gamma, sp.gshunt, sp.srcFact, bad, have);
        bool restart = false;
        if (bad) return CB200_LANE_NONFINITE;
        if (nrm2 < abstol2) return CB200_LANE_OK;
        if (it - it0 >= (lim_on ? 4 * max_nl : max_nl)) {
            if (limit && !lim_on) { lim_on = true; it0 = it + 1; restart = true; }
            else return CB200_LANE_MAXITER;
        }
        if (restart) {                                     // redo the step from u_n, limiting on
            w_copy(pg, w, pg.off_u(), pg.off_un(), pg.n());
            fresh = false;
            continue;
        }
        bool singular;
        const bool ok = w_factor_and_solve(pg, lu, w, hot, singular);
        if (!ok) return singular ? CB200_LANE_SINGULAR : CB200_LANE_NONFINITE;
        w_apply_update(pg, lu, w, hot);
        fresh = false;
        solves++;
        if (lim_on) {                                      // PCNR corrector, solve.jl:686-689
            const int lim0 = pg.n() - pg.n_limits();
            w_copy(pg, w, pg.off_u() + lim0, pg.off_limw(), pg.n_limits());
        }
    }
}

// ---------------------------------------------------------------------------
// tran_fixed_body of lane_kernels.cuh, one lane per warp
// ---------------------------------------------------------------------------
template <typename PG, typename LU, typename W>
__device__ __forceinline__ void w_tran_fixed_body(const PG &pg, const LU &lu, W &w, const Program &p,
                                                  const SpecArgs &sp, const TranArgs &a, int64_t lane,
                                                  const Hot &hot)
{
    const int tl = wtid();
    const int n = pg.n();
    const bool resume = a.k_begin > 1;
    w_load_lane_params(pg, w, p.lanes, p.P, lane);
    for (int i = tl; i < n; i += 32) {
        w(pg.off_u() + i) = a.u[(int64_t)i * p.P + lane];
        w(pg.off_dterm() + i) = resume ? a.hist[(int64_t)(n + i) * p.P + lane] : 0.0;
        w(pg.off_un() + i) = resume ? a.hist[(int64_t)i * p.P + lane] : 0.0;
    }
    __syncwarp();
    w_eval_all(pg, w, a.t0, CB200_MODE_TRAN, false);

    int status = a.status[lane], solves = 0, evals = 0;
    bool fresh = false;
    int64_t tp = a.tp_begin;
    if (!resume) {
        for (int q = tl; q < a.n_save; q += 32)
            a.out[((int64_t)q * a.T + tp) * p.P + lane] = w(pg.off_u() + __ldg(a.save_idx + q));
        tp++;
    }
    const double h = a.h;
    const double abstol2 = a.abstol * a.abstol;
    const int amethod = a.method;
    for (int64_t k = a.k_begin; k <= a.k_end; k++) {
        const double t = a.t0 + (double)k * h;
        const int method = (k == 1) ? CB200_METHOD_BE : amethod;
        const double gamma = method == CB200_METHOD_BE ? 1.0 / h
                           : method == CB200_METHOD_TRAP ? 2.0 / h : 3.0 / (2.0 * h);
        for (int i = tl; i < n; i += 32) {
            const double ui = w(pg.off_u() + i);
            if (method == CB200_METHOD_GEAR2) w(pg.off_dterm() + i) = -(ui - w(pg.off_un() + i)) / (2.0 * h);
            else if (method == CB200_METHOD_BE) w(pg.off_dterm() + i) = 0.0;
            w(pg.off_un() + i) = ui;
        }
        __syncwarp();
        w_eval_sources(pg, w, t, CB200_MODE_TRAN);
        const int st = w_newton_step(pg, lu, w, hot, sp, t, gamma, abstol2, a.max_nl, a.limit, solves, fresh, evals);
        if (st != CB200_LANE_OK && status == CB200_LANE_OK) status = st;
        if (st != CB200_LANE_OK) fresh = false;
        if (st == CB200_LANE_NONFINITE || st == CB200_LANE_SINGULAR)     // dead lane: hold last state
            w_copy(pg, w, pg.off_u(), pg.off_un(), n);
        if (amethod == CB200_METHOD_TRAP) {
            for (int i = tl; i < n; i += 32)
                w(pg.off_dterm() + i) = -(gamma * (w(pg.off_u() + i) - w(pg.off_un() + i)) + w(pg.off_dterm() + i));
            __syncwarp();
        }
        if (k % a.save_every == 0 || k == a.nsteps) {
            for (int q = tl; q < a.n_save; q += 32)
                a.out[((int64_t)q * a.T + tp) * p.P + lane] = w(pg.off_u() + __ldg(a.save_idx + q));
            tp++;
        }
    }
    for (int i = tl; i < n; i += 32) {
        a.u[(int64_t)i * p.P + lane] = w(pg.off_u() + i);
        if (a.hist != nullptr) {
            a.hist[(int64_t)i * p.P + lane] = w(pg.off_un() + i);
            a.hist[(int64_t)(n + i) * p.P + lane] = w(pg.off_dterm() + i);
        }
    }
    const bool any_weak = __any_sync(CB_FULL, w.weak);
    if (tl == 0) {
        a.status[lane] = status;
        if (a.weak && any_weak) a.weak[lane] = 1;
        a.iters[lane] += solves;
        if (a.evals != nullptr) a.evals[lane] += evals;
    }
}

// ---------------------------------------------------------------------------
// tran_adaptive_body of lane_kernels.cuh, one lane per warp.  The LTE norm is summed by every
// thread over all unknowns in index order (the scaled errors are staged in the wv slots).
// ---------------------------------------------------------------------------
template <typename PG, typename LU, typename W>
__device__ __forceinline__ void w_tran_adaptive_body(const PG &pg, const LU &lu, W &w, const Program &p,
                                                     const SpecArgs &sp, const AdaptArgs &a, int64_t lane,
                                                     const Hot &hot)
{
    const int tl = wtid();
    const int n = pg.n();
    w_load_lane_params(pg, w, p.lanes, p.P, lane);
    for (int i = tl; i < n; i += 32) {
        w(pg.off_u() + i) = a.u[(int64_t)i * p.P + lane];
        w(pg.off_dterm() + i) = 0.0;
        w(pg.off_un() + i) = 0.0;
        w(pg.off_h1() + i) = 0.0;
        w(pg.off_h2() + i) = 0.0;
    }
    __syncwarp();
    w_eval_all(pg, w, a.t0, CB200_MODE_TRAN, false);

    int status = a.status[lane], solves = 0, rej = 0, T = 0, evals = 0;
    bool fresh = false;
    if (tl == 0) a.out_t[(int64_t)T * p.P + lane] = a.t0;
    for (int q = tl; q < a.n_save; q += 32)
        a.out[((int64_t)q * a.max_points + T) * p.P + lane] = w(pg.off_u() + __ldg(a.save_idx + q));
    T++;
    const int amethod = a.method;
    const double abstol2 = a.abstol * a.abstol;
    double t = a.t0, h = a.h0, h1 = 0.0, h2 = 0.0;
    int nhist = 0, istop = 0;
    bool finished = !(t < a.t1);
    while (!finished) {
        // ---- choose the step
        while (istop < a.n_tstops && __ldg(a.tstops + istop) <= fma(4.440892098500626e-16, fabs(t), t)) istop++;   // 4 eps(t), either sign of t
        double tnext = istop < a.n_tstops ? __ldg(a.tstops + istop) : a.t1;
        if (tnext > a.t1) tnext = a.t1;
        double hh = h;
        bool hit = false;
        if (t + hh >= tnext - 1e-3 * hh) { hh = tnext - t; hit = true; }
        const double tn = hit ? tnext : t + hh;
        const bool be = (nhist == 0 || amethod == CB200_METHOD_BE);
        const double gamma = be ? 1.0 / hh : 2.0 / hh;
        for (int i = tl; i < n; i += 32) {
            w(pg.off_un() + i) = w(pg.off_u() + i);
            if (be) w(pg.off_dterm() + i) = 0.0;           // trap: dterm holds -du_n
        }
        __syncwarp();
        w_eval_sources(pg, w, tn, CB200_MODE_TRAN);
        const int st = w_newton_step(pg, lu, w, hot, sp, tn, gamma, abstol2, a.max_nl, a.limit, solves, fresh, evals);
        if (st != CB200_LANE_OK) {                            // Newton failed: shrink and retry
            w_copy(pg, w, pg.off_u(), pg.off_un(), n);
            fresh = false;
            rej++;
            h = hh / 4.0;
            if (h < a.dtmin) {
                if (status == CB200_LANE_OK) status = (st == CB200_LANE_MAXITER) ? CB200_LANE_DTMIN : st;
                finished = true;
            }
            continue;
        }
        // ---- local truncation error estimate
        double err = 0.0;
        int pord = 1;
        if (nhist >= 1) {
            if (be || nhist == 1) {
                const double r = hh / h1, c = hh / (2.0 * hh + h1);
                for (int i = tl; i < n; i += 32) {
                    const double ui = w(pg.off_u() + i), uni = w(pg.off_un() + i);
                    const double up = uni + r * (uni - w(pg.off_h1() + i));
                    const double tol = lte_atol(a, i) + a.reltol * fmax(fabs(ui), fabs(uni));
                    hot.wv[i] = c * (ui - up) / tol;
                }
            } else {
                pord = 2;
                const double ta = -(h1 + h2), tb = -h1, tc = 0.0, tx = hh;
                const double la = (tx - tb) * (tx - tc) / ((ta - tb) * (ta - tc));
                const double lb = (tx - ta) * (tx - tc) / ((tb - ta) * (tb - tc));
                const double lc = (tx - ta) * (tx - tb) / ((tc - ta) * (tc - tb));
                const double c = hh * hh / (hh * hh + 2.0 * (hh + h1) * (hh + h1 + h2));
                for (int i = tl; i < n; i += 32) {
                    const double ui = w(pg.off_u() + i), uni = w(pg.off_un() + i);
                    const double up = la * w(pg.off_h2() + i) + lb * w(pg.off_h1() + i) + lc * uni;
                    const double tol = lte_atol(a, i) + a.reltol * fmax(fabs(ui), fabs(uni));
                    hot.wv[i] = c * (ui - up) / tol;
                }
            }
            __syncwarp();
            double acc = 0.0;
            for (int i = 0; i < n; i++) {
                const double e = hot.wv[i];
                acc += e * e;
            }
            err = sqrt(acc / (double)n);
            __syncwarp();
        }
        if (err > 1.0) {                                      // reject
            w_copy(pg, w, pg.off_u(), pg.off_un(), n);
            fresh = false;
            rej++;
            double f = 0.9 * pow(err, -1.0 / (pord + 1));
            if (f < 0.2) f = 0.2;
            h = hh * f;
            if (h < a.dtmin) { if (status == CB200_LANE_OK) status = CB200_LANE_DTMIN; finished = true; }
            continue;
        }
        // ---- accept
        for (int i = tl; i < n; i += 32) {
            const double ui = w(pg.off_u() + i), uni = w(pg.off_un() + i);
            w(pg.off_dterm() + i) = -(gamma * (ui - uni) + w(pg.off_dterm() + i));   // -du_{n+1}
            w(pg.off_h2() + i) = w(pg.off_h1() + i);
            w(pg.off_h1() + i) = uni;
        }
        __syncwarp();
        h2 = h1; h1 = hh;
        t = tn;
        nhist++;
        if (T < a.max_points) {
            if (tl == 0) a.out_t[(int64_t)T * p.P + lane] = t;
            for (int q = tl; q < a.n_save; q += 32)
                a.out[((int64_t)q * a.max_points + T) * p.P + lane] = w(pg.off_u() + __ldg(a.save_idx + q));
        }
        T++;
        double f = err > 0.0 ? 0.9 * pow(err, -1.0 / (pord + 1)) : 2.0;
        if (f > 2.0) f = 2.0;
        if (f < 0.2) f = 0.2;
        h = hh * f;
        if (h > a.dtmax) h = a.dtmax;
        if (hit && tn < a.t1) nhist = 0;                      // restart after a breakpoint
        if (!(t < a.t1)) finished = true;
        else if (T >= a.max_points) { if (status == CB200_LANE_OK) status = CB200_LANE_MAXITER; finished = true; }
    }
    __syncwarp();
    for (int i = tl; i < n; i += 32) a.u[(int64_t)i * p.P + lane] = w(pg.off_u() + i);
    const bool any_weak = __any_sync(CB_FULL, w.weak);
    if (tl == 0) {
        a.status[lane] = status;
        if (a.weak && any_weak) a.weak[lane] = 1;
        a.iters[lane] += solves;
        a.rejected[lane] = rej;
        a.count[lane] = T < a.max_points ? T : a.max_points;
        if (a.evals != nullptr) a.evals[lane] += evals;
    }
}

}  // namespace cb200
