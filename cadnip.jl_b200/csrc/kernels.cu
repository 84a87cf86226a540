// kernels.cu -- sm_100a CUDA kernels of the batched MNA Newton / transient hot path.
//
// Mapping: one sweep lane per thread.  A lane's whole Newton state --
// iterate, history, stamp values, the sparse LU factor -- lives in a private column
// of shared memory, ws[slot * blockDim.x + threadIdx.x] (fp64, conflict-free: a warp
// touches 32 consecutive 8-byte words), or of a global [slot][lane] array when the
// circuit is too large for shared memory (coalesced: lane is the fastest index).
// Lanes never communicate, so the kernels contain no barriers and no atomics.  The
// "program" (device table, segment lists, LU schedule) is read-only, identical for
// every lane and fetched through the read-only path at warp-uniform addresses.
//
// The device-evaluation functions restate the stamp! methods of
// src/mna/devices.jl; they write each stamp VALUE to the slot the reference's
// DirectStampContext would have visited at that program position
// (src/mna/value_only.jl:395-478).  Assembly then sums every matrix entry's segment
// of stamp slots left to right, i.e. in program order: the same floating-point
// summation order as the reference's `nzval[map[pos]] += v`, without atomics.
#include <cfloat>
#include <cmath>

#include "../../include/cadnip_b200.h"
#include "kernels.h"

namespace cb200 {

// ---------------------------------------------------------------------------
// lane workspace accessor
// ---------------------------------------------------------------------------
template <typename IdxT>
struct LaneWs {
    double *ws;
    IdxT stride;
    __device__ __forceinline__ double &operator()(int slot) const { return ws[(IdxT)slot * stride]; }
};

template <typename L>
__device__ __forceinline__ double param(const Program &p, const L &w, const int *pr, int i)
{
    const int r = __ldg(pr + i);
    return r >= 0 ? __ldg(p.uniform + r) : w(p.off_lp + ~r);
}

template <typename L>
__device__ __forceinline__ double xval(const Program &p, const L &w, int idx)
{
    // V_k = node_k == 0 ? 0.0 : x[node_k]   (devices.jl:1373)
    return idx == 0 ? 0.0 : w(p.off_u + idx - 1);
}

// ---------------------------------------------------------------------------
// waveforms (src/mna/devices.jl:30-216)
// ---------------------------------------------------------------------------
__device__ __forceinline__ double d_mod(double x, double y)   // Julia mod, y > 0
{
    double r = fmod(x, y);
    if (r != 0.0 && ((r < 0.0) != (y < 0.0))) r += y;
    return r;
}

// Base.sind restated: exact reduction mod 360, octant folding.
__device__ __forceinline__ double d_sind(double x)
{
    const double d2r = 0.017453292519943295;
    if (isnan(x) || isinf(x)) return NAN;
    const double rx = copysign(fmod(x, 360.0), x);
    const double arx = fabs(rx);
    if (rx == 0.0) return rx;
    else if (arx < 45.0) return sin(rx * d2r);
    else if (arx <= 135.0) return copysign(cos((90.0 - arx) * d2r), rx);
    else if (arx == 180.0) return copysign(0.0, rx);
    else if (arx < 225.0) return sin(((180.0 - arx) * (rx < 0 ? -1.0 : 1.0)) * d2r);
    else if (arx <= 315.0) return -copysign(cos((270.0 - arx) * d2r), rx);
    else return sin((rx - copysign(360.0, rx)) * d2r);
}

// pulse_at_time  devices.jl:85-103
__device__ __forceinline__ double d_pulse(double v1, double v2, double td, double tr, double tf,
                                          double pw, double per, double t)
{
    if (t < td) return v1;
    const double phase = per > 0 ? d_mod(t - td, per) : (t - td);
    if (phase < tr) return tr > 0 ? v1 + (v2 - v1) * (phase / tr) : v2;
    else if (phase < tr + pw) return v2;
    else if (phase < tr + pw + tf) return tf > 0 ? v2 + (v1 - v2) * ((phase - tr - pw) / tf) : v1;
    else return v1;
}

// get_source_value (devices.jl:352-360) over PWLWave / PulseWave / SinWave.
// pr[0] = dc, pr[1..] = wave parameters.
template <typename L>
__device__ __forceinline__ double source_value(const Program &p, const L &w, int wave,
                                               const int *pr, int npar, double t, int mode)
{
    if (wave == CB200_WAVE_NONE) return param(p, w, pr, 0);
    if (mode == CB200_MODE_DCOP || mode == CB200_MODE_AC) return param(p, w, pr, 0);
    if (wave == CB200_WAVE_SIN) {          // SinWave  devices.jl:168-174
        const double vo = param(p, w, pr, 1), va = param(p, w, pr, 2), fr = param(p, w, pr, 3);
        const double td = param(p, w, pr, 4), th = param(p, w, pr, 5), ph = param(p, w, pr, 6);
        if (t < td) return vo + va * d_sind(ph);
        return vo + va * exp(-th * (t - td)) * d_sind(360 * fr * (t - td) + ph);
    }
    if (wave == CB200_WAVE_PULSE) {
        return d_pulse(param(p, w, pr, 1), param(p, w, pr, 2), param(p, w, pr, 3),
                       param(p, w, pr, 4), param(p, w, pr, 5), param(p, w, pr, 6),
                       param(p, w, pr, 7), t);
    }
    // PWL: pwl_at_time devices.jl:47-71 with find_t_in_ts :30-36
    const int np = (npar - 1) / 2;
    int i = 1;                              // 1-based searchsortedfirst
    while (i <= np && param(p, w, pr, 1 + 2 * (i - 1)) < t) i++;
    if (i <= np && param(p, w, pr, 1 + 2 * (i - 1)) == t) i++;
    if (i <= 1) return param(p, w, pr, 2);
    if (i > np) return param(p, w, pr, 2 + 2 * (np - 1));
    const double t0 = param(p, w, pr, 1 + 2 * (i - 2)), y0 = param(p, w, pr, 2 + 2 * (i - 2));
    const double t1 = param(p, w, pr, 1 + 2 * (i - 1)), y1 = param(p, w, pr, 2 + 2 * (i - 1));
    if (y0 == y1) return y1;
    if (t1 == t0) return (y0 + y1) / 2;
    const double slope = (y1 - y0) / (t1 - t0);
    return y0 + (t - t0) * slope;
}

// ---------------------------------------------------------------------------
// limiting primitives (src/mna/devices.jl:1169-1258, :1333-1345)
// ---------------------------------------------------------------------------
__device__ __forceinline__ double d_pnjlim(double vnew, double vold, double vt, double vcrit)
{
    if (vnew > vcrit && fabs(vnew - vold) > vt + vt) {
        if (vold > 0.0) {
            const double arg = (vnew - vold) / vt;
            if (arg > 0.0) return vold + vt * (2.0 + log(arg - 2.0));
            return vold - vt * (2.0 + log(2.0 - arg));
        }
        return vt * log(vnew / vt);
    } else if (vnew < 0.0) {
        const double arg = vold > 0.0 ? -vold - 1.0 : 2.0 * vold - 1.0;
        if (vnew < arg) return arg;
    }
    return vnew;
}

__device__ __forceinline__ void d_diode_iv(double Is, double nVt, double v, double &I0, double &Gd)
{
    const double xarg = v / nVt;
    if (xarg > 80.0) {
        const double e80 = exp(80.0);
        I0 = Is * (e80 * (1.0 + (xarg - 80.0)) - 1.0);
        Gd = Is / nVt * e80;
    } else {
        const double expterm = exp(xarg);
        I0 = Is * (expterm - 1.0);
        Gd = Is / nVt * expterm;
    }
}

// diode_junction_cap  devices.jl:1505-1516
__device__ __forceinline__ double d_junction_cap(double V, double Cj0, double Vj, double m)
{
    const double Vmax = 0.9 * Vj;
    if (V < Vmax) return Cj0 / pow(1 - V / Vj, m);
    const double C_at_max = Cj0 / pow(1 - Vmax / Vj, m);
    const double dC_dV = Cj0 * m / Vj / pow(1 - Vmax / Vj, m + 1);
    return C_at_max + dC_dV * (V - Vmax);
}

// ---------------------------------------------------------------------------
// device evaluation: one stamp! call.  PASS 0 writes every stamp (once per kernel),
// PASS 1 only those that depend on (x, t); positions advance identically in both.
// ---------------------------------------------------------------------------
#define ST_G(i, j, v) do { if ((i) != 0 && (j) != 0) { if (PASS == 0) w(g) = (v); g++; } } while (0)
#define DY_G(i, j, v) do { if ((i) != 0 && (j) != 0) { w(g) = (v); g++; } } while (0)
#define ST_C(i, j, v) do { if ((i) != 0 && (j) != 0) { if (PASS == 0) w(c) = (v); c++; } } while (0)
#define DY_C(i, j, v) do { if ((i) != 0 && (j) != 0) { w(c) = (v); c++; } } while (0)
#define DY_B(i, v)    do { if ((i) != 0) { w(b) = (v); b++; } } while (0)

template <int PASS, typename L>
__device__ __forceinline__ void eval_device(const Program &p, const L &w, int d, double t,
                                            int mode, bool initjct)
{
    const int kind = __ldg(p.dev_kind + d), flags = __ldg(p.dev_flags + d);
    const int *nd = p.dev_nodes + __ldg(p.dev_node_ptr + d);
    const int pbeg = __ldg(p.dev_param_ptr + d);
    const int *pr = p.dev_params + pbeg;
    const int npar = __ldg(p.dev_param_ptr + d + 1) - pbeg;
    int g = p.off_SG + __ldg(p.dev_gbase + d);
    int c = p.off_SC + __ldg(p.dev_cbase + d);
    int b = p.off_SB + __ldg(p.dev_bbase + d);
    (void)npar; (void)c; (void)b; (void)flags;

    switch (kind) {
    case CB200_DEV_RESISTOR: {                       // devices.jl:498-510
        const int pp = __ldg(nd), nn = __ldg(nd + 1);
        const double G = 1.0 / param(p, w, pr, 0);
        ST_G(pp, pp, G); ST_G(pp, nn, -G); ST_G(nn, pp, -G); ST_G(nn, nn, G);
    } break;
    case CB200_DEV_CAPACITOR: {                      // devices.jl:531-534
        const int pp = __ldg(nd), nn = __ldg(nd + 1);
        const double C = param(p, w, pr, 0);
        ST_C(pp, pp, C); ST_C(pp, nn, -C); ST_C(nn, pp, -C); ST_C(nn, nn, C);
    } break;
    case CB200_DEV_INDUCTOR: {                       // devices.jl:569-586
        const int pp = __ldg(nd), nn = __ldg(nd + 1), I = __ldg(nd + 2);
        ST_G(pp, I, 1.0); ST_G(nn, I, -1.0); ST_G(I, pp, 1.0); ST_G(I, nn, -1.0);
        ST_C(I, I, -param(p, w, pr, 0));
    } break;
    case CB200_DEV_VSOURCE: {                        // devices.jl:643-663
        const int pp = __ldg(nd), nn = __ldg(nd + 1), I = __ldg(nd + 2);
        ST_G(pp, I, 1.0); ST_G(nn, I, -1.0); ST_G(I, pp, 1.0); ST_G(I, nn, -1.0);
        DY_B(I, source_value(p, w, flags, pr, npar, t, mode));
    } break;
    case CB200_DEV_ISOURCE: {                        // devices.jl:719-737
        const int pp = __ldg(nd), nn = __ldg(nd + 1);
        const double i = source_value(p, w, flags, pr, npar, t, mode);
        DY_B(pp, i); DY_B(nn, -i);
    } break;
    case CB200_DEV_VCVS: {                           // devices.jl:760-775
        const int op = __ldg(nd), on = __ldg(nd + 1), ip = __ldg(nd + 2), in = __ldg(nd + 3);
        const int I = __ldg(nd + 4);
        const double A = param(p, w, pr, 0);
        ST_G(op, I, 1.0); ST_G(on, I, -1.0); ST_G(I, op, 1.0); ST_G(I, on, -1.0);
        ST_G(I, ip, -A); ST_G(I, in, A);
    } break;
    case CB200_DEV_VCCS: {                           // devices.jl:797-808
        const int op = __ldg(nd), on = __ldg(nd + 1), ip = __ldg(nd + 2), in = __ldg(nd + 3);
        const double gm = param(p, w, pr, 0);
        ST_G(op, ip, -gm); ST_G(op, in, gm); ST_G(on, ip, gm); ST_G(on, in, -gm);
    } break;
    case CB200_DEV_CCVS: {
        const double rm = param(p, w, pr, 0);
        if (flags == 0) {                            // devices.jl:824-849
            const int op = __ldg(nd), on = __ldg(nd + 1), ip = __ldg(nd + 2), in = __ldg(nd + 3);
            const int Iin = __ldg(nd + 4), Iout = __ldg(nd + 5);
            ST_G(ip, Iin, 1.0); ST_G(in, Iin, -1.0); ST_G(Iin, ip, 1.0); ST_G(Iin, in, -1.0);
            ST_G(op, Iout, 1.0); ST_G(on, Iout, -1.0); ST_G(Iout, op, 1.0); ST_G(Iout, on, -1.0);
            ST_G(Iout, Iin, -rm);
        } else {                                     // devices.jl:898-913
            const int op = __ldg(nd), on = __ldg(nd + 1), Iin = __ldg(nd + 2), Iout = __ldg(nd + 3);
            ST_G(op, Iout, 1.0); ST_G(on, Iout, -1.0); ST_G(Iout, op, 1.0); ST_G(Iout, on, -1.0);
            ST_G(Iout, Iin, -rm);
        }
    } break;
    case CB200_DEV_CCCS: {
        const double A = param(p, w, pr, 0);
        if (flags == 0) {                            // devices.jl:865-881
            const int op = __ldg(nd), on = __ldg(nd + 1), ip = __ldg(nd + 2), in = __ldg(nd + 3);
            const int Iin = __ldg(nd + 4);
            ST_G(ip, Iin, 1.0); ST_G(in, Iin, -1.0); ST_G(Iin, ip, 1.0); ST_G(Iin, in, -1.0);
            ST_G(op, Iin, -A); ST_G(on, Iin, A);
        } else {                                     // devices.jl:924-931
            const int op = __ldg(nd), on = __ldg(nd + 1), Iin = __ldg(nd + 2);
            ST_G(op, Iin, -A); ST_G(on, Iin, A);
        }
    } break;
    case CB200_DEV_DIODE: {                          // devices.jl:1370-1428
        const int pp = __ldg(nd), nn = __ldg(nd + 1);
        const double V0 = xval(p, w, pp) - xval(p, w, nn);
        const double Is = param(p, w, pr, 0), Vt = param(p, w, pr, 1), nf = param(p, w, pr, 2);
        const double nVt = nf * Vt;
        if (flags & 1) {
            // limit!  devices.jl:1209-1234
            const int lim = __ldg(nd + 2);
            const double vcrit = param(p, w, pr, 3);
            const double vold = w(p.off_u + lim - 1);
            const double wv = initjct ? (V0 - V0 + vcrit) : d_pnjlim(V0, vold, nVt, vcrit);
            w(p.off_limw + (lim - 1 - (p.n - p.n_limits))) = wv;     // record_limit_w!
            ST_G(lim, lim, 1.0); ST_G(lim, pp, -1.0); ST_G(lim, nn, 1.0);
            double I0, Gd;
            d_diode_iv(Is, nVt, wv, I0, Gd);
            // stamp_limited_companion!  devices.jl:1251-1258
            DY_G(pp, pp, Gd); DY_G(pp, nn, -Gd); DY_G(nn, pp, -Gd); DY_G(nn, nn, Gd);
            const double Ieq = I0 - Gd * wv;
            DY_B(pp, -Ieq); DY_B(nn, Ieq);
        } else {
            const double expterm = exp(V0 / nVt);
            const double I0 = Is * (expterm - 1.0);
            const double Gd = Is / nVt * expterm;
            const double Ieq = I0 - Gd * V0;
            DY_G(pp, pp, Gd); DY_G(pp, nn, -Gd); DY_G(nn, pp, -Gd); DY_G(nn, nn, Gd);
            DY_B(pp, -Ieq); DY_B(nn, Ieq);
        }
    } break;
    case CB200_DEV_DIODECAP: {                       // devices.jl:1558-1602
        const int pp = __ldg(nd), nn = __ldg(nd + 1);
        const double V0 = xval(p, w, pp) - xval(p, w, nn);
        const double Is = param(p, w, pr, 0), Vt = param(p, w, pr, 1), nf = param(p, w, pr, 2);
        const double nVt = nf * Vt;
        const double expterm = exp(V0 / nVt);
        const double I0 = Is * (expterm - 1.0);
        const double G = Is / nVt * expterm;
        const double Ieq = I0 - G * V0;
        DY_G(pp, pp, G); DY_G(pp, nn, -G); DY_G(nn, pp, -G); DY_G(nn, nn, G);
        DY_B(pp, -Ieq); DY_B(nn, Ieq);
        const double Cj = d_junction_cap(V0, param(p, w, pr, 3), param(p, w, pr, 4), param(p, w, pr, 5));
        DY_C(pp, pp, Cj); DY_C(pp, nn, -Cj); DY_C(nn, pp, -Cj); DY_C(nn, nn, Cj);
    } break;
    case CB200_DEV_SIMPLEMOS: {                      // devices.jl:1667-1749
        const int dd = __ldg(nd), gg = __ldg(nd + 1), ss = __ldg(nd + 2);
        const double Vd = xval(p, w, dd), Vg = xval(p, w, gg), Vs = xval(p, w, ss);
        const double Vgs = Vg - Vs, Vds = Vd - Vs;
        const double Vth = param(p, w, pr, 0), K = param(p, w, pr, 1), lambda = param(p, w, pr, 2);
        double Ids, gm, gds;
        if (Vgs <= Vth) { Ids = 0.0; gm = 0.0; gds = 0.0; }
        else if (Vds <= Vgs - Vth) {
            Ids = K * ((Vgs - Vth) * Vds - Vds * Vds / 2);
            gm = K * Vds;
            gds = K * (Vgs - Vth - Vds);
        } else {
            Ids = K / 2 * ((Vgs - Vth) * (Vgs - Vth)) * (1 + lambda * Vds);
            gm = K * (Vgs - Vth) * (1 + lambda * Vds);
            gds = K / 2 * ((Vgs - Vth) * (Vgs - Vth)) * lambda;
        }
        const double Ieq = Ids - gm * Vgs - gds * Vds;
        DY_G(dd, dd, gds); DY_G(dd, gg, gm); DY_G(dd, ss, -(gds + gm));
        DY_G(ss, dd, -gds); DY_G(ss, gg, -gm); DY_G(ss, ss, gds + gm);
        DY_B(dd, -Ieq); DY_B(ss, Ieq);
        const double Cgd = param(p, w, pr, 3), Cgs = param(p, w, pr, 4);
        ST_C(gg, gg, Cgs); ST_C(gg, ss, -Cgs); ST_C(ss, gg, -Cgs); ST_C(ss, ss, Cgs);
        ST_C(gg, gg, Cgd); ST_C(gg, dd, -Cgd); ST_C(dd, gg, -Cgd); ST_C(dd, dd, Cgd);
    } break;
    default: break;
    }
}

template <typename L>
__device__ __forceinline__ void eval_all_static(const Program &p, const L &w, double t, int mode)
{
    for (int d = 0; d < p.n_dev; d++) eval_device<0>(p, w, d, t, mode, false);
}

template <typename L>
__device__ __forceinline__ void eval_dynamic(const Program &p, const L &w, double t, int mode,
                                             bool initjct)
{
    for (int q = 0; q < p.n_dyn; q++) eval_device<1>(p, w, __ldg(p.dyn_list + q), t, mode, initjct);
}

// ---------------------------------------------------------------------------
// assembly (deterministic segmented reduction) fused with the residual
//   F = C*du + G*u - b      (fast_residual!, precompile.jl:546-557; DC: F = G*u - b)
// and the Jacobian scatter  J = G + gamma*C  (fast_jacobian!, :568-585) straight into
// the LU workspace.  Also applies srcFact / gshunt (fast_rebuild!, :517-534).
// Returns ||F||_2^2; *bad is set when F has a non-finite entry.
// ---------------------------------------------------------------------------
template <bool TRAN, typename L>
__device__ __forceinline__ double assemble(const Program &p, const LuProgram &lu, const L &w,
                                           double gamma, double gshunt, double srcFact, bool &bad)
{
    for (int r = 0; r < p.n; r++) w(p.off_F + r) = 0.0;
    for (int q = 0; q < lu.n_fill; q++) w(p.off_LU + __ldg(lu.fill_slots + q)) = 0.0;
    for (int j = 0; j < p.n; j++) {
        const double uj = w(p.off_u + j);
        double duj = 0.0;
        if (TRAN) duj = gamma * (uj - w(p.off_un + j)) + w(p.off_dterm + j);
        const int s1 = __ldg(p.colptr + j + 1);
        for (int s = __ldg(p.colptr + j); s < s1; s++) {
            double gsum = 0.0;
            const int g1 = __ldg(p.gseg_ptr + s + 1);
            for (int q = __ldg(p.gseg_ptr + s); q < g1; q++) gsum += w(p.off_SG + __ldg(p.gseg_idx + q));
            if (gshunt != 0.0 && __ldg(p.nz_is_node_diag + s)) gsum += gshunt;
            double jv = gsum;
            const int r = __ldg(p.rowval + s);
            double f = w(p.off_F + r);
            if (TRAN) {
                double csum = 0.0;
                const int c1 = __ldg(p.cseg_ptr + s + 1);
                for (int q = __ldg(p.cseg_ptr + s); q < c1; q++) csum += w(p.off_SC + __ldg(p.cseg_idx + q));
                f += csum * duj;
                jv += gamma * csum;
            }
            f += gsum * uj;
            w(p.off_F + r) = f;
            w(p.off_LU + __ldg(lu.jmap + s)) = jv;
        }
    }
    double nrm2 = 0.0;
    bad = false;
    for (int r = 0; r < p.n; r++) {
        double bsum = 0.0;
        const int b1 = __ldg(p.bseg_ptr + r + 1);
        for (int q = __ldg(p.bseg_ptr + r); q < b1; q++) bsum += w(p.off_SB + __ldg(p.bseg_idx + q));
        if (srcFact < 1.0) bsum *= srcFact;
        const double f = w(p.off_F + r) - bsum;
        w(p.off_F + r) = f;
        bad |= !isfinite(f);
        nrm2 += f * f;
    }
    return nrm2;
}

// ---------------------------------------------------------------------------
// numeric refactor on the fixed pattern with the host's static pivot order, then
// the triangular solves.  delta (in pivot coordinates) ends up in the wv slots.
// Returns false when a pivot vanished or the solution is non-finite.
// ---------------------------------------------------------------------------
template <typename L>
__device__ __forceinline__ bool factor_and_solve(const Program &p, const LuProgram &lu, const L &w,
                                                 bool &singular)
{
    const int n = lu.n;
    singular = false;
    for (int k = 0; k < n; k++) {
        const int ds = p.off_LU + __ldg(lu.diag_slot + k);
        const double dgl = w(ds);
        if (!(fabs(dgl) >= DBL_MIN) || !isfinite(dgl)) singular = true;
        const double inv = 1.0 / dgl;
        w(ds) = inv;
        const int l0 = __ldg(lu.Lptr + k), l1 = __ldg(lu.Lptr + k + 1);
        const int u0 = __ldg(lu.Uptr + k), u1 = __ldg(lu.Uptr + k + 1);
        int tq = __ldg(lu.tgt_ptr + k);
        for (int e = l0; e < l1; e++) {
            const int ls = p.off_LU + __ldg(lu.L_slot + e);
            const double l = w(ls) * inv;
            w(ls) = l;
            for (int q = u0; q < u1; q++, tq++) {
                const int ts = p.off_LU + __ldg(lu.tgt + tq);
                w(ts) = w(ts) - l * w(p.off_LU + __ldg(lu.U_slot + q));
            }
        }
    }
    // forward: z = L^-1 P F
    for (int k = 0; k < n; k++) w(p.off_wv + k) = w(p.off_F + __ldg(lu.rowperm + k));
    for (int k = 0; k < n; k++) {
        const double zk = w(p.off_wv + k);
        const int l1 = __ldg(lu.Lptr + k + 1);
        for (int e = __ldg(lu.Lptr + k); e < l1; e++) {
            const int i = p.off_wv + __ldg(lu.L_row + e);
            w(i) = w(i) - w(p.off_LU + __ldg(lu.L_slot + e)) * zk;
        }
    }
    // backward: y = U^-1 z
    bool finite = true;
    for (int k = n - 1; k >= 0; k--) {
        double acc = w(p.off_wv + k);
        const int u1 = __ldg(lu.Uptr + k + 1);
        for (int q = __ldg(lu.Uptr + k); q < u1; q++)
            acc -= w(p.off_LU + __ldg(lu.U_slot + q)) * w(p.off_wv + __ldg(lu.U_col + q));
        acc *= w(p.off_LU + __ldg(lu.diag_slot + k));
        w(p.off_wv + k) = acc;
        finite &= isfinite(acc);
    }
    return finite && !singular;
}

// u[colperm[k]] -= delta[k]
template <typename L>
__device__ __forceinline__ void apply_update(const Program &p, const LuProgram &lu, const L &w)
{
    for (int k = 0; k < lu.n; k++) {
        const int j = p.off_u + __ldg(lu.colperm + k);
        w(j) = w(j) - w(p.off_wv + k);
    }
}

template <typename L>
__device__ __forceinline__ void load_lane_params(const Program &p, const L &w, int64_t lane)
{
    for (int c = 0; c < p.n_lane_cols; c++) w(p.off_lp + c) = p.lanes[(int64_t)c * p.P + lane];
}

// ---------------------------------------------------------------------------
// lane <-> workspace plumbing shared by the fused kernels
// ---------------------------------------------------------------------------
template <bool SMEM> struct WsSel;
template <> struct WsSel<true> {
    typedef LaneWs<int> type;
    __device__ static __forceinline__ type make(double *smem, double *, int64_t, int64_t) {
        type w; w.ws = smem + threadIdx.x; w.stride = (int)blockDim.x; return w;
    }
};
template <> struct WsSel<false> {
    typedef LaneWs<long long> type;
    __device__ static __forceinline__ type make(double *, double *glob, int64_t lane, int64_t P) {
        type w; w.ws = glob + lane; w.stride = (long long)P; return w;
    }
};

extern __shared__ double cb200_smem[];

// ---------------------------------------------------------------------------
// DC kernel: _dc_pcnr_newton (solve.jl:599-698) and the plain-Newton restatement of
// _dc_newton_compiled (solve.jl:542-578), one lane per thread, masked lanes.
// ---------------------------------------------------------------------------
template <bool SMEM>
__global__ void __launch_bounds__(128) dc_kernel(Program p, LuProgram lu, SpecArgs sp, DcArgs a)
{
    const int64_t lane0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in_range = lane0 < p.P;
    const int64_t lane = in_range ? lane0 : p.P - 1;      // clamp: compute, never commit
    typename WsSel<SMEM>::type w = WsSel<SMEM>::make(cb200_smem, a.ws_global, lane, p.P);
    const bool act = in_range && (a.active == nullptr || a.active[lane]);

    load_lane_params(p, w, lane);
    bool cold = true;
    for (int i = 0; i < p.n; i++) {
        const double v = a.u[(int64_t)i * p.P + lane];
        w(p.off_u + i) = v;
        cold &= (v == 0.0);
    }
    const double gshunt = a.gshunt_lane ? a.gshunt_lane[lane] : sp.gshunt;
    const double srcFact = a.srcfact_lane ? a.srcfact_lane[lane] : sp.srcFact;
    const int lim0 = p.n - p.n_limits;
    const bool pcnr = (a.algorithm == 0);
    bool initjct = false;
    if (pcnr && cold) {                                   // solve.jl:622-627
        for (int k = 0; k < p.n_limits; k++) {
            const int r = __ldg(p.limit_init_ref + k);
            w(p.off_u + lim0 + k) = r >= 0 ? __ldg(p.uniform + r) : w(p.off_lp + ~r);
        }
        initjct = true;
    }
    eval_all_static(p, w, a.t, sp.mode);

    bool done = !act;
    bool settling = false;
    int status = CB200_LANE_MAXITER, solves = 0, iter = 0;
    bool conv = false;
    while (true) {
        if (__all_sync(0xffffffffu, done)) break;
        if (!settling) iter++;
        // loop bound: PCNR `for iter in 1:maxiters`; Newton allows maxiters solves then a final test
        const int bound = pcnr ? a.maxiters : a.maxiters + 1;
        if (!done && iter > bound) { done = true; status = CB200_LANE_MAXITER; }
        eval_dynamic(p, w, a.t, sp.mode, initjct);
        initjct = false;
        bool bad;
        const double nrm2 = assemble<false>(p, lu, w, 0.0, gshunt, srcFact, bad);
        if (done) continue;
        if (bad) { done = true; status = CB200_LANE_NONFINITE; continue; }
        if (sqrt(nrm2) < a.abstol) {
            if (!pcnr || settling) { done = true; conv = true; status = CB200_LANE_OK; continue; }
            // settle the limit slots and re-verify (solve.jl:640-663)
            for (int k = 0; k < p.n_limits; k++) w(p.off_u + lim0 + k) = w(p.off_limw + k);
            settling = true;
            continue;
        }
        settling = false;
        if (!pcnr && iter > a.maxiters) { done = true; status = CB200_LANE_MAXITER; continue; }
        bool singular;
        const bool ok = factor_and_solve(p, lu, w, singular);
        if (!ok) { done = true; status = singular ? CB200_LANE_SINGULAR : CB200_LANE_NONFINITE; continue; }
        apply_update(p, lu, w);
        solves++;
        if (pcnr)                                           // CORRECT  solve.jl:686-689
            for (int k = 0; k < p.n_limits; k++) w(p.off_u + lim0 + k) = w(p.off_limw + k);
    }
    if (act) {
        for (int i = 0; i < p.n; i++) a.u[(int64_t)i * p.P + lane] = w(p.off_u + i);
        a.status[lane] = status;
        a.iters[lane] += solves;
        a.converged[lane] = conv ? 1 : 0;
    }
}

// ---------------------------------------------------------------------------
// fixed-step transient kernel: the whole time loop of a lane on the device.
//   du = gamma*(u - u_n) + dterm
//   BE:    gamma = 1/h,      dterm = 0
//   trap:  gamma = 2/h,      dterm = -du_n                      (first step BE)
//   Gear2: gamma = 3/(2h),   dterm = -(u_n - u_{n-1})/(2h)      (first step BE)
// Newton per step from u = u_n: rebuild, F, stop when ||F||_2 < abstol, otherwise
// solve J delta = F and update; at most max_nl solves.  t_k = t0 + k*h.
// ---------------------------------------------------------------------------
template <bool SMEM>
__global__ void __launch_bounds__(128) tran_fixed_kernel(Program p, LuProgram lu, SpecArgs sp,
                                                         TranArgs a)
{
    const int64_t lane0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool act = lane0 < p.P;
    const int64_t lane = act ? lane0 : p.P - 1;
    typename WsSel<SMEM>::type w = WsSel<SMEM>::make(cb200_smem, a.ws_global, lane, p.P);

    load_lane_params(p, w, lane);
    for (int i = 0; i < p.n; i++) {
        w(p.off_u + i) = a.u[(int64_t)i * p.P + lane];
        w(p.off_dterm + i) = 0.0;
    }
    eval_all_static(p, w, a.t0, CB200_MODE_TRAN);

    int status = a.status[lane], solves = 0;     // keeps an InitialFailure from the DC init
    int64_t tp = 0;
    if (act) for (int q = 0; q < a.n_save; q++)
        a.out[((int64_t)q * a.T + tp) * p.P + lane] = w(p.off_u + __ldg(a.save_idx + q));
    tp++;
    const double h = a.h;
    for (int64_t k = 1; k <= a.nsteps; k++) {
        const double t = a.t0 + (double)k * h;
        const int method = (k == 1) ? CB200_METHOD_BE : a.method;
        const double gamma = method == CB200_METHOD_BE ? 1.0 / h
                           : method == CB200_METHOD_TRAP ? 2.0 / h : 3.0 / (2.0 * h);
        // history terms; un <- u
        for (int i = 0; i < p.n; i++) {
            const double ui = w(p.off_u + i);
            if (method == CB200_METHOD_GEAR2) w(p.off_dterm + i) = -(ui - w(p.off_un + i)) / (2.0 * h);
            else if (method == CB200_METHOD_BE) w(p.off_dterm + i) = 0.0;
            /* trap: dterm already holds -du_n */
            w(p.off_un + i) = ui;
        }
        bool done = false;
        int st = CB200_LANE_OK;
        for (int it = 0;; it++) {
            eval_dynamic(p, w, t, CB200_MODE_TRAN, false);
            bool bad;
            const double nrm2 = assemble<true>(p, lu, w, gamma, sp.gshunt, sp.srcFact, bad);
            if (!done) {
                if (bad) { done = true; st = CB200_LANE_NONFINITE; }
                else if (sqrt(nrm2) < a.abstol) { done = true; }
                else if (it >= a.max_nl) { done = true; st = CB200_LANE_MAXITER; }
            }
            if (__all_sync(0xffffffffu, done)) break;
            bool singular;
            const bool ok = factor_and_solve(p, lu, w, singular);
            if (!done) {
                if (!ok) { done = true; st = singular ? CB200_LANE_SINGULAR : CB200_LANE_NONFINITE; }
                else { apply_update(p, lu, w); solves++; }
            }
        }
        if (st != CB200_LANE_OK && status == CB200_LANE_OK) status = st;
        if (st == CB200_LANE_NONFINITE || st == CB200_LANE_SINGULAR)   // dead lane: hold last state
            for (int i = 0; i < p.n; i++) w(p.off_u + i) = w(p.off_un + i);
        if (a.method == CB200_METHOD_TRAP)                // dterm <- -du_{n+1}
            for (int i = 0; i < p.n; i++)
                w(p.off_dterm + i) = -(gamma * (w(p.off_u + i) - w(p.off_un + i)) + w(p.off_dterm + i));
        if (k % a.save_every == 0 || k == a.nsteps) {
            if (act) for (int q = 0; q < a.n_save; q++)
                a.out[((int64_t)q * a.T + tp) * p.P + lane] = w(p.off_u + __ldg(a.save_idx + q));
            tp++;
        }
    }
    if (act) {
        for (int i = 0; i < p.n; i++) a.u[(int64_t)i * p.P + lane] = w(p.off_u + i);
        a.status[lane] = status;
        a.iters[lane] += solves;
    }
}

// ---------------------------------------------------------------------------
// evaluation-only kernels: K1 device evaluation over (instance x lane), K2
// segmented assembly over (matrix entry x lane).  Same device functions, global
// [slot][lane] workspace.
// ---------------------------------------------------------------------------
__global__ void eval_load_params_kernel(Program p, double *ws)
{
    const int64_t lane = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= p.P) return;
    LaneWs<long long> w; w.ws = ws + lane; w.stride = p.P;
    load_lane_params(p, w, lane);
}

__global__ void eval_devices_kernel(Program p, SpecArgs sp, EvalArgs a)
{
    const int64_t lane = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int d = blockIdx.y;
    if (lane >= p.P) return;
    LaneWs<long long> w; w.ws = a.ws + lane; w.stride = p.P;
    eval_device<0>(p, w, d, a.t, sp.mode, a.initjct != 0);
}

__global__ void eval_assemble_kernel(Program p, SpecArgs sp, EvalArgs a)
{
    const int64_t lane = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y;                 // 0..nnz-1: matrix entries; nnz..nnz+n-1: b rows
    if (lane >= p.P) return;
    LaneWs<long long> w; w.ws = a.ws + lane; w.stride = p.P;
    if (s < p.nnz) {
        double gsum = 0.0, csum = 0.0;
        for (int q = p.gseg_ptr[s]; q < p.gseg_ptr[s + 1]; q++) gsum += w(p.off_SG + p.gseg_idx[q]);
        for (int q = p.cseg_ptr[s]; q < p.cseg_ptr[s + 1]; q++) csum += w(p.off_SC + p.cseg_idx[q]);
        if (sp.gshunt != 0.0 && p.nz_is_node_diag[s]) gsum += sp.gshunt;
        if (a.G_nz) a.G_nz[(int64_t)s * p.P + lane] = gsum;
        if (a.C_nz) a.C_nz[(int64_t)s * p.P + lane] = csum;
    } else {
        const int r = s - p.nnz;
        double bsum = 0.0;
        for (int q = p.bseg_ptr[r]; q < p.bseg_ptr[r + 1]; q++) bsum += w(p.off_SB + p.bseg_idx[q]);
        if (sp.srcFact < 1.0) bsum *= sp.srcFact;
        if (a.b) a.b[(int64_t)r * p.P + lane] = bsum;
        if (a.limw && r < p.n_limits) a.limw[(int64_t)r * p.P + lane] = w(p.off_limw + r);
    }
}

cudaError_t launch_eval(const Program &p, const SpecArgs &s, const EvalArgs &a, cudaStream_t st,
                        int64_t *launches)
{
    const int block = 128;
    const unsigned gx = (unsigned)((p.P + block - 1) / block);
    eval_load_params_kernel<<<gx, block, 0, st>>>(p, a.ws);
    if (p.n_dev > 0) eval_devices_kernel<<<dim3(gx, p.n_dev), block, 0, st>>>(p, s, a);
    if (p.nnz + p.n > 0) eval_assemble_kernel<<<dim3(gx, p.nnz + p.n), block, 0, st>>>(p, s, a);
    if (launches) *launches += 3;
    return cudaGetLastError();
}

int choose_block(int n_slots, size_t smem_limit, int preferred)
{
    for (int b = preferred; b >= 32; b /= 2)
        if ((size_t)n_slots * b * sizeof(double) <= smem_limit) return b;
    return 0;
}

template <typename K, typename... Args>
static cudaError_t launch_lane_kernel(K ksmem, K kglob, int n_slots, int64_t P, int block,
                                      size_t smem_limit, cudaStream_t st, Args... args)
{
    const int b = choose_block(n_slots, smem_limit, block);
    if (b > 0) {
        const size_t bytes = (size_t)n_slots * b * sizeof(double);
        cudaError_t e = cudaFuncSetAttribute(ksmem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return e;
        ksmem<<<(unsigned)((P + b - 1) / b), b, bytes, st>>>(args...);
    } else {
        const int gb = 64;
        kglob<<<(unsigned)((P + gb - 1) / gb), gb, 0, st>>>(args...);
    }
    return cudaGetLastError();
}

cudaError_t launch_dc(const Program &p, const LuProgram &lu, const SpecArgs &s, const DcArgs &a,
                      int block, size_t smem_limit, cudaStream_t st, int64_t *launches)
{
    if (launches) *launches += 1;
    return launch_lane_kernel(dc_kernel<true>, dc_kernel<false>, p.n_slots, p.P, block, smem_limit,
                              st, p, lu, s, a);
}

cudaError_t launch_tran_fixed(const Program &p, const LuProgram &lu, const SpecArgs &s,
                              const TranArgs &a, int block, size_t smem_limit, cudaStream_t st,
                              int64_t *launches)
{
    if (launches) *launches += 1;
    return launch_lane_kernel(tran_fixed_kernel<true>, tran_fixed_kernel<false>, p.n_slots, p.P,
                              block, smem_limit, st, p, lu, s, a);
}

}  // namespace cb200
