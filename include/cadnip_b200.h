/*
 * cadnip_b200.h -- C ABI of libcadnip_b200.so: the B200-native batched MNA
 * Newton / transient hot path behind Cadnip's MNACircuit / dc! / tran! /
 * CircuitSweep API.
 *
 * Plain pointers and sizes only; no torch / CUDA types cross this boundary.
 * All entry points return 0 on success or a negative CB200_E* code; they never
 * throw.  Per-lane numerical outcomes are reported in status[] arrays
 * (CB200_LANE_*), mirroring the reference's "numerical failure never throws"
 * convention (src/mna/solve.jl:1052, :2406; src/mna/dcop.jl:197-200).
 *
 * A "lane" is one sweep point of a CircuitSweep (src/sweeps.jl:387-424); lane
 * order is the sweep iterator's order (Iterators.product: first axis fastest,
 * src/sweeps.jl:272).  Unknown ordering is [nodes | currents | charges | limits]
 * (src/mna/context.jl:436-438).  Indices that are Julia `Int` in the reference
 * are int64 and 1-based here (0 = ground), so a Julia host can pass its arrays
 * through `ccall` without conversion.
 *
 * Reference interfaces each entry point replaces are cited per declaration
 * (paths relative to the reference repository root).
 */
#ifndef CADNIP_B200_H
#define CADNIP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CB200_ABI_VERSION 2

/* ---- error codes ------------------------------------------------------- */
#define CB200_OK            0
#define CB200_EINVAL       -1   /* bad argument / inconsistent description     */
#define CB200_ENOMEM       -2
#define CB200_ECUDA        -3   /* CUDA runtime error; see cb200_last_error    */
#define CB200_ENODEVICE    -4   /* no sm_100 device: there is NO CPU fallback  */
#define CB200_ESTATE       -5   /* call order violated (e.g. lanes not set)    */
#define CB200_ESINGULAR    -6   /* structurally singular pattern at analysis   */

/* ---- per-lane status ---------------------------------------- */
#define CB200_LANE_OK        0
#define CB200_LANE_MAXITER   1
#define CB200_LANE_SINGULAR  2
#define CB200_LANE_NONFINITE 3
#define CB200_LANE_DTMIN     4

/* ---- analysis mode: MNASpec.mode (src/mna/solve.jl:57-70) ---------------- */
#define CB200_MODE_DCOP   0
#define CB200_MODE_TRAN   1
#define CB200_MODE_TRANOP 2
#define CB200_MODE_AC     3

/* ---- device kinds: one per stamp! method of src/mna/devices.jl ---------- */
/* nodes[] holds resolved unknown indices (1-based, 0 = ground) in the order
 * listed; indices allocated by the stamp itself (alloc_current!/alloc_limit!)
 * follow the terminals.  params[] are parameter references (see below).      */
#define CB200_DEV_RESISTOR   1  /* nodes p,n        params r            devices.jl:498  */
#define CB200_DEV_CAPACITOR  2  /* nodes p,n        params c            devices.jl:531  */
#define CB200_DEV_INDUCTOR   3  /* nodes p,n,I      params l            devices.jl:569  */
#define CB200_DEV_VSOURCE    4  /* nodes p,n,I      params dc,wave...   devices.jl:643  */
#define CB200_DEV_ISOURCE    5  /* nodes p,n        params dc,wave...   devices.jl:719  */
#define CB200_DEV_VCVS       6  /* nodes op,on,ip,in,I      params gain devices.jl:760  */
#define CB200_DEV_VCCS       7  /* nodes op,on,ip,in        params gm   devices.jl:797  */
#define CB200_DEV_CCVS       8  /* nodes op,on,ip,in,Iin,Iout params rm devices.jl:824;
                                   flags=1: nodes op,on,Iin,Iout        devices.jl:898  */
#define CB200_DEV_CCCS       9  /* nodes op,on,ip,in,Iin    params gain devices.jl:865;
                                   flags=1: nodes op,on,Iin             devices.jl:924  */
#define CB200_DEV_DIODE     10  /* nodes p,n[,lim]  params Is,Vt,n,vcrit devices.jl:1370;
                                   flags bit0 = limit (PCNR)                            */
#define CB200_DEV_DIODECAP  11  /* nodes p,n  params Is,Vt,n,Cj0,Vj,m   devices.jl:1558 */
#define CB200_DEV_SIMPLEMOS 12  /* nodes d,g,s params Vth,K,lambda,Cgd,Cgs devices.jl:1667 */
#define CB200_DEV_VA        13  /* emitted Verilog-A module (cadnip_b200.verilog_a / src/vasim.jl:2993-3985):
                                   flags = model index in the header given to
                                   cb200_load_va_models; nodes ports, internal nodes,
                                   charge unknowns; params in declaration order          */
#define CB200_DEV_KIND_MAX  13

/* source waveform selector, stored in dev_flags of V/I sources
 * (PWLWave / PulseWave / SinWave, src/mna/devices.jl:130-216)                */
#define CB200_WAVE_NONE  0      /* params: dc                                   */
#define CB200_WAVE_PWL   1      /* params: dc, t1,y1, t2,y2, ...                */
#define CB200_WAVE_PULSE 2      /* params: dc, v1,v2,td,tr,tf,pw,per            */
#define CB200_WAVE_SIN   3      /* params: dc, vo,va,freq,td,theta,phase        */

/* A parameter reference is an int32: ref >= 0 selects uniform[ref] (same value
 * in every lane); ref < 0 selects lane-parameter column ~ref of the SoA given
 * to cb200_set_lanes (value differs per sweep point).                         */

/* MNASpec (src/mna/solve.jl:57-70) as a POD.                                  */
typedef struct cb200_spec {
    double  temp;      /* 27.0 */
    int32_t mode;      /* CB200_MODE_* */
    int32_t _pad;
    double  gmin;      /* 1e-12 */
    double  gshunt;    /* 0     */
    double  srcFact;   /* 1     */
    double  tnom;      /* 27    */
    double  abstol;    /* 1e-12 */
    double  reltol;    /* 1e-3  */
    double  vntol;     /* 1e-6  */
    double  iabstol;   /* 1e-12 */
} cb200_spec;

/* Circuit description: what build_with_detection + the builder's stamp! calls
 * produce on an MNAContext (src/mna/solve.jl:1793-1822, src/mna/context.jl:248-372),
 * with typed indices already resolved (context.jl:577-581), plus the device
 * table a recording context captures next to each stamp! call.    */
typedef struct cb200_desc {
    int32_t n_nodes, n_currents, n_charges, n_limits;
    int64_t nG, nC, nb;                 /* COO stamp counts (ground stamps excluded) */
    const int64_t *G_I, *G_J;           /* [nG] 1-based                        */
    const int64_t *C_I, *C_J;           /* [nC]                                */
    const int64_t *b_I;                 /* [nb]                                */
    int32_t n_devices;
    int32_t n_uniform;
    const int32_t *dev_kind;            /* [n_devices] CB200_DEV_*             */
    const int32_t *dev_flags;           /* [n_devices]                         */
    const int32_t *dev_node_ptr;        /* [n_devices+1] into dev_nodes        */
    const int32_t *dev_nodes;
    const int32_t *dev_param_ptr;       /* [n_devices+1] into dev_params       */
    const int32_t *dev_params;          /* parameter references                */
    const int64_t *dev_gbase;           /* [n_devices+1] first G COO slot (0-based) */
    const int64_t *dev_cbase;           /* [n_devices+1]                       */
    const int64_t *dev_bbase;           /* [n_devices+1]                       */
    const double  *uniform;             /* [n_uniform]                         */
    const int32_t *limit_init_ref;      /* [n_limits] parameter refs (context.jl:826) */
    int32_t n_lane_cols;                /* columns cb200_set_lanes must supply */
    int32_t _pad;
    const int32_t *dev_state_ptr;       /* [n_devices+1] or NULL: private per-lane state slots of each
                                           device (bias-independent set-up values of Verilog-A
                                           instances, computed once per kernel)            */
} cb200_desc;

typedef struct cb200_handle cb200_handle;
typedef struct cb200_wave   cb200_wave;

/* DC options: keyword arguments of _dc_solve_with_fallbacks
 * (src/mna/solve.jl:871-874) / CedarTranOp (src/mna/dcop.jl:81).             */
typedef struct cb200_dc_opts {
    double  abstol;        /* 1e-10 (dc!) / 1e-9 (transient init)              */
    int32_t maxiters;      /* 100 / 500                                        */
    int32_t use_stepping;  /* 1: gshunt + source stepping fallbacks; 2 / 3 (test hooks):
                              enter the chain at gshunt / source stepping directly    */
} cb200_dc_opts;

#define CB200_METHOD_BE    0
#define CB200_METHOD_TRAP  1
#define CB200_METHOD_GEAR2 2
#define CB200_METHOD_BDF   3   /* adaptive only: variable-order (1..5) variable-step BDF, the IDA family
                                  (src/sweeps.jl:599-601); see DESIGN.md s. 5 */

/* Transient options: the keyword surface of tran! that reaches the integrator
 * (src/sweeps.jl:588-665): solver choice (method), adaptive/dt, tolerances,
 * saveat decimation, initialisation algorithm.                               */
typedef struct cb200_tran_opts {
    int32_t method;        /* CB200_METHOD_*                                   */
    int32_t adaptive;      /* 0: fixed dt; 1: LTE step control per lane        */
    double  dt;            /* fixed step, or initial step when adaptive        */
    double  abstol;        /* Newton residual 2-norm tolerance (1e-10)         */
    double  reltol;        /* LTE relative tolerance when adaptive (1e-8)      */
    double  lte_abstol;    /* LTE absolute tolerance when adaptive             */
    double  dtmin, dtmax;  /* adaptive bounds (0 = derive from tspan)          */
    int32_t max_nl_iters;  /* 10 (IDA max_nonlinear_iters, sweeps.jl:599)      */
    int32_t save_every;    /* fixed-step: keep every k-th point (saveat), >=1  */
    int32_t max_points;    /* adaptive: per-lane output capacity               */
    int32_t init;          /* 0: CedarTranOp (dcop.jl:160); 1: u0 given, used as is;
                              2: CedarUICOp (dcop.jl:311-411): no DC solve, uic_steps fixed
                              backward-Euler steps of uic_dt from u0 (NULL = zeros) relax the
                              algebraic constraints, marching on through Newton failures   */
    double  init_abstol;   /* 1e-9  */
    int32_t init_maxiters; /* 500   */
    int32_t flags;         /* CB200_TRAN_* bits                                */
    int32_t uic_steps;     /* 10    (CedarUICOp warmup_steps)                  */
    int32_t _pad;
    double  uic_dt;        /* 1e-12 (CedarUICOp dt)                            */
    double  vntol;         /* CB200_TRAN_CLASS_ABSTOL: absolute tolerance of the LTE test  */
    double  iabstol;       /*   per class of unknown, state_abstol (build.jl:276-283):      */
    double  chgtol;        /*   node voltages + limit unknowns / branch currents / charges  */
} cb200_tran_opts;
/* Apply the PCNR corrector inside the transient Newton loop as well: after every solve the
 * limit unknowns are set to the recorded limited voltages w (solve.jl:686-689), so the
 * models' $limit functions (pnjlim / fetlim / limvds) damp the iteration as they do in
 * ngspice's transient.  Off = the reference's formulation (limit rows are ordinary
 * algebraic unknowns in transient).  The converged step is the same either way.        */
#define CB200_TRAN_LIMIT 1
/* adaptive mode: the LTE test uses vntol / iabstol / chgtol per class of unknown instead of the
 * one lte_abstol (the reference's `abstol = (vntol=..., iabstol=..., chgtol=...)`, sweeps.jl:556,
 * 615-618; defaults 1e-6 / 1e-12 / 1e-14) */
#define CB200_TRAN_CLASS_ABSTOL 4
/* cb200_specialize flags */
#define CB200_SPEC_COMPILE_ONLY 1   /* generate + compile into the cache, do not load          */
#define CB200_SPEC_FIXED_ONLY   4   /* skip the adaptive transient kernel (a third of the compile time) */
#define CB200_SPEC_TRAN_LIMIT   2   /* include the CB200_TRAN_LIMIT path in the transient kernels
                                       (without it, a transient with that flag runs on the
                                       table-driven kernels)                                     */

/* Counters of the most recent call (all device times from CUDA events on the
 * library's own stream).                                                      */
typedef struct cb200_stats {
    double  kernel_ms;         /* sum of kernel time of the last dc/tran call   */
    double  tran_kernel_ms;    /* the time-loop kernel alone (dominant kernel)  */
    double  dc_kernel_ms;      /* DC / initialisation kernels                    */
    double  h2d_ms, d2h_ms;    /* copies done inside the last call              */
    int64_t launches;          /* kernels launched by the last call             */
    int64_t newton_iters;      /* sum over lanes                                */
    int64_t steps_accepted;    /* sum over lanes                                */
    int64_t steps_rejected;    /* sum over lanes (adaptive)                     */
    int64_t h2d_bytes, d2h_bytes;
    int64_t device_evals;      /* transient: device-model evaluation passes executed, summed over
                                  lanes (a step whose first residual re-uses the stamps of the
                                  previous step's converged check does not count one)          */
    int64_t dc_stepping_lanes; /* lanes the last DC / initialisation sent through gshunt / source
                                  stepping (tiers 2-3 of solve.jl:871-929), on the device       */
} cb200_stats;

/* ---- lifecycle ---------------------------------------------------------- */

/* Replaces compile_structure + create_workspace (src/mna/precompile.jl:312-443,
 * :193): builds the unified G|C CSC pattern (sparse() semantics: rows sorted per
 * column, duplicates merged, explicit zeros kept), COO->nz maps, deferred-b
 * rows, diagonal map, and uploads the device table.  `device` is the CUDA
 * ordinal (one handle per GPU).  Copies everything; caller keeps ownership.   */
int cb200_create(const cb200_desc *desc, int32_t device, cb200_handle **out);
void cb200_destroy(cb200_handle *h);
const char *cb200_last_error(const cb200_handle *h);   /* h may be NULL       */
int cb200_abi_version(void);

/* Structure getters so a host that owns a CompiledStructure can check parity
 * (precompile.jl:253-283, :350-366, :451-467).  Arrays are int64 1-based.
 * Pass NULL pointers to query sizes only.                                     */
int cb200_get_pattern(const cb200_handle *h, int64_t *n, int64_t *nnz,
                      int64_t *colptr /*[n+1]*/, int64_t *rowval /*[nnz]*/);
int cb200_get_maps(const cb200_handle *h, int64_t *G_coo_to_idx /*[nG]*/,
                   int64_t *C_coo_to_idx /*[nC]*/, int64_t *b_rows /*[nb]*/,
                   int64_t *G_diag_idx /*[n_nodes]*/);

/* Replaces alter() per sweep point (src/mna/solve.jl:1719-1732; sweeps.jl:515,
 * :696): lane parameters as a struct of arrays soa[col*P + lane], host memory. */
int cb200_set_lanes(cb200_handle *h, int64_t P, int32_t n_cols, const double *soa);

/* Host symbolic phase of the linear solver that replaces KLU's analyze step
 * (call sites src/mna/solve.jl:612-613; sweeps.jl:600): probes |J| on sample
 * lanes, chooses a static pivot sequence (threshold Markowitz), computes fill
 * and the elimination schedule, uploads it.  gamma is the C-matrix weight the
 * probe uses (0 for DC, 1/dt for transient).  Called implicitly by dc/tran
 * when the schedule for that gamma class is missing.                          */
int cb200_analyze(cb200_handle *h, const cb200_spec *spec, double gamma);
/* perm arrays are int64 1-based: pivot k eliminates row rowperm[k], col colperm[k]. */
int cb200_get_pivot_order(const cb200_handle *h, int64_t *rowperm, int64_t *colperm,
                          int64_t *nnz_lu);

/* Verilog-A device models.  cuda_header is the text emitted for the circuit's modules
 * (va_stamp_<module> functions + va_dispatch): the kernels are rebuilt with it (nvcc,
 * cached under cache_dir by content hash) and the handle switches to them.  Required
 * before any evaluation when the description contains CB200_DEV_VA rows.  With
 * h == NULL the kernel set is only built into the cache (no device needed).           */
int cb200_load_va_models(cb200_handle *h, const char *cuda_header, const char *csrc_dir,
                         const char *cache_dir);

/* The emitter (north_star: "a new emitter lowers each device model's stamp function
 * ... to hand-written-style sm_100a CUDA C"): generates, compiles (nvcc, cached under
 * cache_dir by content hash) and loads kernels specialised for THIS circuit -- device
 * table, stamp segments, pattern and both LU schedules become compile-time
 * constants and the lane state lives in registers.  csrc_dir is the directory that
 * holds lane_kernels.cuh.  (method, dt) select the transient schedule and the integration
 * method compiled into the time-loop kernel.  flags bit0: compile only.  Without this call the table-driven kernels are used;
 * both instantiate the same device code.                                         */
int cb200_specialize(cb200_handle *h, const cb200_spec *spec, int32_t method, double dt,
                     const char *csrc_dir, const char *cache_dir, int32_t flags);
int cb200_is_specialized(const cb200_handle *h);
/* Static-pivot safeguard.  The pivot order is chosen once from sample lanes and reused for every
 * lane, step and gamma (klu_refactor's role, solve.jl:612-613); KLU re-pivots when a refactor meets
 * a bad pivot.  Here every refactor checks its multipliers: a lane on which one exceeded 1e8 is
 * MARKED (the solve carries on; a vanished pivot still ends the lane CB200_LANE_SINGULAR).
 * flags[P] (may be NULL) receives 1 for the marked lanes of the last cb200_dc / cb200_tran* call,
 * *count (may be NULL) their number.  The caller re-solves those lanes on a handle that holds only
 * them, so the order comes from their own values (analysis._retry_lanes does). */
int cb200_weak_pivot_lanes(cb200_handle *h, int32_t *flags, int64_t *count);
/* Which mapping the table-driven kernels use for this circuit (the specialised kernels, when
 * loaded, are always lane-per-thread / registers): 0 = one lane per thread, workspace in
 * shared memory; 1 = one lane per thread, workspace column in HBM; 2 = one lane per WARP,
 * workspace row in HBM / L2 (medium circuits: the lane state leaves < 64 lanes per block in
 * shared memory).  No reference counterpart: the reference solves one point at a time
 * (src/sweeps.jl:511-532).                                                          */
int cb200_lane_mapping(const cb200_handle *h);
/* Host-only test hook (no device): builds the pattern, the static-pivot schedule (pivoting on
 * |J_nz|) and the level schedule of the lane-per-warp kernels for a description, and executes
 * BOTH schedules on the host for the matrix J_nz (pattern order) and right-hand side rhs; the
 * two solutions must be identical.  J_nz == NULL: only the pattern (1-based) and info[0..1] =
 * {n, nnz} are returned.  info[2..5] = {nnz(LU), factor levels, forward levels, backward levels}. */
int cb200_host_lu_check(const cb200_desc *desc, const double *J_nz, const double *rhs,
                        double *x_serial, double *x_level, int64_t *colptr, int64_t *rowval,
                        int32_t *info);
/* The same with the pivot order taken from absJ / absJmin (the largest / smallest magnitude of every
 * entry over the probe states of cb200_analyze; absJmin may be NULL) and the numeric phase run on
 * J_nz, a matrix met later in the Newton loop: checks a STATIC order. */
int cb200_host_lu_check_static(const cb200_desc *d, const double *absJ, const double *absJmin, const double *J_nz,
                               const double *rhs, double *x_serial, double *x_level, int32_t *info);
/* Host-only emitter entry (no device): description + nominal |J| magnitudes for the DC
 * and transient schedules -> generated CUDA source (returns its length; negative =
 * error).  Lets the emitter be tested where no GPU is present.                      */
int64_t cb200_emit_source(const cb200_desc *desc, const double *absJ_dc, const double *absJ_tr,
                          int32_t method, int64_t P, int32_t num_sms, char *out, int64_t cap);

/* ---- evaluation only: fast_rebuild! (src/mna/precompile.jl:493-537) ------ */
/* x: [n][P] host (lane fastest) or NULL for ZERO_VECTOR; outputs host arrays
 * G_nz,C_nz: [nnz][P], b: [n][P]; any output may be NULL.  initjct arms the
 * PCNR seed evaluation (devices.jl:1217-1222).  limit_w_out: [n_limits][P].   */
int cb200_eval(cb200_handle *h, const cb200_spec *spec, double t, int32_t initjct,
               const double *x, double *G_nz, double *C_nz, double *b,
               double *limit_w_out);

/* ---- dc!(::CircuitSweep) (src/sweeps.jl:511-532; solve.jl:871-929) -------- */
/* u0: NULL (cold start, zeros) or [n][P]; x_out [n][P]; status/iters [P].     */
int cb200_dc(cb200_handle *h, const cb200_spec *spec, const cb200_dc_opts *opts,
             const double *u0, double *x_out, int32_t *status, int32_t *iters);

/* ---- tran!(::CircuitSweep, tspan) (src/sweeps.jl:692-707, :588-665) ------- */
/* save_idx: 1-based unknown indices to record.  The waveform stays in HBM
 * inside *out until fetched.  u0 [n][P] is used only when opts->init == 1.    */
int cb200_tran(cb200_handle *h, const cb200_spec *spec, double t0, double t1,
               const cb200_tran_opts *opts, const int64_t *save_idx, int32_t n_save,
               const double *u0, cb200_wave **out);

/* tran! with the result delivered straight into host memory (what a caller of
 * tran!(::CircuitSweep) ultimately reads).  Fixed-step runs are cut into n_segments
 * launches; the waveform of a finished segment is copied to u_out on a second stream
 * while the next segment computes, so only the last segment's copy is exposed.  u_out
 * should be page-locked for the copies to overlap.  Layouts as cb200_wave_fetch.       */
int cb200_tran_fetch(cb200_handle *h, const cb200_spec *spec, double t0, double t1,
                     const cb200_tran_opts *opts, const int64_t *save_idx, int32_t n_save,
                     const double *u0, int32_t n_segments, double *t_out, double *u_out,
                     int32_t *count, int32_t *status, int32_t *newton_iters);

/* Same, with a leading dimension: consecutive saved points of u_out (and, for adaptive runs, of
 * t_out) are u_ld >= P doubles apart, i.e. this handle's P lanes are the column block
 * [lane0, lane0 + P) of a waveform array of u_ld lanes that the caller addresses at
 * u_out = base + lane0.  This is how a sweep sharded over several GPUs (one handle each, contiguous
 * lane blocks in sweep order, src/sweeps.jl:272) delivers ONE [save][T][P_total] result --
 * the final gather of north_star -- without a staging copy.  u_ld = 0 means P.            */
int cb200_tran_fetch_ld(cb200_handle *h, const cb200_spec *spec, double t0, double t1,
                        const cb200_tran_opts *opts, const int64_t *save_idx, int32_t n_save,
                        const double *u0, int32_t n_segments, double *t_out, double *u_out,
                        int64_t u_ld, int32_t *count, int32_t *status, int32_t *newton_iters);

/* Adaptive stepping only: time points the integrator must hit exactly -- the source
 * breakpoints the host derives with expand_breakpoints (src/mna/solve.jl:1847-1918;
 * tran!'s auto_tstops, src/sweeps.jl:620-627).  Copied; persists until replaced.   */
int cb200_set_tstops(cb200_handle *h, const double *tstops, int32_t n);

/* Wave accessors.  Fixed-step layout is u[save][T][P] with a shared t[T];
 * adaptive layout is u[save][max_points][P] plus t[max_points][P] and count[P]. */
int cb200_wave_info(const cb200_wave *w, int64_t *T, int64_t *P, int32_t *n_save,
                    int32_t *adaptive);
int cb200_wave_fetch(cb200_wave *w, double *t /*[T] or [T][P]*/, double *u,
                     int32_t *count /*[P] or NULL*/, int32_t *status /*[P]*/,
                     int32_t *newton_iters /*[P]*/);
/* cb200_wave_fetch with leading dimensions for u (and the adaptive t): see cb200_tran_fetch_ld. */
int cb200_wave_fetch_ld(cb200_wave *w, double *t, double *u, int64_t u_ld, int64_t t_ld,
                        int32_t *count, int32_t *status, int32_t *newton_iters);
/* final state u(t1) of every lane, [n][P]                                      */
int cb200_wave_final_state(cb200_wave *w, double *x_out);
void cb200_wave_free(cb200_wave *w);

int cb200_get_stats(const cb200_handle *h, cb200_stats *out);

/* Measurement support (no reference counterpart).  cb200_measure_fp64_peak times a
 * register-only FP64 FMA kernel (8 independent chains per thread, 64 warps per SM) on `device`
 * and returns 2 x FMA/s in TFLOP/s: the denominator of the FP64 roofline of the fused,
 * register-resident kernels.  cb200_flop_model returns the static flop counts of one Newton
 * iteration's linear algebra for the handle's circuit: out[0..1] assembly + residual (DC,
 * transient), out[2..3] numeric refactor, out[4..5] triangular solves, out[6] update, out[7]
 * nonlinear device evaluations per iteration.                                             */
int cb200_measure_fp64_peak(int32_t device, double *tflops, double *kernel_ms);
int cb200_flop_model(const cb200_handle *h, int64_t out[8]);

/* Test hook: y[i] = the device kernels' junction exp(x[i]) (accuracy tests).          */
int cb200_debug_exp(const double *x, double *y, int32_t n);

#ifdef __cplusplus
}
#endif
#endif /* CADNIP_B200_H */
