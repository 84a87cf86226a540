/*
 * cadnip_oracle.h -- CPU restatement of Cadnip.jl's MNA Newton/transient hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity oracle for cadnip-b200; it is
 * never linked, imported or executed by the product (cadnip.jl_b200/).  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may use it.
 *
 * Parity pinning: the reference (NyanCAD/Cadnip.jl) is 100 % Julia and Julia is
 * not available in the build container, so the reference cannot be executed.
 * The oracle is pinned instead against every known-answer test the reference
 * holds for this path: see tests/test_oracle_golden.py.  The
 * third-party pieces (KLU pivot order, Sundials/OrdinaryDiffEq step control,
 * NonlinearSolve globalisation) are NOT in the reference tree and are unpinned:
 * "parity unpinned" for pivot sequences and adaptive time-point sequences.
 *
 * Each function cites the reference file:line it follows (paths relative to the
 * reference repository root).
 */
#ifndef CADNIP_ORACLE_H
#define CADNIP_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* device kinds / wave kinds: numerically identical to include/cadnip_b200.h so
 * the same netlist table can be fed to both sides (restated, not included).   */
enum {
    ORA_DEV_RESISTOR = 1, ORA_DEV_CAPACITOR = 2, ORA_DEV_INDUCTOR = 3,
    ORA_DEV_VSOURCE = 4, ORA_DEV_ISOURCE = 5, ORA_DEV_VCVS = 6, ORA_DEV_VCCS = 7,
    ORA_DEV_CCVS = 8, ORA_DEV_CCCS = 9, ORA_DEV_DIODE = 10, ORA_DEV_DIODECAP = 11,
    ORA_DEV_SIMPLEMOS = 12, ORA_DEV_VA = 13
};
enum { ORA_WAVE_NONE = 0, ORA_WAVE_PWL = 1, ORA_WAVE_PULSE = 2, ORA_WAVE_SIN = 3 };
enum { ORA_MODE_DCOP = 0, ORA_MODE_TRAN = 1, ORA_MODE_TRANOP = 2, ORA_MODE_AC = 3 };
enum { ORA_LANE_OK = 0, ORA_LANE_MAXITER = 1, ORA_LANE_SINGULAR = 2, ORA_LANE_NONFINITE = 3,
       ORA_LANE_DTMIN = 4 };
enum { ORA_METHOD_BE = 0, ORA_METHOD_TRAP = 1, ORA_METHOD_GEAR2 = 2, ORA_METHOD_BDF = 3 };

/* MNASpec, src/mna/solve.jl:57-70 */
typedef struct ora_spec {
    double temp; int32_t mode; int32_t _pad;
    double gmin, gshunt, srcFact, tnom, abstol, reltol, vntol, iabstol;
} ora_spec;

/* The "builder": an ordered list of stamp! calls on user nodes.  Node indices
 * are those get_node! returned (1-based, 0 = ground); current / limit variables
 * are allocated by running the list, exactly as the reference's builders do.
 * For kinds that reference an existing current variable (CCVS/CCCS flags=1) the
 * node entry holds the 1-based current-variable NUMBER k (CurrentIndex(k)).   */
typedef struct ora_netlist {
    int32_t n_dev;
    int32_t n_nodes;              /* nodes allocated by get_node! before stamping */
    const int32_t *kind;          /* [n_dev] */
    const int32_t *flags;         /* [n_dev] */
    const int32_t *node_ptr;      /* [n_dev+1] */
    const int32_t *nodes;
    const int32_t *par_ptr;       /* [n_dev+1] */
    const double  *par;           /* parameter VALUES of one circuit (one lane)   */
} ora_netlist;

typedef struct ora_ctx ora_ctx;           /* MNAContext           context.jl:248  */
typedef struct ora_structure ora_structure; /* CompiledStructure  precompile.jl:88 */
typedef struct ora_workspace ora_workspace; /* EvalWorkspace      precompile.jl:168 */

/* Verilog-A modules: registers the lookup `fn(model_index) -> stamp function` of a
 * shared object holding the C the product's emitter generated (see cadnip_oracle.c).  */
void ora_set_va_table(void *table_fn);

/* ---- pure functions (known-answer tests) -------------------------------- */
double ora_pwl_at_time(const double *ts, const double *ys, int n, double t);    /* devices.jl:47  */
double ora_pulse_at_time(double v1, double v2, double td, double tr, double tf,
                         double pw, double per, double t);                      /* devices.jl:85  */
double ora_sin_wave(double vo, double va, double freq, double td, double theta,
                    double phase, double t);                                     /* devices.jl:168 */
double ora_sind(double x_degrees);
double ora_pnjlim(double vnew, double vold, double vt, double vcrit, int *limited); /* devices.jl:1169 */
void   ora_diode_iv(double Is, double nVt, double v, double *I0, double *Gd);   /* devices.jl:1333 */
double ora_diode_vcrit(double Is, double Vt, double n);                         /* devices.jl:1319 */
/* sparse(I,J,V,n,n) + compute_coo_to_nz_mapping  (precompile.jl:253-283).
 * colptr[n+1], rowval/nzval sized >= ncoo, mapping[ncoo]; returns nnz.          */
int64_t ora_sparse_coo(int64_t n, int64_t ncoo, const int64_t *I, const int64_t *J,
                       const double *V, int64_t *colptr, int64_t *rowval, double *nzval,
                       int64_t *mapping);
/* expand_breakpoints for one source (solve.jl:1847-1918; devices.jl:145,180,211) */
int64_t ora_expand_breakpoints(int wave, const double *wp, int nwp, double t0, double t1,
                               double *out, int64_t cap);

/* ---- structure discovery and compilation -------------------------------- */
ora_ctx *ora_detect_structure(const ora_netlist *nl, const ora_spec *spec);     /* solve.jl:992-1015 */
void     ora_ctx_free(ora_ctx *c);
void     ora_ctx_sizes(const ora_ctx *c, int64_t *n_nodes, int64_t *n_currents,
                       int64_t *n_charges, int64_t *n_limits, int64_t *nG, int64_t *nC,
                       int64_t *nb);
/* resolved COO (1-based), values of the last stamping pass                     */
void     ora_ctx_coo(const ora_ctx *c, int64_t *G_I, int64_t *G_J, double *G_V,
                     int64_t *C_I, int64_t *C_J, double *C_V, int64_t *b_I, double *b_V,
                     double *limit_init);
/* run the builder once on a fresh MNAContext at state x (NULL = ZERO_VECTOR)    */
ora_ctx *ora_build(const ora_netlist *nl, const ora_spec *spec, double t, const double *x,
                   int64_t nx);

ora_structure *ora_compile_structure(const ora_netlist *nl, const ora_spec *spec,
                                     const ora_ctx *ctx);                        /* precompile.jl:312 */
void     ora_structure_free(ora_structure *s);
void     ora_structure_sizes(const ora_structure *s, int64_t *n, int64_t *nnz);
void     ora_structure_arrays(const ora_structure *s, int64_t *colptr, int64_t *rowval,
                              int64_t *G_coo_to_idx, int64_t *C_coo_to_idx,
                              int64_t *b_resolved, int64_t *G_diag_idx,
                              double *G_nz0, double *C_nz0);

ora_workspace *ora_create_workspace(const ora_structure *s);                    /* precompile.jl:193 */
void     ora_workspace_free(ora_workspace *w);

/* fast_rebuild! (precompile.jl:493-537).  Outputs copied if non-NULL.           */
void ora_fast_rebuild(ora_workspace *w, const ora_structure *s, const ora_spec *spec,
                      const double *u, double t, int initjct,
                      double *G_nz, double *C_nz, double *b, double *limit_w);

/* ---- DC ------------------------------------------------------------------ */
/* _dc_pcnr_newton (solve.jl:599-698). returns converged flag; *iters as the
 * reference reports it.                                                        */
int ora_dc_pcnr_newton(ora_workspace *w, const ora_structure *s, const ora_spec *spec,
                       double *u /*in: u0, out*/, double abstol, int maxiters, int *iters);
/* _dc_solve_with_fallbacks (solve.jl:871-929).  u in: u0, out: solution.
 * total_iters: Newton linear solves spent over all tiers.                      */
int ora_dc_solve_with_fallbacks(ora_workspace *w, const ora_structure *s, const ora_spec *spec,
                                double *u, double abstol, int maxiters, int use_stepping,
                                int *total_iters);
/* solve_dc(builder, params, spec) (solve.jl:2389-2420): detection + compile +
 * fallbacks.  x must hold n entries (query n with ora_ctx_sizes first, or pass
 * cap).  Returns converged flag, -1 on error.                                  */
int ora_solve_dc(const ora_netlist *nl, const ora_spec *spec, const double *u0, int64_t n_u0,
                 double abstol, int maxiters, double *x, int64_t cap, int64_t *n_out,
                 int *iters);

/* ---- transient, fixed step (BE / trapezoidal / Gear-2) -------------------- */
typedef struct ora_tran_opts {
    int32_t method, adaptive;
    double dt, abstol, reltol, lte_abstol, dtmin, dtmax;
    int32_t max_nl_iters, save_every, max_points, init;
    double init_abstol; int32_t init_maxiters; int32_t flags;   /* bit 0: PCNR corrector in transient */
    int32_t uic_steps, _pad; double uic_dt;                     /* init == 2: CedarUICOp warm-up (dcop.jl:311-411) */
    double vntol, iabstol, chgtol;                              /* flags bit 2: state_abstol per class (build.jl:276-283) */
} ora_tran_opts;

/* Linear solver behind every Newton solve: 0 = dense LU with partial pivoting (default: the
 * CHECKER of the parity tests), 1 = fixed-pattern sparse LU with a kept pivot sequence (what the
 * timed CPU baseline uses; KLU's role at solve.jl:612-613, :667-670).  Process-wide.           */
void ora_set_detect_override(const int32_t *flags, int64_t n);   /* replay a detection outcome (test hook) */
int ora_last_dc_tier(void);      /* tier that produced the calling thread's last DC result (-1: none) */
void ora_set_linear_solver(int kind);
int ora_get_linear_solver(void);

/* One circuit.  Fixed step: nsteps = round((t1-t0)/dt); saved points are k = 0,
 * save_every, 2*save_every, ... (the final step is always saved); out_u is
 * [T][n_save] row-major, out_t [T].  Adaptive: up to max_points points, *T_out
 * receives the count.  Returns lane status.                                     */
int ora_tran(const ora_netlist *nl, const ora_spec *spec, double t0, double t1,
             const ora_tran_opts *o, const int64_t *save_idx, int n_save,
             const double *u0, double *out_t, double *out_u, int64_t cap_T,
             int64_t *T_out, int64_t *newton_iters, int64_t *rejected, double *x_final);

/* ---- sweeps: the serial per-point loop of src/sweeps.jl:511-532 / :692-707,
 * parallelised over points with OpenMP (nthreads <= 0: all cores).            */
/* par_lanes: [P][npar] parameter values (row per lane). x_out [P][n].           */
int ora_sweep_dc(const ora_netlist *nl, const double *par_lanes, int64_t P, int64_t npar,
                 const ora_spec *spec, double abstol, int maxiters, int continuation,
                 double *x_out, int64_t n, int32_t *status, int32_t *iters, int nthreads);
/* out_u: [P][T][n_save]                                                        */
int ora_sweep_tran(const ora_netlist *nl, const double *par_lanes, int64_t P, int64_t npar,
                   const ora_spec *spec, double t0, double t1, const ora_tran_opts *o,
                   const int64_t *save_idx, int n_save, double *out_t, double *out_u,
                   int64_t cap_T, int32_t *T_out, int32_t *status, int64_t *newton_iters,
                   int nthreads);
int ora_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
