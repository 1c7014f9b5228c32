/* mpcgpu.h -- C ABI of the B200 batched MPC solve engine (libmpcgpu.so).
 *
 * Drop-in boundary for the ONE hot path of Juleszwanen/oscar_mpc_planner_mr_modification: the SQP-RTI
 * loop that MPCPlanner::Solver::solve() runs once per GuidanceConstraints homotopy.  Every entry
 * point names the reference interface it replaces (paths relative to the reference root).
 * Plain pointers and sizes only; no torch / CUDA types in the signatures (a stream is passed as
 * void*, NULL = the engine's own stream).  All functions return 0 on success, a negative
 * mpcgpu_status otherwise; they never throw and keep no hidden global state besides the registry
 * of compiled problem configurations.  Callers own every array they pass.
 */
#ifndef MPCGPU_H
#define MPCGPU_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct mpcgpu_engine mpcgpu_engine;

enum mpcgpu_status {
    MPCGPU_OK = 0,
    MPCGPU_ERR_ARG = -1,        /* bad argument (NULL, n > max_batch, unknown config) */
    MPCGPU_ERR_CUDA = -2,       /* CUDA runtime error; see mpcgpu_last_error() */
    MPCGPU_ERR_NO_DEVICE = -3   /* no usable CUDA device: there is NO CPU fallback */
};

/* Number of compiled problem configurations and their names ("c1_basic", "tmpc_shipped", ...).
 * Replaces: the one generated solver linked at build time
 *           (mpc_planner_solver/include/mpc_planner_solver/solver_interface.h:4-12). */
int mpcgpu_num_configs(void);
const char *mpcgpu_config_name(int i);

/* Create / destroy an engine for one configuration on one device.
 * Replaces: Solver_acados_create_capsule + Solver_acados_create_with_discretization and
 *           Solver_acados_free + Solver_acados_free_capsule
 *           (mpc_planner_solver/src/acados_solver_interface.cpp:17,33,51-65).
 * One engine serves up to max_batch problems per call; device buffers are owned by the engine. */
int mpcgpu_engine_create(const char *config_name, int device, int max_batch, mpcgpu_engine **out);
int mpcgpu_engine_destroy(mpcgpu_engine *e);

/* Problem dimensions. Replaces: SOLVER_N / SOLVER_NX / SOLVER_NU / SOLVER_NP / SOLVER_NH macros of
 * the generated acados_solver_Solver.h (acados_solver_interface.h:14,20-47) and solver_settings.yaml
 * (acados_solver_interface.cpp:13,20-24). */
int mpcgpu_desc_query(const mpcgpu_engine *e, int *N, int *nx, int *nu, int *npar, int *nh);
/* doubles per problem of the persistent solver memory blob (see mpcgpu_solve_batch) */
int mpcgpu_mem_doubles(const mpcgpu_engine *e);

/* Solve n independent problems; HOST arrays in, HOST arrays out (copies are inside the call).
 * Replaces, per problem: Solver::solve() = initializeOneIteration + num_iter x solveOneIteration +
 *           completeOneIteration (acados_solver_interface.cpp:86-204) preceded by loadWarmstart (:274-284).
 *   xinit   [n * nx]                 AcadosParameters::xinit            (acados_solver_interface.h:53)
 *   x0      [n * (nu+nx) * (N+1)]    AcadosParameters::x0, [u_k, x_k]   (:54)
 *   params  [n * N * npar]           AcadosParameters::all_parameters   (:56), stage N reuses N-1 (:128-134)
 *   num_iter[n] or NULL              SQP-RTI iterations per problem; NULL => num_iter_all for all
 *                                    (Solver::_num_iterations; the wall-clock timeout rule of :108-116
 *                                    is evaluated by the host shim, which passes the resulting count).
 *                                    A NEGATIVE count -k runs k iterations with the completion step deferred: no
 *                                    res_eq demotion and no reset of mem_inout on failure.  That is the stepwise
 *                                    interface (solveOneIteration :145-160 keeps multipliers and QP memory between
 *                                    iterations; the demotion :176-181 and the reset :187-191 happen once, in
 *                                    completeOneIteration -- the host shim applies them there).
 *   mem_inout [n * mem_doubles] or NULL   persistent capsule memory (NLP multipliers + QP warm start):
 *                                    [flag][pi (N+1)*nx][lam N*nc][t N*nc][v (N+1)*(nu+nx)];
 *                                    flag 0 = fresh capsule, 1 = multipliers only (after
 *                                    ocp_nlp_solver_reset_qp_memory, :70), 2 = multipliers + QP warm start.
 *                                    Zeroed on failure like Solver_acados_reset (:187-191).
 *   xtraj   [n * nx * (N+1)], utraj [n * nu * N]   AcadosOutput (:129-130)
 *   pobj[n] AcadosInfo::pobj; exit_code[n] return value of solve() (1 success, 0, 2, 3, 4: :197-203);
 *   qp_status[n] AcadosInfo::qp_status in the numbering Solver::explainExitFlag decodes (:409-420): 0 ok,
 *                2 max iterations, 3 minimal step, 4 NaN; res_eq[n] the value tested at :177;
 *   ipm_iters[n] or NULL: total interior-point iterations (diagnostic).
 * Host pipeline: batches of >= 8192 problems whose INPUT arrays are page-locked (mpcgpu_alloc_pinned, cudaHostAlloc,
 * cudaHostRegister) are solved by ONE persistent launch while the copy stream still delivers the inputs chunk by chunk (a gate
 * word behind the work counter tells the kernel how far they have arrived); results are written by the kernel straight into
 * OUTPUT arrays that are page-locked too, otherwise staged and copied once.  Pageable inputs take chunked launches (copies
 * from pageable memory would serialise the single launch).  Results are identical either way. */
int mpcgpu_solve_batch(mpcgpu_engine *e, int n, const double *xinit, const double *x0, const double *params,
                       const int *num_iter, int num_iter_all, double *mem_inout, double *xtraj, double *utraj,
                       double *pobj, int *exit_code, int *qp_status, double *res_eq, int *ipm_iters);

/* Same contract with DEVICE pointers (inputs already resident in HBM); asynchronous on `stream`
 * (a cudaStream_t passed as void*, NULL = the engine's stream).  Every launch takes its own work counter from a ring
 * of MPCGPU_MAX_INFLIGHT, so up to that many device calls of one engine may be in flight on different streams.
 * mpcgpu_sync() waits for the engine's own streams AND for the most recent device call on a caller's stream. */
#define MPCGPU_MAX_INFLIGHT 16
int mpcgpu_solve_batch_device(mpcgpu_engine *e, int n, const double *xinit, const double *x0, const double *params,
                              const int *num_iter, int num_iter_all, double *mem_inout, double *xtraj, double *utraj,
                              double *pobj, int *exit_code, int *qp_status, double *res_eq, int *ipm_iters,
                              void *stream);
int mpcgpu_sync(mpcgpu_engine *e);

/* Optional arguments of the homotopy-set entries (NULL = none of them).  Plain pointers and sizes; HOST arrays.
 *
 * Consistency-cost post-processing.  Replaces: calculateConsistencyCostForSolver and its two call sites
 *   (mpc_planner_modules/src/guidance_constraints.cpp:384-388,405-408,1025-1050): for a planner with
 *   has_consistency_enabled the objective compared by FindBestPlanner is
 *       pobj - consistency_weight * sum_{k=1}^{N-2} ((x_k - X_k)^2 + (y_k - Y_k)^2)
 *   with (x_k, y_k) the SOLVED trajectory (getOutput(k, "x"/"y")), (X_k, Y_k) = _interpolated_prev_trajectory[k] (the
 *   values setConsistencyParametersForPlanner loaded into prev_traj_x/y, :985-1023), NOT scaled by dt, accumulated in
 *   stage order, multiplied by the weight once at the end, subtracted BEFORE the 0.75 selection weight (obj_scale).
 *   It is evaluated on the device from the solver's output, state components ix / iy (model_map "x", "y").
 * Persistent capsule memory.  Replaces: the per-planner local_solver capsules that live across control cycles
 *   (guidance_constraints.cpp:17-25); `*solver = *_solver` (:323) resets only the QP memory
 *   (acados_solver_interface.cpp:67-77), so the set entries downgrade flag 2 -> 1 (multipliers survive) before solving.
 * Static halfspaces.  Replaces: module_data.static_obstacles in LinearizedConstraints::update
 *   (linearized_constraints.cpp:107-127): n_static rows (a1, a2, b) per set and stage k >= 1, written behind the
 *   obstacle rows of a guided planner (slot n_obs + h) and from slot h on for a non-guided planner (its obstacle list is
 *   empty_data_, :326-329); only used by the entries that build the halfspaces on the device.
 * Selected trajectory.  Replaces: the copy of the best planner's _output into the main solver
 *   (guidance_constraints.cpp:520-522).  With best_xtraj / best_utraj the per-planner xtraj / utraj arguments of the set
 *   entries may be NULL: only the decision record {best_idx, pobj, exit_code, qp_status, res_eq} per planner and ONE
 *   trajectory per set travel back to the host (north_star: "only the per-problem cost and feasibility flags are gathered"). */
typedef struct mpcgpu_set_options {
    double consistency_weight;                 /* CONFIG["weights"]["consistency"] */
    const double *prev_traj;                   /* [n_sets * N * 2] (X_k, Y_k), k = 0..N-1; NULL = no consistency term */
    const unsigned char *consistency_enabled;  /* [n] planner.has_consistency_enabled; NULL = enabled for every planner */
    int ix, iy;                                /* state indices of x and y inside xtraj (0, 1 for both unicycle models) */
    double *mem_inout;                         /* [n * mem_doubles] persistent capsule memory, or NULL (fresh capsules) */
    double *objective_out;                     /* [n] planner.result.objective after the post-processing, or NULL */
    double *consistency_cost_out;              /* [n] the subtracted term (0 where not enabled), or NULL */
    const double *static_halfspaces;           /* [n_sets * N * n_static * 3], or NULL */
    int n_static;
    double *best_xtraj;                        /* [n_sets * nx * (N+1)] trajectory of the selected planner (planner 0 of the */
    double *best_utraj;                        /* [n_sets * nu * N]     set if none succeeded), or NULL; see below           */
} mpcgpu_set_options;

/* Homotopy-SET entry: what GuidanceConstraints::optimize does for one robot --
 * `*solver = *_solver` for every planner (guidance_constraints.cpp:323: all planners start from the main
 * solver's parameter block), planner-specific parameters on top (guidance halfspaces
 * linearized_constraints.cpp:150-189, consistency reference), one solve() each (:369), the objective post-processing
 * (:373-420, incl. the consistency cost of the solved trajectory: `opt`) and FindBestPlanner
 * (:572-590) -- as ONE call for n_sets sets of `planners` planners.  The shared block travels once per
 * set: host->device bytes drop from P*N*npar to N*npar + P*N*nidx doubles per set.
 *   xinit_sets [n_sets*nx]; shared_params [n_sets*N*npar]; x0 [n*(nu+nx)*(N+1)], n = n_sets*planners,
 *   problem index = set*planners + planner; param_idx [nidx] flat parameter indices that differ per planner;
 *   planner_params [n*N*nidx] their values (stage-major); outputs as mpcgpu_solve_batch plus
 *   best_idx [n_sets] with the semantics of mpcgpu_select_best (obj_scale / obj_sub / disabled may be NULL);
 *   obj_sub is a term known BEFORE the solve -- the consistency cost is not: it goes through `opt`. */
int mpcgpu_solve_sets(mpcgpu_engine *e, int n_sets, int planners, const double *xinit_sets, const double *shared_params,
                      const double *x0, int nidx, const int *param_idx, const double *planner_params, const int *num_iter,
                      int num_iter_all, double *xtraj, double *utraj, double *pobj, int *exit_code, int *qp_status,
                      double *res_eq, const double *obj_scale, const double *obj_sub, const unsigned char *disabled,
                      int *best_idx, const mpcgpu_set_options *opt);

/* Guidance halfspaces built ON THE DEVICE.
 * Replaces: LinearizedConstraints::update + projectToSafety + setParameters for the topology constraints of
 *           GuidanceConstraints (mpc_planner_modules/src/linearized_constraints.cpp:43-189 with _use_guidance = true, one
 *           disc, radius 1e-3; called per planner at guidance_constraints.cpp:321-352), i.e. the host loop that fills
 *           3 * (max_obstacles + add_halfspaces) parameters per planner and stage right before solve().
 * Writes the parameter slots lin_base + 3 j + {0,1,2} = (a1, a2, b) of constraint j < lin_count for every stage of every
 * problem (problem = set * planners + planner):
 *   stage 0 and non-guided planners (guided[q] == 0: update(state, empty_data_)): dummies (1, 0, xinit_x + 100)
 *   (static halfspaces, if any: mpcgpu_set_options);
 *   guided planners, k >= 1: pos = warm start (x0) position of stage k, pushed out of the obstacles (3 sweeps of the
 *   Douglas-Rachford step, radius 1e-3 + robot_radius), a = (o - pos)/|o - pos|, b = a.o - (1e-3 + robot_radius) with
 *   o = obst_pred[set][k-1][j] (prediction.modes[0][k-1].position); slots j >= n_obs are dummies.
 * The Douglas-Rachford step itself lives in the un-vendored `ros_tools` package: restated (oracle/mpc_oracle.c), parity
 * unpinned for that sub-step.  DEVICE pointers; xinit_sets [n_sets*nx], x0 [n*(nu+nx)*(N+1)], obst_pred
 * [n_sets*N*n_obs*2], guided [n], params [n*N*npar] in/out; asynchronous on `stream`. */
int mpcgpu_guidance_halfspaces_device(mpcgpu_engine *e, int n_sets, int planners, const double *xinit_sets, const double *x0,
                                      const double *obst_pred, int n_obs, const unsigned char *guided, int lin_base,
                                      int lin_count, double robot_radius, double *params, void *stream);

/* mpcgpu_solve_sets with the guidance halfspaces built on the device (HOST arrays): the per-planner halfspace block is
 * no longer uploaded -- obstacle predictions travel once per set (N * n_obs * 2 doubles) and the warm starts are needed
 * anyway.  Other per-planner parameters (e.g. the consistency reference) still go through param_idx / planner_params
 * (nidx may be 0); they are applied before the halfspaces are written. */
int mpcgpu_solve_sets_guided(mpcgpu_engine *e, int n_sets, int planners, const double *xinit_sets, const double *shared_params,
                             const double *x0, int n_obs, const double *obst_pred, const unsigned char *guided, int lin_base,
                             int lin_count, double robot_radius, int nidx, const int *param_idx, const double *planner_params,
                             const int *num_iter, int num_iter_all, double *xtraj, double *utraj, double *pobj, int *exit_code,
                             int *qp_status, double *res_eq, const double *obj_scale, const double *obj_sub,
                             const unsigned char *disabled, int *best_idx, const mpcgpu_set_options *opt);

/* Struct-of-tables parameter path: what the caller knows once per control cycle, in the shape it knows it.
 * Replaces: the k-loop `for k < N: for module: setParameters(data, module_data, k)` (mpc_planner/src/planner.cpp:153-159) and
 *           the string-keyed / if-chain setters underneath it (acados_solver_interface.cpp:212-225,
 *           solver_generator/generate_cpp_files.py:235-254): N * npar doubles per planner written on the host and copied.
 * Here the per-set parameter block [N][npar] is BUILT ON THE DEVICE from
 *   invariant  [n_sets][n_invariant]     parameters with the same value at every stage (weights, the 5 spline segments
 *                                        contouring.cpp:96-126, disc radius / offsets), flat indices invariant_idx[n_invariant];
 *   stage      [n_sets][N][n_stage]      parameters that change per stage and are shared by the planners of a set (e.g. decomp
 *                                        halfspaces), flat indices stage_idx[n_stage]; may be empty;
 *   obstacles  [n_sets][N][M][ob_stride] prediction step i of obstacle j: (x, y) for ob_stride 2 (psi = 0, radius from
 *                                        obstacle_radius[n_sets][M]) or (x, y, psi, r) for ob_stride 4 -> ellipsoid slots
 *                                        (EllipsoidConstraints::setParameters, ellipsoid_constraints.cpp:34-90: stage k <- step k-1,
 *                                        dummies at k = 0; ell_base < 0: none) and, with guided != NULL, the guidance halfspaces
 *                                        (see mpcgpu_guidance_halfspaces_device);
 * every other parameter is zero unless a per-planner override (param_idx / planner_params) writes it.  The generated
 * mpc_planner_tables.h / tables.yaml of a configuration list the invariant indices and the slot layout.
 * For the benchmark configuration this is 2.4 KB host->device per solve instead of 15 KB (mpcgpu_solve_sets) or 44 KB (flat). */
typedef struct mpcgpu_param_tables {
    int n_invariant;
    const int *invariant_idx;
    const double *invariant;
    int n_stage;
    const int *stage_idx;
    const double *stage;
    int M, ob_stride;
    const double *obstacles;
    const double *obstacle_radius;
    int ell_base, ell_stride;
    const int *ell_offsets;                    /* [7]: x, y, psi, major, minor, chi, r inside an obstacle's block */
    const unsigned char *guided;               /* [n] or NULL: no halfspaces built on the device */
    int lin_base, lin_count;
    double robot_radius;
} mpcgpu_param_tables;
int mpcgpu_solve_sets_tables(mpcgpu_engine *e, int n_sets, int planners, const double *xinit_sets, const mpcgpu_param_tables *tables,
                             const double *x0, int nidx, const int *param_idx, const double *planner_params, const int *num_iter,
                             int num_iter_all, double *xtraj, double *utraj, double *pobj, int *exit_code, int *qp_status,
                             double *res_eq, const double *obj_scale, const double *obj_sub, const unsigned char *disabled,
                             int *best_idx, const mpcgpu_set_options *opt);

/* Pick the best planner of each homotopy set.
 * Replaces: the objective post-processing of GuidanceConstraints::optimize and FindBestPlanner
 *           (mpc_planner_modules/src/guidance_constraints.cpp:373-420,572-590):
 *   obj_i = (pobj_i - obj_sub_i) * obj_scale_i   (consistency-cost subtraction :384-388,405-408 and
 *           selection_weight_consistency_ :418-419; either array may be NULL),
 *   success_i = exit_code_i == 1, disabled planners skipped (:578-579), strict '<' in ascending index
 *   order starting from 1e10 (:575-589); best_idx[s] = index within the set or -1.
 * HOST arrays; set_offsets has n_sets+1 entries. */
int mpcgpu_select_best(mpcgpu_engine *e, int n_sets, const int *set_offsets, const double *pobj, const int *exit_code,
                       const double *obj_scale, const double *obj_sub, const unsigned char *disabled, int *best_idx);
/* DEVICE-pointer variant, asynchronous on `stream`. */
int mpcgpu_select_best_device(mpcgpu_engine *e, int n_sets, const int *set_offsets, const double *pobj,
                              const int *exit_code, const double *obj_scale, const double *obj_sub,
                              const unsigned char *disabled, int *best_idx, void *stream);

/* Diagnostic (tests): evaluate the emitted model device functions at n points (HOST arrays).
 * Replaces nothing at run time; it is the hook that pins the CasADi-generated functions' counterparts
 * (Solver_model / Solver_cost / Solver_constraints of the generated acados solver) to golden vectors.
 *   z [n*(nu+nx)], p [n*npar], pi [n*nx] (multipliers weighting d2Phi), mh [n*nh] (weights of d2h)
 *   out[n*D], D = mpcgpu_model_eval_doubles(): xn[nx] | W[nx*nz] | Hdyn[pk] | cost | g[nz] (dt-scaled) |
 *   Hcost[pk] (dt-scaled) | h[nh] | C[nh*nhs] | Hcon[pk];  pk = nz(nz+1)/2 packed lower triangle,
 *   nhs = number of variables h depends on (returned through *nhs). */
int mpcgpu_model_eval_doubles(const mpcgpu_engine *e, int *nhs);
int mpcgpu_model_eval(mpcgpu_engine *e, int n, const double *z, const double *p, const double *pi, const double *mh,
                      double *out);

/* Measurement helper (no reference counterpart): FP64 FMA peak of `device` in TFLOP/s from a register-
 * resident DFMA kernel (best of 5 after warm-up).  The FP64 roofline denominator of bench.py. */
int mpcgpu_measure_fp64_peak(int device, double *tflops);

/* Synthetic workload on the device (measurement infrastructure, no reference counterpart; SURVEY 8d: "counter-based Philox
 * keyed by problem index so host and device generate identical data").  Fills the reference's own input layouts (xinit, x0,
 * all_parameters: acados_solver_interface.h:51-91) for homotopy sets first_set .. first_set + n_sets - 1 directly in DEVICE
 * memory; a set's data depend on (seed, set index) only, so any shard of any rank reproduces its slice of the global batch.
 * Value conventions as the reference's modules (contouring.cpp:52-126 spline block, ellipsoid_constraints.cpp:34-90 stage k <-
 * prediction k-1 and dummies at k = 0, linearized_constraints.cpp:49-189 halfspaces from the warm-start positions,
 * data_preparation.cpp:64-81 constant-velocity predictions, acados_solver_interface.cpp:303-342 braking roll-out).
 * Host mirror with the same arithmetic: synthetic.make_batch_philox (numpy).  The layout names the parameter indices of the
 * configuration (-1: absent) and carries the constants, so that both sides use the same doubles. */
#define MPCGPU_SYNTH_MAX_OBST 16
typedef struct mpcgpu_synth_layout {
    int N, nx, nu, npar;
    int guided;                 /* 1: planners follow guidance polylines, the last planner of a set (planners > 1) brakes */
    int weights[9];             /* acceleration, angular_velocity, velocity, reference_velocity, contour, lag, terminal_angle,
                                   terminal_contouring, consistency_weight */
    int spline[5][9];           /* segment i: x a b c d, y a b c d, start */
    int ego_disc_radius, ego_disc_0_offset;
    int goal[3];                /* goal_weight, goal_x, goal_y */
    int prev_traj_x, prev_traj_y;
    int n_obst;                 /* ellipsoid obstacles */
    int obst[MPCGPU_SYNTH_MAX_OBST][7];   /* x, y, psi, r, major, minor, chi */
    int n_lin;                  /* guidance halfspaces */
    int lin[MPCGPU_SYNTH_MAX_OBST][3];    /* a1, a2, b */
    double dt, pi, vg, need, deceleration, robot_radius, obstacle_radius, lin_margin;
    double weight_values[9];
    double lateral[8];          /* lateral offset of guidance corridor h % 8 */
    double lat_profile[64];     /* sin^2(pi k / N), k <= N */
} mpcgpu_synth_layout;
/* d_obst_pred (nullable): [n_sets][N][n_obst][2] obstacle predictions, the input of mpcgpu_solve_sets_guided.  stream: a
 * cudaStream_t (NULL: default stream); the call returns after the launch. */
int mpcgpu_generate_synthetic_device(int device, const mpcgpu_synth_layout *layout, unsigned long long seed, long long first_set,
                                     int n_sets, int planners, double *d_xinit, double *d_x0, double *d_params,
                                     double *d_obst_pred, void *stream);
/* the same data into HOST arrays (temporary device buffers inside; synchronous) */
int mpcgpu_generate_synthetic(int device, const mpcgpu_synth_layout *layout, unsigned long long seed, long long first_set,
                              int n_sets, int planners, double *xinit, double *x0, double *params, double *obst_pred);

/* Pinned (page-locked) host memory for the HOST-array entry points: copies from pinned buffers overlap the solve kernels
 * of the chunked pipeline; pageable memory works but serialises.  No reference counterpart (plumbing). */
int mpcgpu_alloc_pinned(size_t bytes, void **out);
int mpcgpu_free_pinned(void *p);

/* Several GPUs of one node behind ONE handle (north_star: "shards naturally across the 8 GPUs with one stream per
 * GPU and no NCCL on the solve path; only the per-problem cost and feasibility flags are gathered").
 * Replaces: the OpenMP team over planners (guidance_constraints.cpp:304) + one ROS node per robot, scaled out.
 * Homotopy sets are partitioned BY SET into contiguous ranges, one per device, the remainder to the last device
 * (mpcgpu_multi_shard_range; a set never straddles two devices because its argmin is taken on the device); every range is
 * solved by that device's engine from its own host thread on its own streams; results land in the caller's arrays at the
 * range's offsets -- the "gather" is the concatenation, no collective.  Arguments as the single-device entries. */
typedef struct mpcgpu_multi mpcgpu_multi;
int mpcgpu_multi_create(const char *config_name, const int *devices, int n_devices, int max_batch_per_device, mpcgpu_multi **out);
int mpcgpu_multi_destroy(mpcgpu_multi *m);
int mpcgpu_multi_num_devices(const mpcgpu_multi *m);
mpcgpu_engine *mpcgpu_multi_engine(mpcgpu_multi *m, int i);
int mpcgpu_multi_shard_range(int n_units, int n_devices, int i, int *begin, int *end);
int mpcgpu_multi_solve_sets(mpcgpu_multi *m, int n_sets, int planners, const double *xinit_sets, const double *shared_params,
                            const double *x0, int nidx, const int *param_idx, const double *planner_params, const int *num_iter,
                            int num_iter_all, double *xtraj, double *utraj, double *pobj, int *exit_code, int *qp_status,
                            double *res_eq, const double *obj_scale, const double *obj_sub, const unsigned char *disabled,
                            int *best_idx, const mpcgpu_set_options *opt);
int mpcgpu_multi_solve_sets_guided(mpcgpu_multi *m, int n_sets, int planners, const double *xinit_sets, const double *shared_params,
                                   const double *x0, int n_obs, const double *obst_pred, const unsigned char *guided, int lin_base,
                                   int lin_count, double robot_radius, int nidx, const int *param_idx, const double *planner_params,
                                   const int *num_iter, int num_iter_all, double *xtraj, double *utraj, double *pobj,
                                   int *exit_code, int *qp_status, double *res_eq, const double *obj_scale, const double *obj_sub,
                                   const unsigned char *disabled, int *best_idx, const mpcgpu_set_options *opt);
int mpcgpu_multi_solve_sets_tables(mpcgpu_multi *m, int n_sets, int planners, const double *xinit_sets, const mpcgpu_param_tables *tables,
                                   const double *x0, int nidx, const int *param_idx, const double *planner_params, const int *num_iter,
                                   int num_iter_all, double *xtraj, double *utraj, double *pobj, int *exit_code, int *qp_status,
                                   double *res_eq, const double *obj_scale, const double *obj_sub, const unsigned char *disabled,
                                   int *best_idx, const mpcgpu_set_options *opt);
/* independent problems (no sets): contiguous ranges of problems */
int mpcgpu_multi_solve_batch(mpcgpu_multi *m, int n, const double *xinit, const double *x0, const double *params,
                             const int *num_iter, int num_iter_all, double *mem_inout, double *xtraj, double *utraj,
                             double *pobj, int *exit_code, int *qp_status, double *res_eq, int *ipm_iters);
/* max over the devices of the solve-kernel time of the last multi call (ms, CUDA events on each device) */
float mpcgpu_multi_last_kernel_ms(mpcgpu_multi *m);

/* Diagnostic build (-DMPC_CHECK=1, lib/libmpcgpu_check.so; no reference counterpart -- it stands in for compute-sanitizer's
 * memcheck / racecheck, which this pool's GPUs refuse): shared-memory accessors with per-array bounds, canary words between the
 * shared-memory regions of both solve kernels, and labelled rendezvous in the role-split kernel (every role posts the protocol
 * point it arrived at; a mismatch is counted).  counters8: [0] index out of range, [1] canary overwritten, [2] roles met at
 * different barrier labels, [3] reserved, [4] problems checked.  Both return MPCGPU_ERR_ARG on a normal build. */
int mpcgpu_check_report(mpcgpu_engine *e, unsigned long long *counters8, int reset);
int mpcgpu_check_selftest(mpcgpu_engine *e);

/* Kernel choice (no reference counterpart).  Two kernels implement the same solve: the thread-per-stage kernel
 * (one warp per problem, 8 problems per SM: throughput) and the role-split kernel (one CTA of several warps
 * per problem: latency; compiled for configurations with enough general constraints).  AUTO takes the role-split
 * kernel for batches of at most SM-count problems (every problem has an SM to itself).  Returns 1 if the configuration has a role-split kernel,
 * 0 if not (SPLIT then falls back to the thread-per-stage kernel), negative on bad arguments. */
#define MPCGPU_KERNEL_AUTO 0
#define MPCGPU_KERNEL_STAGE 1
#define MPCGPU_KERNEL_SPLIT 2
int mpcgpu_set_kernel_mode(mpcgpu_engine *e, int mode);

/* Kernel launches issued by this engine so far (solve + select), for bench accounting. */
long long mpcgpu_launch_count(const mpcgpu_engine *e);
/* Device time in ms of the most recent mpcgpu_solve_batch[_device] solve kernel (CUDA events on the
 * launching stream); valid after the call returned (host variant) or after mpcgpu_sync(). */
float mpcgpu_last_kernel_ms(mpcgpu_engine *e);
const char *mpcgpu_last_error(const mpcgpu_engine *e);

#ifdef __cplusplus
}
#endif
#endif /* MPCGPU_H */
