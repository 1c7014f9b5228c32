/* mpcgpu_wire.h -- ROS 1 wire formats of mpc_planner_msgs <-> engine tables.
 *
 * The reference moves obstacle predictions and solver metrics as ROS 1 messages
 * (mpc_planner_msgs/msg/ObstacleGMM.msg, ObstacleArray.msg, Gaussian.msg, MPCMetrics.msg) and re-marshals them through
 * std::vector<DynamicObstacle> into string-keyed setParameter calls.  These entry points go from the serialized bytes
 * (roscpp serialization: little endian, uint32 length prefixes for strings and arrays, time = 2 x uint32) straight to
 * the tables the engine consumes, and from the engine's result records to a serialized MPCMetrics message.
 * Part of libmpcgpu.so; plain C ABI, no ROS dependency. */
#ifndef MPCGPU_WIRE_H
#define MPCGPU_WIRE_H
#include <stddef.h>
#include "mpcgpu.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct mpcgpu_track {
    int id;             /* ObstacleGMM.id */
    double x, y, angle; /* ObstacleGMM.pose: position and yaw of the orientation (RosTools::quaternionToAngle) */
    double radius;      /* not on the wire: CONFIG["obstacle_radius"] / the peer's robot radius; parsers set 0, caller fills */
    int n_steps;        /* poses of gaussians[0].mean (0: message without a trajectory, ignored by the reference) */
} mpcgpu_track;

/* Parse one serialized mpc_planner_msgs/ObstacleGMM.
 * Replaces: JulesJackalPlanner::trajectoryCallback, the part that turns the message into a DynamicObstacle
 *           (mpc_planner_jackalsimulator/src/jules_ros1_jackalplanner.cpp:566-616): pose -> position/angle, the FIRST
 *           Gaussian's mean path -> prediction.modes[0] (position, quaternion yaw; major = minor = -1 are not kept:
 *           the prediction is DETERMINISTIC).
 * steps [max_steps][3] = (x, y, angle) per pose; poses beyond max_steps are skipped (n_steps is clamped).
 * Returns the number of bytes consumed (> 0) or MPCGPU_ERR_ARG on a truncated / malformed buffer. */
long mpcgpu_wire_parse_obstacle_gmm(const unsigned char *buf, size_t len, mpcgpu_track *track, int max_steps, double *steps);

/* Parse one serialized mpc_planner_msgs/ObstacleArray (std_msgs/Header + ObstacleGMM[]); upstream mpc_planner's
 * obstacleCallback (a no-op in this fork, jules_ros1_jackalplanner.cpp:473-476).  tracks [max_tracks], steps
 * [max_tracks][max_steps][3]; obstacles beyond max_tracks are skipped.  Returns bytes consumed or MPCGPU_ERR_ARG. */
long mpcgpu_wire_parse_obstacle_array(const unsigned char *buf, size_t len, int max_tracks, int max_steps, mpcgpu_track *tracks,
                                      double *steps, int *n_tracks);

/* Obstacle table of one robot from its tracks.
 * Replaces: ensureObstacleSize (mpc_planner/src/data_preparation.cpp:97-170: more than max_obstacles -> keep the closest
 *           by min_k (k+1) 0.6 |pred_k - (pos + v k dir)|, stable order on ties; fewer -> dummies at (x+100, y+100),
 *           radius 0, constant prediction :51-58,159-166) and the per-stage reads of EllipsoidConstraints::setParameters.
 * table [N][max_obstacles][4] = (x, y, psi, r) of prediction step i; state = (x, y, psi, v).  Every kept track needs
 * n_steps >= N (the reference indexes modes[0][k] for k < N unchecked).  Returns the number of real (non-dummy) obstacles
 * or MPCGPU_ERR_ARG. */
int mpcgpu_obstacle_table(const mpcgpu_track *tracks, const double *steps, int n_tracks, int max_steps, int N, int max_obstacles,
                          const double *state_xypsiv, double *table);

/* Ellipsoid parameter slots from obstacle tables, on the device.
 * Replaces: EllipsoidConstraints::setParameters (mpc_planner_modules/src/ellipsoid_constraints.cpp:34-90) for
 *           DETERMINISTIC predictions: stage 0 dummies (x+50, y+50, psi 0, r 0.1, major = minor = 0, chi 1), stage k >= 1
 *           from prediction step k-1 (x, y, psi, r; major = minor = 0, chi = 1).
 * DEVICE pointers: xinit_sets [n_sets*nx], table [n_sets][N][M][4], params [n_sets][N][npar] (the per-set shared block)
 * in/out.  Obstacle j's block starts at ell_base + j*ell_stride; ell_offsets[7] = offsets of (x, y, psi, major, minor,
 * chi, r) inside it (parameter_map.yaml).  Asynchronous on `stream`. */
int mpcgpu_pack_obstacles_device(mpcgpu_engine *e, int n_sets, const double *xinit_sets, const double *table, int M, int ell_base,
                                 int ell_stride, const int *ell_offsets, double *params, void *stream);

/* mpcgpu_solve_sets_guided fed from obstacle tables (HOST arrays): the ellipsoid slots of the shared block AND the guidance
 * halfspaces are written on the device from table [n_sets][N][M][4]; shared_params only needs the remaining parameters
 * (weights, spline, radii).  Everything else as mpcgpu_solve_sets_guided. */
int mpcgpu_solve_sets_tracks(mpcgpu_engine *e, int n_sets, int planners, const double *xinit_sets, const double *shared_params,
                             const double *x0, int M, const double *table, const unsigned char *guided, int lin_base, int lin_count,
                             double robot_radius, int ell_base, int ell_stride, const int *ell_offsets, const int *num_iter,
                             int num_iter_all, double *xtraj, double *utraj, double *pobj, int *exit_code, int *qp_status,
                             double *res_eq, const double *obj_scale, const double *obj_sub, const unsigned char *disabled,
                             int *best_idx, const mpcgpu_set_options *opt);

/* Solver section of mpc_planner_msgs/MPCMetrics from the engine's records of one homotopy set.
 * Replaces: the metrics fill of the planner node (jules_ros1_jackalplanner.cpp, _metrics_pub) for the fields the solve path
 * owns: solve_time_ms, success_rate, iterations, exit_code, objective_value, objective_values_all_planners,
 * selected_planner_index, used_guidance, num_of_guidance_found.  All other fields are serialized empty / zero in message
 * order so that the bytes are a valid MPCMetrics message.  Returns bytes written or MPCGPU_ERR_ARG (buffer too small). */
typedef struct mpcgpu_metrics {
    unsigned int seq, stamp_sec, stamp_nsec;
    const char *frame_id, *robot_name;
    double solve_time_ms, success_rate;
    int iterations, exit_code;
    double objective_value;
    const double *objective_values_all_planners;
    int n_planners;
    int selected_planner_index, num_of_guidance_found;
    unsigned char used_guidance;
} mpcgpu_metrics;
long mpcgpu_wire_serialize_metrics(const mpcgpu_metrics *m, unsigned char *buf, size_t cap);
/* fills objective / exit code / selected index / guidance counts of `m` from one set's result records */
int mpcgpu_metrics_from_set(mpcgpu_metrics *m, int planners, const double *pobj, const int *exit_code, int best_idx,
                            const unsigned char *guided, double *objective_values_out);

#ifdef __cplusplus
}
#endif
#endif /* MPCGPU_WIRE_H */
