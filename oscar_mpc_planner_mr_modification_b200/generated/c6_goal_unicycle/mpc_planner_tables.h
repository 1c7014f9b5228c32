// GENERATED -- struct-of-tables parameter path of configuration c6_goal_unicycle: stage-invariant parameters travel
// once per homotopy set, obstacle predictions as a table; the engine expands them to all_parameters on the device
// (mpcgpu_solve_sets_tables, include/mpcgpu.h).  Replaces: the k-loop over modules->setParameters (mpc_planner/src/planner.cpp:153-159)
// through the generated setSolverParameter<Bundle> if-chains (solver_generator/generate_cpp_files.py:235-254).
#pragma once
namespace MPCPlanner {
struct SolverTables {
    static constexpr int n_invariant = 7;
    // flat parameter indices, in the order of `invariant` below
    static constexpr int invariant_idx[7] = {0, 1, 2, 3, 4, 5, 6};
    double invariant[7] = {};   // acceleration, angular_velocity, goal_weight, goal_x, goal_y, ego_disc_radius, ego_disc_0_offset
    static constexpr int inv_acceleration = 0;
    static constexpr int inv_angular_velocity = 1;
    static constexpr int inv_goal_weight = 2;
    static constexpr int inv_goal_x = 3;
    static constexpr int inv_goal_y = 4;
    static constexpr int inv_ego_disc_radius = 5;
    static constexpr int inv_ego_disc_0_offset = 6;
    static constexpr int max_obstacles = 4, ell_base = 7, ell_stride = 7;
    static constexpr int ell_offsets[7] = {0, 1, 2, 3, 4, 5, 6};   // x, y, psi, major, minor, chi, r inside an obstacle's block
};
}  // namespace MPCPlanner
