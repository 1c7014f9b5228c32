// GENERATED -- struct-of-tables parameter path of configuration c7_linearized: stage-invariant parameters travel
// once per homotopy set, obstacle predictions as a table; the engine expands them to all_parameters on the device
// (mpcgpu_solve_sets_tables, include/mpcgpu.h).  Replaces: the k-loop over modules->setParameters (mpc_planner/src/planner.cpp:153-159)
// through the generated setSolverParameter<Bundle> if-chains (solver_generator/generate_cpp_files.py:235-254).
#pragma once
namespace MPCPlanner {
struct SolverTables {
    static constexpr int n_invariant = 54;
    // flat parameter indices, in the order of `invariant` below
    static constexpr int invariant_idx[54] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 36, 37, 38, 39, 40, 41, 42, 43, 44, 45, 46, 47, 48, 49, 50, 51, 52, 53};
    double invariant[54] = {};   // acceleration, angular_velocity, velocity, reference_velocity, contour, lag, terminal_angle, terminal_contouring, ...
    static constexpr int inv_acceleration = 0;
    static constexpr int inv_angular_velocity = 1;
    static constexpr int inv_velocity = 2;
    static constexpr int inv_reference_velocity = 3;
    static constexpr int inv_contour = 4;
    static constexpr int inv_lag = 5;
    static constexpr int inv_terminal_angle = 6;
    static constexpr int inv_terminal_contouring = 7;
    static constexpr int inv_spline_x0_a = 8;
    static constexpr int inv_spline_x0_b = 9;
    static constexpr int inv_spline_x0_c = 10;
    static constexpr int inv_spline_x0_d = 11;
    static constexpr int inv_spline_y0_a = 12;
    static constexpr int inv_spline_y0_b = 13;
    static constexpr int inv_spline_y0_c = 14;
    static constexpr int inv_spline_y0_d = 15;
    static constexpr int inv_spline0_start = 16;
    static constexpr int inv_spline_x1_a = 17;
    static constexpr int inv_spline_x1_b = 18;
    static constexpr int inv_spline_x1_c = 19;
    static constexpr int inv_spline_x1_d = 20;
    static constexpr int inv_spline_y1_a = 21;
    static constexpr int inv_spline_y1_b = 22;
    static constexpr int inv_spline_y1_c = 23;
    static constexpr int inv_spline_y1_d = 24;
    static constexpr int inv_spline1_start = 25;
    static constexpr int inv_spline_x2_a = 26;
    static constexpr int inv_spline_x2_b = 27;
    static constexpr int inv_spline_x2_c = 28;
    static constexpr int inv_spline_x2_d = 29;
    static constexpr int inv_spline_y2_a = 30;
    static constexpr int inv_spline_y2_b = 31;
    static constexpr int inv_spline_y2_c = 32;
    static constexpr int inv_spline_y2_d = 33;
    static constexpr int inv_spline2_start = 34;
    static constexpr int inv_spline_x3_a = 35;
    static constexpr int inv_spline_x3_b = 36;
    static constexpr int inv_spline_x3_c = 37;
    static constexpr int inv_spline_x3_d = 38;
    static constexpr int inv_spline_y3_a = 39;
    static constexpr int inv_spline_y3_b = 40;
    static constexpr int inv_spline_y3_c = 41;
    static constexpr int inv_spline_y3_d = 42;
    static constexpr int inv_spline3_start = 43;
    static constexpr int inv_spline_x4_a = 44;
    static constexpr int inv_spline_x4_b = 45;
    static constexpr int inv_spline_x4_c = 46;
    static constexpr int inv_spline_x4_d = 47;
    static constexpr int inv_spline_y4_a = 48;
    static constexpr int inv_spline_y4_b = 49;
    static constexpr int inv_spline_y4_c = 50;
    static constexpr int inv_spline_y4_d = 51;
    static constexpr int inv_spline4_start = 52;
    static constexpr int inv_ego_disc_0_offset = 53;
};
}  // namespace MPCPlanner
