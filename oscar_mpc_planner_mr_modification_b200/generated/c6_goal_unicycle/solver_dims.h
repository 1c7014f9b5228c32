// GENERATED -- dimensions the reference takes from Solver/acados_solver_Solver.h
#pragma once
#define SOLVER_N 30
#define SOLVER_NX 4
#define SOLVER_NU 2
#define SOLVER_NP 35
#define SOLVER_NH 4
#define MPCGPU_CONFIG_NAME "c6_goal_unicycle"
