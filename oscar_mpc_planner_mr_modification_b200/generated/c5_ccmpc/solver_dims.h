// GENERATED -- dimensions the reference takes from Solver/acados_solver_Solver.h
#pragma once
#define SOLVER_N 50
#define SOLVER_NX 5
#define SOLVER_NU 2
#define SOLVER_NP 115
#define SOLVER_NH 16
#define MPCGPU_CONFIG_NAME "c5_ccmpc"
