// Drop-in for mpc_planner_solver/include/mpc_planner_solver/solver_interface.h:4-12 with one more branch:
// -DGPU_SOLVER selects the B200 engine behind the unchanged Solver API.
#ifndef __MPC_PLANNER_SOLVER_H__
#define __MPC_PLANNER_SOLVER_H__

#if defined(GPU_SOLVER)
#include <mpc_planner_solver/gpu_solver_interface.h>
#elif defined(ACADOS_SOLVER)
#include <mpc_planner_solver/acados_solver_interface.h>
#else
#include <mpc_planner_solver/forces_solver_interface.h>
#endif

#endif
