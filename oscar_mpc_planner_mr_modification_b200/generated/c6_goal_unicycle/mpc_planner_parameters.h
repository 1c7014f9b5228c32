// GENERATED -- mirrors mpc_planner_solver/include/mpc_planner_solver/mpc_planner_parameters.h
#pragma once
namespace MPCPlanner {
struct AcadosParameters;
inline void setSolverParameterAcceleration(int k, AcadosParameters& params, const double value, int index = 0)
{ static const int idx[1] = {0}; mpcgpu_set_parameter(params, k * 35 + idx[index], value); }
inline void setSolverParameterAngularVelocity(int k, AcadosParameters& params, const double value, int index = 0)
{ static const int idx[1] = {1}; mpcgpu_set_parameter(params, k * 35 + idx[index], value); }
inline void setSolverParameterGoalWeight(int k, AcadosParameters& params, const double value, int index = 0)
{ static const int idx[1] = {2}; mpcgpu_set_parameter(params, k * 35 + idx[index], value); }
inline void setSolverParameterGoalX(int k, AcadosParameters& params, const double value, int index = 0)
{ static const int idx[1] = {3}; mpcgpu_set_parameter(params, k * 35 + idx[index], value); }
inline void setSolverParameterGoalY(int k, AcadosParameters& params, const double value, int index = 0)
{ static const int idx[1] = {4}; mpcgpu_set_parameter(params, k * 35 + idx[index], value); }
inline void setSolverParameterEgoDiscRadius(int k, AcadosParameters& params, const double value, int index = 0)
{ static const int idx[1] = {5}; mpcgpu_set_parameter(params, k * 35 + idx[index], value); }
inline void setSolverParameterEgoDiscOffset(int k, AcadosParameters& params, const double value, int index = 0)
{ static const int idx[1] = {6}; mpcgpu_set_parameter(params, k * 35 + idx[index], value); }
inline void setSolverParameterEllipsoidObstX(int k, AcadosParameters& params, const double value, int index = 0)
{ static const int idx[4] = {7, 14, 21, 28}; mpcgpu_set_parameter(params, k * 35 + idx[index], value); }
inline void setSolverParameterEllipsoidObstY(int k, AcadosParameters& params, const double value, int index = 0)
{ static const int idx[4] = {8, 15, 22, 29}; mpcgpu_set_parameter(params, k * 35 + idx[index], value); }
inline void setSolverParameterEllipsoidObstPsi(int k, AcadosParameters& params, const double value, int index = 0)
{ static const int idx[4] = {9, 16, 23, 30}; mpcgpu_set_parameter(params, k * 35 + idx[index], value); }
inline void setSolverParameterEllipsoidObstMajor(int k, AcadosParameters& params, const double value, int index = 0)
{ static const int idx[4] = {10, 17, 24, 31}; mpcgpu_set_parameter(params, k * 35 + idx[index], value); }
inline void setSolverParameterEllipsoidObstMinor(int k, AcadosParameters& params, const double value, int index = 0)
{ static const int idx[4] = {11, 18, 25, 32}; mpcgpu_set_parameter(params, k * 35 + idx[index], value); }
inline void setSolverParameterEllipsoidObstChi(int k, AcadosParameters& params, const double value, int index = 0)
{ static const int idx[4] = {12, 19, 26, 33}; mpcgpu_set_parameter(params, k * 35 + idx[index], value); }
inline void setSolverParameterEllipsoidObstR(int k, AcadosParameters& params, const double value, int index = 0)
{ static const int idx[4] = {13, 20, 27, 34}; mpcgpu_set_parameter(params, k * 35 + idx[index], value); }
}  // namespace MPCPlanner
