// GENERATED -- mirrors mpc_planner_solver/include/mpc_planner_solver/mpc_planner_parameters.h
#pragma once
namespace MPCPlanner {
struct AcadosParameters;
inline void setSolverParameterAcceleration(int k, AcadosParameters& params, const double value, int index = 0)
{ static const int idx[1] = {0}; mpcgpu_set_parameter(params, k * 72 + idx[index], value); }
inline void setSolverParameterAngularVelocity(int k, AcadosParameters& params, const double value, int index = 0)
{ static const int idx[1] = {1}; mpcgpu_set_parameter(params, k * 72 + idx[index], value); }
inline void setSolverParameterVelocity(int k, AcadosParameters& params, const double value, int index = 0)
{ static const int idx[1] = {2}; mpcgpu_set_parameter(params, k * 72 + idx[index], value); }
inline void setSolverParameterReferenceVelocity(int k, AcadosParameters& params, const double value, int index = 0)
{ static const int idx[1] = {3}; mpcgpu_set_parameter(params, k * 72 + idx[index], value); }
inline void setSolverParameterContour(int k, AcadosParameters& params, const double value, int index = 0)
{ static const int idx[1] = {4}; mpcgpu_set_parameter(params, k * 72 + idx[index], value); }
inline void setSolverParameterLag(int k, AcadosParameters& params, const double value, int index = 0)
{ static const int idx[1] = {5}; mpcgpu_set_parameter(params, k * 72 + idx[index], value); }
inline void setSolverParameterTerminalAngle(int k, AcadosParameters& params, const double value, int index = 0)
{ static const int idx[1] = {6}; mpcgpu_set_parameter(params, k * 72 + idx[index], value); }
inline void setSolverParameterTerminalContouring(int k, AcadosParameters& params, const double value, int index = 0)
{ static const int idx[1] = {7}; mpcgpu_set_parameter(params, k * 72 + idx[index], value); }
inline void setSolverParameterSplineXA(int k, AcadosParameters& params, const double value, int index = 0)
{ static const int idx[5] = {8, 17, 26, 35, 44}; mpcgpu_set_parameter(params, k * 72 + idx[index], value); }
inline void setSolverParameterSplineXB(int k, AcadosParameters& params, const double value, int index = 0)
{ static const int idx[5] = {9, 18, 27, 36, 45}; mpcgpu_set_parameter(params, k * 72 + idx[index], value); }
inline void setSolverParameterSplineXC(int k, AcadosParameters& params, const double value, int index = 0)
{ static const int idx[5] = {10, 19, 28, 37, 46}; mpcgpu_set_parameter(params, k * 72 + idx[index], value); }
inline void setSolverParameterSplineXD(int k, AcadosParameters& params, const double value, int index = 0)
{ static const int idx[5] = {11, 20, 29, 38, 47}; mpcgpu_set_parameter(params, k * 72 + idx[index], value); }
inline void setSolverParameterSplineYA(int k, AcadosParameters& params, const double value, int index = 0)
{ static const int idx[5] = {12, 21, 30, 39, 48}; mpcgpu_set_parameter(params, k * 72 + idx[index], value); }
inline void setSolverParameterSplineYB(int k, AcadosParameters& params, const double value, int index = 0)
{ static const int idx[5] = {13, 22, 31, 40, 49}; mpcgpu_set_parameter(params, k * 72 + idx[index], value); }
inline void setSolverParameterSplineYC(int k, AcadosParameters& params, const double value, int index = 0)
{ static const int idx[5] = {14, 23, 32, 41, 50}; mpcgpu_set_parameter(params, k * 72 + idx[index], value); }
inline void setSolverParameterSplineYD(int k, AcadosParameters& params, const double value, int index = 0)
{ static const int idx[5] = {15, 24, 33, 42, 51}; mpcgpu_set_parameter(params, k * 72 + idx[index], value); }
inline void setSolverParameterSplineStart(int k, AcadosParameters& params, const double value, int index = 0)
{ static const int idx[5] = {16, 25, 34, 43, 52}; mpcgpu_set_parameter(params, k * 72 + idx[index], value); }
inline void setSolverParameterEgoDiscOffset(int k, AcadosParameters& params, const double value, int index = 0)
{ static const int idx[1] = {53}; mpcgpu_set_parameter(params, k * 72 + idx[index], value); }
inline void setSolverParameterLinConstraintA1(int k, AcadosParameters& params, const double value, int index = 0)
{ static const int idx[6] = {54, 57, 60, 63, 66, 69}; mpcgpu_set_parameter(params, k * 72 + idx[index], value); }
inline void setSolverParameterLinConstraintA2(int k, AcadosParameters& params, const double value, int index = 0)
{ static const int idx[6] = {55, 58, 61, 64, 67, 70}; mpcgpu_set_parameter(params, k * 72 + idx[index], value); }
inline void setSolverParameterLinConstraintB(int k, AcadosParameters& params, const double value, int index = 0)
{ static const int idx[6] = {56, 59, 62, 65, 68, 71}; mpcgpu_set_parameter(params, k * 72 + idx[index], value); }
}  // namespace MPCPlanner
