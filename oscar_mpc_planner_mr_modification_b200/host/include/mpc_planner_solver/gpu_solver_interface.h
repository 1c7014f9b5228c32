// gpu_solver_interface.h -- MPCPlanner::Solver backed by the B200 engine (libmpcgpu.so, include/mpcgpu.h).
//
// Public surface = the reference's acados wrapper, member for member
// (mpc_planner_solver/include/mpc_planner_solver/acados_solver_interface.h:51-222), so that Planner and
// every ControllerModule (mpc_planner_modules) compile and behave unchanged: same struct and member
// names (`AcadosParameters`, `_params`, `_info`, `_output`, ...), same exit codes, same warm-start
// helpers.  Only the private acados handles are replaced by an engine handle and the persistent
// capsule memory blob.  Additions (not in the reference): the static Solver::solveBatch().
#ifndef GPU_SOLVER_INTERFACE_H
#define GPU_SOLVER_INTERFACE_H

#include <iostream>
#include <vector>
#include <unordered_map>
#include <utility>

#include <mpc_planner_solver/state.h>

#include <mpc_planner_util/load_yaml.hpp>

#include <ros_tools/logging.h>

#include "solver_dims.h"   // generated: SOLVER_N / SOLVER_NX / SOLVER_NU / SOLVER_NP / SOLVER_NH

#define NX SOLVER_NX
#define NU SOLVER_NU
#define NH SOLVER_NH

struct mpcgpu_engine;

namespace MPCPlanner
{
    struct AcadosParameters
    {
        double xinit[NX];                      // Initial state
        double x0[(NU + NX) * (SOLVER_N + 1)]; // Warmstart: [u0, x0 | u1 x1 | ... | uN xN]

        double all_parameters[SOLVER_NP * SOLVER_N]; // SOLVER_NP parameters for all stages

        double solver_timeout{0.};

        double *getU0() { return x0; }

        AcadosParameters()
        {
            for (int i = 0; i < NX; i++)
                xinit[i] = 0.;
            for (int i = 0; i < (NU + NX) * (SOLVER_N + 1); i++)
                x0[i] = 0.;
            for (int i = 0; i < SOLVER_NP * SOLVER_N; i++)
                all_parameters[i] = 0.;
        }

        void printParameters(YAML::Node &parameter_map)
        {
            LOG_HEADER("Parameters");
            for (int k = 0; k < SOLVER_N; k++)
            {
                LOG_HEADER(k);
                for (YAML::const_iterator it = parameter_map.begin(); it != parameter_map.end(); ++it)
                {
                    if (it->first.as<std::string>() == "num parameters")
                        continue;
                    LOG_VALUE(it->first.as<std::string>(), all_parameters[k * SOLVER_NP + it->second.as<int>()]);
                }
            }
        }
    };

    // used by the generated setSolverParameter<Bundle>() functions (mpc_planner_parameters.h)
    inline void mpcgpu_set_parameter(AcadosParameters &params, int flat_index, double value) { params.all_parameters[flat_index] = value; }

    class Solver
    {
    public:
        struct AcadosInfo
        {
            double min_time;
            double kkt_norm_inf;
            double elapsed_time;
            int sqp_iter;
            double nlp_res;
            double solvetime;

            int qp_status;

            double pobj{0.};

            AcadosInfo()
            {
                min_time = 1e12;
                kkt_norm_inf = 0.; elapsed_time = 0.; sqp_iter = 0; nlp_res = 0.; solvetime = 0.; qp_status = 0;
            }
        };

        struct AcadosOutput
        {
            double xtraj[NX * (SOLVER_N + 1)];
            double utraj[NU * SOLVER_N];

            AcadosOutput()
            {
                for (int i = 0; i < NX * (SOLVER_N + 1); i++)
                    xtraj[i] = 0.;
                for (int i = 0; i < NU * SOLVER_N; i++)
                    utraj[i] = 0.;
            }
        };

    private:
        // name -> index tables built once in the constructor.  Replaces the per-call yaml-cpp lookups of
        // acados_solver_interface.cpp:212-225,229-272,379-389 (`_parameter_map[parameter].as<int>()` in every
        // setParameter / getOutput: the 3.3 ms "SetParameters" scope of the reference's traces)
        struct VarInfo { bool is_state; int index; double lb, ub; };
        std::unordered_map<std::string, int> _param_index;
        std::unordered_map<std::string, VarInfo> _var_index;
        std::vector<std::pair<std::string, VarInfo>> _vars;     // iteration order of _model_map
        int paramIndex(const std::string &name) const;
        const VarInfo &varInfo(const std::string &name) const;
        double _last_res_eq{0.};
        int _stepwise_status{0};

        mpcgpu_engine *_engine{nullptr};   // replaces the acados capsule + ocp_nlp_* handles
        std::vector<double> _mem;          // persistent capsule memory: NLP multipliers + QP warm start
        std::vector<double> _iterate;      // the solver's own (u,x) iterate (acados: nlp_out), x0 layout
        int _exit_code_one_iter{-1};
        int _iterations_requested{0};
        double _avg_iteration_time{0.};

        int numIterationsForTimeout() const;
        int finish(double pobj, int exit_code, int qp_status, double res_eq, int sqp_iter, double seconds);

    public:
        int _solver_id;

        AcadosParameters _params;
        AcadosInfo _info;
        AcadosOutput _output;

        int N;
        unsigned int nu;   // Number of control variables
        unsigned int nx;   // Differentiable variables
        unsigned int nvar; // Total variable count
        unsigned int npar; // Parameters per iteration
        double dt;

        YAML::Node _config, _parameter_map, _model_map;

        int _num_iterations;

    public:
        Solver(int solver_id = 0);
        ~Solver();
        Solver(const Solver &) = delete;

        /** @brief Copy data from another solver. Does not copy solver generic parameters like the horizon N*/
        Solver &operator=(const Solver &rhs);

        void reset();

        int solve();

        // One iteration a time interface
        void initializeOneIteration();
        int solveOneIteration();
        int completeOneIteration();

        // PARAMETERS //
        bool hasParameter(std::string &&parameter);
        void setParameter(int k, std::string &&parameter, double value);
        void setParameter(int k, std::string &parameter, double value);
        double getParameter(int k, std::string &&parameter);

        // XINIT //
        void setXinit(std::string &&state_name, double value);
        void setXinit(const State &state);

        // WARMSTART //
        void setEgoPrediction(unsigned int k, std::string &&var_name, double value);
        double getEgoPrediction(unsigned int k, std::string &&var_name);
        void setEgoPredictionPosition(unsigned int k, const Eigen::Vector2d &value);
        Eigen::Vector2d getEgoPredictionPosition(unsigned int k);

        void loadWarmstart();
        void initializeWarmstart(const State &state, bool shift_previous_solution_forward);
        void initializeWithState(const State &initial_state);
        void initializeWithBraking(const State &initial_state);

        // OUTPUT //
        double getOutput(int k, std::string &&state_name) const;

        // DEBUG //
        std::string explainExitFlag(int exitflag) const;
        void printIfBoundLimited() const;

        // ADDITION: solve several Solver objects (the planners of one homotopy set, or of several robots)
        // in ONE batched engine call; exit_codes[i] is what solvers[i]->solve() would have returned.
        static void solveBatch(const std::vector<Solver *> &solvers, std::vector<int> &exit_codes);
    };
}

#endif // GPU_SOLVER_INTERFACE_H
