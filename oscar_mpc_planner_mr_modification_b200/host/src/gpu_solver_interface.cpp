// gpu_solver_interface.cpp -- MPCPlanner::Solver on the B200 engine.  Method-by-method counterpart of
// the reference's mpc_planner_solver/src/acados_solver_interface.cpp (line numbers cited per method);
// the acados calls are replaced by ONE C-ABI call into libmpcgpu.so (include/mpcgpu.h).
#include <mpc_planner_solver/gpu_solver_interface.h>

#include <mpc_planner_util/parameters.h>

#include <ros_tools/profiling.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "mpcgpu.h"

namespace MPCPlanner
{
    namespace
    {
        int deviceFromEnv()
        {
            const char *e = std::getenv("MPCGPU_DEVICE");
            return e ? std::atoi(e) : 0;
        }
        // Engine(s) shared by Solver::solveBatch (all planners of a set / several robots in one launch).  MPCGPU_DEVICES=0,1,..
        // lists several GPUs: the batch is then partitioned over them by mpcgpu_multi_solve_batch (contiguous ranges).
        std::mutex g_batch_mutex;
        mpcgpu_engine *g_batch_engine = nullptr;
        mpcgpu_multi *g_batch_multi = nullptr;
        const int kBatchCapacity = 256;
        std::vector<int> devicesFromEnv()
        {
            std::vector<int> d;
            if (const char *e = std::getenv("MPCGPU_DEVICES"))
                for (const char *p = e; *p;)
                {
                    char *end = nullptr;
                    const long v = std::strtol(p, &end, 10);
                    if (end == p)
                        break;
                    d.push_back((int)v);
                    p = (*end == ',') ? end + 1 : end;
                }
            return d;
        }
        // pinned staging of solveBatch, grown on demand and reused: no allocation per control cycle
        struct Staging
        {
            void *base = nullptr;
            size_t bytes = 0;
            void *get(size_t need)
            {
                if (need > bytes)
                {
                    if (base)
                        mpcgpu_free_pinned(base);
                    base = nullptr;
                    bytes = 0;
                    if (mpcgpu_alloc_pinned(need + need / 2, &base) != 0)
                    {
                        printf("mpcgpu_alloc_pinned(%zu) failed. Exiting.\n", need);
                        exit(1);
                    }
                    bytes = need + need / 2;
                }
                return base;
            }
        } g_staging;
        double seconds(std::chrono::steady_clock::time_point t0)
        {
            return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        }
    }

    // acados_solver_interface.cpp:9-49
    Solver::Solver(int solver_id)
    {
        _solver_id = solver_id;

        loadConfigYaml(SYSTEM_CONFIG_PATH(__FILE__, "solver_settings"), _config);
        loadConfigYaml(SYSTEM_CONFIG_PATH(__FILE__, "parameter_map"), _parameter_map);
        loadConfigYaml(SYSTEM_CONFIG_PATH(__FILE__, "model_map"), _model_map);

        N = SOLVER_N;
        nu = _config["nu"].as<unsigned int>();
        nx = _config["nx"].as<unsigned int>();
        nvar = _config["nvar"].as<unsigned int>();
        npar = _config["npar"].as<unsigned int>();
        dt = CONFIG["integrator_step"].as<double>();

        _num_iterations = CONFIG["solver_settings"]["acados"]["iterations"].as<int>();
        if (CONFIG["solver_settings"]["acados"]["solver_type"].as<std::string>() == "SQP")
            _num_iterations = 1;

        int status = mpcgpu_engine_create(MPCGPU_CONFIG_NAME, deviceFromEnv(), 1, &_engine);
        int eN = 0, enx = 0, enu = 0, enp = 0, enh = 0;
        if (status == 0)
            mpcgpu_desc_query(_engine, &eN, &enx, &enu, &enp, &enh);
        if (status || eN != N || enx != (int)nx || enu != (int)nu || enp != (int)npar)
        {
            printf("mpcgpu_engine_create(%s) returned status %d (%s). Exiting.\n", MPCGPU_CONFIG_NAME, status,
                   _engine ? mpcgpu_last_error(_engine) : "no engine");
            exit(1); // same behaviour as a failed Solver_acados_create (:35-39)
        }
        _mem.assign(mpcgpu_mem_doubles(_engine), 0.0);
        _iterate.assign((size_t)nvar * (N + 1), 0.0);

        // name -> index tables (one YAML traversal here instead of one lookup per setParameter / getOutput call)
        for (YAML::const_iterator it = _parameter_map.begin(); it != _parameter_map.end(); ++it)
            _param_index[it->first.as<std::string>()] = it->second.as<int>();
        for (YAML::const_iterator it = _model_map.begin(); it != _model_map.end(); ++it)
        {
            VarInfo v{it->second[0].as<std::string>() == "x", it->second[1].as<int>(), it->second[2].as<double>(), it->second[3].as<double>()};
            _var_index[it->first.as<std::string>()] = v;
            _vars.emplace_back(it->first.as<std::string>(), v);
        }

        reset();
    }

    // :51-65
    Solver::~Solver()
    {
        if (_engine)
            mpcgpu_engine_destroy(_engine);
    }

    // :67-77 -- copies the parameters only and resets the QP memory (keeps the NLP multipliers)
    Solver &Solver::operator=(const Solver &rhs)
    {
        _params = rhs._params;
        if (!_mem.empty() && _mem[0] > 1.0)
            _mem[0] = 1.0; // ocp_nlp_solver_reset_qp_memory: multipliers stay, QP warm start is dropped
        return *this;
    }

    // :79-84
    void Solver::reset()
    {
        _params = AcadosParameters();
        _info = AcadosInfo();
        _output = AcadosOutput();
    }

    // The reference stops iterating when `elapsed + average iteration time >= solver_timeout` (:108-116),
    // a wall-clock rule.  The engine takes the iteration count as an explicit input, so the rule is
    // evaluated here with the average iteration time of the previous solves: a non-positive timeout (the
    // fork's effective setting) means exactly one iteration.
    int Solver::numIterationsForTimeout() const
    {
        if (_params.solver_timeout <= 0.)
            return 1;
        if (_avg_iteration_time <= 0.)
            return _num_iterations;
        const int m = (int)std::ceil(_params.solver_timeout / _avg_iteration_time - 1.);
        return std::max(1, std::min(_num_iterations, m));
    }

    // :86-119
    int Solver::solve()
    {
        initializeOneIteration();
        const int iterations = numIterationsForTimeout();
        const auto t0 = std::chrono::steady_clock::now();
        double pobj = 0., res_eq = 0.;
        int exit_code = 0, qp_status = 0, ipm = 0;
        int status = mpcgpu_solve_batch(_engine, 1, _params.xinit, _iterate.data(), _params.all_parameters, nullptr, iterations,
                                        _mem.data(), _output.xtraj, _output.utraj, &pobj, &exit_code, &qp_status, &res_eq, &ipm);
        if (status)
        {
            LOG_ERROR("mpcgpu_solve_batch failed: " << mpcgpu_last_error(_engine));
            return 0;
        }
        return finish(pobj, exit_code, qp_status, res_eq, iterations, seconds(t0));
    }

    int Solver::finish(double pobj, int exit_code, int qp_status, double res_eq, int sqp_iter, double secs)
    {
        _info.pobj = pobj;
        _info.qp_status = qp_status;
        _info.nlp_res = res_eq;
        _info.sqp_iter = sqp_iter;
        _info.elapsed_time = secs;
        _info.solvetime += secs;
        _info.min_time = std::min(secs, _info.min_time);
        if (sqp_iter > 0)
            _avg_iteration_time = (_avg_iteration_time <= 0.) ? secs / sqp_iter : 0.5 * (_avg_iteration_time + secs / sqp_iter);
        _exit_code_one_iter = exit_code;
        // a later solve() without loadWarmstart() continues from this solution, as acados' nlp_out does
        for (int k = 0; k <= N; k++)
        {
            for (unsigned int i = 0; i < nx; i++)
                _iterate[k * nvar + nu + i] = _output.xtraj[k * nx + i];
            if (k < N)
                for (unsigned int i = 0; i < nu; i++)
                    _iterate[k * nvar + i] = _output.utraj[k * nu + i];
        }
        return exit_code; // already mapped: 1 success, 0, 2, 3, 4 (:197-203)
    }

    // :121-143 -- xinit / parameters travel with the engine call; here only the per-solve info is reset
    void Solver::initializeOneIteration()
    {
        _info = AcadosInfo();
        _iterations_requested = 0;
    }

    // :145-160 -- one SQP-RTI iteration, continuing from the stored iterate and capsule memory.  The engine runs it with the
    // completion step deferred (num_iter = -1): multipliers and QP memory survive between iterations exactly as inside
    // Solver_acados_solve; the res_eq demotion and the reset on failure happen once, in completeOneIteration().
    int Solver::solveOneIteration()
    {
        const auto t0 = std::chrono::steady_clock::now();
        double pobj = 0., res_eq = 0.;
        int exit_code = 0, qp_status = 0, ipm = 0;
        int status = mpcgpu_solve_batch(_engine, 1, _params.xinit, _iterate.data(), _params.all_parameters, nullptr, -1, _mem.data(),
                                        _output.xtraj, _output.utraj, &pobj, &exit_code, &qp_status, &res_eq, &ipm);
        if (status)
            return 1;
        _iterations_requested++;
        finish(pobj, exit_code, qp_status, res_eq, _iterations_requested, seconds(t0));
        _last_res_eq = res_eq;
        _stepwise_status = exit_code == 1 ? 0 : (exit_code == 0 ? 1 : exit_code); // back to the acados convention
        return _stepwise_status;
    }

    // :162-204 -- res_eq rule (:176-181), reset on failure (:187-191), exit-code map (:197-203)
    int Solver::completeOneIteration()
    {
        int status = _stepwise_status;
        if (!(_last_res_eq <= 1e-2) && status == 0)
            status = 4; // ACADOS_QP_FAILURE
        if (status != 0)
            std::fill(_mem.begin(), _mem.end(), 0.0); // Solver_acados_reset + ocp_nlp_solver_reset_qp_memory
        _exit_code_one_iter = (status == 0) ? 1 : (status == 1 ? 0 : status);
        return _exit_code_one_iter;
    }

    int Solver::paramIndex(const std::string &name) const
    {
        auto it = _param_index.find(name);
        if (it == _param_index.end())
        {
            LOG_ERROR("unknown solver parameter: " << name);
            exit(1); // the reference throws from yaml-cpp's as<int>() on an undefined node
        }
        return it->second;
    }

    const Solver::VarInfo &Solver::varInfo(const std::string &name) const
    {
        auto it = _var_index.find(name);
        if (it == _var_index.end())
        {
            LOG_ERROR("unknown model variable: " << name);
            exit(1);
        }
        return it->second;
    }

    // PARAMETERS // (:207-225)
    bool Solver::hasParameter(std::string &&parameter) { return _param_index.count(parameter) != 0; }

    void Solver::setParameter(int k, std::string &&parameter, double value)
    {
        _params.all_parameters[k * npar + paramIndex(parameter)] = value;
    }

    void Solver::setParameter(int k, std::string &parameter, double value)
    {
        _params.all_parameters[k * npar + paramIndex(parameter)] = value;
    }

    double Solver::getParameter(int k, std::string &&parameter)
    {
        return _params.all_parameters[k * npar + paramIndex(parameter)];
    }

    // XINIT // (:229-246)
    void Solver::setXinit(std::string &&state_name, double value)
    {
        _params.xinit[varInfo(state_name).index - nu] = value;
    }

    void Solver::setXinit(const State &state)
    {
        for (const auto &v : _vars)
            if (v.second.is_state)
                _params.xinit[v.second.index - nu] = state.get(std::string(v.first));
    }

    // WARMSTART // (:250-376)
    void Solver::setEgoPrediction(unsigned int k, std::string &&var_name, double value)
    {
        _params.x0[k * nvar + varInfo(var_name).index] = value;
    }

    double Solver::getEgoPrediction(unsigned int k, std::string &&var_name)
    {
        return _params.x0[k * nvar + varInfo(var_name).index];
    }

    void Solver::setEgoPredictionPosition(unsigned int k, const Eigen::Vector2d &value)
    {
        setEgoPrediction(k, "x", value(0));
        setEgoPrediction(k, "y", value(1));
    }

    Eigen::Vector2d Solver::getEgoPredictionPosition(unsigned int k)
    {
        return Eigen::Vector2d(getEgoPrediction(k, "x"), getEgoPrediction(k, "y"));
    }

    // :274-284 -- x0 -> the solver's internal iterate (acados: ocp_nlp_out_set "x"/"u")
    void Solver::loadWarmstart()
    {
        std::copy(_params.x0, _params.x0 + (size_t)nvar * (N + 1), _iterate.begin());
    }

    void Solver::initializeWithState(const State &initial_state)
    {
        for (int k = 0; k <= N; k++)
            for (const auto &v : _vars)
                _params.x0[k * nvar + v.second.index] = v.second.is_state ? initial_state.get(std::string(v.first)) : 0.;
    }

    void Solver::initializeWithBraking(const State &initial_state)
    {
        initializeWithState(initial_state);

        double x, y, psi, v, a, spline;
        double deceleration = std::abs(CONFIG["deceleration_at_infeasible"].as<double>());

        x = initial_state.get("x");
        y = initial_state.get("y");
        psi = initial_state.get("psi");
        v = initial_state.get("v");
        spline = initial_state.get("spline");
        a = -deceleration;

        for (int k = 0; k <= N; k++)
        {
            if (k > 0)
            {
                x += v * dt * std::cos(psi);
                y += v * dt * std::sin(psi);
                spline += v * dt;
                v += a * dt;
                v = std::max(v, 0.);
            }
            setEgoPrediction(k, "x", x);
            setEgoPrediction(k, "y", y);
            setEgoPrediction(k, "psi", psi);
            setEgoPrediction(k, "v", v);
            setEgoPrediction(k, "spline", spline);
            setEgoPrediction(k, "a", a);
            setEgoPrediction(k, "w", 0);
        }
    }

    // :344-376
    void Solver::initializeWarmstart(const State &initial_state, bool shift_previous_solution_forward)
    {
        auto output = [&](int k, const VarInfo &v) { return v.is_state ? _output.xtraj[k * nx + v.index - nu] : _output.utraj[k * nu + v.index]; };
        if (shift_previous_solution_forward)
        {
            // [initial_state, x_2, x_3, ..., x_N-1, x_N-1]
            for (int k = 0; k <= N; k++)
                for (const auto &v : _vars)
                {
                    double value;
                    if (k == 0) // the reference also calls initial_state.get() for the INPUT names here (:355), which indexes
                                // State::_state at -nu (state.cpp:22: out of bounds); the shim loads 0 for inputs instead
                        value = v.second.is_state ? initial_state.get(std::string(v.first)) : 0.;
                    else if (k >= N - 1)
                        value = output(N - 1, v.second);
                    else
                        value = output(k + 1, v.second);
                    _params.x0[k * nvar + v.second.index] = value;
                }
        }
        else
        {
            // [initial_state, x_1, x_2, ..., x_N-1, x_N]  (the reference copies stages 0..N-1 of the previous output, :366-375)
            for (int k = 0; k < N; k++)
                for (const auto &v : _vars)
                    _params.x0[k * nvar + v.second.index] = output(k, v.second);
        }
    }

    // OUTPUT // (:379-389)
    double Solver::getOutput(int k, std::string &&state_name) const
    {
        const VarInfo &v = varInfo(state_name);
        return v.is_state ? _output.xtraj[k * nx + v.index - nu] : _output.utraj[k * nu + v.index];
    }

    // :391-424
    std::string Solver::explainExitFlag(int exitflag) const
    {
        switch (exitflag)
        {
        case 1:
            return "Success";
        case 0:
            return "Failure (no more information)";
        case 2:
            return "Failure (maximum number of iterations reached)";
        case 3:
            return "Failure (minimum step size reached)";
        case 4:
            break;
        default:
            return "Unknown exit code; code: " + std::to_string(exitflag);
        }

        switch (_info.qp_status)
        {
        case 1:
            return "QP Failure: No more information on QP failure";
        case 2:
            return "QP Failure: Max Iterations";
        case 3:
            return "QP Failure: Minimal Step Reached";
        case 4:
            return "QP Failure: NAN in solution";
        case 5:
            return "QP Failure: Inconsistent Equality Constraints";
        default:
            return "QP Failure: UNKNOWN";
        }
    }

    // :426-448
    void Solver::printIfBoundLimited() const
    {
        for (int k = 0; k < N; k++)
            for (const auto &v : _vars)
            {
                if (k == 0 && v.second.is_state)
                    continue;
                const double value = v.second.is_state ? _output.xtraj[k * nx + v.second.index - nu] : _output.utraj[k * nu + v.second.index];
                if (std::abs(value - v.second.lb) < 1e-2)
                    LOG_WARN_THROTTLE(500, v.first + " limited by lower bound");
                if (std::abs(value - v.second.ub) < 1e-2)
                    LOG_WARN_THROTTLE(500, v.first + " limited by upper bound");
            }
    }

    // ADDITION -- what `#pragma omp parallel for` over planners (guidance_constraints.cpp:304-370) becomes:
    // one engine call for all planners.
    void Solver::solveBatch(const std::vector<Solver *> &solvers, std::vector<int> &exit_codes)
    {
        const int n = (int)solvers.size();
        exit_codes.assign(n, 0);
        if (n == 0)
            return;
        std::lock_guard<std::mutex> lock(g_batch_mutex);      // one shared engine: callers take turns (a call is one launch)
        if (!g_batch_engine && !g_batch_multi)
        {
            const std::vector<int> devs = devicesFromEnv();
            int rc = devs.size() > 1 ? mpcgpu_multi_create(MPCGPU_CONFIG_NAME, devs.data(), (int)devs.size(), kBatchCapacity, &g_batch_multi)
                                     : mpcgpu_engine_create(MPCGPU_CONFIG_NAME, devs.size() == 1 ? devs[0] : deviceFromEnv(), kBatchCapacity, &g_batch_engine);
            if (rc)
            {
                printf("mpcgpu engine creation (batch) failed with status %d. Exiting.\n", rc);
                exit(1);
            }
        }
        const int capacity = kBatchCapacity * (g_batch_multi ? mpcgpu_multi_num_devices(g_batch_multi) : 1);
        const Solver *s0 = solvers[0];
        const size_t N = s0->N, nx = s0->nx, nu = s0->nu, nz = s0->nvar, np = s0->npar, md = s0->_mem.size(), B = (size_t)n;
        // one pinned block, carved up: [xinit | x0 | par | mem | xt | ut | pobj | req | nit | ec | qs | ipm]
        const size_t nd = B * (nx + nz * (N + 1) + N * np + md + nx * (N + 1) + nu * N + 2);
        char *base = (char *)g_staging.get(nd * sizeof(double) + 4 * B * sizeof(int));
        double *xinit = (double *)base, *x0 = xinit + B * nx, *par = x0 + B * nz * (N + 1), *mem = par + B * N * np, *xt = mem + B * md,
               *ut = xt + B * nx * (N + 1), *pobj = ut + B * nu * N, *req = pobj + B;
        int *nit = (int *)(req + B), *ec = nit + B, *qs = ec + B, *ipm = qs + B;
        for (int i = 0; i < n; i++)
        {
            Solver *s = solvers[i];
            s->initializeOneIteration();
            nit[i] = s->numIterationsForTimeout();
            std::memcpy(xinit + i * nx, s->_params.xinit, sizeof(double) * nx);
            std::memcpy(x0 + i * nz * (N + 1), s->_iterate.data(), sizeof(double) * nz * (N + 1));
            std::memcpy(par + i * N * np, s->_params.all_parameters, sizeof(double) * N * np);
            std::memcpy(mem + i * md, s->_mem.data(), sizeof(double) * md);
        }
        const auto t0 = std::chrono::steady_clock::now();
        for (int b = 0; b < n; b += capacity)
        {
            const int m = std::min(capacity, n - b);
            const size_t o = (size_t)b;
            int status = g_batch_multi
                             ? mpcgpu_multi_solve_batch(g_batch_multi, m, xinit + o * nx, x0 + o * nz * (N + 1), par + o * N * np, nit + o, 0, mem + o * md,
                                                        xt + o * nx * (N + 1), ut + o * nu * N, pobj + o, ec + o, qs + o, req + o, ipm + o)
                             : mpcgpu_solve_batch(g_batch_engine, m, xinit + o * nx, x0 + o * nz * (N + 1), par + o * N * np, nit + o, 0, mem + o * md,
                                                  xt + o * nx * (N + 1), ut + o * nu * N, pobj + o, ec + o, qs + o, req + o, ipm + o);
            if (status)
            {
                LOG_ERROR("mpcgpu_solve_batch failed with status " << status);
                return;
            }
        }
        const double secs = seconds(t0);
        for (int i = 0; i < n; i++)
        {
            Solver *s = solvers[i];
            std::memcpy(s->_output.xtraj, xt + i * nx * (N + 1), sizeof(double) * nx * (N + 1));
            std::memcpy(s->_output.utraj, ut + i * nu * N, sizeof(double) * nu * N);
            std::memcpy(s->_mem.data(), mem + i * md, sizeof(double) * md);
            exit_codes[i] = s->finish(pobj[i], ec[i], qs[i], req[i], nit[i], secs);
        }
    }
}
