// Tests of the GPU-backed MPCPlanner::Solver, written after the reference's own
// mpc_planner_solver/test/test_solver.cpp:52-134 (State get/set, setParameter/getParameter, setXinit,
// setEgoPrediction indices, operator= copies params) for the acados buffer layout, plus -- when a GPU is
// present -- solve() against a direct C-ABI call and solveBatch() against per-solver solve().
//   test_solver_shim <settings.yaml> [--no-gpu]
#include <mpc_planner_solver/solver_interface.h>
#include <mpc_planner_util/parameters.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include "mpc_planner_parameters.h"
#include "mpcgpu.h"

using namespace MPCPlanner;

static int g_fail = 0;
#define CHECK(c) do { if (!(c)) { std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #c); g_fail++; } } while (0)
#define CHECK_NEAR(a, b, t) CHECK(std::abs((a) - (b)) <= (t))

static void fillProblem(Solver &s, State &st, double lateral)
{
    // weights (settings.yaml:78-92) and a straight reference path along +x, identical for every stage
    const char *names[] = {"acceleration", "angular_velocity", "velocity", "reference_velocity", "contour", "lag", "terminal_angle", "terminal_contouring"};
    const double vals[] = {0.34, 0.85, 0.55, 2.0, 0.05, 0.75, 100.0, 10.0};
    for (int k = 0; k < s.N; k++)
    {
        for (int i = 0; i < 8; i++)
            s.setParameter(k, std::string(names[i]), vals[i]);
        for (int seg = 0; seg < 5; seg++)
        {
            setSolverParameterSplineXA(k, s._params, 0.0, seg); setSolverParameterSplineXB(k, s._params, 0.0, seg);
            setSolverParameterSplineXC(k, s._params, 1.0, seg); setSolverParameterSplineXD(k, s._params, 5.0 * seg, seg);
            setSolverParameterSplineYA(k, s._params, 0.0, seg); setSolverParameterSplineYB(k, s._params, 0.0, seg);
            setSolverParameterSplineYC(k, s._params, 0.0, seg); setSolverParameterSplineYD(k, s._params, 0.0, seg);
            setSolverParameterSplineStart(k, s._params, 5.0 * seg, seg);
        }
        setSolverParameterEgoDiscRadius(k, s._params, 0.325);
        setSolverParameterEgoDiscOffset(k, s._params, 0.0, 0);
        for (int o = 0; o < SOLVER_NH && s.hasParameter("ellipsoid_obst_" + std::to_string(o) + "_x"); o++)
        {
            setSolverParameterEllipsoidObstX(k, s._params, 6.0 + 0.1 * k, o); setSolverParameterEllipsoidObstY(k, s._params, 1.5 + lateral + 3.0 * o, o);
            setSolverParameterEllipsoidObstPsi(k, s._params, 0.0, o); setSolverParameterEllipsoidObstMajor(k, s._params, 0.0, o);
            setSolverParameterEllipsoidObstMinor(k, s._params, 0.0, o); setSolverParameterEllipsoidObstChi(k, s._params, 1.0, o);
            setSolverParameterEllipsoidObstR(k, s._params, 0.325, o);
        }
        if (s.hasParameter("lin_constraint_0_a1"))
            for (int h = 0; s.hasParameter("lin_constraint_" + std::to_string(h) + "_a1"); h++)
            {
                setSolverParameterLinConstraintA1(k, s._params, 1.0, h); setSolverParameterLinConstraintA2(k, s._params, 0.0, h);
                setSolverParameterLinConstraintB(k, s._params, 100.0, h);
            }
    }
    st.set("x", 0.1); st.set("y", lateral); st.set("psi", 0.05); st.set("v", 1.5); st.set("spline", 0.0);
    s.setXinit(st);
    s.initializeWithBraking(st);
    for (int k = 0; k <= s.N; k++) s.setEgoPrediction(k, "a", 0.0);
}

int main(int argc, char **argv)
{
    if (argc < 2) { std::printf("usage: %s settings.yaml [--no-gpu]\n", argv[0]); return 2; }
    const bool gpu = !(argc > 2 && std::strcmp(argv[2], "--no-gpu") == 0);
    Configuration::getInstance().initialize(argv[1]);

    // ---- State (test_solver.cpp:52-71)
    State state;
    state.set("x", 2.5); state.set("v", 1.0);
    CHECK(state.get("x") == 2.5 && state.get("v") == 1.0 && state.get("y") == 0.0);
    CHECK(state.validData());
    CHECK(state.getPos()(0) == 2.5);
    State zero;
    CHECK(!zero.validData());

    if (!gpu)
    {
        // without a device the constructor must fail loudly (exit(1) like a failed Solver_acados_create);
        // only the host-side pieces that need no engine are covered here.
        YAML::Node pm = YAML::LoadFile(SYSTEM_CONFIG_PATH(__FILE__, "parameter_map"));
        CHECK(pm["acceleration"].as<int>() == 0);
        CHECK(pm["num parameters"].as<int>() == SOLVER_NP);
        AcadosParameters p;
        setSolverParameterSplineXA(3, p, 7.0, 2);
        CHECK(p.all_parameters[3 * SOLVER_NP + pm["spline_x2_a"].as<int>()] == 7.0);
        setSolverParameterEgoDiscRadius(0, p, 0.3);
        CHECK(p.all_parameters[pm["ego_disc_radius"].as<int>()] == 0.3);
        CHECK(CONFIG["N"].as<int>() == 30 && CONFIG["solver_settings"]["acados"]["iterations"].as<int>() == 10);
        CHECK(CONFIG["weights"]["lag"].as<double>() == 0.75 && CONFIG["name"].as<std::string>() == "jackal");
        std::printf("%s (host-only)\n", g_fail ? "FAILED" : "OK");
        return g_fail ? 1 : 0;
    }

    // ---- Solver buffers (test_solver.cpp:73-134, acados layout)
    Solver solver(0);
    CHECK(solver.N == SOLVER_N && solver.nx == SOLVER_NX && solver.nu == SOLVER_NU && solver.npar == SOLVER_NP && solver.nvar == 7);
    CHECK(solver._num_iterations == 10 && solver.dt == 0.2);
    solver.setParameter(4, "lag", 3.25);
    CHECK(solver.getParameter(4, "lag") == 3.25);
    CHECK(solver._params.all_parameters[4 * solver.npar + 5] == 3.25);
    CHECK(solver.hasParameter("contour") && !solver.hasParameter("does_not_exist"));
    solver.setXinit(state);
    CHECK(solver._params.xinit[0] == 2.5 && solver._params.xinit[3] == 1.0);
    solver.setEgoPrediction(2, "v", 0.7);
    CHECK(solver._params.x0[2 * 7 + 5] == 0.7 && solver.getEgoPrediction(2, "v") == 0.7);
    solver.setEgoPredictionPosition(3, Eigen::Vector2d(1.0, 2.0));
    CHECK(solver._params.x0[3 * 7 + 2] == 1.0 && solver._params.x0[3 * 7 + 3] == 2.0 && solver.getEgoPredictionPosition(3)(1) == 2.0);
    Solver other(1);
    other = solver;
    CHECK(other.getParameter(4, "lag") == 3.25 && other._params.xinit[0] == 2.5);      // operator= copies _params only
    CHECK(solver.explainExitFlag(1) == "Success" && solver.explainExitFlag(2) == "Failure (maximum number of iterations reached)");
    solver.reset();
    CHECK(solver.getParameter(4, "lag") == 0.0);

    // ---- solve(): the shim against a direct C-ABI call with the same buffers
    State st;
    fillProblem(solver, st, 0.2);
    solver._params.solver_timeout = 1.0;          // positive: all _num_iterations are run on the first call
    solver.loadWarmstart();
    int exit_code = solver.solve();
    CHECK(exit_code == 1);
    CHECK(solver._info.pobj > 0 && (solver._info.qp_status == 0 || solver._info.qp_status == 2) && solver._info.sqp_iter == 10);
    CHECK_NEAR(solver.getOutput(0, "x"), 0.1, 1e-9);
    CHECK(solver.getOutput(5, "v") > 1.0 && solver.getOutput(solver.N, "x") > 5.0);
    mpcgpu_engine *eng = nullptr;
    CHECK(mpcgpu_engine_create(MPCGPU_CONFIG_NAME, 0, 4, &eng) == 0);
    std::vector<double> xt(SOLVER_NX * (SOLVER_N + 1)), ut(SOLVER_NU * SOLVER_N);
    double pobj, req; int ec, qs, ipm;
    CHECK(mpcgpu_solve_batch(eng, 1, solver._params.xinit, solver._params.x0, solver._params.all_parameters, nullptr, 10, nullptr,
                             xt.data(), ut.data(), &pobj, &ec, &qs, &req, &ipm) == 0);
    CHECK(ec == exit_code && pobj == solver._info.pobj);
    for (size_t i = 0; i < xt.size(); i++) CHECK(xt[i] == solver._output.xtraj[i]);
    mpcgpu_engine_destroy(eng);

    // non-positive timeout (the fork's effective behaviour): exactly one iteration
    Solver one(2);
    fillProblem(one, st, 0.2);
    one._params.solver_timeout = -1.0;
    one.loadWarmstart();
    one.solve();
    CHECK(one._info.sqp_iter == 1);

    // one-iteration-at-a-time interface == solve() with the same number of iterations
    Solver stepwise(3), whole(4);
    fillProblem(stepwise, st, -0.3); fillProblem(whole, st, -0.3);
    whole._num_iterations = 3; whole._params.solver_timeout = 10.0;
    whole.loadWarmstart();
    int e_whole = whole.solve();
    stepwise.loadWarmstart();
    stepwise.initializeOneIteration();
    for (int i = 0; i < 3; i++) stepwise.solveOneIteration();
    int e_step = stepwise.completeOneIteration();
    CHECK(e_whole == e_step);
    for (int i = 0; i < SOLVER_NX * (SOLVER_N + 1); i++) CHECK_NEAR(whole._output.xtraj[i], stepwise._output.xtraj[i], 1e-9);

    // ---- solveBatch(): all planners in one engine call == one solve() each
    std::vector<std::unique_ptr<Solver>> a, b;
    std::vector<Solver *> pa;
    for (int i = 0; i < 5; i++)
    {
        a.emplace_back(new Solver(10 + i)); b.emplace_back(new Solver(20 + i));
        fillProblem(*a[i], st, 0.1 * i - 0.2); fillProblem(*b[i], st, 0.1 * i - 0.2);
        a[i]->_params.solver_timeout = 1.0; b[i]->_params.solver_timeout = 1.0;
        a[i]->loadWarmstart(); b[i]->loadWarmstart();
        pa.push_back(a[i].get());
    }
    std::vector<int> codes;
    Solver::solveBatch(pa, codes);
    for (int i = 0; i < 5; i++)
    {
        int e = b[i]->solve();
        CHECK(e == codes[i]);
        CHECK(a[i]->_info.pobj == b[i]->_info.pobj);
        CHECK(std::memcmp(a[i]->_output.xtraj, b[i]->_output.xtraj, sizeof(a[i]->_output.xtraj)) == 0);
    }

    // ---- QP failure decoding (acados_solver_interface.cpp:391-424): the engine reports qp_status in the numbering the reference
    //      decodes -- 2 max iterations, 3 minimal step, 4 NaN
    {
        Solver dec(30);
        dec._info.qp_status = 2; CHECK(dec.explainExitFlag(4) == "QP Failure: Max Iterations");
        dec._info.qp_status = 3; CHECK(dec.explainExitFlag(4) == "QP Failure: Minimal Step Reached");
        dec._info.qp_status = 4; CHECK(dec.explainExitFlag(4) == "QP Failure: NAN in solution");
        CHECK(dec.explainExitFlag(0) == "Failure (no more information)" && dec.explainExitFlag(3) == "Failure (minimum step size reached)");
        // a NaN parameter: the interior-point loop stops with the NaN status, solve() returns ACADOS_QP_FAILURE (4)
        Solver nan_case(31);
        fillProblem(nan_case, st, 0.2);
        nan_case._params.solver_timeout = 1.0;
        nan_case.setParameter(3, "lag", std::nan(""));
        nan_case.loadWarmstart();
        int e = nan_case.solve();
        CHECK(e == 4 && nan_case._info.qp_status == 4);
        CHECK(nan_case.explainExitFlag(e) == "QP Failure: NAN in solution");
        // an infeasible QP (two contradicting halfspaces x <= -100 and x >= 100 on every stage): the interior-point iteration
        // diverges; it ends at the iteration limit or with a vanishing step, and the message must name that reason
        if (nan_case.hasParameter("lin_constraint_1_a1"))
        {
            Solver inf_case(32);
            fillProblem(inf_case, st, 0.2);
            inf_case._params.solver_timeout = 1.0;
            for (int k = 1; k < inf_case.N; k++)
            {
                setSolverParameterLinConstraintA1(k, inf_case._params, 1.0, 0); setSolverParameterLinConstraintA2(k, inf_case._params, 0.0, 0);
                setSolverParameterLinConstraintB(k, inf_case._params, -100.0, 0);
                setSolverParameterLinConstraintA1(k, inf_case._params, -1.0, 1); setSolverParameterLinConstraintA2(k, inf_case._params, 0.0, 1);
                setSolverParameterLinConstraintB(k, inf_case._params, -100.0, 1);
            }
            inf_case.loadWarmstart();
            e = inf_case.solve();
            CHECK(e != 1);
            const int q = inf_case._info.qp_status;
            CHECK(q == 2 || q == 3 || q == 4);
            if (e == 4)
                CHECK(inf_case.explainExitFlag(e) == (q == 2 ? "QP Failure: Max Iterations" : (q == 3 ? "QP Failure: Minimal Step Reached" : "QP Failure: NAN in solution")));
            std::printf("infeasible case: exit %d qp_status %d (%s)\n", e, q, inf_case.explainExitFlag(e).c_str());
        }
    }

    // ---- warm starts (acados_solver_interface.cpp:286-376) against hand-computed values
    {
        Solver ws(40);
        for (int k = 0; k <= ws.N; k++)
            for (int i = 0; i < SOLVER_NX; i++) ws._output.xtraj[k * SOLVER_NX + i] = 100.0 * i + k;        // x_k[i] = 100 i + k
        for (int k = 0; k < ws.N; k++)
            for (int i = 0; i < SOLVER_NU; i++) ws._output.utraj[k * SOLVER_NU + i] = -(10.0 * i + k) - 1.0;  // u_k[i] = -(10 i + k) - 1
        State now;
        now.set("x", 7.0); now.set("y", 8.0); now.set("psi", 0.3); now.set("v", 1.1); now.set("spline", 2.2);
        const int N = ws.N;
        ws.initializeWarmstart(now, true);      // [initial_state, x_2, x_3, ..., x_N-1, x_N-1]
        CHECK(ws.getEgoPrediction(0, "x") == 7.0 && ws.getEgoPrediction(0, "y") == 8.0 && ws.getEgoPrediction(0, "spline") == 2.2);
        CHECK(ws.getEgoPrediction(1, "x") == 2.0 && ws.getEgoPrediction(1, "y") == 102.0 && ws.getEgoPrediction(1, "v") == 302.0);
        CHECK(ws.getEgoPrediction(5, "psi") == 206.0 && ws.getEgoPrediction(5, "a") == -7.0 && ws.getEgoPrediction(5, "w") == -17.0);
        CHECK(ws.getEgoPrediction(N - 2, "x") == N - 1.0 && ws.getEgoPrediction(N - 1, "x") == N - 1.0 && ws.getEgoPrediction(N, "x") == N - 1.0);
        CHECK(ws.getEgoPrediction(N - 1, "a") == -(N - 1.0) - 1.0);
        Solver keep(41);
        std::memcpy(keep._output.xtraj, ws._output.xtraj, sizeof(ws._output.xtraj));
        std::memcpy(keep._output.utraj, ws._output.utraj, sizeof(ws._output.utraj));
        keep.setEgoPrediction(N, "x", -5.0);
        keep.initializeWarmstart(now, false);   // previous output kept for stages 0..N-1; stage N is not touched (:366-375)
        CHECK(keep.getEgoPrediction(0, "x") == 0.0 && keep.getEgoPrediction(3, "y") == 103.0 && keep.getEgoPrediction(N - 1, "v") == 300.0 + N - 1);
        CHECK(keep.getEgoPrediction(4, "w") == -15.0 && keep.getEgoPrediction(N, "x") == -5.0);
        Solver flat(42);
        flat.initializeWithState(now);          // every stage = the state, inputs zero (:286-301)
        for (int k = 0; k <= N; k += 7)
            CHECK(flat.getEgoPrediction(k, "x") == 7.0 && flat.getEgoPrediction(k, "v") == 1.1 && flat.getEgoPrediction(k, "a") == 0.0 && flat.getEgoPrediction(k, "w") == 0.0);
        Solver brk(43);
        brk.initializeWithBraking(now);         // :303-342 with deceleration_at_infeasible
        const double dec = std::abs(CONFIG["deceleration_at_infeasible"].as<double>());
        double x = 7.0, y = 8.0, v = 1.1, sp = 2.2;
        for (int k = 1; k <= 3; k++) { x += v * 0.2 * std::cos(0.3); y += v * 0.2 * std::sin(0.3); sp += v * 0.2; v = std::max(v - dec * 0.2, 0.0); }
        CHECK(brk.getEgoPrediction(3, "x") == x && brk.getEgoPrediction(3, "y") == y && brk.getEgoPrediction(3, "v") == v && brk.getEgoPrediction(3, "spline") == sp);
        CHECK(brk.getEgoPrediction(3, "a") == -dec && brk.getEgoPrediction(N, "v") == 0.0);
        brk._output.xtraj[2 * SOLVER_NX + 3] = 3.0 - 0.005;      // v at its upper bound (3.0): printIfBoundLimited must not crash
        brk.printIfBoundLimited();
    }

    // ---- threading contract: solve() from 8 threads on DISTINCT Solver objects, as the OpenMP team of
    //      guidance_constraints.cpp:304,369 does -- bit-identical to the serial results
    {
        const int T = 8;
        std::vector<std::unique_ptr<Solver>> par, ser;
        for (int i = 0; i < T; i++)
        {
            par.emplace_back(new Solver(50 + i)); ser.emplace_back(new Solver(60 + i));
            fillProblem(*par[i], st, 0.07 * i - 0.25); fillProblem(*ser[i], st, 0.07 * i - 0.25);
            par[i]->_params.solver_timeout = 1.0; ser[i]->_params.solver_timeout = 1.0;
        }
        std::vector<int> ep(T), es(T);
        std::vector<std::thread> th;
        for (int i = 0; i < T; i++)
            th.emplace_back([&, i] { for (int rep = 0; rep < 3; rep++) { par[i]->loadWarmstart(); ep[i] = par[i]->solve(); } });
        for (auto &t : th) t.join();
        for (int i = 0; i < T; i++)
        {
            for (int rep = 0; rep < 3; rep++) { ser[i]->loadWarmstart(); es[i] = ser[i]->solve(); }
            CHECK(ep[i] == es[i] && par[i]->_info.pobj == ser[i]->_info.pobj && par[i]->_info.qp_status == ser[i]->_info.qp_status);
            CHECK(std::memcmp(par[i]->_output.xtraj, ser[i]->_output.xtraj, sizeof(ser[i]->_output.xtraj)) == 0);
            CHECK(std::memcmp(par[i]->_output.utraj, ser[i]->_output.utraj, sizeof(ser[i]->_output.utraj)) == 0);
        }
    }
    std::printf("%s\n", g_fail ? "FAILED" : "OK");
    return g_fail ? 1 : 0;
}
