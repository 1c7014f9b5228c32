"""CUDA emitter: the counterpart of the reference's `generate_acados_solver.py`.

    generate_cuda_solver(modules, settings, model, name, out_dir)

takes the SAME three objects the reference hands to `generate_acados_solver(modules, settings, model,
skip)` (solver_generator/generate_acados_solver.py:68) -- a `ModuleManager`, the settings dict and a
`DynamicsModel` -- and instead of calling acados it writes, for the B200 engine:

  model.cuh                    __device__ functions: interval map Phi (ERK4 x 3) with sensitivities and
                               multiplier-weighted second-order term, stage cost with gradient/Hessian,
                               constraint residuals with Jacobian/Hessian, and the sparsity-specialised
                               W'PW / W v / W'y products the Riccati sweeps use
  parameter_map.yaml           name -> index            (util/parameters.py:69-74)
  model_map.yaml               name -> [x|u, idx, lb, ub]   (solver_model.py:118-128)
  solver_settings.yaml         N, nx, nu, nvar, npar    (generate_solver.py:38-46)
  mpc_planner_parameters.h     setSolverParameter<Bundle>() (generate_cpp_files.py:204-260)

The module / model objects are used duck-typed exactly as solver_definition.py:5-76 uses them, so the
reference's own Python modules (imported through the sympy-backed casadi stand-in when casadi is
absent) drop in untouched.  Symbolic differentiation is done with sympy.
"""
import os

import numpy as np
import sympy as sp
import yaml

from .symbolic import emit_block, to_exprs, symbolic_namespace

BOUND_INF = 1e10   # |bound| >= this => bound absent (same contract as the oracle)
BIG = 1e15         # generate_acados_solver.py:17-24


def _idx(i, j):
    """packed lower-triangular index"""
    if i < j:
        i, j = j, i
    return i * (i + 1) // 2 + j


def extract_problem(modules, settings, model):
    """Mirror of generate_acados_solver.py:27-75 + solver_definition.py:5-76: parameters in
    objective-then-constraint order, stage cost and constraints at stage_idx = 1."""
    cd = symbolic_namespace()
    from_params = _Parameters(cd)
    for module in modules.modules:
        if module.type == "objective":
            module.define_parameters(from_params)
    for module in modules.modules:
        if module.type == "constraint":
            module.define_parameters(from_params)
    from_params.make_symbols()
    settings = dict(settings)
    settings["params"] = from_params

    z = model.acados_symbolics()
    f_expl, _ = model.get_acados_dynamics()
    from_params.load(from_params.symbols)
    model.load(z)
    cost = 0.0
    for module in modules.modules:
        if module.type == "objective":
            cost = cost + module.get_value(model, from_params, settings, 1)
    h, lh, uh = [], [], []
    for module in modules.modules:
        if module.type == "constraint":
            for c in module.constraints:
                h += c.get_constraints(model, from_params, settings, 1)
                lh += list(c.get_lower_bound())
                uh += list(c.get_upper_bound())
    hx = []
    for c in h:
        hx.extend(to_exprs(c))
    clip = lambda v: -BIG if v == -np.inf else (BIG if v == np.inf else float(v))
    return dict(
        N=int(settings["N"]), dt=float(settings["integrator_step"]), nx=model.nx, nu=model.nu,
        states=list(model.states), inputs=list(model.inputs),
        lb=[float(v) for v in model.lower_bound], ub=[float(v) for v in model.upper_bound],
        z=to_exprs(z), p=[to_exprs(q)[0] for q in from_params.symbols], param_names=list(from_params.names),
        bundles=dict(from_params.parameter_bundles), f=to_exprs(f_expl), cost=to_exprs(cost)[0], h=hx,
        lh=[clip(v) for v in lh], uh=[clip(v) for v in uh],
    )


class _Parameters:
    """The subset of util/parameters.py `Parameters` the modules use (add / get / has_parameter)."""

    def __init__(self, cd):
        self._cd = cd
        self.names = []
        self._index = {}
        self.parameter_bundles = {}
        self.symbols = None
        self._p = None

    def add(self, parameter, add_to_rqt_reconfigure=False, rqt_config_name=None, bundle_name=None,
            rqt_min_value=0.0, rqt_max_value=100.0):
        if parameter in self._index:
            return
        self._index[parameter] = len(self.names)
        self.parameter_bundles.setdefault(bundle_name or parameter, []).append(len(self.names))
        self.names.append(parameter)

    def length(self):
        return len(self.names)

    def has_parameter(self, parameter):
        return parameter in self._index

    def make_symbols(self):
        self.symbols = [self._cd.SX.sym(n, 1) for n in self.names]

    def load(self, p):
        self._p = p

    def get_p(self):
        return self._p

    def get(self, parameter):
        return self._p[self._index[parameter]]


def _interval_map(pb, sim_steps=3):
    """Phi = sim_steps explicit RK4 steps of f over one shooting interval (symbolic)."""
    z, f, nu, nx = pb["z"], pb["f"], pb["nu"], pb["nx"]
    x = z[nu:]
    h = sp.Float(pb["dt"]) / sim_steps

    def F(xv):
        sub = dict(zip(x, xv))
        return [fi.xreplace(sub) for fi in f]

    y = list(x)
    for _ in range(sim_steps):
        k1 = F(y)
        k2 = F([y[i] + h / 2 * k1[i] for i in range(nx)])
        k3 = F([y[i] + h / 2 * k2[i] for i in range(nx)])
        k4 = F([y[i] + h * k3[i] for i in range(nx)])
        y = [y[i] + h / 6 * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]) for i in range(nx)]
    return y


def _group_rows(h, p, min_len=3):
    """Find runs of consecutive constraint rows that are the SAME expression with a block of parameters
    shifted by a constant stride (the reference defines obstacles / halfspaces in a Python loop, so the
    parameter blocks are contiguous: ellipsoid_constraints.py:41-49, guidance_constraints.py:76-81).
    Returns [(first_row, count, stride, fixed_param_indices)], rows outside any run are emitted flat."""
    pidx = {sym: i for i, sym in enumerate(p)}
    par = [sorted(pidx[q] for q in e.free_symbols if q in pidx) for e in h]
    groups, i = [], 0
    while i < len(h) - 1:
        fixed = sorted(set(par[i]) & set(par[i + 1]))
        s0 = [q for q in par[i] if q not in fixed]
        s1 = [q for q in par[i + 1] if q not in fixed]
        ok = len(s0) == len(s1) and len(s0) > 0 and len({b - a for a, b in zip(s0, s1)}) == 1 and s1[0] > s0[0]
        if not ok:
            i += 1
            continue
        stride = s1[0] - s0[0]
        cnt = 1
        while i + cnt < len(h):
            shift = {p[q]: p[q + cnt * stride] for q in s0 if q + cnt * stride < len(p)}
            if len(shift) != len(s0) or h[i].xreplace(shift) != h[i + cnt]:      # structural equality (no simplify: can be very slow)
                break
            cnt += 1
        if cnt >= min_len:
            groups.append((i, cnt, stride, fixed))
            i += cnt
        else:
            i += 1
    return groups


def emit_model_header(pb, name, sim_steps=3):
    z, p, cost, h = pb["z"], pb["p"], pb["cost"], pb["h"]
    nu, nx = pb["nu"], pb["nx"]
    nz, npar, nh = nu + nx, len(p), len(h)
    npk = nz * (nz + 1) // 2
    zmap = {z[i]: "z[%d]" % i for i in range(nz)}
    pmap = {p[i]: "p[%d]" % i for i in range(npar)}
    m = dict(zmap)
    m.update(pmap)

    out = []
    w = out.append
    w("// GENERATED by oscar_mpc_planner_mr_modification_b200/solver_generator/generate_cuda_solver.py -- do not edit.")
    w("// configuration: %s" % name)
    w("#pragma once")
    w("#ifndef MPCGEN_RCP      // reciprocal used by the emitted expressions; the kernel header may map it to its branch-free version")
    w("#define MPCGEN_RCP(x) (1.0/(x))")
    w("#endif")
    w("namespace mpcgen {")
    w("constexpr int NX = %d, NU = %d, NZ = %d, NP = %d, NH = %d, NSTAGE = %d;" % (nx, nu, nz, npar, nh, pb["N"]))
    w("constexpr int NPK = %d;  // packed lower-triangular NZ x NZ" % npk)
    w("constexpr double DT = %r;" % pb["dt"])
    w("constexpr int SIM_STEPS = %d;" % sim_steps)
    w("#define MPCGEN_CONFIG_NAME \"%s\"" % name)
    w("__device__ constexpr double LBZ[%d] = {%s};" % (nz, ", ".join(repr(v) for v in pb["lb"])))
    w("__device__ constexpr double UBZ[%d] = {%s};" % (nz, ", ".join(repr(v) for v in pb["ub"])))

    # general constraint entries (one per finite side), support variables of h
    hrow, hsgn, hbnd = [], [], []
    for i in range(nh):
        if pb["lh"][i] > -BOUND_INF:
            hrow.append(i); hsgn.append(1.0); hbnd.append(pb["lh"][i])
        if pb["uh"][i] < BOUND_INF:
            hrow.append(i); hsgn.append(-1.0); hbnd.append(pb["uh"][i])
    ncg = len(hrow)
    sup = [i for i in range(nz) if any(e.has(z[i]) for e in h)]
    nhs = len(sup)
    w("constexpr int NCG = %d;   // general inequality entries (finite sides of lh <= h <= uh)" % ncg)
    w("constexpr int NHS = %d;   // variables h depends on" % max(nhs, 1))
    w("__device__ constexpr int HSUP[%d] = {%s};" % (max(nhs, 1), ", ".join(str(i) for i in sup) or "0"))
    w("__device__ constexpr int HROW[%d] = {%s};" % (max(ncg, 1), ", ".join(str(i) for i in hrow) or "0"))
    w("__device__ constexpr double HSGN[%d] = {%s};" % (max(ncg, 1), ", ".join(repr(v) for v in hsgn) or "0"))
    w("__device__ constexpr double HBND[%d] = {%s};" % (max(ncg, 1), ", ".join(repr(v) for v in hbnd) or "0"))

    # ---- dynamics over one interval -----------------------------------------------------------
    phi = _interval_map(pb, sim_steps)
    J = [[sp.diff(phi[i], z[j]) for j in range(nz)] for i in range(nx)]
    wv = [(i, j) for i in range(nx) for j in range(nz) if J[i][j] != 0 and not J[i][j].is_number]
    wconst = [(i, j, float(J[i][j])) for i in range(nx) for j in range(nz) if J[i][j].is_number and J[i][j] != 0]
    w("constexpr int NWV = %d;   // state-dependent entries of W = dPhi/d[u;x]; the rest are constants" % max(len(wv), 1))
    w("// varying entries (row, col): %s" % wv)
    w("// constant entries (row, col, value): %s" % wconst)
    w("\n// x+ = Phi(x,u): %d explicit RK4 steps (generate_acados_solver.py:148-150)" % sim_steps)
    w("__device__ __forceinline__ void dyn_phi(const double* z, double* xn)\n{")
    w(emit_block([("xn[%d]" % i, phi[i]) for i in range(nx)], zmap))
    w("}")
    pis = sp.symbols("pi0:%d" % nx)
    pimap = {pis[i]: "pi[%d]" % i for i in range(nx)}
    L = sum(pis[i] * phi[i] for i in range(nx))
    gL = [sp.diff(L, v) for v in z]
    Hd = {(i, j): sp.diff(gL[i], z[j]) for i in range(nz) for j in range(i + 1)}
    Hd = {k: v for k, v in Hd.items() if v != 0}
    m2 = dict(zmap)
    m2.update(pimap)
    w("\n// xn = Phi, Wv = varying entries of dPhi/dz, H(packed) += sum_j pi_j d2Phi_j/dz2")
    w("__device__ __forceinline__ void dyn_lin(const double* z, const double* pi, double* xn, double* Wv, double* H)\n{")
    w(emit_block([("xn[%d]" % i, phi[i]) for i in range(nx)] +
                 [("Wv[%d]" % n, J[i][j]) for n, (i, j) in enumerate(wv)] +
                 [("H[%d] +" % _idx(i, j), e) for (i, j), e in sorted(Hd.items())], m2))
    w("}")

    # W-products, specialised to the sparsity / constants of W
    Wsym = sp.zeros(nx, nz)
    wsyms = sp.symbols("wv0:%d" % max(len(wv), 1))
    wmap = {wsyms[n]: "Wv[%d]" % n for n in range(len(wv))}
    for n, (i, j) in enumerate(wv):
        Wsym[i, j] = wsyms[n]
    for i, j, v in wconst:
        Wsym[i, j] = sp.Integer(1) if v == 1.0 else sp.Float(v)
    npx = nx * (nx + 1) // 2
    Ps = sp.symbols("P0:%d" % npx)
    Pm = sp.Matrix(nx, nx, lambda i, j: Ps[_idx(i, j)])
    Pmap = {Ps[i]: "P[%d]" % i for i in range(npx)}
    # two-step product T = P W, G += W' T: both steps only touch the structural non-zeros of W
    Ts = sp.Matrix(nx, nz, lambda i, j: sp.Symbol("T%d_%d" % (i, j)))
    Tm = Pm * Wsym
    mm = dict(wmap)
    mm.update(Pmap)
    G = Wsym.T * Ts
    w("\n// G(packed NZ) += W' P W   (P packed NX); T = P W first, then W' T")
    w("__device__ __forceinline__ void wtpw_add(const double* Wv, const double* P, double* G)\n{")
    w(emit_block([("const double T%d_%d" % (i, j), Tm[i, j]) for i in range(nx) for j in range(nz)], mm, tmp_prefix="a"))
    mt = dict(wmap)
    mt.update({Ts[i, j]: "T%d_%d" % (i, j) for i in range(nx) for j in range(nz)})
    w(emit_block([("G[%d] +" % _idx(i, j), G[i, j]) for i in range(nz) for j in range(i + 1) if G[i, j] != 0], mt, tmp_prefix="b"))
    w("}")
    w("\n// dense row-major NX x NZ sensitivity matrix from its varying entries (diagnostics / tests)")
    w("__device__ __forceinline__ void w_to_dense(const double* Wv, double* Wd)\n{")
    w(emit_block([("Wd[%d]" % (i * nz + j), Wsym[i, j]) for i in range(nx) for j in range(nz)], dict(wmap), tmp_prefix="c"))
    w("}")
    vs = sp.symbols("v0:%d" % nz)
    ys = sp.symbols("y0:%d" % nx)
    Wv_ = Wsym * sp.Matrix(vs)
    Wty = Wsym.T * sp.Matrix(ys)
    mv = dict(wmap); mv.update({vs[i]: "v[%d]" % i for i in range(nz)})
    my = dict(wmap); my.update({ys[i]: "y[%d]" % i for i in range(nx)})
    w("\n// out(NX) += W v")
    w("__device__ __forceinline__ void w_mul_add(const double* Wv, const double* v, double* out)\n{")
    w(emit_block([("out[%d] +" % i, Wv_[i]) for i in range(nx) if Wv_[i] != 0], mv))
    w("}")
    w("\n// out(NZ) += W' y")
    w("__device__ __forceinline__ void wt_mul_add(const double* Wv, const double* y, double* out)\n{")
    w(emit_block([("out[%d] +" % i, Wty[i]) for i in range(nz) if Wty[i] != 0], my))
    w("}")

    # ---- cost -----------------------------------------------------------------------------------
    w("\n// stage cost l(z,p), unscaled (solver_definition.py:19-35 with the reference's modules)")
    w("template <class PT>      // PT: const double* (reference layout, index k*npar + idx) or any type with operator[] and operator+ (e.g. a strided view)\n__device__ __forceinline__ double cost_val(const double* z, const PT p)\n{\n    double l;")
    w(emit_block([("l", cost)], m))
    w("    return l;\n}")
    g = [sp.diff(cost, v) for v in z]
    Hc = {(i, j): sp.diff(g[i], z[j]) for i in range(nz) for j in range(i + 1)}
    # Diagonal blocks of the stage Hessian when the constraints add no coupling (e.g. zero disc offset: h does not
    # depend on psi): connected components of the sparsity of (cost Hessian + dynamics second-order term).  The
    # kernel checks the coupling entries at run time and then mirrors the blocks separately.
    adj = {i: {i} for i in range(nz)}
    for (i, j), e in list(Hc.items()) + list(Hd.items()):
        if e != 0:
            adj[i].add(j); adj[j].add(i)
    comp, seen = [], set()
    for i in range(nz):
        if i in seen:
            continue
        stack, c = [i], []
        while stack:
            q_ = stack.pop()
            if q_ in seen:
                continue
            seen.add(q_); c.append(q_); stack.extend(adj[q_] - seen)
        comp.append(sorted(c))
    bmax = max(len(c) for c in comp)
    w("constexpr int HBLK_N = %d, HBLK_MAX = %d;   // Hessian blocks without constraint coupling: %s" % (len(comp), bmax, comp))
    w("__device__ constexpr int HBLK_SIZE[%d] = {%s};" % (len(comp), ", ".join(str(len(c)) for c in comp)))
    w("__device__ constexpr int HBLK_IDX[%d][%d] = {%s};" % (len(comp), bmax, ", ".join(
        "{" + ", ".join(str(v) for v in c + [-1] * (bmax - len(c))) + "}" for c in comp)))
    w("\n// g = DT * grad l ; H(packed) = DT * hess l   (all NPK entries written)")
    w("template <class PT>\n__device__ __forceinline__ void cost_lin(const double* z, const PT p, double* g, double* H)\n{")
    w(emit_block([("g[%d]" % i, pb["dt"] * g[i]) for i in range(nz)] +
                 [("H[%d]" % _idx(i, j), pb["dt"] * Hc[(i, j)]) for i in range(nz) for j in range(i + 1)], m))
    w("}")

    # ---- constraints ------------------------------------------------------------------------------
    groups = _group_rows(h, p) if nh else []
    in_group = set()
    for (r0, cnt, stride, fixed) in groups:
        in_group.update(range(r0, r0 + cnt))
    flat = [i for i in range(nh) if i not in in_group]
    w("// constraint rows emitted as loops over parameter blocks (first row, count, parameter stride): %s"
      % [(g[0], g[1], g[2]) for g in groups])

    def qmap(r0, fixed):
        """row r0 of a group: shifting parameters are read through q = p + r*stride"""
        mq = dict(zmap)
        for i, sym in enumerate(p):
            mq[sym] = ("p[%d]" if i in fixed else "q[%d]") % i
        return mq

    mhs = sp.symbols("mh0:%d" % max(nh, 1))

    def hess_outputs(rows, msyms):
        """packed-Hessian accumulations of sum_r msyms[r] h_r over the support of h"""
        Lh = sum(msyms[i] * h[i] for i in rows)
        gh = {s_: sp.diff(Lh, z[s_]) for s_ in sup}
        outs = []
        for a_ in sup:
            for b_ in sup:
                if b_ <= a_:
                    e = sp.diff(gh[a_], z[b_])
                    if e != 0:
                        outs.append(("H[%d] +" % _idx(a_, b_), e))
        return outs

    # ---- both at once: one traversal of the rows with the common subexpressions (cos / sin / sqrt / reciprocals of the
    #      obstacle parameters) shared between the Jacobian and the multiplier-weighted Hessian
    w("\n// hv[NH] = h(z,p);  C[r*NHS + s] = d h_r / d z_HSUP[s];  H(packed NZ) += sum_r mh[r] d2 h_r / dz2")
    w("template <class PT, class CM>      // C: plain array or any type with operator[] (e.g. a shared-memory column accessor)")
    w("__device__ __forceinline__ void con_lin(const double* z, const PT p, const double* mh, double* H, double* hv, CM&& C)\n{")
    if nh:
        if flat:
            m3 = dict(m)
            m3.update({mhs[i]: "mh[%d]" % i for i in range(nh)})
            w(emit_block([("hv[%d]" % i, h[i]) for i in flat] +
                         [("C[%d]" % (i * nhs + s_), sp.diff(h[i], z[sup[s_]])) for i in flat for s_ in range(nhs)] +
                         hess_outputs(flat, mhs), m3))
        for gi, (r0, cnt, stride, fixed) in enumerate(groups):
            mq = qmap(r0, fixed)
            mq[mhs[r0]] = "mh[%d + r]" % r0
            w("#pragma unroll 1\n    for (int r = 0; r < %d; r++) {\n        const PT q = p + r * %d;" % (cnt, stride))
            w(emit_block([("hv[%d + r]" % r0, h[r0])] +
                         [("C[(%d + r) * %d + %d]" % (r0, nhs, s_), sp.diff(h[r0], z[sup[s_]])) for s_ in range(nhs)] +
                         hess_outputs([r0], mhs), mq, tmp_prefix="l%d_" % gi, indent="        "))
            w("    }")
    w("}")
    w("\n// the same restricted to the rows [r_lo, r_hi): hv, C, mh indexed by (row - r_lo)")
    w("template <class PT>\n__device__ __forceinline__ void con_lin_rows(const double* z, const PT p, int r_lo, int r_hi, const double* mh, double* H, double* hv, double* C)\n{")
    if nh:
        for i in flat:
            m3 = dict(m)
            m3[mhs[i]] = "mh[o]"
            w("    if (r_lo <= %d && %d < r_hi) {\n        const int o = %d - r_lo;" % (i, i, i))
            w(emit_block([("hv[o]", h[i])] + [("C[o * %d + %d]" % (nhs, s_), sp.diff(h[i], z[sup[s_]])) for s_ in range(nhs)] +
                         hess_outputs([i], mhs), m3, tmp_prefix="n%d_" % i, indent="        "))
            w("    }")
        for gi, (r0, cnt, stride, fixed) in enumerate(groups):
            mq = qmap(r0, fixed)
            mq[mhs[r0]] = "mh[o]"
            w("#pragma unroll 1\n    for (int r = (r_lo > %d ? r_lo - %d : 0); r < (r_hi - %d < %d ? r_hi - %d : %d); r++) {" % (r0, r0, r0, cnt, r0, cnt))
            w("        const PT q = p + r * %d;\n        const int o = %d + r - r_lo;" % (stride, r0))
            w(emit_block([("hv[o]", h[r0])] + [("C[o * %d + %d]" % (nhs, s_), sp.diff(h[r0], z[sup[s_]])) for s_ in range(nhs)] +
                         hess_outputs([r0], mhs), mq, tmp_prefix="m%d_" % gi, indent="        "))
            w("    }")
    w("}")

    w("}  // namespace mpcgen")
    return "\n".join(out) + "\n"


def _bundle_function_name(key):
    return "setSolverParameter" + key.replace("_", " ").title().replace(" ", "")


def emit_parameter_header(pb):
    """setSolverParameter<Bundle>(k, params, value, index): generate_cpp_files.py:204-260.
    A constant index table replaces the reference's if-chain."""
    npar = len(pb["p"])
    out = ["// GENERATED -- mirrors mpc_planner_solver/include/mpc_planner_solver/mpc_planner_parameters.h",
           "#pragma once", "namespace MPCPlanner {", "struct AcadosParameters;"]
    for key, idxs in pb["bundles"].items():
        fn = _bundle_function_name(key)
        tbl = ", ".join(str(i) for i in idxs)
        out.append("inline void %s(int k, AcadosParameters& params, const double value, int index = 0)" % fn)
        out.append("{ static const int idx[%d] = {%s}; mpcgpu_set_parameter(params, k * %d + idx[index], value); }"
                   % (len(idxs), tbl, npar))
    out.append("}  // namespace MPCPlanner")
    return "\n".join(out) + "\n"


# ---- struct-of-tables parameter path --------------------------------------------------------------------
# Which parameters the reference's C++ modules write with the SAME value at every stage (weights from CONFIG:
# mpc_base / contouring / goal setParameters; the spline segments: contouring.cpp:96-126; disc radius / offsets) and which are
# per-stage tables (obstacle slots: ellipsoid_constraints.cpp:34-90, gaussian_constraints.cpp:31-79; halfspaces:
# linearized_constraints.cpp:150-189, decomp_constraints.cpp:150-189; the consistency reference: guidance_constraints.cpp:985-1023).
_STAGE_PREFIXES = ("ellipsoid_obst_", "gaussian_obst_", "lin_constraint_", "disc_", "consistency_weight", "prev_traj_")


def classify_parameters(pb):
    names = pb["param_names"]
    per_stage = [i for i, n in enumerate(names) if n.startswith(_STAGE_PREFIXES)]
    invariant = [i for i in range(len(names)) if i not in set(per_stage)]
    pm = {n: i for i, n in enumerate(names)}
    M = sum(1 for n in names if n.startswith("ellipsoid_obst_") and n.endswith("_x"))
    ell = None
    if M:
        base = pm["ellipsoid_obst_0_x"]
        stride = (pm["ellipsoid_obst_1_x"] - base) if M > 1 else 7
        base = min(pm["ellipsoid_obst_0_" + f] for f in ("x", "y", "psi", "major", "minor", "chi", "r"))
        ell = dict(count=M, base=base, stride=stride, offsets=[pm["ellipsoid_obst_0_" + f] - base for f in ("x", "y", "psi", "major", "minor", "chi", "r")])
    nlin = sum(1 for n in names if n.startswith("lin_constraint_") and n.endswith("_a1"))
    lin = dict(count=nlin, base=pm["lin_constraint_0_a1"]) if nlin else None
    return dict(invariant=invariant, per_stage=per_stage, ellipsoid=ell, guidance_halfspaces=lin)


def emit_tables_header(pb, name):
    """mpc_planner_tables.h: the struct-of-tables a caller fills ONCE per control cycle instead of N x npar setParameter calls"""
    c = classify_parameters(pb)
    names = pb["param_names"]
    out = ["// GENERATED -- struct-of-tables parameter path of configuration %s (SURVEY 8 f2): stage-invariant parameters travel" % name,
           "// once per homotopy set, obstacle predictions as a table; the engine expands them to all_parameters on the device",
           "// (mpcgpu_solve_sets_tables, include/mpcgpu.h).  Replaces: the k-loop over modules->setParameters (mpc_planner/src/planner.cpp:153-159)",
           "// through the generated setSolverParameter<Bundle> if-chains (solver_generator/generate_cpp_files.py:235-254).",
           "#pragma once", "namespace MPCPlanner {", "struct SolverTables {",
           "    static constexpr int n_invariant = %d;" % len(c["invariant"]),
           "    // flat parameter indices, in the order of `invariant` below",
           "    static constexpr int invariant_idx[%d] = {%s};" % (max(len(c["invariant"]), 1), ", ".join(str(i) for i in c["invariant"]) or "0"),
           "    double invariant[%d] = {};   // %s" % (max(len(c["invariant"]), 1), ", ".join(names[i] for i in c["invariant"][:8]) + (", ..." if len(c["invariant"]) > 8 else ""))]
    for pos, i in enumerate(c["invariant"]):
        out.append("    static constexpr int inv_%s = %d;" % (names[i], pos))
    if c["ellipsoid"]:
        e = c["ellipsoid"]
        out += ["    static constexpr int max_obstacles = %d, ell_base = %d, ell_stride = %d;" % (e["count"], e["base"], e["stride"]),
                "    static constexpr int ell_offsets[7] = {%s};   // x, y, psi, major, minor, chi, r inside an obstacle's block" % ", ".join(str(v) for v in e["offsets"])]
    if c["guidance_halfspaces"]:
        out.append("    static constexpr int lin_base = %d, lin_count = %d;" % (c["guidance_halfspaces"]["base"], c["guidance_halfspaces"]["count"]))
    out += ["};", "}  // namespace MPCPlanner"]
    return "\n".join(out) + "\n"


def generate_cuda_solver(modules, settings, model, name, out_dir):
    pb = extract_problem(modules, settings, model)
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "model.cuh"), "w") as f:
        f.write(emit_model_header(pb, name))
    with open(os.path.join(out_dir, "mpc_planner_parameters.h"), "w") as f:
        f.write(emit_parameter_header(pb))
    with open(os.path.join(out_dir, "solver_dims.h"), "w") as f:      # the SOLVER_* macros of acados_solver_Solver.h
        f.write("// GENERATED -- dimensions the reference takes from Solver/acados_solver_Solver.h\n#pragma once\n"
                "#define SOLVER_N %d\n#define SOLVER_NX %d\n#define SOLVER_NU %d\n#define SOLVER_NP %d\n#define SOLVER_NH %d\n"
                "#define MPCGPU_CONFIG_NAME \"%s\"\n" % (pb["N"], pb["nx"], pb["nu"], len(pb["p"]), len(pb["h"]), name))
    pmap = {n: i for i, n in enumerate(pb["param_names"])}
    pmap["num parameters"] = len(pb["param_names"])       # util/parameters.py:72
    with open(os.path.join(out_dir, "parameter_map.yaml"), "w") as f:
        yaml.dump(pmap, f, default_flow_style=False)
    mmap = {}
    nu = pb["nu"]
    for i, s in enumerate(pb["states"]):
        mmap[s] = ["x", i + nu, pb["lb"][nu + i], pb["ub"][nu + i]]
    for i, s in enumerate(pb["inputs"]):
        mmap[s] = ["u", i, pb["lb"][i], pb["ub"][i]]
    with open(os.path.join(out_dir, "model_map.yaml"), "w") as f:
        yaml.dump(mmap, f, default_flow_style=False)
    with open(os.path.join(out_dir, "mpc_planner_tables.h"), "w") as f:
        f.write(emit_tables_header(pb, name))
    cls = classify_parameters(pb)
    with open(os.path.join(out_dir, "tables.yaml"), "w") as f:
        yaml.dump(dict(invariant=[pb["param_names"][i] for i in cls["invariant"]], per_stage=[pb["param_names"][i] for i in cls["per_stage"]],
                       ellipsoid=cls["ellipsoid"], guidance_halfspaces=cls["guidance_halfspaces"]), f, default_flow_style=None)
    with open(os.path.join(out_dir, "solver_settings.yaml"), "w") as f:
        yaml.dump(dict(N=pb["N"], nx=pb["nx"], nu=pb["nu"], nvar=pb["nx"] + pb["nu"], npar=len(pb["p"])), f,
                  default_flow_style=False)
    return pb
