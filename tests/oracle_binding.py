"""ctypes binding of the CPU oracle (oracle/_build/liboracle_<config>.so).  TEST INFRASTRUCTURE:
imported only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class Oracle:
    _cache = {}

    def __init__(self, config, variant=""):
        """variant "balance_": the -DMPC_HPIPM_BALANCE build (conditional predictor-corrector)"""
        libname = "liboracle_%s%s.so" % (variant, config)
        path = os.path.join(ORACLE_DIR, "_build", libname)
        if not os.path.exists(path):
            subprocess.run(["make", "-C", ORACLE_DIR, "_build/" + libname], check=True,
                           stdout=subprocess.DEVNULL)
        config_key = variant + config
        if config_key not in Oracle._cache:
            lib = ctypes.CDLL(path)
            lib.oracle_param_name.restype = ctypes.c_char_p
            lib.oracle_var_name.restype = ctypes.c_char_p
            lib.oracle_bounds.restype = ctypes.POINTER(ctypes.c_double)
            Oracle._cache[config_key] = lib
        self.lib = Oracle._cache[config_key]
        d = [ctypes.c_int() for _ in range(5)]
        self.lib.oracle_dims(*[ctypes.byref(v) for v in d])
        self.N, self.nx, self.nu, self.npar, self.nh = [v.value for v in d]
        self.nz = self.nx + self.nu
        self.nc = self.lib.oracle_nc()
        self.mem_doubles = self.lib.oracle_mem_doubles()
        self.parameter_map = {self.lib.oracle_param_name(i).decode(): i for i in range(self.npar)}
        self.var_names = [self.lib.oracle_var_name(i).decode() for i in range(self.nz)]
        self.dims = dict(N=self.N, nx=self.nx, nu=self.nu, npar=self.npar, dt=0.2)

    def bounds(self, which):
        n = self.nz if which < 2 else max(self.nh, 1)
        return np.array([self.lib.oracle_bounds(which)[i] for i in range(n)])

    def solve_batch(self, xinit, x0, params, num_iter=10, mem=None, threads=None):
        n = xinit.shape[0]
        xinit = np.ascontiguousarray(xinit, np.float64)
        x0 = np.ascontiguousarray(x0, np.float64)
        params = np.ascontiguousarray(params, np.float64)
        out = dict(xtraj=np.zeros((n, (self.N + 1) * self.nx)), utraj=np.zeros((n, self.N * self.nu)), pobj=np.zeros(n),
                   exit_code=np.zeros(n, np.int32), qp_status=np.zeros(n, np.int32), res_eq=np.zeros(n),
                   ipm_iters=np.zeros(n, np.int32))
        ni = np.full(n, num_iter, np.int32) if np.ndim(num_iter) == 0 else np.ascontiguousarray(num_iter, np.int32)
        if threads is None:
            threads = min(8, os.cpu_count() or 1)
        self.lib.oracle_solve_batch(n, _ptr(xinit), _ptr(x0), _ptr(params), _ptr(ni), _ptr(mem), _ptr(out["xtraj"]),
                                    _ptr(out["utraj"]), _ptr(out["pobj"]), _ptr(out["exit_code"]), _ptr(out["qp_status"]),
                                    _ptr(out["res_eq"]), _ptr(out["ipm_iters"]), int(threads))
        return out

    def guidance_halfspaces(self, n_sets, planners, xinit_sets, x0, obst_pred, guided, robot_radius, lin_base, lin_count, params,
                            static_halfspaces=None):
        """oracle_guidance_halfspaces[_static]: writes the halfspace slots of `params` [n, N*npar] in place;
        static_halfspaces [n_sets, N, n_static, 3] = module_data.static_obstacles rows (a1, a2, b)"""
        obst_pred = np.ascontiguousarray(obst_pred, np.float64)
        n_obs = obst_pred.shape[2]
        assert params.flags["C_CONTIGUOUS"] and params.dtype == np.float64
        st = None if static_halfspaces is None else np.ascontiguousarray(static_halfspaces, np.float64)
        fn = self.lib.oracle_guidance_halfspaces_static
        fn.argtypes = [ctypes.c_int] * 9 + [ctypes.c_void_p] * 4 + [ctypes.c_double, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        fn.restype = None
        fn(n_sets, planners, self.N, self.nx, self.nu, self.npar, lin_base, lin_count, n_obs,
           _ptr(np.ascontiguousarray(xinit_sets, np.float64)), _ptr(np.ascontiguousarray(x0, np.float64)),
           _ptr(obst_pred), _ptr(np.ascontiguousarray(guided, np.uint8)), float(robot_radius), _ptr(st),
           0 if st is None else int(st.shape[2]), _ptr(params))
        return params

    def select_best_cons(self, set_offsets, pobj, exit_code, xtraj, prev_traj, cons_weight, cons_enabled=None, obj_scale=None,
                         obj_sub=None, disabled=None, ix=0, iy=1):
        """objective post-processing with the consistency cost of the SOLVED trajectory + FindBestPlanner
        (guidance_constraints.cpp:373-420,572-590,1025-1050); returns (best, objective, consistency_cost)"""
        set_offsets = np.ascontiguousarray(set_offsets, np.int32)
        n_sets, n = set_offsets.size - 1, int(set_offsets[-1])
        best = np.zeros(n_sets, np.int32)
        obj, cons = np.zeros(n), np.zeros(n)
        f64 = lambda a: None if a is None else np.ascontiguousarray(a, np.float64)
        u8 = lambda a: None if a is None else np.ascontiguousarray(a, np.uint8)
        sc, sb, ds, en, xt, pv = f64(obj_scale), f64(obj_sub), u8(disabled), u8(cons_enabled), f64(xtraj), f64(prev_traj)
        fn = self.lib.oracle_select_best_cons
        vp = ctypes.c_void_p
        fn.argtypes = [ctypes.c_int] + [vp] * 10 + [ctypes.c_double] + [ctypes.c_int] * 4 + [vp, vp]
        fn(n_sets, _ptr(set_offsets), _ptr(f64(pobj)), _ptr(np.ascontiguousarray(exit_code, np.int32)), _ptr(sc), _ptr(sb), _ptr(ds),
           _ptr(best), _ptr(xt), _ptr(pv), _ptr(en), float(cons_weight), self.N, self.nx, ix, iy, _ptr(obj), _ptr(cons))
        return best, obj, cons

    def select_best(self, set_offsets, pobj, exit_code, obj_scale=None, obj_sub=None, disabled=None):
        set_offsets = np.ascontiguousarray(set_offsets, np.int32)
        n_sets = set_offsets.size - 1
        best = np.zeros(n_sets, np.int32)
        sc = None if obj_scale is None else np.ascontiguousarray(obj_scale, np.float64)
        sb = None if obj_sub is None else np.ascontiguousarray(obj_sub, np.float64)
        ds = None if disabled is None else np.ascontiguousarray(disabled, np.uint8)
        self.lib.oracle_select_best(n_sets, _ptr(set_offsets), _ptr(np.ascontiguousarray(pobj, np.float64)),
                                    _ptr(np.ascontiguousarray(exit_code, np.int32)), _ptr(sc), _ptr(sb), _ptr(ds), _ptr(best))
        return best
