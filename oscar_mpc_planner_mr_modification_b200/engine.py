"""ctypes binding of libmpcgpu.so (include/mpcgpu.h) -- host-side plumbing for tests and bench.

There is NO CPU fallback: if the shared library or a CUDA device is missing, construction raises.
"""
import ctypes
import os

import numpy as np
import yaml

_PKG = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("MPCGPU_LIB") or os.path.join(_PKG, "lib", "libmpcgpu.so")
_lib = None

SYMBOLS = [
    "mpcgpu_num_configs", "mpcgpu_config_name", "mpcgpu_engine_create", "mpcgpu_engine_destroy", "mpcgpu_desc_query",
    "mpcgpu_mem_doubles", "mpcgpu_solve_batch", "mpcgpu_solve_batch_device", "mpcgpu_sync", "mpcgpu_solve_sets", "mpcgpu_select_best",
    "mpcgpu_alloc_pinned", "mpcgpu_free_pinned", "mpcgpu_solve_sets_tables", "mpcgpu_multi_create", "mpcgpu_multi_destroy", "mpcgpu_multi_num_devices", "mpcgpu_multi_engine",
    "mpcgpu_multi_shard_range", "mpcgpu_multi_solve_sets", "mpcgpu_multi_solve_sets_guided", "mpcgpu_multi_solve_sets_tables", "mpcgpu_multi_solve_batch", "mpcgpu_multi_last_kernel_ms",
    "mpcgpu_generate_synthetic", "mpcgpu_generate_synthetic_device",
    "mpcgpu_select_best_device", "mpcgpu_model_eval_doubles", "mpcgpu_model_eval", "mpcgpu_measure_fp64_peak", "mpcgpu_launch_count", "mpcgpu_last_kernel_ms", "mpcgpu_last_error", "mpcgpu_set_kernel_mode", "mpcgpu_guidance_halfspaces_device", "mpcgpu_solve_sets_guided",
]


KERNEL_AUTO, KERNEL_STAGE, KERNEL_SPLIT = 0, 1, 2      # mpcgpu_set_kernel_mode (include/mpcgpu.h)


class SetOptions(ctypes.Structure):
    """struct mpcgpu_set_options (include/mpcgpu.h): optional arguments of the homotopy-set entries"""
    _fields_ = [("consistency_weight", ctypes.c_double), ("prev_traj", ctypes.c_void_p), ("consistency_enabled", ctypes.c_void_p),
                ("ix", ctypes.c_int), ("iy", ctypes.c_int), ("mem_inout", ctypes.c_void_p), ("objective_out", ctypes.c_void_p),
                ("consistency_cost_out", ctypes.c_void_p), ("static_halfspaces", ctypes.c_void_p), ("n_static", ctypes.c_int),
                ("best_xtraj", ctypes.c_void_p), ("best_utraj", ctypes.c_void_p)]


def load_library():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError("libmpcgpu.so is not built (%s); run `python __graft_entry__.py`. "
                               "There is no CPU fallback for the solve path." % _LIB_PATH)
        lib = ctypes.CDLL(_LIB_PATH)
        lib.mpcgpu_config_name.restype = ctypes.c_char_p
        lib.mpcgpu_last_error.restype = ctypes.c_char_p
        lib.mpcgpu_last_error.argtypes = [ctypes.c_void_p]
        lib.mpcgpu_launch_count.restype = ctypes.c_longlong
        lib.mpcgpu_launch_count.argtypes = [ctypes.c_void_p]
        lib.mpcgpu_last_kernel_ms.restype = ctypes.c_float
        lib.mpcgpu_last_kernel_ms.argtypes = [ctypes.c_void_p]
        lib.mpcgpu_engine_create.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
        lib.mpcgpu_engine_destroy.argtypes = [ctypes.c_void_p]
        lib.mpcgpu_desc_query.argtypes = [ctypes.c_void_p] + [ctypes.POINTER(ctypes.c_int)] * 5
        lib.mpcgpu_mem_doubles.argtypes = [ctypes.c_void_p]
        lib.mpcgpu_sync.argtypes = [ctypes.c_void_p]
        vp = ctypes.c_void_p
        lib.mpcgpu_solve_batch.argtypes = [vp, ctypes.c_int, vp, vp, vp, vp, ctypes.c_int, vp, vp, vp, vp, vp, vp, vp, vp]
        lib.mpcgpu_solve_batch_device.argtypes = [vp, ctypes.c_int, vp, vp, vp, vp, ctypes.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        lib.mpcgpu_select_best.argtypes = [vp, ctypes.c_int, vp, vp, vp, vp, vp, vp, vp]
        lib.mpcgpu_select_best_device.argtypes = [vp, ctypes.c_int, vp, vp, vp, vp, vp, vp, vp, vp]
        _lib = lib
    return _lib


def config_dir(config):
    return os.path.join(_PKG, "generated", config)


def load_maps(config):
    d = config_dir(config)
    with open(os.path.join(d, "parameter_map.yaml")) as f:
        pmap = yaml.safe_load(f)
    with open(os.path.join(d, "model_map.yaml")) as f:
        mmap = yaml.safe_load(f)
    with open(os.path.join(d, "solver_settings.yaml")) as f:
        st = yaml.safe_load(f)
    pmap = {k: v for k, v in pmap.items() if k != "num parameters"}
    return pmap, mmap, st


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(ctypes.c_void_p)
    return ctypes.c_void_p(int(a))   # raw device address (e.g. torch.Tensor.data_ptr())


def measure_fp64_peak(device=0):
    lib = load_library()
    v = ctypes.c_double()
    rc = lib.mpcgpu_measure_fp64_peak(int(device), ctypes.byref(v))
    if rc != 0:
        raise RuntimeError("mpcgpu_measure_fp64_peak failed: %d" % rc)
    return v.value


class PinnedArray:
    """numpy view of page-locked host memory from mpcgpu_alloc_pinned (freed with the object).  Copies from pinned arrays run
    asynchronously; mpcgpu_solve_batch then takes its gated single-launch pipeline and writes results straight into pinned
    output arrays."""

    def __init__(self, shape, dtype=np.float64):
        lib = load_library()
        lib.mpcgpu_alloc_pinned.argtypes = [ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p)]
        lib.mpcgpu_free_pinned.argtypes = [ctypes.c_void_p]
        self._lib = lib
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self._ptr = ctypes.c_void_p()
        rc = lib.mpcgpu_alloc_pinned(max(nbytes, 8), ctypes.byref(self._ptr))
        if rc != 0:
            raise RuntimeError("mpcgpu_alloc_pinned(%d bytes) failed: %d" % (nbytes, rc))
        buf = (ctypes.c_char * max(nbytes, 8)).from_address(self._ptr.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def __del__(self):
        try:
            if self._ptr:
                self.array = None
                self._lib.mpcgpu_free_pinned(self._ptr)
                self._ptr = None
        except Exception:
            pass


def pinned_copy(a):
    """(holder, array): a copy of `a` in pinned host memory; keep the holder alive while the array is in use"""
    h = PinnedArray(a.shape, a.dtype)
    h.array[...] = a
    return h, h.array


class ParamTables(ctypes.Structure):
    """struct mpcgpu_param_tables (include/mpcgpu.h): the struct-of-tables parameter path"""
    _fields_ = [("n_invariant", ctypes.c_int), ("invariant_idx", ctypes.c_void_p), ("invariant", ctypes.c_void_p),
                ("n_stage", ctypes.c_int), ("stage_idx", ctypes.c_void_p), ("stage", ctypes.c_void_p),
                ("M", ctypes.c_int), ("ob_stride", ctypes.c_int), ("obstacles", ctypes.c_void_p), ("obstacle_radius", ctypes.c_void_p),
                ("ell_base", ctypes.c_int), ("ell_stride", ctypes.c_int), ("ell_offsets", ctypes.c_void_p),
                ("guided", ctypes.c_void_p), ("lin_base", ctypes.c_int), ("lin_count", ctypes.c_int), ("robot_radius", ctypes.c_double)]


class MpcGpuError(RuntimeError):
    pass


class SynthLayout(ctypes.Structure):
    """struct mpcgpu_synth_layout (include/mpcgpu.h): parameter indices + constants of the counter-based synthetic generator"""
    _fields_ = [("N", ctypes.c_int), ("nx", ctypes.c_int), ("nu", ctypes.c_int), ("npar", ctypes.c_int), ("guided", ctypes.c_int),
                ("weights", ctypes.c_int * 9), ("spline", (ctypes.c_int * 9) * 5), ("ego_disc_radius", ctypes.c_int),
                ("ego_disc_0_offset", ctypes.c_int), ("goal", ctypes.c_int * 3), ("prev_traj_x", ctypes.c_int), ("prev_traj_y", ctypes.c_int),
                ("n_obst", ctypes.c_int), ("obst", (ctypes.c_int * 7) * 16), ("n_lin", ctypes.c_int), ("lin", (ctypes.c_int * 3) * 16),
                ("dt", ctypes.c_double), ("pi", ctypes.c_double), ("vg", ctypes.c_double), ("need", ctypes.c_double),
                ("deceleration", ctypes.c_double), ("robot_radius", ctypes.c_double), ("obstacle_radius", ctypes.c_double),
                ("lin_margin", ctypes.c_double), ("weight_values", ctypes.c_double * 9), ("lateral", ctypes.c_double * 8),
                ("lat_profile", ctypes.c_double * 64)]

    @classmethod
    def from_dict(cls, d):
        """d: synthetic.synth_layout(parameter_map, dims)"""
        L = cls()
        for name, _ in cls._fields_:
            v = d[name]
            if name in ("spline", "obst", "lin"):
                for i, row in enumerate(v):
                    for j, x in enumerate(row):
                        getattr(L, name)[i][j] = x
                for i in range(len(v), len(getattr(L, name))):
                    for j in range(len(getattr(L, name)[i])):
                        getattr(L, name)[i][j] = -1
            elif isinstance(v, (list, tuple)):
                for i, x in enumerate(v):
                    getattr(L, name)[i] = x
            else:
                setattr(L, name, v)
        return L


def generate_synthetic(parameter_map, dims, n_sets, planners, seed=1234, first_set=0, device=0, guided=None, device_buffers=None,
                       stream=None):
    """Synthetic homotopy sets first_set .. first_set + n_sets - 1 of the global batch of `seed`, generated ON THE DEVICE
    (mpcgpu_generate_synthetic[_device]; host mirror: synthetic.make_batch_philox).
    device_buffers = (xinit, x0, params, obst_pred or None) raw device addresses: filled in place, asynchronously on `stream`,
    returns None.  Otherwise the data are copied back and returned as a make_batch-style dict of numpy arrays."""
    from . import synthetic
    lib = load_library()
    ld = synthetic.synth_layout(parameter_map, dims, guided)
    L = SynthLayout.from_dict(ld)
    vp, ull, ll = ctypes.c_void_p, ctypes.c_ulonglong, ctypes.c_longlong
    if device_buffers is not None:
        lib.mpcgpu_generate_synthetic_device.argtypes = [ctypes.c_int, vp, ull, ll, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, vp]
        xi, x0, pa, ob = device_buffers
        rc = lib.mpcgpu_generate_synthetic_device(int(device), ctypes.byref(L), int(seed), int(first_set), int(n_sets), int(planners),
                                                  _ptr(xi), _ptr(x0), _ptr(pa), _ptr(ob), _ptr(stream) if stream else None)
        if rc != 0:
            raise MpcGpuError("mpcgpu_generate_synthetic_device failed: status %d" % rc)
        return None
    N, nx, nu, npar, M = ld["N"], ld["nx"], ld["nu"], ld["npar"], ld["n_obst"]
    B = n_sets * planners
    out = dict(xinit=np.empty((B, nx)), x0=np.empty((B, (N + 1) * (nx + nu))), params=np.empty((B, N * npar)),
               obst_pred=np.empty((n_sets, N, M if ld["n_lin"] else 0, 2)), set_offsets=np.arange(0, B + 1, planners, dtype=np.int32), n=B,
               robot_radius=ld["robot_radius"])
    follow = [bool(ld["guided"]) and not (planners > 1 and h == planners - 1) for h in range(planners)]
    out["guided"] = np.ascontiguousarray(np.tile(np.array([1 if f else 0 for f in follow], np.uint8), n_sets))
    lib.mpcgpu_generate_synthetic.argtypes = [ctypes.c_int, vp, ull, ll, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp]
    rc = lib.mpcgpu_generate_synthetic(int(device), ctypes.byref(L), int(seed), int(first_set), int(n_sets), int(planners),
                                       _ptr(out["xinit"]), _ptr(out["x0"]), _ptr(out["params"]),
                                       _ptr(out["obst_pred"]) if out["obst_pred"].size else None)
    if rc != 0:
        raise MpcGpuError("mpcgpu_generate_synthetic failed: status %d" % rc)
    return out


class Engine:
    """One engine = one problem configuration on one GPU."""

    def __init__(self, config, device=0, max_batch=4096):
        self.lib = load_library()
        self.config = config
        self.parameter_map, self.model_map, st = load_maps(config)
        self.handle = ctypes.c_void_p()
        rc = self.lib.mpcgpu_engine_create(config.encode(), device, max_batch, ctypes.byref(self.handle))
        if rc != 0:
            msg = self.lib.mpcgpu_last_error(self.handle).decode() if self.handle else ""
            if self.handle:
                self.lib.mpcgpu_engine_destroy(self.handle)
                self.handle = ctypes.c_void_p()
            raise MpcGpuError("mpcgpu_engine_create(%s, device %d) failed: status %d %s" % (config, device, rc, msg))
        d = [ctypes.c_int() for _ in range(5)]
        self.lib.mpcgpu_desc_query(self.handle, *[ctypes.byref(v) for v in d])
        self.N, self.nx, self.nu, self.npar, self.nh = [v.value for v in d]
        assert (self.N, self.nx, self.nu, self.npar) == (st["N"], st["nx"], st["nu"], st["npar"])
        self.nz = self.nx + self.nu
        self.max_batch = max_batch
        self.mem_doubles = self.lib.mpcgpu_mem_doubles(self.handle)
        self.dims = dict(N=self.N, nx=self.nx, nu=self.nu, npar=self.npar, dt=0.2)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.mpcgpu_engine_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise MpcGpuError("%s failed: status %d %s" % (what, rc, self.lib.mpcgpu_last_error(self.handle).decode()))

    def alloc_outputs(self, n):
        return dict(xtraj=np.zeros((n, (self.N + 1) * self.nx)), utraj=np.zeros((n, self.N * self.nu)), pobj=np.zeros(n),
                    exit_code=np.zeros(n, np.int32), qp_status=np.zeros(n, np.int32), res_eq=np.zeros(n),
                    ipm_iters=np.zeros(n, np.int32))

    def solve_batch(self, xinit, x0, params, num_iter=10, mem=None, out=None):
        """HOST numpy arrays in/out (H2D + kernel + D2H inside the call)."""
        n = xinit.shape[0]
        xinit = np.ascontiguousarray(xinit, np.float64)
        x0 = np.ascontiguousarray(x0, np.float64)
        params = np.ascontiguousarray(params, np.float64)
        assert x0.size == n * self.nz * (self.N + 1) and params.size == n * self.N * self.npar and xinit.size == n * self.nx
        if out is None:
            out = self.alloc_outputs(n)
        ni = None
        nall = 0
        if np.ndim(num_iter) == 0:
            nall = int(num_iter)
        else:
            ni = np.ascontiguousarray(num_iter, np.int32)
        if mem is not None:
            assert mem.dtype == np.float64 and mem.size == n * self.mem_doubles and mem.flags["C_CONTIGUOUS"]
        rc = self.lib.mpcgpu_solve_batch(self.handle, n, _ptr(xinit), _ptr(x0), _ptr(params), _ptr(ni), nall, _ptr(mem),
                                         _ptr(out["xtraj"]), _ptr(out["utraj"]), _ptr(out["pobj"]), _ptr(out["exit_code"]),
                                         _ptr(out["qp_status"]), _ptr(out["res_eq"]), _ptr(out["ipm_iters"]))
        self._check(rc, "mpcgpu_solve_batch")
        return out

    def solve_batch_device(self, n, xinit, x0, params, num_iter, xtraj, utraj, pobj, exit_code, qp_status, res_eq,
                           ipm_iters=None, mem=None, num_iter_dev=None, stream=None):
        """Raw device addresses (ints); asynchronous."""
        rc = self.lib.mpcgpu_solve_batch_device(self.handle, n, _ptr(xinit), _ptr(x0), _ptr(params), _ptr(num_iter_dev),
                                                int(num_iter), _ptr(mem), _ptr(xtraj), _ptr(utraj), _ptr(pobj), _ptr(exit_code),
                                                _ptr(qp_status), _ptr(res_eq), _ptr(ipm_iters), _ptr(stream))
        self._check(rc, "mpcgpu_solve_batch_device")

    def _set_options(self, out, n, prev_traj=None, cons_weight=0.0, cons_enabled=None, mem=None, static_halfspaces=None, n_sets=None,
                     best_only=False):
        """(ctypes pointer or None, keep-alive list) for struct mpcgpu_set_options; adds out["objective"], out["consistency_cost"]
        and, with best_only, out["best_xtraj"] / out["best_utraj"] (the per-planner trajectories then stay on the device)"""
        if prev_traj is None and mem is None and static_halfspaces is None and not best_only:
            return None, []
        keep = []
        o = SetOptions()
        o.consistency_weight = float(cons_weight)
        o.ix, o.iy = self.model_map["x"][1] - self.nu, self.model_map["y"][1] - self.nu
        if prev_traj is not None:
            pv = np.ascontiguousarray(prev_traj, np.float64); keep.append(pv)
            o.prev_traj = pv.ctypes.data
            if cons_enabled is not None:
                en = np.ascontiguousarray(cons_enabled, np.uint8); keep.append(en)
                assert en.size == n
                o.consistency_enabled = en.ctypes.data
        out["objective"] = np.zeros(n); out["consistency_cost"] = np.zeros(n)
        o.objective_out = out["objective"].ctypes.data
        o.consistency_cost_out = out["consistency_cost"].ctypes.data
        if mem is not None:
            assert mem.dtype == np.float64 and mem.size == n * self.mem_doubles and mem.flags["C_CONTIGUOUS"]
            o.mem_inout = mem.ctypes.data
        if static_halfspaces is not None:
            st = np.ascontiguousarray(static_halfspaces, np.float64); keep.append(st)
            o.static_halfspaces = st.ctypes.data
            o.n_static = int(st.shape[2])
        if best_only:
            out["best_xtraj"] = np.zeros((n_sets, self.nx * (self.N + 1))); out["best_utraj"] = np.zeros((n_sets, self.nu * self.N))
            o.best_xtraj = out["best_xtraj"].ctypes.data
            o.best_utraj = out["best_utraj"].ctypes.data
        keep.append(o)
        return ctypes.byref(o), keep

    def solve_sets(self, n_sets, planners, xinit_sets, shared_params, x0, param_idx, planner_params, num_iter=10, out=None,
                   obj_scale=None, disabled=None, **opts):
        """Compact homotopy-set entry: shared parameter block per set + per-planner overrides; returns the
        per-problem outputs and `best` (selected planner per set).  opts: prev_traj, cons_weight, cons_enabled, mem
        (struct mpcgpu_set_options)."""
        n = n_sets * planners
        if out is None:
            out = self.alloc_outputs(n)
            out["best"] = np.zeros(n_sets, np.int32)
        opt, _keep = self._set_options(out, n, n_sets=n_sets, **opts)
        best_only = bool(opts.get("best_only"))
        sc = None if obj_scale is None else np.ascontiguousarray(obj_scale, np.float64)
        ds = None if disabled is None else np.ascontiguousarray(disabled, np.uint8)
        idx = np.ascontiguousarray(param_idx, np.int32)
        vp = ctypes.c_void_p
        self.lib.mpcgpu_solve_sets.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, vp, vp, ctypes.c_int, vp, vp, vp, ctypes.c_int] + [vp] * 11
        xs, sh, x0_, pv = (np.ascontiguousarray(a, np.float64) for a in (xinit_sets, shared_params, x0, planner_params))
        assert sh.size == n_sets * self.N * self.npar and x0_.size == n * self.nz * (self.N + 1) and pv.size == n * self.N * idx.size
        rc = self.lib.mpcgpu_solve_sets(self.handle, n_sets, planners, _ptr(xs), _ptr(sh), _ptr(x0_), int(idx.size), _ptr(idx), _ptr(pv), None,
                                        int(num_iter), None if best_only else _ptr(out["xtraj"]), None if best_only else _ptr(out["utraj"]),
                                        _ptr(out["pobj"]), _ptr(out["exit_code"]),
                                        _ptr(out["qp_status"]), _ptr(out["res_eq"]), _ptr(sc), None, _ptr(ds), _ptr(out["best"]), opt)
        self._check(rc, "mpcgpu_solve_sets")
        return out

    def lin_constraint_block(self):
        """(lin_base, lin_count): the guidance halfspace parameters lin_constraint_<j>_{a1,a2,b} must be 3 consecutive
        slots per constraint (they are: the generator adds them in that order, guidance_constraints.py / linearized_constraints.py)"""
        pm = self.parameter_map
        cnt = sum(1 for k in pm if k.startswith("lin_constraint_") and k.endswith("_a1"))
        if cnt == 0:
            return 0, 0
        base = pm["lin_constraint_0_a1"]
        for j in range(cnt):
            assert (pm["lin_constraint_%d_a1" % j], pm["lin_constraint_%d_a2" % j], pm["lin_constraint_%d_b" % j]) == (
                base + 3 * j, base + 3 * j + 1, base + 3 * j + 2), "unexpected halfspace parameter layout"
        return base, cnt

    def solve_sets_guided(self, n_sets, planners, xinit_sets, shared_params, x0, obst_pred, guided, robot_radius, num_iter=10,
                          param_idx=None, planner_params=None, out=None, obj_scale=None, disabled=None, **opts):
        """mpcgpu_solve_sets_guided: halfspaces from obstacle predictions + warm starts, built on the device.
        opts: prev_traj, cons_weight, cons_enabled, mem, static_halfspaces (struct mpcgpu_set_options)."""
        n = n_sets * planners
        if out is None:
            out = self.alloc_outputs(n)
            out["best"] = np.zeros(n_sets, np.int32)
        opt, _keep = self._set_options(out, n, n_sets=n_sets, **opts)
        best_only = bool(opts.get("best_only"))
        sc = None if obj_scale is None else np.ascontiguousarray(obj_scale, np.float64)
        ds = None if disabled is None else np.ascontiguousarray(disabled, np.uint8)
        lin_base, lin_count = self.lin_constraint_block()
        obst_pred = np.ascontiguousarray(obst_pred, np.float64)
        n_obs = obst_pred.shape[2] if obst_pred.ndim == 4 else 0
        guided = np.ascontiguousarray(guided, np.uint8)
        nidx = 0 if param_idx is None else int(np.asarray(param_idx).size)
        pidx = None if nidx == 0 else np.ascontiguousarray(param_idx, np.int32)
        pvals = None if nidx == 0 else np.ascontiguousarray(planner_params, np.float64)
        vp = ctypes.c_void_p
        self.lib.mpcgpu_solve_sets_guided.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, vp, vp, ctypes.c_int, vp, vp, ctypes.c_int,
                                                      ctypes.c_int, ctypes.c_double, ctypes.c_int, vp, vp, vp, ctypes.c_int] + [vp] * 11
        rc = self.lib.mpcgpu_solve_sets_guided(self.handle, n_sets, planners, _ptr(np.ascontiguousarray(xinit_sets, np.float64)),
                                               _ptr(np.ascontiguousarray(shared_params, np.float64)), _ptr(np.ascontiguousarray(x0, np.float64)),
                                               n_obs, _ptr(obst_pred), _ptr(guided), lin_base, lin_count, float(robot_radius), nidx,
                                               _ptr(pidx), _ptr(pvals), None, int(num_iter), None if best_only else _ptr(out["xtraj"]),
                                               None if best_only else _ptr(out["utraj"]), _ptr(out["pobj"]), _ptr(out["exit_code"]), _ptr(out["qp_status"]), _ptr(out["res_eq"]),
                                               _ptr(sc), None, _ptr(ds), _ptr(out["best"]), opt)
        self._check(rc, "mpcgpu_solve_sets_guided")
        return out

    def table_layout(self):
        """generated/<config>/tables.yaml: invariant parameter indices and the obstacle / halfspace slot layout"""
        if getattr(self, "_table_layout", None) is None:      # (parsed once: the set entries call this per solve)
            with open(os.path.join(config_dir(self.config), "tables.yaml")) as f:
                t = yaml.safe_load(f)
            t["invariant_idx"] = np.array([self.parameter_map[n] for n in t["invariant"]], np.int32)
            self._table_layout = t
        return self._table_layout

    def solve_sets_tables(self, n_sets, planners, xinit_sets, invariant, obstacles, x0, guided=None, robot_radius=0.0, obstacle_radius=None,
                          stage_idx=None, stage=None, param_idx=None, planner_params=None, num_iter=10, out=None, obj_scale=None, disabled=None,
                          _multi=None, **opts):
        """mpcgpu_solve_sets_tables: stage-invariant parameters [n_sets, n_invariant] (order of table_layout()["invariant"]), obstacle
        table [n_sets, N, M, 2|4], warm starts; the parameter block is built on the device."""
        n = n_sets * planners
        if out is None:
            out = self.alloc_outputs(n)
            out["best"] = np.zeros(n_sets, np.int32)
        opt, _keep = self._set_options(out, n, n_sets=n_sets, **opts)
        best_only = bool(opts.get("best_only"))
        lay = self.table_layout()
        f64 = lambda a: None if a is None else np.ascontiguousarray(a, np.float64)
        inv, ob, rad, stg = f64(invariant), f64(obstacles), f64(obstacle_radius), f64(stage)
        sidx = None if stage_idx is None else np.ascontiguousarray(stage_idx, np.int32)
        g = None if guided is None else np.ascontiguousarray(guided, np.uint8)
        t = ParamTables()
        t.n_invariant = int(lay["invariant_idx"].size); t.invariant_idx = lay["invariant_idx"].ctypes.data; t.invariant = inv.ctypes.data
        assert inv.size == n_sets * t.n_invariant
        if sidx is not None:
            t.n_stage = int(sidx.size); t.stage_idx = sidx.ctypes.data; t.stage = stg.ctypes.data
        t.M = 0 if ob is None else int(ob.shape[2]); t.ob_stride = 2 if ob is None else int(ob.shape[3])
        if ob is not None:
            t.obstacles = ob.ctypes.data
        if rad is not None:
            t.obstacle_radius = rad.ctypes.data
        ell = lay.get("ellipsoid")
        off = np.array(ell["offsets"], np.int32) if ell else None
        t.ell_base = ell["base"] if ell else -1
        t.ell_stride = ell["stride"] if ell else 0
        if ell:
            t.ell_offsets = off.ctypes.data
        lin = lay.get("guidance_halfspaces")
        if g is not None and lin:
            t.guided = g.ctypes.data; t.lin_base = lin["base"]; t.lin_count = lin["count"]; t.robot_radius = float(robot_radius)
        sc = None if obj_scale is None else np.ascontiguousarray(obj_scale, np.float64)
        ds = None if disabled is None else np.ascontiguousarray(disabled, np.uint8)
        nidx = 0 if param_idx is None else int(np.asarray(param_idx).size)
        pidx = None if nidx == 0 else np.ascontiguousarray(param_idx, np.int32)
        pvals = None if nidx == 0 else np.ascontiguousarray(planner_params, np.float64)
        vp = ctypes.c_void_p
        fn = self.lib.mpcgpu_solve_sets_tables if _multi is None else self.lib.mpcgpu_multi_solve_sets_tables      # (same arguments behind the handle)
        fn.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, vp, vp, ctypes.c_int, vp, vp, vp, ctypes.c_int] + [vp] * 11
        rc = fn(self.handle if _multi is None else _multi, n_sets, planners, _ptr(f64(xinit_sets)), ctypes.byref(t), _ptr(f64(x0)), nidx, _ptr(pidx),
                                               _ptr(pvals), None, int(num_iter), None if best_only else _ptr(out["xtraj"]),
                                               None if best_only else _ptr(out["utraj"]), _ptr(out["pobj"]), _ptr(out["exit_code"]),
                                               _ptr(out["qp_status"]), _ptr(out["res_eq"]), _ptr(sc), None, _ptr(ds), _ptr(out["best"]), opt)
        if _multi is not None and rc != 0:
            raise RuntimeError("mpcgpu_multi_solve_sets_tables failed with status %d" % rc)
        self._check(rc, "mpcgpu_solve_sets_tables")
        out["h2d_bytes"] = int(8 * (inv.size + (0 if ob is None else ob.size) + (0 if rad is None else rad.size) + (0 if stg is None else stg.size)
                                    + np.asarray(x0).size + np.asarray(xinit_sets).size + (0 if pvals is None else pvals.size)) + (0 if g is None else g.size))
        return out

    def guidance_halfspaces_device(self, n_sets, planners, xinit_sets, x0, obst_pred, n_obs, guided, robot_radius, params, stream=None):
        """device pointers (ints); asynchronous"""
        lin_base, lin_count = self.lin_constraint_block()
        vp = ctypes.c_void_p
        self.lib.mpcgpu_guidance_halfspaces_device.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, vp, vp, ctypes.c_int, vp, ctypes.c_int,
                                                               ctypes.c_int, ctypes.c_double, vp, vp]
        rc = self.lib.mpcgpu_guidance_halfspaces_device(self.handle, n_sets, planners, xinit_sets, x0, obst_pred, n_obs, guided, lin_base,
                                                        lin_count, float(robot_radius), params, stream)
        self._check(rc, "mpcgpu_guidance_halfspaces_device")

    def select_best(self, set_offsets, pobj, exit_code, obj_scale=None, obj_sub=None, disabled=None):
        set_offsets = np.ascontiguousarray(set_offsets, np.int32)
        n_sets = set_offsets.size - 1
        best = np.zeros(n_sets, np.int32)
        pobj = np.ascontiguousarray(pobj, np.float64)
        exit_code = np.ascontiguousarray(exit_code, np.int32)
        sc = None if obj_scale is None else np.ascontiguousarray(obj_scale, np.float64)
        sb = None if obj_sub is None else np.ascontiguousarray(obj_sub, np.float64)
        ds = None if disabled is None else np.ascontiguousarray(disabled, np.uint8)
        rc = self.lib.mpcgpu_select_best(self.handle, n_sets, _ptr(set_offsets), _ptr(pobj), _ptr(exit_code), _ptr(sc), _ptr(sb),
                                         _ptr(ds), _ptr(best))
        self._check(rc, "mpcgpu_select_best")
        return best

    def select_best_device(self, n_sets, set_offsets, pobj, exit_code, best_idx, obj_scale=None, obj_sub=None, disabled=None,
                           stream=None):
        rc = self.lib.mpcgpu_select_best_device(self.handle, n_sets, _ptr(set_offsets), _ptr(pobj), _ptr(exit_code),
                                                _ptr(obj_scale), _ptr(obj_sub), _ptr(disabled), _ptr(best_idx), _ptr(stream))
        self._check(rc, "mpcgpu_select_best_device")

    def model_eval(self, z, p, pi, mh):
        """Evaluate the emitted device model functions at n points; returns a dict of arrays."""
        n = z.shape[0]
        nhs = ctypes.c_int()
        self.lib.mpcgpu_model_eval_doubles.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)]
        D = self.lib.mpcgpu_model_eval_doubles(self.handle, ctypes.byref(nhs))
        nhs = nhs.value
        out = np.zeros((n, D))
        self.lib.mpcgpu_model_eval.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 5
        zz, pp, pi_, mh_ = (np.ascontiguousarray(a, np.float64) for a in (z, p, pi, mh))
        self._check(self.lib.mpcgpu_model_eval(self.handle, n, _ptr(zz), _ptr(pp), _ptr(pi_), _ptr(mh_), _ptr(out)), "mpcgpu_model_eval")
        nx, nz, nh = self.nx, self.nz, self.nh
        pk = nz * (nz + 1) // 2
        sizes = [("xn", nx), ("W", nx * nz), ("Hdyn", pk), ("cost", 1), ("g", nz), ("Hcost", pk), ("h", nh), ("C", nh * nhs), ("Hcon", pk)]
        res, o = {"nhs": nhs}, 0
        for name, sz in sizes:
            res[name] = out[:, o:o + sz]
            o += sz
        return res

    def sync(self):
        self._check(self.lib.mpcgpu_sync(self.handle), "mpcgpu_sync")

    def last_kernel_ms(self):
        return float(self.lib.mpcgpu_last_kernel_ms(self.handle))

    def set_kernel_mode(self, mode):
        """0 auto, 1 thread-per-stage kernel, 2 role-split kernel; returns True if the configuration has a role-split kernel"""
        self.lib.mpcgpu_set_kernel_mode.argtypes = [ctypes.c_void_p, ctypes.c_int]
        rc = self.lib.mpcgpu_set_kernel_mode(self.handle, int(mode))
        self._check(min(rc, 0), "mpcgpu_set_kernel_mode")
        return rc == 1

    def launch_count(self):
        return int(self.lib.mpcgpu_launch_count(self.handle))


class MultiEngine:
    """mpcgpu_multi_* (include/mpcgpu.h): several GPUs of one node behind one handle; homotopy sets partitioned by set into
    contiguous ranges, one host thread + the engine's streams per device, no collective.  `devices` may repeat a device
    (two engines on one GPU), which exercises the sharding on a single-GPU box."""

    def __init__(self, config, devices, max_batch_per_device):
        self.lib = load_library()
        vp = ctypes.c_void_p
        self.lib.mpcgpu_multi_create.argtypes = [ctypes.c_char_p, vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(vp)]
        self.lib.mpcgpu_multi_destroy.argtypes = [vp]
        self.lib.mpcgpu_multi_engine.argtypes = [vp, ctypes.c_int]
        self.lib.mpcgpu_multi_engine.restype = vp
        self.lib.mpcgpu_multi_last_kernel_ms.argtypes = [vp]
        self.lib.mpcgpu_multi_last_kernel_ms.restype = ctypes.c_float
        self.lib.mpcgpu_multi_num_devices.argtypes = [vp]
        devs = np.ascontiguousarray(devices, np.int32)
        h = vp()
        rc = self.lib.mpcgpu_multi_create(config.encode(), _ptr(devs), int(devs.size), int(max_batch_per_device), ctypes.byref(h))
        if rc != 0:
            raise RuntimeError("mpcgpu_multi_create(%s, %s) failed with status %d (no CPU fallback)" % (config, list(devices), rc))
        self.handle = h
        self.devices = [int(d) for d in devs]
        self.first = Engine.__new__(Engine)            # dimensions / maps / option packing of the first engine (not owned)
        self.first.lib, self.first.handle, self.first.config = self.lib, None, config
        self.first.parameter_map, self.first.model_map, st = load_maps(config)
        N, nx, nu, npar, nh = (ctypes.c_int() for _ in range(5))
        e0 = self.lib.mpcgpu_multi_engine(self.handle, 0)
        self.lib.mpcgpu_desc_query(e0, ctypes.byref(N), ctypes.byref(nx), ctypes.byref(nu), ctypes.byref(npar), ctypes.byref(nh))
        f = self.first
        f.N, f.nx, f.nu, f.npar, f.nh = N.value, nx.value, nu.value, npar.value, nh.value
        f.nz = f.nx + f.nu
        f.mem_doubles = int(self.lib.mpcgpu_mem_doubles(e0))
        self.N, self.nx, self.nu, self.npar, self.nz, self.mem_doubles = f.N, f.nx, f.nu, f.npar, f.nz, f.mem_doubles

    def close(self):
        if getattr(self, "handle", None):
            self.lib.mpcgpu_multi_destroy(self.handle)
            self.handle = None

    __del__ = close

    def set_kernel_mode(self, mode):
        self.lib.mpcgpu_set_kernel_mode.argtypes = [ctypes.c_void_p, ctypes.c_int]
        for i in range(len(self.devices)):
            self.lib.mpcgpu_set_kernel_mode(self.lib.mpcgpu_multi_engine(self.handle, i), int(mode))

    def shard_range(self, n_units, i):
        b, e = ctypes.c_int(), ctypes.c_int()
        self.lib.mpcgpu_multi_shard_range(int(n_units), len(self.devices), int(i), ctypes.byref(b), ctypes.byref(e))
        return b.value, e.value

    def last_kernel_ms(self):
        return float(self.lib.mpcgpu_multi_last_kernel_ms(self.handle))

    def alloc_outputs(self, n):
        return dict(xtraj=np.zeros((n, (self.N + 1) * self.nx)), utraj=np.zeros((n, self.N * self.nu)), pobj=np.zeros(n),
                    exit_code=np.zeros(n, np.int32), qp_status=np.zeros(n, np.int32), res_eq=np.zeros(n), ipm_iters=np.zeros(n, np.int32))

    def solve_batch(self, xinit, x0, params, num_iter=10, mem=None, out=None):
        n = xinit.shape[0]
        out = out or self.alloc_outputs(n)
        vp = ctypes.c_void_p
        self.lib.mpcgpu_multi_solve_batch.argtypes = [vp, ctypes.c_int, vp, vp, vp, vp, ctypes.c_int] + [vp] * 8
        a = [np.ascontiguousarray(v, np.float64) for v in (xinit, x0, params)]
        rc = self.lib.mpcgpu_multi_solve_batch(self.handle, n, _ptr(a[0]), _ptr(a[1]), _ptr(a[2]), None, int(num_iter), _ptr(mem), _ptr(out["xtraj"]),
                                               _ptr(out["utraj"]), _ptr(out["pobj"]), _ptr(out["exit_code"]), _ptr(out["qp_status"]),
                                               _ptr(out["res_eq"]), _ptr(out["ipm_iters"]))
        if rc != 0:
            raise RuntimeError("mpcgpu_multi_solve_batch failed with status %d" % rc)
        return out

    def solve_sets(self, n_sets, planners, xinit_sets, shared_params, x0, param_idx, planner_params, num_iter=10, obj_scale=None,
                   disabled=None, guided_args=None, **opts):
        """mpcgpu_multi_solve_sets / _guided (guided_args = (obst_pred, guided, robot_radius, lin_base, lin_count))"""
        n = n_sets * planners
        out = self.alloc_outputs(n)
        out["best"] = np.zeros(n_sets, np.int32)
        opt, _keep = self.first._set_options(out, n, n_sets=n_sets, **opts)
        best_only = bool(opts.get("best_only"))
        sc = None if obj_scale is None else np.ascontiguousarray(obj_scale, np.float64)
        ds = None if disabled is None else np.ascontiguousarray(disabled, np.uint8)
        nidx = 0 if param_idx is None else int(np.asarray(param_idx).size)
        idx = None if nidx == 0 else np.ascontiguousarray(param_idx, np.int32)
        pv = None if nidx == 0 else np.ascontiguousarray(planner_params, np.float64)
        xs, sh, x0_ = (np.ascontiguousarray(a, np.float64) for a in (xinit_sets, shared_params, x0))
        vp = ctypes.c_void_p
        xt, ut = (None, None) if best_only else (_ptr(out["xtraj"]), _ptr(out["utraj"]))
        tail = [None, int(num_iter), xt, ut, _ptr(out["pobj"]), _ptr(out["exit_code"]), _ptr(out["qp_status"]), _ptr(out["res_eq"]), _ptr(sc), None,
                _ptr(ds), _ptr(out["best"]), opt]
        if guided_args is None:
            self.lib.mpcgpu_multi_solve_sets.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, vp, vp, ctypes.c_int, vp, vp, vp, ctypes.c_int] + [vp] * 11
            rc = self.lib.mpcgpu_multi_solve_sets(self.handle, n_sets, planners, _ptr(xs), _ptr(sh), _ptr(x0_), nidx, _ptr(idx), _ptr(pv), *tail)
        else:
            obst_pred, guided, robot_radius, lin_base, lin_count = guided_args
            ob = np.ascontiguousarray(obst_pred, np.float64); g = np.ascontiguousarray(guided, np.uint8)
            self.lib.mpcgpu_multi_solve_sets_guided.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, vp, vp, ctypes.c_int, vp, vp, ctypes.c_int,
                                                                ctypes.c_int, ctypes.c_double, ctypes.c_int, vp, vp, vp, ctypes.c_int] + [vp] * 11
            rc = self.lib.mpcgpu_multi_solve_sets_guided(self.handle, n_sets, planners, _ptr(xs), _ptr(sh), _ptr(x0_), int(ob.shape[2]), _ptr(ob), _ptr(g),
                                                         int(lin_base), int(lin_count), float(robot_radius), nidx, _ptr(idx), _ptr(pv), *tail)
        if rc != 0:
            raise RuntimeError("mpcgpu_multi_solve_sets failed with status %d" % rc)
        return out

    def solve_sets_tables(self, n_sets, planners, xinit_sets, invariant, obstacles, x0, **kw):
        """mpcgpu_multi_solve_sets_tables: the struct-of-tables entry over all devices (arguments as Engine.solve_sets_tables)"""
        return self.first.solve_sets_tables(n_sets, planners, xinit_sets, invariant, obstacles, x0, _multi=self.handle, **kw)
