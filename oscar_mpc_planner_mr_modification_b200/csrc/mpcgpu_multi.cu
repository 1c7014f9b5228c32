// Several GPUs of one node behind one handle (include/mpcgpu.h: mpcgpu_multi_*).  Host code only: one engine, one host
// thread and the engine's own streams per device; homotopy sets are partitioned BY SET into contiguous ranges (a set's
// argmin is taken on the device that solved it), results land in the caller's arrays at the range offsets.  No collective:
// the problems are independent (the reference's counterpart is the OpenMP team over planners,
// mpc_planner_modules/src/guidance_constraints.cpp:304, and one ROS node per robot).
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/mpcgpu.h"

namespace {
// a persistent worker per device: a homotopy set per control cycle must not pay a thread spawn
struct Worker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<int()> job;
    bool has_job = false, done = true, quit = false;
    int rc = 0;
    void loop()
    {
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv.wait(lk, [&] { return has_job || quit; });
            if (quit) return;
            std::function<int()> j = std::move(job);
            has_job = false;
            lk.unlock();
            const int r = j();
            lk.lock();
            rc = r;
            done = true;
            cv.notify_all();
        }
    }
    void submit(std::function<int()> j)
    {
        std::lock_guard<std::mutex> lk(mu);
        job = std::move(j);
        has_job = true;
        done = false;
        cv.notify_all();
    }
    int wait()
    {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return done; });
        return rc;
    }
};
}  // namespace

struct mpcgpu_multi {
    std::vector<mpcgpu_engine*> eng;
    std::vector<Worker*> workers;
    std::vector<char> used;      // device took part in the last call
    int max_batch = 0;
    std::mutex call_mu;          // one multi call at a time per handle
};

static void shard(int n_units, int n_dev, int i, int* b, int* e)
{
    const int per = n_units / n_dev;
    *b = i * per;
    *e = (i == n_dev - 1) ? n_units : *b + per;
}

// run fn(device index, begin, end) on every device with a non-empty range; first error wins
template <class F>
static int run_sharded(mpcgpu_multi* m, int n_units, F fn)
{
    std::lock_guard<std::mutex> lk(m->call_mu);
    const int D = (int)m->eng.size();
    for (int i = 0; i < D; i++) {
        int b, e;
        shard(n_units, D, i, &b, &e);
        m->used[i] = e > b;
        if (e > b) m->workers[i]->submit([=] { return fn(i, b, e); });
    }
    int rc = MPCGPU_OK;
    for (int i = 0; i < D; i++)
        if (m->used[i]) {
            const int r = m->workers[i]->wait();
            if (r != MPCGPU_OK && rc == MPCGPU_OK) rc = r;
        }
    return rc;
}

extern "C" {

int mpcgpu_multi_shard_range(int n_units, int n_devices, int i, int* begin, int* end)
{
    if (n_units < 0 || n_devices <= 0 || i < 0 || i >= n_devices || !begin || !end) return MPCGPU_ERR_ARG;
    shard(n_units, n_devices, i, begin, end);
    return MPCGPU_OK;
}

int mpcgpu_multi_create(const char* config_name, const int* devices, int n_devices, int max_batch_per_device, mpcgpu_multi** out)
{
    if (!config_name || !devices || n_devices <= 0 || max_batch_per_device <= 0 || !out) return MPCGPU_ERR_ARG;
    *out = nullptr;
    mpcgpu_multi* m = new mpcgpu_multi();
    m->max_batch = max_batch_per_device;
    for (int i = 0; i < n_devices; i++) {
        mpcgpu_engine* e = nullptr;
        const int rc = mpcgpu_engine_create(config_name, devices[i], max_batch_per_device, &e);
        if (rc != MPCGPU_OK) {
            if (e) mpcgpu_engine_destroy(e);
            for (auto* q : m->eng) mpcgpu_engine_destroy(q);
            delete m;
            return rc;
        }
        m->eng.push_back(e);
    }
    m->used.assign(n_devices, 0);
    for (int i = 0; i < n_devices; i++) {
        Worker* w = new Worker();
        w->th = std::thread([w] { w->loop(); });
        m->workers.push_back(w);
    }
    *out = m;
    return MPCGPU_OK;
}

int mpcgpu_multi_destroy(mpcgpu_multi* m)
{
    if (!m) return MPCGPU_ERR_ARG;
    for (auto* w : m->workers) {
        {
            std::lock_guard<std::mutex> lk(w->mu);
            w->quit = true;
            w->cv.notify_all();
        }
        w->th.join();
        delete w;
    }
    for (auto* e : m->eng) mpcgpu_engine_destroy(e);
    delete m;
    return MPCGPU_OK;
}

int mpcgpu_multi_num_devices(const mpcgpu_multi* m) { return m ? (int)m->eng.size() : MPCGPU_ERR_ARG; }
mpcgpu_engine* mpcgpu_multi_engine(mpcgpu_multi* m, int i) { return (m && i >= 0 && i < (int)m->eng.size()) ? m->eng[i] : nullptr; }

float mpcgpu_multi_last_kernel_ms(mpcgpu_multi* m)
{
    if (!m) return -1.0f;
    float mx = 0.0f;
    for (size_t i = 0; i < m->eng.size(); i++)
        if (m->used[i]) {
            const float t = mpcgpu_last_kernel_ms(m->eng[i]);
            if (t < 0.0f) return -1.0f;
            if (t > mx) mx = t;
        }
    return mx;
}

// the slice of the optional arguments that belongs to sets [s0, s1) / problems [p0, ...)
static mpcgpu_set_options slice_options(const mpcgpu_set_options* opt, int s0, size_t p0, int N, int nx, int nu, int md)
{
    mpcgpu_set_options o = *opt;
    if (o.prev_traj) o.prev_traj += (size_t)s0 * N * 2;
    if (o.consistency_enabled) o.consistency_enabled += p0;
    if (o.mem_inout) o.mem_inout += p0 * md;
    if (o.objective_out) o.objective_out += p0;
    if (o.consistency_cost_out) o.consistency_cost_out += p0;
    if (o.static_halfspaces) o.static_halfspaces += (size_t)s0 * N * o.n_static * 3;
    if (o.best_xtraj) o.best_xtraj += (size_t)s0 * nx * (N + 1);
    if (o.best_utraj) o.best_utraj += (size_t)s0 * nu * N;
    return o;
}

int mpcgpu_multi_solve_sets_guided(mpcgpu_multi* m, int n_sets, int planners, const double* xinit_sets, const double* shared_params,
                                   const double* x0, int n_obs, const double* obst_pred, const unsigned char* guided, int lin_base,
                                   int lin_count, double robot_radius, int nidx, const int* param_idx, const double* planner_params,
                                   const int* num_iter, int num_iter_all, double* xtraj, double* utraj, double* pobj, int* exit_code,
                                   int* qp_status, double* res_eq, const double* obj_scale, const double* obj_sub,
                                   const unsigned char* disabled, int* best_idx, const mpcgpu_set_options* opt)
{
    if (!m || n_sets < 0 || planners <= 0) return MPCGPU_ERR_ARG;
    if (n_sets == 0) return MPCGPU_OK;
    int N, nx, nu, np, nh;
    mpcgpu_desc_query(m->eng[0], &N, &nx, &nu, &np, &nh);
    const int md = mpcgpu_mem_doubles(m->eng[0]);
    const size_t nz = (size_t)nx + nu;
    const bool with_guidance = guided != nullptr;
    return run_sharded(m, n_sets, [=](int dev, int s0, int s1) {
        const int ns = s1 - s0;
        const size_t p0 = (size_t)s0 * planners;
        mpcgpu_set_options o;
        if (opt) o = slice_options(opt, s0, p0, N, nx, nu, md);
        const double* xs = xinit_sets + (size_t)s0 * nx;
        const double* sh = shared_params + (size_t)s0 * N * np;
        const double* x0d = x0 + p0 * nz * (N + 1);
        const double* pv = planner_params ? planner_params + p0 * N * nidx : nullptr;
        const int* ni = num_iter ? num_iter + p0 : nullptr;
        double* xt = xtraj ? xtraj + p0 * nx * (N + 1) : nullptr;
        double* ut = utraj ? utraj + p0 * nu * N : nullptr;
        const double* sc = obj_scale ? obj_scale + p0 : nullptr;
        const double* sb = obj_sub ? obj_sub + p0 : nullptr;
        const unsigned char* ds = disabled ? disabled + p0 : nullptr;
        if (with_guidance)
            return mpcgpu_solve_sets_guided(m->eng[dev], ns, planners, xs, sh, x0d, n_obs, obst_pred ? obst_pred + (size_t)s0 * N * n_obs * 2 : nullptr,
                                            guided + p0, lin_base, lin_count, robot_radius, nidx, param_idx, pv, ni, num_iter_all, xt, ut, pobj + p0,
                                            exit_code + p0, qp_status + p0, res_eq + p0, sc, sb, ds, best_idx + s0, opt ? &o : nullptr);
        return mpcgpu_solve_sets(m->eng[dev], ns, planners, xs, sh, x0d, nidx, param_idx, pv, ni, num_iter_all, xt, ut, pobj + p0, exit_code + p0,
                                 qp_status + p0, res_eq + p0, sc, sb, ds, best_idx + s0, opt ? &o : nullptr);
    });
}

int mpcgpu_multi_solve_sets(mpcgpu_multi* m, int n_sets, int planners, const double* xinit_sets, const double* shared_params,
                            const double* x0, int nidx, const int* param_idx, const double* planner_params, const int* num_iter,
                            int num_iter_all, double* xtraj, double* utraj, double* pobj, int* exit_code, int* qp_status, double* res_eq,
                            const double* obj_scale, const double* obj_sub, const unsigned char* disabled, int* best_idx,
                            const mpcgpu_set_options* opt)
{
    return mpcgpu_multi_solve_sets_guided(m, n_sets, planners, xinit_sets, shared_params, x0, 0, nullptr, nullptr, 0, 0, 0.0, nidx, param_idx,
                                          planner_params, num_iter, num_iter_all, xtraj, utraj, pobj, exit_code, qp_status, res_eq, obj_scale,
                                          obj_sub, disabled, best_idx, opt);
}

// struct-of-tables entry (2.4 KB per solve from the host for c2): every table is per set, per (set, stage) or per problem
int mpcgpu_multi_solve_sets_tables(mpcgpu_multi* m, int n_sets, int planners, const double* xinit_sets, const mpcgpu_param_tables* tables,
                                   const double* x0, int nidx, const int* param_idx, const double* planner_params, const int* num_iter,
                                   int num_iter_all, double* xtraj, double* utraj, double* pobj, int* exit_code, int* qp_status,
                                   double* res_eq, const double* obj_scale, const double* obj_sub, const unsigned char* disabled,
                                   int* best_idx, const mpcgpu_set_options* opt)
{
    if (!m || n_sets < 0 || planners <= 0 || !tables) return MPCGPU_ERR_ARG;
    if (n_sets == 0) return MPCGPU_OK;
    int N, nx, nu, np, nh;
    mpcgpu_desc_query(m->eng[0], &N, &nx, &nu, &np, &nh);
    const int md = mpcgpu_mem_doubles(m->eng[0]);
    const size_t nz = (size_t)nx + nu;
    const mpcgpu_param_tables T = *tables;
    return run_sharded(m, n_sets, [=](int dev, int s0, int s1) {
        const int ns = s1 - s0;
        const size_t p0 = (size_t)s0 * planners;
        mpcgpu_set_options o;
        if (opt) o = slice_options(opt, s0, p0, N, nx, nu, md);
        mpcgpu_param_tables t = T;
        if (t.invariant) t.invariant += (size_t)s0 * t.n_invariant;
        if (t.stage) t.stage += (size_t)s0 * N * t.n_stage;
        if (t.obstacles) t.obstacles += (size_t)s0 * N * t.M * t.ob_stride;
        if (t.obstacle_radius) t.obstacle_radius += (size_t)s0 * t.M;
        if (t.guided) t.guided += p0;
        return mpcgpu_solve_sets_tables(m->eng[dev], ns, planners, xinit_sets + (size_t)s0 * nx, &t, x0 + p0 * nz * (N + 1), nidx, param_idx,
                                        planner_params ? planner_params + p0 * N * nidx : nullptr, num_iter ? num_iter + p0 : nullptr, num_iter_all,
                                        xtraj ? xtraj + p0 * nx * (N + 1) : nullptr, utraj ? utraj + p0 * nu * N : nullptr, pobj + p0, exit_code + p0,
                                        qp_status + p0, res_eq + p0, obj_scale ? obj_scale + p0 : nullptr, obj_sub ? obj_sub + p0 : nullptr,
                                        disabled ? disabled + p0 : nullptr, best_idx + s0, opt ? &o : nullptr);
    });
}

int mpcgpu_multi_solve_batch(mpcgpu_multi* m, int n, const double* xinit, const double* x0, const double* params, const int* num_iter,
                             int num_iter_all, double* mem_inout, double* xtraj, double* utraj, double* pobj, int* exit_code,
                             int* qp_status, double* res_eq, int* ipm_iters)
{
    if (!m || n < 0) return MPCGPU_ERR_ARG;
    if (n == 0) return MPCGPU_OK;
    int N, nx, nu, np, nh;
    mpcgpu_desc_query(m->eng[0], &N, &nx, &nu, &np, &nh);
    const size_t md = (size_t)mpcgpu_mem_doubles(m->eng[0]), nz = (size_t)nx + nu;
    return run_sharded(m, n, [=](int dev, int b, int e) {
        const size_t p = (size_t)b;
        return mpcgpu_solve_batch(m->eng[dev], e - b, xinit + p * nx, x0 + p * nz * (N + 1), params + p * N * np, num_iter ? num_iter + p : nullptr,
                                  num_iter_all, mem_inout ? mem_inout + p * md : nullptr, xtraj + p * nx * (N + 1), utraj + p * nu * N, pobj + p,
                                  exit_code + p, qp_status + p, res_eq + p, ipm_iters ? ipm_iters + p : nullptr);
    });
}

}  // extern "C"
