// mpcgpu_wire.inl -- ROS 1 wire formats of mpc_planner_msgs <-> engine tables (include/mpcgpu_wire.h).
// Included at the end of mpcgpu_capi.cu (it uses the engine's buffers).  Host-side parsing / serialization plus one
// device kernel that writes the ellipsoid parameter slots from obstacle tables.
#include "../../include/mpcgpu_wire.h"

#include <algorithm>
#include <cmath>
#include <numeric>

namespace {

// ---- roscpp serialization primitives (little endian; string / array = uint32 count + payload) ----------------------
struct Reader {
    const unsigned char* p;
    size_t left;
    bool ok = true;
    bool take(void* out, size_t n)
    {
        if (!ok || left < n) { ok = false; return false; }
        if (out) memcpy(out, p, n);
        p += n; left -= n;
        return true;
    }
    unsigned u32() { unsigned v = 0; take(&v, 4); return v; }
    int i32() { int v = 0; take(&v, 4); return v; }
    double f64() { double v = 0; take(&v, 8); return v; }
    void skip_string() { const unsigned n = u32(); take(nullptr, n); }
    void skip_header() { u32(); u32(); u32(); skip_string(); }            // seq, stamp.sec, stamp.nsec, frame_id
    void skip_f64_array() { const unsigned n = u32(); if (ok && (size_t)n * 8 <= left) take(nullptr, (size_t)n * 8); else ok = false; }
};

// RosTools::quaternionToAngle (un-vendored ros_tools): yaw of the quaternion
double quat_yaw(double qx, double qy, double qz, double qw)
{
    return atan2(2.0 * (qw * qz + qx * qy), 1.0 - 2.0 * (qy * qy + qz * qz));
}

// geometry_msgs/Pose: Point (x, y, z) + Quaternion (x, y, z, w)
void read_pose(Reader& r, double* x, double* y, double* yaw)
{
    const double px = r.f64(), py = r.f64();
    r.f64();
    const double qx = r.f64(), qy = r.f64(), qz = r.f64(), qw = r.f64();
    *x = px; *y = py; *yaw = quat_yaw(qx, qy, qz, qw);
}

bool read_obstacle_gmm(Reader& r, mpcgpu_track* t, int max_steps, double* steps)
{
    t->id = r.i32();
    read_pose(r, &t->x, &t->y, &t->angle);
    t->radius = 0.0;
    t->n_steps = 0;
    const unsigned n_gauss = r.u32();
    for (unsigned g = 0; g < n_gauss && r.ok; g++) {
        r.skip_header();                                                  // nav_msgs/Path mean: header
        const unsigned n_poses = r.u32();
        for (unsigned k = 0; k < n_poses && r.ok; k++) {
            r.skip_header();                                              // PoseStamped.header
            double x, y, a;
            read_pose(r, &x, &y, &a);
            if (g == 0 && (int)k < max_steps && steps) {                  // only the first Gaussian is used (:593-601)
                steps[3 * k] = x; steps[3 * k + 1] = y; steps[3 * k + 2] = a;
                t->n_steps = (int)k + 1;
            }
        }
        r.skip_f64_array();                                               // major_semiaxis
        r.skip_f64_array();                                               // minor_semiaxis
    }
    r.skip_f64_array();                                                   // probabilities
    return r.ok;
}

struct Writer {
    unsigned char* p;
    size_t left;
    bool ok = true;
    size_t written = 0;
    void put(const void* src, size_t n)
    {
        if (!ok || left < n) { ok = false; return; }
        memcpy(p, src, n);
        p += n; left -= n; written += n;
    }
    void u32(unsigned v) { put(&v, 4); }
    void i32(int v) { put(&v, 4); }
    void f64(double v) { put(&v, 8); }
    void u8(unsigned char v) { put(&v, 1); }
    void str(const char* s) { const unsigned n = s ? (unsigned)strlen(s) : 0u; u32(n); if (n) put(s, n); }
};

__global__ void pack_obstacles_kernel(int n_sets, int N, int nx, int npar, int M, int ell_base, int ell_stride, int o_x, int o_y, int o_psi,
                                      int o_major, int o_minor, int o_chi, int o_r, const double* __restrict__ xinit_sets,
                                      const double* __restrict__ table, int obs, const double* __restrict__ radius, double* __restrict__ params)
{
    const long long total = (long long)n_sets * N * M;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(t % M), k = (int)((t / M) % N), s = (int)(t / ((long long)M * N));
        double* P = params + ((size_t)s * N + k) * npar + ell_base + (size_t)j * ell_stride;
        double x, y, psi, r;
        if (k == 0) {                                                     // ellipsoid_constraints.cpp:30-31,42-56
            x = xinit_sets[(size_t)s * nx] + 50.0; y = xinit_sets[(size_t)s * nx + 1] + 50.0; psi = 0.0; r = 0.1;
        } else {                                                          // :66-77 (prediction step k-1)
            const double* T = table + (((size_t)s * N + (k - 1)) * M + j) * obs;      // obs = 4: (x, y, psi, r); 2: (x, y) + radius table
            x = T[0]; y = T[1];
            psi = obs >= 4 ? T[2] : 0.0;
            r = obs >= 4 ? T[3] : radius[(size_t)s * M + j];
        }
        P[o_x] = x; P[o_y] = y; P[o_psi] = psi; P[o_r] = r; P[o_major] = 0.0; P[o_minor] = 0.0; P[o_chi] = 1.0;
    }
}

}  // namespace

long mpcgpu_wire_parse_obstacle_gmm(const unsigned char* buf, size_t len, mpcgpu_track* track, int max_steps, double* steps)
{
    if (!buf || !track || max_steps < 0 || (max_steps > 0 && !steps)) return MPCGPU_ERR_ARG;
    Reader r{buf, len};
    if (!read_obstacle_gmm(r, track, max_steps, steps)) return MPCGPU_ERR_ARG;
    return (long)(len - r.left);
}

long mpcgpu_wire_parse_obstacle_array(const unsigned char* buf, size_t len, int max_tracks, int max_steps, mpcgpu_track* tracks,
                                      double* steps, int* n_tracks)
{
    if (!buf || !tracks || !n_tracks || max_tracks < 0 || max_steps < 0 || (max_steps > 0 && !steps)) return MPCGPU_ERR_ARG;
    Reader r{buf, len};
    r.skip_header();
    const unsigned n = r.u32();
    *n_tracks = 0;
    for (unsigned i = 0; i < n && r.ok; i++) {
        mpcgpu_track tmp;
        const bool keep = (int)i < max_tracks;
        if (!read_obstacle_gmm(r, keep ? &tracks[i] : &tmp, keep ? max_steps : 0, keep ? steps + (size_t)i * max_steps * 3 : nullptr))
            return MPCGPU_ERR_ARG;
        if (keep) *n_tracks = (int)i + 1;
    }
    if (!r.ok) return MPCGPU_ERR_ARG;
    return (long)(len - r.left);
}

int mpcgpu_obstacle_table(const mpcgpu_track* tracks, const double* steps, int n_tracks, int max_steps, int N, int max_obstacles,
                          const double* st, double* table)
{
    if (n_tracks < 0 || (n_tracks > 0 && (!tracks || !steps)) || N <= 0 || max_obstacles < 0 || !st || !table || max_steps < 0) return MPCGPU_ERR_ARG;
    std::vector<int> idx((size_t)n_tracks);
    std::iota(idx.begin(), idx.end(), 0);
    int kept = n_tracks;
    if (n_tracks > max_obstacles) {                                       // keep the closest (data_preparation.cpp:106-150)
        std::vector<double> dist((size_t)n_tracks);
        const double dx = cos(st[2]), dy = sin(st[2]);
        for (int i = 0; i < n_tracks; i++) {
            if (tracks[i].n_steps < N) return MPCGPU_ERR_ARG;
            double min_dist = 1e5;
            for (int k = 0; k < N; k++) {
                const double* p = steps + ((size_t)i * max_steps + k) * 3;
                const double ex = st[0] + st[3] * (double)k * dx, ey = st[1] + st[3] * (double)k * dy;
                const double d = (double)(k + 1) * 0.6 * sqrt((p[0] - ex) * (p[0] - ex) + (p[1] - ey) * (p[1] - ey));
                if (d < min_dist) min_dist = d;
            }
            dist[(size_t)i] = min_dist;
        }
        std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return dist[(size_t)a] < dist[(size_t)b]; });
        kept = max_obstacles;
    }
    for (int j = 0; j < max_obstacles; j++) {
        if (j < kept) {
            const int i = idx[(size_t)j];
            if (tracks[i].n_steps < N) return MPCGPU_ERR_ARG;
            for (int k = 0; k < N; k++) {
                const double* p = steps + ((size_t)i * max_steps + k) * 3;
                double* T = table + ((size_t)k * max_obstacles + j) * 4;
                T[0] = p[0]; T[1] = p[1]; T[2] = p[2]; T[3] = tracks[i].radius;
            }
        } else {                                                          // getDummyObstacle + constant prediction (:51-58,159-166)
            for (int k = 0; k < N; k++) {
                double* T = table + ((size_t)k * max_obstacles + j) * 4;
                T[0] = st[0] + 100.0; T[1] = st[1] + 100.0; T[2] = 0.0; T[3] = 0.0;
            }
        }
    }
    return kept;
}

static int pack_obstacles_launch(mpcgpu_engine* e, int n_sets, const double* xinit_sets, const double* table, int obs, const double* radius, int M,
                                 int ell_base, int ell_stride, const int* off, double* params, void* stream);
int mpcgpu_pack_obstacles_device(mpcgpu_engine* e, int n_sets, const double* xinit_sets, const double* table, int M, int ell_base,
                                 int ell_stride, const int* off, double* params, void* stream)
{
    return pack_obstacles_launch(e, n_sets, xinit_sets, table, 4, nullptr, M, ell_base, ell_stride, off, params, stream);
}
static int pack_obstacles_launch(mpcgpu_engine* e, int n_sets, const double* xinit_sets, const double* table, int obs, const double* radius, int M,
                                 int ell_base, int ell_stride, const int* off, double* params, void* stream)
{
    if (obs != 4 && !(obs == 2 && radius)) return MPCGPU_ERR_ARG;
    if (!e || n_sets < 0 || M < 0 || !xinit_sets || (M > 0 && !table) || !off || !params || ell_base < 0 || ell_stride <= 0) return MPCGPU_ERR_ARG;
    for (int i = 0; i < 7; i++)
        if (off[i] < 0 || off[i] >= ell_stride) return MPCGPU_ERR_ARG;
    if (M > 0 && ell_base + (M - 1) * ell_stride + ell_stride > e->ops->np) return MPCGPU_ERR_ARG;
    if (n_sets == 0 || M == 0) return MPCGPU_OK;
    CK(cudaSetDevice(e->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : e->stream;
    const long long total = (long long)n_sets * e->ops->N * M;
    const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
    pack_obstacles_kernel<<<blocks, 256, 0, st>>>(n_sets, e->ops->N, e->ops->nx, e->ops->np, M, ell_base, ell_stride, off[0], off[1], off[2], off[3],
                                                   off[4], off[5], off[6], xinit_sets, table, obs, radius, params);
    CK(cudaGetLastError());
    e->launches += 1;
    return MPCGPU_OK;
}

int mpcgpu_solve_sets_tracks(mpcgpu_engine* e, int n_sets, int planners, const double* xinit_sets, const double* shared_params,
                             const double* x0, int M, const double* table, const unsigned char* guided, int lin_base, int lin_count,
                             double robot_radius, int ell_base, int ell_stride, const int* ell_offsets, const int* num_iter,
                             int num_iter_all, double* xtraj, double* utraj, double* pobj, int* exit_code, int* qp_status, double* res_eq,
                             const double* obj_scale, const double* obj_sub, const unsigned char* disabled, int* best_idx,
                             const mpcgpu_set_options* opt)
{
    if (!e || M < 0 || (M > 0 && !table) || !guided || !ell_offsets || lin_base < 0 || lin_count < 0 || lin_base + 3 * lin_count > e->ops->np ||
        ell_base < 0 || ell_stride <= 0 || (M > 0 && ell_base + M * ell_stride > e->ops->np))
        return MPCGPU_ERR_ARG;
    for (int i = 0; i < 7; i++)
        if (ell_offsets[i] < 0 || ell_offsets[i] >= ell_stride) return MPCGPU_ERR_ARG;
    GuidedArgs ga = {M, lin_base, lin_count, table, guided, robot_radius};
    ga.ob_stride = 4;
    ga.ell_base = ell_base; ga.ell_stride = ell_stride;
    for (int i = 0; i < 7; i++) ga.ell_off[i] = ell_offsets[i];
    return drain(e, solve_sets_impl(e, n_sets, planners, xinit_sets, shared_params, x0, 0, nullptr, nullptr, &ga, num_iter, num_iter_all, xtraj, utraj,
                                    pobj, exit_code, qp_status, res_eq, obj_scale, obj_sub, disabled, best_idx, opt));
}

long mpcgpu_wire_serialize_metrics(const mpcgpu_metrics* m, unsigned char* buf, size_t cap)
{
    if (!m || !buf || m->n_planners < 0 || (m->n_planners > 0 && !m->objective_values_all_planners)) return MPCGPU_ERR_ARG;
    Writer w{buf, cap};
    w.u32(m->seq); w.u32(m->stamp_sec); w.u32(m->stamp_nsec); w.str(m->frame_id);          // Header header
    w.str(m->robot_name);                                                                   // string robot_name
    w.f64(m->solve_time_ms); w.f64(m->success_rate); w.i32(m->iterations); w.i32(m->exit_code); w.f64(m->objective_value);
    w.u32((unsigned)m->n_planners);                                                         // float64[] objective_values_all_planners
    for (int i = 0; i < m->n_planners; i++) w.f64(m->objective_values_all_planners[i]);
    w.i32(0); w.i32(0); w.u8(0);                                   // current_topology_id, previous_topology_id, topology_switch
    w.u8(m->used_guidance); w.i32(m->selected_planner_index); w.i32(m->num_of_guidance_found);
    w.str(nullptr); w.str(nullptr);                                // current_state, previous_state
    w.u32(0);                                                      // float64[] current_position
    w.f64(0.0); w.f64(0.0);                                        // current_linear_x, current_angular_vel
    w.str(nullptr); w.i32(0); w.i32(0); w.f64(0.0);                // last_communication_trigger, messages_sent/saved_total, savings
    w.u32(0); w.u32(0); w.u32(0); w.u32(0);                        // topology_selection_counts, topology_labels, trigger counts, labels
    return w.ok ? (long)w.written : (long)MPCGPU_ERR_ARG;
}

int mpcgpu_metrics_from_set(mpcgpu_metrics* m, int planners, const double* pobj, const int* exit_code, int best_idx,
                            const unsigned char* guided, double* objective_values_out)
{
    if (!m || planners <= 0 || !pobj || !exit_code || best_idx < -1 || best_idx >= planners) return MPCGPU_ERR_ARG;
    int found = 0, ok = 0;
    for (int i = 0; i < planners; i++) {
        if (objective_values_out) objective_values_out[i] = (exit_code[i] == 1) ? pobj[i] : -1.0;
        if (guided && guided[i]) found++;
        if (exit_code[i] == 1) ok++;
    }
    m->objective_values_all_planners = objective_values_out;
    m->n_planners = objective_values_out ? planners : 0;
    m->selected_planner_index = best_idx;
    m->exit_code = best_idx >= 0 ? exit_code[best_idx] : exit_code[0];        // all failed: planner 0's code (guidance_constraints.cpp:441)
    m->objective_value = best_idx >= 0 ? pobj[best_idx] : -1.0;
    m->used_guidance = (best_idx >= 0 && guided && guided[best_idx]) ? 1 : 0;
    m->num_of_guidance_found = found;
    m->success_rate = (double)ok / (double)planners;
    return MPCGPU_OK;
}
