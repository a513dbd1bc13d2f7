/*
 * ORACLE - TEST INFRASTRUCTURE ONLY (CPU baseline + checker).  Not product code; never linked into libptzba.so.
 *
 * Plain-C restatement of the reference's bundle-adjustment hot path for sizes the Python loops cannot reach:
 *   residual   r = proj - obs with TransFunction.from_ray_to_image's own algebra
 *              (slam_system/transformation.py:99-135: tan/atan/sqrt round trip, ~20 libm calls per projection),
 *              observation loop of slam_system/bundle_adjustment.py:67-98 in flat form;
 *   Jacobian   analytic 2x3 / 2x2 blocks - the reference lets scipy form a dense
 *              forward-difference Jacobian (bundle_adjustment.py:200-202), infeasible beyond a few keyframes;
 *   assembly   J^T J / J^T r into per-keyframe U (6), g_c (3) and per-landmark V (3), g_l (2) blocks.
 * Parity pinning: checked against oracle/ptz_oracle.py (itself pinned to reference goldens) in tests/test_oracle_c.py.
 * Parallelised over landmark ranges with pthreads (observations are landmark-major), keyframe blocks reduced per thread.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define DEG2RAD 0.017453292519943295

/* transformation.py:99-135 */
static void from_ray_to_image(double u, double v, double f, double c_p, double c_t, double p, double t, double* x, double* y) {
    const double pan = p * DEG2RAD, tilt = t * DEG2RAD, cp = c_p * DEG2RAD, ct = c_t * DEG2RAD;
    const double tp = tan(pan), tt = tan(tilt), sq = sqrt(tp * tp + 1);
    const double num_x = tp * cos(cp) - sin(cp);
    const double den = tp * sin(cp) * cos(ct) + tt * sq * sin(ct) + cos(ct) * cos(cp);
    const double rel_pan = atan(num_x / den);
    const double num_y = -(tp * sin(ct) * sin(cp) - tt * sq * cos(ct) + sin(ct) * cos(cp));
    const double rel_tilt = atan(num_y / sqrt(num_x * num_x + den * den));
    const double dx = f * tan(rel_pan);
    *x = dx + u;
    *y = -sqrt(f * f + dx * dx) * tan(rel_tilt) + v;
}

/* SURVEY.md Appendix A, derivatives per degree / pixel */
static void jac_blocks(double pan, double tilt, double f, double theta, double phi, double jc[6], double jr[4]) {
    const double k = DEG2RAD;
    const double a = (theta - pan) * k, t = tilt * k, th = theta * k, ph = phi * k;
    const double sg = cos(th) < 0 ? -1.0 : 1.0;
    const double T = tan(ph) * sg, S = (1.0 + tan(ph) * tan(ph)) * sg;
    const double sa = sin(a), ca = cos(a), st = sin(t), ct = cos(t);
    const double Ny = -ct * T + st * ca, z = st * T + ct * ca;
    const double px = sa / z, py = Ny / z, fz = f / z;
    const double dxa = fz * (ca + ct * sa * px), dya = fz * sa * (-st + ct * py);
    const double dxt = f * px * py, dyt = f * (1.0 + py * py);
    const double dxp = -fz * px * st * S, dyp = -fz * (ct + st * py) * S;
    jc[0] = -k * dxa; jc[1] = k * dxt; jc[2] = px;
    jc[3] = -k * dya; jc[4] = k * dyt; jc[5] = py;
    jr[0] = k * dxa; jr[1] = k * dxp;
    jr[2] = k * dya; jr[3] = k * dyp;
}

/* ---- pthread work splitting (libgomp is not in this image) -------------------------------------------------------- */
typedef struct {
    int tid, nthr;
    int64_t n_obs;
    const int32_t *cam_idx, *lm_idx;
    const double *obs_xy, *poses, *rays;
    int n_pose, n_lm;
    double u, v;
    double *r, *V, *gl, *priv;
    double cost;
    int fused;
} job_t;

static void* worker(void* arg) {
    job_t* j = (job_t*)arg;
    int64_t b = j->n_obs * j->tid / j->nthr, e = j->n_obs * (j->tid + 1) / j->nthr;
    if (!j->fused) {
        for (int64_t k = b; k < e; ++k) {
            const double* c = j->poses + 3 * (size_t)j->cam_idx[k];
            const double* l = j->rays + 2 * (size_t)j->lm_idx[k];
            double x, y;
            from_ray_to_image(j->u, j->v, c[2], c[0], c[1], l[0], l[1], &x, &y);
            j->r[2 * k] = x - j->obs_xy[2 * k];
            j->r[2 * k + 1] = y - j->obs_xy[2 * k + 1];
        }
        return 0;
    }
    /* split at landmark boundaries so that V/g_l of a landmark belong to one thread */
    const int32_t* lm_idx = j->lm_idx;
    while (b > 0 && b < j->n_obs && lm_idx[b] == lm_idx[b - 1]) ++b;
    while (e > 0 && e < j->n_obs && lm_idx[e] == lm_idx[e - 1]) ++e;
    double* pu = j->priv + (size_t)j->tid * j->n_pose * 9;
    double cost = 0.0;
    for (int64_t k = b; k < e; ++k) {
        const int ci = j->cam_idx[k], li = lm_idx[k];
        const double* c = j->poses + 3 * (size_t)ci;
        const double* l = j->rays + 2 * (size_t)li;
        double x, y, jc[6], jr[4];
        from_ray_to_image(j->u, j->v, c[2], c[0], c[1], l[0], l[1], &x, &y);
        const double rx = x - j->obs_xy[2 * k], ry = y - j->obs_xy[2 * k + 1];
        if (j->r) { j->r[2 * k] = rx; j->r[2 * k + 1] = ry; }
        cost += rx * rx + ry * ry;
        jac_blocks(c[0], c[1], c[2], l[0], l[1], jc, jr);
        double* vv = j->V + 3 * (size_t)li;
        vv[0] += jr[0] * jr[0] + jr[2] * jr[2];
        vv[1] += jr[0] * jr[1] + jr[2] * jr[3];
        vv[2] += jr[1] * jr[1] + jr[3] * jr[3];
        j->gl[2 * (size_t)li] += jr[0] * rx + jr[2] * ry;
        j->gl[2 * (size_t)li + 1] += jr[1] * rx + jr[3] * ry;
        if (ci != 0) {
            double* q = pu + 9 * (size_t)ci;
            q[0] += jc[0] * jc[0] + jc[3] * jc[3];
            q[1] += jc[0] * jc[1] + jc[3] * jc[4];
            q[2] += jc[0] * jc[2] + jc[3] * jc[5];
            q[3] += jc[1] * jc[1] + jc[4] * jc[4];
            q[4] += jc[1] * jc[2] + jc[4] * jc[5];
            q[5] += jc[2] * jc[2] + jc[5] * jc[5];
            q[6] += jc[0] * rx + jc[3] * ry;
            q[7] += jc[1] * rx + jc[4] * ry;
            q[8] += jc[2] * rx + jc[5] * ry;
        }
    }
    j->cost = cost;
    return 0;
}

int oracle_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

static void run_jobs(job_t* proto, int nthr) {
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthr);
    job_t* jobs = (job_t*)malloc(sizeof(job_t) * nthr);
    for (int t = 0; t < nthr; ++t) {
        jobs[t] = *proto;
        jobs[t].tid = t;
        jobs[t].nthr = nthr;
        jobs[t].cost = 0;
        if (t > 0) pthread_create(&th[t], 0, worker, &jobs[t]);
    }
    worker(&jobs[0]);
    proto->cost = jobs[0].cost;
    for (int t = 1; t < nthr; ++t) {
        pthread_join(th[t], 0);
        proto->cost += jobs[t].cost;
    }
    free(th);
    free(jobs);
}

/* residual only, caller order (bundle_adjustment.py:67-98 in flat form) */
void oracle_ba_residual(int64_t n_obs, const int32_t* cam_idx, const int32_t* lm_idx, const double* obs_xy,
                        const double* poses, const double* rays, double u, double v, double* r, int n_threads) {
    job_t j;
    memset(&j, 0, sizeof(j));
    j.n_obs = n_obs; j.cam_idx = cam_idx; j.lm_idx = lm_idx; j.obs_xy = obs_xy; j.poses = poses; j.rays = rays;
    j.u = u; j.v = v; j.r = r; j.fused = 0;
    run_jobs(&j, n_threads > 0 ? n_threads : oracle_max_threads());
}

/* fused pass; lm_idx must be non-decreasing.  U[N*6] (pp,pt,pf,tt,tf,ff), gc[N*3], V[M*3] (tt,tp,pp), gl[M*2].
 * Keyframe 0 is the fixed reference pose: its U/gc stay zero.  Returns 0.5*sum r^2. */
double oracle_ba_fused(int64_t n_obs, const int32_t* cam_idx, const int32_t* lm_idx, const double* obs_xy,
                       int n_pose, int n_lm, const double* poses, const double* rays, double u, double v, double* r,
                       double* U, double* gc, double* V, double* gl, int n_threads) {
    const int nt = n_threads > 0 ? n_threads : oracle_max_threads();
    memset(U, 0, sizeof(double) * 6 * (size_t)n_pose);
    memset(gc, 0, sizeof(double) * 3 * (size_t)n_pose);
    memset(V, 0, sizeof(double) * 3 * (size_t)n_lm);
    memset(gl, 0, sizeof(double) * 2 * (size_t)n_lm);
    job_t j;
    memset(&j, 0, sizeof(j));
    j.n_obs = n_obs; j.cam_idx = cam_idx; j.lm_idx = lm_idx; j.obs_xy = obs_xy; j.poses = poses; j.rays = rays;
    j.n_pose = n_pose; j.n_lm = n_lm; j.u = u; j.v = v; j.r = r; j.V = V; j.gl = gl; j.fused = 1;
    j.priv = (double*)calloc((size_t)nt * n_pose * 9, sizeof(double));
    run_jobs(&j, nt);
    for (int t = 0; t < nt; ++t)
        for (int c = 0; c < n_pose; ++c) {
            const double* q = j.priv + ((size_t)t * n_pose + c) * 9;
            for (int e = 0; e < 6; ++e) U[6 * (size_t)c + e] += q[e];
            for (int e = 0; e < 3; ++e) gc[3 * (size_t)c + e] += q[6 + e];
        }
    free(j.priv);
    return 0.5 * j.cost;
}
