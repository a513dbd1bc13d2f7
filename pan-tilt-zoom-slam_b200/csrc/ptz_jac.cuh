// ptz_jac.cuh - measurement Jacobian blocks d(x,y)/d(pan,tilt,f) (2x3) and d(x,y)/d(theta,phi) (2x2) of one ray,
// shared by projection.cu (compute_h_jacobian entry points) and ekf.cu (EKF update).
//   PTZBA_JAC_ANALYTIC    closed form, displacement allowed
//   PTZBA_JAC_CENTRAL_FD  the reference's central differences operation by operation (ptz_slam.py:87-136):
//                         delta = 0.001 deg for angles, 0.1 px for f, 10 projections per ray
#pragma once
#include "ptz_math.cuh"
#include "../../include/ptzba.h"

#define PTZ_FD_DELTA_ANGLE 0.001   // ptz_slam.py:87
#define PTZ_FD_DELTA_F 0.1         // ptz_slam.py:88

// camera variants used by the central differences: 0 base, 1/2 pan -/+, 3/4 tilt -/+, 5/6 f -/+
__device__ __forceinline__ CamFull h_cam_variant(int which, double p, double t, double f, double u, double v,
                                                 const double* __restrict__ disp) {
    const double da = PTZ_FD_DELTA_ANGLE, df = PTZ_FD_DELTA_F;
    switch (which) {
        case 1: p = p - da; break;
        case 2: p = p + da; break;
        case 3: t = t - da; break;
        case 4: t = t + da; break;
        case 5: f = f - df; break;
        case 6: f = f + df; break;
        default: break;
    }
    return make_cam(p, t, f, u, v, disp);
}

// analytic, general displacement: q = R_tilt R_pan d + disp(f); x = f q0/q2 + u, y = f q1/q2 + v.
__device__ __forceinline__ void jac_analytic_full(const CamFull& c, const double* __restrict__ lam, double th_deg,
                                                  double ph_deg, double* jc, double* jr) {
    const double k = PTZ_DEG2RAD;
    const double tx = tan(th_deg * k), tp = tan(ph_deg * k);
    const double sec2t = 1.0 + tx * tx, sec2p = 1.0 + tp * tp;
    const double sq = sqrt(sec2t);
    const double r0 = tx, r1 = -tp * sq;
    const double a0 = c.cp * r0 - c.sp, a2 = c.sp * r0 + c.cp;
    const double q0 = a0 + c.d0;
    const double q1 = c.ct * r1 + c.st * a2 + c.d1;
    const double q2 = -c.st * r1 + c.ct * a2 + c.d2;
    const double iz = 1.0 / q2;
    const double px = q0 * iz, py = q1 * iz;
    // d q / d pan (rad): d a0 = -sp r0 - cp = -a2 ; d a2 = cp r0 - sp = a0
    const double q0p = -a2, q1p = c.st * a0, q2p = c.ct * a0;
    // d q / d tilt (rad): q1 = ct r1 + st a2 -> -st r1 + ct a2 = q2 - d2 ; q2 -> -ct r1 - st a2 = -(q1 - d1)
    const double q1t = q2 - c.d2, q2t = -(q1 - c.d1);
    // d q / d f : disp derivative (l3, l4, l5)
    const double l3 = lam ? lam[3] : 0.0, l4 = lam ? lam[4] : 0.0, l5 = lam ? lam[5] : 0.0;
    // d q / d theta (rad): d r0 = sec2t ; d r1 = -tp * tx * sec2t / sq = -tp tx sq
    const double r0h = sec2t, r1h = -tp * tx * sq;
    const double q0h = c.cp * r0h, a2h = c.sp * r0h;
    const double q1h = c.ct * r1h + c.st * a2h, q2h = -c.st * r1h + c.ct * a2h;
    // d q / d phi (rad): d r1 = -sec2p sq
    const double r1f = -sec2p * sq;
    const double q1f = c.ct * r1f, q2f = -c.st * r1f;
    const double fi = c.f * iz;
#define DX(dq0, dq2) (fi * ((dq0) - px * (dq2)))
#define DY(dq1, dq2) (fi * ((dq1) - py * (dq2)))
    jc[0] = k * DX(q0p, q2p);  jc[3] = k * DY(q1p, q2p);
    jc[1] = k * DX(0.0, q2t);  jc[4] = k * DY(q1t, q2t);
    jc[2] = px + DX(l3, l5);   jc[5] = py + DY(l4, l5);
    jr[0] = k * DX(q0h, q2h);  jr[2] = k * DY(q1h, q2h);
    jr[1] = k * DX(0.0, q2f);  jr[3] = k * DY(q1f, q2f);
#undef DX
#undef DY
}


// jc[6] row-major 2x3 (cols pan,tilt,f), jr[4] row-major 2x2 (cols theta,phi); cams = the 7 variants above
__device__ __forceinline__ void h_blocks_eval(const CamFull* cams, const double* __restrict__ disp, double th, double ph,
                                              int mode, double* jc, double* jr) {
    if (mode == PTZBA_JAC_ANALYTIC) {
        jac_analytic_full(cams[0], disp, th, ph, jc, jr);
        return;
    }
    const double da = PTZ_FD_DELTA_ANGLE, df = PTZ_FD_DELTA_F;
    double x1, y1, x2, y2, q;
    project_full(cams[1], th, ph, x1, y1, q);
    project_full(cams[2], th, ph, x2, y2, q);
    jc[0] = (x2 - x1) / (2 * da); jc[3] = (y2 - y1) / (2 * da);
    project_full(cams[3], th, ph, x1, y1, q);
    project_full(cams[4], th, ph, x2, y2, q);
    jc[1] = (x2 - x1) / (2 * da); jc[4] = (y2 - y1) / (2 * da);
    project_full(cams[5], th, ph, x1, y1, q);
    project_full(cams[6], th, ph, x2, y2, q);
    jc[2] = (x2 - x1) / (2 * df); jc[5] = (y2 - y1) / (2 * df);
    project_full(cams[0], th - da, ph, x1, y1, q);
    project_full(cams[0], th + da, ph, x2, y2, q);
    jr[0] = (x2 - x1) / (2 * da); jr[2] = (y2 - y1) / (2 * da);
    project_full(cams[0], th, ph - da, x1, y1, q);
    project_full(cams[0], th, ph + da, x2, y2, q);
    jr[1] = (x2 - x1) / (2 * da); jr[3] = (y2 - y1) / (2 * da);
}
