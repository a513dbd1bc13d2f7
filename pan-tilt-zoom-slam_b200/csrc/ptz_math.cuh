// ptz_math.cuh - device-side PTZ camera geometry (FP64).
//
// Follows the camera model of the reference:
//   PTZCamera.project_ray          slam_system/ptz_camera.py:191-210   x = K (R_tilt R_pan d + disp(f))
//   PTZCamera.back_project_to_ray  slam_system/ptz_camera.py:287-312
//   TransFunction.from_ray_to_image / from_image_to_ray  slam_system/transformation.py:99-175  (disp = 0 twins)
// with d = (tan th, -tan ph * sqrt(tan^2 th + 1), 1), R_pan = [[c,0,-s],[0,1,0],[s,0,c]], R_tilt = [[1,0,0],[0,c,s],[0,-s,c]].
// All API angles are degrees; DEG2RAD mirrors math.radians (x * pi/180).
#pragma once
#include <cuda_runtime.h>

#define PTZ_DEG2RAD 0.017453292519943295   // pi / 180
#define PTZ_RAD2DEG 57.29577951308232      // 180 / pi

struct CamFull {          // one camera with optional displacement, trig already evaluated
    double sp, cp, st, ct; // sin/cos pan, sin/cos tilt
    double f, u, v;
    double d0, d1, d2;     // disp(f) = (l0 + l3 f, l1 + l4 f, l2 + l5 f)   ptz_camera.py:106-115
};

__device__ __forceinline__ CamFull make_cam(double pan, double tilt, double f, double u, double v,
                                            const double* __restrict__ lam /* 6 or nullptr */) {
    CamFull c;
    sincos(pan * PTZ_DEG2RAD, &c.sp, &c.cp);
    sincos(tilt * PTZ_DEG2RAD, &c.st, &c.ct);
    c.f = f; c.u = u; c.v = v;
    if (lam) {
        c.d0 = lam[0] + lam[3] * f;
        c.d1 = lam[1] + lam[4] * f;
        c.d2 = lam[2] + lam[5] * f;
    } else {
        c.d0 = c.d1 = c.d2 = 0.0;
    }
    return c;
}

// general projection (displacement allowed).  q2 is the homogeneous depth (reference asserts q2 != 0).
__device__ __forceinline__ void project_full(const CamFull& c, double theta_deg, double phi_deg,
                                             double& x, double& y, double& q2) {
    const double tx = tan(theta_deg * PTZ_DEG2RAD);
    const double tp = tan(phi_deg * PTZ_DEG2RAD);
    const double r0 = tx;
    const double r1 = -tp * sqrt(tx * tx + 1.0);
    // a = R_pan * (r0, r1, 1)
    const double a0 = c.cp * r0 - c.sp;
    const double a2 = c.sp * r0 + c.cp;
    // q = R_tilt * a + disp
    const double q0 = a0 + c.d0;
    const double q1 = c.ct * r1 + c.st * a2 + c.d1;
    q2 = -c.st * r1 + c.ct * a2 + c.d2;
    const double iz = 1.0 / q2;
    x = c.f * q0 * iz + c.u;
    y = c.f * q1 * iz + c.v;
}

// back-projection pixel -> (theta, phi) degrees, displacement allowed (R^-1 = R^T, K^-1 in closed form)
__device__ __forceinline__ void backproject_full(const CamFull& c, double x, double y, double& theta_deg,
                                                 double& phi_deg) {
    const double q0 = (x - c.u) / c.f - c.d0;
    const double q1 = (y - c.v) / c.f - c.d1;
    const double q2 = 1.0 - c.d2;
    const double a1 = c.ct * q1 - c.st * q2;
    const double a2 = c.st * q1 + c.ct * q2;
    const double r0 = c.cp * q0 + c.sp * a2;
    const double r2 = -c.sp * q0 + c.cp * a2;
    theta_deg = atan(r0 / r2) * PTZ_RAD2DEG;
    phi_deg = atan(-a1 / sqrt(r0 * r0 + r2 * r2)) * PTZ_RAD2DEG;
}

// ---------------------------------------------------------------------------------------------------------------
// disp = 0 fast form used by bundle adjustment: no transcendental per observation.
//   per keyframe : sp, cp, st, ct, f                       (CamTrig)
//   per landmark : sin th, cos th, T = tan ph * sgn(cos th), S = (1 + tan^2 ph) * sgn(cos th)     (LmTrig)
//   alpha = th - pan:  sa = sth cp - cth sp, ca = cth cp + sth sp
//   Nx = sa, Ny = -ct T + st ca, z = st T + ct ca, x = u + f Nx / z, y = v + f Ny / z
// ---------------------------------------------------------------------------------------------------------------
struct CamTrig { double sp, cp, st, ct, f; };
struct __align__(32) LmTrig { double sth, cth, T, S; };   // one 256-bit load (LDG.E.ENL2.256) per gather

__device__ __forceinline__ CamTrig make_cam_trig(double pan, double tilt, double f) {
    CamTrig c;
    sincos(pan * PTZ_DEG2RAD, &c.sp, &c.cp);
    sincos(tilt * PTZ_DEG2RAD, &c.st, &c.ct);
    c.f = f;
    return c;
}

__device__ __forceinline__ LmTrig make_lm_trig(double theta_deg, double phi_deg) {
    LmTrig l;
    sincos(theta_deg * PTZ_DEG2RAD, &l.sth, &l.cth);
    double sph, cph;
    sincos(phi_deg * PTZ_DEG2RAD, &sph, &cph);            // tan = sin / cos: one range reduction instead of tan()'s own
    const double tp = sph / cph;
    const double sg = (l.cth < 0.0) ? -1.0 : 1.0;   // sqrt(tan^2+1) = |sec|  (ptz_camera.py:205)
    l.T = tp * sg;
    l.S = (1.0 + tp * tp) * sg;
    return l;
}

struct ObsGeom {        // everything one observation needs, in RADIAN derivative units
    double px, py;      // Nx/z, Ny/z   (= d x/d f, d y/d f)
    double xa, ya;      // d(x,y)/d alpha   (d/d theta = +, d/d pan = -)
    double xt, yt;      // d(x,y)/d tilt
    double xp, yp;      // d(x,y)/d phi
};

// 1 / z without the IEEE special-case branch of the compiler's division (z is a homogeneous depth: finite, non-zero, far
// from the denormal range): hardware seed (>= 20 bits) + two Newton steps in FMA form = full double accuracy (<= 1 ulp),
// 5 instructions and no BSSY/BSYNC region that would stop the scheduler from interleaving neighbouring observations.
__device__ __forceinline__ double fast_rcp(double z) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(z));
    double e = fma(-z, r, 1.0);
    r = fma(r, e, r);
    e = fma(-z, r, 1.0);
    return fma(r, e, r);
}

// shared front end of project_fast / project_fast_jac.  Every product that feeds an addition is written as an explicit
// __dmul_rn / fma, so that the compiler's FMA contraction cannot differ between kernels: the residual-only pass and the
// fused pass return bit-identical residuals (the trust-region loop compares costs computed by the two).
struct ProjCore { double sa, ca, Ny, iz, px, py; };
__device__ __forceinline__ ProjCore project_core(const CamTrig& c, const LmTrig& l) {
    ProjCore p;
    p.sa = fma(l.sth, c.cp, -__dmul_rn(l.cth, c.sp));
    p.ca = fma(l.cth, c.cp, __dmul_rn(l.sth, c.sp));
    p.Ny = fma(c.st, p.ca, -__dmul_rn(c.ct, l.T));
    const double z = fma(c.st, l.T, __dmul_rn(c.ct, p.ca));
    p.iz = fast_rcp(z);
    p.px = __dmul_rn(p.sa, p.iz);
    p.py = __dmul_rn(p.Ny, p.iz);
    return p;
}

__device__ __forceinline__ void project_fast(const CamTrig& c, const LmTrig& l, double u, double v, double& x,
                                             double& y) {
    const ProjCore p = project_core(c, l);
    x = fma(c.f, p.px, u);
    y = fma(c.f, p.py, v);
}

__device__ __forceinline__ void project_fast_jac(const CamTrig& c, const LmTrig& l, double u, double v, double& x,
                                                 double& y, ObsGeom& g) {
    const ProjCore p = project_core(c, l);
    g.px = p.px;
    g.py = p.py;
    x = fma(c.f, g.px, u);
    y = fma(c.f, g.py, v);
    const double fz = c.f * p.iz;
    g.xa = fz * fma(c.ct * p.sa, g.px, p.ca);        // f (ca z + ct sa^2) / z^2
    g.ya = fz * p.sa * fma(c.ct, g.py, -c.st);       // f sa (-st z + ct Ny) / z^2
    g.xt = c.f * g.px * g.py;                        // f Nx Ny / z^2
    g.yt = fma(c.f * g.py, g.py, c.f);               // f (1 + Ny^2/z^2)
    const double fzS = -fz * l.S;
    g.xp = fzS * g.px * c.st;                        // -f Nx st S / z^2
    g.yp = fzS * fma(c.st, g.py, c.ct);              // -f (ct z + st Ny) S / z^2
}
