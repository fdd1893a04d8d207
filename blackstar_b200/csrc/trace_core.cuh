// trace_core.cuh -- per-ray arithmetic of the geodesic tracer (K1) and the sky lookup (K2).
//
// Written as __host__ __device__ so the SAME source is (a) inlined into the sm_100a kernels
// of trace_kernel.cu and (b) compiled for the host by tests/hostcheck (a test-only harness
// that diffs this arithmetic against the oracle without a GPU).  The shipped library never
// runs the host instantiation: there is no CPU fallback.
//
// Reference semantics: src/Raytracer.hs:34-134, src/StarMap.hs:93-115.
//
// B200-first restructuring.  The kernel is bound by FP64 issue, so the
// work per RK4 step is cut from the reference's 141 flops (4 sqrt + 4 div) to 51 DP
// instructions + 4 MUFU without changing the discrete map:
//  * every ray's motion is planar (the force is central) and classical RK4 commutes with
//    linear changes of variables, so the 6-double state (vel, pos) is integrated as 4 doubles
//    (u, v, du, dv) in an orthonormal basis (f1, f2) of the ray's own orbital plane.  In exact
//    arithmetic this is the SAME discrete map as the reference's 3-D RK4 (same truncation
//    error); only rounding (1e-16 per step) differs.
//  * f2 is chosen horizontal, so scene-y = f1y * u: the disk-crossing test (signum y' /=
//    signum y) is a sign-bit comparison of u, no FP64 work;
//  * lengths are divided by L = (3.75 h2 T^2)^(1/5) per ray and time by T = stepSize/2, which
//    makes the force constant -0.4 -- exactly the factor the |pos|^-5 primitive leaves out (one
//    multiply less per force evaluation, one instruction less in the correction) -- and the step 2,
//    so that the step constants are 1, 2, 1/3 and 2/3 (two registers instead of seven);
//  * |pos|^-5 comes from one MUFU.RSQ64H seed and a first-order correction in 5 DP
//    instructions instead of sqrt, three multiplies and a divide (force error <= 1.4e-11);
//  * the stage velocities are eliminated algebraically (p3 = p2 - (h/2)^2 a1, ...) and only a1
//    is ever formed: the other stage accelerations enter the two weighted sums as FMAs;
//  * horizon / escape tests compare the bit patterns of positive doubles as integers.
#pragma once

#include "bsb_common.cuh"

#include <cmath>

namespace bsb {

// ---- correctly-rounded primitives that must never be contracted into FMAs ------------
#if defined(__CUDA_ARCH__)
BSB_D double mul_rn(double a, double b) { return __dmul_rn(a, b); }
BSB_D double add_rn(double a, double b) { return __dadd_rn(a, b); }
BSB_D double sub_rn(double a, double b) { return __dsub_rn(a, b); }
BSB_D double div_rn(double a, double b) { return __ddiv_rn(a, b); }
BSB_D double sqrt_rn(double a) { return __dsqrt_rn(a); }
BSB_D double fma_(double a, double b, double c) { return __fma_rn(a, b, c); }
#else
// host instantiation (tests/hostcheck only), compiled with -ffp-contract=off
inline double mul_rn(double a, double b) { return a * b; }
inline double add_rn(double a, double b) { return a + b; }
inline double sub_rn(double a, double b) { return a - b; }
inline double div_rn(double a, double b) { return a / b; }
inline double sqrt_rn(double a) { return std::sqrt(a); }
inline double fma_(double a, double b, double c) { return std::fma(a, b, c); }
#endif

// Seed for x^-1/2.  Device: MUFU.RSQ64H via rsqrt.approx.ftz.f64 (looks at the high word of
// x only; ~20 good bits).  Host: an emulation with the same information loss, so hostcheck
// exercises the correction polynomial at the accuracy the hardware seed has.
BSB_HD double rsqrt_seed(double x)
{
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
#else
    union { double d; uint64_t u; } a, b;
    a.d = x;
    a.u &= 0xFFFFFFFF00000000ull;             // hardware reads the high 32 bits only
    b.d = (double)(float)(1.0 / std::sqrt(a.d)); // ~24-bit result
    b.u &= 0xFFFFFFFF00000000ull;
    return b.d;
#endif
}

// 0.4 q^(-5/2) (the factor 0.4 is absorbed by the ray's length scale, see ray_frame).  With
// y0 = q^-1/2 (1+d) from the seed and e = 1 - q y0^2:
//   q^(-5/2) = y0^5 (1-e)^(-5/2) = y0^5 (1 + 5/2 e + 35/8 e^2 + ...) ~= 2.5 y0^5 (1.4 - q y0^2).
// The dropped term is 4.375 e^2: |e| <= 2^-19.1 measured on the device (bsb_selftest_rinv5), so
// the force is low by at most 1.4e-11 relative -- as if h2 were that much smaller, i.e. a deflection
// error of ~1e-11 rad against the ~1e-7 rad that the 1e-4 parity bar on the star Gaussians allows
// (SURVEY.md S5).  5 FP64 instructions + 1 MUFU.
// `k14` must hold 1.4 in a register / the constant bank: an FP64 immediate carries only a high word.
// |p|^2 and the seed for its inverse square root in one go.  (Tried and rejected: letting the seed keep
// whatever low word its register pair held, which saves the "clear the low word" move ptxas emits for
// every MUFU.RSQ64H -- 3 instructions per RK4 step fewer, and 3.7 % SLOWER on the B200: the kernel is bound
// by the FP64 pipe and its register-operand traffic, not by issue slots; profiles/r02_trace_variants.txt.)
BSB_HD void norm2_seed(double pu, double pv, double &q, double &y0)
{
    q = fma_(pu, pu, pv * pv);
    y0 = rsqrt_seed(q);
}

// 0.4 q^(-5/2) from q and a seed y0 ~ q^-1/2 (see rinv5)
BSB_HD double rinv5_seeded(double q, double y0, double k14)
{
    const double s = y0 * y0;
    const double c = fma_(-q, s, k14);
    const double s2 = s * s;        // this association, (s^2 y0) c, is the fastest of the four that were
    const double y5 = s2 * y0;      // timed (profiles/r02_trace_variants.txt): up to 4 % between them
    return y5 * c;
}

BSB_HD long long dbits(double x)
{
#if defined(__CUDA_ARCH__)
    return __double_as_longlong(x);
#else
    union { double d; long long i; } t;
    t.d = x;
    return t.i;
#endif
}

BSB_HD int hi32(double x)
{
#if defined(__CUDA_ARCH__)
    return __double2hiint(x);
#else
    return (int)(dbits(x) >> 32);
#endif
}

// 0 for a clear sign bit, -1 for a set one
BSB_HD int sign_code(double x) { return hi32(x) >> 31; }

// ---- ray generation: src/Raytracer.hs:40-51, bit-exact with the reference's op order ----
BSB_HD double pixel_vx(const FrameParams &P, int x)   // :49  fov * (x/w - 0.5)
{
    return mul_rn(P.fov, sub_rn(div_rn((double)x, (double)P.W2), 0.5));
}
BSB_HD double pixel_vy(const FrameParams &P, int y)   // :50  ((fov * (0.5 - y/h)) * h) / w
{
    const double w = (double)P.W2, h = (double)P.H2;
    return div_rn(mul_rn(mul_rn(P.fov, sub_rn(0.5, div_rn((double)y, h))), h), w);
}

BSB_HD void ray_direction(const FrameParams &P, int x, int y, double dir[3])
{
    const double vx = P.vx_tab ? P.vx_tab[x] : pixel_vx(P, x);
    const double vy = P.vy_tab ? P.vy_tab[y] : pixel_vy(P, y);
    double u[3];
#pragma unroll
    for (int k = 0; k < 3; k++)  // :48  (xa_i*vx + ya_i*vy) + (-za_i)*(-1)
        u[k] = add_rn(add_rn(mul_rn(P.xa[k], vx), mul_rn(P.ya[k], vy)), P.za[k]);
    // Linear.normalize: unchanged if |v|^2 is within 1e-12 of 0 or 1
    const double l = add_rn(add_rn(mul_rn(u[0], u[0]), mul_rn(u[1], u[1])), mul_rn(u[2], u[2]));
    if (fabs(l) <= 1e-12 || fabs(sub_rn(1.0, l)) <= 1e-12) {
        dir[0] = u[0]; dir[1] = u[1]; dir[2] = u[2];
    } else {
        const double s = sqrt_rn(l);
        dir[0] = div_rn(u[0], s); dir[1] = div_rn(u[1], s); dir[2] = div_rn(u[2], s);
    }
}

// x^(-1/2) for the per-ray frame (device: MUFU seed + Newton, no division; host: 1 / sqrt)
BSB_HD double inv_sqrt(double x)
{
#if defined(__CUDA_ARCH__)
    return rsqrt(x);
#else
    return 1.0 / std::sqrt(x);
#endif
}

// a^(-1/5) to ~2 ulp: single-precision seed (relative error d ~ 1e-6), two Newton steps
// y <- y (1.2 - 0.2 a y^5) (no division; error 3 d^2 per step: 3e-12, then 3e-23).  Replaces
// pow(a, 0.2) in the per-ray set-up.
BSB_HD double inv_fifth_root(double a)
{
#if defined(__CUDA_ARCH__)
    double y = (double)exp2f(-0.2f * __log2f((float)a));
#else
    double y = (double)std::pow((float)a, -0.2f);
#endif
#pragma unroll
    for (int it = 0; it < 2; it++) {
        const double y2 = y * y;
        const double y5 = (y2 * y2) * y;
        y = y * fma_(-0.2 * a, y5, 1.2);
    }
    return y;
}

// State of one ray between step blocks.  Coordinates are in the ray's own orbital plane,
// basis (f1, f2) with f2 HORIZONTAL (so scene-y = f1y * u and a disk-plane crossing is a sign
// change of u), divided by L = (3.75 h2 T^2)^(1/5) with time in units of T = stepSize / 2, so that
// the equation of motion is p'' = -0.4 p / |p|^5 and the step is 2 for every ray (classical RK4 commutes with this linear change of
// variables, so it is still the reference's discrete map).
struct RayState {
    double u, v;       // position / L in the (f1, f2) basis
    double du, dv;     // velocity / L, per HALF STEP (time unit = stepSize / 2)
    double q;          // u^2 + v^2
    double qh, qs;     // horizon and escape thresholds on q: 1/L^2, safe2/L^2
    double acc[4];     // colour accumulated front-to-back (premultiplied RGBA), Raytracer.hs:86
    uint32_t steps;
    int32_t status;    // kAlive / kBlack / kSky / kCapped
    int32_t side;      // sign code (0 / -1) that u has on the current side of the disk plane;
                       // kSideZero = the reference's signum y is 0 here (any move is a crossing);
                       // kSideNever = this ray's plane is the disk plane (signum y stays 0)
};
enum : int32_t { kSideZero = 1, kSideNever = 2 };
enum : int32_t { kAlive = 0, kBlack = 1, kSky = 2, kCapped = 3, kIdle = 4 };

struct RayFrame {
    double f1[3], f2[3];
    double L;          // length scale (3.75 h2 (stepSize/2)^2)^(1/5)
    double iL;         // 1 / L
    double qh;         // 1 / L^2
    int32_t ysign;
};

// Orbital-plane frame of the ray through `dir`.  n = cam x dir is the plane normal
// (|n|^2 = h2, formed exactly as src/Raytracer.hs:73 forms it).
BSB_HD void ray_frame(const FrameParams &P, const double dir[3], RayFrame &F)
{
    const double n0 = sub_rn(mul_rn(P.cam[1], dir[2]), mul_rn(P.cam[2], dir[1]));
    const double n1 = sub_rn(mul_rn(P.cam[2], dir[0]), mul_rn(P.cam[0], dir[2]));
    const double n2 = sub_rn(mul_rn(P.cam[0], dir[1]), mul_rn(P.cam[1], dir[0]));
    const double h2 = add_rn(add_rn(mul_rn(n0, n0), mul_rn(n1, n1)), mul_rn(n2, n2));  // :73
    // p'' = -1.5 h2 p/|p|^5.  With lengths in units of L and time in units of T = stepSize/2 this is
    // p~'' = -(1.5 h2 T^2 / L^5) p~/|p~|^5.  The kernel's |p|^-5 primitive returns 0.4 |p~|^-5, so the
    // constant wanted is 0.4: L^5 = 3.75 h2 T^2
    double l5 = (3.75 * h2) * P.hh2;
    if (!(l5 > 1e-30)) l5 = 1e-30;    // (near-)radial ray: the force is ~1e-30 of anything else either way
    const double iL = inv_fifth_root(l5);
    const double iL2 = iL * iL;
    F.L = l5 * (iL2 * iL2);           // a^(1/5) = a * a^(-4/5)
    F.iL = iL;
    F.qh = iL2;
    const double s2 = n0 * n0 + n2 * n2;
    // h2 below ~1e-24 |cam|^2 is rounding noise of the cross product (the ray points at the hole to
    // within 1e-12 rad): n is then meaningless as a plane normal, and the motion is a line anyway.
    const bool radial = !(h2 > 1e-24 * P.q0);
    if (radial || !(s2 > 1e-28 * h2)) {
        // radial ray (no plane) or a plane that IS the disk plane: f1 = cam/|cam|, f2 any
        // in-plane unit vector orthogonal to it.
        F.f1[0] = P.e1[0]; F.f1[1] = P.e1[1]; F.f1[2] = P.e1[2];
        double w0, w1, w2;
        if (!radial) {  // f2 = nhat x f1
            const double in = inv_sqrt(h2);
            w0 = (n1 * P.e1[2] - n2 * P.e1[1]) * in;
            w1 = (n2 * P.e1[0] - n0 * P.e1[2]) * in;
            w2 = (n0 * P.e1[1] - n1 * P.e1[0]) * in;
        } else {
            const double ax = fabs(P.e1[0]), ay = fabs(P.e1[1]), az = fabs(P.e1[2]);
            double t0 = 0, t1 = 0, t2 = 0;
            if (ax <= ay && ax <= az) t0 = 1; else if (ay <= az) t1 = 1; else t2 = 1;
            const double d = t0 * P.e1[0] + t1 * P.e1[1] + t2 * P.e1[2];
            w0 = t0 - d * P.e1[0]; w1 = t1 - d * P.e1[1]; w2 = t2 - d * P.e1[2];
        }
        const double iw = inv_sqrt(w0 * w0 + w1 * w1 + w2 * w2);
        F.f2[0] = w0 * iw; F.f2[1] = w1 * iw; F.f2[2] = w2 * iw;
        F.ysign = !radial ? 0 : ((P.e1[1] > 0.0) - (P.e1[1] < 0.0));
        return;
    }
    // general case: f2 along the line of nodes (n x yhat), f1 = nhat x f2; then f1y = -s/|n| < 0
    const double in = inv_sqrt(h2);
    const double nh0 = n0 * in, nh1 = n1 * in, nh2 = n2 * in;
    double g0 = -n2, g1 = 0.0, g2 = n0;
    const double dp = g0 * nh0 + g2 * nh2;                 // re-orthogonalise against nhat
    g0 -= dp * nh0; g1 -= dp * nh1; g2 -= dp * nh2;
    const double ig = inv_sqrt(g0 * g0 + g1 * g1 + g2 * g2);
    g0 *= ig; g1 *= ig; g2 *= ig;
    F.f2[0] = g0; F.f2[1] = g1; F.f2[2] = g2;
    F.f1[0] = nh1 * g2 - nh2 * g1;
    F.f1[1] = nh2 * g0 - nh0 * g2;
    F.f1[2] = nh0 * g1 - nh1 * g0;
    F.ysign = -1;
}

// traceRay's setup (src/Raytracer.hs:69-75): ray, h2 = |pos x vel|^2, acc = 0.  The frame F is
// handed back so the caller can keep it (shared memory in the kernels) for ray_finish.
BSB_HD void ray_init(const FrameParams &P, int x, int y, RayState &s, RayFrame &F)
{
    double dir[3];
    ray_direction(P, x, y, dir);
    ray_frame(P, dir, F);
    const double iL = F.iL;
    s.u = (P.cam[0] * F.f1[0] + P.cam[1] * F.f1[1] + P.cam[2] * F.f1[2]) * iL;
    s.v = (P.cam[0] * F.f2[0] + P.cam[1] * F.f2[1] + P.cam[2] * F.f2[2]) * iL;
    const double iLT = iL * P.hh;     // velocities: length / L per half step
    s.du = (dir[0] * F.f1[0] + dir[1] * F.f1[1] + dir[2] * F.f1[2]) * iLT;
    s.dv = (dir[0] * F.f2[0] + dir[1] * F.f2[1] + dir[2] * F.f2[2]) * iLT;
    s.q = P.q0 * F.qh;               // the reference tests quadrance(cam) on the first step
    s.qh = F.qh;
    s.qs = P.safe2 * F.qh;
    s.acc[0] = s.acc[1] = s.acc[2] = s.acc[3] = 0.0;
    s.steps = 0;
    s.status = kAlive;
    // signum of the exact starting y (:96), expressed as the sign u must have on that side
    const int sy = (P.cam[1] > 0.0) - (P.cam[1] < 0.0);
    if (F.ysign == 0) s.side = kSideNever;            // y = 0 along the whole ray: signum never changes
    else if (sy == 0) s.side = kSideZero;             // signum 0 /= signum y' on the first step
    else s.side = (sy * F.ysign > 0) ? 0 : -1;
}

// massiv-io HSI -> RGB (see oracle/oracle_thirdparty.c for the statement of the formula)
BSB_HD void hsi_to_rgb(double hp, double s, double i, double rgb[3])
{
    const double pi = 3.141592653589793;
    const double h = hp * 2 * pi;
    const double is = i * s;
    const double second = i - is;
    double r, g, b;
    if (h < 0) {
        r = g = b = NAN;
    } else if (h < 2 * pi / 3) {
        r = i + is * cos(h) / cos(pi / 3 - h);
        b = second;
        g = i + 2 * is + b - r;
    } else if (h < 4 * pi / 3) {
        g = i + is * cos(h - 2 * pi / 3) / cos(h + pi);
        r = second;
        b = i + 2 * is + r - g;
    } else if (h < 2 * pi) {
        b = i + is * cos(h - 4 * pi / 3) / cos(2 * pi - pi / 3 - h);
        g = second;
        r = i + 2 * is + g - b;
    } else {
        r = g = b = NAN;
    }
    rgb[0] = r; rgb[1] = g; rgb[2] = b;
}

// blend (src/Raytracer.hs:34-37): acc is on top, c goes below it; all four channels
BSB_HD void blend_under(double acc[4], const double c[4])
{
    const double t = 1.0 - acc[3];
    acc[0] = fma_(c[0], t, acc[0]);
    acc[1] = fma_(c[1], t, acc[1]);
    acc[2] = fma_(c[2], t, acc[2]);
    acc[3] = fma_(c[3], t, acc[3]);
}

// diskColor' (src/Raytracer.hs:104-111)
BSB_HD void disk_layer(const FrameParams &P, double r2ave, double acc[4])
{
    const double pi = 3.141592653589793;
    const double r = sqrt(r2ave);
    const double qn = (P.r_out - r) / (P.r_out - P.r_in);
    const double intensity = sin(pi * (qn * qn));
    const double c[4] = { P.disk_rgb[0] * intensity, P.disk_rgb[1] * intensity, P.disk_rgb[2] * intensity,
                          intensity * P.disk_opacity };
    blend_under(acc, c);
}

// One classical RK4 step of y' = f(y), f(vel,pos) = (-pos/|pos|^5, vel)  (src/Raytracer.hs:113-134)
// from (u, v, du, dv) with q = u^2 + v^2 to (nu, nv, du, dv) with nq.  51 FP64 + 4 MUFU:
//   4 x (2 for |p|^2 + 5 for g = 0.4|p|^-5) + 23 for the stage positions and the two sums
//   pos' = pos + h vel - h^2/6 (a1 + a2 + a3),   vel' = vel - h/6 (a1 + 2 a2 + 2 a3 + a4),
// with a_i = g_i p_i (MINUS the acceleration).  Only a1 is formed explicitly; a2, a3, a4 enter as
// fused multiply-adds:  S = a1 + g2 p2 + g3 p3,  D = g4 p4 - a1,  vel' = vel - h/6 (2 S + D).
// The stage velocities are eliminated (p3 = p2 - (h/2)^2 a1, p4 = pe - h (h/2) g2 p2).
// TIME is measured in half steps (ray_frame), so h = 2, h/2 = (h/2)^2 = 1, h(h/2) = 2, h^2/6 = h/3 =
// 2/3, h/6 = 1/3: of the step constants only 1/3 and 2/3 need registers (an FP64 instruction can carry
// 1 and 2 as immediates), which is what keeps every loop constant resident at 2 CTAs per SM.
BSB_HD void rk4_step(const FrameParams &P, double u, double v, double q, double &du, double &dv,
                     double &nu, double &nv, double &nq, double k14)
{
    const double k13 = P.k13, k23 = P.k23;
    const double g1 = rinv5_seeded(q, rsqrt_seed(q), k14);
    const double a1u = g1 * u, a1v = g1 * v;
    const double p2u = u + du, p2v = v + dv;
    double q2, y2;
    norm2_seed(p2u, p2v, q2, y2);
    const double g2 = rinv5_seeded(q2, y2, k14);
    const double p3u = p2u - a1u, p3v = p2v - a1v;
    double q3, y3;
    norm2_seed(p3u, p3v, q3, y3);
    const double g3 = rinv5_seeded(q3, y3, k14);
    const double peu = fma_(2.0, du, u), pev = fma_(2.0, dv, v);
    const double c2 = g2 + g2;
    const double p4u = fma_(-c2, p2u, peu), p4v = fma_(-c2, p2v, pev);
    double q4, y4;
    norm2_seed(p4u, p4v, q4, y4);
    const double g4 = rinv5_seeded(q4, y4, k14);
    const double su = fma_(g3, p3u, fma_(g2, p2u, a1u)), sv = fma_(g3, p3v, fma_(g2, p2v, a1v));
    const double du4 = fma_(g4, p4u, -a1u), dv4 = fma_(g4, p4v, -a1v);
    nu = fma_(-k23, su, peu);
    nv = fma_(-k23, sv, pev);
    du = fma_(-k13, fma_(2.0, su, du4), du);     // vel' = vel - 1/3 (2 S + D): 0.5 % faster than two FMAs with 2/3 and 1/3
    dv = fma_(-k13, fma_(2.0, sv, dv4), dv);
    nq = fma_(nu, nu, nv * nv);
}

// Advance one ray by at most `max_steps` RK4 steps (colorize', src/Raytracer.hs:80-85).
// The reference takes the step first and then tests the OLD position; testing first and
// skipping the (unused) last step gives the same result with one step less per ray.
// On this machine every FP64 instruction occupies the issue port for two cycles and every other
// instruction for one, so the fast loop is written for the smallest (2 x FP64 + others): it is
// unrolled over two register sets (A -> B -> A, no state copies), has ONE exit test per step
// (sign flip of u | radius outside the safe band of high words | step budget) and resolves what
// happened -- exactly -- outside the loop.
BSB_HD void ray_advance(const FrameParams &P, RayState &s, uint32_t max_steps)
{
    double ua = s.u, va = s.v, qa = s.q, du = s.du, dv = s.dv;
    double ub = ua, vb = va, qb = qa;
    const double k14 = 1.4;      // as a literal: 0.9 % faster than from the constant bank (profiles/r02_trace_variants.txt)
    // q > 0, so doubles order like their bit patterns.  Fast test on the high words: strictly
    // between the two thresholds' high words => neither the horizon nor the escape test fires.
    const long long qh = dbits(s.qh), qs = dbits(s.qs);
    const int lo_hi = hi32(s.qh) + 1;
    const unsigned span = (unsigned)(hi32(s.qs) - lo_hi);
    const bool disk = P.disk_on != 0 && s.side != kSideNever;
    const int dmask = disk ? (int)0x80000000 : 0;              // sign-bit compare enabled?
    const uint32_t left = P.step_cap > s.steps ? P.step_cap - s.steps : 0u;
    const uint32_t budget = max_steps < left ? max_steps : left;
    uint32_t remaining = budget;
    int32_t status = kAlive;
    // `side` as a word whose sign bit is the sign u has on the current side of the disk plane
    int side_word = (s.side == -1) ? (int)0x80000000 : 0;
    bool first_is_zero = disk && s.side == kSideZero;           // signum y = 0 at the start (:96)
    for (;;) {
        // ---- exact tests on the current position (A)
        {
            const long long qi = dbits(qa);
            if (qi < qh) { status = kBlack; break; }               // :93 passed the horizon
            if (qi > qs) { status = kSky; break; }                 // :94 escaped
        }
        if (remaining == 0) break;
        bool newest_is_b;
        if (first_is_zero) {
            rk4_step(P, ua, va, qa, du, dv, ub, vb, qb, k14);       // rare: the camera sits in the disk plane
            remaining--;
            newest_is_b = true;
        } else {
            for (;;) {
                rk4_step(P, ua, va, qa, du, dv, ub, vb, qb, k14);
                remaining--;
                if (((((hi32(ub) ^ side_word) & dmask) < 0) | ((unsigned)(hi32(qb) - lo_hi) >= span) | (remaining == 0)) != 0) {
                    newest_is_b = true;
                    break;
                }
                rk4_step(P, ub, vb, qb, du, dv, ua, va, qa, k14);
                remaining--;
                if (((((hi32(ua) ^ side_word) & dmask) < 0) | ((unsigned)(hi32(qa) - lo_hi) >= span) | (remaining == 0)) != 0) {
                    newest_is_b = false;
                    break;
                }
            }
        }
        // old position o, new position w
        const double uo = newest_is_b ? ua : ub, qo = newest_is_b ? qa : qb;
        const double uw = newest_is_b ? ub : ua, vw = newest_is_b ? vb : va, qw = newest_is_b ? qb : qa;
        if (first_is_zero || (((hi32(uw) ^ side_word) & dmask) < 0)) {
            // disk-plane crossing (:96); :102 r2ave = (y' r2 - y r2') / (y' - y): f1y cancels,
            // 1/qh = L^2 restores the scale
            const double r2ave = ((uw * qo - uo * qw) / (uw - uo)) / s.qh;
            if (r2ave > P.din2 && r2ave < P.dout2) disk_layer(P, r2ave, s.acc);   // :97-98
            side_word = hi32(uw) & (int)0x80000000;
            first_is_zero = false;
        }
        ua = uw; va = vw; qa = qw;
    }
    const uint32_t n = budget - remaining;
    if (status == kAlive && s.steps + n >= P.step_cap) status = kCapped;
    s.u = ua; s.v = va; s.du = du; s.dv = dv; s.q = qa;
    if (disk && n > 0) s.side = side_word < 0 ? -1 : 0;
    s.steps += n;
    s.status = status;
}

// ---- sky: StarMap.starLookup (src/StarMap.hs:93-115) over the bucketed k-d tree --------
// `top` points at the tree's top levels (shared memory on the device).
BSB_HD uint32_t star_lookup(const FrameParams &P, const float *top, const double vel[3], double rgb[3])
{
    rgb[0] = rgb[1] = rgb[2] = 0.0;
    const StarTreeDev &T = P.tree;
    if (T.n_stars <= 0) return 0;  // empty KdMap: inRadius = [] -> PixelRGB 0 0 0
    // :103 nvel = normalize vel
    double n0, n1, n2;
    {
        const double l = add_rn(add_rn(mul_rn(vel[0], vel[0]), mul_rn(vel[1], vel[1])), mul_rn(vel[2], vel[2]));
        if (fabs(l) <= 1e-12 || fabs(sub_rn(1.0, l)) <= 1e-12) {
            n0 = vel[0]; n1 = vel[1]; n2 = vel[2];
        } else {
            const double sl = sqrt_rn(l);
            n0 = div_rn(vel[0], sl); n1 = div_rn(vel[1], sl); n2 = div_rn(vel[2], sl);
        }
    }
    const double w = 0.0005;                      // :101
    const double radius = 3 * w;                  // :104
    const double r2max = mul_rn(radius, radius);  // kdt: distSqr p q <= radius*radius
    // pruning margins: the stored split planes are rounded (float / 2 stolen mantissa bits), the
    // exact inclusion test below decides membership
    const double prune_top = radius + 2e-6, prune_rec = radius + 1e-12;
    const double a_mag = 0.013862943611198907;    // :108 log 2 / dynamic (= ln2/50, correctly rounded)
    const double two_w2 = 2 * (w * w);            // :113 d2 / (2*w^2)
    const int D = T.depth, TL = T.top_levels;
    // the top levels hold FLOAT planes: walk them with single-precision compares (one issue cycle
    // each instead of two); the margin covers the rounding of the plane and of the query
    const float f0 = (float)n0, f1 = (float)n1, f2 = (float)n2;
    const float fmargin = (float)prune_top;
    const float fr2 = (float)((radius + 1e-6) * (radius + 1e-6));
    uint32_t hits = 0;
    uint32_t stack[32];                           // (depth << 26) | index within the level
    int sp = 0;
    int d = 0;
    uint32_t i = 0;
    for (;;) {
        // ---- shared-memory levels: heap index n = 2^d - 1 + i
        if (d == 0 && TL > 0) {
            // a walk from the root: the axis of level l is l mod 3 (build_star_tree), so the chain is
            // unrolled with the query component known at compile time
            uint32_t n = 0;
#pragma unroll
            for (int l = 0; l < kSmemTreeLevels; l++) {
                if (l < TL) {
                    const float qa = (l % 3 == 0) ? f0 : ((l % 3 == 1) ? f1 : f2);
                    const float diff = qa - top[n];
                    const uint32_t right = diff > 0.0f ? 1u : 0u;
                    if (fabsf(diff) <= fmargin) {
                        const uint32_t far = 2u * n + 2u - right;                 // heap index of the other child
                        stack[sp++] = ((uint32_t)(l + 1) << 26) | (far - ((2u << l) - 1u));
                    }
                    n = 2u * n + 1u + right;
                }
            }
            d = TL;
            i = n - ((1u << TL) - 1u);
        } else if (d < TL) {
            uint32_t n = (1u << d) - 1u + i;
            const uint32_t n_top = (1u << TL) - 1u;
            do {
                union { float f; uint32_t b; } cv;
                cv.f = top[n];
                const uint32_t axis = cv.b & 3u;
                const float qa = axis == 0u ? f0 : (axis == 1u ? f1 : f2);
                const float diff = qa - cv.f;
                const uint32_t right = diff > 0.0f ? 1u : 0u;
                if (fabsf(diff) <= fmargin) {
                    const uint32_t far = 2u * n + 2u - right;             // heap index of the other child
                    const int fd = d + 1;
                    stack[sp++] = ((uint32_t)fd << 26) | (far - ((1u << fd) - 1u));
                }
                n = 2u * n + 1u + right;
                d++;
            } while (n < n_top);
            i = n - ((1u << d) - 1u);
        }
        // ---- 3-level records below
        while (d < D) {
            const int g = (d - TL) / 3, l = (d - TL) - 3 * g;
            const uint32_t blk = i >> l, local = (1u << l) + (i & ((1u << l) - 1u));
            union { double d; uint64_t b; } cv;
            cv.d = T.rec[((size_t)T.rec_off[g] + blk) * 8 + local];
            const int axis = (int)(cv.b & 3ull);
            const double qa = axis == 0 ? n0 : (axis == 1 ? n1 : n2);
            const double diff = qa - cv.d;
            const uint32_t right = diff > 0.0 ? 1u : 0u;
            if (fabs(diff) <= prune_rec) stack[sp++] = ((uint32_t)(d + 1) << 26) | (2u * i + (1u - right));
            i = 2u * i + right;
            d++;
        }
        const StarRec *leaf = T.stars + (size_t)i * kLeafSlots;
#pragma unroll
        for (int j = 0; j < kLeafSlots; j++) {
            const StarRec &st = leaf[j];
            // single-precision pre-filter (conservative: radius + 1e-6), then the exact test
            const float ex_ = st.fx - f0, ey_ = st.fy - f1, ez_ = st.fz - f2;
            if (ex_ * ex_ + ey_ * ey_ + ez_ * ez_ > fr2) continue;
            const double dx = sub_rn(st.x, n0), dy = sub_rn(st.y, n1), dz = sub_rn(st.z, n2);
            const double d2 = add_rn(add_rn(mul_rn(dx, dx), mul_rn(dy, dy)), mul_rn(dz, dz)); // :107 qd
            if (d2 <= r2max) {
                const double ex = exp(a_mag * (950.0 - (double)st.mag) - d2 / two_w2);        // :113
                const double val = (ex < 1.0 ? ex : 1.0) * P.star_intensity;                  // :112
                // :114 toPixelRGB (PixelHSI hue (saturation*sat) val): channel = val (1 + S k_c)
                rgb[0] += val * fma_(P.star_saturation, st.kr, 1.0);                          // :115 foldl'
                rgb[1] += val * fma_(P.star_saturation, st.kg, 1.0);
                rgb[2] += val * fma_(P.star_saturation, st.kb, 1.0);
                hits++;
            }
        }
        if (sp == 0) break;
        const uint32_t e = stack[--sp];
        d = (int)(e >> 26);
        i = e & 0x03ffffffu;
    }
    rgb[0] = rgb[0] < 1.0 ? rgb[0] : 1.0;          // :115 fmap (min 1)
    rgb[1] = rgb[1] < 1.0 ? rgb[1] : 1.0;
    rgb[2] = rgb[2] < 1.0 ? rgb[2] : 1.0;
    return hits;
}

// Finish a terminated ray: findColor's Bottom cases (src/Raytracer.hs:93-95) + dropAlpha (:75).
// F is the frame ray_init produced for this ray (needed for the 3-D exit velocity).
BSB_HD uint32_t ray_finish(const FrameParams &P, const float *top, const RayFrame &F, const RayState &s,
                           double rgb[3])
{
    uint32_t hits = 0;
    double acc[4] = { s.acc[0], s.acc[1], s.acc[2], s.acc[3] };
    if (s.status == kSky) {
        double c[4] = { 0, 0, 0, 1.0 };
        if (P.tree.n_stars > 0) {
            const double vu = s.du * F.L, vv = s.dv * F.L;   // :94 the pre-step velocity, unscaled
            const double vel[3] = { fma_(vu, F.f1[0], vv * F.f2[0]), fma_(vu, F.f1[1], vv * F.f2[1]),
                                    fma_(vu, F.f1[2], vv * F.f2[2]) };
            hits = star_lookup(P, top, vel, c);
        }
        blend_under(acc, c);
    }
    // kBlack: blend (0,0,0,1) under acc leaves rgb unchanged; kCapped: whatever was accumulated
    rgb[0] = acc[0]; rgb[1] = acc[1]; rgb[2] = acc[2];
    return hits;
}

}  // namespace bsb
