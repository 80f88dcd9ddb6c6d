// gpd_math.cuh — per-drone device math of the DYN path, templated on the compute type.
//
// Citations are to the reference (paths relative to gym_pybullet_drones/).  In the FP64
// instantiation every expression keeps the reference's (numpy's) order of operations and the
// translation unit is compiled with -fmad=false, so results differ from the CPU reference only
// through libm (sin/cos/atan2/asin/exp).  The FP32 instantiation is the throughput mode: same
// formulas, FMA contraction allowed, rotor prologue in FP64 (see Forcing below).
#pragma once

#include "gpd_internal.h"

namespace gpd {

// FP32 throughput mode: atan2 on MUFU.RCP + the degree-17 odd polynomial of Abramowitz & Stegun 4.4.49
// (|error| <= 1.4e-8 on [0,1], i.e. below FP32 resolution of the result) instead of libdevice's ~115-instruction atan2f.
__device__ __forceinline__ float fast_atan2f(float y, float x)
{
    float ax = fabsf(x), ay = fabsf(y);
    float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    float a = mx > 0.f ? __fdividef(mn, mx) : 0.f;
    float s = a * a;
    float p = 0.0028662257f;
    p = fmaf(p, s, -0.0161657367f); p = fmaf(p, s, 0.0429096138f); p = fmaf(p, s, -0.0752896400f);
    p = fmaf(p, s, 0.1065626393f); p = fmaf(p, s, -0.1420889944f); p = fmaf(p, s, 0.1999355085f);
    p = fmaf(p, s, -0.3333314528f); p = fmaf(p, s, 1.0f);
    float r = a * p;
    if (ay > ax) r = 1.57079632679489662f - r;
    if (x < 0.f) r = 3.14159265358979324f - r;
    return copysignf(r, y);
}

__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float ex2_approx(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

template <typename R> struct M;
template <> struct M<float> {
    static constexpr bool is_double = false;
    static __device__ __forceinline__ float sqrt(float x) { return sqrtf(x); }
    static __device__ __forceinline__ void sincos(float x, float* s, float* c) { sincosf(x, s, c); }
    static __device__ __forceinline__ float atan2(float y, float x) { return fast_atan2f(y, x); }
    static __device__ __forceinline__ float asin(float x) { return fast_atan2f(x, sqrtf(fmaxf(0.f, (1.f - x) * (1.f + x)))); }
    static __device__ __forceinline__ float exp(float x) { return expf(x); }
    static __device__ __forceinline__ float abs(float x) { return fabsf(x); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdividef(a, b); }      // MUFU.RCP, 2 ulp
    static __device__ __forceinline__ float4 make4(float a, float b, float c, float d) { return make_float4(a, b, c, d); }
};
template <> struct M<double> {
    static constexpr bool is_double = true;
    static __device__ __forceinline__ double sqrt(double x) { return ::sqrt(x); }
    static __device__ __forceinline__ void sincos(double x, double* s, double* c) { ::sincos(x, s, c); }
    static __device__ __forceinline__ double atan2(double y, double x) { return ::atan2(y, x); }
    static __device__ __forceinline__ double asin(double x) { return ::asin(x); }
    static __device__ __forceinline__ double exp(double x) { return ::exp(x); }
    static __device__ __forceinline__ double abs(double x) { return fabs(x); }
    static __device__ __forceinline__ double div(double a, double b) { return a / b; }
    static __device__ __forceinline__ double4 make4(double a, double b, double c, double d) { return make_double4(a, b, c, d); }
};

// a / b for a divisor that is constant over the launch, with rb = RN(1/b) computed on the host: q0 = RN(a*rb),
// r = a - q0*b (exact in one FMA), q = RN(q0 + r*rb).  By Markstein's theorem q is the CORRECTLY ROUNDED quotient — the
// same bits as a / b — whenever rb is the correctly rounded reciprocal and nothing over- or underflows; outside a wide
// safe range of |a| (which also covers +-0, inf, NaN) the plain division runs.  3 FMA-pipe instructions instead of the
// ~12-instruction IEEE division sequence, three times per substep in the FP64 kernels (F / M, BaseAviary.py:855).
// tests/test_div_const_cpu.py checks the identity bit for bit on 3e7 random operands for every drone mass.
__device__ __forceinline__ double div_by_const(double a, double b, double rb)
{
    const double m = fabs(a);
    if (!(m > 1e-280 && m < 1e280)) return a / b;
    const double q0 = a * rb;
    const double r = fma(-q0, b, a);
    return fma(r, rb, q0);
}
__device__ __forceinline__ float div_by_const(float a, float b, float) { return a / b; }

#define GPD_PI 3.14159265358979323846

template <typename R>
struct State {
    R px, py, pz;
    R qx, qy, qz, qw;
    R vx, vy, vz;
    R wx, wy, wz;      // rpy_rates (body rates), BaseAviary.py:835
};

template <typename R> __device__ __forceinline__ R clip(R v, R lo, R hi) { return v < lo ? lo : (v > hi ? hi : v); }

// b3Matrix3x3::setRotation as reached through p.getMatrixFromQuaternion (BaseAviary.py:836); row-major m[9].
template <typename R>
__device__ __forceinline__ R quat_to_mat(R x, R y, R z, R w, R* m)
{
    R d = x * x + y * y + z * z + w * w;
    R s;
    if constexpr (M<R>::is_double) s = R(2) / d;
    else s = __fdividef(2.0f, d);      // FP32 throughput mode: MUFU.RCP (2 ulp) instead of the IEEE division sequence
    R xs = x * s, ys = y * s, zs = z * s;
    R wx = w * xs, wy = w * ys, wz = w * zs;
    R xx = x * xs, xy = x * ys, xz = x * zs;
    R yy = y * ys, yz = y * zs, zz = z * zs;
    m[0] = R(1) - (yy + zz); m[1] = xy - wz;          m[2] = xz + wy;
    m[3] = xy + wz;          m[4] = R(1) - (xx + zz); m[5] = yz - wx;
    m[6] = xz - wy;          m[7] = yz + wx;          m[8] = R(1) - (xx + yy);
    return xx + yy;                    // 1 - m[8] without the cancellation (used by the FP32 gravity term)
}

// pybullet_getEulerFromQuaternion (BaseAviary.py:518, DSLPIDControl.py:144,241)
template <typename R>
__device__ __forceinline__ void quat_to_euler(R x, R y, R z, R w, R& roll, R& pitch, R& yaw)
{
    R sqx = x * x, sqy = y * y, sqz = z * z, squ = w * w;
    R sarg = R(-2) * (x * z - w * y);
    if (sarg <= R(-0.99999)) {
        roll = R(0); pitch = R(-0.5 * GPD_PI); yaw = R(2) * M<R>::atan2(x, -y);
    } else if (sarg >= R(0.99999)) {
        roll = R(0); pitch = R(0.5 * GPD_PI); yaw = R(2) * M<R>::atan2(-x, y);
    } else {
        roll = M<R>::atan2(R(2) * (y * z + w * x), squ - sqx - sqy + sqz);
        pitch = M<R>::asin(sarg);
        yaw = M<R>::atan2(R(2) * (x * y + w * z), squ + sqx - sqy - sqz);
    }
}

// b3Quaternion::setEulerZYX + normalize (BaseAviary.py:488)
template <typename R>
__device__ __forceinline__ void euler_to_quat(R roll, R pitch, R yaw, R& x, R& y, R& z, R& w)
{
    R sr, cr, sp, cp, sy, cy;
    M<R>::sincos(roll * R(0.5), &sr, &cr);
    M<R>::sincos(pitch * R(0.5), &sp, &cp);
    M<R>::sincos(yaw * R(0.5), &sy, &cy);
    x = sr * cp * cy - cr * sp * sy;
    y = cr * sp * cy + sr * cp * sy;
    z = cr * cp * sy - sr * sp * cy;
    w = cr * cp * cy + sr * sp * sy;
    R n = M<R>::sqrt(x * x + y * y + z * z + w * w);
    x = x / n; y = y / n; z = z / n; w = w / n;
}

// BaseAviary._integrateQ (BaseAviary.py:876-889).  No renormalisation (reference quirk 13).
template <typename R>
__device__ __forceinline__ void integrate_q(State<R>& s, R dt)
{
    R p = s.wx, q = s.wy, r = s.wz;
    R cs, ap, aq, ar;
    bool series = false;
    if constexpr (!M<R>::is_double) {
        // FP32 throughput mode.  With theta = |omega|*dt/2 the update matrix is cos(theta)*I + (sin(theta)/|omega|)*Lambda:
        // both coefficients are even power series in theta, i.e. polynomials in t = theta^2 = |omega|^2*(dt/2)^2, so the
        // sqrt, the division and sincosf of the literal form disappear.  Truncation error < 3e-10 for theta < 0.5 rad
        // (|omega| < 240 rad/s at 240 Hz); above that the literal path below is taken.  For |omega| <= 1e-8 (:879) the
        // series changes q by < 1e-10 relative, below FP32 resolution.
        R h = dt * R(.5);
        R t = (p * p + q * q + r * r) * (h * h);
        if (t < R(0.25)) {
            series = true;
            cs = R(1) + t * (R(-1. / 2) + t * (R(1. / 24) + t * (R(-1. / 720) + t * R(1. / 40320))));
            R sc = h * (R(1) + t * (R(-1. / 6) + t * (R(1. / 120) + t * (R(-1. / 5040) + t * R(1. / 362880)))));
            ap = p * sc; aq = q * sc; ar = r * sc;
        }
    }
    if (!series) {
        R n = M<R>::sqrt(p * p + q * q + r * r);                 // :877
        if (n <= R(1e-8)) return;                                // :879 np.isclose(n, 0): |n| <= atol
        R theta = n * dt / R(2);                                 // :887
        R sn;
        M<R>::sincos(theta, &sn, &cs);
        R k = R(2) / n;
        ap = k * (p * R(.5)) * sn; aq = k * (q * R(.5)) * sn; ar = k * (r * R(.5)) * sn;   // :881-888
    }
    R x = s.qx, y = s.qy, z = s.qz, w = s.qw;
    s.qx = cs * x + ar * y - aq * z + ap * w;
    s.qy = -ar * x + cs * y + ap * z + aq * w;
    s.qz = aq * x - ap * y + cs * z + ar * w;
    s.qw = -ap * x - aq * y - ar * z + cs * w;
}

// Rotor terms of one ctrl step (rpm is constant over the PYB_STEPS_PER_CTRL substeps, BaseAviary.py:341-343).
//   FP64: f[k] = rpm_k^2*KF (BaseAviary.py:838); T, tx, ty, tz as BaseAviary.py:839-851.
//   FP32: the same quantities formed in FP64 and stored as T - GRAVITY, tx, ty, tz: thrust-minus-weight and
//         the rotor-force differences cancel catastrophically in FP32 near hover.
template <typename R>
struct Forcing {
    R f[4];
    R T;          // FP64: total thrust.  FP32: total thrust minus GRAVITY.
    R tx, ty, tz;
    R Ttot;       // FP32 only: total thrust
};

template <typename R>
__device__ __forceinline__ void make_forcing(const DevDrone<R>& P, const double rpm[4], Forcing<R>& F)
{
    if constexpr (M<R>::is_double) {
        double zt[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            F.f[k] = rpm[k] * rpm[k] * P.KF;                     // :838
            zt[k] = rpm[k] * rpm[k] * P.KM;                      // :842
            if (P.model == GPD_RACE) zt[k] = -zt[k];             // :843-844
        }
        F.T = F.f[0] + ((F.f[1] + F.f[2]) + F.f[3]);             // :839 (np.sum order)
        F.Ttot = F.T;
        F.tz = -zt[0] + zt[1] - zt[2] + zt[3];                   // :845
        if (P.model == GPD_CF2P) {                               // :849-851
            F.tx = (F.f[1] - F.f[3]) * P.L;
            F.ty = (-F.f[0] + F.f[2]) * P.L;
        } else {                                                 // :846-848
            F.tx = (F.f[0] + F.f[1] - F.f[2] - F.f[3]) * P.ARM;
            F.ty = (-F.f[0] + F.f[1] + F.f[2] - F.f[3]) * P.ARM;
        }
    } else {
        double f[4], zt[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double r2 = rpm[k] * rpm[k];
            f[k] = r2 * P.KF_d;
            zt[k] = (P.model == GPD_RACE) ? -(r2 * P.KM_d) : r2 * P.KM_d;
            F.f[k] = (R)f[k];
        }
        const double tt = f[0] + ((f[1] + f[2]) + f[3]);
        F.T = (R)(tt - P.GRAVITY_d);
        F.Ttot = (R)tt;
        F.tz = (R)(-zt[0] + zt[1] - zt[2] + zt[3]);
        if (P.model == GPD_CF2P) {
            F.tx = (R)((f[1] - f[3]) * P.L_d);
            F.ty = (R)((-f[0] + f[2]) * P.L_d);
        } else {
            F.tx = (R)((f[0] + f[1] - f[2] - f[3]) * P.ARM_d);
            F.ty = (R)((-f[0] + f[1] + f[2] - f[3]) * P.ARM_d);
        }
    }
}

// BaseAviary._groundEffect (BaseAviary.py:715-750): +z LINK-frame force at each rotor.  m = rotation (row-major),
// rpm2kf[k] = rpm_k^2*KF.  Returns the gate of line 742.
template <typename R>
__device__ __forceinline__ bool ground_effect(const DevDrone<R>& P, const R rpm[4], R pz, const R* m, R roll, R pitch,
                                              R out[4])
{
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        R h = pz + (m[6] * P.ROTOR[k][0] + m[7] * P.ROTOR[k][1] + m[8] * P.ROTOR[k][2]);   // :732-739
        if (h < P.GND_EFF_H_CLIP) h = P.GND_EFF_H_CLIP;                                    // :740
        R ratio = M<R>::div(P.PROP_RADIUS, R(4) * h);                                      // h >= GND_EFF_H_CLIP > 0
        out[k] = rpm[k] * rpm[k] * P.KF * P.GND_EFF_COEFF * (ratio * ratio);               // :741
    }
    return M<R>::abs(roll) < R(GPD_PI / 2) && M<R>::abs(pitch) < R(GPD_PI / 2);            // :742
}

// BaseAviary._drag (BaseAviary.py:754-781): CoM LINK-frame (body) force.  wsum = sum of the rotor speeds in rad/s (:773).
template <typename R>
__device__ __forceinline__ R drag_wsum(const R rpm_prev[4])
{
    R w0 = R(2) * R(GPD_PI) * rpm_prev[0] / R(60), w1 = R(2) * R(GPD_PI) * rpm_prev[1] / R(60);
    R w2 = R(2) * R(GPD_PI) * rpm_prev[2] / R(60), w3 = R(2) * R(GPD_PI) * rpm_prev[3] / R(60);
    return w0 + ((w1 + w2) + w3);
}

template <typename R>
__device__ __forceinline__ void drag_body_w(const DevDrone<R>& P, R wsum, const R* m, R vx, R vy, R vz, R out[3])
{
    R fx = (R(-1) * P.DRAG[0] * wsum) * vx, fy = (R(-1) * P.DRAG[1] * wsum) * vy, fz = (R(-1) * P.DRAG[2] * wsum) * vz;
    out[0] = (m[0] * fx + m[3] * fy) + m[6] * fz;                                          // :774 base_rot.T · (...)
    out[1] = (m[1] * fx + m[4] * fy) + m[7] * fz;
    out[2] = (m[2] * fx + m[5] * fy) + m[8] * fz;
}

template <typename R>
__device__ __forceinline__ void drag_body(const DevDrone<R>& P, const R rpm_prev[4], const R* m, R vx, R vy, R vz, R out[3])
{
    drag_body_w(P, drag_wsum(rpm_prev), m, vx, vy, vz, out);
}

// One pair term of BaseAviary._downwash (BaseAviary.py:799-804): body-z force on "me" from a drone at (ox,oy,oz).
// XY_SCALED (FP32 step kernel): the caller passes x and y already multiplied by GPD_DW_XY_SCALE = sqrt(0.5 * log2(e)), so that the
// squared horizontal distance arrives as the exponent's factor (one multiply less per pair; the drone scales its position once
// per substep when it writes the env's snapshot).
#define GPD_DW_XY_SCALE 0.84932180028801907f
template <typename R, bool XY_SCALED = false>
__device__ __forceinline__ R downwash_pair(const DevDrone<R>& P, R mx, R my, R mz, R ox, R oy, R oz)
{
    R delta_z = oz - mz;                                                                   // :799
    R dx = ox - mx, dy = oy - my;
    if constexpr (M<R>::is_double) {
        R delta_xy = M<R>::sqrt(dx * dx + dy * dy);                                        // :800
        if (delta_z > R(0) && delta_xy < R(10)) {                                          // :801
            R ratio = P.PROP_RADIUS / (R(4) * delta_z);
            R alpha = P.DW1 * (ratio * ratio);                                             // :802
            R beta = P.DW2 * delta_z + P.DW3;                                              // :803
            R u = delta_xy / beta;
            return -alpha * M<R>::exp(R(-.5) * (u * u));                                   // :804
        }
        return R(0);
    } else {
        // FP32 throughput mode: the same expression on squared distances (u^2 = dxy^2 / beta^2, so the sqrt of :800
        // disappears; delta_xy < 10 <=> dxy^2 < 100) with ONE raw MUFU reciprocal, r = 1/(delta_z*beta), shared by
        // 1/delta_z = r*beta and 1/beta = r*delta_z, and one raw MUFU ex2.  Branch-free: every guard (including beta = 0,
        // where u = inf and the reference's exp(-inf) gives 0) is folded into the final select.
        const float d2 = dx * dx + dy * dy;
        const float beta = fmaf(P.DW2, delta_z, P.DW3);
        const float zb = delta_z * beta;
        const float r = rcp_approx(zb);
        const float rb = r * beta;                                                         // 1 / delta_z
        const float ib = r * delta_z;                                                      // 1 / beta
        const float e = XY_SCALED ? ex2_approx(d2 * -(ib * ib))                            // d2 = 0.5 log2(e) dxy^2 already
                                  : ex2_approx(-0.72134752044448170368f * d2 * (ib * ib)); // exp(-0.5 u^2)
        const float f = P.DW1_NEG_PR2_16 * (rb * rb) * e;                                  // -DW1 * (PROP_RADIUS / 4)^2 / delta_z^2 * e
        // ONE predicate for the three guards (chained setp) and one select: the compiler's own lowering spends three selects
        float out;
        asm("{\n\t.reg .pred p;\n\t"
            "setp.gt.f32 p, %1, 0f00000000;\n\t"
            "setp.lt.and.f32 p, %2, %5, p;\n\t"                  // dxy < 10
            "setp.gt.and.f32 p, %3, 0f0DA24260, p;\n\t"          // 1e-30
            "selp.f32 %0, %4, 0f00000000, p;\n\t}"
            : "=f"(out) : "f"(delta_z), "f"(d2), "f"(fabsf(zb)), "f"(f), "f"(XY_SCALED ? 72.134752044448170368f : 100.f));
        return out;
    }
}

// One DYN substep, BaseAviary._dynamics (BaseAviary.py:831-874).
//   m        rotation matrix of the substep-start quaternion (caller computes it: the force models share it)
//   F        rotor terms of this ctrl step
//   gnd      extra +z rotor forces (DYN_GND) or nullptr
//   fb       extra body-frame force (drag + downwash) or nullptr
//   av*      ang_v = R(old)·rates(new)  (BaseAviary.py:870)
template <typename R>
__device__ __forceinline__ void dyn_substep(const DevDrone<R>& P, R dt, State<R>& s, const R* m, R one_minus_m8,
                                            const Forcing<R>& F, const R* gnd, const R* fb, bool want_av,
                                            R& avx, R& avy, R& avz)
{
    R T = F.T, tx = F.tx, ty = F.ty, tz = F.tz;
    R Fx, Fy, Fz;
    if constexpr (M<R>::is_double) {
        if (gnd) {                                               // rotor forces change -> redo :839-851
            R f0 = F.f[0] + gnd[0], f1 = F.f[1] + gnd[1], f2 = F.f[2] + gnd[2], f3 = F.f[3] + gnd[3];
            T = f0 + ((f1 + f2) + f3);
            if (P.model == GPD_CF2P) { tx = (f1 - f3) * P.L; ty = (-f0 + f2) * P.L; }
            else { tx = (f0 + f1 - f2 - f3) * P.ARM; ty = (-f0 + f1 + f2 - f3) * P.ARM; }
        }
        Fx = m[2] * T; Fy = m[5] * T; Fz = m[8] * T;             // :840
        Fz = Fz - P.GRAVITY;                                     // :841
    } else {
        if (gnd) {                                               // linear in the rotor forces -> add the deltas
            T = T + (gnd[0] + ((gnd[1] + gnd[2]) + gnd[3]));
            if (P.model == GPD_CF2P) { tx += (gnd[1] - gnd[3]) * P.L; ty += (-gnd[0] + gnd[2]) * P.L; }
            else { tx += (gnd[0] + gnd[1] - gnd[2] - gnd[3]) * P.ARM; ty += (-gnd[0] + gnd[1] + gnd[2] - gnd[3]) * P.ARM; }
        }
        // R[:,2]*T_total - [0,0,G] = R[:,2]*(T_total - G) + G*(R02, R12, R22 - 1), with R22 - 1 = -(xx+yy) taken
        // from quat_to_mat before the "1 -" so that nothing cancels near hover.
        Fx = m[2] * T + P.GRAVITY * m[2];
        Fy = m[5] * T + P.GRAVITY * m[5];
        Fz = m[8] * T - P.GRAVITY * one_minus_m8;
    }
    if (fb) {                                                    // CoM LINK-frame forces -> world = R·fb
        Fx = Fx + ((m[0] * fb[0] + m[1] * fb[1]) + m[2] * fb[2]);
        Fy = Fy + ((m[3] * fb[0] + m[4] * fb[1]) + m[5] * fb[2]);
        Fz = Fz + ((m[6] * fb[0] + m[7] * fb[1]) + m[8] * fb[2]);
    }
    // :852-853 torques - cross(rates, J·rates)
    R Jx = P.J[0] * s.wx, Jy = P.J[1] * s.wy, Jz = P.J[2] * s.wz;
    R gx = s.wy * Jz - s.wz * Jy, gy = s.wz * Jx - s.wx * Jz, gz = s.wx * Jy - s.wy * Jx;
    if constexpr (M<R>::is_double) {
        R dwx = P.JINV[0] * (tx - gx), dwy = P.JINV[1] * (ty - gy), dwz = P.JINV[2] * (tz - gz);   // :854
        R ax = div_by_const(Fx, P.M, P.INV_M), ay = div_by_const(Fy, P.M, P.INV_M), az = div_by_const(Fz, P.M, P.INV_M);   // :855, same bits as F / M
        s.vx = s.vx + dt * ax; s.vy = s.vy + dt * ay; s.vz = s.vz + dt * az;               // :857
        s.wx = s.wx + dt * dwx; s.wy = s.wy + dt * dwy; s.wz = s.wz + dt * dwz;            // :858
    } else {                                                     // same updates with dt/M and dt*J^-1 folded on the host
        s.vx = s.vx + P.DT_INV_M * Fx; s.vy = s.vy + P.DT_INV_M * Fy; s.vz = s.vz + P.DT_INV_M * Fz;
        s.wx = s.wx + P.DT_JINV[0] * (tx - gx); s.wy = s.wy + P.DT_JINV[1] * (ty - gy); s.wz = s.wz + P.DT_JINV[2] * (tz - gz);
    }
    s.px = s.px + dt * s.vx; s.py = s.py + dt * s.vy; s.pz = s.pz + dt * s.vz;             // :859
    integrate_q(s, dt);                                                                    // :860
    if (want_av) {                                               // only the last substep's value is observable
        avx = (m[0] * s.wx + m[1] * s.wy) + m[2] * s.wz;                                   // :870
        avy = (m[3] * s.wx + m[4] * s.wy) + m[5] * s.wz;
        avz = (m[6] * s.wx + m[7] * s.wy) + m[8] * s.wz;
    }
}

// FP32 throughput mode, no force models (Physics.DYN): the substeps before the last one of a ctrl step, where the
// angular velocity in the world frame (:870) is not observable.  Same update as dyn_substep, regrouped so that only
// what the state needs is formed:
//   * thrust: R[:,2]*T - [0,0,G] = (2/d)*(xz+wy, yz-wx, -(xx+yy))*T + [0,0,T-G]; (2/d)*(dt/M)*T is one factor
//   * J is diagonal (BaseAviary.py:142), so cross(w, J w)[k] = (J[k+2]-J[k+1])*w[k+1]*w[k+2] (Euler's equations) and
//     dt*J^-1*torque is constant over the ctrl step
//   * _integrateQ as the even power series of integrate_q, one degree lower and valid for theta^2 < 0.04
//     (|omega| < 96 rad/s at 240 Hz, truncation < 7e-11); above that the literal form is used.
struct LeanStep {
    float kt2;            // 2*(dt/M)*T_total
    float ktmg;           // (dt/M)*(T_total - G)
    float cx, cy, cz;     // dt*J^-1*torque
    float ex, ey, ez;     // dt*J^-1[k]*(J[k+2]-J[k+1])
    float h, hh;          // dt/2, (dt/2)^2
};

__device__ __forceinline__ void lean_substep_f32(float dt, State<float>& s, const LeanStep& c)
{
    const float x = s.qx, y = s.qy, z = s.qz, w = s.qw;
    const float d = fmaf(w, w, fmaf(z, z, fmaf(y, y, x * x)));
    const float k = c.kt2 * rcp_approx(d);
    const float r02 = fmaf(x, z, w * y), r12 = fmaf(y, z, -(w * x)), omz = fmaf(x, x, y * y);
    const float yz = s.wy * s.wz, zx = s.wz * s.wx, xy = s.wx * s.wy;      // rates of the substep start (:852)
    s.vx = fmaf(k, r02, s.vx); s.vy = fmaf(k, r12, s.vy); s.vz = fmaf(-k, omz, s.vz + c.ktmg);          // :855-857
    s.wx = fmaf(-c.ex, yz, s.wx + c.cx);                                                       // :852-854,858
    s.wy = fmaf(-c.ey, zx, s.wy + c.cy);
    s.wz = fmaf(-c.ez, xy, s.wz + c.cz);
    s.px = fmaf(dt, s.vx, s.px); s.py = fmaf(dt, s.vy, s.py); s.pz = fmaf(dt, s.vz, s.pz);              // :859
    const float t = (s.wx * s.wx + s.wy * s.wy + s.wz * s.wz) * c.hh;
    float cs = fmaf(t, fmaf(t, fmaf(t, -1.f / 720, 1.f / 24), -.5f), 1.f);                              // :860
    const float sc = c.h * fmaf(t, fmaf(t, fmaf(t, -1.f / 5040, 1.f / 120), -1.f / 6), 1.f);
    float ap = s.wx * sc, aq = s.wy * sc, ar = s.wz * sc;
    if (__builtin_expect(t >= 0.04f, 0)) {          // literal form (:877-888); the norm cannot be ~0 here
        const float n = sqrtf(s.wx * s.wx + s.wy * s.wy + s.wz * s.wz);
        float sn;
        sincosf(n * dt / 2.f, &sn, &cs);
        const float k = 2.f / n;
        ap = k * (s.wx * .5f) * sn; aq = k * (s.wy * .5f) * sn; ar = k * (s.wz * .5f) * sn;
    }
    s.qx = cs * x + ar * y - aq * z + ap * w;
    s.qy = -ar * x + cs * y + ap * z + aq * w;
    s.qz = aq * x - ap * y + cs * z + ar * w;
    s.qw = -ap * x - aq * y - ar * z + cs * w;
}

// DSLPIDControl.computeControl (control/DSLPIDControl.py:82-259).  st[9] = integral_pos_e, integral_rpy_e, last_rpy.
// The scipy matrix -> Euler('XYZ') -> matrix round trip of lines 205,242-244 is an identity to 9e-16;
// the target rotation is used directly and only its intrinsic-XYZ yaw is extracted for the returned yaw error.
template <typename R>
__device__ __forceinline__ void pid_compute(const DevPid<R>& C, R dt, R px, R py, R pz, R qx, R qy, R qz, R qw,
                                            R vx, R vy, R vz, const R tp[3], const R trpy[3], const R tv[3],
                                            const R trates[3], R st[9], R rpm[4], R pos_e[3], R& yaw_e)
{
    R m[9];
    quat_to_mat(qx, qy, qz, qw, m);                                                        // :187
    pos_e[0] = tp[0] - px; pos_e[1] = tp[1] - py; pos_e[2] = tp[2] - pz;                   // :188
    R vel_e[3] = { tv[0] - vx, tv[1] - vy, tv[2] - vz };                                   // :189
    R tt[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        st[k] = clip(st[k] + pos_e[k] * dt, R(-2), R(2));                                  // :190-191
    }
    st[2] = clip(st[2], R(-0.15), R(.15));                                                 // :192
#pragma unroll
    for (int k = 0; k < 3; ++k)                                                            // :194-196
        tt[k] = C.P_FOR[k] * pos_e[k] + C.I_FOR[k] * st[k] + C.D_FOR[k] * vel_e[k] + (k == 2 ? C.GRAVITY : R(0));
    R sthr = (tt[0] * m[2] + tt[1] * m[5]) + tt[2] * m[8];                                 // :197
    if (!(sthr > R(0))) sthr = R(0);
    R thrust = (M<R>::sqrt(sthr / C.KF4) - C.PWM2RPM_CONST) / C.PWM2RPM_SCALE;             // :198
    R ntt = M<R>::sqrt(tt[0] * tt[0] + tt[1] * tt[1] + tt[2] * tt[2]);
    R zx = tt[0] / ntt, zy = tt[1] / ntt, zz = tt[2] / ntt;                                // :199
    R sy, cy;
    M<R>::sincos(trpy[2], &sy, &cy);                                                       // :200 x_c = [cos, sin, 0]
    R c0 = zy * R(0) - zz * sy, c1 = zz * cy - zx * R(0), c2 = zx * sy - zy * cy;          // cross(z, x_c)
    R nc = M<R>::sqrt(c0 * c0 + c1 * c1 + c2 * c2);
    R yx = c0 / nc, yy = c1 / nc, yz = c2 / nc;                                            // :201
    R xx = yy * zz - yz * zy, xy = yz * zx - yx * zz, xz = yx * zy - yy * zx;              // :202 cross(y, z)
    // :203 target_rotation columns (x, y, z):  Rd[r][c]
    R Rd[9] = { xx, yx, zx, xy, yy, zy, xz, yz, zz };
    R target_yaw = M<R>::atan2(-Rd[1], Rd[0]);                                             // :205 as_euler('XYZ')[2]
    R roll, pitch, yaw;
    quat_to_euler(qx, qy, qz, qw, roll, pitch, yaw);                                       // :241
    // :245-246  E = Rd^T·R - R^T·Rd ; rot_e = [E21, E02, E10]
    R a, b, rot_e[3];
    a = (Rd[2] * m[1] + Rd[5] * m[4]) + Rd[8] * m[7]; b = (m[2] * Rd[1] + m[5] * Rd[4]) + m[8] * Rd[7]; rot_e[0] = a - b;
    a = (Rd[0] * m[2] + Rd[3] * m[5]) + Rd[6] * m[8]; b = (m[0] * Rd[2] + m[3] * Rd[5]) + m[6] * Rd[8]; rot_e[1] = a - b;
    a = (Rd[1] * m[0] + Rd[4] * m[3]) + Rd[7] * m[6]; b = (m[1] * Rd[0] + m[4] * Rd[3]) + m[7] * Rd[6]; rot_e[2] = a - b;
    R cur[3] = { roll, pitch, yaw }, tq[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        R rate_e = trates[k] - (cur[k] - st[6 + k]) / dt;                                  // :247
        st[6 + k] = cur[k];                                                                // :248
        st[3 + k] = clip(st[3 + k] - rot_e[k] * dt, R(-1500), R(1500));                    // :249-250
        if (k < 2) st[3 + k] = clip(st[3 + k], R(-1), R(1));                               // :251
        tq[k] = -(C.P_TOR[k] * rot_e[k]) + C.D_TOR[k] * rate_e + C.I_TOR[k] * st[3 + k];   // :253-255
        tq[k] = clip(tq[k], R(-3200), R(3200));                                            // :256
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {                                                          // :257-259
        R mix = (C.MIXER[r][0] * tq[0] + C.MIXER[r][1] * tq[1]) + C.MIXER[r][2] * tq[2];
        R pwm = clip(thrust + mix, C.MIN_PWM, C.MAX_PWM);
        rpm[r] = C.PWM2RPM_SCALE * pwm + C.PWM2RPM_CONST;
    }
    yaw_e = target_yaw - yaw;                                                              // :144-145
}

}  // namespace gpd
