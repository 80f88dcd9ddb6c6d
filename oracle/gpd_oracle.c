/*
 * gpd_oracle.c — CPU restatement (plain C, FP64) of the reference's Physics.DYN hot path.
 * TEST INFRASTRUCTURE; see gpd_oracle.h for the rules and the pinning statement.
 * Citations are relative to /root/reference/gym_pybullet_drones/.
 * Build: gcc -O2 -ffp-contract=off -pthread -shared -fPIC (oracle/Makefile).
 */
#include "gpd_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define ORC_PI 3.14159265358979323846

/* ------------------------------------------------------------------ */
/* Bullet closed forms (pybullet ^3.2.5, double build; not in /root/reference) */

/* b3Matrix3x3::setRotation — reached from BaseAviary.py:836, DSLPIDControl.py:187,240 */
void orc_matrix_from_quaternion(const double q[4], double m[9])
{
    double x = q[0], y = q[1], z = q[2], w = q[3];
    double d = x * x + y * y + z * z + w * w;
    double s = 2.0 / d;
    double xs = x * s, ys = y * s, zs = z * s;
    double wx = w * xs, wy = w * ys, wz = w * zs;
    double xx = x * xs, xy = x * ys, xz = x * zs;
    double yy = y * ys, yz = y * zs, zz = z * zs;
    m[0] = 1.0 - (yy + zz); m[1] = xy - wz;         m[2] = xz + wy;
    m[3] = xy + wz;         m[4] = 1.0 - (xx + zz); m[5] = yz - wx;
    m[6] = xz - wy;         m[7] = yz + wx;         m[8] = 1.0 - (xx + yy);
}

/* pybullet_getEulerFromQuaternion — reached from BaseAviary.py:518, DSLPIDControl.py:144,241 */
void orc_euler_from_quaternion(const double q[4], double rpy[3])
{
    double x = q[0], y = q[1], z = q[2], w = q[3];
    double sqx = x * x, sqy = y * y, sqz = z * z, squ = w * w;
    double sarg = -2.0 * (x * z - w * y);
    if (sarg <= -0.99999) {
        rpy[0] = 0.0; rpy[1] = -0.5 * ORC_PI; rpy[2] = 2.0 * atan2(x, -y);
    } else if (sarg >= 0.99999) {
        rpy[0] = 0.0; rpy[1] = 0.5 * ORC_PI;  rpy[2] = 2.0 * atan2(-x, y);
    } else {
        rpy[0] = atan2(2.0 * (y * z + w * x), squ - sqx - sqy + sqz);
        rpy[1] = asin(sarg);
        rpy[2] = atan2(2.0 * (x * y + w * z), squ + sqx - sqy - sqz);
    }
}

/* b3Quaternion::setEulerZYX + normalize — reached from BaseAviary.py:488 */
void orc_quaternion_from_euler(const double rpy[3], double q[4])
{
    double hr = rpy[0] * 0.5, hp = rpy[1] * 0.5, hy = rpy[2] * 0.5;
    double cy = cos(hy), sy = sin(hy), cp = cos(hp), sp = sin(hp), cr = cos(hr), sr = sin(hr);
    double x = sr * cp * cy - cr * sp * sy;
    double y = cr * sp * cy + sr * cp * sy;
    double z = cr * cp * sy - sr * sp * cy;
    double w = cr * cp * cy + sr * sp * sy;
    double n = sqrt(x * x + y * y + z * z + w * w);
    q[0] = x / n; q[1] = y / n; q[2] = z / n; q[3] = w / n;
}

/* ------------------------------------------------------------------ */
static double norm3(const double v[3]) { return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }

/* numpy.cross for two 3-vectors: c0 = a1*b2 - a2*b1, ... */
static void cross3(const double a[3], const double b[3], double c[3])
{
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

static double clipd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* BaseAviary.py:876-889 */
void orc_integrate_q(const double quat[4], const double omega[3], double dt, double out[4])
{
    double n = norm3(omega);                             /* :877 */
    double p = omega[0], q = omega[1], r = omega[2];     /* :878 */
    if (fabs(n - 0.0) <= 1e-8) {                         /* :879 np.isclose(n, 0): atol 1e-8 + rtol*|0| */
        memcpy(out, quat, 4 * sizeof(double));
        return;
    }
    /* :881-886 lambda_ * .5 */
    double lam[4][4] = {
        { 0.0 * .5,  r * .5, -q * .5, p * .5 },
        { -r * .5, 0.0 * .5,  p * .5, q * .5 },
        {  q * .5, -p * .5, 0.0 * .5, r * .5 },
        { -p * .5, -q * .5, -r * .5, 0.0 * .5 } };
    double theta = n * dt / 2;                           /* :887 */
    double c = cos(theta), s = sin(theta), k = 2 / n;
    for (int i = 0; i < 4; ++i) {                        /* :888 */
        double acc = 0.0;
        for (int j = 0; j < 4; ++j) {
            double mij = (i == j ? 1.0 : 0.0) * c + k * lam[i][j] * s;
            acc += mij * quat[j];
        }
        out[i] = acc;
    }
}

/* BaseAviary.py:831-874 (+ build-defined injection of the force models, see header) */
void orc_dynamics(const orc_drone* d, double dt, const double rpm[4],
                  double pos[3], double quat[4], double vel[3], double rates[3], double ang_v_out[3],
                  const double gnd[4], const double f_ext_world[3], const double f_ext_body[3])
{
    double R[9];
    orc_matrix_from_quaternion(quat, R);                 /* :836 */
    double f[4], zt[4];
    for (int k = 0; k < 4; ++k) {
        f[k] = rpm[k] * rpm[k] * d->KF;                  /* :838 */
        if (gnd) f[k] = f[k] + gnd[k];                   /* DYN_GND: extra +z force at rotor k */
        zt[k] = rpm[k] * rpm[k] * d->KM;                 /* :842 */
        if (d->model == ORC_RACE) zt[k] = -zt[k];        /* :843-844 */
    }
    double T = f[0] + ((f[1] + f[2]) + f[3]);            /* :839 np.sum: first item + pairwise(rest) */
    double Fw[3] = { R[2] * T, R[5] * T, R[8] * T };     /* :840 */
    Fw[2] = Fw[2] - d->GRAVITY;                          /* :841 */
    if (f_ext_world) for (int k = 0; k < 3; ++k) Fw[k] = Fw[k] + f_ext_world[k];
    if (f_ext_body) for (int k = 0; k < 3; ++k)
        Fw[k] = Fw[k] + ((R[3 * k] * f_ext_body[0] + R[3 * k + 1] * f_ext_body[1]) + R[3 * k + 2] * f_ext_body[2]);
    double z_torque = -zt[0] + zt[1] - zt[2] + zt[3];    /* :845 */
    double x_torque, y_torque;
    if (d->model == ORC_CF2X || d->model == ORC_RACE) {  /* :846-848 */
        double arm = d->L / sqrt(2.0);
        x_torque = (f[0] + f[1] - f[2] - f[3]) * arm;
        y_torque = (-f[0] + f[1] + f[2] - f[3]) * arm;
    } else {                                             /* :849-851 */
        x_torque = (f[1] - f[3]) * d->L;
        y_torque = (-f[0] + f[2]) * d->L;
    }
    double Jw[3] = { d->J[0] * rates[0], d->J[1] * rates[1], d->J[2] * rates[2] };
    double gyro[3];
    cross3(rates, Jw, gyro);
    double tau[3] = { x_torque - gyro[0], y_torque - gyro[1], z_torque - gyro[2] };   /* :852-853 */
    double wdot[3] = { d->J_INV[0] * tau[0], d->J_INV[1] * tau[1], d->J_INV[2] * tau[2] }; /* :854 */
    double acc[3] = { Fw[0] / d->M, Fw[1] / d->M, Fw[2] / d->M };                     /* :855 */
    for (int k = 0; k < 3; ++k) vel[k] = vel[k] + dt * acc[k];                        /* :857 */
    for (int k = 0; k < 3; ++k) rates[k] = rates[k] + dt * wdot[k];                   /* :858 */
    for (int k = 0; k < 3; ++k) pos[k] = pos[k] + dt * vel[k];                        /* :859 */
    double qn[4];
    orc_integrate_q(quat, rates, dt, qn);                                             /* :860 */
    memcpy(quat, qn, sizeof qn);
    for (int k = 0; k < 3; ++k)                                                       /* :870 */
        ang_v_out[k] = (R[3 * k] * rates[0] + R[3 * k + 1] * rates[1]) + R[3 * k + 2] * rates[2];
}

/* BaseAviary.py:715-750 */
int orc_ground_effect(const orc_drone* d, const double rpm[4], const double pos[3], const double quat[4],
                      const double rpy[3], double out[4])
{
    double R[9];
    orc_matrix_from_quaternion(quat, R);
    for (int k = 0; k < 4; ++k) {
        const double* o = d->ROTOR_XYZ[k];
        /* :732-739 world z of rotor link k's CoM (getLinkStates()[k][0][2]) */
        double h = pos[2] + (R[6] * o[0] + R[7] * o[1] + R[8] * o[2]);
        if (h < d->GND_EFF_H_CLIP) h = d->GND_EFF_H_CLIP;                /* :740 */
        double ratio = d->PROP_RADIUS / (4 * h);
        out[k] = rpm[k] * rpm[k] * d->KF * d->GND_EFF_COEFF * (ratio * ratio);  /* :741 */
    }
    return (fabs(rpy[0]) < ORC_PI / 2 && fabs(rpy[1]) < ORC_PI / 2) ? 1 : 0;     /* :742 */
}

/* BaseAviary.py:754-781 */
void orc_drag(const orc_drone* d, const double rpm_prev[4], const double quat[4], const double vel[3], double out[3])
{
    double R[9];
    orc_matrix_from_quaternion(quat, R);                                  /* :771 */
    double w[4];
    for (int k = 0; k < 4; ++k) w[k] = 2 * ORC_PI * rpm_prev[k] / 60;     /* :773 */
    double wsum = w[0] + ((w[1] + w[2]) + w[3]);
    double fv[3];
    for (int k = 0; k < 3; ++k) fv[k] = (-1 * d->DRAG_COEFF[k] * wsum) * vel[k];   /* :773-774 */
    for (int k = 0; k < 3; ++k)                                           /* :774 base_rot.T · (...) */
        out[k] = (R[k] * fv[0] + R[3 + k] * fv[1]) + R[6 + k] * fv[2];
}

/* BaseAviary.py:785-811 */
double orc_downwash(const orc_drone* d, int n, const double* pos_all, int i)
{
    double total = 0.0;
    const double* pi = pos_all + 3 * i;
    for (int j = 0; j < n; ++j) {                                         /* :798 */
        const double* pj = pos_all + 3 * j;
        double delta_z = pj[2] - pi[2];                                   /* :799 */
        double dx = pj[0] - pi[0], dy = pj[1] - pi[1];
        double delta_xy = sqrt(dx * dx + dy * dy);                        /* :800 */
        if (delta_z > 0 && delta_xy < 10) {                               /* :801 */
            double ratio = d->PROP_RADIUS / (4 * delta_z);
            double alpha = d->DW_COEFF_1 * (ratio * ratio);               /* :802 */
            double beta = d->DW_COEFF_2 * delta_z + d->DW_COEFF_3;        /* :803 */
            double u = delta_xy / beta;
            total += -alpha * exp(-.5 * (u * u));                         /* :804 */
        }
    }
    return total;
}

/* BaseAviary.py:1105-1147 */
void orc_calculate_next_step(const double cur[3], const double dest[3], double step_size, double out[3])
{
    double dir[3] = { dest[0] - cur[0], dest[1] - cur[1], dest[2] - cur[2] };
    double dist = norm3(dir);
    if (dist <= step_size) { memcpy(out, dest, 3 * sizeof(double)); return; }
    for (int k = 0; k < 3; ++k) out[k] = cur[k] + dir[k] / dist * step_size;
}

/* scipy Rotation.from_matrix(M).as_euler('XYZ') for a proper rotation M (row-major):
 * M = Rx(a)·Ry(b)·Rz(c)  =>  b = asin(M02), a = atan2(-M12, M22), c = atan2(-M01, M00).
 * (DSLPIDControl.py:205; scipy ^1.10 is third-party.) */
static void euler_XYZ_from_matrix(const double M[9], double e[3])
{
    double sb = M[2];
    if (sb > 1.0) sb = 1.0;
    if (sb < -1.0) sb = -1.0;
    e[1] = asin(sb);
    e[0] = atan2(-M[5], M[8]);
    e[2] = atan2(-M[1], M[0]);
}

/* scipy Rotation.from_euler('XYZ', e).as_matrix()  (DSLPIDControl.py:242-244; the w,x,y,z
 * unpack/re-pack there is a no-op) */
static void matrix_from_euler_XYZ(const double e[3], double M[9])
{
    double ca = cos(e[0]), sa = sin(e[0]), cb = cos(e[1]), sb = sin(e[1]), cc = cos(e[2]), sc = sin(e[2]);
    M[0] = cb * cc;                 M[1] = -cb * sc;                M[2] = sb;
    M[3] = ca * sc + sa * sb * cc;  M[4] = ca * cc - sa * sb * sc;  M[5] = -sa * cb;
    M[6] = sa * sc - ca * sb * cc;  M[7] = sa * cc + ca * sb * sc;  M[8] = ca * cb;
}

/* control/DSLPIDControl.py:82-145 = :149-208 + :212-259 */
void orc_pid_compute(const orc_pid* c, double dt, const double cur_pos[3], const double cur_quat[4],
                     const double cur_vel[3], const double target_pos[3], const double target_rpy[3],
                     const double target_vel[3], const double target_rpy_rates[3],
                     double pid_state[9], double rpm_out[4], double pos_e_out[3], double* yaw_e_out)
{
    double* integral_pos_e = pid_state;
    double* integral_rpy_e = pid_state + 3;
    double* last_rpy = pid_state + 6;
    double R[9];
    orc_matrix_from_quaternion(cur_quat, R);                                   /* :187 */
    double pos_e[3], vel_e[3], tt[3];
    for (int k = 0; k < 3; ++k) { pos_e[k] = target_pos[k] - cur_pos[k]; vel_e[k] = target_vel[k] - cur_vel[k]; } /* :188-189 */
    for (int k = 0; k < 3; ++k) {
        integral_pos_e[k] = integral_pos_e[k] + pos_e[k] * dt;                 /* :190 */
        integral_pos_e[k] = clipd(integral_pos_e[k], -2., 2.);                 /* :191 */
    }
    integral_pos_e[2] = clipd(integral_pos_e[2], -0.15, .15);                  /* :192 */
    for (int k = 0; k < 3; ++k)                                                /* :194-196 */
        tt[k] = c->P_FOR[k] * pos_e[k] + c->I_FOR[k] * integral_pos_e[k] + c->D_FOR[k] * vel_e[k]
                + (k == 2 ? c->GRAVITY : 0.0);
    double st = (tt[0] * R[2] + tt[1] * R[5]) + tt[2] * R[8];                  /* :197 */
    if (!(st > 0.)) st = 0.;
    double thrust = (sqrt(st / (4 * c->KF)) - c->PWM2RPM_CONST) / c->PWM2RPM_SCALE;   /* :198 */
    double ntt = norm3(tt);
    double z_ax[3] = { tt[0] / ntt, tt[1] / ntt, tt[2] / ntt };                /* :199 */
    double x_c[3] = { cos(target_rpy[2]), sin(target_rpy[2]), 0.0 };           /* :200 */
    double zx[3], y_ax[3], x_ax[3];
    cross3(z_ax, x_c, zx);
    double nzx = norm3(zx);
    for (int k = 0; k < 3; ++k) y_ax[k] = zx[k] / nzx;                         /* :201 */
    cross3(y_ax, z_ax, x_ax);                                                  /* :202 */
    double Rt[9] = { x_ax[0], y_ax[0], z_ax[0],                                /* :203 columns x,y,z */
                     x_ax[1], y_ax[1], z_ax[1],
                     x_ax[2], y_ax[2], z_ax[2] };
    double target_euler[3];
    euler_XYZ_from_matrix(Rt, target_euler);                                   /* :205 */
    /* ---- attitude loop :240-259 ---- */
    double cur_rpy[3];
    orc_euler_from_quaternion(cur_quat, cur_rpy);                              /* :241 */
    double Rd[9];
    matrix_from_euler_XYZ(target_euler, Rd);                                   /* :242-244 */
    /* :245  E = Rd^T·R − R^T·Rd ; :246 rot_e = [E21, E02, E10] */
    double E21 = 0, E02 = 0, E10 = 0;
    {
        double a = 0, b = 0;
        for (int k = 0; k < 3; ++k) { a += Rd[3 * k + 2] * R[3 * k + 1]; b += R[3 * k + 2] * Rd[3 * k + 1]; }
        E21 = a - b;
        a = 0; b = 0;
        for (int k = 0; k < 3; ++k) { a += Rd[3 * k + 0] * R[3 * k + 2]; b += R[3 * k + 0] * Rd[3 * k + 2]; }
        E02 = a - b;
        a = 0; b = 0;
        for (int k = 0; k < 3; ++k) { a += Rd[3 * k + 1] * R[3 * k + 0]; b += R[3 * k + 1] * Rd[3 * k + 0]; }
        E10 = a - b;
    }
    double rot_e[3] = { E21, E02, E10 };
    double rate_e[3], tq[3];
    for (int k = 0; k < 3; ++k) {
        rate_e[k] = target_rpy_rates[k] - (cur_rpy[k] - last_rpy[k]) / dt;     /* :247 */
        last_rpy[k] = cur_rpy[k];                                              /* :248 */
        integral_rpy_e[k] = integral_rpy_e[k] - rot_e[k] * dt;                 /* :249 */
        integral_rpy_e[k] = clipd(integral_rpy_e[k], -1500., 1500.);           /* :250 */
    }
    integral_rpy_e[0] = clipd(integral_rpy_e[0], -1., 1.);                     /* :251 */
    integral_rpy_e[1] = clipd(integral_rpy_e[1], -1., 1.);
    for (int k = 0; k < 3; ++k) {                                              /* :253-256 */
        tq[k] = -(c->P_TOR[k] * rot_e[k]) + c->D_TOR[k] * rate_e[k] + c->I_TOR[k] * integral_rpy_e[k];
        tq[k] = clipd(tq[k], -3200, 3200);
    }
    for (int m = 0; m < 4; ++m) {                                              /* :257-259 */
        double mix = (c->MIXER[m][0] * tq[0] + c->MIXER[m][1] * tq[1]) + c->MIXER[m][2] * tq[2];
        double pwm = clipd(thrust + mix, c->MIN_PWM, c->MAX_PWM);
        rpm_out[m] = c->PWM2RPM_SCALE * pwm + c->PWM2RPM_CONST;
    }
    if (pos_e_out) memcpy(pos_e_out, pos_e, sizeof pos_e);
    if (yaw_e_out) *yaw_e_out = target_euler[2] - cur_rpy[2];                  /* :144-145 */
}

int orc_action_width(int action_type)                                           /* BaseRLAviary.py:140-145 */
{
    switch (action_type) {
    case ORC_ACT_RPM: case ORC_ACT_VEL: case ORC_ACT_CTRL_RPM: case ORC_ACT_CTRL_VEL: return 4;
    case ORC_ACT_PID: return 3;
    case ORC_ACT_ONE_D_RPM: case ORC_ACT_ONE_D_PID: return 1;
    default: return -1;
    }
}

int orc_max_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

/* state20 field offsets, BaseAviary.py:559-561 */
enum { S_POS = 0, S_QUAT = 3, S_RPY = 7, S_VEL = 10, S_ANGV = 13, S_RPM = 16 };

/* BaseRLAviary.py:307-319 / CtrlAviary.py:117 */
static void write_obs(const orc_env_cfg* cfg, const double* st /*[N][20]*/, const float* ring /*[N][B][A]*/, void* obs_env)
{
    int N = cfg->num_drones;
    if (cfg->env_kind == ORC_ENV_CTRL) {
        memcpy(obs_env, st, (size_t)N * 20 * sizeof(double));
        return;
    }
    int A = orc_action_width(cfg->action_type), B = cfg->action_buffer_size, W = 12 + A * B;
    float* o = (float*)obs_env;
    for (int i = 0; i < N; ++i) {
        const double* s = st + 20 * i;
        float* r = o + (size_t)W * i;
        for (int k = 0; k < 3; ++k) {
            r[k] = (float)s[S_POS + k]; r[3 + k] = (float)s[S_RPY + k];
            r[6 + k] = (float)s[S_VEL + k]; r[9 + k] = (float)s[S_ANGV + k];
        }
        memcpy(r + 12, ring + (size_t)i * B * A, (size_t)B * A * sizeof(float));
    }
}

static void step_one_env(const orc_env_cfg* cfg, double* st, double* rr, double* ps, float* ring, int32_t* counter,
                         const void* act_env, const double* target_pos, void* obs_env,
                         double* reward, uint8_t* terminated, uint8_t* truncated, double* scratch)
{
    const orc_drone* d = &cfg->drone;
    int N = cfg->num_drones, A = orc_action_width(cfg->action_type), B = cfg->action_buffer_size;
    double dt = 1. / cfg->pyb_freq;                                      /* BaseAviary.py:83 */
    double ctrl_dt = 1. / cfg->ctrl_freq;                                /* BaseAviary.py:82 */
    double* rpm = scratch;                 /* [N][4] */
    double* snap_pos = scratch + 4 * N;    /* [N][3] */

    /* ---- _preprocessAction ---- */
    for (int i = 0; i < N; ++i) {
        double* s = st + 20 * i;
        double* r = rpm + 4 * i;
        if (cfg->action_type == ORC_ACT_CTRL_RPM) {                      /* CtrlAviary.py:140 */
            const double* a = (const double*)act_env + 4 * i;
            for (int k = 0; k < 4; ++k) r[k] = clipd(a[k], 0, d->MAX_RPM);
            continue;
        }
        if (cfg->action_type == ORC_ACT_CTRL_VEL) {                      /* VelocityAviary.py:148-169 (float64 action) */
            const double* a = (const double*)act_env + 4 * i;
            double n = norm3(a), u[3] = { 0, 0, 0 };
            if (n != 0) { u[0] = a[0] / n; u[1] = a[1] / n; u[2] = a[2] / n; }
            double sp = cfg->speed_limit * fabs(a[3]);
            double tv[3] = { sp * u[0], sp * u[1], sp * u[2] };
            double trpy[3] = { 0, 0, s[S_RPY + 2] }, zero[3] = { 0, 0, 0 };
            orc_pid_compute(&cfg->pid, ctrl_dt, s + S_POS, s + S_QUAT, s + S_VEL, s + S_POS, trpy, tv, zero,
                            ps + 9 * i, r, NULL, NULL);
            continue;
        }
        const float* a = (const float*)act_env + A * i;
        float* rg = ring + (size_t)i * B * A;                            /* BaseRLAviary.py:187 deque(maxlen=B).append */
        memmove(rg, rg + A, (size_t)(B - 1) * A * sizeof(float));
        memcpy(rg + (size_t)(B - 1) * A, a, A * sizeof(float));
        switch (cfg->action_type) {
        case ORC_ACT_RPM:                                                /* BaseRLAviary.py:191-192 */
            for (int k = 0; k < 4; ++k) { float t = 1.0f + 0.05f * a[k]; r[k] = d->HOVER_RPM * (double)t; }
            break;
        case ORC_ACT_ONE_D_RPM: {                                        /* BaseRLAviary.py:224-225 */
            float t = 1.0f + 0.05f * a[0];
            for (int k = 0; k < 4; ++k) r[k] = d->HOVER_RPM * (double)t;
            break; }
        case ORC_ACT_PID: {                                              /* BaseRLAviary.py:193-207 */
            double dest[3] = { a[0], a[1], a[2] }, nxt[3], zero[3] = { 0, 0, 0 };
            orc_calculate_next_step(s + S_POS, dest, 1, nxt);
            orc_pid_compute(&cfg->pid, ctrl_dt, s + S_POS, s + S_QUAT, s + S_VEL, nxt, zero, zero, zero,
                            ps + 9 * i, r, NULL, NULL);
            break; }
        case ORC_ACT_VEL: {                                              /* BaseRLAviary.py:208-223 (float32 sub-expressions) */
            /* np.linalg.norm(float32[3]) = sqrt(x.dot(x)): OpenBLAS sdot forms float32 products and accumulates its
             * scalar tail (n < 32) in double, then rounds the sum to float32; the sqrt is a float32 sqrt.  Checked
             * bit-for-bit against numpy on 200,000 random vectors (oracle/gen_golden.py environment). */
            float n = sqrtf((float)((double)(a[0] * a[0]) + (double)(a[1] * a[1]) + (double)(a[2] * a[2])));
            float u[3] = { 0.f, 0.f, 0.f };
            if (n != 0.f) { u[0] = a[0] / n; u[1] = a[1] / n; u[2] = a[2] / n; }
            float sp = (float)cfg->speed_limit * fabsf(a[3]);
            double tv[3] = { (double)(sp * u[0]), (double)(sp * u[1]), (double)(sp * u[2]) };
            double trpy[3] = { 0, 0, s[S_RPY + 2] }, zero[3] = { 0, 0, 0 };
            orc_pid_compute(&cfg->pid, ctrl_dt, s + S_POS, s + S_QUAT, s + S_VEL, s + S_POS, trpy, tv, zero,
                            ps + 9 * i, r, NULL, NULL);
            break; }
        case ORC_ACT_ONE_D_PID: {                                        /* BaseRLAviary.py:226-235 */
            double tp[3] = { s[0] + 0.1 * 0.0, s[1] + 0.1 * 0.0, s[2] + 0.1 * (double)a[0] }, zero[3] = { 0, 0, 0 };
            orc_pid_compute(&cfg->pid, ctrl_dt, s + S_POS, s + S_QUAT, s + S_VEL, tp, zero, zero, zero,
                            ps + 9 * i, r, NULL, NULL);
            break; }
        default: break;
        }
    }

    /* ---- substep loop, BaseAviary.py:343-372 ---- */
    for (int sub = 0; sub < cfg->substeps; ++sub) {
        /* :346-347 snapshot: rpy of every drone refreshed from its quaternion (:518) */
        for (int i = 0; i < N; ++i) {
            double* s = st + 20 * i;
            orc_euler_from_quaternion(s + S_QUAT, s + S_RPY);
            memcpy(snap_pos + 3 * i, s + S_POS, 3 * sizeof(double));
        }
        for (int i = 0; i < N; ++i) {                                    /* :349-353 */
            double* s = st + 20 * i;
            double gnd[4], drag_body[3], dw_body[3] = { 0, 0, 0 }, fb[3] = { 0, 0, 0 };
            const double* pg = NULL; const double* pb = NULL;
            if (cfg->physics_flags & ORC_PHY_GND) {
                /* the dynamics use the snapshot position: s[S_POS] is still the snapshot for drone i */
                if (orc_ground_effect(d, rpm + 4 * i, s + S_POS, s + S_QUAT, s + S_RPY, gnd)) pg = gnd;
            }
            if (cfg->physics_flags & ORC_PHY_DRAG) {                     /* :359,366: rpm = last_clipped_action */
                orc_drag(d, s + S_RPM, s + S_QUAT, s + S_VEL, drag_body);
                for (int k = 0; k < 3; ++k) fb[k] += drag_body[k];
                pb = fb;
            }
            if (cfg->physics_flags & ORC_PHY_DW) {
                dw_body[2] = orc_downwash(d, N, snap_pos, i);
                for (int k = 0; k < 3; ++k) fb[k] += dw_body[k];
                pb = fb;
            }
            orc_dynamics(d, dt, rpm + 4 * i, s + S_POS, s + S_QUAT, s + S_VEL, rr + 3 * i, s + S_ANGV, pg, NULL, pb);
        }
        for (int i = 0; i < N; ++i) memcpy(st + 20 * i + S_RPM, rpm + 4 * i, 4 * sizeof(double));   /* :372 */
    }
    /* :374 */
    for (int i = 0; i < N; ++i) orc_euler_from_quaternion(st + 20 * i + S_QUAT, st + 20 * i + S_RPY);
    if (cfg->substeps == 0)
        for (int i = 0; i < N; ++i) memcpy(st + 20 * i + S_RPM, rpm + 4 * i, 4 * sizeof(double));

    /* :376-380 */
    if (obs_env) write_obs(cfg, st, ring, obs_env);
    double rew = -1; int term = 0, trunc = 0;
    if (cfg->env_kind == ORC_ENV_HOVER) {                                /* HoverAviary.py:68-117 */
        const double* s = st;
        double e[3] = { target_pos[0] - s[0], target_pos[1] - s[1], target_pos[2] - s[2] };
        double n = norm3(e);
        double v = 2 - pow(n, 4);
        rew = v > 0 ? v : 0;
        term = n < .0001;
        trunc = (fabs(s[0]) > 1.5 || fabs(s[1]) > 1.5 || s[2] > 2.0 || fabs(s[7]) > .4 || fabs(s[8]) > .4);
        if ((double)*counter / (double)cfg->pyb_freq > cfg->episode_len_sec) trunc = 1;
    } else if (cfg->env_kind == ORC_ENV_MULTIHOVER) {                    /* MultiHoverAviary.py:84-130 */
        double ret = 0, dist = 0;
        for (int i = 0; i < N; ++i) {
            const double* s = st + 20 * i;
            const double* t = target_pos + 3 * i;
            double e[3] = { t[0] - s[0], t[1] - s[1], t[2] - s[2] };
            double n = norm3(e);
            double v = 2 - pow(n, 4);
            ret += v > 0 ? v : 0;
            dist += n;
            if (fabs(s[0]) > 2.0 || fabs(s[1]) > 2.0 || s[2] > 2.0 || fabs(s[7]) > .4 || fabs(s[8]) > .4) trunc = 1;
        }
        rew = ret;
        term = dist < .0001;
        if ((double)*counter / (double)cfg->pyb_freq > cfg->episode_len_sec) trunc = 1;
    }
    if (reward) *reward = rew;
    if (terminated) *terminated = (uint8_t)term;
    if (truncated) *truncated = (uint8_t)trunc;
    *counter = *counter + cfg->substeps;                                 /* :382 */
}

typedef struct step_job {
    const orc_env_cfg* cfg;
    int64_t e0, e1;
    double* state20; double* rpy_rates; double* pid_state; float* ring; int32_t* step_counter;
    const void* actions; const double* target_pos;
    void* obs; double* reward; uint8_t* terminated; uint8_t* truncated;
} step_job;

static void* step_range(void* arg)
{
    const step_job* j = (const step_job*)arg;
    const orc_env_cfg* cfg = j->cfg;
    int N = cfg->num_drones, A = orc_action_width(cfg->action_type), B = cfg->action_buffer_size;
    int is_ctrl = cfg->env_kind == ORC_ENV_CTRL;
    size_t act_stride = (size_t)N * A * ((cfg->action_type == ORC_ACT_CTRL_RPM || cfg->action_type == ORC_ACT_CTRL_VEL) ? sizeof(double) : sizeof(float));
    size_t obs_stride = is_ctrl ? (size_t)N * 20 * sizeof(double) : (size_t)N * (12 + A * B) * sizeof(float);
    double* scratch = (double*)malloc(sizeof(double) * 7 * (size_t)N);
    for (int64_t e = j->e0; e < j->e1; ++e) {
        step_one_env(cfg, j->state20 + (size_t)e * N * 20, j->rpy_rates + (size_t)e * N * 3,
                     j->pid_state ? j->pid_state + (size_t)e * N * 9 : NULL,
                     j->ring ? j->ring + (size_t)e * N * B * A : NULL, j->step_counter + e,
                     (const char*)j->actions + e * act_stride, j->target_pos,
                     j->obs ? (char*)j->obs + e * obs_stride : NULL,
                     j->reward ? j->reward + e : NULL, j->terminated ? j->terminated + e : NULL,
                     j->truncated ? j->truncated + e : NULL, scratch);
    }
    free(scratch);
    return NULL;
}

void orc_step(const orc_env_cfg* cfg, int64_t num_envs,
              double* state20, double* rpy_rates, double* pid_state, float* ring, int32_t* step_counter,
              const void* actions, const double* target_pos,
              void* obs, double* reward, uint8_t* terminated, uint8_t* truncated, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if ((int64_t)nthreads > num_envs) nthreads = num_envs > 0 ? (int)num_envs : 1;
    step_job* jobs = (step_job*)malloc(sizeof(step_job) * (size_t)nthreads);
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
    for (int t = 0; t < nthreads; ++t) {
        step_job j = { cfg, num_envs * t / nthreads, num_envs * (t + 1) / nthreads,
                       state20, rpy_rates, pid_state, ring, step_counter, actions, target_pos,
                       obs, reward, terminated, truncated };
        jobs[t] = j;
    }
    for (int t = 1; t < nthreads; ++t) pthread_create(&th[t], NULL, step_range, &jobs[t]);
    step_range(&jobs[0]);
    for (int t = 1; t < nthreads; ++t) pthread_join(th[t], NULL);
    free(th);
    free(jobs);
}

void orc_reset(const orc_env_cfg* cfg, int64_t num_envs, const uint8_t* mask,
               const double* init_xyz, const double* init_rpy,
               double* state20, double* rpy_rates, const float* ring, int32_t* step_counter, void* obs)
{
    int N = cfg->num_drones, A = orc_action_width(cfg->action_type), B = cfg->action_buffer_size;
    int is_ctrl = cfg->env_kind == ORC_ENV_CTRL;
    size_t obs_stride = is_ctrl ? (size_t)N * 20 * sizeof(double) : (size_t)N * (12 + A * B) * sizeof(float);
    for (int64_t e = 0; e < num_envs; ++e) {
        if (mask && !mask[e]) continue;
        double* st = state20 + (size_t)e * N * 20;
        for (int i = 0; i < N; ++i) {
            double* s = st + 20 * i;
            memset(s, 0, 20 * sizeof(double));                           /* BaseAviary.py:468-475 */
            memcpy(s + S_POS, init_xyz + ((size_t)e * N + i) * 3, 3 * sizeof(double));     /* :487 */
            orc_quaternion_from_euler(init_rpy + ((size_t)e * N + i) * 3, s + S_QUAT);     /* :488 */
            orc_euler_from_quaternion(s + S_QUAT, s + S_RPY);            /* :249 -> :518 */
            memset(rpy_rates + ((size_t)e * N + i) * 3, 0, 3 * sizeof(double));            /* :477 */
        }
        step_counter[e] = 0;                                             /* :460 */
        if (obs) write_obs(cfg, st, ring ? ring + (size_t)e * N * B * A : NULL, (char*)obs + e * obs_stride);
    }
}
