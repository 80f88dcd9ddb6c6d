/*
 * gpd_oracle.h — CPU restatement (plain C, FP64) of the reference's Physics.DYN hot path.
 *
 * TEST INFRASTRUCTURE. This is the parity oracle for the CUDA library in
 * gym-pybullet-drones-routing_b200/csrc. Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it. It is never on the
 * product path; the product fails loudly when its CUDA library is missing.
 *
 * Pinning: the reference's own tests hold no golden vectors for this path
 * (reference tests/test_examples.py:1-15 never run Physics.DYN).  The oracle is
 * pinned instead against outputs of the reference's own unmodified Python run in
 * the build container under the stand-ins in oracle/refshim (generator:
 * oracle/gen_golden.py, fixtures: tests/golden/) and, live, by oracle/fuzz_vs_reference.py
 * (random configurations replayed in the reference and here).  Bullet's three closed-form
 * converters are restated from Bullet's published formulas and could not be
 * checked against a real pybullet wheel ("parity unpinned" w.r.t. Bullet itself).
 *
 * All file:line citations are relative to /root/reference/gym_pybullet_drones/.
 * Quaternions are xyzw (pybullet order).  Every function works in IEEE double with
 * the reference's order of operations; compile with -ffp-contract=off.
 */
#ifndef GPD_ORACLE_H
#define GPD_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_CF2X = 0, ORC_CF2P = 1, ORC_RACE = 2 };                       /* utils/enums.py:3-8  */
enum { ORC_ACT_RPM = 0, ORC_ACT_PID = 1, ORC_ACT_VEL = 2,
       ORC_ACT_ONE_D_RPM = 3, ORC_ACT_ONE_D_PID = 4,                     /* utils/enums.py:35-41 */
       ORC_ACT_CTRL_RPM = 5,                                             /* envs/CtrlAviary.py:140 */
       ORC_ACT_CTRL_VEL = 6 };                                           /* envs/VelocityAviary.py:129-170 */
enum { ORC_ENV_CTRL = 0, ORC_ENV_HOVER = 1, ORC_ENV_MULTIHOVER = 2 };
enum { ORC_PHY_GND = 1, ORC_PHY_DRAG = 2, ORC_PHY_DW = 4 };              /* DYN-form composites (build-defined) */

/* envs/BaseAviary.py:97-128 (+ rotor link CoM offsets from assets/<model>.urdf) */
typedef struct orc_drone {
    int32_t model;
    int32_t _pad;
    double M, L, THRUST2WEIGHT;
    double J[3], J_INV[3];          /* diagonal inertia and its inverse */
    double KF, KM;
    double COLLISION_H, COLLISION_R, COLLISION_Z_OFFSET;
    double MAX_SPEED_KMH, GND_EFF_COEFF, PROP_RADIUS;
    double DRAG_COEFF[3];
    double DW_COEFF_1, DW_COEFF_2, DW_COEFF_3;
    double G, GRAVITY, HOVER_RPM, MAX_RPM, MAX_THRUST, MAX_XY_TORQUE, MAX_Z_TORQUE, GND_EFF_H_CLIP;
    double ROTOR_XYZ[4][3];
} orc_drone;

/* control/DSLPIDControl.py:37-60, control/BaseControl.py:35-39 */
typedef struct orc_pid {
    double P_FOR[3], I_FOR[3], D_FOR[3];
    double P_TOR[3], I_TOR[3], D_TOR[3];
    double PWM2RPM_SCALE, PWM2RPM_CONST, MIN_PWM, MAX_PWM;
    double MIXER[4][3];
    double GRAVITY, KF;             /* of the controller's own drone model */
} orc_pid;

typedef struct orc_env_cfg {
    int32_t num_drones;             /* N */
    int32_t substeps;               /* PYB_STEPS_PER_CTRL, BaseAviary.py:81 */
    int32_t pyb_freq;
    int32_t ctrl_freq;
    int32_t env_kind;
    int32_t action_type;
    int32_t physics_flags;
    int32_t action_buffer_size;     /* B = ctrl_freq//2, BaseRLAviary.py:66 */
    double episode_len_sec;         /* HoverAviary.py:52 */
    double speed_limit;             /* BaseRLAviary.py:95 (ActionType.VEL) */
    orc_drone drone;
    orc_pid pid;
} orc_env_cfg;

/* ---- Bullet closed forms used on the path (third-party; see header note) ---- */
void orc_matrix_from_quaternion(const double q[4], double m[9]);
void orc_euler_from_quaternion(const double q[4], double rpy[3]);
void orc_quaternion_from_euler(const double rpy[3], double q[4]);

/* ---- BaseAviary.py:876-889 ---- */
void orc_integrate_q(const double quat[4], const double omega[3], double dt, double out[4]);

/* ---- BaseAviary.py:831-874. One DYN substep for one drone.
 * pos/quat/vel: the substep-start snapshot (updated in place), rates: rpy_rates.
 * ang_v_out = R(old)·rates(new)  (BaseAviary.py:870).
 * Extra (DYN-form composite, build-defined): gnd[4] added to the rotor forces before
 * thrust and x/y torques are formed; f_ext_world[3] and f_ext_body[3] added to the
 * world force (body one rotated by R(old)).  Pass NULL for plain Physics.DYN. */
void orc_dynamics(const orc_drone* d, double dt, const double rpm[4],
                  double pos[3], double quat[4], double vel[3], double rates[3], double ang_v_out[3],
                  const double gnd[4], const double f_ext_world[3], const double f_ext_body[3]);

/* ---- force models: values exactly as handed to applyExternalForce ---- */
/* BaseAviary.py:715-750: out[4] = +z LINK-frame force on rotor links 0-3; returns 1 if the
 * |roll|,|pitch| < pi/2 gate (line 742) passes, else 0 (forces not applied). */
int orc_ground_effect(const orc_drone* d, const double rpm[4], const double pos[3], const double quat[4],
                      const double rpy[3], double out[4]);
/* BaseAviary.py:754-781: out[3] = CoM LINK-frame (body) drag force. */
void orc_drag(const orc_drone* d, const double rpm_prev[4], const double quat[4], const double vel[3], double out[3]);
/* BaseAviary.py:785-811: returns the sum over the other drones of the body-z (LINK frame) force on drone i;
 * pos_all is [N][3]. */
double orc_downwash(const orc_drone* d, int n, const double* pos_all, int i);

/* ---- control/DSLPIDControl.py:82-259.  pid_state[9] = integral_pos_e3, integral_rpy_e3, last_rpy3 ---- */
void orc_pid_compute(const orc_pid* c, double dt, const double cur_pos[3], const double cur_quat[4],
                     const double cur_vel[3], const double target_pos[3], const double target_rpy[3],
                     const double target_vel[3], const double target_rpy_rates[3],
                     double pid_state[9], double rpm_out[4], double pos_e_out[3], double* yaw_e_out);

/* ---- BaseAviary.py:1105-1147 ---- */
void orc_calculate_next_step(const double cur[3], const double dest[3], double step_size, double out[3]);

/*
 * ---- One env.step() for `num_envs` independent envs (BaseAviary.py:259-383) ----
 * Layout (row-major, all double unless noted):
 *   state20   [E][N][20]  pos3 quat4 rpy3 vel3 ang_v3 last_clipped_action4   (BaseAviary.py:559-561)
 *   rpy_rates [E][N][3]                                                       (BaseAviary.py:477,874)
 *   pid_state [E][N][9]   (NULL unless a PID-family action type)
 *   ring      [E][N][B][A] float32, oldest -> newest                          (BaseRLAviary.py:66-67,187)
 *   step_counter [E] int32                                                    (BaseAviary.py:460,382)
 *   actions   [E][N][A]: float32 for RL envs (SB3 dtype, BaseRLAviary.py:156), double for ORC_ACT_CTRL_RPM/_VEL
 *   target_pos [N][3]  (HoverAviary.py:51, MultiHoverAviary.py:71)
 *   obs: RL envs float32 [E][N][12+A*B] (BaseRLAviary.py:307-319); Ctrl env double [E][N][20] (CtrlAviary.py:117)
 *   reward [E] double, terminated/truncated [E] uint8
 * nthreads > 1 splits envs over pthreads (cpu_baseline only).
 */
void orc_step(const orc_env_cfg* cfg, int64_t num_envs,
              double* state20, double* rpy_rates, double* pid_state, float* ring, int32_t* step_counter,
              const void* actions, const double* target_pos,
              void* obs, double* reward, uint8_t* terminated, uint8_t* truncated, int nthreads);

/* BaseAviary.py:220-255 + _housekeeping :451-477 for the envs with mask[e]!=0 (NULL = all).
 * init_xyz/init_rpy are [E][N][3].  The action ring and the controllers are NOT reset
 * (BaseRLAviary.py:153-154,76).  Writes obs like orc_step when obs != NULL. */
void orc_reset(const orc_env_cfg* cfg, int64_t num_envs, const uint8_t* mask,
               const double* init_xyz, const double* init_rpy,
               double* state20, double* rpy_rates, const float* ring, int32_t* step_counter, void* obs);

int orc_action_width(int action_type);
int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
