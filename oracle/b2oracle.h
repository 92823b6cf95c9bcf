/* TEST INFRASTRUCTURE ONLY — CPU oracle for the gym-ignition env.step hot path.
 *
 * Plain-C fp64 restatement of the reference's per-step path (one world, one robot), used only as
 * the parity checker (tests/, __graft_entry__.smoke()) and as bench.py's cpu_baseline / reference
 * arm. The product (gym-ignition_b200/) never links, loads or calls anything in this directory.
 *
 * PARITY UNPINNED vs Ignition Gazebo + DART: the arithmetic of the reference's step lives in
 * third-party libraries that are not in /root/reference and cannot be built in this image
 * (DART 6.9/6.10 behind ignition-physics3 "dartsim", ignition-math6 PID, iDynTree),
 * and the reference holds no golden trajectories. The bookkeeping below (command application order,
 * one-shot commands, deferred resets, PID gating, task maths) follows files that ARE in the tree and
 * cites them; the dynamics follow the published algorithms of those libraries (Featherstone ABA in
 * body coordinates with DART's implicit joint damping/spring, semi-implicit Euler). The task-level
 * maths are pinned against the reference's own Python task classes through tests/golden/.
 */
#ifndef B2ORACLE_H
#define B2ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2O_MAXB 16

enum { B2O_REVOLUTE = 1, B2O_PRISMATIC = 2 };

/* cpp/scenario/core/include/scenario/core/Joint.h:36-45 (JointControlMode) */
enum {
    B2O_MODE_INVALID = 0,
    B2O_MODE_IDLE = 1,
    B2O_MODE_FORCE = 2,
    B2O_MODE_VELOCITY = 3,
    B2O_MODE_VELOCITY_FOLLOWER_DART = 4,
    B2O_MODE_POSITION = 5,
    B2O_MODE_POSITION_INTERPOLATED = 6
};

/* Tasks of python/gym_ignition_environments/tasks/ */
enum {
    B2O_TASK_PENDULUM_SWINGUP = 1,
    B2O_TASK_CARTPOLE_DISCRETE_BALANCING = 2,
    B2O_TASK_CARTPOLE_CONTINUOUS_BALANCING = 3,
    B2O_TASK_CARTPOLE_CONTINUOUS_SWINGUP = 4
};

/* Fixed-base tree of 1-DoF joints; bodies are sorted parents-first. All arrays row-major. */
typedef struct {
    int32_t nb;
    int32_t parent[B2O_MAXB];      /* -1: child of the fixed base */
    int32_t jtype[B2O_MAXB];
    double axis[B2O_MAXB][3];      /* joint axis in the child (= joint) frame */
    double R[B2O_MAXB][9];         /* child frame orientation in the parent body frame at q = 0 */
    double p[B2O_MAXB][3];         /* child frame origin in the parent body frame */
    double mass[B2O_MAXB];
    double com[B2O_MAXB][3];
    double Ic[B2O_MAXB][9];        /* rotational inertia about the COM, body axes */
    double damping[B2O_MAXB];
    double friction[B2O_MAXB];
    double stiffness[B2O_MAXB];
    double rest[B2O_MAXB];
    double lower[B2O_MAXB];
    double upper[B2O_MAXB];
    double effort[B2O_MAXB];
    double vmax[B2O_MAXB];
    double gravity[3];             /* world frame */
    double base_R[9];              /* world_H_base rotation */
    double base_p[3];
} b2o_model;

/* ignition::math::PID (ign-math6, not in tree; call sites JointController.cpp:309,312) */
typedef struct {
    double p, i, d, i_max, i_min, cmd_max, cmd_min, cmd_offset;
    double p_err_last, p_err, i_err, d_err, cmd;
} b2o_pid;

void b2o_pid_init(b2o_pid* pid, double p, double i, double d, double i_max, double i_min,
                  double cmd_max, double cmd_min, double cmd_offset);
void b2o_pid_reset(b2o_pid* pid);
double b2o_pid_update(b2o_pid* pid, double error, double dt);

/* --- rigid body dynamics ------------------------------------------------------------------- */
/* DART-style articulated-body forward dynamics with implicit damping/spring (dt_implicit = step). */
void b2o_forward_dynamics(const b2o_model* m, double dt_implicit, const double* q, const double* dq,
                          const double* tau, double* ddq);
/* RNEA: tau = M ddq + C dq + g (no damping/spring terms). */
void b2o_inverse_dynamics(const b2o_model* m, const double* q, const double* dq, const double* ddq,
                          int with_gravity, double* tau);
/* Joint-space mass matrix, nb x nb row-major (columns by RNEA). */
void b2o_mass_matrix(const b2o_model* m, const double* q, double* M);
/* World pose of every body frame: R[nb][9], p[nb][3]. */
void b2o_forward_kinematics(const b2o_model* m, const double* q, double* R, double* p);
/* Jacobian of a point fixed in `body` (point given in the body frame), world orientation
 * (iDynTree MIXED): J[6][nb] row-major, rows 0-2 linear, rows 3-5 angular. */
void b2o_point_jacobian(const b2o_model* m, const double* q, int body, const double* point,
                        double* J);
/* One DART World::step restatement: ABA -> dq += ddq dt -> joint constraints -> q += dq dt. */
void b2o_physics_step(const b2o_model* m, double dt, double* q, double* dq, const double* tau,
                      double* ddq);
/* Total mechanical energy (kinetic + gravitational potential), for conservation checks. */
double b2o_energy(const b2o_model* m, const double* q, const double* dq);

/* --- single-world ScenarI/O-semantics simulator ---------------------------------------------- */
typedef struct b2o_sim b2o_sim;
b2o_sim* b2o_sim_create(const b2o_model* m, double step_size, int steps_per_run);
void b2o_sim_destroy(b2o_sim* s);
int b2o_sim_run(b2o_sim* s, int paused);
double b2o_sim_time(const b2o_sim* s);
int b2o_sim_set_control_mode(b2o_sim* s, int joint, int mode);
int b2o_sim_control_mode(const b2o_sim* s, int joint);
int b2o_sim_set_pid(b2o_sim* s, int joint, double p, double i, double d, double i_max,
                    double i_min, double cmd_max, double cmd_min, double cmd_offset);
int b2o_sim_set_controller_period(b2o_sim* s, double period);
int b2o_sim_set_force_target(b2o_sim* s, int joint, double f);
int b2o_sim_set_position_target(b2o_sim* s, int joint, double v);
int b2o_sim_set_velocity_target(b2o_sim* s, int joint, double v);
int b2o_sim_set_acceleration_target(b2o_sim* s, int joint, double v);
int b2o_sim_load_computed_torque(b2o_sim* s, const double* kp, const double* kd, const double* gravity);
int b2o_sim_apply_link_wrench(b2o_sim* s, int body, const double* point, const double* wrench, double duration);
int b2o_sim_reset_position(b2o_sim* s, int joint, double v);
int b2o_sim_reset_velocity(b2o_sim* s, int joint, double v);
double b2o_sim_position(const b2o_sim* s, int joint);
double b2o_sim_velocity(const b2o_sim* s, int joint);
double b2o_sim_acceleration(const b2o_sim* s, int joint);
double b2o_sim_force_target(const b2o_sim* s, int joint, int* has_target);
double b2o_sim_position_target(const b2o_sim* s, int joint, int* has_target);
double b2o_sim_velocity_target(const b2o_sim* s, int joint, int* has_target);

/* --- tasks + batched rollout (the fused hot-path semantics) --------------------------------- */
void b2o_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* uniform doubles in [0,1): n values for (seed, global env index, step index) */
void b2o_reset_uniforms(uint64_t seed, uint64_t env, uint64_t step, int n, double* u);
int b2o_task_nobs(int task);
int b2o_task_nq(int task);
/* Fresh episode state from the task's reset distribution. state = [q.., dq..] */
void b2o_task_sample_reset(int task, uint64_t seed, uint64_t env, uint64_t step, double* state);
void b2o_task_reset_from_uniforms(int task, const double* u, double* state);
/* Task maths on a post-step state. Returns done (task termination only, no time limit). */
int b2o_task_evaluate(int task, const double* state, double force_target_after_step, double* obs,
                      double* reward);
/* Joint force applied by Task.set_action for this action value; *joint receives the actuated dof. */
double b2o_task_action_force(int task, double action, int* joint);

/* Batched rollout: n_envs independent worlds, T steps, actions[T][n_envs].
 * state[n_envs][2nq] in/out, elapsed[n_envs] in/out; obs[T][n_envs][nobs], reward[T][n_envs],
 * done[T][n_envs] may be NULL. first_step is the global step index of the first step (Philox counter).
 * env_offset is the global index of env 0 (multi-GPU sharding). */
void b2o_rollout(const b2o_model* m, int task, double dt, int steps_per_run, int max_episode_steps, uint64_t seed,
                 uint64_t env_offset, uint64_t first_step, int n_envs, int T, const double* actions,
                 double* state, int32_t* elapsed, double* obs, double* reward, uint8_t* done);

void b2o_link_motion(const b2o_model* m, const double* q, const double* dq, const double* ddq, int body,
                     const double* point, double* out);

/* --- KinDyn centre of mass / momentum (kindyncomputations.py:305-342) ------------------------------------------ */
void b2o_centroidal(const b2o_model* m, const double* q, const double* dq, double base_mass, const double* base_mc,
                    double* com, double* com_velocity, double* momentum, double* centroidal, double* Jcom);
void b2o_momentum_jacobian(const b2o_model* m, const double* q, double base_mass, const double* base_mc,
                           const double* base_Io, double* Jmom, double* locked);

/* --- per-env domain randomisation ------------------------------------------------------------------ */
void b2o_sample_rand_params(uint64_t seed, uint64_t env, uint64_t step, int nq, double delta, double sigma,
                            double g0, const double* mass, double* out);
void b2o_rollout_randomized(const b2o_model* m, int task, double dt, int steps_per_run, int max_episode_steps,
                            uint64_t seed, uint64_t env_offset, uint64_t first_step, int n_envs, int T,
                            const double* actions, double* state, int32_t* elapsed, double* rand,
                            double mass_delta, double gravity_sigma, double* obs, double* reward, uint8_t* done);

/* --- free rigid bodies with plane / box contacts ----------------------------------------------- */
#define B2O_MAXFREE 8
#define B2O_MAXSTATIC 16
#define B2O_MAXCONTACTS 32
enum { B2O_SHAPE_BOX = 0, B2O_SHAPE_SPHERE = 1, B2O_SHAPE_CYLINDER = 2, B2O_SHAPE_PLANE = 3 };
typedef struct {
    int32_t type;
    double size[3];   /* box: half extents; sphere: radius; plane: unit normal */
    double R[9];      /* in the body frame (free bodies) or in the world (static shapes) */
    double p[3];
    double mu;
} b2o_shape;
typedef struct {
    double mass;
    double Ic[9];
    double com[3];
    int32_t nshapes;
    b2o_shape shape[2];
} b2o_free_body;
typedef struct {
    int32_t nfree, nstatic, iterations;
    double dt, erp, max_erv;
    double g[3];
    b2o_free_body body[B2O_MAXFREE];
    b2o_shape stat[B2O_MAXSTATIC];
    /* external wrench on free body i for this step: force and torque in the world frame, the force acting at the origin
     * of the body's root link (Link::applyWorldWrench on a free body, Physics.cpp:1483-1532; the caller clears it when its
     * duration has expired) */
    double ext[B2O_MAXFREE][6];
} b2o_world;
typedef struct {
    int32_t a, b;     /* free body index, -1 - static shape (b only), or -1000 - robot shape (coupled worlds) */
    double pos[3], n[3], depth, force[3];
} b2o_contact;
/* X: nfree x 13 (position, quaternion wxyz, linear velocity, angular velocity). Returns the contact count. */
int b2o_world_step(const b2o_world* w, double* X, b2o_contact* out, int max_out);

/* --- coupled world: articulated model with link collision shapes + free bodies + static shapes -------- */
typedef struct {
    int32_t nrobot;
    int32_t rbody[4];      /* body (moving joint) each shape is attached to */
    b2o_shape rshape[4];   /* pose in that body's frame */
} b2o_robot_shapes;
/* Constraint stage + free-body integration of one step. q: joint positions; dq: joint velocities after the
 * unconstrained update on entry, constrained on return (positions are integrated by the caller).
 * servo / servo_target may be NULL. Returns the contact count (-1: singular mass matrix). */
int b2o_coupled_step(const b2o_world* w, const b2o_model* m, const b2o_robot_shapes* rs, const double* q,
                     double* dq, const int* servo, const double* servo_target, double* X, b2o_contact* out,
                     int max_out);
/* Attach free bodies / contacts to a single-world simulator (its model provides the articulated side). */
int b2o_sim_attach_world(b2o_sim* s, const b2o_world* w, const b2o_robot_shapes* rs, const double* X0);
void b2o_sim_world_state(const b2o_sim* s, double* X);
int b2o_sim_contacts(const b2o_sim* s, b2o_contact* out, int max_out);
int b2o_sim_reset_base(b2o_sim* s, int body, int velocity, const double* values);

#ifdef __cplusplus
}
#endif
#endif
