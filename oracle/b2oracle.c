/* TEST INFRASTRUCTURE ONLY — see b2oracle.h for the scope note and the "parity unpinned" statement.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (see oracle/Makefile). -ffp-contract=off keeps
 * every product/sum individually rounded, which is what the reference's Python task code does.
 */
#include "b2oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------- */
/* small dense helpers                                                                           */
/* ------------------------------------------------------------------------------------------- */
typedef double v3[3];
typedef double v6[6];
typedef double m3[9];
typedef double m6[36];

static void cross3(const double* a, const double* b, double* c)
{
    double x = a[1] * b[2] - a[2] * b[1];
    double y = a[2] * b[0] - a[0] * b[2];
    double z = a[0] * b[1] - a[1] * b[0];
    c[0] = x; c[1] = y; c[2] = z;
}
static void m3v(const double* A, const double* x, double* y)
{
    double r[3];
    for (int i = 0; i < 3; i++) r[i] = A[3 * i] * x[0] + A[3 * i + 1] * x[1] + A[3 * i + 2] * x[2];
    memcpy(y, r, sizeof r);
}
static void m3tv(const double* A, const double* x, double* y)
{
    double r[3];
    for (int i = 0; i < 3; i++) r[i] = A[i] * x[0] + A[3 + i] * x[1] + A[6 + i] * x[2];
    memcpy(y, r, sizeof r);
}
static void m3m(const double* A, const double* B, double* C)
{
    double r[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            r[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
    memcpy(C, r, sizeof r);
}
static void skew(const double* p, double* S)
{
    S[0] = 0; S[1] = -p[2]; S[2] = p[1];
    S[3] = p[2]; S[4] = 0; S[5] = -p[0];
    S[6] = -p[1]; S[7] = p[0]; S[8] = 0;
}
static void rodrigues(const double* a, double q, double* R)
{
    double c = cos(q), s = sin(q), t = 1.0 - c;
    R[0] = c + t * a[0] * a[0];        R[1] = t * a[0] * a[1] - s * a[2]; R[2] = t * a[0] * a[2] + s * a[1];
    R[3] = t * a[0] * a[1] + s * a[2]; R[4] = c + t * a[1] * a[1];        R[5] = t * a[1] * a[2] - s * a[0];
    R[6] = t * a[0] * a[2] - s * a[1]; R[7] = t * a[1] * a[2] + s * a[0]; R[8] = c + t * a[2] * a[2];
}
static void m6v(const double* A, const double* x, double* y)
{
    double r[6];
    for (int i = 0; i < 6; i++) {
        double s = 0;
        for (int j = 0; j < 6; j++) s += A[6 * i + j] * x[j];
        r[i] = s;
    }
    memcpy(y, r, sizeof r);
}
static void m6tv(const double* A, const double* x, double* y)
{
    double r[6];
    for (int i = 0; i < 6; i++) {
        double s = 0;
        for (int j = 0; j < 6; j++) s += A[6 * j + i] * x[j];
        r[i] = s;
    }
    memcpy(y, r, sizeof r);
}
static double dot6(const double* a, const double* b)
{
    double s = 0;
    for (int i = 0; i < 6; i++) s += a[i] * b[i];
    return s;
}
/* spatial motion cross product: crm(v) w */
static void crm(const double* v, const double* w, double* r)
{
    double a[3], b[3], c[3], out[6];
    cross3(v, w, a);
    cross3(v, w + 3, b);
    cross3(v + 3, w, c);
    out[0] = a[0]; out[1] = a[1]; out[2] = a[2];
    out[3] = b[0] + c[0]; out[4] = b[1] + c[1]; out[5] = b[2] + c[2];
    memcpy(r, out, sizeof out);
}
/* spatial force cross product: crf(v) f = v x* f */
static void crf(const double* v, const double* f, double* r)
{
    double a[3], b[3], c[3], out[6];
    cross3(v, f, a);          /* w x n */
    cross3(v + 3, f + 3, b);  /* v x f */
    cross3(v, f + 3, c);      /* w x f */
    out[0] = a[0] + b[0]; out[1] = a[1] + b[1]; out[2] = a[2] + b[2];
    out[3] = c[0]; out[4] = c[1]; out[5] = c[2];
    memcpy(r, out, sizeof out);
}

/* Motion transform parent -> child for a child frame placed at (R, p) in the parent:
 *   w_c = R^T w_p ;  v_c = R^T (v_p + w_p x p)  */
static void motion_xform(const double* R, const double* p, double* X)
{
    double Rt[9], S[9], RtS[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) Rt[3 * i + j] = R[3 * j + i];
    skew(p, S);
    m3m(Rt, S, RtS);
    memset(X, 0, sizeof(m6));
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            X[6 * i + j] = Rt[3 * i + j];
            X[6 * (i + 3) + (j + 3)] = Rt[3 * i + j];
            X[6 * (i + 3) + j] = -RtS[3 * i + j];
        }
}

/* spatial inertia about the body origin from (m, c, Ic) */
static void spatial_inertia(double m, const double* c, const double* Ic, double* I6)
{
    double C[9], CC[9];
    skew(c, C);
    m3m(C, C, CC);
    memset(I6, 0, sizeof(m6));
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            I6[6 * i + j] = Ic[3 * i + j] - m * CC[3 * i + j];
            I6[6 * i + (j + 3)] = m * C[3 * i + j];
            I6[6 * (i + 3) + j] = -m * C[3 * i + j];
        }
    I6[6 * 3 + 3] = m; I6[6 * 4 + 4] = m; I6[6 * 5 + 5] = m;
}

/* joint placement at q: child frame pose in the parent body frame, and the motion subspace */
static void joint_pose(const b2o_model* m, int i, double q, double* R, double* p, double* S)
{
    memset(S, 0, sizeof(v6));
    if (m->jtype[i] == B2O_REVOLUTE) {
        double Rq[9];
        rodrigues(m->axis[i], q, Rq);
        m3m(m->R[i], Rq, R);
        memcpy(p, m->p[i], sizeof(v3));
        S[0] = m->axis[i][0]; S[1] = m->axis[i][1]; S[2] = m->axis[i][2];
    } else {
        double d[3] = {m->axis[i][0] * q, m->axis[i][1] * q, m->axis[i][2] * q}, Rd[3];
        memcpy(R, m->R[i], sizeof(m3));
        m3v(m->R[i], d, Rd);
        p[0] = m->p[i][0] + Rd[0]; p[1] = m->p[i][1] + Rd[1]; p[2] = m->p[i][2] + Rd[2];
        S[3] = m->axis[i][0]; S[4] = m->axis[i][1]; S[5] = m->axis[i][2];
    }
}

/* ------------------------------------------------------------------------------------------- */
/* ignition::math::PID (ign-math6). error = current - reference (JointController.cpp:308).       */
/* ------------------------------------------------------------------------------------------- */
void b2o_pid_init(b2o_pid* pid, double p, double i, double d, double i_max, double i_min,
                  double cmd_max, double cmd_min, double cmd_offset)
{
    pid->p = p; pid->i = i; pid->d = d;
    pid->i_max = i_max; pid->i_min = i_min;
    pid->cmd_max = cmd_max; pid->cmd_min = cmd_min; pid->cmd_offset = cmd_offset;
    b2o_pid_reset(pid);
}
void b2o_pid_reset(b2o_pid* pid)
{
    pid->p_err_last = 0; pid->p_err = 0; pid->i_err = 0; pid->d_err = 0; pid->cmd = 0;
}
static double clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }
double b2o_pid_update(b2o_pid* pid, double error, double dt)
{
    if (dt == 0.0 || isnan(error) || isinf(error)) return 0.0;
    pid->p_err = error;
    double p_term = pid->p * pid->p_err;
    pid->i_err = pid->i_err + pid->i * dt * pid->p_err;
    if (pid->i_max >= pid->i_min) pid->i_err = clampd(pid->i_err, pid->i_min, pid->i_max);
    pid->d_err = (pid->p_err - pid->p_err_last) / dt;
    pid->p_err_last = pid->p_err;
    double d_term = pid->d * pid->d_err;
    pid->cmd = pid->cmd_offset - p_term - pid->i_err - d_term;
    if (pid->cmd_max >= pid->cmd_min) pid->cmd = clampd(pid->cmd, pid->cmd_min, pid->cmd_max);
    return pid->cmd;
}

/* ------------------------------------------------------------------------------------------- */
/* kinematics pass shared by the algorithms                                                      */
/* ------------------------------------------------------------------------------------------- */
typedef struct {
    m6 X[B2O_MAXB];     /* parent -> child motion transform */
    v6 S[B2O_MAXB];
    m3 Rw[B2O_MAXB];    /* world orientation of the body frame */
    v3 pw[B2O_MAXB];    /* world position of the body origin */
    m6 I[B2O_MAXB];
} kin_t;

static void kinematics(const b2o_model* m, const double* q, kin_t* k)
{
    for (int i = 0; i < m->nb; i++) {
        double R[9], p[3];
        joint_pose(m, i, q[i], R, p, k->S[i]);
        motion_xform(R, p, k->X[i]);
        const double* Rp = m->parent[i] < 0 ? m->base_R : k->Rw[m->parent[i]];
        const double* pp = m->parent[i] < 0 ? m->base_p : k->pw[m->parent[i]];
        double Rpp[3];
        m3m(Rp, R, k->Rw[i]);
        m3v(Rp, p, Rpp);
        for (int a = 0; a < 3; a++) k->pw[i][a] = pp[a] + Rpp[a];
        spatial_inertia(m->mass[i], m->com[i], m->Ic[i], k->I[i]);
    }
}

/* ------------------------------------------------------------------------------------------- */
/* Forward dynamics. Restates DART 6.x Skeleton::computeForwardDynamics (not in tree; reached    */
/* from cpp/scenario/plugins/Physics/Physics.cpp:1824-1835):                                     */
/*   BodyNode::updateArtInertia / updateBiasForce / updateAccelerationFD with                   */
/*   GenericJoint::updateInvProjArtInertiaImplicitDynamic  psi = (S'AS + dt D + dt^2 K)^-1       */
/*   GenericJoint::updateTotalForceDynamic  u = tau - D dq - K (q - q0 + dt dq) - S'(A eta + B)  */
/* ------------------------------------------------------------------------------------------- */
void b2o_forward_dynamics(const b2o_model* m, double dt, const double* q, const double* dq,
                          const double* tau, double* ddq)
{
    kin_t k;
    v6 V[B2O_MAXB], eta[B2O_MAXB], B[B2O_MAXB], AIS[B2O_MAXB], A[B2O_MAXB];
    m6 AI[B2O_MAXB];
    double psi[B2O_MAXB], u[B2O_MAXB];
    const int nb = m->nb;

    kinematics(m, q, &k);
    for (int i = 0; i < nb; i++) {
        v6 Vp = {0, 0, 0, 0, 0, 0}, Sdq;
        if (m->parent[i] >= 0) m6v(k.X[i], V[m->parent[i]], Vp);
        for (int a = 0; a < 6; a++) { Sdq[a] = k.S[i][a] * dq[i]; V[i][a] = Vp[a] + Sdq[a]; }
        crm(V[i], Sdq, eta[i]);
        /* bias force: V x* (I V) - I [0; Rw' g] */
        v6 IV, g6 = {0, 0, 0, 0, 0, 0}, Fg;
        m6v(k.I[i], V[i], IV);
        crf(V[i], IV, B[i]);
        m3tv(k.Rw[i], m->gravity, g6 + 3);
        m6v(k.I[i], g6, Fg);
        for (int a = 0; a < 6; a++) B[i][a] -= Fg[a];
        memcpy(AI[i], k.I[i], sizeof(m6));
    }
    for (int i = nb - 1; i >= 0; i--) {
        v6 tmp, AIeta;
        m6v(AI[i], k.S[i], AIS[i]);
        double d = dot6(k.S[i], AIS[i]) + dt * m->damping[i] + dt * dt * m->stiffness[i];
        psi[i] = 1.0 / d;
        m6v(AI[i], eta[i], AIeta);
        for (int a = 0; a < 6; a++) tmp[a] = AIeta[a] + B[i][a];
        u[i] = tau[i] - m->damping[i] * dq[i]
               - m->stiffness[i] * (q[i] - m->rest[i] + dt * dq[i]) - dot6(k.S[i], tmp);
        int p = m->parent[i];
        if (p >= 0) {
            /* Pi = AI - AIS psi AIS' ; beta = B + AI (eta + S psi u) */
            m6 Pi, XtPi, XtPiX;
            v6 beta, acc;
            for (int a = 0; a < 6; a++)
                for (int b = 0; b < 6; b++)
                    Pi[6 * a + b] = AI[i][6 * a + b] - AIS[i][a] * psi[i] * AIS[i][b];
            for (int a = 0; a < 6; a++) acc[a] = eta[i][a] + k.S[i][a] * (psi[i] * u[i]);
            m6v(AI[i], acc, beta);
            for (int a = 0; a < 6; a++) beta[a] += B[i][a];
            /* parent += X' Pi X ; parent bias += X' beta */
            for (int a = 0; a < 6; a++)
                for (int b = 0; b < 6; b++) {
                    double s = 0;
                    for (int c = 0; c < 6; c++) s += k.X[i][6 * c + a] * Pi[6 * c + b];
                    XtPi[6 * a + b] = s;
                }
            for (int a = 0; a < 6; a++)
                for (int b = 0; b < 6; b++) {
                    double s = 0;
                    for (int c = 0; c < 6; c++) s += XtPi[6 * a + c] * k.X[i][6 * c + b];
                    XtPiX[6 * a + b] = s;
                }
            for (int a = 0; a < 36; a++) AI[p][a] += XtPiX[a];
            m6tv(k.X[i], beta, tmp);
            for (int a = 0; a < 6; a++) B[p][a] += tmp[a];
        }
    }
    for (int i = 0; i < nb; i++) {
        v6 Ap = {0, 0, 0, 0, 0, 0};
        if (m->parent[i] >= 0) m6v(k.X[i], A[m->parent[i]], Ap);
        ddq[i] = psi[i] * (u[i] - dot6(AIS[i], Ap));
        for (int a = 0; a < 6; a++) A[i][a] = Ap[a] + eta[i][a] + k.S[i][a] * ddq[i];
    }
}

/* Recursive Newton-Euler (Featherstone RBDA ch.5), body coordinates. */
void b2o_inverse_dynamics(const b2o_model* m, const double* q, const double* dq, const double* ddq,
                          int with_gravity, double* tau)
{
    kin_t k;
    v6 V[B2O_MAXB], A[B2O_MAXB], F[B2O_MAXB];
    const int nb = m->nb;
    kinematics(m, q, &k);
    for (int i = 0; i < nb; i++) {
        v6 Vp = {0, 0, 0, 0, 0, 0}, Ap = {0, 0, 0, 0, 0, 0}, Sdq, eta, IA, IV, VxIV;
        if (m->parent[i] >= 0) {
            m6v(k.X[i], V[m->parent[i]], Vp);
            m6v(k.X[i], A[m->parent[i]], Ap);
        } else if (with_gravity) {
            /* fictitious base acceleration -g, expressed in the base frame then moved to the child */
            v6 a0 = {0, 0, 0, 0, 0, 0};
            double gb[3];
            m3tv(m->base_R, m->gravity, gb);
            a0[3] = -gb[0]; a0[4] = -gb[1]; a0[5] = -gb[2];
            m6v(k.X[i], a0, Ap);
        }
        for (int a = 0; a < 6; a++) { Sdq[a] = k.S[i][a] * dq[i]; V[i][a] = Vp[a] + Sdq[a]; }
        crm(V[i], Sdq, eta);
        for (int a = 0; a < 6; a++) A[i][a] = Ap[a] + eta[a] + k.S[i][a] * ddq[i];
        m6v(k.I[i], A[i], IA);
        m6v(k.I[i], V[i], IV);
        crf(V[i], IV, VxIV);
        for (int a = 0; a < 6; a++) F[i][a] = IA[a] + VxIV[a];
    }
    for (int i = nb - 1; i >= 0; i--) {
        tau[i] = dot6(k.S[i], F[i]);
        if (m->parent[i] >= 0) {
            v6 t;
            m6tv(k.X[i], F[i], t);
            for (int a = 0; a < 6; a++) F[m->parent[i]][a] += t[a];
        }
    }
}

void b2o_mass_matrix(const b2o_model* m, const double* q, double* M)
{
    const int nb = m->nb;
    double zero[B2O_MAXB] = {0}, e[B2O_MAXB], col[B2O_MAXB];
    for (int j = 0; j < nb; j++) {
        memset(e, 0, sizeof e);
        e[j] = 1.0;
        b2o_inverse_dynamics(m, q, zero, e, 0, col);
        for (int i = 0; i < nb; i++) M[i * nb + j] = col[i];
    }
}

void b2o_forward_kinematics(const b2o_model* m, const double* q, double* R, double* p)
{
    kin_t k;
    kinematics(m, q, &k);
    for (int i = 0; i < m->nb; i++) {
        memcpy(R + 9 * i, k.Rw[i], sizeof(m3));
        memcpy(p + 3 * i, k.pw[i], sizeof(v3));
    }
}

void b2o_point_jacobian(const b2o_model* m, const double* q, int body, const double* point,
                        double* J)
{
    kin_t k;
    const int nb = m->nb;
    kinematics(m, q, &k);
    double Rp[3], pt[3];
    memset(J, 0, sizeof(double) * 6 * nb);
    if (body < 0) return;
    m3v(k.Rw[body], point, Rp);
    for (int a = 0; a < 3; a++) pt[a] = k.pw[body][a] + Rp[a];
    for (int i = body; i >= 0; i = m->parent[i]) {
        double aw[3], r[3], lin[3];
        m3v(k.Rw[i], m->axis[i], aw);
        if (m->jtype[i] == B2O_REVOLUTE) {
            for (int a = 0; a < 3; a++) r[a] = pt[a] - k.pw[i][a];
            cross3(aw, r, lin);
            for (int a = 0; a < 3; a++) { J[a * nb + i] = lin[a]; J[(a + 3) * nb + i] = aw[a]; }
        } else {
            for (int a = 0; a < 3; a++) J[a * nb + i] = aw[a];
        }
    }
}

double b2o_energy(const b2o_model* m, const double* q, const double* dq)
{
    kin_t k;
    v6 V[B2O_MAXB];
    double E = 0;
    kinematics(m, q, &k);
    for (int i = 0; i < m->nb; i++) {
        v6 Vp = {0, 0, 0, 0, 0, 0}, IV;
        double cw[3];
        if (m->parent[i] >= 0) m6v(k.X[i], V[m->parent[i]], Vp);
        for (int a = 0; a < 6; a++) V[i][a] = Vp[a] + k.S[i][a] * dq[i];
        m6v(k.I[i], V[i], IV);
        E += 0.5 * dot6(V[i], IV);
        m3v(k.Rw[i], m->com[i], cw);
        for (int a = 0; a < 3; a++) E -= m->mass[i] * m->gravity[a] * (k.pw[i][a] + cw[a]);
    }
    return E;
}

/* ------------------------------------------------------------------------------------------- */
/* Joint-space constraint stage of DART's World::step (ConstraintSolver::solve, not in tree):    */
/* JointLimitConstraint (ign-physics enforces SDF limits), JointCoulombFrictionConstraint and     */
/* ServoMotorConstraint are 1-D rows on joint velocities; the boxed LCP                          */
/*   w = A lambda + b,  A = rows/cols of M^-1 (impulse response, no implicit damping)            */
/* is solved here by projected Gauss-Seidel to a tight tolerance (DART: Dantzig, same solution). */
/* ------------------------------------------------------------------------------------------- */
static int cholesky_solve_inplace(int n, double* M, double* Minv)
{
    /* Minv = M^-1 via Cholesky; M is overwritten by L. */
    for (int j = 0; j < n; j++) {
        double s = M[j * n + j];
        for (int k = 0; k < j; k++) s -= M[j * n + k] * M[j * n + k];
        if (s <= 0) return 0;
        M[j * n + j] = sqrt(s);
        for (int i = j + 1; i < n; i++) {
            double t = M[i * n + j];
            for (int k = 0; k < j; k++) t -= M[i * n + k] * M[j * n + k];
            M[i * n + j] = t / M[j * n + j];
        }
    }
    for (int c = 0; c < n; c++) {
        double y[B2O_MAXB];
        for (int i = 0; i < n; i++) {
            double t = (i == c) ? 1.0 : 0.0;
            for (int k = 0; k < i; k++) t -= M[i * n + k] * y[k];
            y[i] = t / M[i * n + i];
        }
        for (int i = n - 1; i >= 0; i--) {
            double t = y[i];
            for (int k = i + 1; k < n; k++) t -= M[k * n + i] * Minv[k * n + c];
            Minv[i * n + c] = t / M[i * n + i];
        }
    }
    return 1;
}

typedef struct { int joint; double b, lo, hi; } row_t;

static void joint_constraints(const b2o_model* m, double dt, const double* q, double* dq,
                              const int* servo, const double* servo_target, double* ddq)
{
    row_t rows[3 * B2O_MAXB];
    int nr = 0;
    const int nb = m->nb;
    for (int j = 0; j < nb; j++) {
        if (servo && servo[j]) {
            rows[nr++] = (row_t){j, dq[j] - servo_target[j], -m->effort[j] * dt, m->effort[j] * dt};
            continue;
        }
        if (m->friction[j] != 0.0)
            rows[nr++] = (row_t){j, dq[j], -m->friction[j] * dt, m->friction[j] * dt};
        if (q[j] <= m->lower[j]) rows[nr++] = (row_t){j, dq[j], 0.0, INFINITY};
        if (q[j] >= m->upper[j]) rows[nr++] = (row_t){j, dq[j], -INFINITY, 0.0};
    }
    if (nr == 0) return;
    double M[B2O_MAXB * B2O_MAXB], Minv[B2O_MAXB * B2O_MAXB], lam[3 * B2O_MAXB] = {0};
    b2o_mass_matrix(m, q, M);
    if (!cholesky_solve_inplace(nb, M, Minv)) return;
    for (int it = 0; it < 200; it++) {
        double change = 0;
        for (int a = 0; a < nr; a++) {
            double w = rows[a].b;
            for (int c = 0; c < nr; c++) w += Minv[rows[a].joint * nb + rows[c].joint] * lam[c];
            double nl = clampd(lam[a] - w / Minv[rows[a].joint * nb + rows[a].joint],
                               rows[a].lo, rows[a].hi);
            change += fabs(nl - lam[a]);
            lam[a] = nl;
        }
        if (change < 1e-18) break;
    }
    for (int i = 0; i < nb; i++) {
        double dv = 0;
        for (int a = 0; a < nr; a++) dv += Minv[i * nb + rows[a].joint] * lam[a];
        dq[i] += dv;
        if (ddq) ddq[i] += dv / dt;
    }
}

static void physics_step_ex(const b2o_model* m, double dt, double* q, double* dq,
                            const double* tau, const int* servo, const double* servo_target,
                            double* ddq)
{
    double acc[B2O_MAXB];
    b2o_forward_dynamics(m, dt, q, dq, tau, acc);
    for (int i = 0; i < m->nb; i++) dq[i] += acc[i] * dt;      /* Skeleton::integrateVelocities */
    joint_constraints(m, dt, q, dq, servo, servo_target, acc);  /* ConstraintSolver::solve + impulses */
    for (int i = 0; i < m->nb; i++) q[i] += dq[i] * dt;        /* Skeleton::integratePositions */
    if (ddq) memcpy(ddq, acc, sizeof(double) * m->nb);
}

void b2o_physics_step(const b2o_model* m, double dt, double* q, double* dq, const double* tau,
                      double* ddq)
{
    physics_step_ex(m, dt, q, dq, tau, NULL, NULL, ddq);
}

/* ------------------------------------------------------------------------------------------- */
/* Single-world simulator with the ScenarI/O bookkeeping semantics                               */
/* ------------------------------------------------------------------------------------------- */
struct b2o_sim {
    b2o_model model;
    /* optional coupled world (free bodies + contacts with the model's link shapes) */
    int has_world;
    b2o_world world;
    b2o_robot_shapes rshapes;
    double X[13 * B2O_MAXFREE], X_reset[13 * B2O_MAXFREE];
    int base_reset[B2O_MAXFREE];   /* bit 0: pose pending, bit 1: velocity pending */
    b2o_contact contacts[B2O_MAXCONTACTS];
    int ncontacts;
    int64_t dt_ns, time_ns, prev_update_ns, period_ns;
    int steps_per_run, controller_loaded;
    double q[B2O_MAXB], dq[B2O_MAXB], ddq[B2O_MAXB], tau_read[B2O_MAXB];
    int mode[B2O_MAXB];
    b2o_pid pid[B2O_MAXB];
    int has_force_cmd[B2O_MAXB], has_vel_cmd[B2O_MAXB];
    double force_cmd[B2O_MAXB], vel_cmd[B2O_MAXB];
    int has_pos_target[B2O_MAXB], has_vel_target[B2O_MAXB];
    double pos_target[B2O_MAXB], vel_target[B2O_MAXB];
    int pos_reset[B2O_MAXB], vel_reset[B2O_MAXB];
    double pos_reset_v[B2O_MAXB], vel_reset_v[B2O_MAXB];
    /* ComputedTorqueFixedBase run by ControllerRunner */
    int ct_loaded, ct_refs_set;
    int64_t ct_prev_update_ns;
    int has_acc_target[B2O_MAXB];
    double acc_target[B2O_MAXB], ct_kp[B2O_MAXB], ct_kd[B2O_MAXB], ct_gravity[3], ct_tau[B2O_MAXB];
    /* external link wrenches with duration: body the link is welded to + link origin in that body */
    int nwrench;
    struct { int body; double point[3]; double w[6]; int64_t expiry_ns; } wrench[8];
};

static int64_t to_ns(double seconds) { return (int64_t)llround(seconds * 1e9); } /* helpers.cpp:98-108 */

b2o_sim* b2o_sim_create(const b2o_model* m, double step_size, int steps_per_run)
{
    if (step_size <= 0 || steps_per_run <= 0) return NULL; /* GazeboSimulator.cpp:169-195 */
    b2o_sim* s = (b2o_sim*)calloc(1, sizeof(b2o_sim));
    s->model = *m;
    s->dt_ns = to_ns(step_size);
    s->steps_per_run = steps_per_run;
    s->period_ns = INT64_MAX;                               /* Model.cpp:180-185 duration::max() */
    for (int j = 0; j < m->nb; j++) {
        s->mode[j] = B2O_MODE_IDLE;                         /* Joint.cpp:126-127 */
        b2o_pid_init(&s->pid[j], 1, 0.1, 0.01, -1, 0, -1, 0, 0); /* DefaultPID, Joint.cpp:63 */
    }
    return s;
}
void b2o_sim_destroy(b2o_sim* s) { free(s); }
double b2o_sim_time(const b2o_sim* s) { return (double)s->time_ns / 1e9; }

/* Joint.cpp:369-460 */
int b2o_sim_set_control_mode(b2o_sim* s, int j, int mode)
{
    if (mode == B2O_MODE_POSITION_INTERPOLATED || mode == B2O_MODE_INVALID) return 0;
    if (mode == B2O_MODE_POSITION || mode == B2O_MODE_VELOCITY ||
        mode == B2O_MODE_VELOCITY_FOLLOWER_DART)
        s->controller_loaded = 1;
    s->mode[j] = mode;
    s->has_pos_target[j] = s->has_vel_target[j] = 0;
    s->has_acc_target[j] = 0;
    s->has_vel_cmd[j] = s->has_force_cmd[j] = 0;
    switch (mode) {
    case B2O_MODE_POSITION: s->has_pos_target[j] = 1; s->pos_target[j] = s->q[j]; break;
    case B2O_MODE_VELOCITY:
    case B2O_MODE_VELOCITY_FOLLOWER_DART: s->has_vel_target[j] = 1; s->vel_target[j] = s->dq[j]; break;
    default: s->has_force_cmd[j] = 1; s->force_cmd[j] = 0.0; break;
    }
    b2o_pid_reset(&s->pid[j]);
    return 1;
}
int b2o_sim_control_mode(const b2o_sim* s, int j) { return s->mode[j]; }

/* Joint.cpp:479-525: limits looser than the effort limit are replaced by +-effort */
int b2o_sim_set_pid(b2o_sim* s, int j, double p, double i, double d, double i_max, double i_min,
                    double cmd_max, double cmd_min, double cmd_offset)
{
    double fmax = s->model.effort[j];
    if (cmd_min < -fmax || cmd_max > fmax) { cmd_min = -fmax; cmd_max = fmax; }
    b2o_pid_init(&s->pid[j], p, i, d, i_max, i_min, cmd_max, cmd_min, cmd_offset);
    return 1;
}
int b2o_sim_set_controller_period(b2o_sim* s, double period)
{
    if (period <= 0) return 0;                              /* Model.cpp:589-602 */
    s->period_ns = to_ns(period);
    return 1;
}
/* Joint.cpp:774-815 */
int b2o_sim_set_force_target(b2o_sim* s, int j, double f)
{
    int md = s->mode[j];
    if (!(md == B2O_MODE_FORCE || md == B2O_MODE_POSITION || md == B2O_MODE_POSITION_INTERPOLATED ||
          md == B2O_MODE_VELOCITY))
        return 0;
    s->has_force_cmd[j] = 1;
    s->force_cmd[j] = f;
    return 1;
}
/* Joint.cpp:683-729 */
int b2o_sim_set_position_target(b2o_sim* s, int j, double v)
{
    int md = s->mode[j];
    if (!(md == B2O_MODE_POSITION || md == B2O_MODE_POSITION_INTERPOLATED || md == B2O_MODE_IDLE ||
          md == B2O_MODE_FORCE))
        return 0;
    s->has_pos_target[j] = 1;
    s->pos_target[j] = v;
    return 1;
}
/* Joint.cpp:731-772 */
int b2o_sim_set_velocity_target(b2o_sim* s, int j, double v)
{
    int md = s->mode[j];
    if (!(md == B2O_MODE_POSITION_INTERPOLATED || md == B2O_MODE_VELOCITY ||
          md == B2O_MODE_VELOCITY_FOLLOWER_DART || md == B2O_MODE_IDLE || md == B2O_MODE_FORCE))
        return 0;
    s->has_vel_target[j] = 1;
    s->vel_target[j] = v;
    return 1;
}
/* Joint.cpp:731-772 */
int b2o_sim_set_acceleration_target(b2o_sim* s, int j, double v)
{
    int md = s->mode[j];
    if (!(md == B2O_MODE_POSITION_INTERPOLATED || md == B2O_MODE_IDLE || md == B2O_MODE_FORCE)) return 0;
    s->has_acc_target[j] = 1;
    s->acc_target[j] = v;
    return 1;
}
/* ControllerRunner + ComputedTorqueFixedBase::initialize (ComputedTorqueFixedBase.cpp:125-203) */
int b2o_sim_load_computed_torque(b2o_sim* s, const double* kp, const double* kd, const double* gravity)
{
    for (int j = 0; j < s->model.nb; j++) {
        s->ct_kp[j] = kp[j];
        s->ct_kd[j] = kd[j];
        s->ct_tau[j] = 0;
        b2o_sim_set_control_mode(s, j, B2O_MODE_FORCE);
    }
    memcpy(s->ct_gravity, gravity, sizeof(v3));
    s->ct_loaded = 1;
    s->ct_refs_set = 0;
    s->ct_prev_update_ns = 0;
    return 1;
}
/* Link::applyWorldWrench (Link.cpp:496-527): `body`, `point` locate the link origin (body frame). */
int b2o_sim_apply_link_wrench(b2o_sim* s, int body, const double* point, const double* wrench, double duration)
{
    if (s->nwrench >= 8) return 0;
    s->wrench[s->nwrench].body = body;
    memcpy(s->wrench[s->nwrench].point, point, sizeof(v3));
    memcpy(s->wrench[s->nwrench].w, wrench, 6 * sizeof(double));
    s->wrench[s->nwrench].expiry_ns = s->time_ns + to_ns(duration);
    s->nwrench++;
    return 1;
}
/* Joint.cpp:132-180 */
int b2o_sim_reset_position(b2o_sim* s, int j, double v)
{
    s->pos_reset[j] = 1; s->pos_reset_v[j] = v;
    b2o_pid_reset(&s->pid[j]);
    return 1;
}
int b2o_sim_reset_velocity(b2o_sim* s, int j, double v)
{
    s->vel_reset[j] = 1; s->vel_reset_v[j] = v;
    b2o_pid_reset(&s->pid[j]);
    return 1;
}
double b2o_sim_position(const b2o_sim* s, int j) { return s->q[j]; }
double b2o_sim_velocity(const b2o_sim* s, int j) { return s->dq[j]; }
double b2o_sim_acceleration(const b2o_sim* s, int j) { return s->ddq[j]; }
double b2o_sim_force_target(const b2o_sim* s, int j, int* has)
{
    if (has) *has = s->has_force_cmd[j];
    return s->force_cmd[j];
}
double b2o_sim_position_target(const b2o_sim* s, int j, int* has)
{
    if (has) *has = s->has_pos_target[j];
    return s->pos_target[j];
}
double b2o_sim_velocity_target(const b2o_sim* s, int j, int* has)
{
    if (has) *has = s->has_vel_target[j];
    return s->vel_target[j];
}

int b2o_sim_attach_world(b2o_sim* s, const b2o_world* w, const b2o_robot_shapes* rs, const double* X0)
{
    s->world = *w;
    s->rshapes = *rs;
    memcpy(s->X, X0, sizeof(double) * 13 * w->nfree);
    memset(s->base_reset, 0, sizeof s->base_reset);
    s->ncontacts = 0;
    s->has_world = 1;
    return 1;
}
void b2o_sim_world_state(const b2o_sim* s, double* X) { memcpy(X, s->X, sizeof(double) * 13 * s->world.nfree); }
int b2o_sim_contacts(const b2o_sim* s, b2o_contact* out, int max_out)
{
    int n = s->ncontacts < max_out ? s->ncontacts : max_out;
    if (n > 0) memcpy(out, s->contacts, sizeof(b2o_contact) * n);
    return n < 0 ? 0 : n;
}
/* Model::resetBasePose / resetBaseWorldVelocity of free body `body` (Model.cpp:256-377): deferred to the next run */
int b2o_sim_reset_base(b2o_sim* s, int body, int velocity, const double* values)
{
    if (!s->has_world || body < 0 || body >= s->world.nfree) return 0;
    if (velocity) { memcpy(s->X_reset + 13 * body + 7, values, 6 * sizeof(double)); s->base_reset[body] |= 2; }
    else { memcpy(s->X_reset + 13 * body, values, 7 * sizeof(double)); s->base_reset[body] |= 1; }
    return 1;
}

/* One simulator iteration = JointController::PreUpdate + Physics::Update. */
static void sim_iteration(b2o_sim* s, int paused)
{
    const b2o_model* m = &s->model;
    const int nb = m->nb;
    const double dt = (double)s->dt_ns / 1e9;
    if (!paused) s->time_ns += s->dt_ns; /* systems see the post-step time, Physics.cpp:656-666 */

    /* JointController::PreUpdate, JointController.cpp:114-287 */
    if (!paused && s->controller_loaded) {
        double elapsed = (double)(s->time_ns - s->prev_update_ns) / 1e9;
        double period = s->period_ns == INT64_MAX ? (double)INT64_MAX / 1e9 : (double)s->period_ns / 1e9;
        if (s->prev_update_ns == 0) elapsed = period;               /* :141-144 */
        int compute_new = elapsed >= period - DBL_EPSILON;          /* :153-156 */
        if (compute_new) s->prev_update_ns = s->time_ns;
        for (int j = 0; j < nb; j++) {
            if (s->mode[j] == B2O_MODE_POSITION || s->mode[j] == B2O_MODE_VELOCITY) {
                double cur = s->mode[j] == B2O_MODE_POSITION ? s->q[j] : s->dq[j];
                double ref = s->mode[j] == B2O_MODE_POSITION ? s->pos_target[j] : s->vel_target[j];
                double f = compute_new ? b2o_pid_update(&s->pid[j], cur - ref, dt) : s->pid[j].cmd;
                s->has_force_cmd[j] = 1;                            /* :316 setGeneralizedForceTarget */
                s->force_cmd[j] = f;
            } else if (s->mode[j] == B2O_MODE_VELOCITY_FOLLOWER_DART) {
                s->has_vel_cmd[j] = 1;                              /* :263-286 */
                s->vel_cmd[j] = s->vel_target[j];
            }
        }
    }

    /* ControllerRunner::PreUpdate (ControllerRunner.cpp:183-282) + ComputedTorqueFixedBase::step */
    if (!paused && s->ct_loaded) {
        double elapsed = (double)(s->time_ns - s->ct_prev_update_ns) / 1e9;
        double period = s->period_ns == INT64_MAX ? (double)INT64_MAX / 1e9 : (double)s->period_ns / 1e9;
        if (s->ct_prev_update_ns == 0) elapsed = period;
        if (elapsed >= period - DBL_EPSILON) {
            s->ct_prev_update_ns = s->time_ns;
            int refs = 1;
            for (int j = 0; j < nb; j++)
                if (!(s->has_pos_target[j] && s->has_vel_target[j] && s->has_acc_target[j])) refs = 0;
            if (refs) {   /* references read, state and M, h refreshed (updateStateFromModel) */
                b2o_model cm = *m;
                double M[B2O_MAXB * B2O_MAXB], h[B2O_MAXB], zero[B2O_MAXB] = {0}, acc[B2O_MAXB];
                memcpy(cm.gravity, s->ct_gravity, sizeof(v3));
                b2o_mass_matrix(&cm, s->q, M);
                b2o_inverse_dynamics(&cm, s->q, s->dq, zero, 1, h);
                for (int j = 0; j < nb; j++)
                    acc[j] = s->acc_target[j] - s->ct_kp[j] * (s->q[j] - s->pos_target[j])
                             - s->ct_kd[j] * (s->dq[j] - s->vel_target[j]);
                for (int i = 0; i < nb; i++) {
                    double t = h[i];
                    for (int j = 0; j < nb; j++) t += M[i * nb + j] * acc[j];
                    s->ct_tau[i] = t;
                }
                s->ct_refs_set = 1;
            }
        }
        if (s->ct_refs_set)
            for (int j = 0; j < nb; j++) { s->has_force_cmd[j] = 1; s->force_cmd[j] = s->ct_tau[j]; }
    }

    /* WorldPoseCmd / WorldVelocityCmd of the free bodies (Physics.cpp:1535-1590,1716-1753): consumed by any run */
    if (s->has_world)
        for (int i = 0; i < s->world.nfree; i++) {
            if (s->base_reset[i] & 1) memcpy(s->X + 13 * i, s->X_reset + 13 * i, 7 * sizeof(double));
            if (s->base_reset[i] & 2) memcpy(s->X + 13 * i + 7, s->X_reset + 13 * i + 7, 6 * sizeof(double));
            s->base_reset[i] = 0;
        }

    /* Physics::Impl::UpdatePhysics joint block, Physics.cpp:1313-1443 */
    double tau[B2O_MAXB] = {0}, servo_target[B2O_MAXB] = {0};
    int servo[B2O_MAXB] = {0};
    for (int j = 0; j < nb; j++) {
        if (s->vel_reset[j]) s->dq[j] = s->vel_reset_v[j];
        if (s->pos_reset[j]) s->q[j] = s->pos_reset_v[j];
        if (s->has_force_cmd[j]) {
            tau[j] = s->force_cmd[j];
        } else if (s->has_vel_cmd[j] && !s->vel_reset[j]) {        /* :1404-1412 */
            servo[j] = 1;
            servo_target[j] = s->vel_cmd[j];
        }
    }
    if (!paused) {
        /* link wrenches (Physics.cpp:1483-1532): equivalent joint forces J^T F, then drop the expired ones */
        for (int k = 0; k < s->nwrench; k++) {
            double J[6 * B2O_MAXB];
            b2o_point_jacobian(m, s->q, s->wrench[k].body, s->wrench[k].point, J);
            for (int j = 0; j < nb; j++)
                for (int a = 0; a < 6; a++) tau[j] += J[a * nb + j] * s->wrench[k].w[a];
        }
        int keep = 0;
        for (int k = 0; k < s->nwrench; k++)
            if (!(s->time_ns >= s->wrench[k].expiry_ns)) s->wrench[keep++] = s->wrench[k];
        s->nwrench = keep;
        if (s->has_world) {
            /* World::step with contacts: ABA, dq += ddq dt, one constraint solve for joint rows + contacts, q += dq dt */
            double acc[B2O_MAXB], before[B2O_MAXB];
            b2o_forward_dynamics(m, dt, s->q, s->dq, tau, acc);
            for (int j = 0; j < nb; j++) { s->dq[j] += acc[j] * dt; before[j] = s->dq[j]; }
            s->ncontacts = b2o_coupled_step(&s->world, m, &s->rshapes, s->q, s->dq, servo, servo_target, s->X,
                                            s->contacts, B2O_MAXCONTACTS);
            for (int j = 0; j < nb; j++) { s->ddq[j] = acc[j] + (s->dq[j] - before[j]) / dt; s->q[j] += s->dq[j] * dt; }
        } else {
            physics_step_ex(m, dt, s->q, s->dq, tau, servo, servo_target, s->ddq); /* :1824-1835 */
        }
    }
    /* UpdateSim, Physics.cpp:2227-2345: drop resets, zero one-shot commands, read back.
     * DART clears joint forces at the end of World::step, so the JointForce readback is the pending
     * command only after a paused run. */
    for (int j = 0; j < nb; j++) {
        s->tau_read[j] = paused ? tau[j] : 0.0;
        s->pos_reset[j] = s->vel_reset[j] = 0;
        s->force_cmd[j] = 0.0;
        s->vel_cmd[j] = 0.0;
    }
}

/* GazeboSimulator::run, GazeboSimulator.cpp:202-251 */
int b2o_sim_run(b2o_sim* s, int paused)
{
    int n = paused ? 1 : s->steps_per_run;
    for (int it = 0; it < n; it++) sim_iteration(s, paused);
    return 1;
}

/* ------------------------------------------------------------------------------------------- */
/* Philox4x32-10 (Salmon et al., SC'11): counter-based RNG used for on-device episode resets.    */
/* The reference draws resets from numpy's MT19937 per env (base/task.py:52-61); a counter-based */
/* generator keyed by (seed, env) is the batched equivalent, same distributions.                 */
/* ------------------------------------------------------------------------------------------- */
void b2o_philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4])
{
    uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
    uint32_t k0 = key_in[0], k1 = key_in[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void b2o_reset_uniforms(uint64_t seed, uint64_t env, uint64_t step, int n, double* u)
{
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    for (int blk = 0; 2 * blk < n; blk++) {
        uint32_t ctr[4] = {(uint32_t)step, (uint32_t)(step >> 32), (uint32_t)env,
                           ((uint32_t)(env >> 32) << 8) | (uint32_t)blk};
        uint32_t r[4];
        b2o_philox4x32_10(ctr, key, r);
        /* 53-bit uniform, the construction numpy's random_sample uses */
        u[2 * blk] = ((double)(r[0] >> 5) * 67108864.0 + (double)(r[1] >> 6)) / 9007199254740992.0;
        if (2 * blk + 1 < n)
            u[2 * blk + 1] =
                ((double)(r[2] >> 5) * 67108864.0 + (double)(r[3] >> 6)) / 9007199254740992.0;
    }
}

/* ------------------------------------------------------------------------------------------- */
/* Tasks (python/gym_ignition_environments/tasks, all four files)                                           */
/* ------------------------------------------------------------------------------------------- */
#define B2O_PI 3.141592653589793
static double deg2rad(double d) { return d * (B2O_PI / 180.0); } /* numpy: x * (pi/180) */

int b2o_task_nobs(int task) { return task == B2O_TASK_PENDULUM_SWINGUP ? 3 : 4; }
int b2o_task_nq(int task) { return task == B2O_TASK_PENDULUM_SWINGUP ? 1 : 2; }

double b2o_task_action_force(int task, double action, int* joint)
{
    if (joint) *joint = 0; /* pendulum: "pivot"; cartpole: "linear" is the first joint */
    if (task == B2O_TASK_CARTPOLE_DISCRETE_BALANCING)
        return action == 1.0 ? 20.0 : -20.0;             /* cartpole_discrete_balancing.py:67-77 */
    return action;
}

void b2o_task_sample_reset(int task, uint64_t seed, uint64_t env, uint64_t step, double* state)
{
    double u[4];
    b2o_reset_uniforms(seed, env, step, 4, u);
    b2o_task_reset_from_uniforms(task, u, state);
}

/* Task.reset_task with the RNG draws given explicitly (u in [0,1), in the task's draw order). */
void b2o_task_reset_from_uniforms(int task, const double* u, double* state)
{
    const double lo = -0.05, range = 0.05 - (-0.05);
    switch (task) {
    case B2O_TASK_PENDULUM_SWINGUP: {
        /* pendulum_swingup.py:118-127: (cos, sin, dq) = observation_space.sample() (float32 Box),
         * q = arctan2(sin, cos) evaluated in float32 */
        float c = (float)(-1.0 + 2.0 * u[0]);
        float s = (float)(-1.0 + 2.0 * u[1]);
        float w = (float)(-10.0 + 20.0 * u[2]);
        float qf = (float)atan2((double)s, (double)c);
        state[0] = (double)qf;
        state[1] = (double)w;
        break;
    }
    case B2O_TASK_CARTPOLE_DISCRETE_BALANCING:
    case B2O_TASK_CARTPOLE_CONTINUOUS_BALANCING: {
        /* cartpole_discrete_balancing.py:137: x, dx, q, dq = U(-0.05, 0.05, 4) */
        double x = lo + range * u[0], dx = lo + range * u[1];
        double q = lo + range * u[2], dq = lo + range * u[3];
        state[0] = x; state[1] = q; state[2] = dx; state[3] = dq;
        break;
    }
    case B2O_TASK_CARTPOLE_CONTINUOUS_SWINGUP: {
        /* cartpole_continuous_swingup.py:145-146 */
        double q = B2O_PI - deg2rad(-60.0 + (60.0 - (-60.0)) * u[0]);
        double x = lo + range * u[1], dx = lo + range * u[2], dq = lo + range * u[3];
        state[0] = x; state[1] = q; state[2] = dx; state[3] = dq;
        break;
    }
    default: break;
    }
}

/* gym.spaces.Box(dtype=float32).contains on a float64 vector: bounds are float32-rounded,
 * comparisons inclusive (gym 0.17 Box.contains) */
static int inside(double v, double high_f64)
{
    double h = (double)(float)high_f64;
    return (v >= -h) && (v <= h);
}

int b2o_task_evaluate(int task, const double* st, double tau_after, double* obs, double* reward)
{
    if (task == B2O_TASK_PENDULUM_SWINGUP) {
        double q = st[0], dq = st[1];
        obs[0] = cos(q); obs[1] = sin(q); obs[2] = dq;            /* pendulum_swingup.py:58-71 */
        int done = !(inside(obs[0], 1.0) && inside(obs[1], 1.0) && inside(obs[2], 10.0));
        double cost = done ? 100.0 : 0.0;                         /* :73-90 */
        cost += (q * q) + 0.1 * (dq * dq) + 0.001 * (tau_after * tau_after);
        *reward = -cost;
        return done;
    }
    const double x = st[0], q = st[1], dx = st[2], dq = st[3];
    obs[0] = x; obs[1] = dx; obs[2] = q; obs[3] = dq;             /* cartpole_*.py:79-92 */
    const double x_thr = 2.4, dx_thr = 20.0;
    const double dq_thr = deg2rad(3 * 360);
    if (task == B2O_TASK_CARTPOLE_CONTINUOUS_SWINGUP) {
        const double q_thr = deg2rad(5 * 360);
        int done = !(inside(x, x_thr) && inside(dx, dx_thr) && inside(q, q_thr) && inside(dq, dq_thr));
        double r = (cos(q) + 1) / 2;                              /* :96-117 */
        r -= 0.1 * (dx * dx);
        r -= 10.0 * (double)(x >= 0.8 * x_thr);
        *reward = r;
        return done;
    }
    const double q_thr = deg2rad(12);
    int done = !(inside(x, x_thr) && inside(dx, dx_thr) && inside(q, q_thr) && inside(dq, dq_thr));
    double r = done ? 0.0 : 1.0;                                  /* cartpole_discrete_balancing.py:94-109 */
    double edge = task == B2O_TASK_CARTPOLE_DISCRETE_BALANCING ? 0.9 * x_thr : x_thr;
    r = r - 0.10 * fabs(x) - 0.10 * fabs(dx) - 10.0 * (double)(x >= edge);
    *reward = r;
    return done;
}

void b2o_rollout(const b2o_model* m, int task, double dt, int steps_per_run, int max_episode_steps, uint64_t seed,
                 uint64_t env_offset, uint64_t first_step, int n_envs, int T, const double* actions,
                 double* state, int32_t* elapsed, double* obs, double* reward, uint8_t* done)
{
    const int nq = m->nb, nobs = b2o_task_nobs(task);
    for (int t = 0; t < T; t++) {
        for (int e = 0; e < n_envs; e++) {
            double* st = state + (size_t)e * 2 * nq;
            double tau[B2O_MAXB] = {0}, o[8], r;
            int joint;
            /* Task.set_action -> Joint.set_generalized_force_target (one-shot command) */
            double f = b2o_task_action_force(task, actions[(size_t)t * n_envs + e], &joint);
            tau[joint] = f;
            /* gazebo.run(): no PID joints in these tasks; Physics applies the force and steps. With
             * steps_per_run > 1 the one-shot command acts on the first iteration only. */
            for (int it = 0; it < steps_per_run; it++) {
                b2o_physics_step(m, dt, st, st + nq, tau, NULL);
                tau[joint] = 0.0;
            }
            /* Physics zeroes JointForceCmd after the step -> the task reads tau = 0 */
            int d = b2o_task_evaluate(task, st, 0.0, o, &r);
            elapsed[e] += 1;
            if (elapsed[e] >= max_episode_steps) d = 1;           /* gym TimeLimit */
            size_t idx = (size_t)t * n_envs + e;
            if (obs) memcpy(obs + idx * nobs, o, sizeof(double) * nobs);
            if (reward) reward[idx] = r;
            if (done) done[idx] = (uint8_t)d;
            if (d) {
                b2o_task_sample_reset(task, seed, env_offset + e, first_step + t, st);
                elapsed[e] = 0;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------- */
/* Free rigid bodies with plane / box contacts.                                */
/* Restates, for free bodies, what DART's World::step does behind Physics.cpp:1824-1835 and what   */
/* Physics.cpp:2351-2540 reads back as contacts: velocity update, contact points, contact rows     */
/* (normal + 2 friction directions, mu = min of the two surfaces), error-reduction velocity        */
/* depth * ERP / dt capped at max_erv, then pose integration. DART solves the LCP with Dantzig on  */
/* ODE/FCL contact points; this oracle uses box corners / sphere points and projected              */
/* Gauss-Seidel with a fixed iteration count, so only trajectory statistics are comparable         */
/* with the reference (BASELINE.json: "contact configs are compared on trajectory statistics").    */
/* ------------------------------------------------------------------------------------------- */
typedef struct {
    double xc[3], vc[3], w[3], R[9], Iinv[9], inv_mass;
} body_work;

typedef struct {
    int a, b, shape_a;
    double pos[3], n[3], depth, mu, t1[3], t2[3], kn, kt1, kt2, ln, lt1, lt2, bias;
} contact_row;

static void quat_R(const double* q, double* R)
{
    double w = q[0], x = q[1], y = q[2], z = q[3];
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w); R[2] = 2 * (x * z + y * w);
    R[3] = 2 * (x * y + z * w); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
    R[6] = 2 * (x * z - y * w); R[7] = 2 * (y * z + x * w); R[8] = 1 - 2 * (x * x + y * y);
}
static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void rot_inertia(const double* R, const double* I, double* out)
{
    double t[9], Rt[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) Rt[3 * i + j] = R[3 * j + i];
    m3m(R, I, t);
    m3m(t, Rt, out);
}
static void push_contact(contact_row* cs, int* nc, int max, int a, int sa, int b, const double* pos,
                         const double* n, double depth, double mu)
{
    if (*nc >= max) return;
    contact_row* c = &cs[(*nc)++];
    memset(c, 0, sizeof *c);
    c->a = a; c->b = b; c->shape_a = sa; c->depth = depth; c->mu = mu;
    memcpy(c->pos, pos, sizeof(v3)); memcpy(c->n, n, sizeof(v3));
}
static void box_contacts(contact_row* cs, int* nc, int max, int a, int sa, int b, const double* Ra,
                         const double* pa, const double* ha, const b2o_shape* sb, const double* Rb,
                         const double* pb, double mu)
{
    for (int k = 0; k < 8; k++) {
        double cl[3] = {(k & 1) ? ha[0] : -ha[0], (k & 2) ? ha[1] : -ha[1], (k & 4) ? ha[2] : -ha[2]};
        double x[3], d[3];
        m3v(Ra, cl, x);
        for (int i = 0; i < 3; i++) { x[i] += pa[i]; d[i] = x[i] - pb[i]; }
        if (sb->type == B2O_SHAPE_PLANE) {
            double dist = dot3(sb->size, d);
            if (dist <= 0) push_contact(cs, nc, max, a, sa, b, x, sb->size, -dist, mu);
        }
    }
    if (sb->type != B2O_SHAPE_BOX) return;
    double dc[3] = {pa[0] - pb[0], pa[1] - pb[1], pa[2] - pb[2]}, cb[3];
    m3tv(Rb, dc, cb);
    int axis = 0;
    for (int j = 1; j < 3; j++)
        if (fabs(cb[j]) - sb->size[j] > fabs(cb[axis]) - sb->size[axis]) axis = j;
    double sgn = cb[axis] >= 0 ? 1.0 : -1.0;
    for (int k = 0; k < 8; k++) {
        double cl[3] = {(k & 1) ? ha[0] : -ha[0], (k & 2) ? ha[1] : -ha[1], (k & 4) ? ha[2] : -ha[2]};
        double x[3], d[3], xl[3];
        m3v(Ra, cl, x);
        for (int i = 0; i < 3; i++) { x[i] += pa[i]; d[i] = x[i] - pb[i]; }
        m3tv(Rb, d, xl);
        int inside = 1;
        for (int j = 0; j < 3; j++)
            if (j != axis && fabs(xl[j]) > sb->size[j] + 1e-6) inside = 0;
        double depth = sb->size[axis] - sgn * xl[axis];
        if (inside && depth >= 0 && depth <= 2 * sb->size[axis]) {
            double nl[3] = {0, 0, 0}, n[3];
            nl[axis] = sgn;
            m3v(Rb, nl, n);
            push_contact(cs, nc, max, a, sa, b, x, n, depth, mu);
        }
    }
}
static void sphere_contacts(contact_row* cs, int* nc, int max, int a, int sa, int b, const double* ca, double r,
                            const b2o_shape* sb, const double* Rb, const double* pb, double mu)
{
    double d[3] = {ca[0] - pb[0], ca[1] - pb[1], ca[2] - pb[2]};
    if (sb->type == B2O_SHAPE_PLANE) {
        double dist = dot3(sb->size, d) - r;
        if (dist <= 0) {
            double x[3] = {ca[0] - r * sb->size[0], ca[1] - r * sb->size[1], ca[2] - r * sb->size[2]};
            push_contact(cs, nc, max, a, sa, b, x, sb->size, -dist, mu);
        }
    } else if (sb->type == B2O_SHAPE_BOX) {
        double cl[3], q[3], dl[3];
        int inside = 1;
        m3tv(Rb, d, cl);
        for (int j = 0; j < 3; j++) {
            q[j] = cl[j] < -sb->size[j] ? -sb->size[j] : (cl[j] > sb->size[j] ? sb->size[j] : cl[j]);
            if (q[j] != cl[j]) inside = 0;
            dl[j] = cl[j] - q[j];
        }
        if (inside) return;
        double dist = sqrt(dot3(dl, dl));
        if (dist <= r) {
            double nl[3] = {dl[0] / dist, dl[1] / dist, dl[2] / dist}, n[3], x[3];
            m3v(Rb, nl, n);
            m3v(Rb, q, x);
            for (int i = 0; i < 3; i++) x[i] += pb[i];
            push_contact(cs, nc, max, a, sa, b, x, n, r - dist, mu);
        }
    }
}
static void rel_velocity(const body_work* bw, const contact_row* c, double* v)
{
    double r[3], t[3];
    for (int i = 0; i < 3; i++) r[i] = c->pos[i] - bw[c->a].xc[i];
    cross3(bw[c->a].w, r, t);
    for (int i = 0; i < 3; i++) v[i] = bw[c->a].vc[i] + t[i];
    if (c->b >= 0) {
        for (int i = 0; i < 3; i++) r[i] = c->pos[i] - bw[c->b].xc[i];
        cross3(bw[c->b].w, r, t);
        for (int i = 0; i < 3; i++) v[i] -= bw[c->b].vc[i] + t[i];
    }
}
static double eff_mass(const body_work* bw, const contact_row* c, const double* d)
{
    double k = 0;
    for (int side = 0; side < 2; side++) {
        int bi = side == 0 ? c->a : c->b;
        if (bi < 0) continue;
        double r[3], rxd[3], t[3], u[3];
        for (int i = 0; i < 3; i++) r[i] = c->pos[i] - bw[bi].xc[i];
        cross3(r, d, rxd);
        m3v(bw[bi].Iinv, rxd, t);
        cross3(t, r, u);
        k += bw[bi].inv_mass + dot3(d, u);
    }
    return k;
}
static void impulse(body_work* bw, const contact_row* c, const double* dir, double mag)
{
    for (int side = 0; side < 2; side++) {
        int bi = side == 0 ? c->a : c->b;
        if (bi < 0) continue;
        double s = side == 0 ? mag : -mag, P[3] = {s * dir[0], s * dir[1], s * dir[2]}, r[3], rxP[3], dw[3];
        for (int i = 0; i < 3; i++) r[i] = c->pos[i] - bw[bi].xc[i];
        cross3(r, P, rxP);
        m3v(bw[bi].Iinv, rxP, dw);
        for (int i = 0; i < 3; i++) { bw[bi].vc[i] += bw[bi].inv_mass * P[i]; bw[bi].w[i] += dw[i]; }
    }
}

/* Loads the free bodies and applies the unconstrained velocity update (gravity, gyroscopic torque). */
static void bodies_begin(const b2o_world* W, const double* X, body_work* bw)
{
    const double dt = W->dt;
    for (int i = 0; i < W->nfree; i++) {
        const b2o_free_body* fb = &W->body[i];
        const double* x = X + 13 * i;
        body_work* b = &bw[i];
        double rc[3], t[3], Iw[9], Iinv_b[9], Iw_w[3], g1[3], g2[3];
        quat_R(x + 3, b->R);
        m3v(b->R, fb->com, rc);
        cross3(x + 10, rc, t);
        for (int k = 0; k < 3; k++) { b->xc[k] = x[k] + rc[k]; b->w[k] = x[10 + k]; b->vc[k] = x[7 + k] + t[k]; }
        b->inv_mass = 1.0 / fb->mass;
        /* inverse of the 3x3 body inertia by cofactors */
        {
            const double* I = fb->Ic;
            double det = I[0] * (I[4] * I[8] - I[5] * I[7]) - I[1] * (I[3] * I[8] - I[5] * I[6]) + I[2] * (I[3] * I[7] - I[4] * I[6]);
            Iinv_b[0] = (I[4] * I[8] - I[5] * I[7]) / det; Iinv_b[1] = (I[2] * I[7] - I[1] * I[8]) / det; Iinv_b[2] = (I[1] * I[5] - I[2] * I[4]) / det;
            Iinv_b[3] = (I[5] * I[6] - I[3] * I[8]) / det; Iinv_b[4] = (I[0] * I[8] - I[2] * I[6]) / det; Iinv_b[5] = (I[2] * I[3] - I[0] * I[5]) / det;
            Iinv_b[6] = (I[3] * I[7] - I[4] * I[6]) / det; Iinv_b[7] = (I[1] * I[6] - I[0] * I[7]) / det; Iinv_b[8] = (I[0] * I[4] - I[1] * I[3]) / det;
        }
        rot_inertia(b->R, Iinv_b, b->Iinv);
        rot_inertia(b->R, fb->Ic, Iw);
        m3v(Iw, b->w, Iw_w);
        cross3(b->w, Iw_w, g1);
        m3v(b->Iinv, g1, g2);
        for (int k = 0; k < 3; k++) { b->vc[k] += dt * W->g[k]; b->w[k] -= dt * g2[k]; }
        {   /* external wrench: force at the root link origin -> force at the COM + moment */
            const double* f = W->ext[i];
            double arm[3], mo[3], dw[3];
            for (int k = 0; k < 3; k++) arm[k] = x[k] - b->xc[k];
            cross3(arm, f, mo);
            for (int k = 0; k < 3; k++) mo[k] += f[3 + k];
            m3v(b->Iinv, mo, dw);
            for (int k = 0; k < 3; k++) { b->vc[k] += dt * b->inv_mass * f[k]; b->w[k] += dt * dw[k]; }
        }
    }
}

/* World pose of shape `sh` of a free body. */
static void free_shape_pose(const b2o_free_body* fb, const body_work* b, const b2o_shape* sh, double* Rs, double* ps)
{
    double off[3], t[3];
    m3m(b->R, sh->R, Rs);
    for (int k = 0; k < 3; k++) off[k] = sh->p[k] - fb->com[k];
    m3v(b->R, off, t);
    for (int k = 0; k < 3; k++) ps[k] = b->xc[k] + t[k];
}

/* Contact points of the free bodies against static shapes and against each other. */
static void free_contacts(const b2o_world* W, const body_work* bw, contact_row* cs, int* nc)
{
    for (int i = 0; i < W->nfree; i++) {
        const b2o_free_body* fb = &W->body[i];
        for (int s = 0; s < fb->nshapes; s++) {
            const b2o_shape* sa = &fb->shape[s];
            double Ra[9], pa[3];
            free_shape_pose(fb, &bw[i], sa, Ra, pa);
            for (int k = 0; k < W->nstatic; k++) {
                const b2o_shape* sb = &W->stat[k];
                double mu = sa->mu < sb->mu ? sa->mu : sb->mu;
                if (sa->type == B2O_SHAPE_BOX) box_contacts(cs, nc, B2O_MAXCONTACTS, i, s, -1 - k, Ra, pa, sa->size, sb, sb->R, sb->p, mu);
                else if (sa->type == B2O_SHAPE_SPHERE) sphere_contacts(cs, nc, B2O_MAXCONTACTS, i, s, -1 - k, pa, sa->size[0], sb, sb->R, sb->p, mu);
            }
            for (int j = 0; j < W->nfree; j++) {
                if (j == i) continue;
                const b2o_free_body* fj = &W->body[j];
                for (int u = 0; u < fj->nshapes; u++) {
                    const b2o_shape* sb = &fj->shape[u];
                    if (sb->type != B2O_SHAPE_BOX) continue;
                    double Rb[9], pb[3];
                    free_shape_pose(fj, &bw[j], sb, Rb, pb);
                    double mu = sa->mu < sb->mu ? sa->mu : sb->mu;
                    if (sa->type == B2O_SHAPE_BOX) box_contacts(cs, nc, B2O_MAXCONTACTS, i, s, j, Ra, pa, sa->size, sb, Rb, pb, mu);
                    else if (sa->type == B2O_SHAPE_SPHERE) sphere_contacts(cs, nc, B2O_MAXCONTACTS, i, s, j, pa, sa->size[0], sb, Rb, pb, mu);
                }
            }
        }
    }
}

/* Tangent directions of a contact normal. */
static void contact_tangents(contact_row* c)
{
    double seed[3] = {fabs(c->n[0]) < 0.9 ? 1.0 : 0.0, fabs(c->n[0]) < 0.9 ? 0.0 : 1.0, 0.0}, nrm;
    cross3(c->n, seed, c->t1);
    nrm = sqrt(dot3(c->t1, c->t1));
    for (int i = 0; i < 3; i++) c->t1[i] /= nrm;
    cross3(c->n, c->t1, c->t2);
}

/* Integrates the poses of the free bodies (rotation by the exponential map). */
static void bodies_end(const b2o_world* W, double* X, body_work* bw)
{
    const double dt = W->dt;
    for (int i = 0; i < W->nfree; i++) {
        const b2o_free_body* fb = &W->body[i];
        double* x = X + 13 * i;
        body_work* b = &bw[i];
        double wn = sqrt(dot3(b->w, b->w)), q[4] = {x[3], x[4], x[5], x[6]};
        if (wn > 0) {
            double s = sin(0.5 * wn * dt), c = cos(0.5 * wn * dt);
            double d[4] = {c, s / wn * b->w[0], s / wn * b->w[1], s / wn * b->w[2]};
            double r0 = d[0] * q[0] - d[1] * q[1] - d[2] * q[2] - d[3] * q[3];
            double r1 = d[0] * q[1] + d[1] * q[0] + d[2] * q[3] - d[3] * q[2];
            double r2 = d[0] * q[2] - d[1] * q[3] + d[2] * q[0] + d[3] * q[1];
            double r3 = d[0] * q[3] + d[1] * q[2] - d[2] * q[1] + d[3] * q[0];
            double nrm = 1.0 / sqrt(r0 * r0 + r1 * r1 + r2 * r2 + r3 * r3);
            q[0] = r0 * nrm; q[1] = r1 * nrm; q[2] = r2 * nrm; q[3] = r3 * nrm;
        }
        double Rn[9], rc[3], t[3];
        for (int k = 0; k < 3; k++) b->xc[k] += dt * b->vc[k];
        quat_R(q, Rn);
        m3v(Rn, fb->com, rc);
        cross3(b->w, rc, t);
        for (int k = 0; k < 3; k++) { x[k] = b->xc[k] - rc[k]; x[7 + k] = b->vc[k] - t[k]; x[10 + k] = b->w[k]; }
        for (int k = 0; k < 4; k++) x[3 + k] = q[k];
    }
}

static void export_contacts(const contact_row* cs, int nc, double dt, b2o_contact* out, int max_out)
{
    for (int k = 0; k < nc && k < max_out; k++) {
        out[k].a = cs[k].a; out[k].b = cs[k].b; out[k].depth = cs[k].depth;
        for (int i = 0; i < 3; i++) {
            out[k].pos[i] = cs[k].pos[i]; out[k].n[i] = cs[k].n[i];
            out[k].force[i] = (cs[k].ln * cs[k].n[i] + cs[k].lt1 * cs[k].t1[i] + cs[k].lt2 * cs[k].t2[i]) / dt;
        }
    }
}

int b2o_world_step(const b2o_world* W, double* X, b2o_contact* out, int max_out)
{
    body_work bw[B2O_MAXFREE];
    contact_row cs[B2O_MAXCONTACTS];
    const double dt = W->dt;
    int nc = 0;
    bodies_begin(W, X, bw);
    free_contacts(W, bw, cs, &nc);
    for (int k = 0; k < nc; k++) {
        contact_row* c = &cs[k];
        contact_tangents(c);
        c->kn = eff_mass(bw, c, c->n); c->kt1 = eff_mass(bw, c, c->t1); c->kt2 = eff_mass(bw, c, c->t2);
        double erv = c->depth * W->erp / dt;
        c->bias = erv > W->max_erv ? W->max_erv : erv;
    }
    for (int it = 0; it < W->iterations; it++) {
        for (int k = 0; k < nc; k++) {
            contact_row* c = &cs[k];
            double rel[3], l;
            rel_velocity(bw, c, rel);
            l = c->ln + (c->bias - dot3(c->n, rel)) / c->kn;
            if (l < 0) l = 0;
            impulse(bw, c, c->n, l - c->ln);
            c->ln = l;
            double lim = c->mu * c->ln;
            rel_velocity(bw, c, rel);
            l = clampd(c->lt1 - dot3(c->t1, rel) / c->kt1, -lim, lim);
            impulse(bw, c, c->t1, l - c->lt1);
            c->lt1 = l;
            rel_velocity(bw, c, rel);
            l = clampd(c->lt2 - dot3(c->t2, rel) / c->kt2, -lim, lim);
            impulse(bw, c, c->t2, l - c->lt2);
            c->lt2 = l;
        }
    }
    bodies_end(W, X, bw);
    export_contacts(cs, nc, dt, out, max_out);
    return nc;
}

/* ------------------------------------------------------------------------------------------- */
/* Coupled world: one articulated model whose moving links carry collision shapes (the Panda's   */
/* fingers) + free bodies + static shapes (examples/panda_pick_and_place.py). DART's constraint   */
/* stage sees the skeleton's joint rows (limits, Coulomb friction, servo) and every contact in one */
/* LCP. Restated here in its textbook dense form: generalized velocity v = [dq, (vc, w) per free   */
/* body], one Jacobian row per constraint, Y = M^-1 J^T with M = blockdiag(M(q), m I, I_w), and     */
/* projected Gauss-Seidel sweeps in the order joint rows, then contacts (normal, t1, t2).         */
/* ------------------------------------------------------------------------------------------- */
#define B2O_ROBOT_SIDE (-1000)
#define B2O_NV (B2O_MAXB + 6 * B2O_MAXFREE)
#define B2O_MAXJOINTROWS 16
#define B2O_MAXROWS (B2O_MAXJOINTROWS + 3 * B2O_MAXCONTACTS)
#define B2O_MAXROBOTCONTACTS 16

int b2o_coupled_step(const b2o_world* W, const b2o_model* m, const b2o_robot_shapes* rs, const double* q,
                     double* dq, const int* servo, const double* servo_target, double* X, b2o_contact* out,
                     int max_out)
{
    body_work bw[B2O_MAXFREE];
    contact_row cs[B2O_MAXCONTACTS];
    static _Thread_local double J[B2O_MAXROWS][B2O_NV], Y[B2O_MAXROWS][B2O_NV];
    double v[B2O_NV], kdiag[B2O_MAXROWS], lam[B2O_MAXROWS], lo[B2O_MAXROWS], hi[B2O_MAXROWS], target[B2O_MAXROWS];
    double Rw[B2O_MAXB * 9], pw[B2O_MAXB * 3], M[B2O_MAXB * B2O_MAXB], Minv[B2O_MAXB * B2O_MAXB];
    const int nb = m->nb, nv = nb + 6 * W->nfree;
    const double dt = W->dt;
    int nc = 0;

    b2o_forward_kinematics(m, q, Rw, pw);
    bodies_begin(W, X, bw);
    free_contacts(W, bw, cs, &nc);
    /* link shapes of the articulated model against static shapes and free bodies (both directions for boxes) */
    for (int r = 0; r < rs->nrobot; r++) {
        const b2o_shape* sr = &rs->rshape[r];
        const int body = rs->rbody[r], side = B2O_ROBOT_SIDE - r;
        double Rr[9], pr[3], t[3];
        m3m(Rw + 9 * body, sr->R, Rr);
        m3v(Rw + 9 * body, sr->p, t);
        for (int k = 0; k < 3; k++) pr[k] = pw[3 * body + k] + t[k];
        for (int k = 0; k < W->nstatic; k++) {
            const b2o_shape* sb = &W->stat[k];
            double mu = sr->mu < sb->mu ? sr->mu : sb->mu;
            if (sr->type == B2O_SHAPE_BOX) box_contacts(cs, &nc, B2O_MAXCONTACTS, side, r, -1 - k, Rr, pr, sr->size, sb, sb->R, sb->p, mu);
            else if (sr->type == B2O_SHAPE_SPHERE) sphere_contacts(cs, &nc, B2O_MAXCONTACTS, side, r, -1 - k, pr, sr->size[0], sb, sb->R, sb->p, mu);
        }
        for (int j = 0; j < W->nfree; j++) {
            const b2o_free_body* fj = &W->body[j];
            for (int u = 0; u < fj->nshapes; u++) {
                const b2o_shape* sb = &fj->shape[u];
                double Rb[9], pb[3];
                free_shape_pose(fj, &bw[j], sb, Rb, pb);
                double mu = sr->mu < sb->mu ? sr->mu : sb->mu;
                if (sb->type == B2O_SHAPE_BOX) {
                    if (sr->type == B2O_SHAPE_BOX) box_contacts(cs, &nc, B2O_MAXCONTACTS, side, r, j, Rr, pr, sr->size, sb, Rb, pb, mu);
                    else if (sr->type == B2O_SHAPE_SPHERE) sphere_contacts(cs, &nc, B2O_MAXCONTACTS, side, r, j, pr, sr->size[0], sb, Rb, pb, mu);
                }
                if (sr->type == B2O_SHAPE_BOX) {
                    if (sb->type == B2O_SHAPE_BOX) box_contacts(cs, &nc, B2O_MAXCONTACTS, j, u, side, Rb, pb, sb->size, sr, Rr, pr, mu);
                    else if (sb->type == B2O_SHAPE_SPHERE) sphere_contacts(cs, &nc, B2O_MAXCONTACTS, j, u, side, pb, sb->size[0], sr, Rr, pr, mu);
                }
            }
        }
    }
    /* at most B2O_MAXROBOTCONTACTS contacts may involve the articulated model: later ones are dropped */
    {
        int keep = 0, nrc = 0;
        for (int k = 0; k < nc; k++) {
            if (cs[k].a <= B2O_ROBOT_SIDE || cs[k].b <= B2O_ROBOT_SIDE) {
                if (nrc >= B2O_MAXROBOTCONTACTS) continue;
                nrc++;
            }
            cs[keep++] = cs[k];
        }
        nc = keep;
    }
    /* generalized velocity and block-diagonal inverse mass */
    for (int j = 0; j < nb; j++) v[j] = dq[j];
    for (int i = 0; i < W->nfree; i++)
        for (int k = 0; k < 3; k++) { v[nb + 6 * i + k] = bw[i].vc[k]; v[nb + 6 * i + 3 + k] = bw[i].w[k]; }
    b2o_mass_matrix(m, q, M);
    if (!cholesky_solve_inplace(nb, M, Minv)) return -1;
    /* rows */
    int nr = 0;
    for (int j = 0; j < nb; j++) {
        for (int pass = 0; pass < 3; pass++) {
            double rlo = 0, rhi = 0, rt = 0;
            int active = 0;
            if (servo && servo[j]) {
                if (pass == 0) { active = 1; rt = servo_target[j]; rlo = -m->effort[j] * dt; rhi = m->effort[j] * dt; }
            } else if (pass == 0) {
                if (m->friction[j] != 0.0) { active = 1; rlo = -m->friction[j] * dt; rhi = m->friction[j] * dt; }
            } else if (pass == 1) {
                if (q[j] <= m->lower[j]) { active = 1; rlo = 0; rhi = INFINITY; }
            } else {
                if (q[j] >= m->upper[j]) { active = 1; rlo = -INFINITY; rhi = 0; }
            }
            if (!active || nr >= B2O_MAXJOINTROWS) continue;
            memset(J[nr], 0, sizeof J[nr]);
            J[nr][j] = 1.0;
            lo[nr] = rlo; hi[nr] = rhi; target[nr] = rt;
            nr++;
        }
    }
    const int njr = nr;
    for (int k = 0; k < nc; k++) {
        contact_row* c = &cs[k];
        contact_tangents(c);
        double erv = c->depth * W->erp / dt;
        c->bias = erv > W->max_erv ? W->max_erv : erv;
        const double* dir[3] = {c->n, c->t1, c->t2};
        for (int d = 0; d < 3; d++) {
            double* row = J[nr + d];
            memset(row, 0, sizeof J[0]);
            for (int sidx = 0; sidx < 2; sidx++) {
                const int sd = sidx == 0 ? c->a : c->b;
                const double sign = sidx == 0 ? 1.0 : -1.0;
                if (sd >= 0) {                                   /* free body: d . (vc + w x r) */
                    double r[3], rxd[3];
                    for (int i = 0; i < 3; i++) r[i] = c->pos[i] - bw[sd].xc[i];
                    cross3(r, dir[d], rxd);
                    for (int i = 0; i < 3; i++) { row[nb + 6 * sd + i] += sign * dir[d][i]; row[nb + 6 * sd + 3 + i] += sign * rxd[i]; }
                } else if (sd <= B2O_ROBOT_SIDE) {               /* link of the articulated model: d . J_lin(pos) dq */
                    const int body = rs->rbody[B2O_ROBOT_SIDE - sd];
                    for (int i = body; i >= 0; i = m->parent[i]) {
                        double aw[3], lin[3], r[3];
                        m3v(Rw + 9 * i, m->axis[i], aw);
                        if (m->jtype[i] == B2O_REVOLUTE) {
                            for (int e = 0; e < 3; e++) r[e] = c->pos[e] - pw[3 * i + e];
                            cross3(aw, r, lin);
                        } else {
                            memcpy(lin, aw, sizeof lin);
                        }
                        row[i] += sign * dot3(dir[d], lin);
                    }
                }
            }
        }
        nr += 3;
    }
    for (int a = 0; a < nr; a++) {
        memset(Y[a], 0, sizeof Y[a]);
        for (int i = 0; i < nb; i++) {
            double y = 0;
            for (int j = 0; j < nb; j++) y += Minv[i * nb + j] * J[a][j];
            Y[a][i] = y;
        }
        for (int i = 0; i < W->nfree; i++) {
            const double* jl = &J[a][nb + 6 * i];
            double t[3];
            m3v(bw[i].Iinv, jl + 3, t);
            for (int k = 0; k < 3; k++) { Y[a][nb + 6 * i + k] = bw[i].inv_mass * jl[k]; Y[a][nb + 6 * i + 3 + k] = t[k]; }
        }
        double kd = 0;
        for (int i = 0; i < nv; i++) kd += J[a][i] * Y[a][i];
        kdiag[a] = kd;
        lam[a] = 0;
    }
    for (int it = 0; it < W->iterations; it++) {
        for (int a = 0; a < nr; a++) {
            double w = 0, l, l_lo, l_hi, bias = 0;
            for (int i = 0; i < nv; i++) w += J[a][i] * v[i];
            if (a < njr) {
                w -= target[a]; l_lo = lo[a]; l_hi = hi[a];
            } else {
                const contact_row* c = &cs[(a - njr) / 3];
                const int d = (a - njr) % 3;
                if (d == 0) { bias = c->bias; l_lo = 0; l_hi = INFINITY; }
                else { const double lim = c->mu * lam[a - d]; l_lo = -lim; l_hi = lim; }
            }
            l = clampd(lam[a] + (bias - w) / kdiag[a], l_lo, l_hi);
            const double dl = l - lam[a];
            lam[a] = l;
            for (int i = 0; i < nv; i++) v[i] += Y[a][i] * dl;
        }
    }
    for (int j = 0; j < nb; j++) dq[j] = v[j];
    for (int i = 0; i < W->nfree; i++)
        for (int k = 0; k < 3; k++) { bw[i].vc[k] = v[nb + 6 * i + k]; bw[i].w[k] = v[nb + 6 * i + 3 + k]; }
    for (int k = 0; k < nc; k++) { cs[k].ln = lam[njr + 3 * k]; cs[k].lt1 = lam[njr + 3 * k + 1]; cs[k].lt2 = lam[njr + 3 * k + 2]; }
    bodies_end(W, X, bw);
    export_contacts(cs, nc, dt, out, max_out);
    return nc;
}

/* ------------------------------------------------------------------------------------------- */
/* Per-env domain randomisation (python/gym_ignition_environments/randomizers/cartpole.py:51-56, */
/* 100-135): link masses + U(-delta, delta), gravity_z ~ N(g_z, sigma), redrawn at every reset.   */
/* rand[n_envs][nq+1] = body mass offsets and gravity scale g_z / g_z0 (in/out).                 */
/* ------------------------------------------------------------------------------------------- */
void b2o_sample_rand_params(uint64_t seed, uint64_t env, uint64_t step, int nq, double delta, double sigma,
                            double g0, const double* mass, double* out)
{
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    double u[4];
    for (int blk = 0; blk < 2; blk++) {
        uint32_t ctr[4] = {(uint32_t)step, (uint32_t)(step >> 32), (uint32_t)env,
                           ((uint32_t)(env >> 32) << 8) | (uint32_t)(blk + 2)};
        uint32_t r[4];
        b2o_philox4x32_10(ctr, key, r);
        u[2 * blk] = ((double)(r[0] >> 5) * 67108864.0 + (double)(r[1] >> 6)) / 9007199254740992.0;
        u[2 * blk + 1] = ((double)(r[2] >> 5) * 67108864.0 + (double)(r[3] >> 6)) / 9007199254740992.0;
    }
    for (int k = 0; k < nq; k++) {
        double dm = -delta + 2.0 * delta * u[k], lo = -0.9 * mass[k];
        out[k] = dm < lo ? lo : dm;
    }
    double z = sqrt(-2.0 * log(1.0 - u[2])) * cos(6.283185307179586 * u[3]);
    out[nq] = (g0 + sigma * z) / g0;
}

void b2o_rollout_randomized(const b2o_model* m, int task, double dt, int steps_per_run, int max_episode_steps,
                            uint64_t seed, uint64_t env_offset, uint64_t first_step, int n_envs, int T,
                            const double* actions, double* state, int32_t* elapsed, double* rand,
                            double mass_delta, double gravity_sigma, double* obs, double* reward, uint8_t* done)
{
    const int nq = m->nb, nobs = b2o_task_nobs(task);
    for (int e = 0; e < n_envs; e++) {
        double* st = state + (size_t)e * 2 * nq;
        double* rp = rand + (size_t)e * (nq + 1);
        for (int t = 0; t < T; t++) {
            b2o_model me = *m;   /* this env's own model: randomised masses and gravity */
            for (int k = 0; k < nq; k++) me.mass[k] = m->mass[k] + rp[k];
            me.gravity[2] = m->gravity[2] * rp[nq];
            double tau[B2O_MAXB] = {0}, o[8], r;
            int joint;
            double f = b2o_task_action_force(task, actions[(size_t)t * n_envs + e], &joint);
            tau[joint] = f;
            for (int it = 0; it < steps_per_run; it++) {
                b2o_physics_step(&me, dt, st, st + nq, tau, NULL);
                tau[joint] = 0.0;
            }
            int d = b2o_task_evaluate(task, st, 0.0, o, &r);
            elapsed[e] += 1;
            if (elapsed[e] >= max_episode_steps) d = 1;
            size_t idx = (size_t)t * n_envs + e;
            if (obs) memcpy(obs + idx * nobs, o, sizeof(double) * nobs);
            if (reward) reward[idx] = r;
            if (done) done[idx] = (uint8_t)d;
            if (d) {
                b2o_task_sample_reset(task, seed, env_offset + e, first_step + t, st);
                b2o_sample_rand_params(seed, env_offset + e, first_step + t, nq, mass_delta, gravity_sigma,
                                       m->gravity[2], m->mass, rp);
                elapsed[e] = 0;
            }
        }
    }
}


/* Link velocity and classical acceleration of a point fixed in `body` (point in the body frame), world
 * orientation: out = [v(3), w(3), a(3), alpha(3)]. What Physics.cpp:1989-2079 reads back per link. */
void b2o_link_motion(const b2o_model* m, const double* q, const double* dq, const double* ddq, int body,
                     const double* point, double* out)
{
    kin_t k;
    v6 V[B2O_MAXB], A[B2O_MAXB];
    kinematics(m, q, &k);
    for (int i = 0; i < m->nb; i++) {
        v6 Vp = {0, 0, 0, 0, 0, 0}, Ap = {0, 0, 0, 0, 0, 0}, Sdq, eta;
        if (m->parent[i] >= 0) {
            m6v(k.X[i], V[m->parent[i]], Vp);
            m6v(k.X[i], A[m->parent[i]], Ap);
        }
        for (int a = 0; a < 6; a++) { Sdq[a] = k.S[i][a] * dq[i]; V[i][a] = Vp[a] + Sdq[a]; }
        crm(V[i], Sdq, eta);
        for (int a = 0; a < 6; a++) A[i][a] = Ap[a] + eta[a] + k.S[i][a] * ddq[i];
    }
    memset(out, 0, 12 * sizeof(double));
    if (body < 0) return;
    double wxr[3], vpt[3], axr[3], wxv[3], apt[3];
    cross3(V[body], point, wxr);
    for (int a = 0; a < 3; a++) vpt[a] = V[body][3 + a] + wxr[a];
    cross3(A[body], point, axr);
    cross3(V[body], vpt, wxv);
    for (int a = 0; a < 3; a++) apt[a] = A[body][3 + a] + axr[a] + wxv[a];
    m3v(k.Rw[body], vpt, out);
    m3v(k.Rw[body], V[body], out + 3);
    m3v(k.Rw[body], apt, out + 6);
    m3v(k.Rw[body], A[body], out + 9);
}


/* ------------------------------------------------------------------------------------------- */
/* KinDynComputations centre of mass / momentum (kindyncomputations.py:305-342; iDynTree is not in */
/* the tree). Restated through the point Jacobians of the body centres of mass:                     */
/*   v_ci = J_lin(c_i) dq, w_i = J_ang dq,  L = sum m_i v_ci,  H_O = sum (R I_ci R^T w_i + c_i x m v_ci) */
/*   com = sum m_i c_i / M,  H_G = H_O - com x L,  J_com = sum m_i J_lin(c_i) / M.                  */
/* base_mass / base_mc: mass and first moment (base frame) of the links welded to the fixed base.    */
/* out: com[3], com_velocity[3], momentum[6] (about the world origin), centroidal[6], Jcom[3][nb].    */
/* ------------------------------------------------------------------------------------------- */
void b2o_centroidal(const b2o_model* m, const double* q, const double* dq, double base_mass, const double* base_mc,
                    double* com, double* com_velocity, double* momentum, double* centroidal, double* Jcom)
{
    const int nb = m->nb;
    double R[B2O_MAXB * 9], p[B2O_MAXB * 3], J[6 * B2O_MAXB];
    double M = base_mass, first[3], L[3] = {0, 0, 0}, H[3] = {0, 0, 0}, t[3];
    b2o_forward_kinematics(m, q, R, p);
    m3v(m->base_R, base_mc, t);
    for (int k = 0; k < 3; k++) first[k] = base_mass * m->base_p[k] + t[k];
    for (int k = 0; k < 3 * nb; k++) Jcom[k] = 0;
    for (int i = 0; i < nb; i++) {
        double cw[3], v[3] = {0, 0, 0}, w[3] = {0, 0, 0}, wb[3], Iw[3], Iww[3], cxv[3];
        b2o_point_jacobian(m, q, i, m->com[i], J);
        for (int j = 0; j < nb; j++)
            for (int k = 0; k < 3; k++) { v[k] += J[k * nb + j] * dq[j]; w[k] += J[(3 + k) * nb + j] * dq[j]; }
        m3v(R + 9 * i, m->com[i], cw);
        for (int k = 0; k < 3; k++) cw[k] += p[3 * i + k];
        M += m->mass[i];
        for (int k = 0; k < 3; k++) { first[k] += m->mass[i] * cw[k]; L[k] += m->mass[i] * v[k]; }
        m3tv(R + 9 * i, w, wb);                 /* angular velocity in body axes */
        m3v(m->Ic[i], wb, Iw);
        m3v(R + 9 * i, Iw, Iww);
        cross3(cw, v, cxv);
        for (int k = 0; k < 3; k++) H[k] += Iww[k] + m->mass[i] * cxv[k];
        for (int j = 0; j < nb; j++)
            for (int k = 0; k < 3; k++) Jcom[k * nb + j] += m->mass[i] * J[k * nb + j];
    }
    for (int k = 0; k < 3; k++) { com[k] = first[k] / M; com_velocity[k] = L[k] / M; }
    for (int k = 0; k < 3 * nb; k++) Jcom[k] /= M;
    cross3(com, L, t);
    for (int k = 0; k < 3; k++) {
        momentum[k] = L[k]; momentum[3 + k] = H[k];
        centroidal[k] = L[k]; centroidal[3 + k] = H[k] - t[k];
    }
}

/* ------------------------------------------------------------------------------------------- */
/* Momentum Jacobian of KinDynComputations (kindyncomputations.py:379-427; iDynTree is not in the */
/* tree; MIXED representation: world orientation, momentum taken about the origin of the base).  */
/* Restated link by link through the point Jacobians of the body centres of mass:                 */
/*   h = [sum m_i v_ci ; sum (R I_ci R^T w_i + (c_i - p_base) x m_i v_ci)] = Jmom dq              */
/*   Jmom[:, j] = [sum m_i Jlin_i[:, j] ; sum (R I_ci R^T Jang_i[:, j] + r_i x m_i Jlin_i[:, j])]  */
/* and the locked inertia about the base origin as the plain sum over the links (parallel-axis).  */
/* base_Io: rotational inertia (xx, xy, xz, yy, yz, zz) of the links welded to the base about the  */
/* base origin, base frame. out: Jmom[6][nb] (linear rows first), locked[10] (xx, xy, xz, yy, yz, */
/* zz, m c, m), world orientation.                                                               */
/* ------------------------------------------------------------------------------------------- */
void b2o_momentum_jacobian(const b2o_model* m, const double* q, double base_mass, const double* base_mc,
                           const double* base_Io, double* Jmom, double* locked)
{
    const int nb = m->nb;
    double R[B2O_MAXB * 9], p[B2O_MAXB * 3], J[6 * B2O_MAXB];
    b2o_forward_kinematics(m, q, R, p);
    for (int k = 0; k < 6 * nb; k++) Jmom[k] = 0;
    double A[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, mc[3] = {0, 0, 0}, M = base_mass;
    {   /* links welded to the base: R_b Io R_b^T, R_b mc */
        const double Ib[9] = {base_Io[0], base_Io[1], base_Io[2], base_Io[1], base_Io[3], base_Io[4],
                              base_Io[2], base_Io[4], base_Io[5]};
        double t1[9];
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) {
                t1[3 * r + c] = 0;
                for (int k = 0; k < 3; k++) t1[3 * r + c] += m->base_R[3 * r + k] * Ib[3 * k + c];
            }
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++)
                for (int k = 0; k < 3; k++) A[3 * r + c] += t1[3 * r + k] * m->base_R[3 * c + k];
        m3v(m->base_R, base_mc, mc);
    }
    for (int i = 0; i < nb; i++) {
        double cw[3], r[3], Iw[9], t1[9];
        const double mi = m->mass[i];
        b2o_point_jacobian(m, q, i, m->com[i], J);
        m3v(R + 9 * i, m->com[i], cw);
        for (int k = 0; k < 3; k++) { cw[k] += p[3 * i + k]; r[k] = cw[k] - m->base_p[k]; }
        for (int a = 0; a < 3; a++)
            for (int c = 0; c < 3; c++) {
                t1[3 * a + c] = 0;
                for (int k = 0; k < 3; k++) t1[3 * a + c] += R[9 * i + 3 * a + k] * m->Ic[i][3 * k + c];
            }
        for (int a = 0; a < 3; a++)
            for (int c = 0; c < 3; c++) {
                Iw[3 * a + c] = 0;
                for (int k = 0; k < 3; k++) Iw[3 * a + c] += t1[3 * a + k] * R[9 * i + 3 * c + k];
            }
        for (int j = 0; j < nb; j++) {
            double lin[3], ang[3], Iwa[3], rxl[3];
            for (int k = 0; k < 3; k++) { lin[k] = J[k * nb + j]; ang[k] = J[(3 + k) * nb + j]; }
            m3v(Iw, ang, Iwa);
            cross3(r, lin, rxl);
            for (int k = 0; k < 3; k++) {
                Jmom[k * nb + j] += mi * lin[k];
                Jmom[(3 + k) * nb + j] += Iwa[k] + mi * rxl[k];
            }
        }
        const double rr = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
        for (int a = 0; a < 3; a++)
            for (int c = 0; c < 3; c++) A[3 * a + c] += Iw[3 * a + c] + mi * ((a == c ? rr : 0.0) - r[a] * r[c]);
        for (int k = 0; k < 3; k++) mc[k] += mi * r[k];
        M += mi;
    }
    locked[0] = A[0]; locked[1] = A[1]; locked[2] = A[2]; locked[3] = A[4]; locked[4] = A[5]; locked[5] = A[8];
    locked[6] = mc[0]; locked[7] = mc[1]; locked[8] = mc[2]; locked[9] = M;
}
