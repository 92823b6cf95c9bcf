// Test helper: runs the engine's templated rigid-body code (csrc/b2_rbd.hpp) on the host so that the
// CPU test-suite can compare it with the oracle without a GPU. Not part of the product library.
#include <cstring>

#include "../../gym-ignition_b200/csrc/b2_model.hpp"
#include "../../gym-ignition_b200/csrc/b2_rbd.hpp"
#include "../../gym-ignition_b200/csrc/b2_tree_fast.hpp"

using namespace b2;

static bool tables(const char* xml, const double* pose7, const double* g, ModelDev<double>& md)
{
    try {
        b2model* m = parse_model(xml, strlen(xml));
        Pose base = pose7 ? pose_from_xyz_quat(pose7) : Pose();
        m->to_device_tables<double>(base, g, md);
        delete m;
        return true;
    } catch (...) {
        return false;
    }
}

extern "C" {

int rbd_forward_dynamics(const char* xml, const double* pose7, const double* g, double dt, const double* q,
                         const double* dq, const double* tau, double* ddq)
{
    ModelDev<double> md;
    if (!tables(xml, pose7, g, md)) return -1;
    forward_dynamics<double, kMaxDofs>(md, dt, q, dq, tau, ddq);
    return md.nq;
}
// the low-footprint variant the Panda kernels use, on a plain (stride 1) scratch array
int rbd_forward_dynamics_fast(const char* xml, const double* pose7, const double* g, double dt, const double* q,
                              const double* dq, const double* tau, double* ddq)
{
    ModelDev<double> md;
    if (!tables(xml, pose7, g, md)) return -1;
    int slot[kMaxDofs];
    const int nbranch = branch_slots(md.nq, md.parent, slot);
    if (nbranch < 0) return -2;
    double buf[kSlotsPerBody * kMaxDofs + kSlotsPerBranch * kMaxBranch];
    Scratch<double> w{buf, 1};
    for (int i = 0; i < md.nq; ++i) {
        w[kSlotsPerBody * i + SL_Q] = q[i];
        w[kSlotsPerBody * i + SL_DQ] = dq[i];
        w[kSlotsPerBody * i + SL_TAU] = tau[i];
    }
    clear_parking(md.nq, nbranch, w);
    forward_dynamics_fast(md, slot, dt, w);
    for (int i = 0; i < md.nq; ++i) ddq[i] = w[kSlotsPerBody * i + SL_TAU];
    return md.nq;
}
// one full physics iteration (ABA -> dq += ddq dt -> constraint rows by impulse responses -> q += dq dt)
int rbd_step_fast(const char* xml, const double* pose7, const double* g, double dt, double* q, double* dq,
                  const double* tau)
{
    ModelDev<double> md;
    if (!tables(xml, pose7, g, md)) return -1;
    int slot[kMaxDofs];
    const int nbranch = branch_slots(md.nq, md.parent, slot);
    if (nbranch < 0) return -2;
    double buf[kSlotsPerBody * kMaxDofs + kSlotsPerBranch * kMaxBranch];
    Scratch<double> w{buf, 1};
    for (int i = 0; i < md.nq; ++i) {
        w[kSlotsPerBody * i + SL_Q] = q[i];
        w[kSlotsPerBody * i + SL_DQ] = dq[i];
        w[kSlotsPerBody * i + SL_TAU] = tau[i];
    }
    clear_parking(md.nq, nbranch, w);
    forward_dynamics_fast(md, slot, dt, w);
    for (int i = 0; i < md.nq; ++i) w[kSlotsPerBody * i + SL_DQ] += w[kSlotsPerBody * i + SL_TAU] * dt;
    int rj[kMaxRows];
    double rb[kMaxRows], rlo[kMaxRows], rhi[kMaxRows], none[kMaxDofs] = {0};
    const int nr = collect_rows(md, dt, w, 0u, none, rj, rb, rlo, rhi);
    if (nr < 0) return -3;
    if (nr > 0) constraints_fast(md, dt, w, nr, rj, rb, rlo, rhi);
    for (int i = 0; i < md.nq; ++i) {
        w[kSlotsPerBody * i + SL_Q] += w[kSlotsPerBody * i + SL_DQ] * dt;
        q[i] = w[kSlotsPerBody * i + SL_Q];
        dq[i] = w[kSlotsPerBody * i + SL_DQ];
    }
    return nr;
}
int rbd_inverse_dynamics(const char* xml, const double* pose7, const double* g, const double* q, const double* dq,
                         const double* ddq, int gravity, double* tau)
{
    ModelDev<double> md;
    if (!tables(xml, pose7, g, md)) return -1;
    inverse_dynamics<double, kMaxDofs>(md, q, dq, ddq, gravity != 0, tau);
    return md.nq;
}
int rbd_mass_matrix(const char* xml, const double* pose7, const double* g, const double* q, double* M)
{
    ModelDev<double> md;
    if (!tables(xml, pose7, g, md)) return -1;
    mass_matrix<double, kMaxDofs>(md, q, M);
    return md.nq;
}
int rbd_forward_kinematics(const char* xml, const double* pose7, const double* g, const double* q, double* R, double* p)
{
    ModelDev<double> md;
    if (!tables(xml, pose7, g, md)) return -1;
    M3<double> Rw[kMaxDofs];
    V3<double> pw[kMaxDofs];
    forward_kinematics<double, kMaxDofs>(md, q, Rw, pw);
    for (int i = 0; i < md.nq; ++i) {
        memcpy(R + 9 * i, Rw[i].m, sizeof(double) * 9);
        p[3 * i] = pw[i].x; p[3 * i + 1] = pw[i].y; p[3 * i + 2] = pw[i].z;
    }
    return md.nq;
}
// closed-form chain step (the arithmetic of the fused task kernels) for (pose, gravity, dt)
int rbd_chain_step(const char* xml, const double* pose7, const double* g, double dt, double* state, const double* tau)
{
    try {
        b2model* m = parse_model(xml, strlen(xml));
        Pose base = pose7 ? pose_from_xyz_quat(pose7) : Pose();
        int kind = m->fit(base, g, dt);
        double a0, a1;
        if (kind == B2_KIND_CHAIN1) chain1_step(m->coef, state[0], state[1], tau[0], a0);
        else if (kind == B2_KIND_CHAIN_PR) chain_pr_step(m->coef, state[0], state[1], state[2], state[3], tau[0], tau[1], a0, a1);
        delete m;
        return kind;
    } catch (...) {
        return -1;
    }
}
}
