// Test helper: runs the engine's templated rigid-body code (csrc/b2_rbd.hpp) on the host so that the
// CPU test-suite can compare it with the oracle without a GPU. Not part of the product library.
#include <cstring>

#include "../../gym-ignition_b200/csrc/b2_model.hpp"
#include "../../gym-ignition_b200/csrc/b2_rbd.hpp"
#include "../../gym-ignition_b200/csrc/b2_tree_fast.hpp"
#include "../../gym-ignition_b200/csrc/b2_contact.hpp"

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
int rbd_centroidal(const char* xml, const double* pose7, const double* g, const double* q, const double* dq, double* com,
                   double* vel, double* mom, double* jac)
{
    ModelDev<double> md;
    if (!tables(xml, pose7, g, md)) return -1;
    centroidal<double, kMaxDofs>(md, q, dq, com, vel, mom, jac);
    return md.nq;
}
int rbd_momentum(const char* xml, const double* pose7, const double* g, const double* q, double* jmom, double* locked)
{
    ModelDev<double> md;
    if (!tables(xml, pose7, g, md)) return -1;
    momentum_matrices<double, kMaxDofs>(md, q, jmom, locked);
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

// free bodies with contacts: the engine's templated world step on the host. The world is described with flat
// arrays: per body mass, Ic[9], com[3], box half extents[3], mu; static shapes: type, size[3], R[9], p[3], mu.
static void fill_world(WorldDev<double>& W, int nfree, const double* body_params, int nstatic, const double* static_params,
                       double dt, const double* g, int iterations, double erp, double max_erv);
static int export_contacts(const Contact<double>* cs, int nc, double dt, double* contacts_out);

int contact_world_step(int nfree, const double* body_params /* [nfree][17] */, int nstatic,
                       const double* static_params /* [nstatic][17]: type,size3,R9,p3,mu */, double dt, const double* g,
                       int iterations, double erp, double max_erv, double* X, double* contacts_out /* [32][12] */)
{
    static WorldDev<double> W;
    fill_world(W, nfree, body_params, nstatic, static_params, dt, g, iterations, erp, max_erv);
    Contact<double> cs[kMaxContacts];
    const int nc = world_step(W, X, cs);
    return export_contacts(cs, nc, dt, contacts_out);
}

// coupled world: one physics iteration of an articulated model (URDF `xml`) whose link shapes touch free bodies and
// static shapes. robot_params: [nrobot][18] = body, type, size3 (box: half extents), R9, p3 (body frame), mu.
int coupled_world_step(const char* xml, const double* pose7, int nfree, const double* body_params, int nstatic,
                       const double* static_params, int nrobot, const double* robot_params, double dt, const double* g,
                       int iterations, double erp, double max_erv, double* q, double* dq, const double* tau, double* ddq,
                       double* X, double* contacts_out)
{
    static WorldDev<double> W;
    static ModelDev<double> md;
    if (!tables(xml, pose7, g, md)) return -1;
    fill_world(W, nfree, body_params, nstatic, static_params, dt, g, iterations, erp, max_erv);
    W.nrobot = nrobot;
    for (int r = 0; r < nrobot; ++r) {
        const double* rp = robot_params + 18 * r;
        W.rbody[r] = (int)rp[0];
        ShapeDev<double>& s = W.rshape[r];
        s.type = (int)rp[1];
        for (int k = 0; k < 3; ++k) s.size[k] = rp[2 + k];
        for (int k = 0; k < 9; ++k) s.R[k] = rp[5 + k];
        for (int k = 0; k < 3; ++k) s.p[k] = rp[14 + k];
        s.mu = rp[17];
    }
    double acc[kMaxDofs], before[kMaxDofs];
    forward_dynamics<double, kMaxDofs>(md, dt, q, dq, tau, acc);
    for (int j = 0; j < md.nq; ++j) { dq[j] += acc[j] * dt; before[j] = dq[j]; }
    Contact<double> cs[kMaxContacts];
    static RobotWork<double> rw;
    const int nc = coupled_step(W, md, q, dq, 0u, (const double*)nullptr, X, cs, rw);
    for (int j = 0; j < md.nq; ++j) { ddq[j] = acc[j] + (dq[j] - before[j]) / dt; q[j] += dq[j] * dt; }
    return export_contacts(cs, nc, dt, contacts_out);
}
}

static void fill_world(WorldDev<double>& W, int nfree, const double* body_params, int nstatic, const double* static_params,
                       double dt, const double* g, int iterations, double erp, double max_erv)
{
    memset(&W, 0, sizeof W);
    W.nfree = nfree; W.nstatic = nstatic; W.iterations = iterations;
    W.dt = dt; W.erp = erp; W.max_erv = max_erv;
    for (int k = 0; k < 3; ++k) W.g[k] = g[k];
    for (int i = 0; i < nfree; ++i) {
        const double* bp = body_params + 17 * i;
        FreeBodyDev<double>& b = W.body[i];
        b.mass = bp[0];
        for (int k = 0; k < 9; ++k) b.Ic[k] = bp[1 + k];
        const double* I = b.Ic;
        const double det = I[0] * (I[4] * I[8] - I[5] * I[7]) - I[1] * (I[3] * I[8] - I[5] * I[6]) + I[2] * (I[3] * I[7] - I[4] * I[6]);
        const double inv[9] = {(I[4] * I[8] - I[5] * I[7]) / det, (I[2] * I[7] - I[1] * I[8]) / det, (I[1] * I[5] - I[2] * I[4]) / det,
                               (I[5] * I[6] - I[3] * I[8]) / det, (I[0] * I[8] - I[2] * I[6]) / det, (I[2] * I[3] - I[0] * I[5]) / det,
                               (I[3] * I[7] - I[4] * I[6]) / det, (I[1] * I[6] - I[0] * I[7]) / det, (I[0] * I[4] - I[1] * I[3]) / det};
        for (int k = 0; k < 9; ++k) b.Ic_inv[k] = inv[k];
        for (int k = 0; k < 3; ++k) b.com[k] = bp[10 + k];
        b.nshapes = 1;
        b.shape[0].type = kShapeBox;
        for (int k = 0; k < 3; ++k) { b.shape[0].size[k] = bp[13 + k]; b.shape[0].p[k] = 0; }
        const double eye[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        for (int k = 0; k < 9; ++k) b.shape[0].R[k] = eye[k];
        b.shape[0].mu = bp[16];
    }
    for (int i = 0; i < nstatic; ++i) {
        const double* sp = static_params + 17 * i;
        ShapeDev<double>& s = W.stat[i];
        s.type = (int)sp[0];
        for (int k = 0; k < 3; ++k) s.size[k] = sp[1 + k];
        for (int k = 0; k < 9; ++k) s.R[k] = sp[4 + k];
        for (int k = 0; k < 3; ++k) s.p[k] = sp[13 + k];
        s.mu = sp[16];
    }
}

static int export_contacts(const Contact<double>* cs, int nc, double dt, double* contacts_out)
{
    for (int k = 0; k < nc; ++k) {
        double* o = contacts_out + 12 * k;
        o[0] = cs[k].a; o[1] = cs[k].b;
        o[2] = cs[k].pos.x; o[3] = cs[k].pos.y; o[4] = cs[k].pos.z;
        o[5] = cs[k].n.x; o[6] = cs[k].n.y; o[7] = cs[k].n.z;
        o[8] = cs[k].depth;
        const V3<double> f = contact_force(cs[k], dt);
        o[9] = f.x; o[10] = f.y; o[11] = f.z;
    }
    return nc;
}
