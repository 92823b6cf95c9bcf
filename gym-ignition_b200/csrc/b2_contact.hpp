// Free-floating rigid bodies with ground / box contacts: one world = a few free bodies + static shapes.
//
// What it replaces for these bodies: DART World::step as reached from
// cpp/scenario/plugins/Physics/Physics.cpp:1824-1835 (free-joint dynamics, collision detection, contact
// constraints) and the contact readback of Physics.cpp:2351-2540 / cpp/scenario/gazebo/src/Link.cpp:360-482.
// DART solves the contact LCP with Dantzig (PGS fallback) on ODE/FCL contact points; this is the "LCP-lite"
// the north star allows: box-corner / sphere contact points against planes and boxes, friction pyramid,
// projected Gauss-Seidel with a fixed iteration count, DART's error-reduction velocity for penetration.
// Contact configurations are compared on trajectory statistics (resting force, height, contact count), not on
// trajectories. Host+device templates: the CUDA kernel, the host test helper and nothing else use them.
#pragma once

#include "b2_rbd.hpp"

namespace b2 {

constexpr int kMaxFree = 8;
constexpr int kMaxStaticShapes = 16;
constexpr int kMaxBodyShapes = 2;
constexpr int kMaxContacts = 32;
constexpr int kContactRec = 10;  // position(3), normal(3), depth, force(3)

enum { kShapeBox = 0, kShapeSphere = 1, kShapeCylinder = 2, kShapePlane = 3 };

template <typename T>
struct ShapeDev {
    int type;
    int owner_model, owner_link;  // for reporting
    T size[3];                    // box: half extents; sphere: radius; plane: unit normal (world)
    T R[9];                       // body frame (free bodies) or world frame (static shapes)
    T p[3];
    T mu;
};

template <typename T>
struct FreeBodyDev {
    T mass;
    T Ic[9];      // about the COM, body axes
    T Ic_inv[9];
    T com[3];     // COM in the body (root link) frame
    int nshapes;
    ShapeDev<T> shape[kMaxBodyShapes];
};

template <typename T>
struct WorldDev {
    int nfree, nstatic, iterations;
    T dt, erp, max_erv;
    T g[3];
    FreeBodyDev<T> body[kMaxFree];
    ShapeDev<T> stat[kMaxStaticShapes];
};

template <typename T> B2_HD M3<T> quat_to_rot(const T* q)
{
    const T w = q[0], x = q[1], y = q[2], z = q[3];
    return M3<T>{{T(1) - T(2) * (y * y + z * z), T(2) * (x * y - z * w), T(2) * (x * z + y * w),
                  T(2) * (x * y + z * w), T(1) - T(2) * (x * x + z * z), T(2) * (y * z - x * w),
                  T(2) * (x * z - y * w), T(2) * (y * z + x * w), T(1) - T(2) * (x * x + y * y)}};
}

template <typename T>
struct Contact {
    int a, b;          // a: free body index; b: free body index, or -1 - static shape index
    int shape_a;       // shape of body a that generated the point (reporting)
    V3<T> pos, n;      // world; n points from b towards a
    T depth, mu;
    V3<T> t1, t2;
    T kn, kt1, kt2;    // effective masses along n, t1, t2
    T ln, lt1, lt2;    // accumulated impulses
    T bias;
};

// per-body working set during the solve
template <typename T>
struct BodyWork {
    V3<T> xc, vc, w;   // COM position, COM velocity, angular velocity (world)
    M3<T> R, Iinv;     // orientation, world inverse inertia
    T inv_mass;
};

template <typename T>
B2_HD void add_contact(Contact<T>* cs, int& nc, int a, int shape_a, int b, V3<T> pos, V3<T> n, T depth, T mu)
{
    if (nc >= kMaxContacts) return;
    Contact<T>& c = cs[nc++];
    c.a = a; c.b = b; c.shape_a = shape_a;
    c.pos = pos; c.n = n; c.depth = depth; c.mu = mu;
    c.ln = c.lt1 = c.lt2 = T(0);
}

// Corners of `box` (half extents h, world pose Rb / pb) against the reference face of box B, or against a plane.
template <typename T>
B2_HD void box_vs_shape(Contact<T>* cs, int& nc, int a, int shape_a, int b, const M3<T>& Ra, V3<T> pa, const T* ha,
                        const ShapeDev<T>& sb, const M3<T>& Rb, V3<T> pb, T mu)
{
    if (sb.type == kShapePlane) {
        const V3<T> n = ld3(sb.size);
        for (int k = 0; k < 8; ++k) {
            const V3<T> cl = v3((k & 1) ? ha[0] : -ha[0], (k & 2) ? ha[1] : -ha[1], (k & 4) ? ha[2] : -ha[2]);
            const V3<T> x = pa + mul(Ra, cl);
            const T d = dot(n, x - pb);
            if (d <= T(0)) add_contact(cs, nc, a, shape_a, b, x, n, -d, mu);
        }
        return;
    }
    if (sb.type != kShapeBox) return;
    // reference face of B: the one A's centre is most outside of
    const V3<T> cb = mulT(Rb, pa - pb);
    const T ex[3] = {fabs(cb.x) - sb.size[0], fabs(cb.y) - sb.size[1], fabs(cb.z) - sb.size[2]};
    int axis = 0;
    if (ex[1] > ex[axis]) axis = 1;
    if (ex[2] > ex[axis]) axis = 2;
    const T cbv[3] = {cb.x, cb.y, cb.z};
    const T sgn = cbv[axis] >= T(0) ? T(1) : T(-1);
    const T margin = T(1e-6);
    for (int k = 0; k < 8; ++k) {
        const V3<T> cl = v3((k & 1) ? ha[0] : -ha[0], (k & 2) ? ha[1] : -ha[1], (k & 4) ? ha[2] : -ha[2]);
        const V3<T> x = pa + mul(Ra, cl);
        const V3<T> xl = mulT(Rb, x - pb);
        const T xv[3] = {xl.x, xl.y, xl.z};
        bool inside = true;
        for (int j = 0; j < 3; ++j)
            if (j != axis && fabs(xv[j]) > sb.size[j] + margin) inside = false;
        const T depth = sb.size[axis] - sgn * xv[axis];
        if (inside && depth >= T(0) && depth <= T(2) * sb.size[axis]) {
            const V3<T> nl = v3(axis == 0 ? sgn : T(0), axis == 1 ? sgn : T(0), axis == 2 ? sgn : T(0));
            add_contact(cs, nc, a, shape_a, b, x, mul(Rb, nl), depth, mu);
        }
    }
}

template <typename T>
B2_HD void sphere_vs_shape(Contact<T>* cs, int& nc, int a, int shape_a, int b, V3<T> ca, T r, const ShapeDev<T>& sb,
                           const M3<T>& Rb, V3<T> pb, T mu)
{
    if (sb.type == kShapePlane) {
        const V3<T> n = ld3(sb.size);
        const T d = dot(n, ca - pb) - r;
        if (d <= T(0)) add_contact(cs, nc, a, shape_a, b, ca - r * n, n, -d, mu);
    } else if (sb.type == kShapeBox) {
        const V3<T> cl = mulT(Rb, ca - pb);
        const T cv[3] = {cl.x, cl.y, cl.z};
        T q[3];
        bool inside = true;
        for (int j = 0; j < 3; ++j) {
            q[j] = cv[j] < -sb.size[j] ? -sb.size[j] : (cv[j] > sb.size[j] ? sb.size[j] : cv[j]);
            if (q[j] != cv[j]) inside = false;
        }
        if (inside) return;  // centre inside the box: not handled
        const V3<T> dl = v3(cv[0] - q[0], cv[1] - q[1], cv[2] - q[2]);
        const T dist = sqrt(dot(dl, dl));
        if (dist <= r) {
            const V3<T> n = mul(Rb, (T(1) / dist) * dl);
            add_contact(cs, nc, a, shape_a, b, pb + mul(Rb, v3(q[0], q[1], q[2])), n, r - dist, mu);
        }
    }
}

template <typename T> B2_HD V3<T> point_velocity(const BodyWork<T>& b, V3<T> r) { return b.vc + cross(b.w, r); }

template <typename T>
B2_HD T effective_mass(const BodyWork<T>* bw, const Contact<T>& c, V3<T> d)
{
    const BodyWork<T>& A = bw[c.a];
    const V3<T> ra = c.pos - A.xc;
    T k = A.inv_mass + dot(d, cross(mul(A.Iinv, cross(ra, d)), ra));
    if (c.b >= 0) {
        const BodyWork<T>& B = bw[c.b];
        const V3<T> rb = c.pos - B.xc;
        k += B.inv_mass + dot(d, cross(mul(B.Iinv, cross(rb, d)), rb));
    }
    return k;
}

template <typename T>
B2_HD void apply_impulse(BodyWork<T>* bw, const Contact<T>& c, V3<T> P)
{
    BodyWork<T>& A = bw[c.a];
    A.vc = A.vc + A.inv_mass * P;
    A.w = A.w + mul(A.Iinv, cross(c.pos - A.xc, P));
    if (c.b >= 0) {
        BodyWork<T>& B = bw[c.b];
        B.vc = B.vc - B.inv_mass * P;
        B.w = B.w - mul(B.Iinv, cross(c.pos - B.xc, P));
    }
}

template <typename T>
B2_HD V3<T> relative_velocity(const BodyWork<T>* bw, const Contact<T>& c)
{
    V3<T> v = point_velocity(bw[c.a], c.pos - bw[c.a].xc);
    if (c.b >= 0) v = v - point_velocity(bw[c.b], c.pos - bw[c.b].xc);
    return v;
}

// One step of every free body of one world. X: nfree x 13 (position, quaternion wxyz, linear and angular velocity
// of the body frame, world coordinates). cs: caller-provided contact workspace (kMaxContacts). Returns the number
// of contacts; their impulses divided by dt are the contact forces on body a.
template <typename T>
B2_HD int world_step(const WorldDev<T>& W, T* X, Contact<T>* cs)
{
    BodyWork<T> bw[kMaxFree];
    const T dt = W.dt;
    const V3<T> g = ld3(W.g);
    for (int i = 0; i < W.nfree; ++i) {
        const FreeBodyDev<T>& fb = W.body[i];
        T* x = X + 13 * i;
        BodyWork<T>& b = bw[i];
        b.R = quat_to_rot(x + 3);
        const V3<T> rc = mul(b.R, ld3(fb.com));
        b.xc = ld3(x) + rc;
        b.w = ld3(x + 10);
        b.vc = ld3(x + 7) + cross(b.w, rc);
        b.inv_mass = T(1) / fb.mass;
        b.Iinv = mulBt(mul(b.R, ld9(fb.Ic_inv)), b.R);
        // unconstrained velocity update: gravity and the gyroscopic torque
        const M3<T> Iw = mulBt(mul(b.R, ld9(fb.Ic)), b.R);
        b.vc = b.vc + dt * g;
        b.w = b.w - dt * mul(b.Iinv, cross(b.w, mul(Iw, b.w)));
    }
    // ---- contact generation ----
    int nc = 0;
    for (int i = 0; i < W.nfree; ++i) {
        const FreeBodyDev<T>& fb = W.body[i];
        for (int s = 0; s < fb.nshapes; ++s) {
            const ShapeDev<T>& sa = fb.shape[s];
            const M3<T> Ra = mul(bw[i].R, ld9(sa.R));
            const V3<T> pa = bw[i].xc + mul(bw[i].R, ld3(sa.p) - ld3(fb.com));
            for (int k = 0; k < W.nstatic; ++k) {
                const ShapeDev<T>& sb = W.stat[k];
                const T mu = sa.mu < sb.mu ? sa.mu : sb.mu;
                if (sa.type == kShapeBox) box_vs_shape(cs, nc, i, s, -1 - k, Ra, pa, sa.size, sb, ld9(sb.R), ld3(sb.p), mu);
                else if (sa.type == kShapeSphere) sphere_vs_shape(cs, nc, i, s, -1 - k, pa, sa.size[0], sb, ld9(sb.R), ld3(sb.p), mu);
            }
            for (int j = 0; j < W.nfree; ++j) {
                if (j == i) continue;
                const FreeBodyDev<T>& fj = W.body[j];
                for (int u = 0; u < fj.nshapes; ++u) {
                    const ShapeDev<T>& sb = fj.shape[u];
                    if (sb.type != kShapeBox) continue;
                    const M3<T> Rb = mul(bw[j].R, ld9(sb.R));
                    const V3<T> pb = bw[j].xc + mul(bw[j].R, ld3(sb.p) - ld3(fj.com));
                    const T mu = sa.mu < sb.mu ? sa.mu : sb.mu;
                    if (sa.type == kShapeBox) box_vs_shape(cs, nc, i, s, j, Ra, pa, sa.size, sb, Rb, pb, mu);
                    else if (sa.type == kShapeSphere) sphere_vs_shape(cs, nc, i, s, j, pa, sa.size[0], sb, Rb, pb, mu);
                }
            }
        }
    }
    // ---- contact rows ----
    for (int k = 0; k < nc; ++k) {
        Contact<T>& c = cs[k];
        const V3<T> seed = fabs(c.n.x) < T(0.9) ? v3(T(1), T(0), T(0)) : v3(T(0), T(1), T(0));
        V3<T> t1 = cross(c.n, seed);
        t1 = (T(1) / sqrt(dot(t1, t1))) * t1;
        c.t1 = t1;
        c.t2 = cross(c.n, t1);
        c.kn = effective_mass(bw, c, c.n);
        c.kt1 = effective_mass(bw, c, c.t1);
        c.kt2 = effective_mass(bw, c, c.t2);
        T erv = c.depth * W.erp / dt;  // DART: penetration * ERP / dt, capped
        c.bias = erv > W.max_erv ? W.max_erv : erv;
    }
    // ---- projected Gauss-Seidel ----
    for (int it = 0; it < W.iterations; ++it) {
        for (int k = 0; k < nc; ++k) {
            Contact<T>& c = cs[k];
            V3<T> rel = relative_velocity(bw, c);
            const T ln = c.ln + (c.bias - dot(c.n, rel)) / c.kn;
            const T ln_new = ln > T(0) ? ln : T(0);
            apply_impulse(bw, c, (ln_new - c.ln) * c.n);
            c.ln = ln_new;
            const T lim = c.mu * c.ln;
            rel = relative_velocity(bw, c);
            T l1 = c.lt1 - dot(c.t1, rel) / c.kt1;
            l1 = l1 < -lim ? -lim : (l1 > lim ? lim : l1);
            apply_impulse(bw, c, (l1 - c.lt1) * c.t1);
            c.lt1 = l1;
            rel = relative_velocity(bw, c);
            T l2 = c.lt2 - dot(c.t2, rel) / c.kt2;
            l2 = l2 < -lim ? -lim : (l2 > lim ? lim : l2);
            apply_impulse(bw, c, (l2 - c.lt2) * c.t2);
            c.lt2 = l2;
        }
    }
    // ---- integrate poses (rotation by the exponential map) ----
    for (int i = 0; i < W.nfree; ++i) {
        const FreeBodyDev<T>& fb = W.body[i];
        T* x = X + 13 * i;
        BodyWork<T>& b = bw[i];
        const T wn = sqrt(dot(b.w, b.w));
        T q[4] = {x[3], x[4], x[5], x[6]};
        if (wn > T(0)) {
            T s, c;
            sincos_t(T(0.5) * wn * dt, &s, &c);
            const V3<T> ax = (s / wn) * b.w;
            const T d[4] = {c, ax.x, ax.y, ax.z};
            const T r0 = d[0] * q[0] - d[1] * q[1] - d[2] * q[2] - d[3] * q[3];
            const T r1 = d[0] * q[1] + d[1] * q[0] + d[2] * q[3] - d[3] * q[2];
            const T r2 = d[0] * q[2] - d[1] * q[3] + d[2] * q[0] + d[3] * q[1];
            const T r3 = d[0] * q[3] + d[1] * q[2] - d[2] * q[1] + d[3] * q[0];
            const T nrm = T(1) / sqrt(r0 * r0 + r1 * r1 + r2 * r2 + r3 * r3);
            q[0] = r0 * nrm; q[1] = r1 * nrm; q[2] = r2 * nrm; q[3] = r3 * nrm;
        }
        b.xc = b.xc + dt * b.vc;
        const M3<T> Rn = quat_to_rot(q);
        const V3<T> rc = mul(Rn, ld3(fb.com));
        const V3<T> p = b.xc - rc, v = b.vc - cross(b.w, rc);
        x[0] = p.x; x[1] = p.y; x[2] = p.z;
        x[3] = q[0]; x[4] = q[1]; x[5] = q[2]; x[6] = q[3];
        x[7] = v.x; x[8] = v.y; x[9] = v.z;
        x[10] = b.w.x; x[11] = b.w.y; x[12] = b.w.z;
    }
    return nc;
}

// Force on body a transmitted through contact c (world frame).
template <typename T> B2_HD V3<T> contact_force(const Contact<T>& c, T dt)
{
    return (T(1) / dt) * (c.ln * c.n + c.lt1 * c.t1 + c.lt2 * c.t2);
}

}  // namespace b2
