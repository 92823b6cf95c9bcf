"""TEST INFRASTRUCTURE ONLY — ctypes front-end of the CPU oracle (oracle/b2oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import urdf_tables

_HERE = os.path.dirname(os.path.abspath(__file__))
MAXB = 16

TASK_PENDULUM_SWINGUP = 1
TASK_CARTPOLE_DISCRETE_BALANCING = 2
TASK_CARTPOLE_CONTINUOUS_BALANCING = 3
TASK_CARTPOLE_CONTINUOUS_SWINGUP = 4

MODE_IDLE, MODE_FORCE, MODE_VELOCITY, MODE_VELOCITY_FOLLOWER_DART, MODE_POSITION = 1, 2, 3, 4, 5


class Model(C.Structure):
    _fields_ = [
        ("nb", C.c_int32),
        ("parent", C.c_int32 * MAXB),
        ("jtype", C.c_int32 * MAXB),
        ("axis", C.c_double * (MAXB * 3)),
        ("R", C.c_double * (MAXB * 9)),
        ("p", C.c_double * (MAXB * 3)),
        ("mass", C.c_double * MAXB),
        ("com", C.c_double * (MAXB * 3)),
        ("Ic", C.c_double * (MAXB * 9)),
        ("damping", C.c_double * MAXB),
        ("friction", C.c_double * MAXB),
        ("stiffness", C.c_double * MAXB),
        ("rest", C.c_double * MAXB),
        ("lower", C.c_double * MAXB),
        ("upper", C.c_double * MAXB),
        ("effort", C.c_double * MAXB),
        ("vmax", C.c_double * MAXB),
        ("gravity", C.c_double * 3),
        ("base_R", C.c_double * 9),
        ("base_p", C.c_double * 3),
    ]


def build(force=False):
    """Compile oracle/libb2oracle.so with gcc (building the checker is not using it)."""
    so = os.path.join(_HERE, "libb2oracle.so")
    src = os.path.join(_HERE, "b2oracle.c")
    hdr = os.path.join(_HERE, "b2oracle.h")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src),
                                                                       os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libb2oracle.so"],
                              stdout=subprocess.DEVNULL)
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        dp = C.POINTER(C.c_double)
        mp = C.POINTER(Model)
        L.b2o_forward_dynamics.argtypes = [mp, C.c_double, dp, dp, dp, dp]
        L.b2o_inverse_dynamics.argtypes = [mp, dp, dp, dp, C.c_int, dp]
        L.b2o_mass_matrix.argtypes = [mp, dp, dp]
        L.b2o_forward_kinematics.argtypes = [mp, dp, dp, dp]
        L.b2o_point_jacobian.argtypes = [mp, dp, C.c_int, dp, dp]
        L.b2o_physics_step.argtypes = [mp, C.c_double, dp, dp, dp, dp]
        L.b2o_energy.argtypes = [mp, dp, dp]
        L.b2o_energy.restype = C.c_double
        L.b2o_sim_create.argtypes = [mp, C.c_double, C.c_int]
        L.b2o_sim_create.restype = C.c_void_p
        L.b2o_sim_destroy.argtypes = [C.c_void_p]
        L.b2o_sim_run.argtypes = [C.c_void_p, C.c_int]
        L.b2o_sim_time.argtypes = [C.c_void_p]
        L.b2o_sim_time.restype = C.c_double
        L.b2o_sim_set_control_mode.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.b2o_sim_control_mode.argtypes = [C.c_void_p, C.c_int]
        L.b2o_sim_set_pid.argtypes = [C.c_void_p, C.c_int] + [C.c_double] * 8
        L.b2o_sim_set_controller_period.argtypes = [C.c_void_p, C.c_double]
        L.b2o_sim_load_computed_torque.argtypes = [C.c_void_p, dp, dp, dp]
        L.b2o_sim_apply_link_wrench.argtypes = [C.c_void_p, C.c_int, dp, dp, C.c_double]
        for n in ("set_force_target", "set_position_target", "set_velocity_target", "set_acceleration_target",
                  "reset_position", "reset_velocity"):
            getattr(L, "b2o_sim_" + n).argtypes = [C.c_void_p, C.c_int, C.c_double]
        for n in ("position", "velocity", "acceleration"):
            f = getattr(L, "b2o_sim_" + n)
            f.argtypes = [C.c_void_p, C.c_int]
            f.restype = C.c_double
        for n in ("force_target", "position_target", "velocity_target"):
            f = getattr(L, "b2o_sim_" + n)
            f.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
            f.restype = C.c_double
        L.b2o_pid_init.argtypes = [C.c_void_p] + [C.c_double] * 8
        L.b2o_pid_update.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.b2o_pid_update.restype = C.c_double
        L.b2o_philox4x32_10.argtypes = [C.POINTER(C.c_uint32)] * 3
        L.b2o_reset_uniforms.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, dp]
        L.b2o_task_sample_reset.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, dp]
        L.b2o_task_reset_from_uniforms.argtypes = [C.c_int, dp, dp]
        L.b2o_task_evaluate.argtypes = [C.c_int, dp, C.c_double, dp, dp]
        L.b2o_task_action_force.argtypes = [C.c_int, C.c_double, C.POINTER(C.c_int)]
        L.b2o_task_action_force.restype = C.c_double
        L.b2o_rollout.argtypes = [mp, C.c_int, C.c_double, C.c_int, C.c_int, C.c_uint64, C.c_uint64,
                                  C.c_uint64, C.c_int, C.c_int, dp, dp, C.POINTER(C.c_int32), dp, dp,
                                  C.POINTER(C.c_uint8)]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def model_from_tables(t):
    m = Model()
    m.nb = int(t["nb"])
    for name in ("parent", "jtype"):
        getattr(m, name)[:] = [int(v) for v in t[name]]
    for name in ("axis", "R", "p", "mass", "com", "Ic", "damping", "friction", "stiffness", "rest",
                 "lower", "upper", "effort", "vmax", "gravity", "base_R", "base_p"):
        getattr(m, name)[:] = np.asarray(t[name], float).ravel().tolist()
    return m


def load_urdf(path_or_xml, **kw):
    xml = path_or_xml
    if not path_or_xml.lstrip().startswith("<"):
        with open(path_or_xml) as f:
            xml = f.read()
    t = urdf_tables.flatten(xml, **kw)
    return t, model_from_tables(t)


class Dynamics:
    """Thin numpy API over the rigid-body functions of the oracle."""

    def __init__(self, model):
        self.m = model
        self.nb = model.nb

    def _v(self, x):
        a = np.zeros(MAXB)
        a[:self.nb] = np.asarray(x, float)
        return a

    def forward_dynamics(self, q, dq, tau, dt=0.0):
        q, dq, tau, out = self._v(q), self._v(dq), self._v(tau), np.zeros(MAXB)
        lib().b2o_forward_dynamics(C.byref(self.m), dt, _dp(q), _dp(dq), _dp(tau), _dp(out))
        return out[:self.nb].copy()

    def inverse_dynamics(self, q, dq, ddq, gravity=True):
        q, dq, ddq, out = self._v(q), self._v(dq), self._v(ddq), np.zeros(MAXB)
        lib().b2o_inverse_dynamics(C.byref(self.m), _dp(q), _dp(dq), _dp(ddq), int(gravity), _dp(out))
        return out[:self.nb].copy()

    def mass_matrix(self, q):
        q, M = self._v(q), np.zeros(self.nb * self.nb)
        lib().b2o_mass_matrix(C.byref(self.m), _dp(q), _dp(M))
        return M.reshape(self.nb, self.nb).copy()

    def forward_kinematics(self, q):
        q, R, p = self._v(q), np.zeros(self.nb * 9), np.zeros(self.nb * 3)
        lib().b2o_forward_kinematics(C.byref(self.m), _dp(q), _dp(R), _dp(p))
        return R.reshape(self.nb, 3, 3).copy(), p.reshape(self.nb, 3).copy()

    def point_jacobian(self, q, body, point=(0.0, 0.0, 0.0)):
        q, pt, J = self._v(q), np.array(point, float), np.zeros(6 * self.nb)
        lib().b2o_point_jacobian(C.byref(self.m), _dp(q), int(body), _dp(pt), _dp(J))
        return J.reshape(6, self.nb).copy()

    def step(self, q, dq, tau, dt):
        q, dq, tau, acc = self._v(q), self._v(dq), self._v(tau), np.zeros(MAXB)
        lib().b2o_physics_step(C.byref(self.m), dt, _dp(q), _dp(dq), _dp(tau), _dp(acc))
        return q[:self.nb].copy(), dq[:self.nb].copy(), acc[:self.nb].copy()

    def link_motion(self, q, dq, ddq, body, point=(0.0, 0.0, 0.0)):
        """[v, w, a, alpha] (world orientation) of a point fixed in `body`."""
        q, dq, ddq, pt, out = self._v(q), self._v(dq), self._v(ddq), np.array(point, float), np.zeros(12)
        L = lib()
        L.b2o_link_motion.argtypes = [C.POINTER(Model), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                      C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.b2o_link_motion(C.byref(self.m), _dp(q), _dp(dq), _dp(ddq), int(body), _dp(pt), _dp(out))
        return out.reshape(4, 3).copy()

    def centroidal(self, q, dq, base_mass=0.0, base_mc=(0.0, 0.0, 0.0)):
        """com, com velocity, momentum about the world origin [lin, ang], centroidal momentum, CoM Jacobian [3, nb]."""
        L = lib()
        dp = C.POINTER(C.c_double)
        L.b2o_centroidal.argtypes = [C.POINTER(Model), dp, dp, C.c_double, dp, dp, dp, dp, dp, dp]
        q, dq, bmc = self._v(q), self._v(dq), np.array(base_mc, float)
        com, vel, mom, cen, J = np.zeros(3), np.zeros(3), np.zeros(6), np.zeros(6), np.zeros(3 * self.nb)
        L.b2o_centroidal(C.byref(self.m), _dp(q), _dp(dq), float(base_mass), _dp(bmc), _dp(com), _dp(vel), _dp(mom),
                         _dp(cen), _dp(J))
        return com, vel, mom, cen, J.reshape(3, self.nb)

    def momentum_jacobian(self, q, base_mass=0.0, base_mc=(0.0, 0.0, 0.0), base_Io=(0.0,) * 6):
        """Joint block [6, nb] of the momentum Jacobian (linear rows first, about the base origin, world orientation) and
        the locked inertia [xx, xy, xz, yy, yz, zz, m c (3), m] about the base origin."""
        L = lib()
        dp = C.POINTER(C.c_double)
        L.b2o_momentum_jacobian.argtypes = [C.POINTER(Model), dp, C.c_double, dp, dp, dp, dp]
        q, bmc, bio = self._v(q), np.array(base_mc, float), np.array(base_Io, float)
        J, locked = np.zeros(6 * self.nb), np.zeros(10)
        L.b2o_momentum_jacobian(C.byref(self.m), _dp(q), float(base_mass), _dp(bmc), _dp(bio), _dp(J), _dp(locked))
        return J.reshape(6, self.nb), locked

    def energy(self, q, dq):
        q, dq = self._v(q), self._v(dq)
        return lib().b2o_energy(C.byref(self.m), _dp(q), _dp(dq))


class Sim:
    """Single-world simulator with ScenarI/O bookkeeping semantics (joint indices, not names)."""

    def __init__(self, model, step_size=0.001, steps_per_run=1):
        self.m = model
        self.h = lib().b2o_sim_create(C.byref(model), step_size, steps_per_run)
        if not self.h:
            raise ValueError("invalid step size or steps per run")

    def __del__(self):
        if getattr(self, "h", None):
            lib().b2o_sim_destroy(self.h)
            self.h = None

    def run(self, paused=False):
        return bool(lib().b2o_sim_run(self.h, int(paused)))

    def time(self):
        return lib().b2o_sim_time(self.h)

    def load_computed_torque(self, kp, kd, gravity=(0, 0, -9.80665)):
        kp, kd, g = (np.ascontiguousarray(v, float) for v in (kp, kd, gravity))
        return bool(lib().b2o_sim_load_computed_torque(self.h, _dp(kp), _dp(kd), _dp(g)))

    def apply_link_wrench(self, body, point, wrench, duration):
        p, w = np.ascontiguousarray(point, float), np.ascontiguousarray(wrench, float)
        return bool(lib().b2o_sim_apply_link_wrench(self.h, int(body), _dp(p), _dp(w), float(duration)))

    def __getattr__(self, name):
        fn = getattr(lib(), "b2o_sim_" + name)
        if name in ("force_target", "position_target", "velocity_target"):
            def getter(j):
                has = C.c_int(0)
                v = fn(self.h, j, C.byref(has))
                if not has.value:
                    raise RuntimeError("target not set")
                return v
            return getter
        return lambda *a: fn(self.h, *a)


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().b2o_philox4x32_10(c, k, o)
    return list(o)


def task_nq(task):
    return 1 if task == TASK_PENDULUM_SWINGUP else 2


def task_nobs(task):
    return 3 if task == TASK_PENDULUM_SWINGUP else 4


def sample_reset(task, seed, env, step):
    st = np.zeros(2 * task_nq(task))
    lib().b2o_task_sample_reset(task, seed, env, step, _dp(st))
    return st


def reset_from_uniforms(task, u):
    u4 = np.zeros(4)
    u4[:len(u)] = u
    st = np.zeros(2 * task_nq(task))
    lib().b2o_task_reset_from_uniforms(task, _dp(u4), _dp(st))
    return st


def action_force(task, action):
    j = C.c_int(0)
    return lib().b2o_task_action_force(task, float(action), C.byref(j)), j.value


def sample_reset_batch(task, seed, env_offset, n_envs, step):
    out = np.zeros((n_envs, 2 * task_nq(task)))
    for e in range(n_envs):
        out[e] = sample_reset(task, seed, env_offset + e, step)
    return out


def task_evaluate(task, state, tau_after=0.0):
    st = np.ascontiguousarray(state, float)
    obs = np.zeros(8)
    rew = C.c_double(0)
    d = lib().b2o_task_evaluate(task, _dp(st), tau_after, _dp(obs), C.byref(rew))
    return obs[:task_nobs(task)].copy(), rew.value, bool(d)


def rollout(model, task, actions, state, elapsed, dt=0.001, max_episode_steps=5000, seed=0,
            env_offset=0, first_step=1, record=True, steps_per_run=1):
    """actions[T, n]; state[n, 2nq] and elapsed[n] are updated in place."""
    actions = np.ascontiguousarray(actions, float)
    T, n = actions.shape
    assert state.flags.c_contiguous and state.dtype == np.float64 and state.shape[0] == n
    assert elapsed.dtype == np.int32
    nobs = task_nobs(task)
    if record:
        obs = np.zeros((T, n, nobs))
        rew = np.zeros((T, n))
        done = np.zeros((T, n), np.uint8)
        po, pr, pd = _dp(obs), _dp(rew), done.ctypes.data_as(C.POINTER(C.c_uint8))
    else:
        obs = rew = done = None
        po = pr = pd = None
    lib().b2o_rollout(C.byref(model), task, dt, steps_per_run, max_episode_steps, seed, env_offset, first_step, n, T,
                      _dp(actions), _dp(state), elapsed.ctypes.data_as(C.POINTER(C.c_int32)),
                      po, pr, pd)
    return obs, rew, done


# ---- free rigid bodies with contacts ----------------------------------------------------------------
SHAPE_BOX, SHAPE_SPHERE, SHAPE_CYLINDER, SHAPE_PLANE = 0, 1, 2, 3


class Shape(C.Structure):
    _fields_ = [("type", C.c_int32), ("size", C.c_double * 3), ("R", C.c_double * 9), ("p", C.c_double * 3),
                ("mu", C.c_double)]


class FreeBody(C.Structure):
    _fields_ = [("mass", C.c_double), ("Ic", C.c_double * 9), ("com", C.c_double * 3), ("nshapes", C.c_int32),
                ("shape", Shape * 2)]


class World(C.Structure):
    _fields_ = [("nfree", C.c_int32), ("nstatic", C.c_int32), ("iterations", C.c_int32), ("dt", C.c_double),
                ("erp", C.c_double), ("max_erv", C.c_double), ("g", C.c_double * 3), ("body", FreeBody * 8),
                ("stat", Shape * 16), ("ext", (C.c_double * 6) * 8)]


class ContactRec(C.Structure):
    _fields_ = [("a", C.c_int32), ("b", C.c_int32), ("pos", C.c_double * 3), ("n", C.c_double * 3),
                ("depth", C.c_double), ("force", C.c_double * 3)]


#: contact solver constants shared by the oracle and the engine (DART's ContactConstraint defaults for the error
#: reduction: ERP 0.01, maximum error-reduction velocity 1e-3 m/s... see DESIGN.md)
CONTACT_DEFAULTS = dict(iterations=50, erp=0.01, max_erv=1e-3)


def make_shape(type, size, R=np.eye(3), p=(0, 0, 0), mu=1.0):
    s = Shape()
    s.type = type
    s.size[:] = [float(v) for v in size]
    s.R[:] = np.asarray(R, float).ravel().tolist()
    s.p[:] = [float(v) for v in p]
    s.mu = mu
    return s


def make_box_body(mass, extents, inertia=None, com=(0, 0, 0), mu=1.0):
    b = FreeBody()
    b.mass = mass
    ex = np.asarray(extents, float)
    if inertia is None:
        inertia = np.diag([mass / 12 * (ex[1] ** 2 + ex[2] ** 2), mass / 12 * (ex[0] ** 2 + ex[2] ** 2),
                           mass / 12 * (ex[0] ** 2 + ex[1] ** 2)])
    b.Ic[:] = np.asarray(inertia, float).ravel().tolist()
    b.com[:] = list(com)
    b.nshapes = 1
    b.shape[0] = make_shape(SHAPE_BOX, ex / 2, mu=mu)
    return b


def make_world(bodies, statics, dt=0.001, gravity=(0, 0, -9.8), **kw):
    w = World()
    cfg = dict(CONTACT_DEFAULTS)
    cfg.update(kw)
    w.nfree, w.nstatic = len(bodies), len(statics)
    w.iterations, w.dt, w.erp, w.max_erv = cfg["iterations"], dt, cfg["erp"], cfg["max_erv"]
    w.g[:] = list(gravity)
    for i, b in enumerate(bodies):
        w.body[i] = b
    for i, s in enumerate(statics):
        w.stat[i] = s
    return w


def ground_plane(mu=1.0):
    return make_shape(SHAPE_PLANE, (0, 0, 1), mu=mu)


def world_step(world, X):
    """X: [nfree, 13] updated in place. Returns the list of contact records of the step."""
    L = lib()
    L.b2o_world_step.argtypes = [C.POINTER(World), C.POINTER(C.c_double), C.POINTER(ContactRec), C.c_int]
    assert X.flags.c_contiguous and X.dtype == np.float64
    out = (ContactRec * 32)()
    n = L.b2o_world_step(C.byref(world), _dp(X), out, 32)
    return [dict(a=out[k].a, b=out[k].b, pos=np.array(out[k].pos), n=np.array(out[k].n), depth=out[k].depth,
                 force=np.array(out[k].force)) for k in range(min(n, 32))]


# ---- per-env domain randomisation ---------------------------------------------------------------------
def sample_rand_params(model, seed, env, step, mass_delta, gravity_sigma):
    nq = model.nb
    mass = np.array(model.mass[:], float)
    out = np.zeros(nq + 1)
    L = lib()
    L.b2o_sample_rand_params.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_double, C.c_double, C.c_double,
                                         C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.b2o_sample_rand_params(seed, env, step, nq, mass_delta, gravity_sigma, model.gravity[2], _dp(mass), _dp(out))
    return out


def rollout_randomized(model, task, actions, state, elapsed, rand, mass_delta, gravity_sigma, dt=0.001,
                       max_episode_steps=5000, seed=0, env_offset=0, first_step=1, steps_per_run=1):
    """Like rollout(), every env with its own masses / gravity (rand[n, nq+1], updated in place at resets)."""
    actions = np.ascontiguousarray(actions, float)
    T, n = actions.shape
    nobs = task_nobs(task)
    obs, rew, done = np.zeros((T, n, nobs)), np.zeros((T, n)), np.zeros((T, n), np.uint8)
    L = lib()
    dp = C.POINTER(C.c_double)
    L.b2o_rollout_randomized.argtypes = [C.POINTER(Model), C.c_int, C.c_double, C.c_int, C.c_int, C.c_uint64, C.c_uint64,
                                         C.c_uint64, C.c_int, C.c_int, dp, dp, C.POINTER(C.c_int32), dp, C.c_double,
                                         C.c_double, dp, dp, C.POINTER(C.c_uint8)]
    L.b2o_rollout_randomized(C.byref(model), task, dt, steps_per_run, max_episode_steps, seed, env_offset, first_step, n,
                             T, _dp(actions), _dp(state), elapsed.ctypes.data_as(C.POINTER(C.c_int32)), _dp(rand),
                             mass_delta, gravity_sigma, _dp(obs), _dp(rew), done.ctypes.data_as(C.POINTER(C.c_uint8)))
    return obs, rew, done


# ---- coupled world: articulated model with link shapes + free bodies ------------------------------------
ROBOT_SIDE = -1000


class RobotShapes(C.Structure):
    _fields_ = [("nrobot", C.c_int32), ("rbody", C.c_int32 * 4), ("rshape", Shape * 4)]


def robot_shapes_from_tables(t):
    """Box / sphere shapes on the moving links of a flattened URDF (urdf_tables.flatten)."""
    rs = RobotShapes()
    k = 0
    for sh in t["shapes"]:
        if sh["body"] < 0 or k >= 4:
            continue
        size = np.asarray(sh["size"], float) / 2 if sh["type"] == "box" else sh["size"]
        rs.rbody[k] = int(sh["body"])
        rs.rshape[k] = make_shape(SHAPE_BOX if sh["type"] == "box" else SHAPE_SPHERE, size, R=sh["R"], p=sh["p"], mu=sh["mu"])
        k += 1
    rs.nrobot = k
    return rs


def make_box_static(extents, position, R=np.eye(3), mu=1.0):
    return make_shape(SHAPE_BOX, np.asarray(extents, float) / 2, R=R, p=position, mu=mu)


def _contact_list(out, n):
    return [dict(a=out[k].a, b=out[k].b, pos=np.array(out[k].pos), n=np.array(out[k].n), depth=out[k].depth,
                 force=np.array(out[k].force)) for k in range(max(0, min(n, 32)))]


def coupled_physics_step(model, world, rshapes, q, dq, tau, X):
    """One physics iteration of a coupled world on raw arrays: q, dq [nb] and X [nfree, 13] updated in place.
    Returns (ddq, contacts)."""
    L = lib()
    dp = C.POINTER(C.c_double)
    L.b2o_coupled_step.argtypes = [C.POINTER(World), C.POINTER(Model), C.POINTER(RobotShapes), dp, dp, C.c_void_p,
                                   C.c_void_p, dp, C.POINTER(ContactRec), C.c_int]
    nb = model.nb
    dt = world.dt
    qq, dd, tt, acc = (np.zeros(MAXB) for _ in range(4))
    qq[:nb], dd[:nb], tt[:nb] = q, dq, tau
    L.b2o_forward_dynamics(C.byref(model), dt, _dp(qq), _dp(dd), _dp(tt), _dp(acc))
    dd[:nb] += acc[:nb] * dt
    before = dd.copy()
    out = (ContactRec * 32)()
    n = L.b2o_coupled_step(C.byref(world), C.byref(model), C.byref(rshapes), _dp(qq), _dp(dd), None, None, _dp(X), out, 32)
    ddq = acc[:nb] + (dd[:nb] - before[:nb]) / dt
    qq[:nb] += dd[:nb] * dt
    q[:] = qq[:nb]
    dq[:] = dd[:nb]
    return ddq, _contact_list(out, n)


def sim_attach_world(sim, world, rshapes, X0):
    L = lib()
    L.b2o_sim_attach_world.argtypes = [C.c_void_p, C.POINTER(World), C.POINTER(RobotShapes), C.POINTER(C.c_double)]
    X0 = np.ascontiguousarray(X0, float)
    sim._nfree = world.nfree
    return bool(L.b2o_sim_attach_world(sim.h, C.byref(world), C.byref(rshapes), _dp(X0)))


def sim_world_state(sim):
    L = lib()
    L.b2o_sim_world_state.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    X = np.zeros((sim._nfree, 13))
    L.b2o_sim_world_state(sim.h, _dp(X))
    return X


def sim_contacts(sim):
    L = lib()
    L.b2o_sim_contacts.argtypes = [C.c_void_p, C.POINTER(ContactRec), C.c_int]
    out = (ContactRec * 32)()
    n = L.b2o_sim_contacts(sim.h, out, 32)
    return _contact_list(out, n)


def sim_reset_base(sim, body, values, velocity=False):
    L = lib()
    L.b2o_sim_reset_base.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]
    v = np.ascontiguousarray(values, float)
    return bool(L.b2o_sim_reset_base(sim.h, int(body), int(velocity), _dp(v)))
