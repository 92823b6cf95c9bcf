"""Free bodies and contacts on the B200: the reference's contact tests through the scenario API
(tests/test_scenario/test_contacts.py:58-236) and the world kernel against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MASS, EDGE = 5.0, 0.2
I = 1 / 12 * MASS * (EDGE ** 2 + EDGE ** 2)
CUBE_URDF = f"""
    <robot name="cube_robot" xmlns:xacro="http://www.ros.org/wiki/xacro">
        <link name="cube">
            <inertial>
              <origin rpy="0 0 0" xyz="0 0 0"/>
              <mass value="{MASS}"/>
              <inertia ixx="{I}" ixy="0" ixz="0" iyy="{I}" iyz="0" izz="{I}"/>
            </inertial>
            <visual><geometry><box size="{EDGE} {EDGE} {EDGE}"/></geometry><origin rpy="0 0 0" xyz="0 0 0"/></visual>
            <collision><geometry><box size="{EDGE} {EDGE} {EDGE}"/></geometry><origin rpy="0 0 0" xyz="0 0 0"/></collision>
        </link>
    </robot>"""


@pytest.fixture()
def world_with_ground(model_files):
    from scenario import gazebo as scenario
    gazebo = scenario.GazeboSimulator(0.001, 1.0, 1)
    assert gazebo.initialize()
    world = gazebo.get_world().to_gazebo()
    assert world.set_physics_engine(scenario.PhysicsEngine_dart)
    assert world.insert_model(model_files["ground_plane"])
    yield gazebo, world
    gazebo.close()


def test_cube_contact(world_with_ground):
    """tests/test_scenario/test_contacts.py:63-122."""
    from scenario import core
    gazebo, world = world_with_ground
    assert world.insert_model_from_string(CUBE_URDF, core.Pose([0, 0, 0.15], [1., 0, 0, 0]), "cube")
    assert len(world.model_names()) == 2
    cube = world.get_model("cube")
    assert cube.dofs() == 0 and cube.joint_names() == () and cube.total_mass() == MASS
    assert not cube.contacts_enabled()
    assert cube.enable_contacts(enable=True)
    assert cube.contacts_enabled()
    gazebo.run(paused=True)
    assert cube.base_position() == pytest.approx([0, 0, 0.15])
    assert not cube.get_link("cube").in_contact()
    assert len(cube.contacts()) == 0
    for _ in range(150):
        gazebo.run()
    assert cube.get_link("cube").in_contact()
    assert len(cube.contacts()) == 1
    contact = cube.contacts()[0]
    assert contact.body_a == "cube::cube" and contact.body_b == "ground_plane::link"
    for point in contact.points:
        assert point.normal == pytest.approx([0, 0, 1])
    z_forces = [point.force[2] for point in contact.points]
    assert np.sum(z_forces) == pytest.approx(-5 * world.gravity()[2], abs=0.1)
    assert cube.get_link("cube").contact_wrench() == pytest.approx([0, 0, np.sum(z_forces), 0, 0, 0], abs=1e-6)
    assert cube.links_in_contact() == ("cube",)
    assert cube.base_position()[2] == pytest.approx(EDGE / 2, abs=2e-3)


def test_cube_multiple_contacts(world_with_ground):
    """tests/test_scenario/test_contacts.py:130-236: two stacked cubes."""
    from scenario import core
    gazebo, world = world_with_ground
    assert world.insert_model_from_string(CUBE_URDF, core.Pose([0, 0, 0.15], [1., 0, 0, 0]), "cube1")
    assert world.insert_model_from_string(CUBE_URDF, core.Pose([0, 0, 0.4], [1., 0, 0, 0]), "cube2")
    cube1, cube2 = world.get_model("cube1"), world.get_model("cube2")
    assert cube1.enable_contacts(True) and cube2.enable_contacts(True)
    for _ in range(600):
        gazebo.run()
    contacts1, contacts2 = cube1.contacts(), cube2.contacts()
    assert len(contacts1) == 2 and len(contacts2) == 1
    by_other = {c.body_b: c for c in contacts1}
    assert set(by_other) == {"ground_plane::link", "cube2::cube"} and all(c.body_a == "cube1::cube" for c in contacts1)
    assert contacts2[0].body_a == "cube2::cube" and contacts2[0].body_b == "cube1::cube"
    ground_fz = sum(p.force[2] for p in by_other["ground_plane::link"].points)
    upper_on_lower = sum(p.force[2] for p in by_other["cube2::cube"].points)
    lower_on_upper = sum(p.force[2] for p in contacts2[0].points)
    assert ground_fz == pytest.approx(2 * MASS * 9.8, abs=1.1)
    assert lower_on_upper == pytest.approx(MASS * 9.8, abs=1.1)
    assert upper_on_lower == pytest.approx(-lower_on_upper, abs=1e-6)
    for p in contacts2[0].points:
        assert p.normal == pytest.approx([0, 0, 1], abs=1e-3)
    for p in by_other["cube2::cube"].points:
        assert p.normal == pytest.approx([0, 0, -1], abs=1e-3)
    # net wrench on the lower cube: ground pushes up with 2 m g, the upper cube pushes down with m g
    assert cube1.get_link("cube").contact_wrench()[2] == pytest.approx(MASS * 9.8, abs=1.5)


def test_base_reset_is_deferred_and_velocity_reset_works(world_with_ground):
    """tests/test_scenario/test_model.py:117-214: base pose / velocity resets are applied by the next run."""
    from scenario import core
    gazebo, world = world_with_ground
    assert world.insert_model_from_string(CUBE_URDF, core.Pose([0, 0, 1.0], [1., 0, 0, 0]), "cube")
    cube = world.get_model("cube").to_gazebo()
    gazebo.run(paused=True)
    assert cube.reset_base_pose([1.0, -2.0, 3.0], [0.0, 1.0, 0.0, 0.0])
    assert cube.base_position() == pytest.approx([0, 0, 1.0])
    gazebo.run(paused=True)
    assert cube.base_position() == pytest.approx([1.0, -2.0, 3.0])
    assert cube.base_orientation() == pytest.approx([0.0, 1.0, 0.0, 0.0])
    assert cube.reset_base_world_velocity([0.5, 0.0, 2.0], [0.0, 0.0, 1.0])
    gazebo.run(paused=True)
    assert cube.base_world_linear_velocity() == pytest.approx([0.5, 0.0, 2.0])
    assert cube.base_world_angular_velocity() == pytest.approx([0.0, 0.0, 1.0])
    gazebo.run()
    assert cube.base_world_linear_velocity() == pytest.approx([0.5, 0.0, 2.0 - 9.8e-3], abs=1e-9)
    assert cube.base_position()[0] == pytest.approx(1.0 + 0.5e-3, abs=1e-9)
    panda_like_fixed = world.get_model("ground_plane")
    assert not panda_like_fixed.to_gazebo().reset_base_pose([0, 0, 1], [1, 0, 0, 0])


def test_world_kernel_matches_oracle(oracle, model_files):
    """Batched free-body worlds (two tumbling cubes + ground) against the oracle, env by env: states to 1e-8 and
    identical contact counts over 400 steps."""
    import torch
    import b2sim
    n, T = 64, 400
    sim = b2sim.Simulator(n, 0.001, 1)
    sim.insert_model_file(model_files["ground_plane"])
    a = sim.insert_model(CUBE_URDF, name="a")
    b = sim.insert_model(CUBE_URDF, name="b")
    rng = np.random.default_rng(0)
    X0 = np.zeros((n, 2, 13))
    X0[:, :, 3] = 1.0
    X0[:, 0, :3] = np.c_[rng.uniform(-0.02, 0.02, (n, 2)), rng.uniform(0.12, 0.3, n)]
    X0[:, 1, :3] = np.c_[rng.uniform(-0.02, 0.02, (n, 2)), rng.uniform(0.45, 0.7, n)]
    quat = rng.normal(size=(n, 2, 4)); quat /= np.linalg.norm(quat, axis=2, keepdims=True)
    X0[::2, :, 3:7] = quat[::2]                       # every other env starts with random orientations
    X0[:, :, 7:13] = rng.uniform(-1, 1, (n, 2, 6))
    sim.tensor(a, 14).copy_(torch.as_tensor(X0[:, 0], device="cuda"))
    sim.tensor(b, 14).copy_(torch.as_tensor(X0[:, 1], device="cuda"))
    world = oracle.make_world([oracle.make_box_body(MASS, [EDGE] * 3, inertia=np.eye(3) * I),
                               oracle.make_box_body(MASS, [EDGE] * 3, inertia=np.eye(3) * I)], [oracle.ground_plane()])
    ref = X0.copy()
    counts = np.zeros(n, int)
    for step in range(T):
        sim.run()
        for e in range(n):
            counts[e] = len(oracle.world_step(world, ref[e]))
    got = np.stack([sim.tensor(a, 14).cpu().numpy(), sim.tensor(b, 14).cpu().numpy()], axis=1)
    # impacts amplify rounding differences: compare tightly where the motion is regular, loosely elsewhere
    close = np.abs(got - ref).reshape(n, -1).max(axis=1)
    assert np.median(close) < 1e-8 and (close < 1e-5).mean() > 0.9
    gpu_counts = np.array([len(sim.contacts(e)) for e in range(n)])
    assert (gpu_counts == counts).mean() > 0.9
    assert got[:, :, 2].min() > 0.09            # nothing fell through the ground
    sim.close()


@pytest.mark.parametrize("cubes", [3, 4])
def test_four_cubes_many_contacts(cubes, oracle, model_files):
    """Three / four cubes settling flat on the ground: 12 / 16 contact points (three-row units) and 18 / 24 generalized
    velocities. Under the warp-cooperative solver (see the next test) three cubes run the 32-lane variant of the
    shared-memory form, four cubes the streaming path taken by worlds with more units than that form holds; states
    against the oracle env by env."""
    import torch
    import b2sim
    n, T = 32, 300
    sim = b2sim.Simulator(n, 0.001, 1)
    sim.insert_model_file(model_files["ground_plane"])
    ids = [sim.insert_model(CUBE_URDF, name=f"c{k}") for k in range(cubes)]
    rng = np.random.default_rng(3)
    X0 = np.zeros((n, cubes, 13))
    X0[:, :, 3] = 1.0
    for k in range(cubes):
        X0[:, k, :3] = np.c_[0.5 * k + rng.uniform(-0.02, 0.02, n), rng.uniform(-0.02, 0.02, n), rng.uniform(0.101, 0.13, n)]
    X0[:, :, 7:10] = rng.uniform(-0.2, 0.2, (n, cubes, 3))
    for k, mid in enumerate(ids):
        sim.tensor(mid, 14).copy_(torch.as_tensor(X0[:, k], device="cuda"))
    body = lambda: oracle.make_box_body(MASS, [EDGE] * 3, inertia=np.eye(3) * I)
    world = oracle.make_world([body() for _ in range(cubes)], [oracle.ground_plane()])
    ref = X0.copy()
    counts = np.zeros(n, int)
    for step in range(T):
        sim.run()
        for e in range(n):
            counts[e] = len(oracle.world_step(world, ref[e]))
    got = np.stack([sim.tensor(mid, 14).cpu().numpy() for mid in ids], axis=1)
    close = np.abs(got - ref).reshape(n, -1).max(axis=1)
    assert np.median(close) < 1e-7 and (close < 1e-4).mean() > 0.9, close
    gpu_counts = np.array([len(sim.contacts(e)) for e in range(n)])
    assert counts.max() == 4 * cubes and (gpu_counts == counts).mean() > 0.9
    assert got[:, :, 2].min() > 0.09
    sim.close()


@pytest.mark.parametrize("solver", ["warp", "thread"])
def test_both_contact_solvers_on_free_bodies(solver):
    """Free-body worlds pick the prepare / solve / finish pipeline up to 8,192 envs and the single-thread kernel
    (k_world_free) above; the tests of this file run at small env counts, so each selection is forced here for the
    same oracle comparisons, in a subprocess because the selection is read once per process."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, B2_CONTACT_SOLVER=solver)
    here = os.path.abspath(__file__)
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", here, "-k",
                        "test_world_kernel_matches_oracle or test_cube_multiple_contacts or test_base_reset or test_four_cubes "
                        "or test_free_body_link_accelerations or inserted_and_removed"],
                       env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("steps_per_run", [1, 3])
def test_free_body_wrench_matches_oracle(steps_per_run, oracle, model_files):
    """Link::applyWorldWrench on a free body (Physics.cpp:1483-1532): a cube resting on the ground is pushed and twisted
    for 25 ms in every env, a second cube only in env 5; both against the oracle's world step with the same wrench, which
    stops acting once the pre-step time has reached the expiry (helpers.h:300-345), also in the middle of a run."""
    import torch
    import b2sim
    n, dt = 8, 0.001
    sim = b2sim.Simulator(n, dt, steps_per_run)
    sim.insert_model_file(model_files["ground_plane"])
    a = sim.insert_model(CUBE_URDF, pose=(0, 0, 0.1, 1, 0, 0, 0), name="a")
    b = sim.insert_model(CUBE_URDF, pose=(1.0, 0, 0.1, 1, 0, 0, 0), name="b")
    world = oracle.make_world([oracle.make_box_body(MASS, [EDGE] * 3, inertia=np.eye(3) * I),
                               oracle.make_box_body(MASS, [EDGE] * 3, inertia=np.eye(3) * I)], [oracle.ground_plane()])
    X = np.zeros((n, 2, 13)); X[:, :, 3] = 1.0; X[:, 0, 2] = 0.1; X[:, 1, 0] = 1.0; X[:, 1, 2] = 0.1
    for _ in range(40 // steps_per_run):      # settle
        sim.run()
        for e in range(n):
            for _ in range(steps_per_run):
                oracle.world_step(world, X[e])
    wa, wb = [60.0, -4.0, 30.0, 0.3, 0.1, 0.8], [0.0, 25.0, 70.0, 0.0, 0.0, 0.0]   # a slides (mu (m g - 30) < 60), b lifts off
    duration = 0.025
    sim.apply_link_wrench(a, -1, 0, wa, duration)
    sim.apply_link_wrench(b, 5, 0, wb, duration)
    with pytest.raises(b2sim.B2Error):
        sim.apply_link_wrench(a, -1, 0, wa, duration)     # one wrench per free body at a time
    t, t_exp = 0.0, duration
    for run in range(45 // steps_per_run):
        sim.run()
        for it in range(steps_per_run):
            active = (run == 0 and it == 0) or t < t_exp - 1e-12
            for e in range(n):
                world.ext[0][:] = wa if active else [0.0] * 6
                world.ext[1][:] = wb if (active and e == 5) else [0.0] * 6
                oracle.world_step(world, X[e])
            t += dt
    got = np.stack([sim.tensor(a, 14).cpu().numpy(), sim.tensor(b, 14).cpu().numpy()], axis=1)
    np.testing.assert_allclose(got, X, rtol=1e-8, atol=1e-9)
    assert got[0, 0, 0] > 1e-3                                         # cube a was pushed along x
    assert np.allclose(got[0, 1, :3], [1.0, 0, 0.1], atol=1e-5)        # cube b was only pushed in env 5
    assert got[5, 1, 1] > 1e-3
    # the wrench is gone: a second application is accepted again
    sim.apply_link_wrench(a, -1, 0, wa, 0.0)
    sim.close()


def test_free_body_force_through_the_scenario_api(world_with_ground):
    """Link.apply_world_force on a free-floating cube: 100 N upwards on 5 kg for 50 ms lifts it by
    0.5 (F/m - g) t^2 within the first-order error of the integrator, then it falls back."""
    from scenario import core
    gazebo, world = world_with_ground
    assert world.insert_model_from_string(CUBE_URDF, core.Pose([0, 0, 0.1], [1., 0, 0, 0]), "cube")
    cube = world.get_model("cube")
    for _ in range(30):
        gazebo.run()
    z0 = cube.base_position()[2]
    assert cube.get_link("cube").apply_world_force([0, 0, 100.0], 0.05)
    for _ in range(50):
        gazebo.run()
    acc = 100.0 / MASS + world.gravity()[2]
    assert cube.base_position()[2] - z0 == pytest.approx(0.5 * acc * 0.05 ** 2, rel=0.06)
    assert cube.base_world_linear_velocity()[2] == pytest.approx(acc * 0.05, rel=0.03)
    for _ in range(10):
        gazebo.run()
    assert cube.base_world_linear_velocity()[2] < acc * 0.05      # only gravity acts now


def test_free_body_link_accelerations(world_with_ground):
    """Link::world{Linear,Angular}Acceleration of a free-floating body (Link.cpp:240-294, filled by Physics.cpp:2040-2079):
    gravity in free fall, the velocity change of the step over dt (contact impulses included) at the impact, zero at rest;
    an off-centre link point adds alpha x r + w x (w x r)."""
    from scenario import core
    gazebo, world = world_with_ground
    assert world.insert_model_from_string(CUBE_URDF, core.Pose([0, 0, 0.5], [1., 0, 0, 0]), "cube")
    cube = world.get_model("cube")
    assert cube.enable_contacts(enable=True)
    link = cube.get_link("cube")
    gazebo.run(paused=True)
    assert cube.to_gazebo().reset_base_world_angular_velocity([0.0, 0.0, 3.0])
    dt = gazebo.step_size()
    for _ in range(20):
        v0 = np.array(link.world_linear_velocity())
        gazebo.run()
        v1 = np.array(link.world_linear_velocity())
        a = np.array(link.world_linear_acceleration())
        np.testing.assert_allclose(a, world.gravity(), atol=1e-9)                     # free fall
        np.testing.assert_allclose(a, (v1 - v0) / dt, atol=1e-8)
        np.testing.assert_allclose(link.world_angular_acceleration(), [0, 0, 0], atol=1e-9)
        R = np.array(core_rotation(link.orientation()))
        np.testing.assert_allclose(link.body_linear_acceleration(), R.T @ a, atol=1e-9)
    seen_impact = False
    for _ in range(600):
        v0 = np.array(link.world_linear_velocity()); w0 = np.array(link.world_angular_velocity())
        gazebo.run()
        v1 = np.array(link.world_linear_velocity()); w1 = np.array(link.world_angular_velocity())
        a = np.array(link.world_linear_acceleration())
        np.testing.assert_allclose(a, (v1 - v0) / dt, atol=1e-7)
        np.testing.assert_allclose(link.world_angular_acceleration(), (w1 - w0) / dt, atol=1e-7)
        seen_impact = seen_impact or a[2] > 100.0       # the contact impulse stops the fall within one step
    assert seen_impact
    assert link.in_contact()
    np.testing.assert_allclose(link.world_linear_acceleration(), [0, 0, 0], atol=0.05)   # at rest on the ground


def core_rotation(quat_wxyz):
    w, x, y, z = quat_wxyz
    return [[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
            [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
            [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]]


def test_free_bodies_inserted_and_removed_while_the_world_runs(world_with_ground):
    """Models come and go during a simulation (examples/panda_pick_and_place.py:133-145 inserts a cube per episode, the
    randomizers remove and re-insert the model at every reset, World.cpp:169-177,431-453): a cube settles, a second one
    is inserted above it after 300 steps and lands on it, the first is removed (processed by the next run) and the
    second falls to the ground. The bodies that stay keep their state across the re-built world description."""
    from scenario import core
    gazebo, world = world_with_ground
    assert world.insert_model_from_string(CUBE_URDF, core.Pose([0, 0, 0.12], [1., 0, 0, 0]), "lower")
    lower = world.get_model("lower")
    assert lower.enable_contacts(True)
    for _ in range(300):
        gazebo.run()
    z_lower = lower.base_position()[2]
    assert z_lower == pytest.approx(EDGE / 2, abs=3e-3)
    t_before = world.time()
    assert world.insert_model_from_string(CUBE_URDF, core.Pose([0.01, 0, 0.45], [1., 0, 0, 0]), "upper")
    assert set(world.model_names()) == {"ground_plane", "lower", "upper"}
    upper = world.get_model("upper")
    assert upper.enable_contacts(True)
    gazebo.run(paused=True)
    assert world.time() == t_before                                   # a paused run does not advance time
    assert lower.base_position()[2] == pytest.approx(z_lower, abs=1e-12)   # the settled cube kept its state
    assert upper.base_position() == pytest.approx([0.01, 0, 0.45])
    for _ in range(500):
        gazebo.run()
    assert upper.base_position()[2] == pytest.approx(1.5 * EDGE, abs=6e-3)  # resting on the lower cube
    assert lower.base_position()[2] == pytest.approx(EDGE / 2, abs=4e-3)
    pairs = {(c.body_a, c.body_b) for c in upper.contacts()}
    assert ("upper::cube", "lower::cube") in pairs
    assert world.remove_model("lower")
    assert "lower" in world.model_names()                             # removal is deferred to the next run
    gazebo.run(paused=True)
    assert set(world.model_names()) == {"ground_plane", "upper"}
    for _ in range(400):
        gazebo.run()
    assert upper.base_position()[2] == pytest.approx(EDGE / 2, abs=4e-3)    # fell onto the ground
    pairs = {(c.body_a, c.body_b) for c in upper.contacts()}
    assert pairs == {("upper::cube", "ground_plane::link")}
