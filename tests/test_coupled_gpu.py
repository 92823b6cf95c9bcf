"""Coupled worlds on the B200 (BASELINE config 5): the Panda's finger pads grasp a cube that rests on a table.
The pick-and-place flow of examples/panda_pick_and_place.py through the scenario API, and the coupled kernel
(k_world_coupled) against the oracle env by env."""
import numpy as np
import pytest

from test_coupled_cpu import BASE, CUBE_MASS, EDGE, Q0, pick_scene

pytestmark = pytest.mark.gpu

I = CUBE_MASS / 12 * 2 * EDGE ** 2
CUBE_URDF = f"""
    <robot name="cube_robot">
        <link name="cube">
            <inertial><origin rpy="0 0 0" xyz="0 0 0"/><mass value="{CUBE_MASS}"/>
              <inertia ixx="{I}" ixy="0" ixz="0" iyy="{I}" iyz="0" izz="{I}"/></inertial>
            <collision><geometry><box size="{EDGE} {EDGE} {EDGE}"/></geometry><origin rpy="0 0 0" xyz="0 0 0"/></collision>
        </link>
    </robot>"""
TABLE_SDF = """<?xml version="1.0"?>
<sdf version="1.7"><model name="table"><static>true</static><link name="top">
  <collision name="c"><geometry><box><size>0.4 0.4 0.05</size></box></geometry></collision>
</link></model></sdf>"""
KP, KD = [100.0] * 7 + [10000.0] * 2, [17.5] * 7 + [100.0] * 2   # examples/panda_pick_and_place.py:34-40


def test_pick_and_place_flow_through_scenario_api(model_files, oracle):
    """examples/panda_pick_and_place.py:213-385 without the IK solver: finger contact detection, computed-torque
    controller, grasp detected from the finger contact wrenches, lift, release."""
    from scenario import core
    from scenario import gazebo as scenario
    from gym_ignition.context.gazebo import controllers
    t, m, D, _, _, X0, table_centre = pick_scene(oracle, model_files)
    gazebo = scenario.GazeboSimulator(0.001, 1.0, 1)
    assert gazebo.initialize()
    world = gazebo.get_world().to_gazebo()
    assert world.set_physics_engine(scenario.PhysicsEngine_dart)
    assert world.insert_model(model_files["ground_plane"])
    assert world.insert_model(model_files["panda"], core.Pose(list(BASE), [1.0, 0, 0, 0]), "panda")
    panda = world.get_model("panda").to_gazebo()
    left, right = panda.get_link("panda_leftfinger").to_gazebo(), panda.get_link("panda_rightfinger").to_gazebo()
    assert left.enable_contact_detection(True) and right.enable_contact_detection(True)
    assert panda.reset_joint_positions(list(Q0))
    assert gazebo.run(paused=True)
    assert panda.set_controller_period(gazebo.step_size())
    for name in ("panda_finger_joint1", "panda_finger_joint2"):
        assert panda.get_joint(name).to_gazebo().set_max_generalized_force(500.0)
    assert panda.insert_model_plugin(*controllers.ComputedTorqueFixedBase(
        kp=KP, ki=[0.0] * 9, kd=KD, urdf=model_files["panda"], joints=panda.joint_names()).args())
    assert panda.set_joint_position_targets(panda.joint_positions())
    assert panda.set_joint_velocity_targets(panda.joint_velocities())
    assert panda.set_joint_acceleration_targets(panda.joint_accelerations())
    assert world.insert_model_from_string(TABLE_SDF, core.Pose(table_centre, [1.0, 0, 0, 0]), "table")
    assert world.insert_model_from_string(CUBE_URDF, core.Pose(list(X0[0, :3]), [1.0, 0, 0, 0]), "cube")
    cube = world.get_model("cube").to_gazebo()
    assert cube.enable_contacts(True)
    assert gazebo.run(paused=True)
    for _ in range(200):
        assert gazebo.run()
    assert not left.in_contact() and not right.in_contact()
    assert {c.body_b for c in cube.contacts()} == {"table::top"}
    assert cube.base_position()[2] == pytest.approx(X0[0, 2], abs=1e-3)
    # close the fingers until both report a grasp (panda_pick_and_place.py:316-326, threshold scaled to this gripper)
    f1, f2 = panda.get_joint("panda_finger_joint1"), panda.get_joint("panda_finger_joint2")
    assert f1.set_position_target(f1.position_limit().min) and f2.set_position_target(f2.position_limit().min)
    steps = 0
    while not (np.linalg.norm(left.contact_wrench()) >= 10.0 and np.linalg.norm(right.contact_wrench()) >= 10.0):
        assert gazebo.run()
        steps += 1
        assert steps < 1000, "no grasp detected"
    assert left.in_contact() and right.in_contact()
    assert left.contacts()[0].body_a == "panda::panda_leftfinger" and left.contacts()[0].body_b == "cube::cube"
    assert {c.body_b for c in cube.contacts()} >= {"panda::panda_leftfinger", "panda::panda_rightfinger"}
    # Newton's third law in the report: what the finger feels is minus what the cube feels from that finger
    on_cube = [c for c in cube.contacts() if c.body_b == "panda::panda_leftfinger"][0]
    f_cube = np.sum([p.force for p in on_cube.points], axis=0)
    assert np.allclose(left.contact_wrench()[:3], -f_cube, atol=1e-9)
    # lift by 6 cm
    q = np.array(panda.joint_positions())
    J = D.point_jacobian(q, 6)[:3, :7]
    dq = np.linalg.pinv(J) @ np.array([0.0, 0.0, 0.06])
    arm = [f"panda_joint{k}" for k in range(1, 8)]
    assert panda.set_joint_position_targets(list(q[:7] + dq), arm)
    for _ in range(1500):
        assert gazebo.run()
    assert cube.base_position()[2] - X0[0, 2] > 0.04 and abs(cube.base_position()[1]) < 5e-3
    assert "table::top" not in {c.body_b for c in cube.contacts()}
    # open: the cube falls back onto the table
    assert f1.set_position_target(f1.position_limit().max) and f2.set_position_target(f2.position_limit().max)
    for _ in range(600):
        assert gazebo.run()
    assert not left.in_contact() and not right.in_contact()
    assert cube.base_position()[2] == pytest.approx(X0[0, 2], abs=3e-3)
    gazebo.close()


def test_coupled_kernel_matches_oracle(oracle, model_files):
    """32 worlds (Panda + table + cube at a per-env offset) under the computed-torque controller: open, close on the
    cube, lift. Joint and cube states against the oracle's single-world simulator, env by env."""
    import torch
    import b2sim
    from b2sim import _lib
    n = 32
    t, m, D, world, rs, X0, table_centre = pick_scene(oracle, model_files)
    m.effort[7] = m.effort[8] = 500.0
    sim = b2sim.Simulator(n, 0.001, 1)
    sim.insert_model_file(model_files["ground_plane"])
    panda = sim.insert_model_file(model_files["panda"], pose=list(BASE) + [1.0, 0, 0, 0], name="panda")
    sim.insert_model(TABLE_SDF, pose=table_centre + [1.0, 0, 0, 0], name="table")
    cube = sim.insert_model(CUBE_URDF, pose=list(X0[0, :3]) + [1.0, 0, 0, 0], name="cube")
    rng = np.random.default_rng(5)
    Xc = np.tile(X0[0], (n, 1))
    Xc[:, 0] += rng.uniform(-0.004, 0.004, n)
    Xc[:, 1] += rng.uniform(-0.01, 0.01, n)
    sim.tensor(cube, _lib.BUF_BASE_STATE).copy_(torch.as_tensor(Xc, device="cuda"))
    state = sim.tensor(panda, _lib.BUF_STATE)
    state[:, :9] = torch.as_tensor(Q0, device="cuda")
    for j in (7, 8):
        check = sim.lib.b2sim_set_max_generalized_force(sim.handle, panda, j, 500.0)
        assert check == 0
    sim.set_controller_period(panda, 0.001)
    sim.set_computed_torque(panda, KP, KD)
    pos_t = sim.tensor(panda, _lib.BUF_POS_TARGET)
    for j in range(9):
        sim.set_joint(panda, _lib.FIELD_POSITION_TARGET, -1, j, Q0[j])
        sim.set_joint(panda, _lib.FIELD_VELOCITY_TARGET, -1, j, 0.0)
        sim.set_joint(panda, _lib.FIELD_ACCELERATION_TARGET, -1, j, 0.0)
    refs = []
    for e in range(n):
        ref = oracle.Sim(m, 0.001, 1)
        assert oracle.sim_attach_world(ref, world, rs, Xc[e:e + 1])
        ref.set_controller_period(0.001)
        for j in range(9):
            ref.reset_position(j, Q0[j])
        ref.run(True)
        ref.load_computed_torque(KP, KD)
        for j in range(9):
            ref.set_position_target(j, Q0[j]); ref.set_velocity_target(j, 0.0); ref.set_acceleration_target(j, 0.0)
        refs.append(ref)
    lift = np.linalg.pinv(D.point_jacobian(np.r_[Q0[:7], 0.025, 0.025], 6)[:3, :7]) @ np.array([0.0, 0.0, 0.05])

    def compare(tag, tol_q, tol_x):
        got_q = state.cpu().numpy()
        got_x = sim.tensor(cube, _lib.BUF_BASE_STATE).cpu().numpy()
        want_q = np.array([[r.position(j) for j in range(9)] + [r.velocity(j) for j in range(9)] for r in refs])
        want_x = np.array([oracle.sim_world_state(r)[0] for r in refs])
        err_q = np.abs(got_q - want_q).max(axis=1)
        err_x = np.abs(got_x - want_x).max(axis=1)
        assert np.median(err_q) < tol_q and np.median(err_x) < tol_x, (tag, err_q, err_x)
        assert (err_q < 1e3 * tol_q).mean() > 0.9 and (err_x < 1e3 * tol_x).mean() > 0.9, (tag, err_q, err_x)
        return got_x

    for phase, steps in (("open", 100), ("close", 300), ("lift", 600)):
        if phase == "close":
            pos_t[:, 7:] = 0.0
            for r in refs:
                r.set_position_target(7, 0.0); r.set_position_target(8, 0.0)
        if phase == "lift":
            pos_t[:, :7] += torch.as_tensor(lift, device="cuda")
            for r in refs:
                for j in range(7):
                    r.set_position_target(j, Q0[j] + lift[j])
        for _ in range(steps):
            sim.run()
            for r in refs:
                r.run(False)
        got_x = compare(phase, 1e-8, 1e-8)
        counts = np.array([len(sim.contacts(e)) for e in range(n)])
        want_counts = np.array([len(oracle.sim_contacts(r)) for r in refs])
        assert (counts == want_counts).mean() > 0.9, (phase, counts, want_counts)
    assert (got_x[:, 2] - X0[0, 2] > 0.03).all()       # every env lifted its cube
    sim.close()


def test_split_prepare_equals_single_kernel_prepare(monkeypatch):
    """The three concurrent prepare kernels (k_coupled_dynamics | k_coupled_rows | k_coupled_minv on forked streams)
    against the single k_coupled_prepare they replace, same solver: open gripper, closing, grasp, lift with a pending
    joint reset in the middle. They run the same functions on the same inputs, so every state bit must agree."""
    import torch
    import os
    import subprocess
    import sys
    if os.environ.get("B2_RUN_KERNEL") != "thread":
        # bit equality needs the same dynamics arithmetic on both sides: the split pipeline's default for this batch size
        # is the lane kernel (CRBA + LDL^T), the single prepare kernel runs the articulated-body recursion. The lane
        # pipeline is compared with the oracle in test_coupled_kernel_matches_oracle; here the thread kernels are forced
        # (B2_RUN_KERNEL is read once per process).
        r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", os.path.abspath(__file__), "-k",
                            "test_split_prepare_equals_single_kernel_prepare"], env=dict(os.environ, B2_RUN_KERNEL="thread"),
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
        return
    import b2sim
    from b2sim import _lib
    n = 257
    scenes = []
    for split in ("1", "0"):
        monkeypatch.setenv("B2_COUPLED_SPLIT", split)
        sc = b2sim.PandaPickScene(n, seed=5)
        sc.step(30)
        sc.set_fingers(0.0)
        sc.step(150)
        # a pending position / velocity reset of one finger in a few envs is consumed by the next run
        for env in (3, 100):
            sc.sim.set_joint(sc.panda, _lib.FIELD_POSITION_RESET, env, 7, 0.03)
            sc.sim.set_joint(sc.panda, _lib.FIELD_VELOCITY_RESET, env, 7, 0.0)
        sc.step(20)
        sc.targets[:, 3] -= 0.1
        sc.step(100)
        torch.cuda.synchronize()
        scenes.append(sc)
    a, b = scenes
    assert torch.equal(a.state, b.state) and torch.equal(a.cube_state, b.cube_state)
    acc = [s.sim.tensor(s.panda, _lib.BUF_ACCELERATION) for s in scenes]
    assert torch.equal(acc[0], acc[1])
    assert len(a.sim.contacts(0)) == len(b.sim.contacts(0)) > 0
    assert (a.cube_state[:, 2] > 1.4).all()  # nothing fell through the table or flew away
    for sc in scenes:
        sc.close()


def test_coupled_world_with_two_physics_iterations_per_run():
    """steps_per_run = 2 (GazeboSimulator::run advances two iterations, GazeboSimulator.cpp:202-251) against twice as
    many runs of one iteration: the coupled pipeline is launched once per iteration, the controller recomputes its
    torque at its 1 kHz period in both, so the states agree bit for bit."""
    import torch
    import b2sim
    n = 130
    one = b2sim.PandaPickScene(n, seed=2)
    two = b2sim.PandaPickScene(n, seed=2, steps_per_run=2)
    one.step(40); two.step(20)
    one.set_fingers(0.0); two.set_fingers(0.0)
    one.step(200); two.step(100)
    one.targets[:, 3] += 0.1; two.targets[:, 3] += 0.1
    one.step(120); two.step(60)
    torch.cuda.synchronize()
    assert one.sim.time() == pytest.approx(two.sim.time(), abs=1e-12)
    assert torch.equal(one.state, two.state) and torch.equal(one.cube_state, two.cube_state)
    assert (one.cube_state[:, 2] > 1.5).all()   # lifted
    one.close(); two.close()
