"""Model loader (C++ host code of the engine) against the oracle's independent numpy URDF flattener."""
import numpy as np
import pytest

import b2sim
from b2sim import _lib


@pytest.mark.parametrize("name", ["pendulum", "cartpole", "panda"])
def test_tables_match_the_independent_loader(name, model_files, oracle):
    info = b2sim.ModelInfo.from_file(model_files[name])
    t = info.tables()
    ref, _ = oracle.load_urdf(model_files[name])
    assert info.joint_names == ref["joint_names"]
    assert info.link_names == ref["link_names"]
    nq = t["nq"]
    assert nq == ref["nb"]
    assert np.array_equal(t["parent"], ref["parent"][:nq])
    assert np.array_equal(t["jtype"] - 1, ref["jtype"][:nq])  # 2/3 here, 1/2 in the oracle
    for key in ("axis", "R", "p", "mass", "com", "Ic", "damping", "friction", "effort"):
        np.testing.assert_allclose(t[key], np.asarray(ref[key])[:nq], rtol=0, atol=1e-15, err_msg=key)
    for key in ("lower", "upper"):
        assert np.array_equal(t[key], np.asarray(ref[key])[:nq])
    nl = t["nlinks"]
    assert np.array_equal(t["link_body"], ref["link_body"][:nl])
    np.testing.assert_allclose(t["link_R"], ref["link_R"], atol=1e-15)
    np.testing.assert_allclose(t["link_p"], ref["link_p"], atol=1e-15)
    # links welded to the fixed base: mass, first moment and rotational inertia about the base origin
    np.testing.assert_allclose(t["base_mass"], ref["base_mass"], atol=1e-15)
    np.testing.assert_allclose(t["base_mc"], ref["base_mc"], atol=1e-15)
    np.testing.assert_allclose(t["base_Io"], ref["base_Io"], atol=1e-15)


def test_names_pinned_by_the_reference(model_files):
    """joint / link names the reference's tasks, tests and examples rely on."""
    cp = b2sim.ModelInfo.from_file(model_files["cartpole"])
    assert cp.name == "cartpole" and cp.joint_names == ["linear", "pivot"]
    pe = b2sim.ModelInfo.from_file(model_files["pendulum"])
    assert pe.name == "pendulum" and pe.joint_names == ["pivot"] and "support" in pe.link_names
    pa = b2sim.ModelInfo.from_file(model_files["panda"])
    assert pa.joint_names == [f"panda_joint{i}" for i in range(1, 8)] + ["panda_finger_joint1", "panda_finger_joint2"]
    for link in ("panda_link4", "panda_link7", "panda_leftfinger", "panda_rightfinger", "end_effector_frame"):
        assert link in pa.link_names
    gp = b2sim.ModelInfo.from_file(model_files["ground_plane"])
    assert gp.name == "ground_plane" and gp.link_names == ["link"] and gp.dofs == 0


def test_kind_classification(model_files):
    assert b2sim.ModelInfo.from_file(model_files["pendulum"]).kind == _lib.KIND_CHAIN1
    assert b2sim.ModelInfo.from_file(model_files["cartpole"]).kind == _lib.KIND_CHAIN_PR
    assert b2sim.ModelInfo.from_file(model_files["panda"]).kind == _lib.KIND_TREE
    assert b2sim.ModelInfo.from_file(model_files["ground_plane"]).kind == _lib.KIND_STATIC


def test_joint_names_skip_fixed_joints(model_files):
    """Model::jointNames skips 0-DoF joints (cpp/scenario/gazebo/src/Model.cpp:555-559)."""
    pa = b2sim.ModelInfo.from_file(model_files["panda"])
    assert "panda_joint8" not in pa.joint_names and "panda_hand_joint" not in pa.joint_names
    assert pa.dofs == 9


def test_malformed_models_are_rejected():
    for bad in ("", "<robot name='x'>", "<robot name='x'><link name='a'/><joint name='j' type='revolute'>"
                "<parent link='a'/><child link='zzz'/></joint></robot>", "<foo/>"):
        with pytest.raises(b2sim.B2Error):
            b2sim.ModelInfo.from_string(bad)


def test_sdf_subset_static_model():
    sdf = """<sdf version='1.7'><model name='table'><static>true</static>
      <link name='top'><pose>0 0 1.0 0 0 0</pose>
        <collision name='c'><geometry><box><size>1 2 0.1</size></box></geometry></collision></link>
      </model></sdf>"""
    info = b2sim.ModelInfo.from_string(sdf)
    assert info.name == "table" and info.dofs == 0 and info.link_names == ["top"]
    t = info.tables()
    np.testing.assert_allclose(t["link_p"][0], [0, 0, 1.0])


def test_sdf_joint_frames_match_urdf_semantics():
    """An SDF pendulum written with link poses + joint pose in the child frame flattens to the same
    tables as the equivalent URDF."""
    urdf = """<robot name='p'><link name='world'/><link name='base'/>
      <joint name='f' type='fixed'><parent link='world'/><child link='base'/><origin xyz='0 0 1'/></joint>
      <link name='arm'><inertial><origin xyz='0 0 0.25'/><mass value='1'/>
        <inertia ixx='0.02' iyy='0.02' izz='0.001' ixy='0' ixz='0' iyz='0'/></inertial></link>
      <joint name='pivot' type='revolute'><parent link='base'/><child link='arm'/><origin xyz='0 0.1 0.2' rpy='0.3 0 0'/>
        <axis xyz='1 0 0'/><limit lower='-1' upper='1' effort='10' velocity='5'/></joint></robot>"""
    sdf = """<sdf version='1.7'><model name='p'>
      <link name='base'><pose>0 0 1 0 0 0</pose></link>
      <joint name='f' type='fixed'><parent>world</parent><child>base</child></joint>
      <link name='arm'><pose>0 0.1 1.2 0.3 0 0</pose><inertial><pose>0 0 0.25 0 0 0</pose><mass>1</mass>
        <inertia><ixx>0.02</ixx><iyy>0.02</iyy><izz>0.001</izz></inertia></inertial></link>
      <joint name='pivot' type='revolute'><parent>base</parent><child>arm</child>
        <axis><xyz>1 0 0</xyz><limit><lower>-1</lower><upper>1</upper><effort>10</effort></limit></axis></joint>
      </model></sdf>"""
    a = b2sim.ModelInfo.from_string(urdf).tables()
    b = b2sim.ModelInfo.from_string(sdf).tables()
    for key in ("axis", "R", "p", "mass", "com", "Ic", "lower", "upper", "effort"):
        np.testing.assert_allclose(a[key], b[key], atol=1e-15, err_msg=key)
