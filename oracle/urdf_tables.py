"""TEST INFRASTRUCTURE ONLY — independent URDF -> link/joint table flattening for the CPU oracle.

This is the oracle-side counterpart of the product's C++ loader (gym-ignition_b200/csrc/b2_model.cpp).
It is written separately, in numpy, so that a bug in the product loader shows up as a parity failure
instead of being shared by both sides. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
leg may import it.

Reference behaviour restated:
  * model insertion, cpp/scenario/gazebo/src/World.cpp:394-429 (URDF accepted through sdformat),
  * joint_names() skips 0-DoF joints, cpp/scenario/gazebo/src/Model.cpp:555-559,
  * the initial model pose is written verbatim, cpp/scenario/gazebo/src/World.cpp:169-177.
sdformat itself (URDF->SDF conversion, fixed-joint lumping) is a third-party dependency that is not in
the reference tree; the lumping below restates its documented behaviour: links joined by fixed joints
are merged into one rigid body, their frames are kept as fixed offsets.
"""
import xml.etree.ElementTree as ET

import numpy as np

MAXB = 16
JT_REVOLUTE = 1
JT_PRISMATIC = 2


def rpy_to_R(rpy):
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def quat_wxyz_to_R(q):
    w, x, y, z = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _floats(text, n):
    vals = [float(v) for v in text.split()]
    assert len(vals) == n, text
    return np.array(vals)


def _origin(elem):
    o = elem.find("origin") if elem is not None else None
    xyz = np.zeros(3)
    rpy = np.zeros(3)
    if o is not None:
        if o.get("xyz"):
            xyz = _floats(o.get("xyz"), 3)
        if o.get("rpy"):
            rpy = _floats(o.get("rpy"), 3)
    return rpy_to_R(rpy), xyz


def parse_urdf(xml_string):
    """Plain parse: returns (name, links, joints) with every quantity in its own URDF frame."""
    root = ET.fromstring(xml_string)
    assert root.tag == "robot", "only URDF is handled by the oracle loader"
    links = {}
    link_order = []
    for le in root.findall("link"):
        name = le.get("name")
        ine = le.find("inertial")
        mass, com, Ic = 0.0, np.zeros(3), np.zeros((3, 3))
        if ine is not None:
            Ri, com = _origin(ine)
            mass = float(ine.find("mass").get("value"))
            ie = ine.find("inertia")
            g = lambda k: float(ie.get(k, "0"))
            I0 = np.array([[g("ixx"), g("ixy"), g("ixz")],
                           [g("ixy"), g("iyy"), g("iyz")],
                           [g("ixz"), g("iyz"), g("izz")]])
            Ic = Ri @ I0 @ Ri.T
        shapes = []
        for ce in le.findall("collision"):
            Rc, pc = _origin(ce)
            geom = ce.find("geometry")
            if geom is None:
                continue
            if geom.find("box") is not None:
                shapes.append(dict(type="box", size=_floats(geom.find("box").get("size"), 3), R=Rc, p=pc))
            elif geom.find("sphere") is not None:
                shapes.append(dict(type="sphere", size=np.array([float(geom.find("sphere").get("radius")), 0, 0]), R=Rc, p=pc))
        links[name] = dict(mass=mass, com=com, Ic=Ic, shapes=shapes)
        link_order.append(name)
    joints = []
    for je in root.findall("joint"):
        R, p = _origin(je)
        ax = je.find("axis")
        axis = _floats(ax.get("xyz"), 3) if ax is not None else np.array([1.0, 0, 0])
        nrm = np.linalg.norm(axis)
        if nrm > 0:
            axis = axis / nrm
        lim = je.find("limit")
        dyn = je.find("dynamics")
        jtype = je.get("type")
        lower, upper = -np.inf, np.inf
        effort, velocity = np.inf, np.inf
        if lim is not None:
            if jtype != "continuous":
                lower = float(lim.get("lower", "0"))
                upper = float(lim.get("upper", "0"))
            effort = float(lim.get("effort", "inf"))
            velocity = float(lim.get("velocity", "inf"))
        damping = float(dyn.get("damping", "0")) if dyn is not None else 0.0
        friction = float(dyn.get("friction", "0")) if dyn is not None else 0.0
        joints.append(dict(name=je.get("name"), type=jtype,
                           parent=je.find("parent").get("link"), child=je.find("child").get("link"),
                           R=R, p=p, axis=axis, lower=lower, upper=upper, effort=effort,
                           velocity=velocity, damping=damping, friction=friction))
    return root.get("name"), links, link_order, joints


def flatten(xml_string, base_position=(0.0, 0.0, 0.0), base_orientation_wxyz=(1.0, 0.0, 0.0, 0.0),
            gravity=(0.0, 0.0, -9.8)):
    """URDF -> tables of a fixed-base tree of 1-DoF joints (bodies sorted parents-first)."""
    name, links, link_order, joints = parse_urdf(xml_string)
    children = {j["child"] for j in joints}
    roots = [l for l in link_order if l not in children]
    assert len(roots) == 1, "URDF must have a single root link"
    root = roots[0]
    assert root == "world", "oracle loader handles fixed-base models (root link 'world')"

    # link -> (body index, R, p): pose of the link frame in its body frame. body -1 is the fixed base.
    frame = {root: (-1, np.eye(3), np.zeros(3))}
    bodies = []          # dict per moving body
    pending = list(joints)
    progress = True
    while pending and progress:
        progress = False
        for j in list(pending):
            if j["parent"] not in frame:
                continue
            pb, pR, pp = frame[j["parent"]]
            R = pR @ j["R"]
            p = pR @ j["p"] + pp
            if j["type"] == "fixed":
                frame[j["child"]] = (pb, R, p)
                pending.remove(j)
                progress = True
                break  # restart the scan: keeps "first eligible joint in file order"
            if j["type"] in ("revolute", "continuous", "prismatic"):
                b = len(bodies)
                bodies.append(dict(joint=j, parent=pb, R=R, p=p, links=[]))
                frame[j["child"]] = (b, np.eye(3), np.zeros(3))
                pending.remove(j)
                progress = True
                break
            raise ValueError("unsupported joint type " + j["type"])
    assert not pending, "disconnected joints: " + str([j["name"] for j in pending])
    nb = len(bodies)
    assert nb <= MAXB

    # lump the inertia of every link into its body
    mass = np.zeros(MAXB)
    first = np.zeros((MAXB, 3))       # sum m*c
    for l in link_order:
        b, R, p = frame[l]
        if b < 0:
            continue
        m = links[l]["mass"]
        mass[b] += m
        first[b] += m * (R @ links[l]["com"] + p)
    com = np.zeros((MAXB, 3))
    for b in range(nb):
        if mass[b] > 0:
            com[b] = first[b] / mass[b]
    Ic = np.zeros((MAXB, 3, 3))
    for l in link_order:
        b, R, p = frame[l]
        if b < 0:
            continue
        m = links[l]["mass"]
        c = R @ links[l]["com"] + p - com[b]
        Ic[b] += R @ links[l]["Ic"] @ R.T + m * (np.dot(c, c) * np.eye(3) - np.outer(c, c))

    t = dict(name=name, nb=nb,
             parent=np.full(MAXB, -1, np.int32), jtype=np.zeros(MAXB, np.int32),
             axis=np.zeros((MAXB, 3)), R=np.tile(np.eye(3), (MAXB, 1, 1)), p=np.zeros((MAXB, 3)),
             mass=mass, com=com, Ic=Ic,
             damping=np.zeros(MAXB), friction=np.zeros(MAXB), stiffness=np.zeros(MAXB),
             rest=np.zeros(MAXB), lower=np.full(MAXB, -np.inf), upper=np.full(MAXB, np.inf),
             effort=np.full(MAXB, np.inf), vmax=np.full(MAXB, np.inf),
             gravity=np.array(gravity, float),
             base_R=quat_wxyz_to_R(base_orientation_wxyz), base_p=np.array(base_position, float))
    for b, bd in enumerate(bodies):
        j = bd["joint"]
        t["parent"][b] = bd["parent"]
        t["jtype"][b] = JT_PRISMATIC if j["type"] == "prismatic" else JT_REVOLUTE
        t["axis"][b] = j["axis"]
        t["R"][b] = bd["R"]
        t["p"][b] = bd["p"]
        for k in ("damping", "friction", "lower", "upper", "effort"):
            t[k][b] = j[k]
        t["vmax"][b] = j["velocity"]
    t["joint_names"] = [bd["joint"]["name"] for bd in bodies]
    t["link_names"] = [l for l in link_order if l != "world"]
    t["link_body"] = np.array([frame[l][0] for l in t["link_names"]], np.int32)
    t["link_R"] = np.array([frame[l][1] for l in t["link_names"]])
    t["link_p"] = np.array([frame[l][2] for l in t["link_names"]])
    t["link_mass"] = np.array([links[l]["mass"] for l in t["link_names"]])
    # links welded to the fixed base: mass and first moment in the base frame (they count in the centre of mass)
    t["base_mass"] = float(sum(links[l]["mass"] for l in link_order if frame[l][0] < 0))
    t["base_mc"] = sum((links[l]["mass"] * (frame[l][1] @ links[l]["com"] + frame[l][2]) for l in link_order if frame[l][0] < 0),
                       np.zeros(3))
    Io = np.zeros((3, 3))
    for l in link_order:
        if frame[l][0] < 0:
            c = frame[l][1] @ links[l]["com"] + frame[l][2]
            Io += frame[l][1] @ links[l]["Ic"] @ frame[l][1].T + links[l]["mass"] * (c @ c * np.eye(3) - np.outer(c, c))
    t["base_Io"] = np.array([Io[0, 0], Io[0, 1], Io[0, 2], Io[1, 1], Io[1, 2], Io[2, 2]])
    # box / sphere collision shapes, pose expressed in the frame of the body the link is attached to
    t["shapes"] = []
    for l in t["link_names"]:
        b, R, p = frame[l]
        for sh in links[l]["shapes"]:
            t["shapes"].append(dict(link=l, body=b, type=sh["type"], size=sh["size"], R=R @ sh["R"], p=R @ sh["p"] + p, mu=1.0))
    return t
