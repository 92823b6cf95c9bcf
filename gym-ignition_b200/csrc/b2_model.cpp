// URDF / SDF-subset loader. See b2_model.hpp.
//
// Flattening rules (restating what sdformat + ScenarI/O do for the reference, World.cpp:70-180):
//   * links joined by fixed joints are one rigid body (sdformat's URDF fixed-joint lumping); their
//     frames are kept as fixed offsets so Link getters keep working;
//   * every revolute / continuous / prismatic joint creates one body whose frame is the joint frame;
//   * bodies are ordered parents-first, ties in file order, which is also the joint serialisation of
//     Model::jointNames() (0-DoF joints are skipped, Model.cpp:555-559).
#include "b2_model.hpp"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>
#include <map>
#include <random>
#include <sstream>
#include <stdexcept>

#include "b2_xml.hpp"

namespace b2 {

static std::vector<double> numbers(const std::string& text)
{
    std::vector<double> out;
    std::istringstream ss(text);
    std::string tok;
    while (ss >> tok) {
        char* end = nullptr;
        double v = strtod(tok.c_str(), &end);
        if (end == tok.c_str() || *end != '\0') throw std::runtime_error("not a number: '" + tok + "'");
        out.push_back(v);
    }
    return out;
}
static std::vector<double> numbers_n(const char* text, size_t n, const char* what)
{
    if (!text) throw std::runtime_error(std::string("missing ") + what);
    auto v = numbers(text);
    if (v.size() != n) throw std::runtime_error(std::string("wrong number of values in ") + what);
    return v;
}

static M3<double> rot_rpy(double r, double p, double y)
{
    const double cr = cos(r), sr = sin(r), cp = cos(p), sp = sin(p), cy = cos(y), sy = sin(y);
    // Rz(y) Ry(p) Rx(r)
    return M3<double>{{cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr,
                       sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr,
                       -sp, cp * sr, cp * cr}};
}

Pose pose_from_xyz_rpy(const double xyz[3], const double rpy[3])
{
    Pose P;
    P.R = rot_rpy(rpy[0], rpy[1], rpy[2]);
    P.p = {xyz[0], xyz[1], xyz[2]};
    return P;
}

Pose pose_from_xyz_quat(const double q7[7])
{
    const double w = q7[3], x = q7[4], y = q7[5], z = q7[6];
    Pose P;
    P.R = M3<double>{{1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                      2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                      2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)}};
    P.p = {q7[0], q7[1], q7[2]};
    return P;
}

namespace {

struct LinkDesc {
    std::string name;
    double mass = 0;
    V3<double> com{0, 0, 0};                       // link frame
    M3<double> Ic{{0, 0, 0, 0, 0, 0, 0, 0, 0}};    // about the COM, link axes
    Pose in_model;                                 // link frame in the model frame at q = 0
    bool placed = false;
};
struct JointDesc {
    std::string name, type, parent, child;
    Pose urdf_origin;   // URDF: joint (= child) frame in the parent link frame
    Pose sdf_in_child;  // SDF: joint frame in the child link frame
    V3<double> axis{1, 0, 0};
    double lower = -std::numeric_limits<double>::infinity();
    double upper = std::numeric_limits<double>::infinity();
    double effort = std::numeric_limits<double>::infinity();
    double velocity = std::numeric_limits<double>::infinity();
    double damping = 0, friction = 0, stiffness = 0, rest = 0;
    Pose in_model;      // joint frame in the model frame at q = 0
};
struct Desc {
    std::string name;
    bool is_urdf = true;
    bool is_static = false;
    std::vector<LinkDesc> links;
    std::vector<JointDesc> joints;
    std::vector<CollisionShape> shapes;  // link index filled later by name
    std::vector<std::string> shape_link;
};

Pose origin_of(const XmlNode* e)
{
    double xyz[3] = {0, 0, 0}, rpy[3] = {0, 0, 0};
    const XmlNode* o = e ? e->child("origin") : nullptr;
    if (o) {
        if (o->attr("xyz")) { auto v = numbers_n(o->attr("xyz"), 3, "origin xyz"); xyz[0] = v[0]; xyz[1] = v[1]; xyz[2] = v[2]; }
        if (o->attr("rpy")) { auto v = numbers_n(o->attr("rpy"), 3, "origin rpy"); rpy[0] = v[0]; rpy[1] = v[1]; rpy[2] = v[2]; }
    }
    return pose_from_xyz_rpy(xyz, rpy);
}
Pose sdf_pose_of(const XmlNode* e)
{
    double xyz[3] = {0, 0, 0}, rpy[3] = {0, 0, 0};
    const XmlNode* p = e ? e->child("pose") : nullptr;
    if (p && !p->text.empty()) {
        auto v = numbers(p->text);
        if (v.size() != 6) throw std::runtime_error("<pose> needs 6 values");
        for (int i = 0; i < 3; ++i) { xyz[i] = v[i]; rpy[i] = v[i + 3]; }
    }
    return pose_from_xyz_rpy(xyz, rpy);
}
double attr_d(const XmlNode* e, const char* name, double def)
{
    if (!e || !e->attr(name)) return def;
    auto v = numbers(e->attr(name));
    if (v.size() != 1) throw std::runtime_error(std::string("bad attribute ") + name);
    return v[0];
}
double child_d(const XmlNode* e, const char* name, double def)
{
    const XmlNode* c = e ? e->child(name) : nullptr;
    if (!c || c->text.empty()) return def;
    auto v = numbers(c->text);
    if (v.size() != 1) throw std::runtime_error(std::string("bad element <") + name + ">");
    return v[0];
}

void parse_geometry(const XmlNode* geom, bool urdf, CollisionShape& s)
{
    if (!geom) throw std::runtime_error("collision without geometry");
    if (const XmlNode* b = geom->child("box")) {
        s.type = ShapeType::Box;
        auto v = urdf ? numbers_n(b->attr("size"), 3, "box size")
                      : numbers_n(b->child("size") ? b->child("size")->text.c_str() : nullptr, 3, "box size");
        for (int i = 0; i < 3; ++i) s.size[i] = v[i];
    } else if (const XmlNode* sp = geom->child("sphere")) {
        s.type = ShapeType::Sphere;
        s.size[0] = urdf ? attr_d(sp, "radius", 0) : child_d(sp, "radius", 0);
    } else if (const XmlNode* c = geom->child("cylinder")) {
        s.type = ShapeType::Cylinder;
        s.size[0] = urdf ? attr_d(c, "radius", 0) : child_d(c, "radius", 0);
        s.size[1] = urdf ? attr_d(c, "length", 0) : child_d(c, "length", 0);
    } else if (const XmlNode* pl = geom->child("plane")) {
        s.type = ShapeType::Plane;
        s.size[0] = 0; s.size[1] = 0; s.size[2] = 1;
        if (pl->child("normal")) {
            auto v = numbers_n(pl->child("normal")->text.c_str(), 3, "plane normal");
            for (int i = 0; i < 3; ++i) s.size[i] = v[i];
        }
    } else {
        // meshes and other shapes carry no collision in this engine
        s.type = ShapeType::Sphere;
        s.size[0] = 0;
    }
}

Desc parse_urdf(const XmlNode& robot)
{
    Desc d;
    d.is_urdf = true;
    d.name = robot.attr("name") ? robot.attr("name") : "";
    for (const XmlNode* le : robot.all("link")) {
        LinkDesc L;
        if (!le->attr("name")) throw std::runtime_error("link without a name");
        L.name = le->attr("name");
        if (const XmlNode* in = le->child("inertial")) {
            Pose io = origin_of(in);
            L.com = io.p;
            L.mass = attr_d(in->child("mass"), "value", 0);
            const XmlNode* I = in->child("inertia");
            const double xx = attr_d(I, "ixx", 0), xy = attr_d(I, "ixy", 0), xz = attr_d(I, "ixz", 0),
                         yy = attr_d(I, "iyy", 0), yz = attr_d(I, "iyz", 0), zz = attr_d(I, "izz", 0);
            M3<double> I0{{xx, xy, xz, xy, yy, yz, xz, yz, zz}};
            L.Ic = mulBt(mul(io.R, I0), io.R);
        }
        for (const XmlNode* ce : le->all("collision")) {
            CollisionShape s;
            s.name = ce->attr("name") ? ce->attr("name") : (L.name + "_collision");
            s.pose = origin_of(ce);
            parse_geometry(ce->child("geometry"), true, s);
            d.shapes.push_back(s);
            d.shape_link.push_back(L.name);
        }
        d.links.push_back(L);
    }
    for (const XmlNode* je : robot.all("joint")) {
        JointDesc J;
        if (!je->attr("name") || !je->attr("type")) throw std::runtime_error("joint without name/type");
        J.name = je->attr("name");
        J.type = je->attr("type");
        if (!je->child("parent") || !je->child("child")) throw std::runtime_error("joint " + J.name + " lacks parent/child");
        J.parent = je->child("parent")->attr("link") ? je->child("parent")->attr("link") : "";
        J.child = je->child("child")->attr("link") ? je->child("child")->attr("link") : "";
        J.urdf_origin = origin_of(je);
        if (const XmlNode* ax = je->child("axis")) {
            auto v = numbers_n(ax->attr("xyz"), 3, "axis xyz");
            J.axis = {v[0], v[1], v[2]};
        }
        if (const XmlNode* lim = je->child("limit")) {
            if (J.type != "continuous") {
                J.lower = attr_d(lim, "lower", 0);
                J.upper = attr_d(lim, "upper", 0);
            }
            J.effort = attr_d(lim, "effort", J.effort);
            J.velocity = attr_d(lim, "velocity", J.velocity);
        }
        if (const XmlNode* dyn = je->child("dynamics")) {
            J.damping = attr_d(dyn, "damping", 0);
            J.friction = attr_d(dyn, "friction", 0);
        }
        d.joints.push_back(J);
    }
    return d;
}

Desc parse_sdf(const XmlNode& sdf)
{
    const XmlNode* model = sdf.tag == "model" ? &sdf : sdf.child("model");
    if (!model) throw std::runtime_error("SDF without a <model>");
    Desc d;
    d.is_urdf = false;
    d.name = model->attr("name") ? model->attr("name") : "";
    if (const XmlNode* st = model->child("static")) d.is_static = (st->text == "true" || st->text == "1");
    for (const XmlNode* le : model->all("link")) {
        LinkDesc L;
        if (!le->attr("name")) throw std::runtime_error("link without a name");
        L.name = le->attr("name");
        L.in_model = sdf_pose_of(le);
        L.placed = true;
        if (const XmlNode* in = le->child("inertial")) {
            Pose io = sdf_pose_of(in);
            L.com = io.p;
            L.mass = child_d(in, "mass", 0);
            const XmlNode* I = in->child("inertia");
            const double xx = child_d(I, "ixx", 0), xy = child_d(I, "ixy", 0), xz = child_d(I, "ixz", 0),
                         yy = child_d(I, "iyy", 0), yz = child_d(I, "iyz", 0), zz = child_d(I, "izz", 0);
            M3<double> I0{{xx, xy, xz, xy, yy, yz, xz, yz, zz}};
            L.Ic = mulBt(mul(io.R, I0), io.R);
        }
        for (const XmlNode* ce : le->all("collision")) {
            CollisionShape s;
            s.name = ce->attr("name") ? ce->attr("name") : (L.name + "_collision");
            s.pose = sdf_pose_of(ce);
            parse_geometry(ce->child("geometry"), false, s);
            if (const XmlNode* sf = ce->child("surface"))
                if (const XmlNode* fr = sf->child("friction"))
                    if (const XmlNode* ode = fr->child("ode")) s.mu = child_d(ode, "mu", 1.0);
            d.shapes.push_back(s);
            d.shape_link.push_back(L.name);
        }
        d.links.push_back(L);
    }
    for (const XmlNode* je : model->all("joint")) {
        JointDesc J;
        if (!je->attr("name") || !je->attr("type")) throw std::runtime_error("joint without name/type");
        J.name = je->attr("name");
        J.type = je->attr("type");
        if (!je->child("parent") || !je->child("child")) throw std::runtime_error("joint " + J.name + " lacks parent/child");
        J.parent = je->child("parent")->text;
        J.child = je->child("child")->text;
        J.sdf_in_child = sdf_pose_of(je);
        if (const XmlNode* ax = je->child("axis")) {
            if (ax->child("xyz")) {
                auto v = numbers_n(ax->child("xyz")->text.c_str(), 3, "axis xyz");
                J.axis = {v[0], v[1], v[2]};
            }
            if (const XmlNode* lim = ax->child("limit")) {
                J.lower = child_d(lim, "lower", J.lower);
                J.upper = child_d(lim, "upper", J.upper);
                J.effort = child_d(lim, "effort", J.effort);
                J.velocity = child_d(lim, "velocity", J.velocity);
                if (J.effort < 0) J.effort = std::numeric_limits<double>::infinity();   // SDF: -1 = unlimited
                if (J.velocity < 0) J.velocity = std::numeric_limits<double>::infinity();
            }
            if (const XmlNode* dyn = ax->child("dynamics")) {
                J.damping = child_d(dyn, "damping", 0);
                J.friction = child_d(dyn, "friction", 0);
                J.stiffness = child_d(dyn, "spring_stiffness", 0);
                J.rest = child_d(dyn, "spring_reference", 0);
            }
        }
        if (J.type == "revolute" && J.lower <= -1e16 && J.upper >= 1e16) {
            J.lower = -std::numeric_limits<double>::infinity();
            J.upper = std::numeric_limits<double>::infinity();
        }
        d.joints.push_back(J);
    }
    return d;
}

int find_link(const Desc& d, const std::string& name)
{
    for (size_t i = 0; i < d.links.size(); ++i)
        if (d.links[i].name == name) return (int)i;
    return -1;
}

bool moving(const std::string& type) { return type == "revolute" || type == "continuous" || type == "prismatic"; }

}  // namespace

b2model* parse_model(const char* xml, size_t len)
{
    XmlParser parser(xml, len);
    std::unique_ptr<XmlNode> root = parser.parse();
    Desc d;
    if (root->tag == "robot") d = parse_urdf(*root);
    else if (root->tag == "sdf" || root->tag == "model") d = parse_sdf(*root);
    else throw std::runtime_error("unsupported root element <" + root->tag + ">");
    if (d.name.empty()) throw std::runtime_error("model without a name");
    if (d.links.empty()) throw std::runtime_error("model without links");

    // --- root link -------------------------------------------------------------------------------
    std::map<std::string, int> as_child;
    for (auto& j : d.joints) {
        if (find_link(d, j.child) < 0 || (j.parent != "world" && find_link(d, j.parent) < 0))
            throw std::runtime_error("joint " + j.name + " references an unknown link");
        if (as_child.count(j.child)) throw std::runtime_error("link " + j.child + " has two parent joints");
        as_child[j.child] = 1;
    }
    std::vector<int> roots;
    for (size_t i = 0; i < d.links.size(); ++i)
        if (!as_child.count(d.links[i].name)) roots.push_back((int)i);
    bool has_world = find_link(d, "world") >= 0;
    for (auto& j : d.joints)
        if (j.parent == "world") has_world = true;  // SDF: joints may name the implicit world frame

    // --- place every link / joint frame in the model frame at q = 0 ------------------------------
    if (d.is_urdf) {
        if (roots.size() != 1) throw std::runtime_error("URDF must have exactly one root link");
        d.links[roots[0]].placed = true;
        bool progress = true;
        size_t placed = 1;
        while (progress) {
            progress = false;
            for (auto& j : d.joints) {
                int pi = find_link(d, j.parent), ci = find_link(d, j.child);
                if (d.links[pi].placed && !d.links[ci].placed) {
                    d.links[ci].in_model = compose(d.links[pi].in_model, j.urdf_origin);
                    d.links[ci].placed = true;
                    j.in_model = d.links[ci].in_model;  // URDF: the joint frame is the child link frame
                    ++placed;
                    progress = true;
                }
            }
        }
        if (placed != d.links.size()) throw std::runtime_error("URDF has links that are not connected to the root");
    } else {
        for (auto& j : d.joints) j.in_model = compose(d.links[find_link(d, j.child)].in_model, j.sdf_in_child);
    }

    auto* m = new b2model();
    m->name = d.name;
    m->is_static = d.is_static;
    b2_model_tables& t = m->t;
    memset(&t, 0, sizeof t);

    // --- bodies ----------------------------------------------------------------------------------
    // body_of[link]: -1 base, >= 0 body index, -2 not assigned yet
    std::vector<int> body_of(d.links.size(), -2);
    std::vector<Pose> body_frame;  // body frame in the model frame at q = 0
    if (d.is_static) {
        for (auto& b : body_of) b = -1;
        m->fixed_base = true;
    } else if (has_world) {
        m->fixed_base = true;
        int w = find_link(d, "world");
        if (w >= 0) body_of[w] = -1;
    } else {
        m->fixed_base = false;
    }
    if (!m->fixed_base) {
        // free-floating model: supported when every link is welded to the root link (one rigid body)
        for (auto& j : d.joints)
            if (j.type != "fixed") {
                delete m;
                throw std::runtime_error("floating-base models with moving joints are not supported yet (model '" + d.name + "')");
            }
        const int rootl = roots.empty() ? 0 : roots[0];
        if (!d.is_urdf) {
            // SDF: poses are given in the model frame; re-express every frame relative to the root link
            const Pose root_inv = inverse(d.links[rootl].in_model);
            for (auto& l : d.links) l.in_model = compose(root_inv, l.in_model);
            for (auto& j : d.joints) j.in_model = compose(root_inv, j.in_model);
        }
        body_of[rootl] = -1;  // "base" = the floating body itself
    }
    std::vector<bool> used(d.joints.size(), false);
    auto body_of_name = [&](const std::string& n) { return n == "world" ? -1 : body_of[find_link(d, n)]; };
    for (;;) {
        bool found = false;
        for (size_t k = 0; k < d.joints.size(); ++k) {
            if (used[k]) continue;
            JointDesc& j = d.joints[k];
            int pb = body_of_name(j.parent);
            if (pb == -2) continue;
            int ci = find_link(d, j.child);
            if (j.type == "fixed") {
                body_of[ci] = pb;
            } else if (moving(j.type)) {
                int b = (int)body_frame.size();
                if (b >= B2_MAX_DOFS) { delete m; throw std::runtime_error("too many degrees of freedom"); }
                body_of[ci] = b;
                body_frame.push_back(j.in_model);
                const Pose parent_frame = pb >= 0 ? body_frame[pb] : Pose();
                const Pose X = compose(inverse(parent_frame), j.in_model);
                t.parent[b] = pb;
                t.jtype[b] = j.type == "prismatic" ? B2_JOINT_PRISMATIC : B2_JOINT_REVOLUTE;
                double n = sqrt(dot(j.axis, j.axis));
                V3<double> a = n > 0 ? (1.0 / n) * j.axis : j.axis;
                t.axis[b][0] = a.x; t.axis[b][1] = a.y; t.axis[b][2] = a.z;
                for (int i = 0; i < 9; ++i) t.R[b][i] = X.R.m[i];
                t.p[b][0] = X.p.x; t.p[b][1] = X.p.y; t.p[b][2] = X.p.z;
                t.damping[b] = j.damping; t.friction[b] = j.friction;
                t.stiffness[b] = j.stiffness; t.rest[b] = j.rest;
                t.lower[b] = j.lower; t.upper[b] = j.upper;
                t.effort[b] = j.effort; t.vmax[b] = j.velocity;
                m->joint_names.push_back(j.name);
            } else {
                delete m;
                throw std::runtime_error("unsupported joint type '" + j.type + "' (joint " + j.name + ")");
            }
            used[k] = true;
            found = true;
            break;  // restart: always take the first eligible joint in file order
        }
        if (!found) break;
    }
    for (size_t k = 0; k < d.joints.size(); ++k)
        if (!used[k]) { delete m; throw std::runtime_error("joint " + d.joints[k].name + " is not connected to the base"); }
    // SDF links that no joint reaches are attached to the base when the model is static
    for (size_t i = 0; i < d.links.size(); ++i)
        if (body_of[i] == -2) {
            if (d.is_static || d.links.size() == 1) body_of[i] = -1;
            else { delete m; throw std::runtime_error("link " + d.links[i].name + " is not connected to the base"); }
        }
    t.nq = (int)body_frame.size();
    t.fixed_base = 1;

    // --- links: fixed offsets + lumped inertia ----------------------------------------------------
    std::vector<V3<double>> first(t.nq, V3<double>{0, 0, 0});
    int nl = 0;
    std::vector<int> link_slot(d.links.size(), -1);
    for (size_t i = 0; i < d.links.size(); ++i) {
        if (d.links[i].name == "world") continue;
        if (nl >= B2_MAX_LINKS) { delete m; throw std::runtime_error("too many links"); }
        const int b = body_of[i];
        const Pose bf = b >= 0 ? body_frame[b] : Pose();
        const Pose off = compose(inverse(bf), d.links[i].in_model);
        t.link_body[nl] = b;
        for (int k = 0; k < 9; ++k) t.link_R[nl][k] = off.R.m[k];
        t.link_p[nl][0] = off.p.x; t.link_p[nl][1] = off.p.y; t.link_p[nl][2] = off.p.z;
        t.link_mass[nl] = d.links[i].mass;
        t.link_com[nl][0] = d.links[i].com.x; t.link_com[nl][1] = d.links[i].com.y; t.link_com[nl][2] = d.links[i].com.z;
        t.total_mass += d.links[i].mass;
        m->link_names.push_back(d.links[i].name);
        link_slot[i] = nl++;
        if (b >= 0) {
            t.mass[b] += d.links[i].mass;
            first[b] = first[b] + d.links[i].mass * (mul(off.R, d.links[i].com) + off.p);
        } else {
            const V3<double> mc = d.links[i].mass * (mul(off.R, d.links[i].com) + off.p);
            t.base_mass += d.links[i].mass;
            t.base_mc[0] += mc.x; t.base_mc[1] += mc.y; t.base_mc[2] += mc.z;
            // rotational inertia about the base origin: R Ic R^T + m (|c|^2 1 - c c^T)
            const V3<double> c = mul(off.R, d.links[i].com) + off.p;
            const M3<double> Irot = mulBt(mul(off.R, d.links[i].Ic), off.R);
            const double cc = dot(c, c), lm = d.links[i].mass;
            const int idx[6][2] = {{0, 0}, {0, 1}, {0, 2}, {1, 1}, {1, 2}, {2, 2}};
            for (int k = 0; k < 6; ++k) {
                const int r = idx[k][0], q = idx[k][1];
                t.base_Io[k] += Irot.m[3 * r + q] + lm * ((r == q ? cc : 0.0) - (&c.x)[r] * (&c.x)[q]);
            }
        }
    }
    t.nlinks = nl;
    for (int b = 0; b < t.nq; ++b) {
        V3<double> c = t.mass[b] > 0 ? (1.0 / t.mass[b]) * first[b] : V3<double>{0, 0, 0};
        t.com[b][0] = c.x; t.com[b][1] = c.y; t.com[b][2] = c.z;
    }
    for (size_t i = 0; i < d.links.size(); ++i) {
        const int b = body_of[i];
        if (b < 0 || link_slot[i] < 0) continue;
        const int l = link_slot[i];
        M3<double> R;
        for (int k = 0; k < 9; ++k) R.m[k] = t.link_R[l][k];
        const V3<double> off{t.link_p[l][0], t.link_p[l][1], t.link_p[l][2]};
        const V3<double> c = mul(R, d.links[i].com) + off - V3<double>{t.com[b][0], t.com[b][1], t.com[b][2]};
        const M3<double> Irot = mulBt(mul(R, d.links[i].Ic), R);
        const double cc = dot(c, c), mass = d.links[i].mass;
        const M3<double> par = outer(c, c);
        for (int r = 0; r < 3; ++r)
            for (int s = 0; s < 3; ++s)
                t.Ic[b][3 * r + s] += Irot.m[3 * r + s] + mass * ((r == s ? cc : 0.0) - par.m[3 * r + s]);
    }
    for (size_t k = 0; k < d.shapes.size(); ++k) {
        CollisionShape s = d.shapes[k];
        int li = find_link(d, d.shape_link[k]);
        s.link = li >= 0 ? link_slot[li] : -1;
        if (s.link >= 0) m->shapes.push_back(s);
    }
    // collision shapes
    t.nshapes = 0;
    for (const CollisionShape& sh : m->shapes) {
        if (t.nshapes >= B2_MAX_SHAPES) break;
        const int k = t.nshapes++;
        t.shape_type[k] = sh.type == ShapeType::Box ? B2_SHAPE_BOX : sh.type == ShapeType::Sphere ? B2_SHAPE_SPHERE
                          : sh.type == ShapeType::Cylinder ? B2_SHAPE_CYLINDER : B2_SHAPE_PLANE;
        t.shape_link[k] = sh.link;
        for (int a = 0; a < 3; ++a) { t.shape_size[k][a] = sh.size[a]; t.shape_p[k][a] = (&sh.pose.p.x)[a]; }
        for (int a = 0; a < 9; ++a) t.shape_R[k][a] = sh.pose.R.m[a];
        t.shape_mu[k] = sh.mu;
    }
    t.fixed_base = m->fixed_base ? 1 : 0;
    if (!m->fixed_base) {
        // lump every link into one rigid body expressed in the root-link frame
        V3<double> first{0, 0, 0};
        double mass = 0;
        for (int l = 0; l < t.nlinks; ++l) {
            M3<double> R;
            for (int k = 0; k < 9; ++k) R.m[k] = t.link_R[l][k];
            const V3<double> off{t.link_p[l][0], t.link_p[l][1], t.link_p[l][2]};
            const LinkDesc& L = d.links[find_link(d, m->link_names[l])];
            mass += L.mass;
            first = first + L.mass * (mul(R, L.com) + off);
        }
        const V3<double> c = mass > 0 ? (1.0 / mass) * first : V3<double>{0, 0, 0};
        t.body_mass = mass;
        t.body_com[0] = c.x; t.body_com[1] = c.y; t.body_com[2] = c.z;
        for (int l = 0; l < t.nlinks; ++l) {
            M3<double> R;
            for (int k = 0; k < 9; ++k) R.m[k] = t.link_R[l][k];
            const V3<double> off{t.link_p[l][0], t.link_p[l][1], t.link_p[l][2]};
            const LinkDesc& L = d.links[find_link(d, m->link_names[l])];
            const V3<double> r = mul(R, L.com) + off - c;
            const M3<double> Irot = mulBt(mul(R, L.Ic), R), par = outer(r, r);
            const double rr = dot(r, r);
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b)
                    t.body_Ic[3 * a + b] += Irot.m[3 * a + b] + L.mass * ((a == b ? rr : 0.0) - par.m[3 * a + b]);
        }
        t.kind = B2_KIND_FREE;
        return m;
    }
    t.kind = t.nq == 0 ? B2_KIND_STATIC : B2_KIND_TREE;
    return m;
}

}  // namespace b2

// ---------------------------------------------------------------------------------------------------
template <typename T>
void b2model::to_device_tables(const b2::Pose& base, const double g[3], b2::ModelDev<T>& o) const
{
    memset(&o, 0, sizeof o);
    o.nq = t.nq;
    o.nlinks = t.nlinks;
    for (int b = 0; b < t.nq; ++b) {
        o.parent[b] = t.parent[b];
        o.jtype[b] = t.jtype[b];
        for (int k = 0; k < 3; ++k) {
            o.axis[b][k] = (T)t.axis[b][k];
            o.p[b][k] = (T)t.p[b][k];
            o.mc[b][k] = (T)(t.mass[b] * t.com[b][k]);
        }
        for (int k = 0; k < 9; ++k) o.R[b][k] = (T)t.R[b][k];
        o.mass[b] = (T)t.mass[b];
        // inertia about the body origin: Ic + m (c.c 1 - c c^T)
        const double* c = t.com[b];
        const double cc = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
        for (int r = 0; r < 3; ++r)
            for (int s = 0; s < 3; ++s)
                o.Io[b][3 * r + s] = (T)(t.Ic[b][3 * r + s] + t.mass[b] * ((r == s ? cc : 0.0) - c[r] * c[s]));
        const int sym[6] = {0, 1, 2, 4, 5, 8};
        for (int k = 0; k < 6; ++k) o.Icom[b][k] = (T)t.Ic[b][sym[k]];
        for (int k = 0; k < 3; ++k) o.com[b][k] = (T)t.com[b][k];
        o.damping[b] = (T)t.damping[b];
        o.friction[b] = (T)t.friction[b];
        o.stiffness[b] = (T)t.stiffness[b];
        o.rest[b] = (T)t.rest[b];
        o.lower[b] = (T)t.lower[b];
        o.upper[b] = (T)t.upper[b];
        o.effort[b] = (T)t.effort[b];
    }
    for (int k = 0; k < 3; ++k) {
        o.g[k] = (T)g[k];
        o.basep[k] = (T)(&base.p.x)[k];
    }
    for (int k = 0; k < 9; ++k) o.baseR[k] = (T)base.R.m[k];
    o.base_mass = (T)t.base_mass;
    for (int k = 0; k < 3; ++k) o.base_mc[k] = (T)t.base_mc[k];
    for (int k = 0; k < 6; ++k) o.base_Io[k] = (T)t.base_Io[k];
    for (int l = 0; l < t.nlinks; ++l) {
        o.link_body[l] = t.link_body[l];
        for (int k = 0; k < 9; ++k) o.link_R[l][k] = (T)t.link_R[l][k];
        for (int k = 0; k < 3; ++k) o.link_p[l][k] = (T)t.link_p[l][k];
    }
}
template void b2model::to_device_tables<double>(const b2::Pose&, const double*, b2::ModelDev<double>&) const;
template void b2model::to_device_tables<float>(const b2::Pose&, const double*, b2::ModelDev<float>&) const;

namespace {
// Closed-form coefficient vector (m11, m22, A, B, G1, E, F) of a chain model given as device tables.
void chain_coef_vector(const b2::ModelDev<double>& md, int nq, double out[7])
{
    using namespace b2;
    const double zero[2] = {0, 0}, half_pi = 1.5707963267948966;
    double q0[2] = {0, 0}, q1[2] = {0, 0}, M0[4] = {0}, M1[4] = {0}, g0[2] = {0}, g1[2] = {0};
    q1[nq - 1] = half_pi;
    mass_matrix<double, 2>(md, q0, M0);
    mass_matrix<double, 2>(md, q1, M1);
    inverse_dynamics<double, 2>(md, q0, zero, zero, true, g0);
    inverse_dynamics<double, 2>(md, q1, zero, zero, true, g1);
    if (nq == 1) {
        out[0] = M0[0]; out[1] = 0; out[2] = 0; out[3] = 0; out[4] = 0; out[5] = g0[0];
        out[6] = md.jtype[0] == kRevolute ? g1[0] : 0.0;
    } else {
        out[0] = M0[0]; out[1] = M0[3]; out[2] = M0[1]; out[3] = M1[1]; out[4] = g0[0]; out[5] = g0[1]; out[6] = g1[1];
    }
}
}  // namespace

int b2model::fit(const b2::Pose& base, const double g[3], double dt)
{
    using namespace b2;
    if (!fixed_base) return t.kind = B2_KIND_FREE;
    if (t.nq == 0) return t.kind = B2_KIND_STATIC;
    t.kind = B2_KIND_TREE;
    bool plain = true;  // closed forms cover neither springs, Coulomb friction nor limits
    for (int b = 0; b < t.nq; ++b)
        if (t.stiffness[b] != 0.0 || t.friction[b] != 0.0) plain = false;
    if (!plain || t.nq > 2) return t.kind;

    ModelDev<double> md;
    to_device_tables<double>(base, g, md);
    const double zero[2] = {0, 0};
    const double half_pi = 1.5707963267948966;
    auto grav = [&](const double* q, double* out) { inverse_dynamics<double, 2>(md, q, zero, zero, true, out); };
    ChainCoef<double> c{};
    c.dt = dt;
    std::mt19937_64 rng(12345);
    std::uniform_real_distribution<double> U(-1.0, 1.0);

    if (t.nq == 1) {
        double q0[1] = {0}, q1[1] = {half_pi}, M[1], g0[1], g1[1];
        mass_matrix<double, 2>(md, q0, M);
        grav(q0, g0);
        grav(q1, g1);
        c.revolute = t.jtype[0] == B2_JOINT_REVOLUTE;
        c.m11 = M[0];
        c.d1 = t.damping[0];
        c.E = g0[0];
        c.F = c.revolute ? g1[0] : 0.0;
        for (int k = 0; k < 16; ++k) {
            double q[1] = {6.0 * U(rng)}, dq[1] = {5.0 * U(rng)}, tau[1] = {20.0 * U(rng)}, ref[1];
            forward_dynamics<double, 2>(md, dt, q, dq, tau, ref);
            double qq = q[0], dd = dq[0], acc;
            chain1_step(c, qq, dd, tau[0], acc);
            if (!(fabs(acc - ref[0]) <= 1e-10 * (1.0 + fabs(ref[0])))) return t.kind;
        }
        coef = c;
        fit_basis(md);
        return t.kind = B2_KIND_CHAIN1;
    }
    if (t.jtype[0] == B2_JOINT_PRISMATIC && t.jtype[1] == B2_JOINT_REVOLUTE && t.parent[1] == 0) {
        double q0[2] = {0, 0}, q1[2] = {0, half_pi}, M0[4], M1[4], g0[2], g1[2];
        mass_matrix<double, 2>(md, q0, M0);
        mass_matrix<double, 2>(md, q1, M1);
        grav(q0, g0);
        grav(q1, g1);
        c.m11 = M0[0];
        c.m22 = M0[3];
        c.A = M0[1];
        c.B = M1[1];
        c.G1 = g0[0];
        c.E = g0[1];
        c.F = g1[1];
        c.d1 = t.damping[0];
        c.d2 = t.damping[1];
        for (int k = 0; k < 32; ++k) {
            double q[2] = {3.0 * U(rng), 6.0 * U(rng)}, dq[2] = {5.0 * U(rng), 10.0 * U(rng)};
            double tau[2] = {100.0 * U(rng), 20.0 * U(rng)}, ref[2];
            forward_dynamics<double, 2>(md, dt, q, dq, tau, ref);
            double x = q[0], th = q[1], dx = dq[0], dth = dq[1], ax, ath;
            chain_pr_step(c, x, th, dx, dth, tau[0], tau[1], ax, ath);
            if (!(fabs(ax - ref[0]) <= 1e-10 * (1.0 + fabs(ref[0])) && fabs(ath - ref[1]) <= 1e-10 * (1.0 + fabs(ref[1]))))
                return t.kind;
        }
        coef = c;
        fit_basis(md);
        return t.kind = B2_KIND_CHAIN_PR;
    }
    return t.kind;
}

// d(coef)/d(mass_k): the coefficients of the same chain with a unit point mass at body k's centre of mass and
// nothing else (every coefficient is linear in each body mass at fixed COM and rotational inertia).
void b2model::fit_basis(const b2::ModelDev<double>& md)
{
    using namespace b2;
    const int nq = t.nq;
    for (int k = 0; k < nq; ++k) {
        ModelDev<double> u = md;
        for (int b = 0; b < nq; ++b) {
            u.mass[b] = 0;
            for (int a = 0; a < 3; ++a) u.mc[b][a] = 0;
            for (int a = 0; a < 9; ++a) u.Io[b][a] = 0;
        }
        const double* c = t.com[k];
        const double cc = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
        u.mass[k] = 1.0;
        for (int a = 0; a < 3; ++a) u.mc[k][a] = c[a];
        for (int r = 0; r < 3; ++r)
            for (int q = 0; q < 3; ++q) u.Io[k][3 * r + q] = (r == q ? cc : 0.0) - c[r] * c[q];
        chain_coef_vector(u, nq, basis.dmass[k]);
        basis.mass[k] = t.mass[k];
    }
}
