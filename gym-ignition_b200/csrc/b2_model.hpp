// Model loader: URDF / SDF-subset -> flattened link/joint tables (north_star (a)).
// Replaces sdformat + SdfEntityCreator + Model::createECMResources for the hot path
// (reference: cpp/scenario/gazebo/src/World.cpp:70-180,394-429, Model.cpp:143-188,527-579).
#pragma once

#include <string>
#include <vector>

#include "../../include/b2sim.h"
#include "b2_rbd.hpp"

namespace b2 {

struct Pose {
    M3<double> R{{1, 0, 0, 0, 1, 0, 0, 0, 1}};
    V3<double> p{0, 0, 0};
};
inline Pose compose(const Pose& a, const Pose& b) { return Pose{mul(a.R, b.R), a.p + mul(a.R, b.p)}; }
inline Pose inverse(const Pose& a)
{
    Pose r;
    r.R = transpose(a.R);
    r.p = -1.0 * mul(r.R, a.p);
    return r;
}
Pose pose_from_xyz_rpy(const double xyz[3], const double rpy[3]);
Pose pose_from_xyz_quat(const double pose7[7]);  // xyz + quaternion wxyz (not normalised by the reference)

enum class ShapeType { Box, Sphere, Cylinder, Plane };
struct CollisionShape {
    std::string name;
    int link = -1;
    ShapeType type = ShapeType::Box;
    double size[3] = {0, 0, 0};  // box: xyz; sphere: r; cylinder: r, length; plane: normal
    Pose pose;                   // in the link frame
    double mu = 1.0;
};

}  // namespace b2

// Opaque type of the C ABI.
struct b2model {
    std::string name;
    bool is_static = false;
    bool fixed_base = true;
    b2_model_tables t{};
    std::vector<std::string> joint_names;  // moving joints, body order
    std::vector<std::string> link_names;   // all links except "world", file order
    std::vector<b2::CollisionShape> shapes;
    // closed-form coefficients when kind is CHAIN1 / CHAIN_PR (gravity- and pose-dependent: see fit())
    b2::ChainCoef<double> coef{};
    b2::ChainBasis<double> basis{};  // mass sensitivities of coef (per-env domain randomisation)

    // Fills a kernel-side table in scalar type T for a model placed at `base` under gravity g.
    template <typename T>
    void to_device_tables(const b2::Pose& base, const double g[3], b2::ModelDev<T>& out) const;
    // Classifies the model and fits the closed-form coefficients for (base pose, gravity, dt).
    // Returns the kind actually usable (falls back to B2_KIND_TREE when the fit does not verify).
    int fit(const b2::Pose& base, const double g[3], double dt);
    void fit_basis(const b2::ModelDev<double>& md);
};

namespace b2 {
// Throws std::runtime_error on malformed input.
b2model* parse_model(const char* xml, size_t len);
}  // namespace b2
