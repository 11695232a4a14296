// Synthetic kd-tree stress scenes — BASELINE.json configs[2] ("a synthetic
// N=10^4–10^6 random sphere/triangle scene"), generator of SURVEY §8d M3b:
// the big-scene.rs recipe on an n = ceil(N^(1/3)) grid whose box grows with n,
// and a single KDMesh of N random small triangles.  SplitMix64, fixed seed.
#include <cmath>

#include "examples.hpp"
using namespace portrayer;

namespace {
struct SplitMix64 {
    uint64_t s;
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double f64() { return static_cast<double>(next() >> 11) * (1.0 / 9007199254740992.0); }
    size_t below(size_t n) { return static_cast<size_t>(next() % n); }
};
}  // namespace

namespace portrayer {

ExampleScene make_synthetic_instances(size_t n_instances, uint64_t seed) {
    SplitMix64 rng{seed};
    std::vector<MaterialRef> materials;
    for (int i = 0; i < 15; ++i) {
        Material m;
        m.diffuse = {rng.f64(), rng.f64(), rng.f64()};
        m.specular = {0.3, 0.3, 0.3};
        m.shininess = 25.0;
        materials.push_back(Arc(std::move(m)));
    }
    const std::vector<Primitive> primitives = {Sphere{}, Cube{}, Cone{}, Cylinder{}};
    size_t n = static_cast<size_t>(std::ceil(std::cbrt(static_cast<double>(n_instances))));
    while (n * n * n < n_instances) ++n;
    if (n < 2) n = 2;
    const double box = 800.0 * static_cast<double>(n) / 10.0;

    std::vector<NodeRef> nodes;
    nodes.reserve(n_instances);
    for (size_t i = 0; i < n && nodes.size() < n_instances; ++i) {
        const double x = (double)i / (double)(n - 1) * box - box / 2.0;
        for (size_t j = 0; j < n && nodes.size() < n_instances; ++j) {
            const double y = (double)j / (double)(n - 1) * box - box / 2.0;
            for (size_t k = 0; k < n && nodes.size() < n_instances; ++k) {
                const double z = (double)k / (double)(n - 1) * box - box / 2.0;
                const Primitive& prim = primitives[rng.below(primitives.size())];
                const MaterialRef& mat = materials[rng.below(materials.size())];
                const double scale = 30.0 * rng.f64() + 30.0;
                const Radians angle = Radians::from_degrees(360.0 * rng.f64());
                const double yj = rng.f64() * 50.0;
                nodes.push_back(SceneNode::from(Geometry(prim, mat)).scaled(scale).rotated_xzy(angle)
                                    .translated({x, y + yj, z}).into());
            }
        }
    }
    ExampleScene ex;
    ex.name = "synthetic-instances";
    ex.scene = HierScene{
        .root = SceneNode::from(std::move(nodes)).into(),
        .lights = {
            Light{.position = {-0.125 * box, 0.1875 * box, 0.5 * box}, .color = {0.9, 0.9, 0.9}},
            Light{.position = {0.125 * box, -0.1875 * box, box}, .color = {0.7, 0.7, 0.7}},
            Light{.position = {0.5 * box, 0.125 * box, 0.1875 * box}, .color = {0.7, 0.0, 0.7}},
        },
        .ambient = {0.3, 0.3, 0.3},
    };
    ex.cam = CameraSettings{.eye = {0.0, 0.0, 1.5 * box}, .center = {0.0, 0.0, 0.0}, .up = Vec3::up(),
                            .fovy = Radians::from_degrees(50.0)};
    ex.width = 1980;
    ex.height = 1020;
    ex.background = sky_gradient;
    return ex;
}

ExampleScene make_synthetic_triangles(size_t n_triangles, uint64_t seed, size_t kd_mesh_depth) {
    SplitMix64 rng{seed};
    const double box = 800.0, edge = 0.005 * box * std::cbrt(1.0e4 / std::max<double>(1.0, (double)n_triangles)) * 4.0;
    std::vector<Vec3> positions;
    std::vector<std::array<size_t, 3>> tris;
    positions.reserve(n_triangles * 3);
    for (size_t t = 0; t < n_triangles; ++t) {
        Vec3 c{(rng.f64() - 0.5) * box, (rng.f64() - 0.5) * box, (rng.f64() - 0.5) * box};
        for (int v = 0; v < 3; ++v)
            positions.push_back(c + Vec3{(rng.f64() - 0.5) * edge, (rng.f64() - 0.5) * edge, (rng.f64() - 0.5) * edge});
        tris.push_back({3 * t, 3 * t + 1, 3 * t + 2});
    }
    MeshData data(std::move(positions), std::move(tris), {}, {});
    auto mat = Arc(Material{.diffuse = {0.8, 0.6, 0.2}, .specular = {0.3, 0.3, 0.3}, .shininess = 25.0});

    ExampleScene ex;
    ex.name = "synthetic-triangles";
    ex.scene = HierScene{
        .root = SceneNode::from(Geometry(KDMesh(data, Shading::Flat, kd_mesh_depth), mat)).into(),
        .lights = {
            Light{.position = {-100.0, 150.0, 400.0}, .color = {0.9, 0.9, 0.9}},
            Light{.position = {400.0, 100.0, 150.0}, .color = {0.7, 0.0, 0.7}},
        },
        .ambient = {0.3, 0.3, 0.3},
    };
    ex.cam = CameraSettings{.eye = {0.0, 0.0, 1.5 * box}, .center = {0.0, 0.0, 0.0}, .up = Vec3::up(),
                            .fovy = Radians::from_degrees(50.0)};
    ex.width = 1980;
    ex.height = 1020;
    ex.background = sky_gradient;
    return ex;
}

}  // namespace portrayer
