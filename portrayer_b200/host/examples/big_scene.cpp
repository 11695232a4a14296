// examples/big-scene.rs — BASELINE.json configs[2]; the one scene the
// reference publishes timings for (render/09_kdtree_timing_data.csv).
#include "../rand07.hpp"
#include "examples.hpp"
using namespace portrayer;

namespace portrayer {
ExampleScene make_big_scene(size_t n) {
    // Want the result to be random but also completely reproducible
    StdRng rng = StdRng::seed_from_u64(1234939301ull);

    std::vector<MaterialRef> materials;
    for (int i = 0; i < 15; ++i) {
        Material m;
        // struct-literal fields are evaluated in source order: r, g, b
        double r = rng.gen_f64(), g = rng.gen_f64(), b = rng.gen_f64();
        m.diffuse = {r, g, b};
        m.specular = {0.3, 0.3, 0.3};
        m.shininess = 25.0;
        materials.push_back(Arc(std::move(m)));
    }
    const std::vector<Primitive> primitives = {Sphere{}, Cube{}, Cone{}, Cylinder{}};

    const double width = 800.0, length = 800.0, height = 800.0;

    std::vector<NodeRef> nodes;
    for (size_t i = 0; i < n; ++i) {
        const double x = (double)i / (double)(n - 1) * width - width / 2.0;
        for (size_t j = 0; j < n; ++j) {
            const double y = (double)j / (double)(n - 1) * length - length / 2.0;
            for (size_t k = 0; k < n; ++k) {
                const double z = (double)k / (double)(n - 1) * height - height / 2.0;

                const Primitive& prim = rng.choose(primitives);
                const MaterialRef& mat = rng.choose(materials);

                const double scale_base = 30.0, scale_increase = 30.0;
                const double scale = scale_increase * rng.gen_f64() + scale_base;
                const Radians angle = Radians::from_degrees(360.0 * rng.gen_f64());
                const double y_jitter = rng.gen_f64() * 50.0;
                nodes.push_back(SceneNode::from(Geometry(prim, mat))
                                    .scaled(scale)
                                    .rotated_xzy(angle)
                                    .translated({x, y + y_jitter, z})
                                    .into());
            }
        }
    }

    ExampleScene ex;
    ex.name = "big-scene";
    ex.scene = HierScene{
        .root = SceneNode::from(std::move(nodes)).into(),
        .lights = {
            Light{.position = {-100.0, 150.0, 400.0}, .color = {0.9, 0.9, 0.9}},
            Light{.position = {100.0, -150.0, 800.0}, .color = {0.7, 0.7, 0.7}},
            Light{.position = {400.0, 100.0, 150.0}, .color = {0.7, 0.0, 0.7}},
        },
        .ambient = {0.3, 0.3, 0.3},
    };
    ex.cam = CameraSettings{.eye = {0.0, 0.0, 1200.0}, .center = {0.0, 0.0, 0.0}, .up = Vec3::up(),
                            .fovy = Radians::from_degrees(50.0)};
    ex.width = 1980;
    ex.height = 1020;
    ex.background = sky_gradient;
    return ex;
}
}  // namespace portrayer

PORTRAYER_EXAMPLE(big_scene, "big-scene") { return make_big_scene(10); }
