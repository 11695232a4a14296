// Registry of scene programs: C++ restatements of the reference's
// examples/*.rs (same objects, same transforms in the same order, same
// camera, image size and background closure).  They are the workloads of
// BASELINE.json `configs`; each file cites the example it mirrors.
#pragma once
#include <functional>
#include <map>
#include <string>
#include <vector>

#include "../kdtree.hpp"
#include "../scene.hpp"

namespace portrayer {

struct ExampleScene {
    std::string name;
    HierScene scene;
    CameraSettings cam;
    size_t width = 0, height = 0;
    std::function<Rgb(Uv)> background;
    // Set only by the known-answer scenes that hand-build a kd-tree
    // (src/kdtree/node.rs:219-352); otherwise the tree comes from KDTreeScene::from.
    std::shared_ptr<KDTreeScene> prebuilt;
};

using ExampleFn = std::function<ExampleScene()>;

struct ExampleRegistrar {
    ExampleRegistrar(const std::string& name, ExampleFn fn);
};
const std::map<std::string, ExampleFn>& example_registry();

// |uv| Rgb {r: 0.2, g: 0.4, b: 0.6} * (1.0 - uv.v) + Rgb::blue() * uv.v   — used by almost every example
inline Rgb sky_gradient(Uv uv) { return Rgb{0.2, 0.4, 0.6} * (1.0 - uv.v) + Rgb::blue() * uv.v; }

// Synthetic kd-tree stress scenes (BASELINE.json configs[2])
ExampleScene make_synthetic_instances(size_t n_instances, uint64_t seed);
ExampleScene make_synthetic_triangles(size_t n_triangles, uint64_t seed, size_t kd_mesh_depth);
ExampleScene make_big_scene(size_t n);  // examples/big-scene.rs with its `n` exposed

}  // namespace portrayer

#define PORTRAYER_EXAMPLE(ident, name_str)                                      \
    static ::portrayer::ExampleScene example_fn_##ident();                           \
    static ::portrayer::ExampleRegistrar registrar_##ident(name_str, example_fn_##ident); \
    static ::portrayer::ExampleScene example_fn_##ident()
