// examples/texture-mapping.rs — BASELINE.json configs[1].
// assets/earth_cube.png is missing from the reference checkout
// (.MISSING_LARGE_BLOBS); tools/bake_assets.py registers a stand-in 4x3 atlas
// tiled from earth.jpg under that name (make-cube-map.sh recipe).
#include "examples.hpp"
using namespace portrayer;

PORTRAYER_EXAMPLE(texture_mapping, "texture-mapping") {
    auto mat_mirror = Arc(Material{.diffuse = {0.0, 0.0, 0.0}, .specular = {0.6, 0.6, 0.6}, .shininess = 1000.0,
                                   .reflectivity = 1.0});
    auto mat_wood = Arc(Material{.diffuse = {0.545, 0.353, 0.169}, .specular = {0.5, 0.7, 0.5}, .shininess = 25.0});
    auto earth = ImageTexture::open("assets/earth.jpg");
    auto mat_tex = Arc(Material{.diffuse = {0.506, 0.78, 0.518}, .specular = {0.5, 0.5, 0.5}, .shininess = 25.0,
                                .texture = earth});
    auto earth_cubemap = ImageTexture::open("assets/earth_cube.png");
    auto mat_tex_cube = Arc(Material{.diffuse = {0.506, 0.78, 0.518}, .specular = {0.5, 0.5, 0.5}, .shininess = 25.0,
                                     .texture = earth_cubemap});

    NodeRef mirror = SceneNode::from(Geometry(Cube{}, mat_wood))
        .scaled({9.0, 0.5, 6.0})
        .rotated_x(Radians::from_degrees(10.0))
        .with_child(SceneNode::from(Geometry(Cube{}, mat_mirror))
                        .scaled({8.1 / 9.0, 0.05 / 0.5, 5.4 / 6.0})
                        .translated({0.0, 0.27 / 0.5, 0.0}).into())
        .into();

    ExampleScene ex;
    ex.name = "texture-mapping";
    ex.scene = HierScene{
        .root = SceneNode::from(std::vector<NodeRef>{
            mirror,
            SceneNode::from(Geometry(Plane{}, mat_tex))
                .scaled({8.0, 1.0, 2.0}).rotated_x(Radians::from_degrees(90.0)).translated({0.0, 2.0, -2.0}).into(),
            SceneNode::from(Geometry(Cube{}, mat_tex_cube)).scaled(1.4).translated({-2.0, 2.0, 0.0}).into(),
            SceneNode::from(Geometry(Sphere{}, mat_tex)).translated({2.0, 2.0, 0.0}).into(),
        }).into(),
        .lights = {
            Light{.position = {-6.0, 5.0, 4.0}, .color = {0.5, 0.5, 0.5}},
            Light{.position = {6.0, 5.0, 4.0}, .color = {0.5, 0.5, 0.5}},
            Light{.position = {0.0, 1.0, -4.0}, .color = {0.5, 0.5, 0.5}},
        },
        .ambient = {0.3, 0.3, 0.3},
    };
    ex.cam = CameraSettings{.eye = {0.0, 10.15667, 11.579666}, .center = {0.0, -5.913023, -7.571445},
                            .up = Vec3::up(), .fovy = Radians::from_degrees(25.0)};
    ex.width = 910;
    ex.height = 512;
    ex.background = sky_gradient;
    return ex;
}
