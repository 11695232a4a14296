// examples/nonhier.rs — BASELINE.json configs[0]
#include "examples.hpp"
using namespace portrayer;

PORTRAYER_EXAMPLE(nonhier, "nonhier") {
    auto mat1 = Arc(Material{.diffuse = {0.7, 1.0, 0.7}, .specular = {0.5, 0.7, 0.5}, .shininess = 25.0});
    auto mat2 = Arc(Material{.diffuse = {0.5, 0.5, 0.5}, .specular = {0.5, 0.7, 0.5}, .shininess = 25.0});
    auto mat3 = Arc(Material{.diffuse = {1.0, 0.6, 0.1}, .specular = {0.5, 0.7, 0.5}, .shininess = 25.0});
    auto mat4 = Arc(Material{.diffuse = {0.7, 0.6, 1.0}, .specular = {0.5, 0.4, 0.8}, .shininess = 25.0});

    auto monkey = MeshData::load_obj("assets/monkey.obj");

    ExampleScene ex;
    ex.name = "nonhier";
    ex.scene = HierScene{
        .root = SceneNode::from(std::vector<NodeRef>{
            SceneNode::from(Geometry(Sphere{}, mat1)).scaled(100.0).translated({0.0, 0.0, -400.0}).into(),
            SceneNode::from(Geometry(Sphere{}, mat1)).scaled(150.0).translated({200.0, 50.0, -100.0}).into(),
            SceneNode::from(Geometry(Sphere{}, mat2)).scaled(1000.0).translated({0.0, -1200.0, -500.0}).into(),
            SceneNode::from(Geometry(Cube{}, mat4)).scaled(100.0).translated({-150.0, -75.0, 50.0}).into(),
            SceneNode::from(Geometry(Sphere{}, mat3)).scaled(50.0).translated({-100.0, 25.0, -300.0}).into(),
            SceneNode::from(Geometry(Sphere{}, mat1)).scaled(25.0).translated({0.0, 100.0, -250.0}).into(),
            SceneNode::from(Geometry(Mesh(monkey, Shading::Flat), mat3))
                .scaled(100.0).translated({-150.0, 200.0, -100.0}).into(),
        }).into(),
        .lights = {
            Light{.position = {-100.0, 150.0, 400.0}, .color = {0.9, 0.9, 0.9}},  // white_light
            Light{.position = {400.0, 100.0, 150.0}, .color = {0.7, 0.0, 0.7}},   // magenta_light
        },
        .ambient = {0.3, 0.3, 0.3},
    };
    ex.cam = CameraSettings{.eye = {0.0, 0.0, 800.0}, .center = {0.0, 0.0, 0.0}, .up = Vec3::up(),
                            .fovy = Radians::from_degrees(50.0)};
    ex.width = 256;
    ex.height = 256;
    ex.background = sky_gradient;
    return ex;
}
