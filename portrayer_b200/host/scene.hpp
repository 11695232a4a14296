// Host mirror of the reference's scene-building API.  In the target design
// this layer stays Rust and unchanged (BASELINE.json north_star); there is no
// Rust toolchain in this image, so it is restated in C++ with the same names,
// argument meaning and composition order so example scenes read like
// examples/*.rs.  Nothing here runs per pixel.
//
//   Material            src/material.rs:51-86
//   Light/Falloff/...   src/light.rs:11-85
//   Primitive enum      src/primitive.rs:67-81
//   Geometry/SceneNode  src/scene.rs:21-205
//   BoundingBox/Bounds  src/bounding_box.rs:10-148
//   Triangle/MeshData   src/primitive/triangle.rs:9-36, src/primitive/mesh.rs:12-143
//   CameraSettings      src/camera.rs:5-14
#pragma once
#include <cstring>
#include <cassert>
#include <functional>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "math.hpp"

namespace portrayer {

// ---------------------------------------------------------------- textures
// RgbImageBuffer, src/texture.rs:78-102.  Decoding (image::open(..).to_rgb())
// is host I/O outside the hot path; texels come from the decoded-asset store
// (see assets.hpp) so that host, oracle and device read identical bytes.
struct RgbImageBuffer {
    uint32_t width = 0, height = 0;
    std::vector<uint8_t> data;  // RGB8 row-major
    std::string name;
    static std::shared_ptr<RgbImageBuffer> open(const std::string& path);
    // Identity of the texel content for the device-side texture residency cache (PtTexture.key):
    // 64-bit FNV-1a over (width, height, texels), computed once per loaded image.  The Rust glue
    // would use the Arc's address instead; a content hash also survives re-loading the same file.
    uint64_t content_key() const {
        if (key_ == 0) {
            uint64_t h = 0xCBF29CE484222325ull;
            auto mix = [&h](uint64_t v) { h = (h ^ v) * 0x100000001B3ull; };
            mix(width);
            mix(height);
            size_t i = 0;
            for (; i + 8 <= data.size(); i += 8) {
                uint64_t w;
                memcpy(&w, data.data() + i, 8);
                mix(w);
            }
            for (; i < data.size(); ++i) mix(data[i]);
            key_ = h ? h : 1;
        }
        return key_;
    }

private:
    mutable uint64_t key_ = 0;
};
struct ImageTexture {  // src/texture.rs:149-169
    std::shared_ptr<RgbImageBuffer> buffer;
    static std::shared_ptr<ImageTexture> open(const std::string& path) {
        return std::make_shared<ImageTexture>(ImageTexture{RgbImageBuffer::open(path)});
    }
};
struct NormalMap {  // src/texture.rs:175-187
    std::shared_ptr<RgbImageBuffer> buffer;
    static std::shared_ptr<NormalMap> open(const std::string& path) {
        return std::make_shared<NormalMap>(NormalMap{RgbImageBuffer::open(path)});
    }
};
// Texture::Image only. Texture::FnTex (texture.rs:22-27) cannot cross the FFI
// boundary; no example uses one. The glue rejects it.
using Texture = ImageTexture;

// ---------------------------------------------------------------- material
constexpr double AIR_REFRACTION_INDEX = 1.00;
constexpr double WATER_REFRACTION_INDEX = 1.33;
constexpr double WINDOW_GLASS_REFRACTION_INDEX = 1.51;
constexpr double OPTICAL_GLASS_REFRACTION_INDEX = 1.92;
constexpr double DIAMOND_REFRACTION_INDEX = 2.42;

struct Material {
    Rgb diffuse{};
    Rgb specular{};
    double shininess = 0.0;
    double reflectivity = 0.0;
    double glossy_side_length = 0.0;
    double refraction_index = 0.0;
    std::shared_ptr<Texture> texture{};
    Mat3 uv_trans = Mat3::identity();
    std::shared_ptr<NormalMap> normals{};
};
using MaterialRef = std::shared_ptr<const Material>;
inline MaterialRef Arc(Material m) { return std::make_shared<const Material>(std::move(m)); }

// ---------------------------------------------------------------- lights
struct Falloff {
    double c0 = 1.0, c1 = 0.0, c2 = 0.0;
};
struct Parallelogram {
    Vec3 a{}, b{};
    bool is_empty() const { return a == Vec3::zero() || b == Vec3::zero(); }
};
struct Light {
    Vec3 position{};
    Rgb color{};
    Falloff falloff{};
    Parallelogram area{};
};

// ---------------------------------------------------------------- camera
struct CameraSettings {
    Vec3 eye{};
    Vec3 center{};
    Vec3 up{};
    Radians fovy{};
};

// ---------------------------------------------------------------- bounds
class BoundingBox {
  public:
    BoundingBox() : BoundingBox(Vec3::zero(), Vec3::zero()) {}
    // src/bounding_box.rs:57-82
    BoundingBox(Vec3 min, Vec3 max) : min_(min), max_(max) {
        if (!(min.x <= max.x && min.y <= max.y && min.z <= max.z))
            throw std::runtime_error("bounding box min must be less than max");
        Vec3 size = max - min;
        size = Vec3::partial_max(size, Vec3(EPSILON));
        Vec3 center = (min + max) / 2.0;
        Mat4 trans = Mat4::scaling_3d(size).translated_3d(center);
        invtrans_ = trans.inverted();
    }
    Vec3 min() const { return min_; }
    Vec3 max() const { return max_; }
    const Mat4& invtrans() const { return invtrans_; }
    // HACK in the reference: squared diagonal. src/bounding_box.rs:95-99
    double extent() const { return (max_ - min_).magnitude_squared(); }
    bool operator==(const BoundingBox& o) const { return min_ == o.min_ && max_ == o.max_; }

  private:
    Vec3 min_, max_;
    Mat4 invtrans_;
};

// Mat4 * BoundingBox: AABB of the 8 transformed corners. src/bounding_box.rs:123-148
inline BoundingBox operator*(const Mat4& t, const BoundingBox& b) {
    Vec3 mn(INFINITY_F64), mx(-INFINITY_F64);
    const double xs[2] = {b.min().x, b.max().x}, ys[2] = {b.min().y, b.max().y}, zs[2] = {b.min().z, b.max().z};
    for (double x : xs)
        for (double y : ys)
            for (double z : zs) {
                Vec3 v = transformed_point(Vec3{x, y, z}, t);
                mn = Vec3::partial_min(mn, v);
                mx = Vec3::partial_max(mx, v);
            }
    return BoundingBox(mn, mx);
}

// Bounds for Vec<T>, src/bounding_box.rs:21-35
template <class It, class F>
BoundingBox bounds_of(It first, It last, F&& get_bounds) {
    if (first == last) return BoundingBox(Vec3::zero(), Vec3::zero());
    BoundingBox b0 = get_bounds(*first);
    Vec3 mn = b0.min(), mx = b0.max();
    for (It it = std::next(first); it != last; ++it) {
        BoundingBox b = get_bounds(*it);
        mn = Vec3::partial_min(mn, b.min());
        mx = Vec3::partial_max(mx, b.max());
    }
    return BoundingBox(mn, mx);
}

// ---------------------------------------------------------------- triangles / meshes
struct Triangle {  // src/primitive/triangle.rs:9-19
    Vec3 a, b, c;
    std::optional<std::array<Vec3, 3>> normals{};
    std::optional<std::array<Uv, 3>> tex_coords{};
    static Triangle flat(Vec3 a, Vec3 b, Vec3 c) { return Triangle{a, b, c, std::nullopt, std::nullopt}; }
    BoundingBox bounds() const {  // triangle.rs:29-36
        Vec3 mn = Vec3::partial_min(a, Vec3::partial_min(b, c));
        Vec3 mx = Vec3::partial_max(a, Vec3::partial_max(b, c));
        return BoundingBox(mn, mx);
    }
};

enum class Shading { Flat, Smooth };  // src/primitive/mesh.rs:11-18

class MeshData {  // src/primitive/mesh.rs:20-113
  public:
    MeshData(std::vector<Vec3> positions, std::vector<std::array<size_t, 3>> triangles, std::vector<Vec3> normals,
             std::vector<Uv> tex_coords);
    // tobj 0.1.7 semantics: first model only, one vertex per unique (v, vt, vn)
    // triple in first-use order, f32 parse then widen. src/primitive/mesh.rs:57-61
    static std::shared_ptr<const MeshData> load_obj(const std::string& path);

    void save_baked(const std::string& file) const;  // tools/sync_assets.py
    std::vector<Triangle> triangles(Shading shading) const;  // mesh.rs:95-113
    const BoundingBox& bounds() const { return bounds_; }
    size_t num_triangles() const { return triangles_.size(); }
    size_t num_positions() const { return positions_.size(); }
    size_t num_normals() const { return normals_.size(); }
    bool has_tex_coords() const { return !tex_coords_.empty(); }

  private:
    std::vector<std::array<size_t, 3>> triangles_;
    std::vector<Vec3> positions_;
    std::vector<Vec3> normals_;
    std::vector<Uv> tex_coords_;
    BoundingBox bounds_;
};

struct Mesh {  // src/primitive/mesh.rs:115-143
    std::shared_ptr<const MeshData> data;
    Shading shading;
    Mesh(std::shared_ptr<const MeshData> d, Shading s) : data(std::move(d)), shading(s) {
        if (shading == Shading::Smooth && data->num_positions() != data->num_normals())
            throw std::runtime_error(
                "Meshes must have a vertex normal for each vertex if they are to be used with smooth shading");
    }
};

struct KDMeshTree;  // kdtree.hpp
struct KDMesh {     // src/kdtree/kdmesh.rs:17-58
    std::shared_ptr<const KDMeshTree> triangles;
    KDMesh(const MeshData& data, Shading shading);  // reads KD_MESH_DEPTH (kdmesh.rs:51-53)
    KDMesh(const MeshData& data, Shading shading, size_t max_tree_depth);
};

struct Sphere {};
struct Plane {};
struct Cube {};
struct Cylinder {};
struct Cone {};

enum class PrimKind { Sphere, Triangle, Mesh, KDMesh, Plane, Cube, Cylinder, Cone };

struct Primitive {
    PrimKind kind;
    std::shared_ptr<const Triangle> triangle{};
    std::shared_ptr<const MeshData> mesh{};
    Shading shading = Shading::Flat;
    std::shared_ptr<const KDMeshTree> kdmesh{};

    Primitive(Sphere) : kind(PrimKind::Sphere) {}
    Primitive(Plane) : kind(PrimKind::Plane) {}
    Primitive(Cube) : kind(PrimKind::Cube) {}
    Primitive(Cylinder) : kind(PrimKind::Cylinder) {}
    Primitive(Cone) : kind(PrimKind::Cone) {}
    Primitive(Triangle t) : kind(PrimKind::Triangle), triangle(std::make_shared<const Triangle>(std::move(t))) {}
    Primitive(Mesh m) : kind(PrimKind::Mesh), mesh(std::move(m.data)), shading(m.shading) {}
    Primitive(KDMesh m) : kind(PrimKind::KDMesh), kdmesh(std::move(m.triangles)) {}

    BoundingBox bounds() const;  // per-primitive Bounds impls
};

// ---------------------------------------------------------------- scene graph
struct Geometry {  // src/scene.rs:21-33
    Primitive primitive;
    MaterialRef material;
    Geometry(Primitive p, MaterialRef m) : primitive(std::move(p)), material(std::move(m)) {}
};

class SceneNode;
using NodeRef = std::shared_ptr<const SceneNode>;

class SceneNode {  // src/scene.rs:36-205
  public:
    SceneNode() = default;
    static SceneNode from(Geometry g) {
        SceneNode n;
        n.geometry_ = std::move(g);
        return n;
    }
    static SceneNode from(std::vector<NodeRef> children) {
        SceneNode n;
        n.children_ = std::move(children);
        return n;
    }
    static SceneNode from(NodeRef child) {
        SceneNode n;
        n.children_.push_back(std::move(child));
        return n;
    }
    NodeRef into() && { return std::make_shared<const SceneNode>(std::move(*this)); }

    const std::optional<Geometry>& geometry() const { return geometry_; }
    const Mat4& trans() const { return trans_; }
    const Mat4& inverse_trans() const { return invtrans_; }
    const Mat4& normal_trans() const { return normal_trans_; }
    const std::vector<NodeRef>& children() const { return children_; }

    SceneNode with_child(NodeRef c) && {
        children_.push_back(std::move(c));
        return std::move(*this);
    }
    SceneNode with_children(std::vector<NodeRef> cs) && {
        for (auto& c : cs) children_.push_back(std::move(c));
        return std::move(*this);
    }
    SceneNode scaled(Vec3 s) && {
        set_transform(trans_.scaled_3d(s));
        return std::move(*this);
    }
    SceneNode translated(Vec3 t) && {
        set_transform(trans_.translated_3d(t));
        return std::move(*this);
    }
    // rotate about x, then z, then y. src/scene.rs:177-180
    SceneNode rotated_xzy(Radians x, Radians y, Radians z) && {
        return std::move(*this).rotated_x(x).rotated_z(z).rotated_y(y);
    }
    SceneNode rotated_xzy(Radians all) && { return std::move(*this).rotated_xzy(all, all, all); }
    SceneNode rotated_x(Radians a) && {
        set_transform(trans_.rotated_x(a.get()));
        return std::move(*this);
    }
    SceneNode rotated_y(Radians a) && {
        set_transform(trans_.rotated_y(a.get()));
        return std::move(*this);
    }
    SceneNode rotated_z(Radians a) && {
        set_transform(trans_.rotated_z(a.get()));
        return std::move(*this);
    }
    void set_transform(const Mat4& t) {  // src/scene.rs:200-204
        trans_ = t;
        invtrans_ = t.inverted();
        normal_trans_ = invtrans_.transposed();
    }

  private:
    std::optional<Geometry> geometry_{};
    Mat4 trans_ = Mat4::identity();
    Mat4 invtrans_ = Mat4::identity();
    Mat4 normal_trans_ = Mat4::identity();
    std::vector<NodeRef> children_{};
};

struct HierScene {  // Scene<Arc<SceneNode>>, src/scene.rs:11-18
    NodeRef root;
    std::vector<Light> lights;
    Rgb ambient{};
};

}  // namespace portrayer
