// Primitive bounds, MeshData / OBJ loading, KDMesh construction, texture store.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <mutex>
#include <sstream>
#include <tuple>

#include "assets.hpp"
#include "kdtree.hpp"
#include "scene.hpp"

namespace portrayer {

// ------------------------------------------------------------------ bounds
BoundingBox Primitive::bounds() const {
    switch (kind) {
        case PrimKind::Sphere: return BoundingBox(Vec3(-1.0), Vec3(1.0));                    // sphere.rs:18-24
        case PrimKind::Cube: return BoundingBox(Vec3(-0.5), Vec3(0.5));                      // cube.rs:30-36
        case PrimKind::Plane: return BoundingBox(Vec3{-0.5, 0.0, -0.5}, Vec3{0.5, 0.0, 0.5});  // plane.rs:18-24
        case PrimKind::Cylinder:                                                              // cylinder.rs:19-25
        case PrimKind::Cone:                                                                  // cone.rs:19-25
            return BoundingBox(Vec3{-0.5, -0.5, -0.5}, Vec3{0.5, 0.5, 0.5});
        case PrimKind::Triangle: return triangle->bounds();
        case PrimKind::Mesh: return mesh->bounds();             // mesh.rs:123-127
        case PrimKind::KDMesh: return kdmesh->root->node_bounds();  // kdmesh.rs:26-30
    }
    throw std::logic_error("unreachable");
}

// ------------------------------------------------------------------ MeshData
MeshData::MeshData(std::vector<Vec3> positions, std::vector<std::array<size_t, 3>> triangles, std::vector<Vec3> normals,
                   std::vector<Uv> tex_coords)
    : triangles_(std::move(triangles)),
      positions_(std::move(positions)),
      normals_(std::move(normals)),
      tex_coords_(std::move(tex_coords)) {
    if (positions_.empty()) throw std::runtime_error("Meshes must have at least one vertex");
    Vec3 mn = positions_[0], mx = positions_[0];
    for (size_t i = 1; i < positions_.size(); ++i) {
        mn = Vec3::partial_min(mn, positions_[i]);
        mx = Vec3::partial_max(mx, positions_[i]);
    }
    if (!tex_coords_.empty() && tex_coords_.size() != positions_.size())
        throw std::runtime_error("If meshes have texture coordinates, they must have enough for all vertices");
    bounds_ = BoundingBox(mn, mx);
}

std::vector<Triangle> MeshData::triangles(Shading shading) const {
    std::vector<Triangle> out;
    out.reserve(triangles_.size());
    for (const auto& t : triangles_) {
        Triangle tri{positions_[t[0]], positions_[t[1]], positions_[t[2]], std::nullopt, std::nullopt};
        if (shading == Shading::Smooth) tri.normals = std::array<Vec3, 3>{normals_[t[0]], normals_[t[1]], normals_[t[2]]};
        if (!tex_coords_.empty()) tri.tex_coords = std::array<Uv, 3>{tex_coords_[t[0]], tex_coords_[t[1]], tex_coords_[t[2]]};
        out.push_back(tri);
    }
    return out;
}

namespace {
// one "v/vt/vn" corner; indices are 0-based, -1 = absent
struct Corner {
    long v = -1, vt = -1, vn = -1;
    bool operator<(const Corner& o) const { return std::tie(v, vt, vn) < std::tie(o.v, o.vt, o.vn); }
};

bool parse_corner(const std::string& tok, size_t n_pos, size_t n_tex, size_t n_nrm, Corner& out) {
    long idx[3] = {0, 0, 0};
    bool have[3] = {false, false, false};
    size_t field = 0, start = 0;
    for (size_t i = 0; i <= tok.size() && field < 3; ++i) {
        if (i == tok.size() || tok[i] == '/') {
            if (i > start) {
                idx[field] = std::strtol(tok.substr(start, i - start).c_str(), nullptr, 10);
                have[field] = true;
            }
            ++field;
            start = i + 1;
        }
    }
    if (!have[0]) return false;
    auto fix = [](long i, size_t n) -> long { return i < 0 ? static_cast<long>(n) + i : i - 1; };  // OBJ relative indices
    out.v = fix(idx[0], n_pos);
    out.vt = have[1] ? fix(idx[1], n_tex) : -1;
    out.vn = have[2] ? fix(idx[2], n_nrm) : -1;
    return true;
}
}  // namespace

// Baked form of an OBJ (tools/sync_assets.py): exactly the arrays the loader
// below produces, so a checkout without the reference's assets/ directory can
// still build the scenes whose meshes are committed under tests/golden/meshes.
//   "PTMS" u32 n_pos n_nrm n_uv n_tri | f32 pos[3 n_pos] | f32 nrm[3 n_nrm] | f32 uv[2 n_uv] | u32 idx[3 n_tri]
static std::shared_ptr<const MeshData> load_baked_mesh(const std::string& file) {
    std::ifstream in(file, std::ios::binary);
    if (!in) return nullptr;
    char magic[4];
    uint32_t n[4];
    in.read(magic, 4);
    in.read(reinterpret_cast<char*>(n), sizeof n);
    if (!in || std::memcmp(magic, "PTMS", 4) != 0) throw std::runtime_error("bad baked mesh: " + file);
    std::vector<float> pos(3 * (size_t)n[0]), nrm(3 * (size_t)n[1]), uv(2 * (size_t)n[2]);
    std::vector<uint32_t> idx(3 * (size_t)n[3]);
    in.read(reinterpret_cast<char*>(pos.data()), pos.size() * 4);
    in.read(reinterpret_cast<char*>(nrm.data()), nrm.size() * 4);
    in.read(reinterpret_cast<char*>(uv.data()), uv.size() * 4);
    in.read(reinterpret_cast<char*>(idx.data()), idx.size() * 4);
    if (!in) throw std::runtime_error("truncated baked mesh: " + file);
    std::vector<Vec3> positions, normals;
    std::vector<Uv> tex_coords;
    std::vector<std::array<size_t, 3>> triangles;
    for (size_t i = 0; i < n[0]; ++i) positions.push_back(Vec3{(double)pos[3 * i], (double)pos[3 * i + 1], (double)pos[3 * i + 2]});
    for (size_t i = 0; i < n[1]; ++i) normals.push_back(Vec3{(double)nrm[3 * i], (double)nrm[3 * i + 1], (double)nrm[3 * i + 2]});
    for (size_t i = 0; i < n[2]; ++i) tex_coords.push_back(Uv{(double)uv[2 * i], (double)uv[2 * i + 1]});
    for (size_t i = 0; i < n[3]; ++i) triangles.push_back({idx[3 * i], idx[3 * i + 1], idx[3 * i + 2]});
    return std::make_shared<const MeshData>(std::move(positions), std::move(triangles), std::move(normals),
                                            std::move(tex_coords));
}

void MeshData::save_baked(const std::string& file) const {
    std::ofstream out(file, std::ios::binary);
    if (!out) throw std::runtime_error("cannot write " + file);
    uint32_t n[4] = {(uint32_t)positions_.size(), (uint32_t)normals_.size(), (uint32_t)tex_coords_.size(),
                     (uint32_t)triangles_.size()};
    out.write("PTMS", 4);
    out.write(reinterpret_cast<const char*>(n), sizeof n);
    auto put3 = [&](const std::vector<Vec3>& v) {
        for (const Vec3& p : v) {
            float f[3] = {(float)p.x, (float)p.y, (float)p.z};  // values came from f32: lossless
            out.write(reinterpret_cast<const char*>(f), sizeof f);
        }
    };
    put3(positions_);
    put3(normals_);
    for (const Uv& t : tex_coords_) {
        float f[2] = {(float)t.u, (float)t.v};
        out.write(reinterpret_cast<const char*>(f), sizeof f);
    }
    for (const auto& t : triangles_) {
        uint32_t i[3] = {(uint32_t)t[0], (uint32_t)t[1], (uint32_t)t[2]};
        out.write(reinterpret_cast<const char*>(i), sizeof i);
    }
}

std::shared_ptr<const MeshData> MeshData::load_obj(const std::string& path) {
    std::ifstream in(resolve_asset_path(path));
    if (!in) {
        if (auto baked = load_baked_mesh(resolve_baked_mesh_path(path))) return baked;
        throw std::runtime_error("cannot open OBJ file: " + path);
    }

    // tobj parses every coordinate as f32 (mesh.rs:41-49 widens them to f64)
    std::vector<float> pos, tex, nrm;
    std::vector<std::vector<Corner>> faces;
    bool first_model_done = false;
    std::string line;
    while (std::getline(in, line) && !first_model_done) {
        std::istringstream ss(line);
        std::string tag;
        if (!(ss >> tag)) continue;
        if (tag == "v") {
            float x, y, z;
            ss >> x >> y >> z;
            pos.insert(pos.end(), {x, y, z});
        } else if (tag == "vt") {
            float u = 0, v = 0;
            ss >> u >> v;
            tex.insert(tex.end(), {u, v});
        } else if (tag == "vn") {
            float x, y, z;
            ss >> x >> y >> z;
            nrm.insert(nrm.end(), {x, y, z});
        } else if (tag == "f") {
            std::vector<Corner> face;
            std::string tok;
            while (ss >> tok) {
                Corner c;
                if (parse_corner(tok, pos.size() / 3, tex.size() / 2, nrm.size() / 3, c)) face.push_back(c);
            }
            if (face.size() >= 3) faces.push_back(std::move(face));
        } else if (tag == "o" || tag == "g") {
            // a new object/group after faces were seen ends the first model (models[0], mesh.rs:59-60)
            if (!faces.empty()) first_model_done = true;
        }
    }

    // unify (v, vt, vn) corners in first-use order; fan-triangulate polygons
    std::map<Corner, size_t> index_of;
    std::vector<Vec3> positions, normals;
    std::vector<Uv> tex_coords;
    std::vector<std::array<size_t, 3>> triangles;
    auto add_vertex = [&](const Corner& c) -> size_t {
        auto it = index_of.find(c);
        if (it != index_of.end()) return it->second;
        size_t i = positions.size();
        index_of.emplace(c, i);
        positions.push_back(Vec3{(double)pos[3 * c.v], (double)pos[3 * c.v + 1], (double)pos[3 * c.v + 2]});
        if (c.vt >= 0 && !tex.empty()) tex_coords.push_back(Uv{(double)tex[2 * c.vt], (double)tex[2 * c.vt + 1]});
        if (c.vn >= 0 && !nrm.empty())
            normals.push_back(Vec3{(double)nrm[3 * c.vn], (double)nrm[3 * c.vn + 1], (double)nrm[3 * c.vn + 2]});
        return i;
    };
    for (const auto& f : faces) {
        for (size_t i = 1; i + 1 < f.size(); ++i) {
            size_t a = add_vertex(f[0]), b = add_vertex(f[i]), c = add_vertex(f[i + 1]);
            triangles.push_back({a, b, c});
        }
    }
    return std::make_shared<const MeshData>(std::move(positions), std::move(triangles), std::move(normals),
                                            std::move(tex_coords));
}

// ------------------------------------------------------------------ KDMesh
KDMesh::KDMesh(const MeshData& data, Shading shading) : KDMesh(data, shading, env_depth("KD_MESH_DEPTH", MAX_TREE_DEPTH)) {}

KDMesh::KDMesh(const MeshData& data, Shading shading, size_t max_tree_depth) {
    auto tree = std::make_shared<KDMeshTree>();
    tree->has_normals = shading == Shading::Smooth;
    tree->has_uvs = data.has_tex_coords();
    tree->tris = data.triangles(shading);
    const auto& tris = tree->tris;
    tree->root = build_kdtree(tris.size(), [&tris](size_t i) { return tris[i].bounds(); }, max_tree_depth);
    triangles = std::move(tree);
}

// ------------------------------------------------------------------ textures / assets
namespace {
std::mutex g_assets_mutex;
std::string g_assets_dir = "assets";
std::string g_baked_dir = "tests/golden/meshes";
TextureLoaderFn g_texture_loader = nullptr;
std::map<std::string, std::shared_ptr<RgbImageBuffer>> g_textures;

std::string basename_of(const std::string& path) {
    size_t p = path.find_last_of('/');
    return p == std::string::npos ? path : path.substr(p + 1);
}
}  // namespace

void set_assets_dir(const std::string& dir) {
    std::lock_guard<std::mutex> lock(g_assets_mutex);
    g_assets_dir = dir;
}

std::string resolve_asset_path(const std::string& path) {
    // the examples say "assets/xyz"; map that prefix onto the configured directory
    std::lock_guard<std::mutex> lock(g_assets_mutex);
    const std::string prefix = "assets/";
    if (path.compare(0, prefix.size(), prefix) == 0) return g_assets_dir + "/" + path.substr(prefix.size());
    return path;
}

void set_baked_mesh_dir(const std::string& dir) {
    std::lock_guard<std::mutex> lock(g_assets_mutex);
    g_baked_dir = dir;
}
void set_texture_loader(TextureLoaderFn fn) {
    std::lock_guard<std::mutex> lock(g_assets_mutex);
    g_texture_loader = fn;
}
std::string resolve_baked_mesh_path(const std::string& path) {
    std::string name = basename_of(path);
    size_t dot = name.find_last_of('.');
    if (dot != std::string::npos) name = name.substr(0, dot);
    std::lock_guard<std::mutex> lock(g_assets_mutex);
    return g_baked_dir + "/" + name + ".ptmesh";
}

void register_texture(const std::string& path, uint32_t width, uint32_t height, const uint8_t* rgb8) {
    auto buf = std::make_shared<RgbImageBuffer>();
    buf->width = width;
    buf->height = height;
    buf->data.assign(rgb8, rgb8 + static_cast<size_t>(width) * height * 3);
    buf->name = basename_of(path);
    std::lock_guard<std::mutex> lock(g_assets_mutex);
    g_textures[buf->name] = std::move(buf);
}

std::shared_ptr<RgbImageBuffer> RgbImageBuffer::open(const std::string& path) {
    const std::string key = basename_of(path);
    {
        std::lock_guard<std::mutex> lock(g_assets_mutex);
        auto it = g_textures.find(key);
        if (it != g_textures.end()) return it->second;
    }
    // ask the embedding process (Python + Pillow) to decode and register it
    TextureLoaderFn loader;
    {
        std::lock_guard<std::mutex> lock(g_assets_mutex);
        loader = g_texture_loader;
    }
    if (loader) {
        loader(path.c_str());
        std::lock_guard<std::mutex> lock(g_assets_mutex);
        auto it = g_textures.find(key);
        if (it != g_textures.end()) return it->second;
    }
    // decoded cache: "PTEX" u32 w u32 h, then RGB8
    const std::string raw = resolve_asset_path("assets/_decoded/" + key + ".ptex");
    std::ifstream in(raw, std::ios::binary);
    if (!in) throw std::runtime_error("texture not decoded/registered: " + path + " (expected " + raw + ")");
    char magic[4];
    uint32_t w = 0, h = 0;
    in.read(magic, 4);
    in.read(reinterpret_cast<char*>(&w), 4);
    in.read(reinterpret_cast<char*>(&h), 4);
    if (std::memcmp(magic, "PTEX", 4) != 0) throw std::runtime_error("bad decoded texture file: " + raw);
    auto buf = std::make_shared<RgbImageBuffer>();
    buf->width = w;
    buf->height = h;
    buf->data.resize(static_cast<size_t>(w) * h * 3);
    in.read(reinterpret_cast<char*>(buf->data.data()), buf->data.size());
    if (!in) throw std::runtime_error("truncated decoded texture file: " + raw);
    buf->name = key;
    std::lock_guard<std::mutex> lock(g_assets_mutex);
    g_textures[key] = buf;
    return buf;
}

}  // namespace portrayer
