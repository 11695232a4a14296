#include "capi.h"

#include <chrono>
#include <cstring>
#include <string>

#include "assets.hpp"
#include "examples/examples.hpp"
#include "pack.hpp"

using namespace portrayer;

#include "capi_internal.hpp"

struct PthKdTree {
    std::vector<PtKdNode> nodes;
    std::vector<uint32_t> items;
    uint32_t depth = 0;
    double extent = 0.0;
    double seconds = 0.0;
};

namespace portrayer {
std::string& capi_error() {
    thread_local std::string error;
    return error;
}
}  // namespace portrayer

namespace {
#define g_error (portrayer::capi_error())

PthScene* finish(ExampleScene ex, int64_t kd_depth, int linear_tlas) {
    auto out = std::make_unique<PthScene>();
    auto t0 = std::chrono::steady_clock::now();
    auto keep_bounds = [&out](const std::vector<FlatSceneNode>& nodes) {
        out->item_bounds.reserve(nodes.size() * 6);
        for (const FlatSceneNode& n : nodes) {
            const BoundingBox b = n.bounds();
            for (double v : {b.min().x, b.min().y, b.min().z, b.max().x, b.max().y, b.max().z}) out->item_bounds.push_back(v);
        }
    };
    if (ex.prebuilt) {
        out->blob = pack_scene(*ex.prebuilt);
        keep_bounds(ex.prebuilt->nodes);
    } else {
        FlatScene flat = FlatScene::from(ex.scene);
        out->flatten_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (linear_tlas) {
            // one unpartitioned leaf: the flat_scene feature's linear fold (flat_scene.rs:71-99, ray.rs:87-99)
            KDTreeScene kd = KDTreeScene::from(std::move(flat), 0);
            out->blob = pack_scene(kd);
            keep_bounds(kd.nodes);
        } else {
            KDTreeScene kd = kd_depth < 0 ? KDTreeScene::from(std::move(flat))
                                          : KDTreeScene::from(std::move(flat), static_cast<size_t>(kd_depth));
            out->blob = pack_scene(kd);
            keep_bounds(kd.nodes);
        }
    }
    out->prepare_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    out->example = std::move(ex);
    return out.release();
}

template <class F>
PthScene* guarded(F&& f) {
    try {
        return f();
    } catch (const std::exception& e) {
        g_error = e.what();
        return nullptr;
    }
}
}  // namespace

extern "C" {

const char* pth_last_error(void) { return g_error.c_str(); }
void pth_set_assets_dir(const char* dir) { set_assets_dir(dir); }
void pth_register_texture(const char* path, uint32_t width, uint32_t height, const uint8_t* rgb8) {
    register_texture(path, width, height, rgb8);
}

void pth_set_baked_mesh_dir(const char* dir) { set_baked_mesh_dir(dir); }
void pth_set_texture_loader(PthTextureLoader fn) { set_texture_loader(fn); }
int pth_bake_obj(const char* obj_path, const char* out_path) {
    try {
        MeshData::load_obj(obj_path)->save_baked(out_path);
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return -1;
    }
}
int64_t pth_mesh_info(const char* obj_path, uint64_t* n_positions, uint64_t* n_normals, uint64_t* n_uvs) {
    try {
        auto m = MeshData::load_obj(obj_path);
        *n_positions = m->num_positions();
        *n_normals = m->num_normals();
        *n_uvs = m->has_tex_coords() ? m->num_positions() : 0;
        return static_cast<int64_t>(m->num_triangles());
    } catch (const std::exception& e) {
        g_error = e.what();
        return -1;
    }
}

int pth_example_count(void) { return static_cast<int>(example_registry().size()); }
const char* pth_example_name(int index) {
    int i = 0;
    for (const auto& kv : example_registry())
        if (i++ == index) return kv.first.c_str();
    return nullptr;
}

PthScene* pth_example_build(const char* name, int64_t kd_depth, int linear_tlas) {
    return guarded([&]() -> PthScene* {
        auto it = example_registry().find(name);
        if (it == example_registry().end()) throw std::runtime_error(std::string("unknown example: ") + name);
        return finish(it->second(), kd_depth, linear_tlas);
    });
}
PthScene* pth_big_scene_build(uint64_t n, int64_t kd_depth, int linear_tlas) {
    return guarded([&] { return finish(make_big_scene(n), kd_depth, linear_tlas); });
}
PthScene* pth_synthetic_instances_build(uint64_t n_instances, uint64_t seed, int64_t kd_depth) {
    return guarded([&] { return finish(make_synthetic_instances(n_instances, seed), kd_depth, 0); });
}
PthScene* pth_synthetic_triangles_build(uint64_t n_triangles, uint64_t seed, int64_t kd_mesh_depth) {
    return guarded([&] {
        return finish(make_synthetic_triangles(n_triangles, seed, static_cast<size_t>(kd_mesh_depth)), -1, 0);
    });
}
void pth_scene_free(PthScene* s) { delete s; }

double pth_flatten_seconds(const PthScene* s) { return s->flatten_seconds; }
uint64_t pth_scene_item_count(const PthScene* s) { return s->item_bounds.size() / 6; }
void pth_scene_item_bounds(const PthScene* s, double* out) {
    std::copy(s->item_bounds.begin(), s->item_bounds.end(), out);
}

static const HierarchyExport* hierarchy_of(const PthScene* s) {
    if (s->example.prebuilt) return nullptr;
    auto* m = const_cast<PthScene*>(s);
    if (!m->hierarchy) m->hierarchy = std::make_unique<HierarchyExport>(export_hierarchy(s->example.scene));
    return m->hierarchy.get();
}
int pth_scene_hierarchy_sizes(const PthScene* s, uint32_t* n_nodes, uint32_t* n_children, uint32_t* n_geometries, uint32_t* root) {
    const HierarchyExport* h = hierarchy_of(s);
    if (!h) return -1;
    *n_nodes = static_cast<uint32_t>(h->nodes.size());
    *n_children = static_cast<uint32_t>(h->children.size());
    *n_geometries = static_cast<uint32_t>(h->geometries.size());
    *root = h->root;
    return 0;
}
int pth_scene_hierarchy(const PthScene* s, PtHierNode* nodes_out, uint32_t* children_out, PtGeometryRec* geometries_out) {
    const HierarchyExport* h = hierarchy_of(s);
    if (!h) return -1;
    std::copy(h->nodes.begin(), h->nodes.end(), nodes_out);
    std::copy(h->children.begin(), h->children.end(), children_out);
    std::copy(h->geometries.begin(), h->geometries.end(), geometries_out);
    return 0;
}

PthKdTree* pth_kd_build(const double* bounds, uint64_t n, uint32_t max_depth, uint32_t target_max_nodes,
                        int32_t target_max_merit, uint32_t max_tries) {
    try {
        auto out = std::make_unique<PthKdTree>();
        auto t0 = std::chrono::steady_clock::now();
        PartitionConfig conf{target_max_nodes, target_max_merit, max_tries};
        auto root = build_kdtree(static_cast<size_t>(n), [bounds](size_t i) {
            const double* b = bounds + i * 6;
            return BoundingBox(Vec3{b[0], b[1], b[2]}, Vec3{b[3], b[4], b[5]});
        }, max_depth, conf);
        out->depth = serialise_kd_tree(*root, out->nodes, out->items);
        out->extent = root->extent();
        out->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        return out.release();
    } catch (const std::exception& e) {
        g_error = e.what();
        return nullptr;
    }
}
void pth_kd_tree_free(PthKdTree* t) { delete t; }
uint64_t pth_kd_tree_node_count(const PthKdTree* t) { return t->nodes.size(); }
uint64_t pth_kd_tree_item_count(const PthKdTree* t) { return t->items.size(); }
uint32_t pth_kd_tree_depth(const PthKdTree* t) { return t->depth; }
double pth_kd_tree_extent(const PthKdTree* t) { return t->extent; }
double pth_kd_tree_build_seconds(const PthKdTree* t) { return t->seconds; }
const PtKdNode* pth_kd_tree_nodes(const PthKdTree* t) { return t->nodes.data(); }
const uint32_t* pth_kd_tree_items(const PthKdTree* t) { return t->items.data(); }

uint64_t pth_blob_size(const PthScene* s) { return s->blob.size(); }
const void* pth_blob_data(const PthScene* s) { return s->blob.data(); }
void pth_image_size(const PthScene* s, uint32_t* width, uint32_t* height) {
    *width = static_cast<uint32_t>(s->example.width);
    *height = static_cast<uint32_t>(s->example.height);
}
void pth_camera(const PthScene* s, double width, double height, PtCamera* out) {
    *out = make_camera(s->example.cam, width, height);
}
void pth_background(const PthScene* s, uint32_t width, uint32_t height, double* out) {
    for (uint32_t y = 0; y < height; ++y)
        for (uint32_t x = 0; x < width; ++x) {
            Rgb c = s->example.background(Uv{(double)x / (double)width, (double)y / (double)height});
            double* p = out + (static_cast<size_t>(y) * width + x) * 3;
            p[0] = c.r; p[1] = c.g; p[2] = c.b;
        }
}
double pth_prepare_seconds(const PthScene* s) { return s->prepare_seconds; }

}  // extern "C"
