#include "pack.hpp"

#include <cstring>
#include <deque>
#include <map>
#include <stdexcept>

namespace portrayer {

PtCamera make_camera(const CameraSettings& cam, double width, double height) {
    PtCamera out{};
    out.eye[0] = cam.eye.x; out.eye[1] = cam.eye.y; out.eye[2] = cam.eye.z;
    // Need to invert because look_at returns a world-to-view matrix. camera.rs:38
    Mat4 v2w = Mat4::look_at_rh(cam.eye, cam.center, cam.up).inverted();
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) out.view_to_world[r * 4 + c] = v2w.m[r][c];
    out.fov_factor = std::tan(cam.fovy.get() / 2.0);  // camera.rs:40
    out.aspect_ratio = width / height;                 // camera.rs:41
    out.width = width;
    out.height = height;
    return out;
}

namespace {

void rows3x4(const Mat4& m, double* out) {
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c) out[r * 4 + c] = m.m[r][c];
}

// Breadth-first serialisation so the top levels of a tree are contiguous in
// memory (they are the part every ray touches).
struct TreeOut {
    std::vector<PtKdNode> nodes;
    std::vector<uint32_t> items;
    uint32_t depth = 0;
};

TreeOut serialise_tree(const KDIndexTree& root) {
    TreeOut out;
    struct Pending { const KDIndexTree* node; uint32_t index; uint32_t depth; };
    std::deque<Pending> queue;
    out.nodes.push_back(PtKdNode{});
    queue.push_back({&root, 0, 0});
    while (!queue.empty()) {
        Pending p = queue.front();
        queue.pop_front();
        if (p.depth > out.depth) out.depth = p.depth;
        PtKdNode rec{};
        if (p.node->is_leaf) {
            const uint32_t first = static_cast<uint32_t>(out.items.size());
            for (const auto& nb : p.node->leaf.nodes) out.items.push_back(nb->node);
            rec.split = 0.0;
            rec.a = 3u | (first << 2);
            rec.b = static_cast<uint32_t>(p.node->leaf.nodes.size());
        } else {
            const uint32_t front = static_cast<uint32_t>(out.nodes.size());
            out.nodes.push_back(PtKdNode{});
            const uint32_t back = static_cast<uint32_t>(out.nodes.size());
            out.nodes.push_back(PtKdNode{});
            if (front >= (1u << 30)) throw std::runtime_error("kd-tree too large");
            queue.push_back({p.node->front_nodes.get(), front, p.depth + 1});
            queue.push_back({p.node->back_nodes.get(), back, p.depth + 1});
            rec.split = p.node->plane;
            rec.a = static_cast<uint32_t>(p.node->axis) | (front << 2);
            rec.b = back;
        }
        out.nodes[p.index] = rec;
    }
    return out;
}

}  // namespace

uint32_t serialise_kd_tree(const KDIndexTree& root, std::vector<PtKdNode>& nodes_out, std::vector<uint32_t>& items_out) {
    TreeOut t = serialise_tree(root);
    nodes_out = std::move(t.nodes);
    items_out = std::move(t.items);
    return t.depth;
}

namespace {

struct Builder {
    std::vector<PtKdNode> tlas_nodes, blas_nodes;
    std::vector<uint32_t> tlas_items, blas_items;
    std::vector<PtInstance> instances;
    std::vector<PtInstanceTrans> instance_trans;
    std::vector<PtMesh> meshes;
    std::vector<PtTriPos> tri_pos;
    std::vector<PtTriNormals> tri_normals;
    std::vector<PtTriUvs> tri_uvs;
    std::vector<PtMaterial> materials;
    std::vector<PtLight> lights;
    std::vector<PtTexture> textures;
    std::vector<uint8_t> texels;

    std::map<const void*, uint32_t> material_ids, texture_ids;
    std::map<std::pair<const void*, int>, uint32_t> mesh_ids;

    int32_t texture_id(const std::shared_ptr<RgbImageBuffer>& buf) {
        auto it = texture_ids.find(buf.get());
        if (it != texture_ids.end()) return static_cast<int32_t>(it->second);
        // keep every texture 128-byte aligned in the pool
        while (texels.size() % 128) texels.push_back(0);
        PtTexture t{buf->width, buf->height, texels.size(), buf->content_key(), 0};
        texels.insert(texels.end(), buf->data.begin(), buf->data.end());
        uint32_t id = static_cast<uint32_t>(textures.size());
        textures.push_back(t);
        texture_ids[buf.get()] = id;
        return static_cast<int32_t>(id);
    }

    uint32_t material_id(const MaterialRef& m) {
        auto it = material_ids.find(m.get());
        if (it != material_ids.end()) return it->second;
        PtMaterial r{};
        r.diffuse[0] = m->diffuse.r; r.diffuse[1] = m->diffuse.g; r.diffuse[2] = m->diffuse.b;
        r.specular[0] = m->specular.r; r.specular[1] = m->specular.g; r.specular[2] = m->specular.b;
        r.shininess = m->shininess;
        r.reflectivity = m->reflectivity;
        r.glossy_side_length = m->glossy_side_length;
        r.refraction_index = m->refraction_index;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) r.uv_trans[i * 3 + j] = m->uv_trans.m[i][j];
        r.texture = m->texture ? texture_id(m->texture->buffer) : -1;
        r.normals = m->normals ? texture_id(m->normals->buffer) : -1;
        uint32_t id = static_cast<uint32_t>(materials.size());
        materials.push_back(r);
        material_ids[m.get()] = id;
        return id;
    }

    void add_triangles(const std::vector<Triangle>& tris, PtMesh& rec) {
        rec.tri_first = static_cast<uint32_t>(tri_pos.size());
        rec.tri_count = static_cast<uint32_t>(tris.size());
        rec.nrm_first = 0xFFFFFFFFu;
        rec.uv_first = 0xFFFFFFFFu;
        const bool normals = !tris.empty() && tris[0].normals.has_value();
        const bool uvs = !tris.empty() && tris[0].tex_coords.has_value();
        if (normals) { rec.flags |= PT_MESH_FLAG_NORMALS; rec.nrm_first = static_cast<uint32_t>(tri_normals.size()); }
        if (uvs) { rec.flags |= PT_MESH_FLAG_UVS; rec.uv_first = static_cast<uint32_t>(tri_uvs.size()); }
        for (const Triangle& t : tris) {
            tri_pos.push_back(PtTriPos{{t.a.x, t.a.y, t.a.z}, {t.b.x, t.b.y, t.b.z}, {t.c.x, t.c.y, t.c.z}});
            if (normals) {
                const auto& n = *t.normals;
                tri_normals.push_back(
                    PtTriNormals{{n[0].x, n[0].y, n[0].z}, {n[1].x, n[1].y, n[1].z}, {n[2].x, n[2].y, n[2].z}});
            }
            if (uvs) {
                const auto& u = *t.tex_coords;
                tri_uvs.push_back(PtTriUvs{{u[0].u, u[0].v}, {u[1].u, u[1].v}, {u[2].u, u[2].v}});
            }
        }
    }

    uint32_t mesh_id(const Primitive& p) {
        const void* key_ptr = p.kind == PrimKind::Mesh       ? static_cast<const void*>(p.mesh.get())
                              : p.kind == PrimKind::KDMesh   ? static_cast<const void*>(p.kdmesh.get())
                                                             : static_cast<const void*>(p.triangle.get());
        const int key_tag = p.kind == PrimKind::Mesh ? (p.shading == Shading::Smooth ? 1 : 0) : 2;
        auto key = std::make_pair(key_ptr, key_tag);
        auto it = mesh_ids.find(key);
        if (it != mesh_ids.end()) return it->second;

        PtMesh rec{};
        BoundingBox bounds;
        if (p.kind == PrimKind::Mesh) {
            rec.kind = PT_MESH_LINEAR;
            add_triangles(p.mesh->triangles(p.shading), rec);
            bounds = p.mesh->bounds();  // MeshData.bounds over ALL positions, mesh.rs:66-70,153
        } else if (p.kind == PrimKind::KDMesh) {
            rec.kind = PT_MESH_KD;
            add_triangles(p.kdmesh->tris, rec);
            TreeOut tree = serialise_tree(*p.kdmesh->root);
            rec.node_first = static_cast<uint32_t>(blas_nodes.size());
            rec.node_count = static_cast<uint32_t>(tree.nodes.size());
            rec.item_first = static_cast<uint32_t>(blas_items.size());
            rec.item_count = static_cast<uint32_t>(tree.items.size());
            rec.kd_depth = tree.depth;
            blas_nodes.insert(blas_nodes.end(), tree.nodes.begin(), tree.nodes.end());
            blas_items.insert(blas_items.end(), tree.items.begin(), tree.items.end());
            bounds = p.kdmesh->root->node_bounds();  // kdmesh.rs:26-30,67
            rec.extent = p.kdmesh->root->extent();   // node.rs:39
        } else {
            rec.kind = PT_MESH_TRIANGLE;
            add_triangles({*p.triangle}, rec);
            bounds = p.triangle->bounds();
        }
        rows3x4(bounds.invtrans(), rec.bbox_invtrans);
        uint32_t id = static_cast<uint32_t>(meshes.size());
        meshes.push_back(rec);
        mesh_ids[key] = id;
        return id;
    }

    void add_instance(const FlatSceneNode& n) {
        PtInstance rec{};
        rows3x4(n.invtrans, rec.invtrans);
        rec.mesh = 0xFFFFFFFFu;
        switch (n.geometry.primitive.kind) {
            case PrimKind::Sphere: rec.prim = PT_PRIM_SPHERE; break;
            case PrimKind::Plane: rec.prim = PT_PRIM_PLANE; break;
            case PrimKind::Cube: rec.prim = PT_PRIM_CUBE; break;
            case PrimKind::Cylinder: rec.prim = PT_PRIM_CYLINDER; break;
            case PrimKind::Cone: rec.prim = PT_PRIM_CONE; break;
            case PrimKind::Triangle: rec.prim = PT_PRIM_TRIANGLE; rec.mesh = mesh_id(n.geometry.primitive); break;
            case PrimKind::Mesh: rec.prim = PT_PRIM_MESH; rec.mesh = mesh_id(n.geometry.primitive); break;
            case PrimKind::KDMesh: rec.prim = PT_PRIM_KDMESH; rec.mesh = mesh_id(n.geometry.primitive); break;
        }
        rec.material = material_id(n.geometry.material);
        instances.push_back(rec);
        PtInstanceTrans tr{};
        rows3x4(n.trans, tr.trans);
        instance_trans.push_back(tr);
    }
};

}  // namespace

std::vector<uint8_t> pack_scene(const std::vector<FlatSceneNode>& nodes, const KDIndexTree& root,
                                const std::vector<Light>& lights, Rgb ambient) {
    Builder b;
    for (const auto& n : nodes) b.add_instance(n);
    TreeOut tlas = serialise_tree(root);
    b.tlas_nodes = std::move(tlas.nodes);
    b.tlas_items = std::move(tlas.items);
    for (const Light& l : lights) {
        PtLight r{};
        r.position[0] = l.position.x; r.position[1] = l.position.y; r.position[2] = l.position.z;
        r.color[0] = l.color.r; r.color[1] = l.color.g; r.color[2] = l.color.b;
        r.falloff[0] = l.falloff.c0; r.falloff[1] = l.falloff.c1; r.falloff[2] = l.falloff.c2;
        r.area_a[0] = l.area.a.x; r.area_a[1] = l.area.a.y; r.area_a[2] = l.area.a.z;
        r.area_b[0] = l.area.b.x; r.area_b[1] = l.area.b.y; r.area_b[2] = l.area.b.z;
        b.lights.push_back(r);
    }

    PtSceneDesc d{};
    d.ambient[0] = ambient.r; d.ambient[1] = ambient.g; d.ambient[2] = ambient.b;
    d.tlas_extent = root.extent();  // node.rs:29
    d.tlas_depth = tlas.depth;
    d.n_tlas_nodes = static_cast<uint32_t>(b.tlas_nodes.size());   d.tlas_nodes = b.tlas_nodes.data();
    d.n_tlas_items = static_cast<uint32_t>(b.tlas_items.size());   d.tlas_items = b.tlas_items.data();
    d.n_instances = static_cast<uint32_t>(b.instances.size());     d.instances = b.instances.data();
    d.instance_trans = b.instance_trans.data();
    d.n_meshes = static_cast<uint32_t>(b.meshes.size());           d.meshes = b.meshes.data();
    d.n_blas_nodes = static_cast<uint32_t>(b.blas_nodes.size());   d.blas_nodes = b.blas_nodes.data();
    d.n_blas_items = static_cast<uint32_t>(b.blas_items.size());   d.blas_items = b.blas_items.data();
    d.n_triangles = static_cast<uint32_t>(b.tri_pos.size());       d.tri_pos = b.tri_pos.data();
    d.n_tri_normals = static_cast<uint32_t>(b.tri_normals.size()); d.tri_normals = b.tri_normals.data();
    d.n_tri_uvs = static_cast<uint32_t>(b.tri_uvs.size());         d.tri_uvs = b.tri_uvs.data();
    d.n_materials = static_cast<uint32_t>(b.materials.size());     d.materials = b.materials.data();
    d.n_lights = static_cast<uint32_t>(b.lights.size());           d.lights = b.lights.data();
    d.n_textures = static_cast<uint32_t>(b.textures.size());       d.textures = b.textures.data();
    d.n_texel_bytes = b.texels.size();                             d.texels = b.texels.data();

    std::vector<uint8_t> blob(pt_scene_blob_size(&d));
    int rc = pt_scene_pack(&d, blob.data(), blob.size());
    if (rc != PT_OK) throw std::runtime_error("pt_scene_pack failed");
    return blob;
}

HierarchyExport export_hierarchy(const HierScene& scene) {
    // ids exactly as Builder::add_instance hands them out: walk the flat instances in order
    FlatScene flat = FlatScene::from(scene);
    std::map<const void*, uint32_t> material_ids;
    std::map<std::pair<const void*, int>, uint32_t> mesh_ids;
    auto mesh_key = [](const Primitive& p) {
        const void* key_ptr = p.kind == PrimKind::Mesh       ? static_cast<const void*>(p.mesh.get())
                              : p.kind == PrimKind::KDMesh   ? static_cast<const void*>(p.kdmesh.get())
                                                             : static_cast<const void*>(p.triangle.get());
        const int key_tag = p.kind == PrimKind::Mesh ? (p.shading == Shading::Smooth ? 1 : 0) : 2;
        return std::make_pair(key_ptr, key_tag);
    };
    auto is_mesh_kind = [](PrimKind k) { return k == PrimKind::Triangle || k == PrimKind::Mesh || k == PrimKind::KDMesh; };
    for (const FlatSceneNode& n : flat.root) {
        const Primitive& p = n.geometry.primitive;
        if (is_mesh_kind(p.kind)) mesh_ids.emplace(mesh_key(p), static_cast<uint32_t>(mesh_ids.size()));
        material_ids.emplace(n.geometry.material.get(), static_cast<uint32_t>(material_ids.size()));
    }

    HierarchyExport out;
    std::map<const SceneNode*, uint32_t> node_ids;
    std::deque<const SceneNode*> queue;
    auto id_of = [&](const SceneNode* n) {
        auto it = node_ids.find(n);
        if (it != node_ids.end()) return it->second;
        const uint32_t id = static_cast<uint32_t>(out.nodes.size());
        node_ids[n] = id;
        out.nodes.push_back(PtHierNode{});
        queue.push_back(n);
        return id;
    };
    out.root = id_of(scene.root.get());
    while (!queue.empty()) {
        const SceneNode* n = queue.front();
        queue.pop_front();
        PtHierNode rec{};
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) rec.trans[r * 4 + c] = n->trans().m[r][c];
        rec.geometry = 0xFFFFFFFFu;
        if (n->geometry()) {
            const Geometry& g = *n->geometry();
            PtGeometryRec gr{};
            const BoundingBox b = g.primitive.bounds();
            gr.bounds[0] = b.min().x; gr.bounds[1] = b.min().y; gr.bounds[2] = b.min().z;
            gr.bounds[3] = b.max().x; gr.bounds[4] = b.max().y; gr.bounds[5] = b.max().z;
            switch (g.primitive.kind) {
                case PrimKind::Sphere: gr.prim = PT_PRIM_SPHERE; break;
                case PrimKind::Plane: gr.prim = PT_PRIM_PLANE; break;
                case PrimKind::Cube: gr.prim = PT_PRIM_CUBE; break;
                case PrimKind::Cylinder: gr.prim = PT_PRIM_CYLINDER; break;
                case PrimKind::Cone: gr.prim = PT_PRIM_CONE; break;
                case PrimKind::Triangle: gr.prim = PT_PRIM_TRIANGLE; break;
                case PrimKind::Mesh: gr.prim = PT_PRIM_MESH; break;
                case PrimKind::KDMesh: gr.prim = PT_PRIM_KDMESH; break;
            }
            gr.mesh = is_mesh_kind(g.primitive.kind) ? mesh_ids.at(mesh_key(g.primitive)) : 0xFFFFFFFFu;
            gr.material = material_ids.at(g.material.get());
            rec.geometry = static_cast<uint32_t>(out.geometries.size());
            out.geometries.push_back(gr);
        }
        rec.first_child = static_cast<uint32_t>(out.children.size());
        rec.child_count = static_cast<uint32_t>(n->children().size());
        for (const NodeRef& c : n->children()) out.children.push_back(id_of(c.get()));
        out.nodes[node_ids[n]] = rec;
    }
    return out;
}

std::vector<uint8_t> pack_scene(const KDTreeScene& scene) {
    return pack_scene(scene.nodes, *scene.root, scene.lights, scene.ambient);
}

}  // namespace portrayer
