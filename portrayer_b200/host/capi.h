/* C surface of the host mirror (libportrayer_host.so) for the Python test and
 * benchmark harness: build an example scene, get its packed blob, camera,
 * image size and background.  Not part of the drop-in boundary (that is
 * include/portrayer_gpu.h); this stands in for "the Rust example program". */
#ifndef PORTRAYER_HOST_CAPI_H
#define PORTRAYER_HOST_CAPI_H
#include <stdint.h>

#include "portrayer_gpu.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct PthScene PthScene;

const char* pth_last_error(void);
void pth_set_assets_dir(const char* dir);
void pth_register_texture(const char* path, uint32_t width, uint32_t height, const uint8_t* rgb8);

void pth_set_baked_mesh_dir(const char* dir);
typedef void (*PthTextureLoader)(const char* path);
void pth_set_texture_loader(PthTextureLoader fn);
/* parse an OBJ with the tobj-0.1.7-style loader and write its baked form; 0 on success */
int pth_bake_obj(const char* obj_path, const char* out_path);
/* number of triangles / unified vertices of an OBJ or baked mesh (for loader tests); -1 on error */
int64_t pth_mesh_info(const char* obj_path, uint64_t* n_positions, uint64_t* n_normals, uint64_t* n_uvs);

int pth_example_count(void);
const char* pth_example_name(int index);

/* kd_depth < 0: KD_DEPTH env or 10 (kdscene.rs:13,36-38). linear_tlas: one root leaf, flat order. */
PthScene* pth_example_build(const char* name, int64_t kd_depth, int linear_tlas);
PthScene* pth_big_scene_build(uint64_t n, int64_t kd_depth, int linear_tlas);
PthScene* pth_synthetic_instances_build(uint64_t n_instances, uint64_t seed, int64_t kd_depth);
PthScene* pth_synthetic_triangles_build(uint64_t n_triangles, uint64_t seed, int64_t kd_mesh_depth);
void pth_scene_free(PthScene* s);

uint64_t pth_blob_size(const PthScene* s);
const void* pth_blob_data(const PthScene* s);
void pth_image_size(const PthScene* s, uint32_t* width, uint32_t* height);
/* Camera::new for an arbitrary target size (camera.rs:34-45) */
void pth_camera(const PthScene* s, double width, double height, PtCamera* out);
/* background.at(x/w, y/h) for integer pixels; out = W*H*3 doubles */
void pth_background(const PthScene* s, uint32_t width, uint32_t height, double* out);
/* seconds spent in FlatScene::from + KDTreeScene::from + pack */
double pth_prepare_seconds(const PthScene* s);

/* world-space bounds of the scene tree's items = FlatSceneNode::bounds of every flat instance (flat_scene.rs:63-69),
 * n x {min x, y, z, max x, y, z}: the input of the scene-tree build (kdscene.rs:24-28) */
/* seconds spent in FlatScene::from alone (flat_scene.rs:18-46: products, inverses) */
double pth_flatten_seconds(const PthScene* s);
uint64_t pth_scene_item_count(const PthScene* s);
void pth_scene_item_bounds(const PthScene* s, double* out);

/* the scene GRAPH in the records pt_flatten takes (include/portrayer_gpu.h): sizes, then the arrays.
 * Returns -1 for the hand-built known-answer scenes, which have no graph. */
int pth_scene_hierarchy_sizes(const PthScene* s, uint32_t* n_nodes, uint32_t* n_children, uint32_t* n_geometries, uint32_t* root);
int pth_scene_hierarchy(const PthScene* s, PtHierNode* nodes_out, uint32_t* children_out, PtGeometryRec* geometries_out);

/* KDLeaf::partitioned on the host (the C++ mirror of leaf.rs:89-231) over n items given by their bounds,
 * serialised like the boundary's PtKdNode / leaf-item records: the comparison arm of pt_kd_build. */
typedef struct PthKdTree PthKdTree;
PthKdTree* pth_kd_build(const double* bounds, uint64_t n, uint32_t max_depth, uint32_t target_max_nodes,
                        int32_t target_max_merit, uint32_t max_tries);
void pth_kd_tree_free(PthKdTree* t);
uint64_t pth_kd_tree_node_count(const PthKdTree* t);
uint64_t pth_kd_tree_item_count(const PthKdTree* t);
uint32_t pth_kd_tree_depth(const PthKdTree* t);
double pth_kd_tree_extent(const PthKdTree* t);      /* BoundingBox::extent of the root */
double pth_kd_tree_build_seconds(const PthKdTree* t);
const PtKdNode* pth_kd_tree_nodes(const PthKdTree* t);
const uint32_t* pth_kd_tree_items(const PthKdTree* t);

/* Image::render through the C++ mirror (the call an example's main() makes). rgb_inout = W*H*3. */
int pth_image_render(const PthScene* s, uint32_t width, uint32_t height, uint32_t samples, uint32_t rng_mode,
                     uint64_t seed, uint8_t* rgb_inout, PtStats* stats);
#ifdef __cplusplus
}
#endif
#endif
