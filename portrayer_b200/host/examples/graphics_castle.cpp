// examples/graphics-castle.rs — "The Computer Graphics Castle", BASELINE.json configs[4]
// (3840x2160, SAMPLES=64, tiles across 1/2/4/8 GPUs).  13 KDMesh instances, dielectric windows /
// glass ceiling / glossy lake, normal-mapped door + tapestries + dock, and an outdoor maze of
// textured cubes grown by a randomised depth-first walk (StdRng::seed_from_u64(19392103958),
// examples/graphics-castle.rs:277-470).  assets/shrub.png is missing upstream
// (.MISSING_LARGE_BLOBS): the asset layer substitutes a procedural stand-in and says so.
#include <deque>
#include <optional>
#include <set>

#include "../rand07.hpp"
#include "examples.hpp"
using namespace portrayer;

namespace {

using CellPos = std::pair<size_t, size_t>;
using MaybeCell = std::optional<CellPos>;

enum class Cell { Empty, Wall };

// struct Maze, graphics-castle.rs:359-470
struct Maze {
    std::vector<std::vector<Cell>> cells;  // rows of the maze, stored row-wise

    Maze(size_t rows, size_t cols) : cells(rows, std::vector<Cell>(cols, Cell::Wall)) {}

    // inclusive on both ends (:377-383)
    void reserve(CellPos a, CellPos b) {
        for (size_t row = a.first; row <= b.first; ++row)
            for (size_t col = a.second; col <= b.second; ++col) cells[row][col] = Cell::Empty;
    }

    // :386-469
    void fill_maze(CellPos start) {
        const size_t rows = cells.size(), cols = cells[0].size();
        // leave the first and last row / column untouched
        auto find_adjacents = [&](std::vector<MaybeCell>& adj, size_t row, size_t col) {
            adj[0] = row > 1 ? MaybeCell({row - 1, col}) : std::nullopt;
            adj[1] = row < rows - 2 ? MaybeCell({row + 1, col}) : std::nullopt;
            adj[2] = col > 1 ? MaybeCell({row, col - 1}) : std::nullopt;
            adj[3] = col < cols - 2 ? MaybeCell({row, col + 1}) : std::nullopt;
        };
        auto find_diagonal_adjacents = [&](std::vector<MaybeCell>& adj, size_t row, size_t col) {
            adj[0] = (row > 1 && col > 1) ? MaybeCell({row - 1, col - 1}) : std::nullopt;
            adj[1] = (row < rows - 2 && col > 1) ? MaybeCell({row + 1, col - 1}) : std::nullopt;
            adj[2] = (row > 1 && col < cols - 2) ? MaybeCell({row - 1, col + 1}) : std::nullopt;
            adj[3] = (row < rows - 2 && col < cols - 2) ? MaybeCell({row + 1, col + 1}) : std::nullopt;
        };
        auto count_empty = [&](const std::vector<MaybeCell>& adj) {
            size_t n = 0;
            for (const auto& a : adj)
                if (a && cells[a->first][a->second] == Cell::Empty) ++n;
            return n;
        };

        // Want a random maze but want the same one every time
        StdRng rng = StdRng::seed_from_u64(19392103958ull);
        std::vector<MaybeCell> adjacents(4);
        std::deque<CellPos> walls;
        std::set<CellPos> seen;

        cells[start.first][start.second] = Cell::Empty;
        find_adjacents(adjacents, start.first, start.second);
        for (const auto& a : adjacents)
            if (a) walls.push_back(*a);

        while (!walls.empty()) {
            const CellPos cur = walls.front();
            walls.pop_front();
            if (seen.count(cur)) continue;
            seen.insert(cur);
            const size_t row = cur.first, col = cur.second;
            if (cells[row][col] == Cell::Empty) continue;  // probably reserved

            // diagonal lines of empty cells look ugly
            find_diagonal_adjacents(adjacents, row, col);
            if (count_empty(adjacents) > 1) continue;
            find_adjacents(adjacents, row, col);
            if (count_empty(adjacents) > 1) continue;  // no loops

            cells[row][col] = Cell::Empty;

            // add its adjacent walls to the queue in a random order; depth first for longer paths
            rng.shuffle(adjacents);
            bool first = true;
            for (const auto& a : adjacents) {
                if (!a || cells[a->first][a->second] != Cell::Wall) continue;
                if (first) {
                    walls.push_front(*a);
                    first = false;
                } else {
                    walls.push_back(*a);
                }
            }
        }
    }
};

std::shared_ptr<const MeshData> obj(const char* name) { return MeshData::load_obj(std::string("assets/") + name); }

SceneNode castle() {
    auto mat_castle_walls = Arc(Material{.diffuse = {0.25, 0.25, 0.25}});
    auto wood = ImageTexture::open("assets/old_planks_02_diff_1k.png");
    auto wood_normals = NormalMap::open("assets/old_planks_02_nor_1k.png");
    auto mat_castle_door = Arc(Material{.texture = wood, .normals = wood_normals});
    auto mat_castle_window_frames = Arc(Material{.diffuse = {0.132866, 0.132866, 0.132866}});
    auto mat_ceiling_glass = Arc(Material{.diffuse = {0.147337, 0.239555, 0.034547}, .specular = {0.3, 0.3, 0.3}, .shininess = 100.0,
                                          .reflectivity = 0.8, .refraction_index = WINDOW_GLASS_REFRACTION_INDEX});
    auto mat_window_glass = Arc(Material{.diffuse = {0.147337, 0.239555, 0.034547}, .specular = {0.3, 0.3, 0.3}, .shininess = 100.0,
                                         .reflectivity = 1.0, .refraction_index = WINDOW_GLASS_REFRACTION_INDEX});
    auto mat_stairs_side = Arc(Material{.diffuse = {0.132866, 0.132866, 0.132866}, .specular = {0.3, 0.3, 0.3}, .shininess = 25.0});
    auto mat_tapestry = Arc(Material{.texture = wood, .normals = wood_normals});
    auto mat_puppet = Arc(Material{.diffuse = {0.06998, 0.06998, 0.06998}, .specular = {0.3, 0.3, 0.3}, .shininess = 25.0});

    auto castle_model = obj("castle.obj");
    auto castle_window_frames_model = obj("castle_window_frames.obj");
    auto castle_glass_ceilings_model = obj("castle_glass_ceilings.obj");
    auto castle_door_model = obj("castle_door.obj");
    auto castle_door_arch_model = obj("castle_door_arch.obj");
    auto castle_tapestry_model = obj("castle_tapestry.obj");
    auto castle_stairs_side_model = obj("castle_stairs_side.obj");
    KDMesh castle_stairs_side(*castle_stairs_side_model, Shading::Flat);
    auto puppet_left = obj("puppet_castle_left_tower.obj");
    auto puppet_right = obj("puppet_castle_right_tower.obj");

    auto window = [&](Vec3 scale, Vec3 at) {
        return SceneNode::from(Geometry(Cube{}, mat_window_glass)).scaled(scale).rotated_x(Radians::from_degrees(90.0)).translated(at).into();
    };
    return SceneNode::from(std::vector<NodeRef>{
        // Main castle body
        SceneNode::from(Geometry(KDMesh(*castle_model, Shading::Flat), mat_castle_walls)).translated({0.0, 30.0, -30.0}).into(),
        SceneNode::from(Geometry(KDMesh(*castle_window_frames_model, Shading::Flat), mat_castle_window_frames))
            .translated({0.0, 83.5746, -2.25}).into(),
        SceneNode::from(Geometry(KDMesh(*castle_glass_ceilings_model, Shading::Flat), mat_ceiling_glass))
            .translated({0.0, 96.0, -23.0}).into(),
        // Windows
        window({9.1, 1.0, 12.7}, {-30.0, 70.7, 12.7}),
        window({9.1, 1.0, 12.7}, {30.0, 70.7, 12.7}),
        window({13.4, 1.0, 18.8}, {0.0, 79.4, -2.9}),
        // Door
        SceneNode::from(Geometry(KDMesh(*castle_door_model, Shading::Flat), mat_castle_door)).translated({0.0, 21.739681, 10.0}).into(),
        SceneNode::from(Geometry(KDMesh(*castle_door_arch_model, Shading::Flat), mat_castle_door)).translated({0.0, 42.0, 9.0}).into(),
        // Stairs
        SceneNode::from(Geometry(castle_stairs_side, mat_stairs_side)).translated({-11.0, 5.0, 19.0}).into(),
        SceneNode::from(Geometry(castle_stairs_side, mat_stairs_side)).translated({11.0, 5.0, 19.0}).into(),
        // Statues / Guardians
        SceneNode::from(Geometry(KDMesh(*puppet_left, Shading::Smooth), mat_puppet)).translated({30.0, 33.6, 19.0}).into(),
        SceneNode::from(Geometry(Cylinder{}, mat_castle_walls)).scaled(10.0).translated({30.0, 5.0, 20.0}).into(),
        SceneNode::from(Geometry(KDMesh(*puppet_right, Shading::Smooth), mat_puppet)).translated({-30.0, 33.6, 19.0}).into(),
        SceneNode::from(Geometry(Cylinder{}, mat_castle_walls)).scaled(10.0).translated({-30.0, 5.0, 20.0}).into(),
        // Tapestries
        SceneNode::from(Geometry(KDMesh(*castle_tapestry_model, Shading::Smooth), mat_tapestry)).translated({60.0, 37.0, 10.0}).into(),
        SceneNode::from(Geometry(KDMesh(*castle_tapestry_model, Shading::Smooth), mat_tapestry)).translated({-60.0, 37.0, 10.0}).into(),
    });
}

SceneNode lake() {
    auto mat_water = Arc(Material{.diffuse = {0.0, 0.0, 0.1}, .specular = {0.5, 0.5, 0.5}, .shininess = 100.0, .reflectivity = 0.9,
                                  .glossy_side_length = 0.5, .refraction_index = WATER_REFRACTION_INDEX});
    auto dock = ImageTexture::open("assets/Wood_018_basecolor_cubemap.jpg");
    auto dock_normals = NormalMap::open("assets/Wood_018_normal_cubemap.jpg");
    auto mat_dock = Arc(Material{.specular = {0.5, 0.5, 0.5}, .shininess = 100.0, .texture = dock, .normals = dock_normals});
    // Color of algae makes the water blue!
    auto mat_dirt = Arc(Material{.diffuse = {0.592, 0.671, 0.055}});
    auto castle_water_dirt_model = obj("castle_water_dirt.obj");

    return SceneNode::from(std::vector<NodeRef>{
        SceneNode::from(Geometry(KDMesh(*castle_water_dirt_model, Shading::Flat), mat_dirt)).translated({0.0, -62.0, 125.0}).into(),
        SceneNode::from(Geometry(Cube{}, mat_water)).scaled({640.0, 125.0, 250.0}).translated({0.0, -62.0, 125.0}).into(),
        // Dock
        SceneNode::from(Geometry(Cube{}, mat_dock)).scaled({30.0, 4.0, 36.0}).translated({-100.0, 0.0, 18.0}).into(),
    });
}

SceneNode land() {
    auto mat_grass = Arc(Material{.diffuse = {0.116971, 0.278894, 0.0}});
    auto castle_hill_model = obj("castle_hill.obj");
    return SceneNode::from(std::vector<NodeRef>{
        SceneNode::from(Geometry(KDMesh(*castle_hill_model, Shading::Smooth), mat_grass))
            .translated({0.0, 3.75, -15.75}).scaled(1.4).translated({0.0, 0.0, -229.0}).into(),
        SceneNode::from(Geometry(Cube{}, mat_grass)).scaled({2560.0, 132.0, 1040.0}).translated({0.0, -65.0, -520.0}).into(),
    });
}

SceneNode outdoor_maze() {
    const double cell_width = 12.0, cell_length = cell_width;
    const double maze_width = 1572.0, maze_length = 1284.0, maze_height = 8.0;
    const Vec3 maze_pos{-450.0, maze_height / 2.0 + 1.0, -660.0};
    // area around the castle
    const double castle_area_width = 276.0, castle_area_length = 264.0;
    const Vec3 castle_pos{0.0 - maze_pos.x, 0.0, -260.0 - maze_pos.z};
    const double entrance_x = -100.0 - maze_pos.x;

    const size_t maze_cols = (size_t)(maze_width / cell_width);
    const size_t maze_rows = (size_t)(maze_length / cell_length);
    const size_t entrance_row = maze_rows - 1;
    const size_t entrance_col = (size_t)((entrance_x + maze_width / 2.0) / cell_width);

    const size_t back_corner_row = (size_t)((castle_pos.z - castle_area_length / 2.0 + maze_length / 2.0) / cell_length);
    const size_t back_corner_col = (size_t)((castle_pos.x - castle_area_width / 2.0 + maze_width / 2.0) / cell_width);
    const size_t front_corner_row = (size_t)((castle_pos.z + castle_area_length / 2.0 + maze_length / 2.0) / cell_length);
    const size_t front_corner_col = (size_t)((castle_pos.x + castle_area_width / 2.0 + maze_width / 2.0) / cell_width);

    Maze maze(maze_rows, maze_cols);
    maze.reserve({back_corner_row, back_corner_col}, {front_corner_row, front_corner_col});
    maze.fill_maze({entrance_row, entrance_col});

    auto shrub = ImageTexture::open("assets/shrub.png");
    auto mat_maze = Arc(Material{.texture = shrub, .uv_trans = Mat3::scaling_3d({1.0, maze_height, 1.0})});

    std::vector<NodeRef> nodes;
    for (size_t i = 0; i < maze.cells.size(); ++i) {
        const double z = (double)i * cell_length - maze_length / 2.0;
        for (size_t j = 0; j < maze.cells[i].size(); ++j) {
            if (maze.cells[i][j] == Cell::Empty) continue;
            const double x = (double)j * cell_width - maze_width / 2.0;
            nodes.push_back(SceneNode::from(Geometry(Cube{}, mat_maze)).scaled({cell_width, maze_height, cell_length}).translated({x, 0.0, z}).into());
        }
    }
    // Translate the maze to its correct position in the scene
    return SceneNode::from(std::move(nodes)).translated(maze_pos);
}

}  // namespace

PORTRAYER_EXAMPLE(graphics_castle, "graphics-castle") {
    ExampleScene ex;
    ex.name = "graphics-castle";
    ex.scene = HierScene{
        .root = SceneNode::from(std::vector<NodeRef>{
            castle().scaled(1.4).translated({0.0, 0.0, -229.0}).into(),
            lake().into(),
            land().into(),
            outdoor_maze().into(),
        }).into(),
        .lights = {Light{.position = {65.0, 130.0, -120.0}, .color = {0.9, 0.9, 0.9}}},
        .ambient = {0.3, 0.3, 0.3},
    };
    ex.cam = CameraSettings{.eye = {110.877441, 30.43659, 373.276886}, .center = {-412.953094, 65.409714, -1390.236328},
                            .up = Vec3::up(), .fovy = Radians::from_degrees(24.0)};
    ex.width = 1920;
    ex.height = 1080;
    ex.background = [](Uv uv) { return Rgb{0.529, 0.808, 0.922} * (1.0 - uv.v) + Rgb{0.086, 0.38, 0.745} * uv.v; };
    return ex;
}
