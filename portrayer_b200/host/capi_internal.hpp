// Shared by the two halves of the C ABI of the host mirror: capi.cpp (scene programs, flatten, k-d build, packing —
// no GPU library behind it) and capi_render.cpp (Image::render, which calls libportrayer_gpu.so).
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "examples/examples.hpp"
#include "pack.hpp"

struct PthScene {
    portrayer::ExampleScene example;
    std::vector<uint8_t> blob;
    double prepare_seconds = 0.0;
    double flatten_seconds = 0.0;     // FlatScene::from alone (matrix products, inverses)
    std::vector<double> item_bounds;  // n x 6: FlatSceneNode::bounds of every flat instance
    std::unique_ptr<portrayer::HierarchyExport> hierarchy;  // built on demand
};


namespace portrayer {
std::string& capi_error();  // thread-local message behind pth_last_error()
}
