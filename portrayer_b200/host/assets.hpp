// Asset lookup for the host mirror: OBJ files are read from the configured
// assets directory; textures come from a registry of DECODED RGB8 buffers
// (filled through the C API by Python/Pillow, or from assets/_decoded/*.ptex),
// so host, oracle and device consume identical texel bytes (SURVEY §2a:
// jpeg-decoder 0.1.15 parity is unpinned, therefore decode once and share).
#pragma once
#include <cstdint>
#include <string>

namespace portrayer {
void set_assets_dir(const std::string& dir);
std::string resolve_asset_path(const std::string& path);
void set_baked_mesh_dir(const std::string& dir);
std::string resolve_baked_mesh_path(const std::string& path);
// called with the example's path ("assets/earth.jpg") when a texture is not registered yet;
// the callee decodes it and calls register_texture
typedef void (*TextureLoaderFn)(const char* path);
void set_texture_loader(TextureLoaderFn fn);
void register_texture(const std::string& path, uint32_t width, uint32_t height, const uint8_t* rgb8);
}  // namespace portrayer
