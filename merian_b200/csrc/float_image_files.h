// float_image_files.h — PFM reader / writer (host side, no CUDA): the float wire format at either end of the path
// (env maps in: src/merian-shaders/scene/scene.cpp:780-795 -> image_load_f32; images out: the Image Write node's PFM leg,
// src/merian-graph/nodes/image_write/image_write.cpp:23-30 -> image_save_f32). Format as src/merian/io/float_image_files.cpp:53-168.
#pragma once

#include <string>
#include <vector>

namespace mpt {

struct PfmInfo {
    int width = 0, height = 0, channels = 0; // channels of the FILE: 3 ("PF") or 1 ("Pf")
    bool little_endian = true;
};

// All return an empty string on success, else the error text.
std::string pfm_read_info(const char* path, PfmInfo& info);
// Top-down rows, `desired_channels` (0 = the file's) per pixel: grey is replicated into g and b, alpha = 1.
std::string pfm_load(const char* path, int desired_channels, PfmInfo& info, std::vector<float>& out);
// data: top-down rows of `channels` (1, 3 or 4) floats; RGBA is written as RGB (PFM has no alpha); little endian, scale -1.
std::string pfm_save(const char* path, const float* data, int width, int height, int channels);

// Radiance RGBE (.hdr) as image_load_f32 decodes it through stb_image (src/merian/io/float_image_files.cpp:226-240 -> stbi_loadf): header
// "#?RADIANCE" / "#?RGBE", "FORMAT=32-bit_rle_rgbe", "-Y h +X w", flat or new-style RLE scanlines, top-down; component =
// mantissa * 2^(e - 136), (0,0,0,0) -> 0. `desired_channels` 3 or 4 (alpha = 1).
std::string hdr_load(const char* path, int desired_channels, int& width, int& height, std::vector<float>& out);

} // namespace mpt
