// texture_mips.h — host-side mip chain of a material texture, the way the reference produces it: create_texture_from_rgba8 /
// _rgba32f with generate_mipmaps (src/merian/vk/memory/resource_allocator.cpp:298-358) allocates floor(log2(max(w, h))) + 1 levels and
// fills level i with a LINEAR vkCmdBlitImage of level i - 1 into max(1, dim >> i) texels (include/merian/vk/utils/blits.hpp:101-168).
// sRGB formats are filtered in linear space and re-encoded (Vulkan: blits convert through the format), alpha stays linear.
#pragma once

#include <cstdint>
#include <vector>

namespace mpt {

struct MipLevel {
    uint32_t width = 0, height = 0;
    std::vector<uint8_t> bytes; // RGBA8 (4 B/texel) or RGBA32F (16 B/texel), tightly packed rows
};

uint32_t mip_level_count(uint32_t width, uint32_t height);
// levels[0] is a copy of the input; levels[i] the blit of levels[i - 1]
void build_mip_chain(const void* level0, uint32_t width, uint32_t height, bool rgba32f, bool srgb, std::vector<MipLevel>& levels);

} // namespace mpt
