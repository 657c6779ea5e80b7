// texture_mips.cpp — see texture_mips.h.
#include "texture_mips.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace mpt {
namespace {

float srgb_to_linear(float c) { return c <= 0.04045f ? c / 12.92f : std::pow((c + 0.055f) / 1.055f, 2.4f); }
float linear_to_srgb(float l) { return l <= 0.0031308f ? 12.92f * l : 1.055f * std::pow(l, 1.0f / 2.4f) - 0.055f; }

// texel (x, y) of a level as 4 floats in the space the blit filters in
void fetch(const MipLevel& L, bool f32, bool srgb, const float* lut, int x, int y, float out[4]) {
    x = std::min(std::max(x, 0), int(L.width) - 1);
    y = std::min(std::max(y, 0), int(L.height) - 1);
    if (f32) {
        std::memcpy(out, L.bytes.data() + (size_t(y) * L.width + size_t(x)) * 16, 16);
        return;
    }
    const uint8_t* p = L.bytes.data() + (size_t(y) * L.width + size_t(x)) * 4;
    for (int c = 0; c < 3; c++) out[c] = srgb ? lut[p[c]] : float(p[c]) / 255.0f;
    out[3] = float(p[3]) / 255.0f;
}

} // namespace

uint32_t mip_level_count(uint32_t width, uint32_t height) {
    uint32_t n = 1;
    for (uint32_t m = std::max(width, height); m > 1; m >>= 1) n++;
    return n; // floor(log2(max(w, h))) + 1
}

void build_mip_chain(const void* level0, uint32_t width, uint32_t height, bool f32, bool srgb, std::vector<MipLevel>& levels) {
    const uint32_t n = mip_level_count(width, height);
    levels.assign(n, MipLevel());
    const size_t texel = f32 ? 16 : 4;
    levels[0].width = width;
    levels[0].height = height;
    levels[0].bytes.assign(static_cast<const uint8_t*>(level0), static_cast<const uint8_t*>(level0) + size_t(width) * height * texel);
    float lut[256];
    for (int i = 0; i < 256; i++) lut[i] = srgb_to_linear(float(i) / 255.0f);
    for (uint32_t i = 1; i < n; i++) {
        const MipLevel& S = levels[i - 1];
        MipLevel& D = levels[i];
        D.width = std::max(1u, width >> i);
        D.height = std::max(1u, height >> i);
        D.bytes.resize(size_t(D.width) * D.height * texel);
        const float sx = float(S.width) / float(D.width), sy = float(S.height) / float(D.height);
#pragma omp parallel for schedule(static) if (D.width * D.height > 65536)
        for (int64_t y = 0; y < int64_t(D.height); y++) {
            for (uint32_t x = 0; x < D.width; x++) {
                // VK_FILTER_LINEAR at the source coordinate of the destination texel centre, clamp to edge
                const float u = (float(x) + 0.5f) * sx - 0.5f, v = (float(y) + 0.5f) * sy - 0.5f;
                const float fu = std::floor(u), fv = std::floor(v);
                const float a = u - fu, b = v - fv;
                float t00[4], t10[4], t01[4], t11[4], r[4];
                fetch(S, f32, srgb, lut, int(fu), int(fv), t00);
                fetch(S, f32, srgb, lut, int(fu) + 1, int(fv), t10);
                fetch(S, f32, srgb, lut, int(fu), int(fv) + 1, t01);
                fetch(S, f32, srgb, lut, int(fu) + 1, int(fv) + 1, t11);
                for (int c = 0; c < 4; c++) r[c] = (1 - a) * (1 - b) * t00[c] + a * (1 - b) * t10[c] + (1 - a) * b * t01[c] + a * b * t11[c];
                uint8_t* o = D.bytes.data() + (size_t(y) * D.width + x) * texel;
                if (f32) {
                    std::memcpy(o, r, 16);
                } else {
                    for (int c = 0; c < 4; c++) {
                        const float e = (srgb && c < 3) ? linear_to_srgb(r[c]) : r[c];
                        o[c] = uint8_t(std::min(std::max(e * 255.0f + 0.5f, 0.0f), 255.0f));
                    }
                }
            }
        }
    }
}

} // namespace mpt
