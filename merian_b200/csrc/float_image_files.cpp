// float_image_files.cpp — see float_image_files.h.
#include "float_image_files.h"

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>

namespace mpt {
// Headers are not trusted: a width * height beyond this (2^28 texels = 4 GiB of RGBA32F, 16x the largest env map of the configs) is
// reported as a malformed file instead of being handed to std::vector.
constexpr uint64_t kMaxImagePixels = uint64_t(1) << 28;

namespace {

struct FileCloser {
    void operator()(FILE* f) const { if (f) fclose(f); }
};
using File = std::unique_ptr<FILE, FileCloser>;

// magic, width, height, scale separated by whitespace, then exactly ONE whitespace byte before the pixels
std::string parse_header(FILE* f, const char* path, PfmInfo& info) {
    char magic[3] = {0, 0, 0};
    float scale = 0.f;
    if (fscanf(f, "%2s", magic) != 1) return std::string("image_io: bad PFM magic in ") + path;
    if (std::strcmp(magic, "PF") == 0) info.channels = 3;
    else if (std::strcmp(magic, "Pf") == 0) info.channels = 1;
    else return std::string("image_io: bad PFM magic in ") + path;
    if (fscanf(f, "%d %d %f", &info.width, &info.height, &scale) != 3 || info.width <= 0 || info.height <= 0)
        return std::string("image_io: malformed PFM header in ") + path;
    if (uint64_t(info.width) * uint64_t(info.height) > kMaxImagePixels) return std::string("image_io: implausible image size in the header of ") + path;
    if (fgetc(f) == EOF) return std::string("image_io: malformed PFM header in ") + path;
    info.little_endian = scale < 0.f;
    return {};
}

} // namespace

std::string pfm_read_info(const char* path, PfmInfo& info) {
    File f(fopen(path, "rb"));
    if (!f) return std::string("image_io: cannot open ") + path;
    return parse_header(f.get(), path, info);
}

std::string pfm_load(const char* path, int desired_channels, PfmInfo& info, std::vector<float>& out) {
    File f(fopen(path, "rb"));
    if (!f) return std::string("image_io: cannot open ") + path;
    std::string err = parse_header(f.get(), path, info);
    if (!err.empty()) return err;
    if (desired_channels < 0 || desired_channels > 4) return "image_io: desired_channels must be 0..4";
    const int sc = info.channels, oc = desired_channels == 0 ? sc : desired_channels;
    const size_t w = size_t(info.width), h = size_t(info.height);
    out.assign(w * h * size_t(oc), 0.f);
    std::vector<float> row(w * size_t(sc));
    for (size_t y_file = 0; y_file < h; y_file++) { // the file is bottom-up
        if (fread(row.data(), sizeof(float), row.size(), f.get()) != row.size()) return std::string("image_io: truncated PFM in ") + path;
        if (!info.little_endian) {
            for (float& v : row) {
                uint32_t b;
                std::memcpy(&b, &v, 4);
                b = (b << 24) | ((b & 0xFF00u) << 8) | ((b >> 8) & 0xFF00u) | (b >> 24);
                std::memcpy(&v, &b, 4);
            }
        }
        float* dst = out.data() + (h - 1 - y_file) * w * size_t(oc);
        if (oc == sc) {
            std::memcpy(dst, row.data(), row.size() * sizeof(float));
            continue;
        }
        for (size_t x = 0; x < w; x++) {
            const float r = row[x * sc], g = sc >= 2 ? row[x * sc + 1] : r, b = sc >= 3 ? row[x * sc + 2] : r;
            const float px[4] = {r, g, b, 1.f};
            for (int c = 0; c < oc; c++) dst[x * oc + c] = px[c];
        }
    }
    return {};
}

std::string pfm_save(const char* path, const float* data, int width, int height, int channels) {
    if (channels != 1 && channels != 3 && channels != 4) return "image_io: PFM supports only 1, 3 or 4 channels";
    if (width <= 0 || height <= 0 || !data) return "image_io: empty image";
    File f(fopen(path, "wb"));
    if (!f) return std::string("image_io: cannot open ") + path + " for write";
    const int oc = channels == 1 ? 1 : 3;
    fprintf(f.get(), "%s\n%d %d\n-1\n", oc == 3 ? "PF" : "Pf", width, height);
    std::vector<float> row(size_t(width) * oc);
    for (int y = height - 1; y >= 0; --y) {
        const float* src = data + size_t(y) * width * channels;
        const float* wr = src;
        if (channels != oc) {
            for (int x = 0; x < width; x++)
                for (int c = 0; c < oc; c++) row[size_t(x) * oc + c] = src[size_t(x) * channels + c];
            wr = row.data();
        }
        if (fwrite(wr, sizeof(float), row.size(), f.get()) != row.size()) return std::string("image_io: short write to ") + path;
    }
    return {};
}

std::string hdr_load(const char* path, int desired_channels, int& width, int& height, std::vector<float>& out) {
    if (desired_channels != 3 && desired_channels != 4) return "image_io: HDR loads as 3 or 4 channels";
    File f(fopen(path, "rb"));
    if (!f) return std::string("image_io: cannot open ") + path;
    auto line = [&](std::string& s) {
        s.clear();
        int c;
        while ((c = fgetc(f.get())) != EOF && c != '\n') s += char(c);
        return c != EOF || !s.empty();
    };
    std::string l;
    if (!line(l) || (l != "#?RADIANCE" && l != "#?RGBE")) return std::string("image_io: not a Radiance HDR file: ") + path;
    bool format_ok = false;
    while (line(l) && !l.empty())
        if (l == "FORMAT=32-bit_rle_rgbe") format_ok = true;
    if (!format_ok) return std::string("image_io: unsupported HDR format in ") + path;
    if (!line(l) || std::sscanf(l.c_str(), "-Y %d +X %d", &height, &width) != 2 || width <= 0 || height <= 0)
        return std::string("image_io: unsupported HDR data layout in ") + path;
    if (uint64_t(width) * uint64_t(height) > kMaxImagePixels) return std::string("image_io: implausible image size in the header of ") + path;
    const size_t w = size_t(width), h = size_t(height);
    out.assign(w * h * size_t(desired_channels), 0.f);
    std::vector<unsigned char> scan(w * 4);
    auto convert = [&](const unsigned char* rgbe, float* dst) {
        if (rgbe[3] != 0) {
            const float f1 = std::ldexp(1.0f, int(rgbe[3]) - (128 + 8));
            dst[0] = rgbe[0] * f1; dst[1] = rgbe[1] * f1; dst[2] = rgbe[2] * f1;
        } else {
            dst[0] = dst[1] = dst[2] = 0.f;
        }
        if (desired_channels == 4) dst[3] = 1.f;
    };
    for (size_t y = 0; y < h; y++) {
        unsigned char head[4];
        if (fread(head, 1, 4, f.get()) != 4) return std::string("image_io: truncated HDR in ") + path;
        const bool rle = w >= 8 && w < 32768 && head[0] == 2 && head[1] == 2 && !(head[2] & 0x80) && size_t((head[2] << 8) | head[3]) == w;
        if (!rle) { // flat scanline: the 4 bytes read are the first pixel
            std::memcpy(scan.data(), head, 4);
            if (w > 1 && fread(scan.data() + 4, 4, w - 1, f.get()) != w - 1) return std::string("image_io: truncated HDR in ") + path;
        } else {
            std::vector<unsigned char> comp(w);
            for (int k = 0; k < 4; k++) {
                size_t i = 0;
                while (i < w) {
                    int count = fgetc(f.get());
                    if (count == EOF) return std::string("image_io: truncated HDR in ") + path;
                    if (count > 128) { // run
                        const int v = fgetc(f.get());
                        count -= 128;
                        if (v == EOF || count == 0 || i + size_t(count) > w) return std::string("image_io: corrupt HDR in ") + path;
                        for (int z = 0; z < count; z++) comp[i++] = (unsigned char)v;
                    } else { // dump
                        if (count == 0 || i + size_t(count) > w) return std::string("image_io: corrupt HDR in ") + path;
                        if (fread(comp.data() + i, 1, size_t(count), f.get()) != size_t(count)) return std::string("image_io: truncated HDR in ") + path;
                        i += size_t(count);
                    }
                }
                for (size_t x = 0; x < w; x++) scan[x * 4 + size_t(k)] = comp[x];
            }
        }
        for (size_t x = 0; x < w; x++) convert(&scan[x * 4], &out[(y * w + x) * size_t(desired_channels)]);
    }
    return {};
}

} // namespace mpt
