// node.cu — host side of the `Render (Path-traced)` node (include/mpt_node.h) on top of the mpt_* entry points.
//
// Mirrors RenderPT (include/merian-graph/nodes/render_pt/render_pt.hpp:26-87, src/merian-graph/nodes/render_pt/
// render_pt.cpp): same property keys / defaults / load semantics, the GBuffer-sized output, process() = one render of
// `samples per pixel` paths per pixel with path length <= `max path length`. What is a compile-time Slang constant in the
// reference (render_pt.cpp:126-145) is a run-time parameter here, so changing a property never recompiles anything.
#include "../../include/mpt_node.h"

#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <array>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>

namespace {

// ---------------------------------------------------------------------------------------------- tiny JSON (flat objects)
struct JValue {
    enum Kind { Null, Bool, Number, String, Other } kind = Null;
    bool b = false;
    double num = 0.0;
    bool is_integer = false;
    std::string str;
};

struct JParser {
    const char* p;
    std::string err;
    explicit JParser(const char* s) : p(s) {}
    void ws() { while (*p && std::isspace((unsigned char)*p)) p++; }
    bool lit(const char* s) { const size_t n = std::strlen(s); if (std::strncmp(p, s, n) == 0) { p += n; return true; } return false; }
    bool string(std::string& out) {
        if (*p != '"') { err = "expected string"; return false; }
        p++;
        out.clear();
        while (*p && *p != '"') {
            if (*p == '\\') {
                p++;
                switch (*p) {
                case 'n': out += '\n'; break;
                case 't': out += '\t'; break;
                case 'r': out += '\r'; break;
                case 'b': out += '\b'; break;
                case 'f': out += '\f'; break;
                case 'u': { // \uXXXX: keep ASCII, replace the rest (property keys are ASCII)
                    unsigned v = 0;
                    for (int i = 1; i <= 4 && p[i]; i++) v = v * 16 + (std::isdigit((unsigned char)p[i]) ? p[i] - '0' : (std::tolower(p[i]) - 'a' + 10));
                    out += v < 128 ? char(v) : '?';
                    p += 4;
                    break;
                }
                case 0: err = "unterminated escape"; return false;
                default: out += *p;
                }
                p++;
            } else {
                out += *p++;
            }
        }
        if (*p != '"') { err = "unterminated string"; return false; }
        p++;
        return true;
    }
    bool skip_compound(char open, char close) { // nested arrays / objects are not properties of this node: skip them
        int depth = 0;
        do {
            if (*p == '"') { std::string s; if (!string(s)) return false; continue; }
            if (*p == open) depth++;
            else if (*p == close) depth--;
            else if (*p == 0) { err = "unterminated value"; return false; }
            p++;
        } while (depth > 0);
        return true;
    }
    bool value(JValue& v) {
        ws();
        if (*p == '"') { v.kind = JValue::String; return string(v.str); }
        if (lit("true")) { v.kind = JValue::Bool; v.b = true; return true; }
        if (lit("false")) { v.kind = JValue::Bool; v.b = false; return true; }
        if (lit("null")) { v.kind = JValue::Null; return true; }
        if (*p == '{') { v.kind = JValue::Other; return skip_compound('{', '}'); }
        if (*p == '[') { v.kind = JValue::Other; return skip_compound('[', ']'); }
        char* end = nullptr;
        const double d = std::strtod(p, &end);
        if (end == p) { err = "unexpected character in value"; return false; }
        v.kind = JValue::Number;
        v.num = d;
        v.is_integer = true;
        for (const char* q = p; q < end; q++)
            if (*q == '.' || *q == 'e' || *q == 'E') v.is_integer = false;
        p = end;
        return true;
    }
    bool object(std::map<std::string, JValue>& out) {
        ws();
        if (*p != '{') { err = "expected a JSON object"; return false; }
        p++;
        ws();
        if (*p == '}') { p++; return true; }
        while (true) {
            ws();
            std::string key;
            if (!string(key)) return false;
            ws();
            if (*p != ':') { err = "expected ':'"; return false; }
            p++;
            JValue v;
            if (!value(v)) return false;
            out[key] = v;
            ws();
            if (*p == ',') { p++; continue; }
            if (*p == '}') { p++; return true; }
            err = "expected ',' or '}'";
            return false;
        }
    }
};

__global__ void k_rgba32f_to_rgba16f(const float4* __restrict__ src, uint2* __restrict__ dst, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 v = src[i];
    const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    dst[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
}

const char* format_name(uint32_t f) { return f == MPTN_FORMAT_R16G16B16A16_SFLOAT ? "R16G16B16A16Sfloat" : "R32G32B32A32Sfloat"; }

} // namespace

struct mptn_node {
    mpt_context* ctx = nullptr;
    std::string error;
    // properties, defaults of render_pt.hpp:53-67
    int32_t spp = 1;
    int32_t max_path_length = 5;
    int32_t emitted_max_path_length = 5;
    bool emission_on_primary = true;
    bool enable_ser = false;
    bool demodulate_albedo = false;
    std::array<bool, 8> mask_enabled{{true, true, true, true, true, true, true, true}};
    uint32_t irradiance_format = MPTN_FORMAT_R32G32B32A32_SFLOAT;
    // describe_outputs
    uint32_t width = 1920, height = 1080; // render_pt.hpp:62
    // sharding
    uint32_t rank = 0, world = 1;
    // scratch for 16-bit output
    void* d_scratch = nullptr;
    size_t scratch_bytes = 0;
};

namespace {
mpt_status nfail(mptn_node* n, mpt_status s, const std::string& msg) {
    if (n) n->error = msg;
    return s;
}
} // namespace

extern "C" {

MPT_API mpt_status mptn_create(mpt_context* ctx, mptn_node** out_node) {
    if (!out_node) return MPT_ERR_INVALID_ARGUMENT; // ctx may be NULL: a node that only loads / dumps its configuration
    *out_node = new mptn_node();
    (*out_node)->ctx = ctx;
    return MPT_OK;
}

MPT_API void mptn_destroy(mptn_node* node) {
    if (!node) return;
    if (node->d_scratch) cudaFree(node->d_scratch);
    delete node;
}

MPT_API const char* mptn_last_error(const mptn_node* node) { return node ? node->error.c_str() : "NULL node"; }

MPT_API mpt_status mptn_load_config(mptn_node* node, const char* json, uint32_t* out_status) {
    if (!node || !json) return nfail(node, MPT_ERR_INVALID_ARGUMENT, "NULL argument");
    if (out_status) *out_status = 0;
    std::map<std::string, JValue> o;
    JParser ps(json);
    if (!ps.object(o)) return nfail(node, MPT_ERR_INVALID_ARGUMENT, "properties JSON: " + ps.err);
    // validate every known key first so that a bad value leaves the node unchanged
    auto find = [&](const char* key) -> const JValue* { auto it = o.find(key); return it == o.end() ? nullptr : &it->second; };
    auto want_int = [&](const char* key, int32_t& dst, std::string& e) {
        if (const JValue* v = find(key)) {
            if (v->kind != JValue::Number) { e = std::string("property '") + key + "' must be a number"; return; }
            dst = int32_t(v->num); // nlohmann::json::get<int32_t> truncates a float
        }
    };
    auto want_bool = [&](const char* key, bool& dst, std::string& e) {
        if (const JValue* v = find(key)) {
            if (v->kind != JValue::Bool) { e = std::string("property '") + key + "' must be a boolean"; return; }
            dst = v->b;
        }
    };
    mptn_node t = *node;
    std::string e;
    want_int("samples per pixel", t.spp, e);
    want_int("max path length", t.max_path_length, e);
    want_bool("emission on primary", t.emission_on_primary, e);
    want_bool("shader execution reordering", t.enable_ser, e);
    want_bool("demodulate albedo", t.demodulate_albedo, e);
    for (int bit = 0; bit < 8; bit++) {
        const std::string k = std::to_string(bit);
        want_bool(k.c_str(), t.mask_enabled[size_t(bit)], e);
    }
    if (const JValue* v = find("irradiance format")) {
        if (v->kind != JValue::String) e = "property 'irradiance format' must be a string";
        else if (v->str == "R32G32B32A32Sfloat") t.irradiance_format = MPTN_FORMAT_R32G32B32A32_SFLOAT;
        else if (v->str == "R16G16B16A16Sfloat") t.irradiance_format = MPTN_FORMAT_R16G16B16A16_SFLOAT;
        else if (!v->str.empty()) {
            // config_options keeps the old selection for an unknown option (properties_json_load.cpp:164-178); formats the
            // CUDA node cannot write are reported instead of silently producing garbage
            return nfail(node, MPT_ERR_UNSUPPORTED, "irradiance format '" + v->str + "' is not supported (R32G32B32A32Sfloat, R16G16B16A16Sfloat)");
        }
    }
    if (!e.empty()) return nfail(node, MPT_ERR_INVALID_ARGUMENT, e);
    const bool reconnect = t.irradiance_format != node->irradiance_format;
    t.d_scratch = node->d_scratch;
    t.scratch_bytes = node->scratch_bytes;
    t.error.clear();
    *node = t;
    if (out_status && reconnect) *out_status |= MPTN_NEEDS_RECONNECT;
    return MPT_OK;
}

MPT_API mpt_status mptn_dump_config(mptn_node* node, char* buf, size_t capacity, size_t* out_len) {
    if (!node) return MPT_ERR_INVALID_ARGUMENT;
    std::string s = "{";
    for (int bit = 0; bit < 8; bit++) s += "\"" + std::to_string(bit) + "\": " + (node->mask_enabled[size_t(bit)] ? "true" : "false") + ", ";
    s += std::string("\"demodulate albedo\": ") + (node->demodulate_albedo ? "true" : "false") + ", ";
    s += std::string("\"emission on primary\": ") + (node->emission_on_primary ? "true" : "false") + ", ";
    s += std::string("\"irradiance format\": \"") + format_name(node->irradiance_format) + "\", ";
    s += "\"max path length\": " + std::to_string(node->max_path_length) + ", ";
    s += "\"samples per pixel\": " + std::to_string(node->spp) + ", ";
    s += std::string("\"shader execution reordering\": ") + (node->enable_ser ? "true" : "false") + "}";
    if (out_len) *out_len = s.size();
    if (buf && capacity) {
        const size_t n = std::min(capacity - 1, s.size());
        std::memcpy(buf, s.data(), n);
        buf[n] = 0;
        if (n < s.size()) return nfail(node, MPT_ERR_LIMIT, "buffer too small for the configuration");
    }
    return MPT_OK;
}

MPT_API mpt_status mptn_describe_outputs(mptn_node* node, uint32_t gw, uint32_t gh, uint32_t* out_vk_format, uint32_t* out_w, uint32_t* out_h) {
    if (!node || !gw || !gh) return nfail(node, MPT_ERR_INVALID_ARGUMENT, "describe_outputs: the gbuffer extent must be non-zero");
    node->width = gw;
    node->height = gh;
    if (out_vk_format) *out_vk_format = node->irradiance_format;
    if (out_w) *out_w = gw;
    if (out_h) *out_h = gh;
    return MPT_OK;
}

MPT_API mpt_status mptn_set_shard(mptn_node* node, uint32_t rank, uint32_t world) {
    if (!node || world == 0 || rank >= world) return nfail(node, MPT_ERR_INVALID_ARGUMENT, "set_shard: need rank < world");
    node->rank = rank;
    node->world = world;
    return MPT_OK;
}

MPT_API mpt_status mptn_process(mptn_node* node, const mpt_camera* camera, uint32_t frame, const void* d_tex1, const void* d_albedo,
                                void* d_irradiance, void* stream, uint32_t* out_events) {
    if (!node || !camera || !d_irradiance) return nfail(node, MPT_ERR_INVALID_ARGUMENT, "NULL argument");
    if (!node->ctx) return nfail(node, MPT_ERR_NOT_READY, "the node was created without a device context (initialize)");
    if (out_events) *out_events = 0;
    if (node->max_path_length != node->emitted_max_path_length) { // render_pt.cpp:102-105
        node->emitted_max_path_length = node->max_path_length;
        if (out_events) *out_events |= MPTN_EVENT_BOUNCES_CHANGED;
    }
    if (node->spp < 0 || node->max_path_length < 1) return nfail(node, MPT_ERR_INVALID_ARGUMENT, "samples per pixel must be >= 0 and max path length >= 1");
    if (node->demodulate_albedo && !d_albedo) return nfail(node, MPT_ERR_INVALID_ARGUMENT, "'demodulate albedo' needs the gbuffer albedo texture");
    mpt_render_params P{};
    P.width = node->width;
    P.height = node->height;
    P.frame = frame;
    P.max_path_length = node->max_path_length;
    uint32_t mask = 0; // render_pt.cpp:127-131
    for (uint32_t bit = 0; bit < 8; bit++)
        if (node->mask_enabled[bit]) mask |= 1u << bit;
    P.instance_mask = mask;
    P.flags = (node->emission_on_primary ? MPT_RENDER_EMISSION_ON_PRIMARY : 0u) | (node->demodulate_albedo ? MPT_RENDER_DEMODULATE_ALBEDO : 0u);
    P.sample_offset = 0;
    P.sample_stride = 1;
    P.spp = node->spp;
    if (node->world > 1) { // interleaved sample shard, sums out
        const int32_t total = node->spp;
        P.spp = (total > int32_t(node->rank)) ? (total - int32_t(node->rank) + int32_t(node->world) - 1) / int32_t(node->world) : 0;
        P.sample_offset = node->rank;
        P.sample_stride = node->world;
        P.flags |= MPT_RENDER_OUTPUT_SUM;
    }
    void* target = d_irradiance;
    const size_t npix = size_t(node->width) * node->height;
    if (node->irradiance_format == MPTN_FORMAT_R16G16B16A16_SFLOAT) {
        if (node->scratch_bytes < npix * 16) {
            if (node->d_scratch) cudaFree(node->d_scratch);
            node->d_scratch = nullptr;
            node->scratch_bytes = 0;
            if (cudaMalloc(&node->d_scratch, npix * 16) != cudaSuccess) return nfail(node, MPT_ERR_CUDA, "cudaMalloc of the fp32 staging image failed");
            node->scratch_bytes = npix * 16;
        }
        target = node->d_scratch;
    }
    const mpt_status s = mpt_render(node->ctx, camera, &P, d_tex1, node->demodulate_albedo ? d_albedo : nullptr, target, stream);
    if (s == MPT_ERR_NOT_READY) return MPT_OK; // `if (!scene || !scene->is_ready()) return {};` render_pt.cpp:98-99
    if (s != MPT_OK) return nfail(node, s, mpt_last_error(node->ctx));
    if (target != d_irradiance) {
        k_rgba32f_to_rgba16f<<<unsigned((npix + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const float4*)target, (uint2*)d_irradiance, uint32_t(npix));
        if (cudaGetLastError() != cudaSuccess) return nfail(node, MPT_ERR_CUDA, "rgba16f pack kernel launch failed");
    }
    return MPT_OK;
}

} // extern "C"
