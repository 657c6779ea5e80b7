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
    bool nee = false; // "next event estimation": CUDA-node extension (MPT_RENDER_NEE), off = the reference estimator
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

MPT_API mpt_status mptn_create(mpt_context* ctx, mptn_node** out_node) try {
    if (!out_node) return MPT_ERR_INVALID_ARGUMENT; // ctx may be NULL: a node that only loads / dumps its configuration
    *out_node = new mptn_node();
    (*out_node)->ctx = ctx;
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return MPT_ERR_LIMIT;
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return MPT_ERR_INVALID_ARGUMENT;
} catch (...) {
    return MPT_ERR_INVALID_ARGUMENT;
}

MPT_API void mptn_destroy(mptn_node* node) {
    if (!node) return;
    if (node->d_scratch) cudaFree(node->d_scratch);
    delete node;
}

MPT_API const char* mptn_last_error(const mptn_node* node) { return node ? node->error.c_str() : "NULL node"; }

MPT_API mpt_status mptn_load_config(mptn_node* node, const char* json, uint32_t* out_status) try {
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
            // nlohmann::json::get<int32_t> truncates a float; out-of-range / NaN values would be undefined behaviour in the cast
            if (!(v->num >= -2147483648.0 && v->num <= 2147483647.0)) { e = std::string("property '") + key + "' is out of the int32 range"; return; }
            dst = int32_t(v->num);
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
    want_bool("next event estimation", t.nee, e);
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
    // UI min / max are not applied on load (properties_json_load.cpp:126-133), but values no frame can be rendered with are refused here
    // instead of failing later in process()
    if (t.spp < 0) return nfail(node, MPT_ERR_INVALID_ARGUMENT, "property 'samples per pixel' must not be negative");
    if (t.max_path_length < 1) return nfail(node, MPT_ERR_INVALID_ARGUMENT, "property 'max path length' must be at least 1");
    const bool reconnect = t.irradiance_format != node->irradiance_format;
    t.d_scratch = node->d_scratch;
    t.scratch_bytes = node->scratch_bytes;
    t.error.clear();
    *node = t;
    if (out_status && reconnect) *out_status |= MPTN_NEEDS_RECONNECT;
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return nfail(node, MPT_ERR_LIMIT, "mptn_load_config: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return nfail(node, MPT_ERR_INVALID_ARGUMENT, std::string("mptn_load_config: ") + e.what());
} catch (...) {
    return nfail(node, MPT_ERR_INVALID_ARGUMENT, "mptn_load_config: unknown exception");
}

MPT_API mpt_status mptn_dump_config(mptn_node* node, char* buf, size_t capacity, size_t* out_len) try {
    if (!node) return MPT_ERR_INVALID_ARGUMENT;
    std::string s = "{";
    for (int bit = 0; bit < 8; bit++) s += "\"" + std::to_string(bit) + "\": " + (node->mask_enabled[size_t(bit)] ? "true" : "false") + ", ";
    s += std::string("\"demodulate albedo\": ") + (node->demodulate_albedo ? "true" : "false") + ", ";
    s += std::string("\"emission on primary\": ") + (node->emission_on_primary ? "true" : "false") + ", ";
    s += std::string("\"irradiance format\": \"") + format_name(node->irradiance_format) + "\", ";
    s += "\"max path length\": " + std::to_string(node->max_path_length) + ", ";
    s += std::string("\"next event estimation\": ") + (node->nee ? "true" : "false") + ", ";
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
} catch (const std::bad_alloc&) {
    return nfail(node, MPT_ERR_LIMIT, "mptn_dump_config: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return nfail(node, MPT_ERR_INVALID_ARGUMENT, std::string("mptn_dump_config: ") + e.what());
} catch (...) {
    return nfail(node, MPT_ERR_INVALID_ARGUMENT, "mptn_dump_config: unknown exception");
}

MPT_API mpt_status mptn_describe_outputs(mptn_node* node, uint32_t gw, uint32_t gh, uint32_t* out_vk_format, uint32_t* out_w, uint32_t* out_h) try {
    if (!node || !gw || !gh) return nfail(node, MPT_ERR_INVALID_ARGUMENT, "describe_outputs: the gbuffer extent must be non-zero");
    node->width = gw;
    node->height = gh;
    if (out_vk_format) *out_vk_format = node->irradiance_format;
    if (out_w) *out_w = gw;
    if (out_h) *out_h = gh;
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return nfail(node, MPT_ERR_LIMIT, "mptn_describe_outputs: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return nfail(node, MPT_ERR_INVALID_ARGUMENT, std::string("mptn_describe_outputs: ") + e.what());
} catch (...) {
    return nfail(node, MPT_ERR_INVALID_ARGUMENT, "mptn_describe_outputs: unknown exception");
}

MPT_API mpt_status mptn_set_shard(mptn_node* node, uint32_t rank, uint32_t world) try {
    if (!node || world == 0 || rank >= world) return nfail(node, MPT_ERR_INVALID_ARGUMENT, "set_shard: need rank < world");
    node->rank = rank;
    node->world = world;
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return nfail(node, MPT_ERR_LIMIT, "mptn_set_shard: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return nfail(node, MPT_ERR_INVALID_ARGUMENT, std::string("mptn_set_shard: ") + e.what());
} catch (...) {
    return nfail(node, MPT_ERR_INVALID_ARGUMENT, "mptn_set_shard: unknown exception");
}

MPT_API mpt_status mptn_process(mptn_node* node, const mpt_camera* camera, uint32_t frame, const void* d_tex1, const void* d_albedo,
                                void* d_irradiance, void* stream, uint32_t* out_events) try {
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
    P.flags = (node->emission_on_primary ? MPT_RENDER_EMISSION_ON_PRIMARY : 0u) | (node->demodulate_albedo ? MPT_RENDER_DEMODULATE_ALBEDO : 0u) |
              (node->nee ? MPT_RENDER_NEE : 0u);
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
    // a shard writes SUMS over its samples (and of luminance^2 in .w): they leave the fp16 range (65504) long before the reduce has
    // normalised them, so partial sums stay fp32 whatever the configured output format is; the host converts after the reduce
    if (node->world > 1 && node->irradiance_format == MPTN_FORMAT_R16G16B16A16_SFLOAT)
        return nfail(node, MPT_ERR_UNSUPPORTED, "a sample shard (mptn_set_shard, world > 1) writes fp32 partial sums: set \"irradiance format\" to R32G32B32A32Sfloat and convert after the reduce");
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
} catch (const std::bad_alloc&) {
    return nfail(node, MPT_ERR_LIMIT, "mptn_process: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return nfail(node, MPT_ERR_INVALID_ARGUMENT, std::string("mptn_process: ") + e.what());
} catch (...) {
    return nfail(node, MPT_ERR_INVALID_ARGUMENT, "mptn_process: unknown exception");
}

} // extern "C"


// ====================================================================================== GBuffer (Raytraced) node, host side
// GBufferRTNode (include/merian-graph/nodes/gbuffer_rt/gbuffer.hpp:20-75, src/merian-graph/nodes/gbuffer_rt/gbuffer.cpp): properties
// "width" / "height" (when the scene camera dictates no resolution), "0".."7", "emission format"; outputs "gbuffer" + "emission".
struct mptn_gbuffer {
    mpt_context* ctx = nullptr;
    std::string error;
    uint32_t width = 1920, height = 1080; // gbuffer.hpp:64
    std::array<bool, 8> mask_enabled{{true, true, true, true, true, true, true, true}};
    uint32_t emission_format = MPTN_FORMAT_R32G32B32A32_SFLOAT; // gbuffer.hpp:70
};

// ====================================================================================== Accumulate node (plain mode), host side
// Accumulate (include/merian-graph/nodes/accumulate/accumulate.hpp, src/merian-graph/nodes/accumulate/accumulate.cpp:153-250).
struct mptn_accumulate {
    mpt_context* ctx = nullptr;
    std::string error;
    float alpha = 0.0f;            // accumulate.hpp:34
    float max_history = INFINITY;  // :35
    float normal_reject_cos = 0.8f; // :36
    float depth_reject_percent = 0.02f;
    float hard_clamp = INFINITY;   // :32
    bool clear = false;
    bool use_motion_vectors = true; // reference default; the CUDA node supports false only (same-pixel reuse)
    int gbuffer_check_mode = 0;     // "always" | "only if moving" | "never"
    bool firefly_filter_enable = false;
    float adaptivity = 0.0f;
    bool extended_search = true;    // accepted; the same-pixel check needs no neighbourhood search
    bool history_f32 = false;       // CUDA-node extension: "f32 history" (no 2048 cap)
};

namespace {
template <class N> mpt_status nfail2(N* n, mpt_status s, const std::string& msg) {
    if (n) n->error = msg;
    return s;
}
struct PropReader {
    std::map<std::string, JValue> o;
    std::string err;
    const JValue* find(const char* key) const { auto it = o.find(key); return it == o.end() ? nullptr : &it->second; }
    void num(const char* key, float& dst) {
        if (const JValue* v = find(key)) {
            if (v->kind != JValue::Number) err = std::string("property '") + key + "' must be a number";
            else dst = float(v->num);
        }
    }
    void uint(const char* key, uint32_t& dst) {
        if (const JValue* v = find(key)) {
            if (v->kind != JValue::Number || v->num < 0) err = std::string("property '") + key + "' must be a non-negative number";
            else dst = uint32_t(v->num);
        }
    }
    void boolean(const char* key, bool& dst) {
        if (const JValue* v = find(key)) {
            if (v->kind != JValue::Bool) err = std::string("property '") + key + "' must be a boolean";
            else dst = v->b;
        }
    }
};
std::string jnum(float f) {
    if (std::isinf(f)) return f > 0 ? "1e999" : "-1e999"; // what nlohmann::json cannot represent either; round-trips through strtod as inf
    char b[64];
    std::snprintf(b, sizeof b, "%.9g", double(f));
    return b;
}
mpt_status copy_out(const std::string& s, char* buf, size_t capacity, size_t* out_len) {
    if (out_len) *out_len = s.size();
    if (buf && capacity) {
        const size_t n = std::min(capacity - 1, s.size());
        std::memcpy(buf, s.data(), n);
        buf[n] = 0;
        if (n < s.size()) return MPT_ERR_LIMIT;
    }
    return MPT_OK;
}
} // namespace

extern "C" {

MPT_API mpt_status mptn_gbuffer_create(mpt_context* ctx, mptn_gbuffer** out_node) try {
    if (!out_node) return MPT_ERR_INVALID_ARGUMENT;
    *out_node = new mptn_gbuffer();
    (*out_node)->ctx = ctx;
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return MPT_ERR_LIMIT;
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return MPT_ERR_INVALID_ARGUMENT;
} catch (...) {
    return MPT_ERR_INVALID_ARGUMENT;
}
MPT_API void mptn_gbuffer_destroy(mptn_gbuffer* node) { delete node; }
MPT_API const char* mptn_gbuffer_last_error(const mptn_gbuffer* node) { return node ? node->error.c_str() : "NULL node"; }

MPT_API mpt_status mptn_gbuffer_load_config(mptn_gbuffer* node, const char* json, uint32_t* out_status) try {
    if (!node || !json) return nfail2(node, MPT_ERR_INVALID_ARGUMENT, "NULL argument");
    if (out_status) *out_status = 0;
    PropReader r;
    JParser ps(json);
    if (!ps.object(r.o)) return nfail2(node, MPT_ERR_INVALID_ARGUMENT, "properties JSON: " + ps.err);
    mptn_gbuffer t = *node;
    r.uint("width", t.width);   // gbuffer.cpp:160-161
    r.uint("height", t.height);
    for (int bit = 0; bit < 8; bit++) r.boolean(std::to_string(bit).c_str(), t.mask_enabled[size_t(bit)]);
    if (const JValue* v = r.find("emission format")) {
        if (v->kind != JValue::String) r.err = "property 'emission format' must be a string";
        else if (v->str == "R32G32B32A32Sfloat") t.emission_format = MPTN_FORMAT_R32G32B32A32_SFLOAT;
        else if (!v->str.empty()) return nfail2(node, MPT_ERR_UNSUPPORTED, "emission format '" + v->str + "' is not supported (R32G32B32A32Sfloat)");
    }
    if (!r.err.empty()) return nfail2(node, MPT_ERR_INVALID_ARGUMENT, r.err);
    if (!t.width || !t.height) return nfail2(node, MPT_ERR_INVALID_ARGUMENT, "width and height must be non-zero");
    const bool reconnect = t.width != node->width || t.height != node->height || t.emission_format != node->emission_format;
    t.error.clear();
    *node = t;
    if (out_status && reconnect) *out_status |= MPTN_NEEDS_RECONNECT;
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return nfail2(node, MPT_ERR_LIMIT, "mptn_gbuffer_load_config: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return nfail2(node, MPT_ERR_INVALID_ARGUMENT, std::string("mptn_gbuffer_load_config: ") + e.what());
} catch (...) {
    return nfail2(node, MPT_ERR_INVALID_ARGUMENT, "mptn_gbuffer_load_config: unknown exception");
}

MPT_API mpt_status mptn_gbuffer_dump_config(mptn_gbuffer* node, char* buf, size_t capacity, size_t* out_len) try {
    if (!node) return MPT_ERR_INVALID_ARGUMENT;
    std::string s = "{";
    for (int bit = 0; bit < 8; bit++) s += "\"" + std::to_string(bit) + "\": " + (node->mask_enabled[size_t(bit)] ? "true" : "false") + ", ";
    s += std::string("\"emission format\": \"") + format_name(node->emission_format) + "\", ";
    s += "\"height\": " + std::to_string(node->height) + ", \"width\": " + std::to_string(node->width) + "}";
    const mpt_status st = copy_out(s, buf, capacity, out_len);
    return st == MPT_OK ? st : nfail2(node, st, "buffer too small for the configuration");
} catch (const std::bad_alloc&) {
    return nfail2(node, MPT_ERR_LIMIT, "mptn_gbuffer_dump_config: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return nfail2(node, MPT_ERR_INVALID_ARGUMENT, std::string("mptn_gbuffer_dump_config: ") + e.what());
} catch (...) {
    return nfail2(node, MPT_ERR_INVALID_ARGUMENT, "mptn_gbuffer_dump_config: unknown exception");
}

MPT_API mpt_status mptn_gbuffer_set_resolution(mptn_gbuffer* node, uint32_t width, uint32_t height, uint32_t* out_status) try {
    if (!node || !width || !height) return nfail2(node, MPT_ERR_INVALID_ARGUMENT, "resolution must be non-zero");
    if (out_status) *out_status = (width != node->width || height != node->height) ? MPTN_NEEDS_RECONNECT : 0u; // pre_process, gbuffer.cpp:104-115
    node->width = width;
    node->height = height;
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return nfail2(node, MPT_ERR_LIMIT, "mptn_gbuffer_set_resolution: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return nfail2(node, MPT_ERR_INVALID_ARGUMENT, std::string("mptn_gbuffer_set_resolution: ") + e.what());
} catch (...) {
    return nfail2(node, MPT_ERR_INVALID_ARGUMENT, "mptn_gbuffer_set_resolution: unknown exception");
}

MPT_API mpt_status mptn_gbuffer_describe_outputs(mptn_gbuffer* node, uint32_t* out_width, uint32_t* out_height, uint32_t* out_emission_vk_format) try {
    if (!node) return MPT_ERR_INVALID_ARGUMENT;
    if (out_width) *out_width = node->width;
    if (out_height) *out_height = node->height;
    if (out_emission_vk_format) *out_emission_vk_format = node->emission_format;
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return nfail2(node, MPT_ERR_LIMIT, "mptn_gbuffer_describe_outputs: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return nfail2(node, MPT_ERR_INVALID_ARGUMENT, std::string("mptn_gbuffer_describe_outputs: ") + e.what());
} catch (...) {
    return nfail2(node, MPT_ERR_INVALID_ARGUMENT, "mptn_gbuffer_describe_outputs: unknown exception");
}

MPT_API mpt_status mptn_gbuffer_process(mptn_gbuffer* node, const mpt_camera* camera, const mpt_camera* prev_camera, const mpt_gbuffer_images* images,
                                        void* stream) try {
    if (!node || !camera || !images) return nfail2(node, MPT_ERR_INVALID_ARGUMENT, "NULL argument");
    if (!node->ctx) return nfail2(node, MPT_ERR_NOT_READY, "the node was created without a device context (initialize)");
    uint32_t mask = 0; // gbuffer.cpp:129-134
    for (uint32_t bit = 0; bit < 8; bit++)
        if (node->mask_enabled[bit]) mask |= 1u << bit;
    const mpt_status s = mpt_gbuffer(node->ctx, camera, prev_camera, node->width, node->height, mask, images, stream);
    if (s == MPT_ERR_NOT_READY) return MPT_OK; // `if (scene->is_ready())`, gbuffer.cpp:136
    if (s != MPT_OK) return nfail2(node, s, mpt_last_error(node->ctx));
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return nfail2(node, MPT_ERR_LIMIT, "mptn_gbuffer_process: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return nfail2(node, MPT_ERR_INVALID_ARGUMENT, std::string("mptn_gbuffer_process: ") + e.what());
} catch (...) {
    return nfail2(node, MPT_ERR_INVALID_ARGUMENT, "mptn_gbuffer_process: unknown exception");
}

MPT_API mpt_status mptn_accumulate_create(mpt_context* ctx, mptn_accumulate** out_node) try {
    if (!out_node) return MPT_ERR_INVALID_ARGUMENT;
    *out_node = new mptn_accumulate();
    (*out_node)->ctx = ctx;
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return MPT_ERR_LIMIT;
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return MPT_ERR_INVALID_ARGUMENT;
} catch (...) {
    return MPT_ERR_INVALID_ARGUMENT;
}
MPT_API void mptn_accumulate_destroy(mptn_accumulate* node) { delete node; }
MPT_API const char* mptn_accumulate_last_error(const mptn_accumulate* node) { return node ? node->error.c_str() : "NULL node"; }

MPT_API mpt_status mptn_accumulate_load_config(mptn_accumulate* node, const char* json, uint32_t* out_status) try {
    if (!node || !json) return nfail2(node, MPT_ERR_INVALID_ARGUMENT, "NULL argument");
    if (out_status) *out_status = 0;
    PropReader r;
    JParser ps(json);
    if (!ps.object(r.o)) return nfail2(node, MPT_ERR_INVALID_ARGUMENT, "properties JSON: " + ps.err);
    mptn_accumulate t = *node;
    r.num("alpha", t.alpha);                      // accumulate.cpp:175-177
    r.num("max history", t.max_history);          // :178-180
    bool inf_history = std::isinf(t.max_history), clear = false, inf_clamp = false;
    r.boolean("inf history", inf_history);        // :182-183
    if (inf_history) t.max_history = INFINITY;
    r.boolean("clear", clear);                    // :184
    r.boolean("use motion vectors", t.use_motion_vectors);
    if (const JValue* v = r.find("gbuffer check")) {
        if (v->kind != JValue::String) r.err = "property 'gbuffer check' must be a string";
        else if (v->str == "always") t.gbuffer_check_mode = 0;
        else if (v->str == "only if moving") t.gbuffer_check_mode = 1;
        else if (v->str == "never") t.gbuffer_check_mode = 2;
    }
    float angle = std::acos(t.normal_reject_cos);
    r.num("normal threshold", angle);             // radians, :199-202
    t.normal_reject_cos = std::cos(angle);
    r.num("depth threshold", t.depth_reject_percent);
    r.boolean("extended search", t.extended_search);
    r.boolean("firefly filter enable", t.firefly_filter_enable);
    r.num("hard clamp", t.hard_clamp);            // :229-233
    r.boolean("inf clamp", inf_clamp);
    if (inf_clamp) t.hard_clamp = INFINITY;
    r.num("adaptivity", t.adaptivity);
    r.boolean("f32 history", t.history_f32);
    if (!r.err.empty()) return nfail2(node, MPT_ERR_INVALID_ARGUMENT, r.err);
    if (!(t.alpha >= 0.f && t.alpha <= 1.f)) return nfail2(node, MPT_ERR_INVALID_ARGUMENT, "alpha must be in [0, 1]");
    const bool rebuild = t.use_motion_vectors != node->use_motion_vectors || t.gbuffer_check_mode != node->gbuffer_check_mode || t.extended_search != node->extended_search;
    t.clear = node->clear || clear;
    t.error.clear();
    *node = t;
    if (out_status && rebuild) *out_status |= MPTN_NEEDS_RECONNECT;
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return nfail2(node, MPT_ERR_LIMIT, "mptn_accumulate_load_config: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return nfail2(node, MPT_ERR_INVALID_ARGUMENT, std::string("mptn_accumulate_load_config: ") + e.what());
} catch (...) {
    return nfail2(node, MPT_ERR_INVALID_ARGUMENT, "mptn_accumulate_load_config: unknown exception");
}

MPT_API mpt_status mptn_accumulate_dump_config(mptn_accumulate* node, char* buf, size_t capacity, size_t* out_len) try {
    if (!node) return MPT_ERR_INVALID_ARGUMENT;
    static const char* modes[3] = {"always", "only if moving", "never"};
    auto b = [](bool v) { return v ? "true" : "false"; };
    std::string s = "{";
    s += "\"adaptivity\": " + jnum(node->adaptivity) + ", \"alpha\": " + jnum(node->alpha) + ", \"clear\": false, ";
    s += "\"depth threshold\": " + jnum(node->depth_reject_percent) + ", \"extended search\": " + b(node->extended_search) + ", ";
    s += std::string("\"f32 history\": ") + b(node->history_f32) + ", \"firefly filter enable\": " + b(node->firefly_filter_enable) + ", ";
    s += std::string("\"gbuffer check\": \"") + modes[node->gbuffer_check_mode] + "\", ";
    s += "\"hard clamp\": " + (std::isinf(node->hard_clamp) ? std::string("0") : jnum(node->hard_clamp)) + ", \"inf clamp\": " + b(std::isinf(node->hard_clamp)) + ", ";
    s += std::string("\"inf history\": ") + b(std::isinf(node->max_history)) + ", \"max history\": " + (std::isinf(node->max_history) ? std::string("0") : jnum(node->max_history)) + ", ";
    s += "\"normal threshold\": " + jnum(std::acos(node->normal_reject_cos)) + ", \"use motion vectors\": " + b(node->use_motion_vectors) + "}";
    const mpt_status st = copy_out(s, buf, capacity, out_len);
    return st == MPT_OK ? st : nfail2(node, st, "buffer too small for the configuration");
} catch (const std::bad_alloc&) {
    return nfail2(node, MPT_ERR_LIMIT, "mptn_accumulate_dump_config: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return nfail2(node, MPT_ERR_INVALID_ARGUMENT, std::string("mptn_accumulate_dump_config: ") + e.what());
} catch (...) {
    return nfail2(node, MPT_ERR_INVALID_ARGUMENT, "mptn_accumulate_dump_config: unknown exception");
}

MPT_API mpt_status mptn_accumulate_request_clear(mptn_accumulate* node) try { // Accumulate::request_clear, accumulate.cpp:255-257
    if (!node) return MPT_ERR_INVALID_ARGUMENT;
    node->clear = true;
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return nfail2(node, MPT_ERR_LIMIT, "mptn_accumulate_request_clear: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return nfail2(node, MPT_ERR_INVALID_ARGUMENT, std::string("mptn_accumulate_request_clear: ") + e.what());
} catch (...) {
    return nfail2(node, MPT_ERR_INVALID_ARGUMENT, "mptn_accumulate_request_clear: unknown exception");
}

namespace {
mpt_status accumulate_params(mptn_accumulate* node, uint64_t iteration, uint32_t width, uint32_t height, mpt_accumulate_params& A, uint32_t* out_events) {
    if (node->use_motion_vectors) return nfail2(node, MPT_ERR_UNSUPPORTED, "the CUDA Accumulate node reuses the same pixel only: set 'use motion vectors' to false");
    if (node->firefly_filter_enable || node->adaptivity > 0.f) return nfail2(node, MPT_ERR_UNSUPPORTED, "firefly percentile filter / adaptive alpha stay in the reference node (SURVEY.md 8(f) N2)");
    if (node->gbuffer_check_mode == 0 && node->history_f32) return nfail2(node, MPT_ERR_INVALID_ARGUMENT, "'f32 history' stores no depth: use 'gbuffer check' = never / only if moving");
    A = mpt_accumulate_params{};
    A.width = width; A.height = height;
    A.alpha = node->alpha; A.max_history = node->max_history; A.hard_clamp = node->hard_clamp;
    A.normal_reject_cos = node->normal_reject_cos; A.z_reject_percent = node->depth_reject_percent;
    A.clear = (iteration == 0 || node->clear) ? 1u : 0u;                                          // accumulate.cpp:154-160
    // same pixel => the motion vector is zero => "only if moving" never checks (accumulate.slang:186-187)
    A.flags = (node->gbuffer_check_mode == 0 ? MPT_ACC_GBUFFER_CHECK : 0u) | (node->history_f32 ? MPT_ACC_HISTORY_F32 : 0u);
    if (out_events) *out_events = A.clear ? MPTN_EVENT_CLEAR : 0u;
    node->clear = false;
    return MPT_OK;
}
} // namespace

MPT_API mpt_status mptn_accumulate_process(mptn_accumulate* node, uint64_t iteration, uint32_t width, uint32_t height, const void* d_src,
                                           const mpt_accumulate_buffers* buffers, void* stream, uint32_t* out_events) try {
    if (!node || !d_src || !buffers) return nfail2(node, MPT_ERR_INVALID_ARGUMENT, "NULL argument");
    if (!node->ctx) return nfail2(node, MPT_ERR_NOT_READY, "the node was created without a device context (initialize)");
    mpt_accumulate_params A;
    const mpt_status ps = accumulate_params(node, iteration, width, height, A, out_events);
    if (ps != MPT_OK) return ps;
    const mpt_status s = mpt_accumulate(node->ctx, &A, d_src, buffers, stream);
    if (s != MPT_OK) return nfail2(node, s, mpt_last_error(node->ctx));
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return nfail2(node, MPT_ERR_LIMIT, "mptn_accumulate_process: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return nfail2(node, MPT_ERR_INVALID_ARGUMENT, std::string("mptn_accumulate_process: ") + e.what());
} catch (...) {
    return nfail2(node, MPT_ERR_INVALID_ARGUMENT, "mptn_accumulate_process: unknown exception");
}

MPT_API mpt_status mptn_process_accumulate(mptn_node* pt, mptn_accumulate* acc, const mpt_camera* camera, uint32_t frame, uint64_t iteration,
                                           const void* d_tex1, const void* d_albedo, const mpt_accumulate_buffers* buffers, void* stream,
                                           uint32_t* out_events) try {
    if (!pt || !acc || !camera || !buffers) return nfail2(acc, MPT_ERR_INVALID_ARGUMENT, "NULL argument");
    if (!pt->ctx || pt->ctx != acc->ctx) return nfail2(acc, MPT_ERR_INVALID_ARGUMENT, "both nodes must live on the same device context");
    if (pt->world > 1) return nfail2(acc, MPT_ERR_INVALID_ARGUMENT, "a sharded pt node writes partial sums: reduce first, then mptn_accumulate_process");
    if (pt->irradiance_format != MPTN_FORMAT_R32G32B32A32_SFLOAT) return nfail2(acc, MPT_ERR_UNSUPPORTED, "the fused path keeps the irradiance in fp32 registers: set 'irradiance format' to R32G32B32A32Sfloat");
    uint32_t ev_pt = 0, ev_acc = 0;
    if (pt->max_path_length != pt->emitted_max_path_length) { pt->emitted_max_path_length = pt->max_path_length; ev_pt |= MPTN_EVENT_BOUNCES_CHANGED; }
    if (pt->spp < 0 || pt->max_path_length < 1) return nfail2(acc, MPT_ERR_INVALID_ARGUMENT, "samples per pixel must be >= 0 and max path length >= 1");
    if (pt->demodulate_albedo && !d_albedo) return nfail2(acc, MPT_ERR_INVALID_ARGUMENT, "'demodulate albedo' needs the gbuffer albedo texture");
    mpt_accumulate_params A;
    const mpt_status ps = accumulate_params(acc, iteration, pt->width, pt->height, A, &ev_acc);
    if (ps != MPT_OK) return ps;
    mpt_render_params P{};
    P.width = pt->width; P.height = pt->height; P.frame = frame; P.spp = pt->spp; P.max_path_length = pt->max_path_length;
    for (uint32_t bit = 0; bit < 8; bit++)
        if (pt->mask_enabled[bit]) P.instance_mask |= 1u << bit;
    P.flags = (pt->emission_on_primary ? MPT_RENDER_EMISSION_ON_PRIMARY : 0u) | (pt->demodulate_albedo ? MPT_RENDER_DEMODULATE_ALBEDO : 0u) |
              (pt->nee ? MPT_RENDER_NEE : 0u);
    P.sample_offset = 0; P.sample_stride = 1;
    if (out_events) *out_events = ev_pt | ev_acc;
    const mpt_status s = mpt_render_accumulate(pt->ctx, camera, &P, d_tex1, pt->demodulate_albedo ? d_albedo : nullptr, &A, buffers, stream);
    if (s == MPT_ERR_NOT_READY) return MPT_OK;
    if (s != MPT_OK) return nfail2(acc, s, mpt_last_error(pt->ctx));
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return nfail(pt, MPT_ERR_LIMIT, "mptn_process_accumulate: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return nfail(pt, MPT_ERR_INVALID_ARGUMENT, std::string("mptn_process_accumulate: ") + e.what());
} catch (...) {
    return nfail(pt, MPT_ERR_INVALID_ARGUMENT, "mptn_process_accumulate: unknown exception");
}

} // extern "C"
