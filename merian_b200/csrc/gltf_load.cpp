// gltf_load.cpp — mptt_gltf_* (include/mpt_tables.h): glTF 2.0 (.gltf / .glb) straight into the source tables of mptt_build /
// mptt_materials_build, for hosts without a merian::Scene. Host-only. Restates what GLTFScene::load hands to
// merian::Scene (src/merian-shaders/scene/gltf_scene.cpp):
//   materials incl. the KHR extension gates                         :265-536  (load_materials)
//   texture slots in first-use order, sRGB / linear per use          :538-604  (get_or_load_texture)
//   one Scene::Mesh per TRIANGLES primitive, flags, default material :607-694  (load_meshes)
//   node transforms (matrix or T*R*S), DFS node order                :146-196, 696-716 (gltf_node_transform, load_node)
//   perspective camera from the node's global transform              :719-775  (load_cameras)
// The reference parses with tinygltf and decodes images with stb_image, neither of which is vendored: the container is parsed here (JSON,
// GLB chunks, base64 data URIs, strided accessors) and images are handed to the caller as the encoded bytes found in the file.
// Same rules as the Python mirror merian_b200/gltf.py; tests/test_gltf_cabi.py holds the two against each other. Nothing in a file is
// trusted: every index and byte range is checked, and no exception crosses the C ABI.
#include "../../include/mpt_tables.h"
#include "mpt_internal.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace {

struct GltfError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// ------------------------------------------------------------------------------------------------------------------- JSON
struct Json {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    bool b = false;
    double num = 0.0;
    std::string str;
    std::vector<Json> arr;
    std::vector<std::pair<std::string, Json>> obj;

    const Json* get(const char* key) const {
        if (kind != Obj) return nullptr;
        for (const auto& kv : obj)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
    bool has(const char* key) const { return get(key) != nullptr; }
    double number(const char* key, double def) const {
        const Json* v = get(key);
        return v && v->kind == Num ? v->num : def;
    }
    int64_t integer(const char* key, int64_t def) const {
        const Json* v = get(key);
        return v && v->kind == Num ? int64_t(v->num) : def;
    }
    const std::vector<Json>& array(const char* key) const {
        static const std::vector<Json> empty;
        const Json* v = get(key);
        return v && v->kind == Arr ? v->arr : empty;
    }
    const Json& object(const char* key) const {
        static const Json empty_obj = [] { Json j; j.kind = Obj; return j; }();
        const Json* v = get(key);
        return v && v->kind == Obj ? *v : empty_obj;
    }
};

struct JsonParser {
    const char* p;
    const char* end;
    int depth = 0;
    JsonParser(const char* s, size_t n) : p(s), end(s + n) {}
    [[noreturn]] void fail(const char* what) const { throw GltfError(std::string("JSON: ") + what); }
    void ws() { while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) p++; }
    bool lit(const char* s) {
        const size_t n = std::strlen(s);
        if (size_t(end - p) >= n && std::memcmp(p, s, n) == 0) { p += n; return true; }
        return false;
    }
    static void append_utf8(std::string& out, unsigned cp) {
        if (cp < 0x80) out += char(cp);
        else if (cp < 0x800) { out += char(0xC0 | (cp >> 6)); out += char(0x80 | (cp & 0x3F)); }
        else if (cp < 0x10000) { out += char(0xE0 | (cp >> 12)); out += char(0x80 | ((cp >> 6) & 0x3F)); out += char(0x80 | (cp & 0x3F)); }
        else { out += char(0xF0 | (cp >> 18)); out += char(0x80 | ((cp >> 12) & 0x3F)); out += char(0x80 | ((cp >> 6) & 0x3F)); out += char(0x80 | (cp & 0x3F)); }
    }
    unsigned hex4() {
        if (end - p < 4) fail("short \\u escape");
        unsigned v = 0;
        for (int i = 0; i < 4; i++) {
            const char c = *p++;
            v = v * 16 + (c >= '0' && c <= '9' ? unsigned(c - '0') : c >= 'a' && c <= 'f' ? unsigned(c - 'a' + 10) : c >= 'A' && c <= 'F' ? unsigned(c - 'A' + 10) : (fail("bad \\u escape"), 0u));
        }
        return v;
    }
    std::string string() {
        if (p >= end || *p != '"') fail("expected a string");
        p++;
        std::string out;
        while (p < end && *p != '"') {
            if (*p == '\\') {
                p++;
                if (p >= end) fail("unterminated escape");
                const char c = *p++;
                switch (c) {
                case 'n': out += '\n'; break;
                case 't': out += '\t'; break;
                case 'r': out += '\r'; break;
                case 'b': out += '\b'; break;
                case 'f': out += '\f'; break;
                case 'u': {
                    unsigned cp = hex4();
                    if (cp >= 0xD800 && cp < 0xDC00 && end - p >= 6 && p[0] == '\\' && p[1] == 'u') { // surrogate pair
                        p += 2;
                        const unsigned lo = hex4();
                        cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                    }
                    append_utf8(out, cp);
                    break;
                }
                default: out += c; // \" \\ \/
                }
            } else {
                out += *p++;
            }
        }
        if (p >= end) fail("unterminated string");
        p++;
        return out;
    }
    Json value() {
        if (++depth > 200) fail("nesting too deep");
        ws();
        if (p >= end) fail("unexpected end");
        Json v;
        if (*p == '"') {
            v.kind = Json::Str;
            v.str = string();
        } else if (*p == '{') {
            v.kind = Json::Obj;
            p++;
            ws();
            if (p < end && *p == '}') { p++; }
            else {
                while (true) {
                    ws();
                    std::string key = string();
                    ws();
                    if (p >= end || *p != ':') fail("expected ':'");
                    p++;
                    Json item = value();
                    v.obj.emplace_back(std::move(key), std::move(item));
                    ws();
                    if (p < end && *p == ',') { p++; continue; }
                    if (p < end && *p == '}') { p++; break; }
                    fail("expected ',' or '}'");
                }
            }
        } else if (*p == '[') {
            v.kind = Json::Arr;
            p++;
            ws();
            if (p < end && *p == ']') { p++; }
            else {
                while (true) {
                    v.arr.push_back(value());
                    ws();
                    if (p < end && *p == ',') { p++; continue; }
                    if (p < end && *p == ']') { p++; break; }
                    fail("expected ',' or ']'");
                }
            }
        } else if (lit("true")) { v.kind = Json::Bool; v.b = true; }
        else if (lit("false")) { v.kind = Json::Bool; v.b = false; }
        else if (lit("null")) { v.kind = Json::Null; }
        else {
            // a number: copy the token (strtod needs a terminator; the buffer may not have one)
            const char* q = p;
            while (q < end && (std::strchr("+-.eE", *q) || (*q >= '0' && *q <= '9'))) q++;
            if (q == p || q - p > 64) fail("unexpected character");
            char tok[72];
            std::memcpy(tok, p, size_t(q - p));
            tok[q - p] = 0;
            char* e = nullptr;
            v.num = std::strtod(tok, &e);
            if (e == tok || *e) fail("malformed number");
            v.kind = Json::Num;
            p = q;
        }
        depth--;
        return v;
    }
};

// ------------------------------------------------------------------------------------------------------------------- container
std::vector<uint8_t> read_file(const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw GltfError("cannot open " + path);
    std::vector<uint8_t> data;
    uint8_t buf[1 << 16];
    size_t n;
    while ((n = std::fread(buf, 1, sizeof(buf), f)) > 0) {
        if (data.size() + n > (size_t(1) << 34)) { std::fclose(f); throw GltfError(path + ": larger than 16 GiB"); }
        data.insert(data.end(), buf, buf + n);
    }
    std::fclose(f);
    return data;
}

std::vector<uint8_t> base64_decode(const char* s, size_t n) {
    std::vector<uint8_t> out;
    out.reserve(n / 4 * 3);
    uint32_t acc = 0;
    int bits = 0;
    for (size_t i = 0; i < n; i++) {
        const char c = s[i];
        int v;
        if (c >= 'A' && c <= 'Z') v = c - 'A';
        else if (c >= 'a' && c <= 'z') v = c - 'a' + 26;
        else if (c >= '0' && c <= '9') v = c - '0' + 52;
        else if (c == '+' || c == '-') v = 62;
        else if (c == '/' || c == '_') v = 63;
        else continue; // padding, line breaks
        acc = (acc << 6) | uint32_t(v);
        bits += 6;
        if (bits >= 8) {
            bits -= 8;
            out.push_back(uint8_t((acc >> bits) & 0xFFu));
        }
    }
    return out;
}

std::string dirname_of(const std::string& path) {
    const size_t s = path.find_last_of('/');
    return s == std::string::npos ? std::string(".") : path.substr(0, s);
}

std::vector<uint8_t> load_uri(const std::string& uri, const std::string& base) {
    if (uri.compare(0, 5, "data:") == 0) {
        const size_t comma = uri.find(',');
        if (comma == std::string::npos) throw GltfError("data URI without a payload");
        return base64_decode(uri.c_str() + comma + 1, uri.size() - comma - 1);
    }
    return read_file(base + "/" + uri);
}

uint32_t rd_u32(const uint8_t* p) { uint32_t v; std::memcpy(&v, p, 4); return v; }

struct Doc {
    Json j;
    std::string base;
    std::vector<std::vector<uint8_t>> buffers;

    explicit Doc(const std::string& path) : base(dirname_of(path)) {
        const std::vector<uint8_t> data = read_file(path);
        std::vector<uint8_t> glb;
        bool have_glb = false;
        if (data.size() >= 12 && std::memcmp(data.data(), "glTF", 4) == 0) {
            if (rd_u32(&data[4]) != 2) throw GltfError(path + ": not a glTF 2.0 binary container");
            const size_t length = std::min<size_t>(rd_u32(&data[8]), data.size());
            size_t off = 12;
            bool have_json = false;
            while (off + 8 <= length) {
                const size_t clen = rd_u32(&data[off]);
                const uint32_t ctype = rd_u32(&data[off + 4]);
                if (clen > length - off - 8) throw GltfError(path + ": GLB chunk overruns the file");
                const uint8_t* chunk = &data[off + 8];
                if (ctype == 0x4E4F534Au) { // JSON
                    JsonParser P(reinterpret_cast<const char*>(chunk), clen);
                    j = P.value();
                    have_json = true;
                } else if (ctype == 0x004E4942u && !have_glb) { // BIN
                    glb.assign(chunk, chunk + clen);
                    have_glb = true;
                }
                off += 8 + clen;
            }
            if (!have_json) throw GltfError(path + ": GLB without a JSON chunk");
        } else {
            JsonParser P(reinterpret_cast<const char*>(data.data()), data.size());
            j = P.value();
        }
        if (j.kind != Json::Obj) throw GltfError(path + ": the document is not a JSON object");
        const auto& bufs = j.array("buffers");
        for (size_t i = 0; i < bufs.size(); i++) {
            const Json* uri = bufs[i].get("uri");
            if (uri && uri->kind == Json::Str) buffers.push_back(load_uri(uri->str, base));
            else if (i == 0 && have_glb) buffers.push_back(glb);
            else throw GltfError("buffer " + std::to_string(i) + " has no data");
        }
    }

    // the bytes of a buffer view (pointer + length) and its stride
    void view_bytes(int64_t index, const uint8_t*& ptr, size_t& len, size_t& stride) const {
        const auto& views = j.array("bufferViews");
        if (index < 0 || size_t(index) >= views.size()) throw GltfError("bufferView index out of range");
        const Json& v = views[size_t(index)];
        const int64_t b = v.integer("buffer", -1), off = v.integer("byteOffset", 0), n = v.integer("byteLength", -1);
        if (b < 0 || size_t(b) >= buffers.size() || off < 0 || n < 0) throw GltfError("malformed bufferView");
        const std::vector<uint8_t>& buf = buffers[size_t(b)];
        const size_t o = std::min<size_t>(size_t(off), buf.size());
        ptr = buf.data() + o;
        len = std::min<size_t>(size_t(n), buf.size() - o); // like a Python slice: clipped to what exists
        stride = size_t(std::max<int64_t>(v.integer("byteStride", 0), 0));
    }
};

struct Accessor {
    int component = 0;  // 5120 ... 5126
    size_t comp_size = 0, width = 0, count = 0;
    std::vector<uint8_t> data; // count x width components, tightly packed
};

size_t component_size(int c) {
    switch (c) {
    case 5120: case 5121: return 1;
    case 5122: case 5123: return 2;
    case 5125: case 5126: return 4;
    default: throw GltfError("unknown accessor component type " + std::to_string(c));
    }
}
size_t type_width(const std::string& t) {
    if (t == "SCALAR") return 1;
    if (t == "VEC2") return 2;
    if (t == "VEC3") return 3;
    if (t == "VEC4") return 4;
    if (t == "MAT4") return 16;
    throw GltfError("unsupported accessor type " + t);
}

// get_accessor_base / get_accessor_byte_stride + read_strided (gltf_scene.cpp:30-60)
Accessor read_accessor(const Doc& d, int64_t index, const char* expect_float) {
    const auto& accs = d.j.array("accessors");
    if (index < 0 || size_t(index) >= accs.size()) throw GltfError("accessor index out of range");
    const Json& a = accs[size_t(index)];
    Accessor out;
    out.component = int(a.integer("componentType", 0));
    out.comp_size = component_size(out.component);
    const Json* ty = a.get("type");
    out.width = type_width(ty && ty->kind == Json::Str ? ty->str : std::string("?"));
    if (expect_float && out.component != 5126)
        throw GltfError("accessor " + std::to_string(index) + " (" + expect_float + ") is not FLOAT: the reference reads these attributes as float (gltf_scene.cpp:127-146)");
    const int64_t count = a.integer("count", -1);
    if (count < 0 || count > int64_t(1) << 31) throw GltfError("accessor count out of range");
    out.count = size_t(count);
    const size_t elem = out.comp_size * out.width;
    out.data.assign(out.count * elem, uint8_t(0));
    if (!a.has("bufferView") || out.count == 0) return out;
    const uint8_t* raw;
    size_t len, stride;
    d.view_bytes(a.integer("bufferView", -1), raw, len, stride);
    const int64_t off = a.integer("byteOffset", 0);
    if (off < 0) throw GltfError("negative accessor byteOffset");
    if (stride == 0) stride = elem;
    const size_t need = size_t(off) + (out.count - 1) * stride + elem;
    if (need > len) throw GltfError("accessor " + std::to_string(index) + " overruns its buffer view");
    for (size_t i = 0; i < out.count; i++) std::memcpy(&out.data[i * elem], raw + size_t(off) + i * stride, elem);
    return out;
}

// ------------------------------------------------------------------------------------------------------------------- transforms
struct M4 { float m[16]; }; // row-major
M4 identity4() { return M4{{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}}; }
// hlslpp mul(float4x4, float4x4) in fp32: rows of a dotted with columns of b, accumulated left to right
M4 mul4(const M4& a, const M4& b) {
    M4 r;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            float acc = a.m[i * 4] * b.m[j];
            for (int k = 1; k < 4; k++) acc = acc + a.m[i * 4 + k] * b.m[k * 4 + j];
            r.m[i * 4 + j] = acc;
        }
    return r;
}
// glm::mat4_cast (gltf_scene.cpp:173-180), fp32
M4 quat_to_matrix(float x, float y, float z, float w) {
    const float qxx = x * x, qyy = y * y, qzz = z * z, qxz = x * z, qxy = x * y, qyz = y * z, qwx = w * x, qwy = w * y, qwz = w * z;
    M4 R = identity4();
    R.m[0] = 1.f - 2.f * (qyy + qzz); R.m[4] = 2.f * (qxy + qwz); R.m[8] = 2.f * (qxz - qwy);
    R.m[1] = 2.f * (qxy - qwz); R.m[5] = 1.f - 2.f * (qxx + qzz); R.m[9] = 2.f * (qyz + qwx);
    R.m[2] = 2.f * (qxz + qwy); R.m[6] = 2.f * (qyz - qwx); R.m[10] = 1.f - 2.f * (qxx + qyy);
    return R;
}
// gltf_node_transform, gltf_scene.cpp:146-196
M4 node_transform(const Json& n) {
    const auto& mat = n.array("matrix");
    if (mat.size() == 16) {
        M4 r;
        for (int c = 0; c < 4; c++)
            for (int rr = 0; rr < 4; rr++) r.m[rr * 4 + c] = float(mat[size_t(c * 4 + rr)].num); // column-major in the file
        return r;
    }
    M4 T = identity4(), R = identity4(), S = identity4();
    const auto& t = n.array("translation");
    if (t.size() == 3) { T.m[3] = float(t[0].num); T.m[7] = float(t[1].num); T.m[11] = float(t[2].num); }
    const auto& q = n.array("rotation");
    if (q.size() == 4) R = quat_to_matrix(float(q[0].num), float(q[1].num), float(q[2].num), float(q[3].num));
    const auto& s = n.array("scale");
    if (s.size() == 3) { S.m[0] = float(s[0].num); S.m[5] = float(s[1].num); S.m[10] = float(s[2].num); }
    return mul4(mul4(T, R), S);
}

uint32_t wrap_mode(int64_t gl) { return gl == 33071 ? MPT_WRAP_CLAMP_TO_EDGE : (gl == 33648 ? MPT_WRAP_MIRRORED_REPEAT : MPT_WRAP_REPEAT); }

} // namespace

// ------------------------------------------------------------------------------------------------------------------- the loaded scene
struct mpt_gltf {
    // storage
    struct MeshStore { std::vector<float> pos, nrm, uv, tan; std::vector<uint8_t> idx; };
    std::vector<MeshStore> mesh_store;
    std::vector<std::vector<uint32_t>> node_meshes;
    std::vector<std::vector<uint8_t>> images; // encoded bytes per texture slot
    // views
    std::vector<mpt_src_mesh> meshes;
    std::vector<mpt_src_node> nodes;
    std::vector<mpt_src_material> materials;
    std::vector<mpt_gltf_texture> textures;
    mpt_gltf_view view{};
};

namespace {

mpt_status gltf_fail(mpt_status s, const std::string& msg) { return mpt_internal_fail(nullptr, s, msg.c_str()); }

void load_into(const std::string& path, mpt_gltf& G) {
    const Doc doc(path);
    const Json& j = doc.j;
    const auto& jtextures = j.array("textures");
    const auto& jimages = j.array("images");
    const auto& jsamplers = j.array("samplers");

    // ---- texture slots in order of first use (get_or_load_texture, gltf_scene.cpp:538-604)
    std::map<std::pair<int64_t, bool>, uint32_t> slots;
    auto get_or_load_texture = [&](int64_t index, bool linear) -> uint32_t {
        if (index < 0 || size_t(index) >= jtextures.size()) return MPT_TEXTURE_NONE;
        const auto key = std::make_pair(index, linear);
        const auto it = slots.find(key);
        if (it != slots.end()) return it->second;
        const Json& tex = jtextures[size_t(index)];
        const int64_t src = tex.integer("source", -1);
        if (src < 0 || size_t(src) >= jimages.size()) return MPT_TEXTURE_NONE;
        const Json& img = jimages[size_t(src)];
        std::vector<uint8_t> bytes;
        const Json* uri = img.get("uri");
        if (uri && uri->kind == Json::Str) {
            bytes = load_uri(uri->str, doc.base);
        } else if (img.has("bufferView")) {
            const uint8_t* p;
            size_t len, stride;
            doc.view_bytes(img.integer("bufferView", -1), p, len, stride);
            bytes.assign(p, p + len);
        }
        if (bytes.empty()) return MPT_TEXTURE_NONE; // (the reference also leaves the slot empty when stb_image rejects the bytes: that check is the caller's here)
        mpt_gltf_texture t{};
        t.srgb = linear ? 0u : 1u;
        t.filter = MPT_FILTER_LINEAR;
        t.wrap_u = t.wrap_v = MPT_WRAP_REPEAT;
        t.mipmap = MPT_MIPMAP_NONE;
        const int64_t s = tex.integer("sampler", -1);
        if (s >= 0 && size_t(s) < jsamplers.size()) {
            const Json& sm = jsamplers[size_t(s)];
            if (sm.integer("magFilter", -1) == 9728) t.filter = MPT_FILTER_NEAREST; // libmpt has one filter per texture: mag wins
            switch (sm.integer("minFilter", -1)) { // gltf_min_filter_to_vk, gltf_scene.cpp:212-248: the *_MIPMAP_* values ask for a chain
            case 9984: case 9985: t.mipmap = MPT_MIPMAP_NEAREST; break;
            case 9986: case 9987: t.mipmap = MPT_MIPMAP_LINEAR; break;
            default: break;
            }
            t.wrap_u = wrap_mode(sm.integer("wrapS", 10497));
            t.wrap_v = wrap_mode(sm.integer("wrapT", 10497));
        }
        const uint32_t id = uint32_t(G.textures.size());
        if (id >= MPT_TEXTURE_NONE) throw GltfError("more than 65534 textures");
        G.images.push_back(std::move(bytes));
        G.textures.push_back(t);
        slots[key] = id;
        return id;
    };
    auto tex_index = [](const Json& o, const char* key) -> int64_t {
        const Json* t = o.get(key);
        return t && t->kind == Json::Obj ? t->integer("index", -1) : -1;
    };
    auto fill3 = [](float* dst, const std::vector<Json>& a, size_t n, const float* def) {
        for (size_t k = 0; k < n; k++) dst[k] = k < a.size() && a[k].kind == Json::Num ? float(a[k].num) : def[k];
    };

    // ---- materials (load_materials, gltf_scene.cpp:265-536)
    const auto& jmaterials = j.array("materials");
    auto default_material = [] {
        mpt_src_material m{};
        m.model = MPT_MATERIAL_GLTF;
        for (float& c : m.base_color_factor) c = 1.f;
        m.metallic_factor = m.roughness_factor = m.normal_scale = m.occlusion_strength = 1.f;
        m.ior = 1.5f;
        m.base_color_texture = m.metallic_roughness_texture = m.normal_texture = m.occlusion_texture = m.emissive_texture = MPT_TEXTURE_NONE;
        m.clearcoat_texture = m.clearcoat_roughness_texture = m.sheen_color_texture = m.sheen_roughness_texture = MPT_TEXTURE_NONE;
        m.iridescence_texture = m.iridescence_thickness_texture = m.anisotropy_texture = MPT_TEXTURE_NONE;
        return m;
    };
    static const float ones4[4] = {1, 1, 1, 1}, zeros3[3] = {0, 0, 0}, ones3[3] = {1, 1, 1};
    for (const Json& gm : jmaterials) {
        mpt_src_material m = default_material();
        const Json& pbr = gm.object("pbrMetallicRoughness");
        if (pbr.has("baseColorFactor")) fill3(m.base_color_factor, pbr.array("baseColorFactor"), 4, ones4);
        m.metallic_factor = float(pbr.number("metallicFactor", 1.0));
        m.roughness_factor = float(pbr.number("roughnessFactor", 1.0));
        if (gm.has("emissiveFactor")) fill3(m.emissive_factor, gm.array("emissiveFactor"), 3, zeros3);
        if (tex_index(pbr, "baseColorTexture") >= 0) m.base_color_texture = get_or_load_texture(tex_index(pbr, "baseColorTexture"), false);
        if (tex_index(pbr, "metallicRoughnessTexture") >= 0) m.metallic_roughness_texture = get_or_load_texture(tex_index(pbr, "metallicRoughnessTexture"), true);
        if (tex_index(gm, "normalTexture") >= 0) {
            m.normal_texture = get_or_load_texture(tex_index(gm, "normalTexture"), true);
            m.normal_scale = float(gm.object("normalTexture").number("scale", 1.0));
        }
        if (tex_index(gm, "emissiveTexture") >= 0) m.emissive_texture = get_or_load_texture(tex_index(gm, "emissiveTexture"), false);
        if (tex_index(gm, "occlusionTexture") >= 0) {
            m.occlusion_texture = get_or_load_texture(tex_index(gm, "occlusionTexture"), true);
            m.occlusion_strength = float(gm.object("occlusionTexture").number("strength", 1.0));
        }
        const Json& ext = gm.object("extensions");
        // an extension only produces a payload when its weight gate passes
        if (const Json* e = ext.get("KHR_materials_transmission")) {
            if (e->has("transmissionFactor") && float(e->number("transmissionFactor", 0.0)) > 0.f) {
                m.features |= MPT_GLTF_TRANSMISSION;
                m.transmission_weight = float(e->number("transmissionFactor", 0.0));
            }
        }
        if (const Json* e = ext.get("KHR_materials_ior")) {
            if (e->has("ior")) m.ior = float(e->number("ior", 1.5));
        }
        if (const Json* e = ext.get("KHR_materials_volume")) {
            const double thickness = e->number("thicknessFactor", 0.0);
            const double dist = e->number("attenuationDistance", std::numeric_limits<double>::infinity());
            float col[3];
            fill3(col, e->array("attenuationColor"), 3, ones3);
            if (thickness > 0.0 && std::isfinite(dist) && dist > 0.0) {
                m.features |= MPT_GLTF_VOLUME;
                for (int k = 0; k < 3; k++) // absorption = -ln(clamp(c, 1e-4, 1)) / distance, fp32
                    m.volume_absorption[k] = -std::log(std::min(std::max(col[k], 1e-4f), 1.0f)) / float(dist);
            }
        }
        if (const Json* e = ext.get("KHR_materials_clearcoat")) {
            const float weight = float(e->number("clearcoatFactor", 0.0));
            uint32_t t0 = MPT_TEXTURE_NONE, t1 = MPT_TEXTURE_NONE;
            if (tex_index(*e, "clearcoatTexture") >= 0) t0 = get_or_load_texture(tex_index(*e, "clearcoatTexture"), true);
            if (tex_index(*e, "clearcoatRoughnessTexture") >= 0) t1 = get_or_load_texture(tex_index(*e, "clearcoatRoughnessTexture"), true);
            if (weight > 0.f) {
                m.features |= MPT_GLTF_CLEARCOAT;
                m.clearcoat_weight = weight;
                m.clearcoat_roughness = float(e->number("clearcoatRoughnessFactor", 0.0));
                m.clearcoat_texture = t0;
                m.clearcoat_roughness_texture = t1;
            }
        }
        if (const Json* e = ext.get("KHR_materials_sheen")) {
            float col[3];
            fill3(col, e->array("sheenColorFactor"), 3, zeros3);
            uint32_t t0 = MPT_TEXTURE_NONE, t1 = MPT_TEXTURE_NONE;
            if (tex_index(*e, "sheenColorTexture") >= 0) t0 = get_or_load_texture(tex_index(*e, "sheenColorTexture"), false);
            if (tex_index(*e, "sheenRoughnessTexture") >= 0) t1 = get_or_load_texture(tex_index(*e, "sheenRoughnessTexture"), true);
            if (col[0] > 0.f || col[1] > 0.f || col[2] > 0.f) {
                m.features |= MPT_GLTF_SHEEN;
                std::memcpy(m.sheen_color, col, sizeof(col));
                m.sheen_roughness = float(e->number("sheenRoughnessFactor", 0.0));
                m.sheen_color_texture = t0;
                m.sheen_roughness_texture = t1;
            }
        }
        if (const Json* e = ext.get("KHR_materials_iridescence")) {
            const float factor = float(e->number("iridescenceFactor", 0.0));
            uint32_t t0 = MPT_TEXTURE_NONE, t1 = MPT_TEXTURE_NONE;
            if (tex_index(*e, "iridescenceTexture") >= 0) t0 = get_or_load_texture(tex_index(*e, "iridescenceTexture"), true);
            if (tex_index(*e, "iridescenceThicknessTexture") >= 0) t1 = get_or_load_texture(tex_index(*e, "iridescenceThicknessTexture"), true);
            if (factor > 0.f) {
                m.features |= MPT_GLTF_IRIDESCENCE;
                m.iridescence_factor = factor;
                m.iridescence_ior = float(e->number("iridescenceIor", 1.3));
                m.iridescence_thickness_min = float(e->number("iridescenceThicknessMinimum", 100.0));
                m.iridescence_thickness_max = float(e->number("iridescenceThicknessMaximum", 400.0));
                m.iridescence_texture = t0;
                m.iridescence_thickness_texture = t1;
            }
        }
        if (const Json* e = ext.get("KHR_materials_anisotropy")) {
            const float strength = float(e->number("anisotropyStrength", 0.0));
            uint32_t t0 = MPT_TEXTURE_NONE;
            if (tex_index(*e, "anisotropyTexture") >= 0) t0 = get_or_load_texture(tex_index(*e, "anisotropyTexture"), true);
            if (strength > 0.f) {
                m.features |= MPT_GLTF_ANISOTROPY;
                m.anisotropy_strength = strength;
                m.anisotropy_rotation = float(e->number("anisotropyRotation", 0.0));
                m.anisotropy_texture = t0;
            }
        }
        G.materials.push_back(m);
    }
    if (G.materials.empty()) G.materials.push_back(default_material()); // default material for primitives without one (gltf_scene.cpp:530-535)

    // ---- meshes (load_meshes): one Scene::Mesh per TRIANGLES primitive with a POSITION attribute
    std::vector<std::vector<uint32_t>> mesh_map;
    for (const Json& gmesh : j.array("meshes")) {
        std::vector<uint32_t> ids;
        for (const Json& prim : gmesh.array("primitives")) {
            const Json& at = prim.object("attributes");
            if (prim.integer("mode", 4) != 4 || !at.has("POSITION")) continue;
            mpt_gltf::MeshStore st;
            auto floats = [&](const char* name, size_t want, std::vector<float>& dst) {
                const Accessor a = read_accessor(doc, at.integer(name, -1), name);
                if (a.width < want) throw GltfError(std::string(name) + " accessor is narrower than expected");
                dst.resize(a.count * want);
                for (size_t i = 0; i < a.count; i++) std::memcpy(&dst[i * want], &a.data[i * a.width * 4], want * 4);
                return a.count;
            };
            const size_t n = floats("POSITION", 3, st.pos);
            if (at.has("NORMAL")) {
                if (floats("NORMAL", 3, st.nrm) != n) throw GltfError("NORMAL count differs from POSITION count");
            } else {
                st.nrm.resize(n * 3);
                for (size_t i = 0; i < n; i++) { st.nrm[i * 3] = 0.f; st.nrm[i * 3 + 1] = 1.f; st.nrm[i * 3 + 2] = 0.f; }
            }
            if (at.has("TEXCOORD_0") && floats("TEXCOORD_0", 2, st.uv) != n) throw GltfError("TEXCOORD_0 count differs from POSITION count");
            if (at.has("TANGENT") && floats("TANGENT", 4, st.tan) != n) throw GltfError("TANGENT count differs from POSITION count");
            mpt_src_mesh m{};
            m.vertex_count = uint32_t(n);
            m.primitive_count = uint32_t(n / 3);
            if (prim.integer("indices", -1) >= 0) {
                const Accessor a = read_accessor(doc, prim.integer("indices", -1), nullptr);
                if (a.component != 5121 && a.component != 5123 && a.component != 5125) throw GltfError("index accessor must be UNSIGNED_BYTE / UNSIGNED_SHORT / UNSIGNED_INT");
                const size_t total = a.count * a.width, tris = total / 3;
                st.idx.assign(a.data.begin(), a.data.begin() + long(tris * 3 * a.comp_size));
                m.index_size = uint32_t(a.comp_size);
                m.primitive_count = uint32_t(tris);
            }
            const int64_t mi = prim.integer("material", -1);
            uint32_t flags = 0;
            if (mi >= 0 && size_t(mi) < jmaterials.size()) {
                const Json& gm = jmaterials[size_t(mi)];
                const Json* am = gm.get("alphaMode");
                if (!am || am->kind != Json::Str || am->str == "OPAQUE") flags |= MPT_MESH_IS_OPAQUE;
                const Json* ds = gm.get("doubleSided");
                if ((ds && ds->kind == Json::Bool && ds->b) || gm.object("extensions").has("KHR_materials_transmission"))
                    flags |= MPT_MESH_TWO_SIDED; // two-sided, or transmissive: refraction must hit the exit interface
            } else {
                flags |= MPT_MESH_IS_OPAQUE;
            }
            if (!st.tan.empty()) flags |= MPT_MESH_HAS_TANGENTS;
            m.flags = flags;
            m.material_id = (mi >= 0 && size_t(mi) < G.materials.size()) ? uint32_t(mi) : 0u;
            m.instance_mask = 0x01; // Scene::Mesh's default (scene.hpp:193): glTF cannot say "instance mask"
            ids.push_back(uint32_t(G.meshes.size()));
            G.mesh_store.push_back(std::move(st));
            G.meshes.push_back(m);
        }
        mesh_map.push_back(std::move(ids));
    }
    for (size_t i = 0; i < G.meshes.size(); i++) { // pointers after the stores stopped moving
        const mpt_gltf::MeshStore& st = G.mesh_store[i];
        G.meshes[i].positions = st.pos.data();
        G.meshes[i].normals = st.nrm.data();
        G.meshes[i].uvs = st.uv.empty() ? nullptr : st.uv.data();
        G.meshes[i].tangents = st.tan.empty() ? nullptr : st.tan.data();
        G.meshes[i].indices = G.meshes[i].index_size ? st.idx.data() : nullptr;
    }

    // ---- scene graph (load_node): NodeIDs in DFS pre-order over the default scene's roots
    const auto& gnodes = j.array("nodes");
    const auto& scenes = j.array("scenes");
    if (scenes.empty()) throw GltfError("glTF has no scene");
    int64_t scene_index = j.integer("scene", -1);
    if (scene_index < 0 || size_t(scene_index) >= scenes.size()) scene_index = 0;
    std::vector<M4> globals;
    std::map<int64_t, uint32_t> node_map;
    struct Item { int64_t gi; int parent; }; // parent = index into globals, -1 = root
    std::vector<Item> stack;
    const auto& roots = scenes[size_t(scene_index)].array("nodes");
    for (size_t r = roots.size(); r-- > 0;) stack.push_back(Item{int64_t(roots[r].num), -1});
    size_t visited = 0;
    while (!stack.empty()) {
        const Item it = stack.back();
        stack.pop_back();
        if (it.gi < 0 || size_t(it.gi) >= gnodes.size()) throw GltfError("node index out of range");
        if (++visited > gnodes.size() * 4 + 16) throw GltfError("the node hierarchy is not a tree");
        const Json& gn = gnodes[size_t(it.gi)];
        const M4 local = node_transform(gn);
        const M4 glob = it.parent < 0 ? local : mul4(globals[size_t(it.parent)], local);
        const uint32_t id = uint32_t(G.nodes.size());
        node_map[it.gi] = id;
        globals.push_back(glob);
        mpt_src_node n{};
        std::memcpy(n.transform, glob.m, sizeof(glob.m));
        G.nodes.push_back(n);
        const int64_t mesh = gn.integer("mesh", -1);
        G.node_meshes.push_back(mesh >= 0 && size_t(mesh) < mesh_map.size() ? mesh_map[size_t(mesh)] : std::vector<uint32_t>());
        const auto& ch = gn.array("children");
        for (size_t c = ch.size(); c-- > 0;) stack.push_back(Item{int64_t(ch[c].num), int(id)}); // reversed: the first child is visited first
    }
    for (size_t i = 0; i < G.nodes.size(); i++) {
        G.nodes[i].meshes = G.node_meshes[i].empty() ? nullptr : G.node_meshes[i].data();
        G.nodes[i].mesh_count = uint32_t(G.node_meshes[i].size());
    }

    // ---- camera (load_cameras): the first perspective camera node of the scene becomes the active camera
    const auto& jcameras = j.array("cameras");
    G.view.has_camera = 0;
    for (size_t gi = 0; gi < gnodes.size() && !G.view.has_camera; gi++) {
        const int64_t ci = gnodes[gi].integer("camera", -1);
        const auto nm = node_map.find(int64_t(gi));
        if (ci < 0 || size_t(ci) >= jcameras.size() || nm == node_map.end()) continue;
        const Json& gc = jcameras[size_t(ci)];
        const Json* ty = gc.get("type");
        if (!ty || ty->kind != Json::Str || ty->str != "perspective") continue;
        const float* M = G.nodes[nm->second].transform;
        auto column = [&](int c, float* out) {
            const float x = M[c], y = M[4 + c], z = M[8 + c];
            const float l = std::sqrt((x * x + y * y) + z * z);
            out[0] = x / l; out[1] = y / l; out[2] = z / l;
        };
        float up[3], fwd[3];
        column(1, up);
        column(2, fwd);
        const float eye[3] = {M[3], M[7], M[11]};
        const float target[3] = {eye[0] - fwd[0], eye[1] - fwd[1], eye[2] - fwd[2]}; // the target one unit in front of the eye, gltf_scene.cpp:736
        const Json& p = gc.object("perspective");
        const float yfov = float(p.number("yfov", 60.0 * 3.14159265358979323846 / 180.0));
        G.view.yfov_degrees = yfov * 57.29577951308232f; // rad2deg in fp32
        G.view.aspect = p.number("aspectRatio", 0.0) > 0.0 ? float(p.number("aspectRatio", 1.0)) : 1.0f;
        G.view.znear = float(p.number("znear", 0.01));
        G.view.zfar = p.number("zfar", 0.0) > 0.0 ? float(p.number("zfar", 1000.0)) : 1000.0f;
        if (mptt_camera_look_at(eye, target, up, G.view.yfov_degrees, G.view.aspect, G.view.znear, G.view.zfar, 0.f, nullptr, &G.view.camera) != MPT_OK)
            throw GltfError("degenerate camera node transform");
        G.view.has_camera = 1;
    }

    for (size_t i = 0; i < G.textures.size(); i++) {
        G.textures[i].encoded = G.images[i].data();
        G.textures[i].encoded_size = G.images[i].size();
    }
    G.view.meshes = G.meshes.data(); G.view.mesh_count = uint32_t(G.meshes.size());
    G.view.nodes = G.nodes.data(); G.view.node_count = uint32_t(G.nodes.size());
    G.view.materials = G.materials.data(); G.view.material_count = uint32_t(G.materials.size());
    G.view.textures = G.textures.data(); G.view.texture_count = uint32_t(G.textures.size());
}

} // namespace

MPT_API mpt_status mptt_gltf_load(const char* path, mpt_gltf** out) try {
    if (!path || !out) return gltf_fail(MPT_ERR_INVALID_ARGUMENT, "mptt_gltf_load: NULL argument");
    *out = nullptr;
    std::unique_ptr<mpt_gltf> G(new mpt_gltf);
    load_into(path, *G);
    *out = G.release();
    return MPT_OK;
} catch (const GltfError& e) {
    return gltf_fail(MPT_ERR_IO, std::string("mptt_gltf_load: ") + e.what());
} catch (const std::bad_alloc&) {
    return gltf_fail(MPT_ERR_LIMIT, "mptt_gltf_load: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return gltf_fail(MPT_ERR_IO, std::string("mptt_gltf_load: ") + e.what());
} catch (...) {
    return gltf_fail(MPT_ERR_IO, "mptt_gltf_load: unknown exception");
}

MPT_API mpt_status mptt_gltf_get(const mpt_gltf* g, mpt_gltf_view* out) {
    if (!g || !out) return gltf_fail(MPT_ERR_INVALID_ARGUMENT, "mptt_gltf_get: NULL argument");
    *out = g->view;
    return MPT_OK;
}

MPT_API void mptt_gltf_free(mpt_gltf* g) { delete g; }
