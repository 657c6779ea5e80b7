// oracle_math.hpp — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// CPU restatement of the reference's Slang utility layer: vector helpers with HLSL/GLSL semantics,
// hash + RNG, sampling warps, orthonormal frames, normal/tangent encodings, Fresnel terms.
// Every function cites the reference file:line (relative to /root/reference/) it follows.
//
// PARITY UNPINNED for everything except the BSDF layer: the reference's tests hold no golden vectors
// for RNG, camera rays, hits or images; the BSDF layer is pinned by the property
// suite of tests/test_bsdf.cpp, restated in tests/test_bsdf_oracle.py.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <optional>

namespace orc {

constexpr float PI = 3.14159265358979323846f;        // MERIAN_PI
constexpr float TWO_PI = 6.28318530717958647692f;    // MERIAN_TWO_PI
constexpr float INV_PI = 0.31830988618379067154f;    // MERIAN_INV_PI / MERIAN_M_1_PI
constexpr float INV_TWO_PI = 0.15915494309189533577f; // MERIAN_1_2PI

// ---------------------------------------------------------------- vectors (HLSL semantics)
struct float2 {
    float x = 0, y = 0;
    float2() = default;
    float2(float a, float b) : x(a), y(b) {}
    explicit float2(float a) : x(a), y(a) {}
};
struct float3 {
    float x = 0, y = 0, z = 0;
    float3() = default;
    float3(float a, float b, float c) : x(a), y(b), z(c) {}
    explicit float3(float a) : x(a), y(a), z(a) {}
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
struct float4 {
    float x = 0, y = 0, z = 0, w = 0;
    float4() = default;
    float4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    float4(float3 v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
    float3 xyz() const { return {x, y, z}; }
};

inline float2 operator+(float2 a, float2 b) { return {a.x + b.x, a.y + b.y}; }
inline float2 operator-(float2 a, float2 b) { return {a.x - b.x, a.y - b.y}; }
inline float2 operator*(float2 a, float2 b) { return {a.x * b.x, a.y * b.y}; }
inline float2 operator*(float2 a, float s) { return {a.x * s, a.y * s}; }
inline float2 operator*(float s, float2 a) { return {a.x * s, a.y * s}; }
inline float2 operator/(float2 a, float s) { return {a.x / s, a.y / s}; }
inline float dot(float2 a, float2 b) { return a.x * b.x + a.y * b.y; }

inline float3 operator+(float3 a, float3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline float3 operator-(float3 a, float3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float3 operator-(float3 a) { return {-a.x, -a.y, -a.z}; }
inline float3 operator*(float3 a, float3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline float3 operator*(float3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float3 operator*(float s, float3 a) { return {a.x * s, a.y * s, a.z * s}; }
inline float3 operator/(float3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline float3 operator/(float3 a, float3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
inline float3 operator+(float3 a, float s) { return {a.x + s, a.y + s, a.z + s}; }
inline float3 operator+(float s, float3 a) { return {a.x + s, a.y + s, a.z + s}; }
inline float3 operator-(float3 a, float s) { return {a.x - s, a.y - s, a.z - s}; }
inline float3 operator-(float s, float3 a) { return {s - a.x, s - a.y, s - a.z}; }
inline float3& operator+=(float3& a, float3 b) { a = a + b; return a; }
inline float3& operator*=(float3& a, float3 b) { a = a * b; return a; }
inline float3& operator*=(float3& a, float s) { a = a * s; return a; }
inline float3& operator/=(float3& a, float s) { a = a / s; return a; }
inline float4 operator+(float4 a, float4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline float4 operator*(float s, float4 a) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
inline float4 operator*(float4 a, float4 b) { return {a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w}; }

inline float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float3 cross(float3 a, float3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline float length(float3 a) { return std::sqrt(dot(a, a)); }
inline float distance(float3 a, float3 b) { return length(a - b); }
inline float3 normalize(float3 a) { return a / length(a); }
inline float2 normalize(float2 a) { return a / std::sqrt(dot(a, a)); }
inline float3 abs3(float3 a) { return {std::fabs(a.x), std::fabs(a.y), std::fabs(a.z)}; }
inline float3 max3v(float3 a, float3 b) { return {std::max(a.x, b.x), std::max(a.y, b.y), std::max(a.z, b.z)}; }
inline float3 max3v(float3 a, float b) { return {std::max(a.x, b), std::max(a.y, b), std::max(a.z, b)}; }
inline float3 sqrt3(float3 a) { return {std::sqrt(a.x), std::sqrt(a.y), std::sqrt(a.z)}; }
inline float3 exp3(float3 a) { return {std::exp(a.x), std::exp(a.y), std::exp(a.z)}; }
inline float3 cos3(float3 a) { return {std::cos(a.x), std::cos(a.y), std::cos(a.z)}; }
inline float clampf(float x, float lo, float hi) { return std::min(std::max(x, lo), hi); }
inline float3 clamp3(float3 a, float lo, float hi) { return {clampf(a.x, lo, hi), clampf(a.y, lo, hi), clampf(a.z, lo, hi)}; }
inline float saturate(float x) { return clampf(x, 0.f, 1.f); }
// HLSL lerp / GLSL mix: x*(1-s) + y*s
inline float lerp(float a, float b, float s) { return a * (1.0f - s) + b * s; }
inline float3 lerp(float3 a, float3 b, float s) { return a * (1.0f - s) + b * s; }
inline float signf(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }
inline float smoothstep(float e0, float e1, float x) {
    const float t = saturate((x - e0) / (e1 - e0));
    return t * t * (3.0f - 2.0f * t);
}
inline float frac(float x) { return x - std::floor(x); }
inline bool all_lt(float3 a, float b) { return a.x < b && a.y < b && a.z < b; }
inline bool any_nan(float3 a) { return std::isnan(a.x) || std::isnan(a.y) || std::isnan(a.z); }
inline bool any_inf(float3 a) { return std::isinf(a.x) || std::isinf(a.y) || std::isinf(a.z); }

inline uint32_t as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline int32_t as_int(float f) { int32_t u; std::memcpy(&u, &f, 4); return u; }
inline float as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline float as_float_i(int32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

// HLSL reflect(i, n) = i - 2 n dot(i, n); refract per HLSL/GLSL spec (returns 0 on TIR).
inline float3 reflect(float3 i, float3 n) { return i - 2.0f * dot(n, i) * n; }
inline float3 refract(float3 i, float3 n, float eta) {
    const float ndoti = dot(n, i);
    const float k = 1.0f - eta * eta * (1.0f - ndoti * ndoti);
    if (k < 0.0f) return float3(0.f);
    return eta * i - (eta * ndoti + std::sqrt(k)) * n;
}

// include/merian-shaders/utils/math.slang:14-20, 112-127
inline float pow2(float x) { return x * x; }
inline float3 pow2(float3 x) { return x * x; }
inline float max3(float3 v) { return std::max(v.x, std::max(v.y, v.z)); }
inline float3 reflect2(float3 i, float3 n) { return 2.0f * dot(n, i) * n - i; }
inline float3 reflect2(float3 i, float3 n, float ndoti) { return 2.0f * ndoti * n - i; }

// ---------------------------------------------------------------- binary16
inline float half_to_float(uint16_t h) {
    const uint32_t s = (h >> 15) & 1u, e = (h >> 10) & 0x1Fu, m = h & 0x3FFu;
    uint32_t out;
    if (e == 0) {
        if (m == 0) {
            out = s << 31;
        } else { // subnormal
            int exp = -1;
            uint32_t mm = m;
            do { exp++; mm <<= 1; } while ((mm & 0x400u) == 0);
            out = (s << 31) | (uint32_t(127 - 15 - exp) << 23) | ((mm & 0x3FFu) << 13);
        }
    } else if (e == 31) {
        out = (s << 31) | 0x7F800000u | (m << 13);
    } else {
        out = (s << 31) | ((e + 112u) << 23) | (m << 13);
    }
    return as_float(out);
}
// round-to-nearest-even float -> half
inline uint16_t float_to_half(float f) {
    const uint32_t x = as_uint(f);
    const uint32_t s = (x >> 16) & 0x8000u;
    const uint32_t a = x & 0x7FFFFFFFu;
    if (a >= 0x7F800000u) return uint16_t(s | 0x7C00u | ((a > 0x7F800000u) ? 0x200u : 0u));
    if (a >= 0x477FF000u) return uint16_t(s | 0x7C00u); // rounds to inf (>= 65520)
    if (a < 0x33000001u) return uint16_t(s);            // rounds to zero (<= 2^-25)
    int e = int(a >> 23) - 127;
    uint32_t m = (a & 0x7FFFFFu) | 0x800000u;
    int shift;
    uint32_t he;
    if (e < -14) { shift = 13 + (-14 - e); he = 0; } else { shift = 13; he = uint32_t(e + 15); }
    const uint32_t lsb = 1u << shift, half_ulp = lsb >> 1;
    uint32_t r = m >> shift;
    const uint32_t rem = m & (lsb - 1);
    if (rem > half_ulp || (rem == half_ulp && (r & 1u))) r++;
    uint32_t out = (he == 0) ? r : (((he - 1) << 10) + r); // implicit bit carries into the exponent
    return uint16_t(s | out);
}

// ---------------------------------------------------------------- hash + RNG
// include/merian-shaders/utils/hash.slang:138-146
inline uint32_t pcg4d16(uint32_t vx, uint32_t vy, uint32_t vz, uint32_t vw) {
    vx = vx * 1664525u + 1013904223u;
    vy = vy * 1664525u + 1013904223u;
    vz = vz * 1664525u + 1013904223u;
    vw = vw * 1664525u + 1013904223u;
    vx += vy * vw;
    vy += vz * vx;
    vz += vx * vy;
    vw += vy * vz;
    vx += vy * vw;
    return vx;
}

// include/merian-shaders/utils/bit-twiddling.slang:23-27
inline float uint_to_zero_one_float_rust(uint32_t source) {
    return float(source >> 8) * 5.9604644775390625e-08f;
}

// RandomGenerator interface, include/merian-shaders/utils/random.slang:6-62. Dynamic dispatch is
// avoided by a tagged struct: kind 0 = PCG (random.slang:108-128), 1 = XorShift32 (:64-82).
struct Rng {
    uint32_t state;
    int kind;
    static Rng pcg(uint32_t seed) { return {seed, 0}; }
    static Rng xorshift(uint32_t seed) { return {seed, 1}; }
    uint32_t next() {
        if (kind == 0) {
            const uint32_t s = state;
            state = state * 747796405u + 2891336453u;
            const uint32_t word = ((s >> ((s >> 28u) + 4u)) ^ s) * 277803737u;
            return (word >> 22u) ^ word;
        }
        // xorshift32: Marsaglia "xor" (hash.slang xorshift32: x^=x<<13; x^=x>>17; x^=x<<5)
        uint32_t x = state;
        x ^= x << 13;
        x ^= x >> 17;
        x ^= x << 5;
        state = x;
        return x;
    }
    float next_float() { return uint_to_zero_one_float_rust(next()); }
    float2 next_float2() {
        float2 s;
        s.x = next_float();
        s.y = next_float();
        return s;
    }
};

// ---------------------------------------------------------------- sampling
// include/merian-shaders/utils/sampling.slang:24-29
inline float3 sample_cos(Rng& rng) {
    const float2 xi = rng.next_float2();
    const float su = std::sqrt(xi.x);
    return float3(su * std::cos(TWO_PI * xi.y), su * std::sin(TWO_PI * xi.y), std::sqrt(1.0f - xi.x));
}

// ---------------------------------------------------------------- frames
// include/merian-shaders/utils/frame.slang:7-13 (Duff et al. 2017); rows = (tangent, bitangent, z)
struct Frame {
    float3 t, b, n;

    Frame() = default;
    // frame.slang:88-90
    explicit Frame(float3 z) {
        const float sgn = (z.z >= 0) ? 1.f : -1.f;
        const float a = -1.0f / (sgn + z.z);
        const float bb = z.x * z.y * a;
        t = float3(1.0f + sgn * z.x * z.x * a, sgn * bb, -sgn * z.x);
        b = float3(bb, sgn + z.y * z.y * a, -z.y);
        n = z;
    }
    // frame.slang:94-99
    Frame(float3 nn, float3 target_tangent, float handedness = 1.0f) {
        const float3 N = normalize(nn);
        const float3 T = normalize(target_tangent - N * dot(target_tangent, N));
        const float3 B = cross(N, T) * handedness;
        t = T; b = B; n = N;
    }
    float3 to_local(float3 v) const { return float3(dot(t, v), dot(b, v), dot(n, v)); } // :42-44
    float3 to_world(float3 v) const {                                                   // :46-48 mul(v, frame)
        return float3(v.x * t.x + v.y * b.x + v.z * n.x, v.x * t.y + v.y * b.y + v.z * n.y,
                      v.x * t.z + v.y * b.z + v.z * n.z);
    }
    void flip_normal() { n = -n; } // :122-125
    // frame.slang:129-137
    Frame apply_normal_map(float3 n_local) const {
        const float3 new_N = to_world(n_local);
        const float3 old_T = t;
        const float3 new_T = normalize(old_T - new_N * dot(old_T, new_N));
        const float3 new_B = cross(new_N, new_T);
        Frame f;
        f.t = new_T; f.b = new_B; f.n = new_N;
        return f;
    }
};

// ---------------------------------------------------------------- encodings
// GLSL unpackSnorm2x16: clamp(int16 / 32767, -1, 1); packSnorm2x16: round(clamp(v,-1,1) * 32767)
inline float2 unpack_snorm2x16(uint32_t e) {
    const int16_t a = int16_t(e & 0xFFFFu), b = int16_t(e >> 16);
    return float2(clampf(float(a) / 32767.0f, -1.f, 1.f), clampf(float(b) / 32767.0f, -1.f, 1.f));
}
inline uint32_t pack_snorm2x16(float2 v) {
    const int a = int(std::nearbyint(clampf(v.x, -1.f, 1.f) * 32767.0f));
    const int b = int(std::nearbyint(clampf(v.y, -1.f, 1.f) * 32767.0f));
    return (uint32_t(a) & 0xFFFFu) | (uint32_t(b) << 16);
}
// include/merian-shaders/utils/encoding.slang:6-13
inline float3 decode_normal(uint32_t enc) {
    const float2 p = unpack_snorm2x16(enc);
    float3 vec(p.x, p.y, (1.0f - std::fabs(p.x)) - std::fabs(p.y));
    if (vec.z < 0.0f) {
        const float nx = std::copysign(1.0f - std::fabs(vec.y), vec.x);
        const float ny = std::copysign(1.0f - std::fabs(vec.x), vec.y);
        vec.x = nx;
        vec.y = ny;
    }
    return normalize(vec);
}
// encoding.slang:17-23
inline float2 octahedral_project(float3 n) {
    const float inv = 1.0f / ((std::fabs(n.x) + std::fabs(n.y)) + std::fabs(n.z));
    if (n.z < 0.0f) {
        return float2(std::copysign(1.0f - std::fabs(n.y * inv), n.x),
                      std::copysign(1.0f - std::fabs(n.x * inv), n.y));
    }
    return float2(n.x * inv, n.y * inv);
}
inline uint32_t encode_normal(float3 n) { return pack_snorm2x16(octahedral_project(n)); } // :28-30
// encoding.slang:46-55
inline uint32_t encode_tangent(float4 t) {
    const uint32_t dir = encode_normal(normalize(t.xyz())) & ~1u;
    return dir | (t.w < 0.f ? 1u : 0u);
}
inline float4 decode_tangent(uint32_t enc) {
    const float w = (enc & 1u) != 0u ? -1.f : 1.f;
    return float4(decode_normal(enc & ~1u), w);
}

// include/merian-shaders/utils/normal-map.slang:5-11
inline float3 keller_clamp_normal(float3 shading_normal, float3 face_normal, float3 wi) {
    const float3 r = reflect(-wi, shading_normal);
    if (dot(r, face_normal) < 0) {
        return normalize(wi + normalize(r - face_normal * dot(face_normal, r)));
    }
    return shading_normal;
}

// include/merian-shaders/utils/math.slang:243-249
inline float2 world_to_latlong_map(float3 dir) {
    const float3 p = normalize(dir);
    float2 uv;
    uv.x = std::atan2(p.x, -p.z) * INV_TWO_PI + 0.5f;
    uv.y = std::acos(p.y) * INV_PI;
    return uv;
}

// include/merian-shaders/colors/colorspace_yuv.slang:5-7
inline float yuv_luminance(float3 rgb) { return dot(rgb, float3(0.299f, 0.587f, 0.114f)); }

// ---------------------------------------------------------------- Fresnel
// include/merian-shaders/utils/fresnel.slang:14-26
inline float fresnel_schlick(float cos_theta, float F0, float F90 = 1.0f) {
    return F0 + (F90 - F0) * std::pow(std::max(1.0f - cos_theta, 0.f), 5.0f);
}
inline float3 fresnel_schlick(float cos_theta, float3 F0) {
    const float p = std::pow(std::max(1.0f - cos_theta, 0.f), 5.0f);
    return F0 + (float3(1.0f) - F0) * p;
}
// fresnel.slang:33-42
inline float fresnel_dielectric(float cos_i, float eta) {
    const float sin2_t = eta * eta * (1.0f - cos_i * cos_i);
    if (sin2_t >= 1.0f) return 1.0f;
    const float cos_t = std::sqrt(1.0f - sin2_t);
    const float r_parl = (eta * cos_i - cos_t) / (eta * cos_i + cos_t);
    const float r_perp = (cos_i - eta * cos_t) / (cos_i + eta * cos_t);
    return 0.5f * (r_parl * r_parl + r_perp * r_perp);
}
// fresnel.slang:47-59
inline float iridescence_ior_to_f0(float t_ior, float i_ior) { return pow2((t_ior - i_ior) / (t_ior + i_ior)); }
inline float3 iridescence_ior_to_f0(float3 t_ior, float i_ior) {
    const float3 r = (t_ior - i_ior) / (t_ior + i_ior);
    return r * r;
}
inline float3 iridescence_f0_to_ior(float3 f0) {
    const float3 s = sqrt3(f0);
    return (1.0f + s) / (1.0f - s);
}
// fresnel.slang:62-75
inline float3 iridescence_sensitivity(float opd, float3 shift) {
    const float phase = 2.0f * PI * opd * 1.0e-9f;
    const float3 val(5.4856e-13f, 4.4201e-13f, 5.2481e-13f);
    const float3 pos(1.6810e+06f, 1.7953e+06f, 2.2084e+06f);
    const float3 var(4.3278e+09f, 9.3046e+09f, 6.6121e+09f);
    float3 xyz = val * sqrt3(2.0f * PI * var) * cos3(pos * phase + shift) * exp3(-phase * phase * var);
    xyz.x += 9.7470e-14f * std::sqrt(2.0f * PI * 4.5282e+09f) * std::cos(2.2399e+06f * phase + shift.x) *
             std::exp(-4.5282e+09f * phase * phase);
    xyz /= 1.0685e-7f;
    return float3(3.2404542f * xyz.x - 1.5371385f * xyz.y - 0.4985314f * xyz.z,
                  -0.9692660f * xyz.x + 1.8760108f * xyz.y + 0.0415560f * xyz.z,
                  0.0556434f * xyz.x - 0.2040259f * xyz.y + 1.0572252f * xyz.z);
}
// fresnel.slang:79-129
inline float3 fresnel_iridescent(float outside_ior, float film_ior, float3 base_f0, float thickness,
                                 float cos_theta1) {
    const float iri_ior = lerp(outside_ior, film_ior, smoothstep(0.0f, 0.03f, thickness));
    const float sin_theta2_sq = pow2(outside_ior / iri_ior) * (1.0f - pow2(cos_theta1));
    const float cos_theta2_sq = 1.0f - sin_theta2_sq;
    if (!(cos_theta2_sq > 0.0f)) return float3(1.0f);
    const float cos_theta2 = std::sqrt(cos_theta2_sq);

    const float r12 = fresnel_schlick(cos_theta1, iridescence_ior_to_f0(iri_ior, outside_ior));
    const float t121 = 1.0f - r12;
    const float phi12 = iri_ior < outside_ior ? PI : 0.0f;
    const float phi21 = PI - phi12;

    const float3 base_ior = iridescence_f0_to_ior(clamp3(base_f0, 0.0f, 0.9999f));
    const float3 r23 = fresnel_schlick(cos_theta2, iridescence_ior_to_f0(base_ior, iri_ior));
    const float3 phi23(base_ior.x < iri_ior ? PI : 0.0f, base_ior.y < iri_ior ? PI : 0.0f,
                       base_ior.z < iri_ior ? PI : 0.0f);

    const float opd = 2.0f * iri_ior * thickness * cos_theta2;
    const float3 phi = float3(phi21) + phi23;

    const float3 r123 = clamp3(r12 * r23, 1e-5f, 0.9999f);
    const float3 r123_sqrt = sqrt3(r123);
    const float3 rs = pow2(t121) * r23 / (float3(1.0f) - r123);

    float3 result = float3(r12) + rs;
    float3 cm = rs - float3(t121);
    for (int m = 1; m <= 2; ++m) {
        cm *= r123_sqrt;
        const float3 sm = 2.0f * iridescence_sensitivity(float(m) * opd, float(m) * phi);
        result += cm * sm;
    }
    return clamp3(result, 0.0f, 1.0f);
}

} // namespace orc
