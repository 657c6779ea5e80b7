// vecmath.cuh — device-side vector helpers, RNG, encodings, frames (product code; sm_100a).
// Semantics follow the reference's Slang utility modules; each block cites the file:line it replaces.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define MPT_DEV __device__ __forceinline__

namespace mpt {

constexpr float kPi = 3.14159265358979323846f;
constexpr float kTwoPi = 6.28318530717958647692f;
constexpr float kInvPi = 0.31830988618379067154f;
constexpr float kInvTwoPi = 0.15915494309189533577f;

struct f2 { float x, y; };
struct f3 { float x, y, z; };
struct f4 { float x, y, z, w; };

MPT_DEV f2 mk2(float x, float y) { f2 r; r.x = x; r.y = y; return r; }
MPT_DEV f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
MPT_DEV f3 mk3(float s) { return mk3(s, s, s); }
MPT_DEV f4 mk4(float x, float y, float z, float w) { f4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
MPT_DEV f3 xyz(float4 v) { return mk3(v.x, v.y, v.z); }

MPT_DEV f2 operator+(f2 a, f2 b) { return mk2(a.x + b.x, a.y + b.y); }
MPT_DEV f2 operator-(f2 a, f2 b) { return mk2(a.x - b.x, a.y - b.y); }
MPT_DEV f2 operator*(float s, f2 a) { return mk2(a.x * s, a.y * s); }
MPT_DEV float dot(f2 a, f2 b) { return a.x * b.x + a.y * b.y; }

MPT_DEV f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
MPT_DEV f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
MPT_DEV f3 operator-(f3 a) { return mk3(-a.x, -a.y, -a.z); }
MPT_DEV f3 operator*(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
MPT_DEV f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
MPT_DEV f3 operator*(float s, f3 a) { return mk3(a.x * s, a.y * s, a.z * s); }
MPT_DEV f3 operator/(f3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
MPT_DEV f3 operator/(f3 a, f3 b) { return mk3(a.x / b.x, a.y / b.y, a.z / b.z); }
MPT_DEV f3 operator+(f3 a, float s) { return mk3(a.x + s, a.y + s, a.z + s); }
MPT_DEV f3 operator+(float s, f3 a) { return mk3(a.x + s, a.y + s, a.z + s); }
MPT_DEV f3 operator-(float s, f3 a) { return mk3(s - a.x, s - a.y, s - a.z); }
MPT_DEV f3 operator-(f3 a, float s) { return mk3(a.x - s, a.y - s, a.z - s); }
MPT_DEV f3& operator+=(f3& a, f3 b) { a = a + b; return a; }
MPT_DEV f3& operator*=(f3& a, f3 b) { a = a * b; return a; }
MPT_DEV f3& operator*=(f3& a, float s) { a = a * s; return a; }

MPT_DEV float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
MPT_DEV f3 cross(f3 a, f3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
MPT_DEV float length(f3 a) { return sqrtf(dot(a, a)); }
MPT_DEV f3 normalize(f3 a) { return a / length(a); }
MPT_DEV float max3(f3 v) { return fmaxf(v.x, fmaxf(v.y, v.z)); }
MPT_DEV f3 vmax(f3 a, float b) { return mk3(fmaxf(a.x, b), fmaxf(a.y, b), fmaxf(a.z, b)); }
MPT_DEV f3 vsqrt(f3 a) { return mk3(sqrtf(a.x), sqrtf(a.y), sqrtf(a.z)); }
MPT_DEV f3 vexp(f3 a) { return mk3(expf(a.x), expf(a.y), expf(a.z)); }
MPT_DEV f3 vcos(f3 a) { return mk3(cosf(a.x), cosf(a.y), cosf(a.z)); }
MPT_DEV float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
MPT_DEV f3 vclamp(f3 a, float lo, float hi) { return mk3(clampf(a.x, lo, hi), clampf(a.y, lo, hi), clampf(a.z, lo, hi)); }
MPT_DEV float sq(float x) { return x * x; }
// HLSL lerp: x*(1-s) + y*s
MPT_DEV float mix(float a, float b, float s) { return a * (1.0f - s) + b * s; }
MPT_DEV f3 mix(f3 a, f3 b, float s) { return a * (1.0f - s) + b * s; }
MPT_DEV float sgn(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }
MPT_DEV bool all_below(f3 a, float b) { return a.x < b && a.y < b && a.z < b; }
MPT_DEV bool finite3(f3 a) { return isfinite(a.x) && isfinite(a.y) && isfinite(a.z); }
// Component i of v as two SELs. Written in PTX because the ternary form `i == 0 ? v.x : (i == 1 ? v.y : v.z)` came out of nvcc 12.9 as a
// divergent branch per call (BSSY / BRA / BSYNC, 7-8 instructions each): the nine selections of the watertight triangle test were 72 of
// its 135 SASS instructions (cuobjdump of k_extend_pooled, round 2 session M).
MPT_DEV float comp(f3 v, int i) {
#if defined(__CUDA_ARCH__) && !defined(MPT_COMP_TERNARY)
    float r;
    asm("{\n\t.reg .pred p0, p1;\n\t.reg .f32 t;\n\tsetp.eq.s32 p0, %4, 0;\n\tsetp.eq.s32 p1, %4, 1;\n\tselp.f32 t, %2, %3, p1;\n\tselp.f32 %0, %1, t, p0;\n\t}"
        : "=f"(r)
        : "f"(v.x), "f"(v.y), "f"(v.z), "r"(i));
    return r;
#else
    return i == 0 ? v.x : (i == 1 ? v.y : v.z);
#endif
}

// ---- hash.slang:138-146
MPT_DEV uint32_t pcg4d16(uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    x = x * 1664525u + 1013904223u;
    y = y * 1664525u + 1013904223u;
    z = z * 1664525u + 1013904223u;
    w = w * 1664525u + 1013904223u;
    x += y * w;
    y += z * x;
    z += x * y;
    w += y * z;
    x += y * w;
    return x;
}

// ---- random.slang:108-128 (PCG) with XorShift32 (:64-82) selectable for the BSDF property suite;
//      float conversion bit-twiddling.slang:23-27
struct Rng {
    uint32_t state;
    bool xorshift;
    MPT_DEV uint32_t next() {
        if (xorshift) {
            uint32_t v = state;
            v ^= v << 13;
            v ^= v >> 17;
            v ^= v << 5;
            state = v;
            return v;
        }
        const uint32_t s = state;
        state = s * 747796405u + 2891336453u;
        const uint32_t word = ((s >> ((s >> 28u) + 4u)) ^ s) * 277803737u;
        return (word >> 22u) ^ word;
    }
    MPT_DEV float next_float() { return float(next() >> 8) * 5.9604644775390625e-08f; }
};
MPT_DEV Rng make_pcg(uint32_t seed) { Rng r; r.state = seed; r.xorshift = false; return r; }

// ---- frame.slang:7-13, 88-137: rows t, b, n
struct Frame {
    f3 t, b, n;
    MPT_DEV f3 to_local(f3 v) const { return mk3(dot(t, v), dot(b, v), dot(n, v)); }
    MPT_DEV f3 to_world(f3 v) const {
        return mk3(v.x * t.x + v.y * b.x + v.z * n.x, v.x * t.y + v.y * b.y + v.z * n.y, v.x * t.z + v.y * b.z + v.z * n.z);
    }
};
MPT_DEV Frame frame_from_normal(f3 z) { // Duff et al. 2017
    const float s = (z.z >= 0.f) ? 1.f : -1.f;
    const float a = -1.0f / (s + z.z);
    const float b = z.x * z.y * a;
    Frame f;
    f.t = mk3(1.0f + s * z.x * z.x * a, s * b, -s * z.x);
    f.b = mk3(b, s + z.y * z.y * a, -z.y);
    f.n = z;
    return f;
}
MPT_DEV Frame frame_from_tangent(f3 n, f3 target_tangent, float handedness) {
    Frame f;
    f.n = normalize(n);
    f.t = normalize(target_tangent - f.n * dot(target_tangent, f.n));
    f.b = cross(f.n, f.t) * handedness;
    return f;
}
MPT_DEV Frame frame_apply_normal_map(const Frame& fr, f3 n_local) {
    Frame f;
    f.n = fr.to_world(n_local);
    f.t = normalize(fr.t - f.n * dot(fr.t, f.n));
    f.b = cross(f.n, f.t);
    return f;
}

// ---- encoding.slang:6-55 (oct-encoded unit vectors, snorm16 x2)
MPT_DEV f3 decode_normal(uint32_t enc) {
    const float px = clampf(float(int16_t(enc & 0xFFFFu)) / 32767.0f, -1.f, 1.f);
    const float py = clampf(float(int16_t(enc >> 16)) / 32767.0f, -1.f, 1.f);
    f3 v = mk3(px, py, (1.0f - fabsf(px)) - fabsf(py));
    if (v.z < 0.0f) {
        const float nx = copysignf(1.0f - fabsf(v.y), v.x);
        const float ny = copysignf(1.0f - fabsf(v.x), v.y);
        v.x = nx;
        v.y = ny;
    }
    return normalize(v);
}
MPT_DEV uint32_t encode_normal(f3 n) {
    const float inv = 1.0f / ((fabsf(n.x) + fabsf(n.y)) + fabsf(n.z));
    float px, py;
    if (n.z < 0.0f) {
        px = copysignf(1.0f - fabsf(n.y * inv), n.x);
        py = copysignf(1.0f - fabsf(n.x * inv), n.y);
    } else {
        px = n.x * inv;
        py = n.y * inv;
    }
    const int a = __float2int_rn(clampf(px, -1.f, 1.f) * 32767.0f);
    const int b = __float2int_rn(clampf(py, -1.f, 1.f) * 32767.0f);
    return (uint32_t(a) & 0xFFFFu) | (uint32_t(b) << 16);
}
MPT_DEV f4 decode_tangent(uint32_t enc) {
    const f3 d = decode_normal(enc & ~1u);
    return mk4(d.x, d.y, d.z, (enc & 1u) ? -1.f : 1.f);
}

MPT_DEV f2 half2_bits_to_f2(uint32_t bits) {
    const __half2 h = *reinterpret_cast<const __half2*>(&bits);
    const float2 f = __half22float2(h);
    return mk2(f.x, f.y);
}

// ---- hit.slang:14-23 (Waechter & Binder)
MPT_DEV float offset_component(float p, float n) {
    const int o = int(n * 256.f);
    const float ip = __int_as_float(__float_as_int(p) + (p < 0.f ? -o : o));
    return fabsf(p) < (1.f / 16.f) ? p + n * (1.f / 65536.f) : ip;
}
MPT_DEV f3 offset_ray_origin(f3 pos, f3 n) {
    return mk3(offset_component(pos.x, n.x), offset_component(pos.y, n.y), offset_component(pos.z, n.z));
}

// ---- colorspace_yuv.slang:5-7
MPT_DEV float yuv_luminance(f3 c) { return dot(c, mk3(0.299f, 0.587f, 0.114f)); }

// ---- math.slang:243-249
MPT_DEV f2 world_to_latlong(f3 dir) {
    const f3 p = normalize(dir);
    return mk2(atan2f(p.x, -p.z) * kInvTwoPi + 0.5f, acosf(p.y) * kInvPi);
}

} // namespace mpt
