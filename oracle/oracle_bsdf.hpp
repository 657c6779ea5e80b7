// oracle_bsdf.hpp — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the BSDF layer reachable from merian::GltfBSDF, kept as
// the same generic nest the Slang source declares (gltf-material.slang:18-27) so the evaluation order,
// RNG consumption and the one-sample-MIS pdf/weight arithmetic follow the reference statement by
// statement. The product (merian_b200/csrc/bsdf.cuh) flattens and specialises this nest; the two are
// written independently and compared by tests/.
//
// Pinned by the reference's own property suite (tests/test_bsdf.cpp, tests/bsdf/bsdf-checks.slang),
// restated in tests/test_bsdf_oracle.py.
#pragma once

#include "oracle_math.hpp"

namespace orc {

// include/merian-shaders/shading/bsdfs/bsdf.slang:13
constexpr float BSDF_MIN_COS_THETA = 1e-6f;

// include/merian-shaders/shading/shading-function.slang:5-33
enum class Lobe { Delta, Glossy, Diffuse };
struct BsdfSample {
    float3 wo;
    float pdf;
    float3 weight;
    Lobe lobe = Lobe::Glossy;
    BsdfSample(float3 wo_, float pdf_, float3 w_, Lobe l = Lobe::Glossy) : wo(wo_), pdf(pdf_), weight(w_), lobe(l) {}
};
using OptSample = std::optional<BsdfSample>;

// shading-function.slang:76-85
template <class T> OptSample sample_eval_impl(const T& sf, float3 wi, Rng& rng) {
    const std::optional<float3> wo = sf.sample(wi, rng);
    if (wo) {
        const float p = sf.pdf(wi, *wo);
        return BsdfSample(*wo, p, sf.eval(wi, *wo, rng) / p);
    }
    return std::nullopt;
}

// ------------------------------------------------------------------ GGX helpers (brdf-ggx.slang)
// brdf-ggx.slang:35-44
inline float smith_g1_over_widotn(float widotn, float alpha_squared) {
    return 2.0f / (widotn + std::sqrt(alpha_squared + ((1.0f - alpha_squared) * (widotn * widotn))));
}
// brdf-ggx.slang:51-55
inline float smith_g2_over_widotn(float wodotn, float widotn, float alpha_squared) {
    const float g1 = widotn * std::sqrt(alpha_squared + ((1.0f - alpha_squared) * (wodotn * wodotn)));
    const float g2 = wodotn * std::sqrt(alpha_squared + ((1.0f - alpha_squared) * (widotn * widotn)));
    return (2.0f * wodotn) / (g1 + g2);
}
// brdf-ggx.slang:71-73 (per-axis alpha)
inline float smith_g1_over_widotn(float3 wi, float2 alpha) {
    return 2.0f / (wi.z + length(float3(alpha.x * wi.x, alpha.y * wi.y, wi.z)));
}
// brdf-ggx.slang:78-82
inline float smith_g2_over_widotn(float3 wo, float3 wi, float2 alpha) {
    const float g1 = wi.z * length(float3(alpha.x * wo.x, alpha.y * wo.y, wo.z));
    const float g2 = wo.z * length(float3(alpha.x * wi.x, alpha.y * wi.y, wi.z));
    return (2.0f * wo.z) / (g1 + g2);
}
// brdf-ggx.slang:88-91
inline float ggx_NDF(float ndoth_squared, float alpha_squared) {
    return alpha_squared / (PI * pow2((ndoth_squared * alpha_squared) + (1.0f - ndoth_squared)));
}
// brdf-ggx.slang:95-100
inline float ggx_NDF(float3 h, float2 alpha) {
    const float alpha_xy = alpha.x * alpha.y;
    const float3 v(alpha.y * h.x, alpha.x * h.y, alpha_xy * h.z);
    const float w2 = alpha_xy / dot(v, v);
    return alpha_xy * w2 * w2 / PI;
}
// brdf-ggx.slang:137-152 (Dupuy & Benyoub 2023, spherical caps)
inline float3 ggx_VNDF_sample_H(float3 wi, float2 alpha, float2 random) {
    const float3 wiStd = normalize(float3(wi.x * alpha.x, wi.y * alpha.y, wi.z));
    const float phi = ((2.0f * random.x) - 1.0f) * PI;
    const float z = std::fma(1.0f - random.y, 1.0f + wiStd.z, -wiStd.z); // mad()
    const float sinTheta = std::sqrt(clampf(1.0f - (z * z), 0.0f, 1.0f));
    const float x = sinTheta * std::cos(phi);
    const float y = sinTheta * std::sin(phi);
    const float3 c(x, y, z);
    const float3 wmStd = c + wiStd;
    return normalize(float3(wmStd.x * alpha.x, wmStd.y * alpha.y, wmStd.z));
}
// brdf-ggx.slang:234-239
inline float ggx_VNDF_H_pdf(float3 wi, float3 h, float alpha_squared) {
    const float G1 = smith_g1_over_widotn(wi.z, alpha_squared);
    const float D = ggx_NDF(pow2(h.z), alpha_squared);
    return D * G1 * std::max(0.f, dot(wi, h));
}
// brdf-ggx.slang:246-251
inline float ggx_VNDF_pdf_wo(float widotn, float ndoth_squared, float alpha_squared) {
    const float G1 = smith_g1_over_widotn(widotn, alpha_squared);
    const float D = ggx_NDF(ndoth_squared, alpha_squared);
    return (D * G1) / 4.0f;
}
// brdf-ggx.slang:276-278
inline float ggx_VNDF_pdf_wo(float3 wi, float3 h, float2 alpha) {
    return (ggx_NDF(h, alpha) * smith_g1_over_widotn(wi, alpha)) / 4.0f;
}

// brdf-ggx.slang:285-318 (RT Gems ch. 32 fit; height-correlated variant).
// HLSL matrices are row-major constructors; mul(M, v) = rows dotted with v.
inline float3 ggx_integral_approx_correlated(float3 F0, float alpha, float cosTheta) {
    const float c = std::fabs(cosTheta);
    const float X[4] = {1.0f, c, c * c, c * c * c};
    const float Y[4] = {1.0f, alpha, alpha * alpha, alpha * alpha * alpha};
    // M1 (2x2) * X.xy . Y.xy
    const float m1x = 0.995367f * X[0] + -1.38839f * X[1];
    const float m1y = -0.24751f * X[0] + 1.97442f * X[1];
    // M2 (3x3) * X.xyw . Y.xyw
    const float m2x = 1.0f * X[0] + 2.68132f * X[1] + 52.366f * X[3];
    const float m2y = 16.0932f * X[0] + -3.98452f * X[1] + 59.3013f * X[3];
    const float m2z = -5.18731f * X[0] + 255.259f * X[1] + 2544.07f * X[3];
    float bias = (m1x * Y[0] + m1y * Y[1]) / (m2x * Y[0] + m2y * Y[1] + m2z * Y[3]);
    // M3 (2x2) * X.xy . Y.xy
    const float m3x = -0.0564526f * X[0] + 3.82901f * X[1];
    const float m3y = 16.91f * X[0] + -11.0303f * X[1];
    // M4 (3x3) * X.xzw . Y.xyw
    const float m4x = 1.0f * X[0] + 4.11118f * X[2] + -1.37886f * X[3];
    const float m4y = 19.3254f * X[0] + -28.9947f * X[2] + 16.9514f * X[3];
    const float m4z = 0.545386f * X[0] + 96.0994f * X[2] + -79.4492f * X[3];
    const float scale = (m3x * Y[0] + m3y * Y[1]) / (m4x * Y[0] + m4y * Y[1] + m4z * Y[3]);
    const float F0_luma = (F0.x + F0.y + F0.z) * (1.0f / 3.0f);
    bias *= saturate(F0_luma * 50.0f);
    return F0 * std::max(0.0f, scale) + std::max(0.0f, bias);
}

// ------------------------------------------------------------------ Lambert (brdf-lambert-diffuse.slang:9-61)
struct LambertDiffuseBRDF {
    float3 albedo;
    explicit LambertDiffuseBRDF(float3 a) : albedo(a) {}
    float3 eval(float3 wi, float3 wo, Rng&) const {
        if (std::min(wi.z, wo.z) < BSDF_MIN_COS_THETA) return float3(0.f);
        return INV_PI * albedo * wo.z;
    }
    std::optional<float3> sample(float3 wi, Rng& rng) const {
        const float3 wo = sample_cos(rng);
        if (std::min(wi.z, wo.z) < BSDF_MIN_COS_THETA) return std::nullopt;
        return wo;
    }
    OptSample sample_eval(float3 wi, Rng& rng) const {
        const float3 wo = sample_cos(rng);
        if (std::min(wi.z, wo.z) < BSDF_MIN_COS_THETA) return std::nullopt;
        return BsdfSample(wo, INV_PI * wo.z, albedo, Lobe::Diffuse);
    }
    float pdf(float3 wi, float3 wo) const {
        if (std::min(wi.z, wo.z) < BSDF_MIN_COS_THETA) return 0.0f;
        return INV_PI * wo.z;
    }
    float3 get_albedo(float3) const { return albedo; }
    float3 get_preintegrated_specular(float3, float3) const { return float3(0.f); }
    float alpha() const { return 1.0f; }
};

// ------------------------------------------------------------------ GGX BRDF (brdf-ggx.slang:331-412)
struct GGXBRDF {
    float2 alpha_xy;
    explicit GGXBRDF(float a) : alpha_xy(a, a) {}
    explicit GGXBRDF(float2 a) : alpha_xy(a) {}
    float3 eval(float3 wi, float3 wo, Rng&) const {
        if (std::min(wi.z, wo.z) < BSDF_MIN_COS_THETA) return float3(0.f);
        const float3 h = normalize(wo + wi);
        const float G2 = smith_g2_over_widotn(wo, wi, alpha_xy);
        const float D = ggx_NDF(h, alpha_xy);
        return float3((D * G2) / 4.0f);
    }
    std::optional<float3> sample(float3 wi, Rng& rng) const {
        if (wi.z < BSDF_MIN_COS_THETA) return std::nullopt;
        const float3 h = ggx_VNDF_sample_H(wi, alpha_xy, rng.next_float2());
        const float widoth = dot(wi, h);
        const float3 wo = reflect2(wi, h, widoth);
        if (wo.z < BSDF_MIN_COS_THETA) return std::nullopt;
        return wo;
    }
    float pdf(float3 wi, float3 wo) const {
        if (std::min(wi.z, wo.z) < BSDF_MIN_COS_THETA) return 0.0f;
        const float3 h = normalize(wi + wo);
        return ggx_VNDF_pdf_wo(wi, h, alpha_xy);
    }
    OptSample sample_eval(float3 wi, Rng& rng) const {
        if (wi.z < BSDF_MIN_COS_THETA) return std::nullopt;
        const float3 h = ggx_VNDF_sample_H(wi, alpha_xy, rng.next_float2());
        const float3 wo = reflect2(wi, h, dot(wi, h));
        if (wo.z < BSDF_MIN_COS_THETA) return std::nullopt;
        const float pdf = ggx_VNDF_pdf_wo(wi, h, alpha_xy);
        const float weight = smith_g2_over_widotn(wo, wi, alpha_xy) / smith_g1_over_widotn(wi, alpha_xy);
        return BsdfSample(wo, pdf, float3(weight), Lobe::Glossy);
    }
    float3 get_preintegrated_specular(float3 F0, float3 wi) const {
        return ggx_integral_approx_correlated(F0, std::sqrt(alpha_xy.x * alpha_xy.y), wi.z);
    }
    float3 get_albedo(float3 wi) const { return get_preintegrated_specular(float3(1.0f), wi); }
    float alpha() const { return std::min(alpha_xy.x, alpha_xy.y); }
};

// ------------------------------------------------------------------ GGX BTDF (brdf-ggx.slang:416-502)
struct GGXBTDF {
    float alpha_;
    float eta;
    float3 tint;
    GGXBTDF(float a, float e, float3 t) : alpha_(a), eta(e), tint(t) {}
    float3 eval(float3 wi, float3 wo, Rng&) const {
        if (wi.z < BSDF_MIN_COS_THETA || wo.z > -BSDF_MIN_COS_THETA) return float3(0.f);
        float3 h = normalize(wo + (wi * eta));
        h *= signf(h.z);
        const float widoth = dot(wi, h);
        const float wodoth = dot(wo, h);
        const float a2 = pow2(alpha_);
        const float G2 = smith_g2_over_widotn(std::fabs(wo.z), wi.z, a2);
        const float D = ggx_NDF(pow2(h.z), a2);
        const float denom = wodoth + (eta * widoth);
        const float jacobian = pow2(eta) * widoth * wodoth / pow2(denom);
        return tint * (D * G2 * std::fabs(jacobian));
    }
    float pdf(float3 wi, float3 wo) const {
        if (wi.z < BSDF_MIN_COS_THETA || wo.z > -BSDF_MIN_COS_THETA) return 0.0f;
        float3 h = normalize(wo + (wi * eta));
        h *= signf(h.z);
        const float wodoth = dot(wo, h);
        const float denom = wodoth + (eta * dot(wi, h));
        return ggx_VNDF_H_pdf(wi, h, pow2(alpha_)) * std::fabs(wodoth) / pow2(denom);
    }
};

// ------------------------------------------------------------------ Mix (bsdf-mix.slang:6-65)
template <class B0, class B1> struct MixBSDF {
    B0 bsdf0;
    B1 bsdf1;
    float weight;
    MixBSDF(B0 b0, B1 b1, float w) : bsdf0(b0), bsdf1(b1), weight(w) {}
    float3 eval(float3 wi, float3 wo, Rng& rng) const {
        return lerp(bsdf0.eval(wi, wo, rng), bsdf1.eval(wi, wo, rng), weight);
    }
    std::optional<float3> sample(float3 wi, Rng& rng) const {
        if (weight > 0.0f && rng.next_float() < weight) return bsdf1.sample(wi, rng);
        return bsdf0.sample(wi, rng);
    }
    float pdf(float3 wi, float3 wo) const { return lerp(bsdf0.pdf(wi, wo), bsdf1.pdf(wi, wo), weight); }
    OptSample sample_eval(float3 wi, Rng& rng) const {
        if (weight > 0.0f && rng.next_float() < weight) {
            if (auto s = bsdf1.sample_eval(wi, rng)) {
                const float mix_pdf = lerp(bsdf0.pdf(wi, s->wo), s->pdf, weight);
                const float3 mix_eval = lerp(bsdf0.eval(wi, s->wo, rng), s->weight * s->pdf, weight);
                return BsdfSample(s->wo, mix_pdf, mix_eval / mix_pdf, s->lobe);
            }
            return std::nullopt;
        }
        if (auto s = bsdf0.sample_eval(wi, rng)) {
            const float mix_pdf = lerp(s->pdf, bsdf1.pdf(wi, s->wo), weight);
            const float3 mix_eval = lerp(s->weight * s->pdf, bsdf1.eval(wi, s->wo, rng), weight);
            return BsdfSample(s->wo, mix_pdf, mix_eval / mix_pdf, s->lobe);
        }
        return std::nullopt;
    }
    float3 get_albedo(float3 wi) const { return lerp(bsdf0.get_albedo(wi), bsdf1.get_albedo(wi), weight); }
    float3 get_preintegrated_specular(float3 F0, float3 wi) const {
        return lerp(bsdf0.get_preintegrated_specular(F0, wi), bsdf1.get_preintegrated_specular(F0, wi), weight);
    }
    float alpha() const { return std::sqrt(lerp(pow2(bsdf0.alpha()), pow2(bsdf1.alpha()), weight)); }
};

// ------------------------------------------------------------------ FresnelMix (bsdf-fresnel-mix.slang:7-63)
template <class Layer, class Base> struct FresnelMixBSDF {
    Layer layer;
    Base base;
    float ior_f0;
    FresnelMixBSDF(Layer l, Base b, float f0) : layer(l), base(b), ior_f0(f0) {}
    float3 eval(float3 wi, float3 wo, Rng& rng) const {
        const float3 h = normalize(wi + wo);
        const float F = fresnel_schlick(dot(wi, h), ior_f0);
        return lerp(base.eval(wi, wo, rng), layer.eval(wi, wo, rng), F);
    }
    std::optional<float3> sample(float3 wi, Rng& rng) const {
        const float F = fresnel_schlick(wi.z, ior_f0);
        if (rng.next_float() < F) return layer.sample(wi, rng);
        return base.sample(wi, rng);
    }
    float pdf(float3 wi, float3 wo) const {
        const float F = fresnel_schlick(wi.z, ior_f0);
        return lerp(base.pdf(wi, wo), layer.pdf(wi, wo), F);
    }
    OptSample sample_eval(float3 wi, Rng& rng) const {
        const float F = fresnel_schlick(wi.z, ior_f0);
        OptSample sub = rng.next_float() < F ? layer.sample_eval(wi, rng) : base.sample_eval(wi, rng);
        if (sub) {
            const float p = this->pdf(wi, sub->wo);
            return BsdfSample(sub->wo, p, this->eval(wi, sub->wo, rng) / p, sub->lobe);
        }
        return std::nullopt;
    }
    float3 get_albedo(float3 wi) const {
        const float3 spec = layer.get_preintegrated_specular(float3(ior_f0), wi);
        return spec + (1.0f - spec) * base.get_albedo(wi);
    }
    float3 get_preintegrated_specular(float3 F0, float3 wi) const { return layer.get_preintegrated_specular(F0, wi); }
    float alpha() const { return std::sqrt(lerp(pow2(base.alpha()), pow2(layer.alpha()), ior_f0)); }
};

// ------------------------------------------------------------------ Conductor (bsdf-conductor-fresnel.slang:6-50)
template <class Inner> struct ConductorFresnelBSDF {
    Inner bsdf;
    float3 f0;
    ConductorFresnelBSDF(Inner b, float3 f) : bsdf(b), f0(f) {}
    float3 eval(float3 wi, float3 wo, Rng& rng) const {
        const float3 h = normalize(wi + wo);
        return fresnel_schlick(dot(wi, h), f0) * bsdf.eval(wi, wo, rng);
    }
    std::optional<float3> sample(float3 wi, Rng& rng) const { return bsdf.sample(wi, rng); }
    float pdf(float3 wi, float3 wo) const { return bsdf.pdf(wi, wo); }
    OptSample sample_eval(float3 wi, Rng& rng) const {
        if (auto s = bsdf.sample_eval(wi, rng)) {
            BsdfSample r = *s;
            const float3 h = normalize(wi + r.wo);
            r.weight *= fresnel_schlick(dot(wi, h), f0);
            return r;
        }
        return std::nullopt;
    }
    float3 get_albedo(float3 wi) const { return bsdf.get_preintegrated_specular(f0, wi); }
    float3 get_preintegrated_specular(float3 F0, float3 wi) const { return bsdf.get_preintegrated_specular(F0, wi); }
    float alpha() const { return bsdf.alpha(); }
};

// ------------------------------------------------------------------ Iridescent conductor (bsdf-iridescence.slang:11-77)
template <class Inner> struct IridescentConductorBSDF {
    Inner bsdf;
    float3 f0;
    float film_ior, thickness, iridescence;
    IridescentConductorBSDF(Inner b, float3 f, float fi, float th, float ir)
        : bsdf(b), f0(f), film_ior(fi), thickness(th), iridescence(ir) {}
    float3 fresnel(float vdoth) const {
        const float3 schlick = fresnel_schlick(vdoth, f0);
        if (iridescence <= 0.0f) return schlick;
        return lerp(schlick, fresnel_iridescent(1.0f, film_ior, f0, thickness, vdoth), iridescence);
    }
    float3 eval(float3 wi, float3 wo, Rng& rng) const {
        const float3 h = normalize(wi + wo);
        return fresnel(dot(wi, h)) * bsdf.eval(wi, wo, rng);
    }
    std::optional<float3> sample(float3 wi, Rng& rng) const { return bsdf.sample(wi, rng); }
    float pdf(float3 wi, float3 wo) const { return bsdf.pdf(wi, wo); }
    OptSample sample_eval(float3 wi, Rng& rng) const {
        if (auto s = bsdf.sample_eval(wi, rng)) {
            BsdfSample r = *s;
            const float3 h = normalize(wi + r.wo);
            r.weight *= fresnel(dot(wi, h));
            return r;
        }
        return std::nullopt;
    }
    float3 get_albedo(float3 wi) const {
        const float3 spec = bsdf.get_preintegrated_specular(f0, wi);
        if (iridescence <= 0.0f) return spec;
        return spec * fresnel(wi.z) / max3v(fresnel_schlick(wi.z, f0), 1e-4f);
    }
    float3 get_preintegrated_specular(float3 F0, float3 wi) const { return bsdf.get_preintegrated_specular(F0, wi); }
    float alpha() const { return bsdf.alpha(); }
};

// ------------------------------------------------------------------ Iridescent Fresnel mix (bsdf-iridescence.slang:82-178)
template <class Layer, class Base> struct IridescentFresnelMixBSDF {
    Layer layer;
    Base base;
    float ior_f0, eta, film_ior, thickness, iridescence;
    IridescentFresnelMixBSDF(Layer l, Base b, float f0, float e, float fi, float th, float ir)
        : layer(l), base(b), ior_f0(f0), eta(e), film_ior(fi), thickness(th), iridescence(ir) {}
    float3 eval(float3 wi, float3 wo, Rng& rng) const {
        const float3 base_e = base.eval(wi, wo, rng);
        const float3 layer_e = layer.eval(wi, wo, rng);
        const float widoth = std::fabs(dot(wi, normalize(wi + wo)));
        const float f = fresnel_dielectric(widoth, eta);
        const float3 dielectric = lerp(base_e, layer_e, f);
        if (iridescence <= 0.0f) return dielectric;
        const float3 irf = fresnel_iridescent(1.0f, film_ior, float3(ior_f0), thickness, widoth);
        const float3 iridescent = (1.0f - irf) * base_e + irf * layer_e;
        return lerp(dielectric, iridescent, iridescence);
    }
    std::optional<float3> sample(float3 wi, Rng& rng) const {
        const float f = fresnel_dielectric(wi.z, eta);
        if (rng.next_float() < f) return layer.sample(wi, rng);
        return base.sample(wi, rng);
    }
    float pdf(float3 wi, float3 wo) const {
        const float f = fresnel_dielectric(wi.z, eta);
        return lerp(base.pdf(wi, wo), layer.pdf(wi, wo), f);
    }
    OptSample sample_eval(float3 wi, Rng& rng) const {
        const float f = fresnel_dielectric(wi.z, eta);
        OptSample sub = rng.next_float() < f ? layer.sample_eval(wi, rng) : base.sample_eval(wi, rng);
        if (sub) {
            const float p = this->pdf(wi, sub->wo);
            return BsdfSample(sub->wo, p, this->eval(wi, sub->wo, rng) / p, sub->lobe);
        }
        return std::nullopt;
    }
    float3 get_albedo(float3 wi) const {
        float3 spec = layer.get_preintegrated_specular(float3(ior_f0), wi);
        if (iridescence > 0.0f) {
            const float3 irf = fresnel_iridescent(1.0f, film_ior, float3(ior_f0), thickness, wi.z);
            const float die = fresnel_dielectric(wi.z, eta);
            spec = clamp3(spec * lerp(float3(1.0f), irf / std::max(die, 1e-4f), iridescence), 0.0f, 1.0f);
        }
        return spec + (1.0f - spec) * base.get_albedo(wi);
    }
    float3 get_preintegrated_specular(float3 F0, float3 wi) const { return layer.get_preintegrated_specular(F0, wi); }
    float alpha() const { return std::sqrt(lerp(pow2(base.alpha()), pow2(layer.alpha()), ior_f0)); }
};

// ------------------------------------------------------------------ Iridescent rough dielectric (bsdf-iridescence.slang:184-323)
struct IridescentRoughDielectricBSDF {
    float alpha_, eta;
    float3 tint, base_f0;
    float film_ior, thickness, iridescence;
    IridescentRoughDielectricBSDF(float a, float e, float3 t, float3 bf0, float fi, float th, float ir)
        : alpha_(a), eta(e), tint(t), base_f0(bf0), film_ior(fi), thickness(th), iridescence(ir) {}
    float3 reflectance(float widoth) const {
        const float dielectric = fresnel_dielectric(widoth, eta);
        if (iridescence <= 0.0f) return float3(dielectric);
        return lerp(float3(dielectric), fresnel_iridescent(1.0f, film_ior, base_f0, thickness, widoth), iridescence);
    }
    float3 eval(float3 wi, float3 wo, Rng& rng) const {
        if (wi.z < BSDF_MIN_COS_THETA) return float3(0.f);
        if (wo.z > BSDF_MIN_COS_THETA) {
            const float3 h = normalize(wi + wo);
            return reflectance(std::fabs(dot(wi, h))) * GGXBRDF(alpha_).eval(wi, wo, rng);
        }
        if (wo.z < -BSDF_MIN_COS_THETA) {
            float3 h = normalize(wo + wi * eta);
            h *= signf(h.z);
            return (1.0f - reflectance(std::fabs(dot(wi, h)))) * GGXBTDF(alpha_, eta, tint).eval(wi, wo, rng);
        }
        return float3(0.f);
    }
    float pdf(float3 wi, float3 wo) const {
        if (wi.z < BSDF_MIN_COS_THETA) return 0.0f;
        const float a2 = pow2(alpha_);
        if (wo.z > BSDF_MIN_COS_THETA) {
            const float3 h = normalize(wi + wo);
            const float p = max3(reflectance(std::fabs(dot(wi, h))));
            return p * ggx_VNDF_pdf_wo(wi.z, pow2(std::max(0.0f, h.z)), a2);
        }
        if (wo.z < -BSDF_MIN_COS_THETA) {
            float3 h = normalize(wo + wi * eta);
            h *= signf(h.z);
            const float p = max3(reflectance(std::fabs(dot(wi, h))));
            const float wodoth = dot(wo, h);
            const float denom = wodoth + eta * dot(wi, h);
            return (1.0f - p) * ggx_VNDF_H_pdf(wi, h, a2) * std::fabs(wodoth) / pow2(denom);
        }
        return 0.0f;
    }
    OptSample sample_eval(float3 wi, Rng& rng) const {
        if (wi.z < BSDF_MIN_COS_THETA) return std::nullopt;
        const float a2 = pow2(alpha_);
        const float3 wm = ggx_VNDF_sample_H(wi, float2(alpha_), rng.next_float2());
        const float widoth = dot(wi, wm);
        const float3 R = reflectance(std::fabs(widoth));
        const float p = max3(R);
        if (rng.next_float() < p) {
            const float3 wo = reflect2(wi, wm, widoth);
            if (wo.z < BSDF_MIN_COS_THETA) return std::nullopt;
            const float g = smith_g2_over_widotn(wo.z, wi.z, a2) / smith_g1_over_widotn(wi.z, a2);
            const float pdf = p * ggx_VNDF_pdf_wo(wi.z, pow2(wm.z), a2);
            return BsdfSample(wo, pdf, (R / p) * g, Lobe::Glossy);
        }
        const float3 wo = refract(-wi, wm, eta);
        if ((wo.x == 0.f && wo.y == 0.f && wo.z == 0.f) || wo.z > -BSDF_MIN_COS_THETA) return std::nullopt;
        const float wodoth = dot(wo, wm);
        const float denom = wodoth + eta * widoth;
        const float pdf = (1.0f - p) * ggx_VNDF_H_pdf(wi, wm, a2) * std::fabs(wodoth) / pow2(denom);
        const float g = smith_g2_over_widotn(std::fabs(wo.z), wi.z, a2) / smith_g1_over_widotn(wi.z, a2);
        return BsdfSample(wo, pdf, ((1.0f - R) / (1.0f - p)) * pow2(eta) * tint * g, Lobe::Glossy);
    }
    std::optional<float3> sample(float3 wi, Rng& rng) const {
        if (auto s = sample_eval(wi, rng)) return s->wo;
        return std::nullopt;
    }
    float3 get_preintegrated_specular(float3 F0, float3 wi) const { return GGXBRDF(alpha_).get_preintegrated_specular(F0, wi); }
    float3 get_albedo(float3 wi) const {
        const float3 f0 = float3(pow2((1.0f - eta) / (1.0f + eta)));
        float3 spec = GGXBRDF(alpha_).get_preintegrated_specular(f0, wi);
        if (iridescence > 0.0f) {
            spec = clamp3(spec * reflectance(wi.z) / std::max(fresnel_dielectric(wi.z, eta), 1e-4f), 0.0f, 1.0f);
        }
        return spec + (1.0f - spec) * pow2(eta) * tint;
    }
    // The property is not declared in the Slang struct; GltfBSDF::alpha is never read on the pt path.
    float alpha() const { return alpha_; }
};

// ------------------------------------------------------------------ Sheen (brdf-sheen.slang:9-118)
// brdf-sheen.slang:94-118
inline float sheen_directional_albedo(float ndotv, float alpha) {
    static const float coeff[11][5] = {
        {1.01108f, 75.2286f, -9.4671f, 0.000973445f, -1.01325f},
        {8.44447f, 0.0135054f, 1.11637f, 0.479139f, -9.30873f},
        {5.67124f, 0.0194668f, 1.12489f, 0.453551f, -6.53104f},
        {4.50086f, 0.0237758f, 1.09376f, 0.423665f, -5.36521f},
        {3.77781f, 0.0273286f, 1.0666f, 0.400267f, -4.65601f},
        {3.35338f, 0.0307392f, 0.989189f, 0.371074f, -4.24089f},
        {3.39605f, 0.0308828f, 0.845455f, 0.333813f, -4.28662f},
        {3.62583f, 0.0303531f, 0.79754f, 0.325588f, -4.51725f},
        {4.11524f, 0.0172797f, 0.701166f, 0.275242f, -5.04394f},
        {4.4012f, 0.0159469f, 0.685892f, 0.270221f, -5.32997f},
        {5.8912f, 0.0171461f, 0.66366f, 0.288735f, -6.78867f},
    };
    const float x = clampf(ndotv, 0.f, 1.f) * 20.f + 1.f;
    const int ib = std::min(std::max(int(std::floor(alpha * 10.f)), 0), 10);
    const int it = std::min(ib + 1, 10);
    const float wb = float(it) - alpha * 10.f;
    const float tb = coeff[ib][4] + std::pow(x, coeff[ib][3]) + coeff[ib][0] / (1.f + coeff[ib][1] * std::pow(x, coeff[ib][2]));
    const float tt = coeff[it][4] + std::pow(x, coeff[it][3]) + coeff[it][0] / (1.f + coeff[it][1] * std::pow(x, coeff[it][2]));
    return 1.f - clampf(tb * wb + tt * (1.f - wb), 0.f, 1.f);
}

struct SheenBRDF {
    float3 color;
    float alpha_;
    SheenBRDF(float3 c, float a) : color(c), alpha_(std::max(a, 1e-6f)) {}
    float sheen_D(float ndoth) const {
        const float inv_alpha = 1.f / alpha_;
        const float cos2h = ndoth * ndoth;
        const float sin2h = std::max(1.f - cos2h, 0.0078125f);
        return (2.f + inv_alpha) * std::pow(sin2h, inv_alpha * 0.5f) * (0.5f * INV_PI);
    }
    float sheen_L(float x) const {
        const float r = (1.f - alpha_) * (1.f - alpha_);
        const float inv_r = 1.f - r;
        const float a = 25.3245f * r + 21.5473f * inv_r;
        const float b = 3.32435f * r + 3.82987f * inv_r;
        const float c = 0.16801f * r + 0.19823f * inv_r;
        const float d = -1.27393f * r - 1.97760f * inv_r;
        const float e = -4.85967f * r - 4.32054f * inv_r;
        return a / (1.f + b * std::pow(x, c)) + d * x + e;
    }
    float sheen_lambda(float c) const {
        if (c < 0.5f) return std::exp(sheen_L(c));
        return std::exp(2.f * sheen_L(0.5f) - sheen_L(1.f - c));
    }
    float sheen_G(float ndoto, float ndoti) const {
        return 1.f / (1.f + std::pow(sheen_lambda(ndoto), 1.f + 2.f * std::pow(1.f - ndoto, 8.f)) + sheen_lambda(ndoti));
    }
    float3 eval(float3 wi, float3 wo, Rng&) const {
        if (std::min(wi.z, wo.z) < BSDF_MIN_COS_THETA) return float3(0.f);
        const float3 h = normalize(wi + wo);
        return sheen_D(h.z) * sheen_G(wo.z, wi.z) * color / (4.f * wi.z);
    }
    std::optional<float3> sample(float3 wi, Rng& rng) const {
        const float3 wo = sample_cos(rng);
        if (std::min(wi.z, wo.z) < BSDF_MIN_COS_THETA) return std::nullopt;
        return wo;
    }
    float pdf(float3 wi, float3 wo) const {
        if (std::min(wi.z, wo.z) < BSDF_MIN_COS_THETA) return 0.f;
        return INV_PI * wo.z;
    }
    OptSample sample_eval(float3 wi, Rng& rng) const {
        if (auto s = sample_eval_impl(*this, wi, rng)) {
            BsdfSample r = *s;
            r.lobe = Lobe::Diffuse;
            return r;
        }
        return std::nullopt;
    }
    float3 get_albedo(float3 wi) const { return color * sheen_directional_albedo(wi.z, alpha_); }
    float3 get_preintegrated_specular(float3, float3) const { return float3(0.f); }
    float alpha() const { return alpha_; }
};

// ------------------------------------------------------------------ Sheen layer (bsdf-sheen-layer.slang:13-80)
template <class Base> struct SheenLayerBSDF {
    SheenBRDF sheen;
    Base base;
    float sheen_alpha, max_sheen_color;
    SheenLayerBSDF(SheenBRDF s, Base b, float sa, float mc) : sheen(s), base(b), sheen_alpha(sa), max_sheen_color(mc) {}
    float albedo_scaling(float ndot) const { return 1.f - max_sheen_color * sheen_directional_albedo(ndot, sheen_alpha); }
    float prob_sheen(float ndotv) const {
        return clampf(max_sheen_color * sheen_directional_albedo(ndotv, sheen_alpha), 0.f, 0.9f);
    }
    float3 eval(float3 wi, float3 wo, Rng& rng) const {
        const float scaling = std::min(albedo_scaling(wi.z), albedo_scaling(wo.z));
        return sheen.eval(wi, wo, rng) + base.eval(wi, wo, rng) * scaling;
    }
    std::optional<float3> sample(float3 wi, Rng& rng) const {
        const float p = prob_sheen(wi.z);
        if (p > 0.0f && rng.next_float() < p) return sheen.sample(wi, rng);
        return base.sample(wi, rng);
    }
    float pdf(float3 wi, float3 wo) const { return lerp(base.pdf(wi, wo), sheen.pdf(wi, wo), prob_sheen(wi.z)); }
    OptSample sample_eval(float3 wi, Rng& rng) const {
        const float p = prob_sheen(wi.z);
        OptSample sub = (p > 0.0f && rng.next_float() < p) ? sheen.sample_eval(wi, rng) : base.sample_eval(wi, rng);
        if (sub) {
            const float pdf = this->pdf(wi, sub->wo);
            return BsdfSample(sub->wo, pdf, this->eval(wi, sub->wo, rng) / pdf, sub->lobe);
        }
        return std::nullopt;
    }
    float3 get_albedo(float3 wi) const { return sheen.get_albedo(wi) + albedo_scaling(wi.z) * base.get_albedo(wi); }
    float3 get_preintegrated_specular(float3 F0, float3 wi) const { return base.get_preintegrated_specular(F0, wi); }
    float alpha() const { return base.alpha(); }
};

// ------------------------------------------------------------------ the glTF nest (gltf-material.slang:18-27)
using SpecularBRDF = GGXBRDF;
using GltfOpaqueDielectric = IridescentFresnelMixBSDF<SpecularBRDF, LambertDiffuseBRDF>;
using GltfDielectric = MixBSDF<GltfOpaqueDielectric, IridescentRoughDielectricBSDF>;
using GltfConductor = IridescentConductorBSDF<SpecularBRDF>;
using GltfReflective = MixBSDF<GltfDielectric, GltfConductor>;
using GltfSheened = SheenLayerBSDF<GltfReflective>;
using GltfCoated = FresnelMixBSDF<SpecularBRDF, GltfSheened>;
using GltfBSDF = MixBSDF<GltfSheened, GltfCoated>;

constexpr float GLTF_DIELECTRIC_IOR_F0 = 0.04f; // gltf-material.slang:29

} // namespace orc
