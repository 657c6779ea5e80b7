// bsdf.cuh — the glTF metallic-roughness BSDF (and Lambert) flattened for the GPU.
//
// The reference builds a 7-deep generic nest (include/merian-shaders/shading/materials/gltf-material.slang:18-27):
//   GltfBSDF = Mix< Sheened, FresnelMix<GGX_coat, Sheened> >(coat_weight),  Sheened = SheenLayer< Reflective >,
//   Reflective = Mix< Mix< IridescentFresnelMix<GGX, Lambert>, IridescentRoughDielectric >(transmission),
//                     IridescentConductor<GGX> >(metallic)
// and every sample_eval re-evaluates its sibling lobes recursively for one-sample MIS. Here the nest is flattened into
//   (1) bsdf_sample_direction(): one walk down the tree that consumes the same random numbers in the same order and
//       returns only the sampled direction, and
//   (2) bsdf_eval_pdf(): ONE evaluation of f*cos and pdf of the whole mixture for (wi, wo), sharing the half vector
//       and the D/G terms between lobes.
// weight = eval / pdf is, lobe by lobe, what the nest computes (bsdf-mix.slang:31-52, bsdf-fresnel-mix.slang:42-52,
// bsdf-iridescence.slang:147-158, bsdf-sheen-layer.slang:57-68). Disabled lobes (weight 0) drop out exactly as the
// reference's `weight > 0` guards (bsdf-mix.slang:34) make them. FEAT = union of MPT_GLTF_* bits present in the scene:
// FEAT == 0 compiles the extension code out, mirroring Slang's Conditional<> specialisation (gltf-material.slang:64-90).
#pragma once

#include "vecmath.cuh"

namespace mpt {

constexpr float kMinCos = 1e-6f; // BSDF_MIN_COS_THETA, bsdf.slang:13
enum : uint32_t { F_TRANSMISSION = 1, F_VOLUME = 2, F_CLEARCOAT = 4, F_SHEEN = 8, F_IRIDESCENCE = 16, F_ANISOTROPY = 32 };

struct Bsdf {
    uint32_t gltf;      // 0: LambertDiffuseBRDF(base) only (diffuse-material.slang)
    f3 base;            // albedo / base colour
    float ax, ay;       // specular_alpha (per axis)
    float alpha;        // isotropic alpha (rough dielectric)
    float metallic;
    float f0;           // dielectric_f0
    float eta;          // relative ior incident / transmitted
    float tw;           // transmission weight
    float coat_w, coat_alpha;
    f3 sheen_col;
    float sheen_alpha, max_sheen;
    float iri, iri_ior, iri_thick;
};

// ---------------------------------------------------------------- Fresnel (utils/fresnel.slang)
MPT_DEV float pow5(float x) { return powf(x, 5.0f); }
MPT_DEV float schlick(float c, float F0) { return F0 + (1.0f - F0) * pow5(fmaxf(1.0f - c, 0.f)); } // :14-26
MPT_DEV f3 schlick3(float c, f3 F0) {
    const float p = pow5(fmaxf(1.0f - c, 0.f));
    return F0 + (mk3(1.0f) - F0) * p;
}
MPT_DEV float fresnel_dielectric(float cos_i, float eta) { // :33-42
    const float sin2_t = eta * eta * (1.0f - cos_i * cos_i);
    if (sin2_t >= 1.0f) return 1.0f;
    const float cos_t = sqrtf(1.0f - sin2_t);
    const float r_parl = (eta * cos_i - cos_t) / (eta * cos_i + cos_t);
    const float r_perp = (cos_i - eta * cos_t) / (cos_i + eta * cos_t);
    return 0.5f * (r_parl * r_parl + r_perp * r_perp);
}
MPT_DEV f3 iri_sensitivity(float opd, f3 shift) { // :62-75
    const float phase = 2.0f * kPi * opd * 1.0e-9f;
    const f3 val = mk3(5.4856e-13f, 4.4201e-13f, 5.2481e-13f);
    const f3 pos = mk3(1.6810e+06f, 1.7953e+06f, 2.2084e+06f);
    const f3 var = mk3(4.3278e+09f, 9.3046e+09f, 6.6121e+09f);
    f3 xyz = val * vsqrt(2.0f * kPi * var) * vcos(pos * phase + shift) * vexp(-phase * phase * var);
    xyz.x += 9.7470e-14f * sqrtf(2.0f * kPi * 4.5282e+09f) * cosf(2.2399e+06f * phase + shift.x) * expf(-4.5282e+09f * phase * phase);
    xyz = xyz / 1.0685e-7f;
    return mk3(3.2404542f * xyz.x - 1.5371385f * xyz.y - 0.4985314f * xyz.z, -0.9692660f * xyz.x + 1.8760108f * xyz.y + 0.0415560f * xyz.z,
               0.0556434f * xyz.x - 0.2040259f * xyz.y + 1.0572252f * xyz.z);
}
MPT_DEV f3 fresnel_iridescent(float outside_ior, float film_ior, f3 base_f0, float thickness, float cos1) { // :79-129
    const float ts = clampf(thickness / 0.03f, 0.f, 1.f);
    const float iri_ior = mix(outside_ior, film_ior, ts * ts * (3.0f - 2.0f * ts));
    const float sin2 = sq(outside_ior / iri_ior) * (1.0f - sq(cos1));
    const float cos2_sq = 1.0f - sin2;
    if (!(cos2_sq > 0.0f)) return mk3(1.0f);
    const float cos2 = sqrtf(cos2_sq);
    const float r12 = schlick(cos1, sq((iri_ior - outside_ior) / (iri_ior + outside_ior)));
    const float t121 = 1.0f - r12;
    const float phi12 = iri_ior < outside_ior ? kPi : 0.0f;
    const float phi21 = kPi - phi12;
    const f3 s = vsqrt(vclamp(base_f0, 0.0f, 0.9999f));
    const f3 base_ior = (1.0f + s) / (1.0f - s);
    const f3 rr = (base_ior - iri_ior) / (base_ior + iri_ior);
    const f3 r23 = schlick3(cos2, rr * rr);
    const f3 phi23 = mk3(base_ior.x < iri_ior ? kPi : 0.f, base_ior.y < iri_ior ? kPi : 0.f, base_ior.z < iri_ior ? kPi : 0.f);
    const float opd = 2.0f * iri_ior * thickness * cos2;
    const f3 phi = mk3(phi21) + phi23;
    const f3 r123 = vclamp(r12 * r23, 1e-5f, 0.9999f);
    const f3 r123_sqrt = vsqrt(r123);
    const f3 rs = sq(t121) * r23 / (mk3(1.0f) - r123);
    f3 result = mk3(r12) + rs;
    f3 cm = rs - mk3(t121);
#pragma unroll
    for (int m = 1; m <= 2; ++m) {
        cm *= r123_sqrt;
        result += cm * (2.0f * iri_sensitivity(float(m) * opd, float(m) * phi));
    }
    return vclamp(result, 0.0f, 1.0f);
}

// ---------------------------------------------------------------- GGX terms (bsdfs/brdf-ggx.slang)
MPT_DEV float g1_over_wiz_iso(float wiz, float a2) { return 2.0f / (wiz + sqrtf(a2 + ((1.0f - a2) * (wiz * wiz)))); } // :42-44
MPT_DEV float g2_over_wiz_iso(float woz, float wiz, float a2) {                                                     // :51-55
    const float g1 = wiz * sqrtf(a2 + ((1.0f - a2) * (woz * woz)));
    const float g2 = woz * sqrtf(a2 + ((1.0f - a2) * (wiz * wiz)));
    return (2.0f * woz) / (g1 + g2);
}
MPT_DEV float g1_over_wiz(f3 wi, float ax, float ay) { return 2.0f / (wi.z + length(mk3(ax * wi.x, ay * wi.y, wi.z))); } // :71-73
MPT_DEV float g2_over_wiz(f3 wo, f3 wi, float ax, float ay) {                                                        // :78-82
    const float g1 = wi.z * length(mk3(ax * wo.x, ay * wo.y, wo.z));
    const float g2 = wo.z * length(mk3(ax * wi.x, ay * wi.y, wi.z));
    return (2.0f * wo.z) / (g1 + g2);
}
MPT_DEV float ndf_iso(float nh2, float a2) { return a2 / (kPi * sq((nh2 * a2) + (1.0f - nh2))); } // :88-91
MPT_DEV float ndf(f3 h, float ax, float ay) {                                                    // :95-100
    const float axy = ax * ay;
    const f3 v = mk3(ay * h.x, ax * h.y, axy * h.z);
    const float w2 = axy / dot(v, v);
    return axy * w2 * w2 / kPi;
}
// Dupuy & Benyoub spherical-cap VNDF sampling, :137-152
MPT_DEV f3 vndf_sample_h(f3 wi, float ax, float ay, float r0, float r1) {
    const f3 wis = normalize(mk3(wi.x * ax, wi.y * ay, wi.z));
    const float phi = ((2.0f * r0) - 1.0f) * kPi;
    const float z = fmaf(1.0f - r1, 1.0f + wis.z, -wis.z);
    const float st = sqrtf(clampf(1.0f - (z * z), 0.0f, 1.0f));
    const f3 c = mk3(st * cosf(phi), st * sinf(phi), z);
    const f3 wm = c + wis;
    return normalize(mk3(wm.x * ax, wm.y * ay, wm.z));
}
MPT_DEV float vndf_h_pdf(f3 wi, f3 h, float a2) { // :234-239
    return ndf_iso(sq(h.z), a2) * g1_over_wiz_iso(wi.z, a2) * fmaxf(0.f, dot(wi, h));
}

// GGXBRDF::eval / pdf for a given reflection half vector (:345-377): returns (f*cos, pdf)
MPT_DEV void ggx_reflect(f3 wi, f3 wo, f3 h, float ax, float ay, float& e, float& p) {
    if (fminf(wi.z, wo.z) < kMinCos) { e = 0.f; p = 0.f; return; }
    const float D = ndf(h, ax, ay);
    e = (D * g2_over_wiz(wo, wi, ax, ay)) / 4.0f;
    p = (D * g1_over_wiz(wi, ax, ay)) / 4.0f;
}

// ---------------------------------------------------------------- sheen (brdf-sheen.slang, bsdf-sheen-layer.slang)
MPT_DEV float sheen_albedo(float ndotv, float alpha) { // brdf-sheen.slang:94-118
    const float coeff[11][5] = {
        {1.01108f, 75.2286f, -9.4671f, 0.000973445f, -1.01325f}, {8.44447f, 0.0135054f, 1.11637f, 0.479139f, -9.30873f},
        {5.67124f, 0.0194668f, 1.12489f, 0.453551f, -6.53104f},  {4.50086f, 0.0237758f, 1.09376f, 0.423665f, -5.36521f},
        {3.77781f, 0.0273286f, 1.0666f, 0.400267f, -4.65601f},   {3.35338f, 0.0307392f, 0.989189f, 0.371074f, -4.24089f},
        {3.39605f, 0.0308828f, 0.845455f, 0.333813f, -4.28662f}, {3.62583f, 0.0303531f, 0.79754f, 0.325588f, -4.51725f},
        {4.11524f, 0.0172797f, 0.701166f, 0.275242f, -5.04394f}, {4.4012f, 0.0159469f, 0.685892f, 0.270221f, -5.32997f},
        {5.8912f, 0.0171461f, 0.66366f, 0.288735f, -6.78867f},
    };
    const float x = clampf(ndotv, 0.f, 1.f) * 20.f + 1.f;
    const int ib = min(max(int(floorf(alpha * 10.f)), 0), 10);
    const int it = min(ib + 1, 10);
    const float wb = float(it) - alpha * 10.f;
    const float tb = coeff[ib][4] + powf(x, coeff[ib][3]) + coeff[ib][0] / (1.f + coeff[ib][1] * powf(x, coeff[ib][2]));
    const float tt = coeff[it][4] + powf(x, coeff[it][3]) + coeff[it][0] / (1.f + coeff[it][1] * powf(x, coeff[it][2]));
    return 1.f - clampf(tb * wb + tt * (1.f - wb), 0.f, 1.f);
}
MPT_DEV float sheen_L(float x, float alpha) { // :29-38
    const float r = (1.f - alpha) * (1.f - alpha), ir = 1.f - r;
    const float a = 25.3245f * r + 21.5473f * ir, b = 3.32435f * r + 3.82987f * ir, c = 0.16801f * r + 0.19823f * ir;
    const float d = -1.27393f * r - 1.97760f * ir, e = -4.85967f * r - 4.32054f * ir;
    return a / (1.f + b * powf(x, c)) + d * x + e;
}
MPT_DEV float sheen_lambda(float c, float alpha) { // :40-45
    if (c < 0.5f) return expf(sheen_L(c, alpha));
    return expf(2.f * sheen_L(0.5f, alpha) - sheen_L(1.f - c, alpha));
}
MPT_DEV f3 sheen_eval(f3 wi, f3 wo, f3 h, f3 color, float alpha_in) { // :53-60
    if (fminf(wi.z, wo.z) < kMinCos) return mk3(0.f);
    const float alpha = fmaxf(alpha_in, 1e-6f);
    const float inv_alpha = 1.f / alpha;
    const float sin2h = fmaxf(1.f - h.z * h.z, 0.0078125f);
    const float D = (2.f + inv_alpha) * powf(sin2h, inv_alpha * 0.5f) * (0.5f * kInvPi);
    const float G = 1.f / (1.f + powf(sheen_lambda(wo.z, alpha), 1.f + 2.f * powf(1.f - wo.z, 8.f)) + sheen_lambda(wi.z, alpha));
    return D * G * color / (4.f * wi.z);
}

// ---------------------------------------------------------------- rough dielectric (bsdf-iridescence.slang:184-323)
template <uint32_t FEAT> MPT_DEV f3 rd_reflectance(const Bsdf& b, float widoth) { // :212-219
    const float d = fresnel_dielectric(widoth, b.eta);
    if (!(FEAT & F_IRIDESCENCE) || b.iri <= 0.0f) return mk3(d);
    return mix(mk3(d), fresnel_iridescent(1.0f, b.iri_ior, mk3(b.f0), b.iri_thick, widoth), b.iri);
}
template <uint32_t FEAT> MPT_DEV void rd_eval_pdf(const Bsdf& b, f3 wi, f3 wo, f3& e, float& p) { // :221-260
    e = mk3(0.f);
    p = 0.f;
    if (wi.z < kMinCos) return;
    const float a2 = sq(b.alpha);
    if (wo.z > kMinCos) {
        const f3 h = normalize(wi + wo);
        const f3 R = rd_reflectance<FEAT>(b, fabsf(dot(wi, h)));
        float ge, gp;
        ggx_reflect(wi, wo, h, b.alpha, b.alpha, ge, gp);
        e = R * ge;
        p = max3(R) * (ndf_iso(sq(fmaxf(0.0f, h.z)), a2) * g1_over_wiz_iso(wi.z, a2)) / 4.0f;
    } else if (wo.z < -kMinCos) {
        f3 h = normalize(wo + wi * b.eta);
        h = h * sgn(h.z);
        const float widoth = dot(wi, h), wodoth = dot(wo, h);
        const f3 R = rd_reflectance<FEAT>(b, fabsf(widoth));
        const float denom = wodoth + b.eta * widoth;
        // GGXBTDF::eval, brdf-ggx.slang:428-446
        const float G2 = g2_over_wiz_iso(fabsf(wo.z), wi.z, a2);
        const float D = ndf_iso(sq(h.z), a2);
        const float jac = sq(b.eta) * widoth * wodoth / sq(denom);
        e = (mk3(1.0f) - R) * (b.base * (D * G2 * fabsf(jac)));
        p = (1.0f - max3(R)) * vndf_h_pdf(wi, h, a2) * fabsf(wodoth) / sq(denom);
    }
}

// ---------------------------------------------------------------- full mixture: f*cos and pdf for (wi, wo)
template <uint32_t FEAT> MPT_DEV void bsdf_eval_pdf(const Bsdf& b, f3 wi, f3 wo, f3& out_e, float& out_p) {
    if (!b.gltf) { // LambertDiffuseBRDF, brdf-lambert-diffuse.slang:16-47
        if (fminf(wi.z, wo.z) < kMinCos) { out_e = mk3(0.f); out_p = 0.f; return; }
        out_e = kInvPi * b.base * wo.z;
        out_p = kInvPi * wo.z;
        return;
    }
    const f3 h = normalize(wi + wo);
    const float widoth_s = dot(wi, h);
    float ggx_e, ggx_p;
    ggx_reflect(wi, wo, h, b.ax, b.ay, ggx_e, ggx_p);

    // ---- Reflective = Mix(dielectric, conductor, metallic)
    f3 refl_e = mk3(0.f);
    float refl_p = 0.f;
    if (b.metallic < 1.0f) {
        // opaque dielectric: IridescentFresnelMix<GGX, Lambert> (bsdf-iridescence.slang:113-145)
        const bool up = fminf(wi.z, wo.z) >= kMinCos;
        const f3 lam_e = up ? kInvPi * b.base * wo.z : mk3(0.f);
        const float lam_p = up ? kInvPi * wo.z : 0.f;
        const float widoth = fabsf(widoth_s);
        const float f = fresnel_dielectric(widoth, b.eta);
        f3 die_e = mix(lam_e, mk3(ggx_e), f);
        if ((FEAT & F_IRIDESCENCE) && b.iri > 0.0f) {
            const f3 irf = fresnel_iridescent(1.0f, b.iri_ior, mk3(b.f0), b.iri_thick, widoth);
            die_e = mix(die_e, (mk3(1.0f) - irf) * lam_e + irf * ggx_e, b.iri);
        }
        float die_p = mix(lam_p, ggx_p, fresnel_dielectric(wi.z, b.eta));
        if ((FEAT & F_TRANSMISSION) && b.tw > 0.0f) { // Mix(opaque, rough dielectric, tw)
            f3 rd_e;
            float rd_p;
            rd_eval_pdf<FEAT>(b, wi, wo, rd_e, rd_p);
            die_e = mix(die_e, rd_e, b.tw);
            die_p = mix(die_p, rd_p, b.tw);
        }
        refl_e = die_e;
        refl_p = die_p;
    }
    if (b.metallic > 0.0f) {
        // IridescentConductor<GGX> (bsdf-iridescence.slang:27-52): Fresnel at the microfacet angle
        f3 F = schlick3(widoth_s, b.base);
        if ((FEAT & F_IRIDESCENCE) && b.iri > 0.0f) F = mix(F, fresnel_iridescent(1.0f, b.iri_ior, b.base, b.iri_thick, widoth_s), b.iri);
        const f3 con_e = F * ggx_e;
        if (b.metallic < 1.0f) {
            refl_e = mix(refl_e, con_e, b.metallic);
            refl_p = mix(refl_p, ggx_p, b.metallic);
        } else {
            refl_e = con_e;
            refl_p = ggx_p;
        }
    }
    // ---- Sheened = SheenLayer<Reflective> (bsdf-sheen-layer.slang:38-55)
    f3 sh_e = refl_e;
    float sh_p = refl_p;
    if ((FEAT & F_SHEEN) && b.max_sheen > 0.0f) {
        const float Ei = sheen_albedo(wi.z, b.sheen_alpha), Eo = sheen_albedo(wo.z, b.sheen_alpha);
        const float scaling = fminf(1.f - b.max_sheen * Ei, 1.f - b.max_sheen * Eo);
        sh_e = sheen_eval(wi, wo, h, b.sheen_col, b.sheen_alpha) + refl_e * scaling;
        const float cos_p = (fminf(wi.z, wo.z) < kMinCos) ? 0.f : kInvPi * wo.z;
        sh_p = mix(refl_p, cos_p, clampf(b.max_sheen * Ei, 0.f, 0.9f));
    }
    // ---- GltfBSDF = Mix(Sheened, FresnelMix<GGX_coat, Sheened>(0.04), coat_w) (bsdf-fresnel-mix.slang:19-40)
    out_e = sh_e;
    out_p = sh_p;
    if ((FEAT & F_CLEARCOAT) && b.coat_w > 0.0f) {
        float ce, cp;
        ggx_reflect(wi, wo, h, b.coat_alpha, b.coat_alpha, ce, cp);
        const f3 coated_e = mix(sh_e, mk3(ce), schlick(widoth_s, 0.04f));
        const float coated_p = mix(sh_p, cp, schlick(wi.z, 0.04f));
        out_e = mix(sh_e, coated_e, b.coat_w);
        out_p = mix(sh_p, coated_p, b.coat_w);
    }
}

// ---------------------------------------------------------------- directional albedo (BSDF::get_albedo, bsdf.slang:19)
// A denoiser guide, not part of the pt estimate: it feeds GBuffer tex3 (gbuffer_rt.slang:71-72) and through it
// merian_render_demodulate_albedo. ggx_integral_approx_correlated: brdf-ggx.slang:285-318 (RT Gems ch. 32 fit).
MPT_DEV f3 ggx_preintegrated(f3 F0, float alpha, float cos_theta) {
    const float c = fabsf(cos_theta), c2 = c * c, c3 = c2 * c, a = alpha, a3 = alpha * alpha * alpha;
    const float m1x = 0.995367f + -1.38839f * c, m1y = -0.24751f + 1.97442f * c;
    const float m2x = 1.0f + 2.68132f * c + 52.366f * c3, m2y = 16.0932f + -3.98452f * c + 59.3013f * c3, m2z = -5.18731f + 255.259f * c + 2544.07f * c3;
    float bias = (m1x + m1y * a) / (m2x + m2y * a + m2z * a3);
    const float m3x = -0.0564526f + 3.82901f * c, m3y = 16.91f + -11.0303f * c;
    const float m4x = 1.0f + 4.11118f * c2 + -1.37886f * c3, m4y = 19.3254f + -28.9947f * c2 + 16.9514f * c3, m4z = 0.545386f + 96.0994f * c2 + -79.4492f * c3;
    const float scale = (m3x + m3y * a) / (m4x + m4y * a + m4z * a3);
    bias *= clampf((F0.x + F0.y + F0.z) * (1.0f / 3.0f) * 50.0f, 0.f, 1.f);
    return F0 * fmaxf(0.0f, scale) + fmaxf(0.0f, bias);
}
template <uint32_t FEAT> MPT_DEV f3 bsdf_albedo(const Bsdf& b, f3 wi) {
    if (!b.gltf) return b.base; // LambertDiffuseBRDF::get_albedo
    const float spec_alpha = sqrtf(b.ax * b.ay);
    const bool iri = (FEAT & F_IRIDESCENCE) && b.iri > 0.0f;
    // Reflective = Mix(Dielectric, Conductor, metallic) (bsdf-mix.slang:51-54: a plain lerp, both branches evaluated)
    // opaque dielectric: IridescentFresnelMix<GGX, Lambert>::get_albedo, bsdf-iridescence.slang:160-170
    f3 spec = ggx_preintegrated(mk3(b.f0), spec_alpha, wi.z);
    if (iri) {
        const f3 irf = fresnel_iridescent(1.0f, b.iri_ior, mk3(b.f0), b.iri_thick, wi.z);
        const float die = fresnel_dielectric(wi.z, b.eta);
        spec = vclamp(spec * mix(mk3(1.0f), irf / fmaxf(die, 1e-4f), b.iri), 0.0f, 1.0f);
    }
    f3 dielectric = spec + (mk3(1.0f) - spec) * b.base;
    {   // Mix(opaque, IridescentRoughDielectric, transmission weight); rough dielectric albedo: bsdf-iridescence.slang:305-315
        const f3 f0 = mk3(sq((1.0f - b.eta) / (1.0f + b.eta)));
        f3 rs = ggx_preintegrated(f0, b.alpha, wi.z);
        if (iri) rs = vclamp(rs * rd_reflectance<FEAT>(b, wi.z) / fmaxf(fresnel_dielectric(wi.z, b.eta), 1e-4f), 0.0f, 1.0f);
        const f3 rough = rs + (mk3(1.0f) - rs) * sq(b.eta) * b.base;
        dielectric = mix(dielectric, rough, (FEAT & F_TRANSMISSION) ? b.tw : 0.f);
    }
    // IridescentConductor<GGX>::get_albedo, bsdf-iridescence.slang:59-65
    f3 conductor = ggx_preintegrated(b.base, spec_alpha, wi.z);
    if (iri) {
        const f3 F = mix(schlick3(wi.z, b.base), fresnel_iridescent(1.0f, b.iri_ior, b.base, b.iri_thick, wi.z), b.iri);
        const f3 sch = schlick3(wi.z, b.base);
        conductor = conductor * F / mk3(fmaxf(sch.x, 1e-4f), fmaxf(sch.y, 1e-4f), fmaxf(sch.z, 1e-4f));
    }
    const f3 reflective = mix(dielectric, conductor, b.metallic);
    // SheenLayer::get_albedo, bsdf-sheen-layer.slang:69-71 (sheen colour 0 => sheen term 0, scaling 1)
    f3 sheened = reflective;
    if (FEAT & F_SHEEN) {
        // SheenBRDF clamps its own alpha to 1e-6 (brdf-sheen.slang:15), the layer's albedo scaling does not
        sheened = b.sheen_col * sheen_albedo(wi.z, fmaxf(b.sheen_alpha, 1e-6f)) + (1.f - b.max_sheen * sheen_albedo(wi.z, b.sheen_alpha)) * reflective;
    }
    // GltfBSDF = Mix(Sheened, FresnelMix<GGX coat, Sheened>(0.04), coat weight), bsdf-fresnel-mix.slang:50-53
    const f3 cs = ggx_preintegrated(mk3(0.04f), b.coat_alpha, wi.z);
    const f3 coated = cs + (mk3(1.0f) - cs) * sheened;
    return mix(sheened, coated, (FEAT & F_CLEARCOAT) ? b.coat_w : 0.f);
}

// ---------------------------------------------------------------- direction sampling: same RNG draws, same order as the nest
MPT_DEV bool sample_cos_dir(f3 wi, Rng& rng, f3& wo) { // sampling.slang:24-29 + Lambert/Sheen guards
    const float x0 = rng.next_float(), x1 = rng.next_float();
    const float su = sqrtf(x0);
    wo = mk3(su * cosf(kTwoPi * x1), su * sinf(kTwoPi * x1), sqrtf(1.0f - x0));
    return fminf(wi.z, wo.z) >= kMinCos;
}
MPT_DEV bool sample_ggx_dir(f3 wi, float ax, float ay, Rng& rng, f3& wo) { // GGXBRDF::sample_eval, brdf-ggx.slang:379-396
    if (wi.z < kMinCos) return false;
    const float r0 = rng.next_float(), r1 = rng.next_float();
    const f3 h = vndf_sample_h(wi, ax, ay, r0, r1);
    wo = 2.0f * dot(wi, h) * h - wi;
    return wo.z >= kMinCos;
}

template <uint32_t FEAT> MPT_DEV bool bsdf_sample_direction(const Bsdf& b, f3 wi, Rng& rng, f3& wo) {
    if (!b.gltf) return sample_cos_dir(wi, rng, wo);
    // Mix(Sheened, Coated, coat_w): bsdf-mix.slang:31-42
    if ((FEAT & F_CLEARCOAT) && b.coat_w > 0.0f && rng.next_float() < b.coat_w) {
        // FresnelMix::sample_eval, bsdf-fresnel-mix.slang:42-52
        const float F = schlick(wi.z, 0.04f);
        if (rng.next_float() < F) return sample_ggx_dir(wi, b.coat_alpha, b.coat_alpha, rng, wo);
    }
    // SheenLayer::sample_eval, bsdf-sheen-layer.slang:57-68
    if ((FEAT & F_SHEEN) && b.max_sheen > 0.0f) {
        const float p = clampf(b.max_sheen * sheen_albedo(wi.z, b.sheen_alpha), 0.f, 0.9f);
        if (p > 0.0f && rng.next_float() < p) return sample_cos_dir(wi, rng, wo);
    }
    // Mix(dielectric, conductor, metallic)
    if (b.metallic > 0.0f && rng.next_float() < b.metallic) return sample_ggx_dir(wi, b.ax, b.ay, rng, wo);
    // Mix(opaque, rough dielectric, tw)
    if ((FEAT & F_TRANSMISSION) && b.tw > 0.0f && rng.next_float() < b.tw) {
        // IridescentRoughDielectricBSDF::sample_eval, bsdf-iridescence.slang:262-301
        if (wi.z < kMinCos) return false;
        const float r0 = rng.next_float(), r1 = rng.next_float();
        const f3 wm = vndf_sample_h(wi, b.alpha, b.alpha, r0, r1);
        const float widoth = dot(wi, wm);
        const float p = max3(rd_reflectance<FEAT>(b, fabsf(widoth)));
        if (rng.next_float() < p) {
            wo = 2.0f * widoth * wm - wi;
            return wo.z >= kMinCos;
        }
        // refract(-wi, wm, eta)
        const float k = 1.0f - b.eta * b.eta * (1.0f - widoth * widoth);
        if (k < 0.0f) return false;
        wo = b.eta * (-wi) - (b.eta * (-widoth) + sqrtf(k)) * wm;
        if ((wo.x == 0.f && wo.y == 0.f && wo.z == 0.f) || wo.z > -kMinCos) return false;
        return true;
    }
    // IridescentFresnelMix::sample_eval, bsdf-iridescence.slang:147-158
    const float f = fresnel_dielectric(wi.z, b.eta);
    if (rng.next_float() < f) return sample_ggx_dir(wi, b.ax, b.ay, rng, wo);
    return sample_cos_dir(wi, rng, wo);
}

// sample_eval of the whole BSDF: direction, pdf and weight = f*cos / pdf.
template <uint32_t FEAT> MPT_DEV bool bsdf_sample_eval(const Bsdf& b, f3 wi, Rng& rng, f3& wo, float& pdf, f3& weight) {
    if (!bsdf_sample_direction<FEAT>(b, wi, rng, wo)) return false;
    if (!b.gltf) { // LambertDiffuseBRDF::sample_eval returns the albedo directly (brdf-lambert-diffuse.slang:30-38)
        pdf = kInvPi * wo.z;
        weight = b.base;
        return true;
    }
    f3 e;
    bsdf_eval_pdf<FEAT>(b, wi, wo, e, pdf);
    weight = e / pdf;
    return true;
}

} // namespace mpt
