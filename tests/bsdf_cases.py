"""The reference's BSDF property suite (tests/test_bsdf.cpp:343-596 + tests/bsdf/bsdf-checks.slang:30-135), restated
once and shared by the CPU-oracle tests (test_bsdf_oracle.py) and the CUDA tests (test_gpu_bsdf.py).

A `run(**params) -> float[16]` callable evaluates the check kernel; slots as in bsdf-checks.slang:117-134:
  0..2 furnace, 3..5 furnace stderr, 6 mean cos/pdf, 7 its stderr, 8 valid samples, 9 max |weight - eval/pdf|,
  10 max relative |pdf - pdf()|, 11 INT pdf dwo, 12 max reciprocity error, 13..15 get_albedo.
"""
import math

import numpy as np

# constants of tests/test_bsdf.cpp:30-58
INTEGRAL_COS_HEMISPHERE = math.pi
MC_SIGMA = 6.0
MC_NUMERIC_FLOOR = 1e-3
FLOAT_EQ_TOL = 1e-4
PDF_REL_TOL = 1e-3
QUADRATURE_TOL = 1e-3
ALBEDO_REL_TOL = 0.15

# mpt_bsdf_check_params.bsdf
LAMBERT, GGX, GGX_ANISO, FRESNEL_MIX, IRIDESCENT_MIX, ROUGH_DIELECTRIC, SHEEN, GLTF_FULL, CONDUCTOR, SHEEN_LAYER = 0, 1, 2, 3, 4, 5, 6, 7, 8, 9


def wi_zenith(deg, azimuth_deg=0.0):
    t, p = math.radians(deg), math.radians(azimuth_deg)
    return (math.sin(t) * math.cos(p), math.sin(t) * math.sin(p), math.cos(t))


WI_NORMAL, WI_GRAZING = wi_zenith(0.0), wi_zenith(80.0)


class R:
    def __init__(self, out):
        o = np.asarray(out, np.float64)
        self.furnace, self.furnace_se = o[0:3], o[3:6]
        self.mean_cos, self.se_cos, self.valid = o[6], o[7], o[8]
        self.werr, self.perr, self.pdf_int, self.recip = o[9], o[10], o[11], o[12]
        self.albedo = o[13:16]


def sample_pdf_match(r):  # test_bsdf.cpp:286-291
    assert r.valid > 0
    assert abs(r.mean_cos - INTEGRAL_COS_HEMISPHERE) <= MC_SIGMA * r.se_cos + MC_NUMERIC_FLOOR, (r.mean_cos, r.se_cos)


def sample_eval_consistent(r):  # :295-298
    assert r.perr < PDF_REL_TOL, r.perr
    assert r.werr < FLOAT_EQ_TOL, r.werr


def reciprocal(r):  # :302-304
    assert r.recip < FLOAT_EQ_TOL, r.recip


def energy_conserving(r):  # :308-314
    assert np.isfinite(r.furnace).all()
    assert (r.furnace <= 1.0 + MC_SIGMA * r.furnace_se).all(), (r.furnace, r.furnace_se)


def pdf_normalized(r):  # :318-320
    assert abs(r.pdf_int - 1.0) <= QUADRATURE_TOL, r.pdf_int


def albedo_matches_furnace(r):  # :326-334
    for a, f, se in zip(r.albedo, r.furnace, r.furnace_se):
        assert abs(a - f) <= ALBEDO_REL_TOL * max(abs(f), 0.05) + MC_SIGMA * se, (a, f)


# (name, params, checks, extra) — one entry per TEST_F / TEST_P instantiation of the reference
def cases(include_albedo=True):
    A = [albedo_matches_furnace] if include_albedo else []
    out = []
    lam = dict(bsdf=LAMBERT, albedo=(0.8, 0.5, 0.2))
    out.append(("LambertNormal", dict(lam, wi=WI_NORMAL), [sample_pdf_match, sample_eval_consistent, reciprocal, energy_conserving, pdf_normalized] + A, "lambert_furnace"))
    out.append(("LambertGrazing", dict(lam, wi=WI_GRAZING), [sample_pdf_match, sample_eval_consistent, reciprocal, energy_conserving, pdf_normalized], "lambert_furnace_x"))
    for a in (0.09, 0.3, 0.5, 1.0):
        for nm, wi in (("Normal", WI_NORMAL), ("Grazing", WI_GRAZING)):
            out.append((f"SheenAlpha{a}{nm}", dict(bsdf=SHEEN, albedo=(0.8, 0.5, 0.2), alpha=a, wi=wi),
                        [sample_pdf_match, sample_eval_consistent, energy_conserving, pdf_normalized] + A, None))
    for a in (0.05, 0.3, 0.8):
        for nm, wi in (("Normal", WI_NORMAL), ("Grazing", WI_GRAZING)):
            out.append((f"GGXRoughness{a}{nm}", dict(bsdf=GGX, alpha=a, wi=wi), [sample_pdf_match, sample_eval_consistent, reciprocal, energy_conserving], None))
    for at, ab in ((0.5, 0.05), (0.05, 0.5), (1.0, 0.01), (0.3, 0.1)):
        for nm, wi in (("Normal", wi_zenith(30, 40)), ("Grazing", wi_zenith(80, 40))):
            out.append((f"GGXAniso{at}_{ab}{nm}", dict(bsdf=GGX_ANISO, alpha=at, alpha_bitangent=ab, wi=wi),
                        [sample_pdf_match, sample_eval_consistent, reciprocal, energy_conserving], None))
    for a in (0.05, 0.3, 0.8):
        for nm, wi in (("Normal", WI_NORMAL), ("Grazing", WI_GRAZING)):
            out.append((f"FresnelMix{a}{nm}", dict(bsdf=FRESNEL_MIX, alpha=a, albedo=(0.7, 0.7, 0.7), wi=wi), [sample_pdf_match, sample_eval_consistent, reciprocal], None))
    for s in (0.0, 1.0):
        for nm, wi in (("Normal", WI_NORMAL), ("Grazing", WI_GRAZING)):
            out.append((f"IridescentMix{s}{nm}", dict(bsdf=IRIDESCENT_MIX, alpha=0.3, albedo=(0.7, 0.7, 0.7), iridescence=s, wi=wi),
                        [sample_pdf_match, sample_eval_consistent, reciprocal], None))
    # RoughDielectric plain = the iridescent variant at strength 0 is a distinct config in the reference
    # (bsdf-rough-dielectric.slang is not reachable from GltfBSDF); the iridescent one is.
    for a, s in ((0.3, 0.0), (0.3, 1.0), (0.5, 1.0)):
        for nm, wi in (("Normal", WI_NORMAL), ("Grazing", WI_GRAZING)):
            chk = [sample_eval_consistent] + (A if wi[2] > 0.9 else [])
            out.append((f"RoughDielectricIrid{a}_{s}{nm}", dict(bsdf=ROUGH_DIELECTRIC, alpha=a, iridescence=s, albedo=(0.9, 0.95, 1.0), wi=wi), chk, None))
    for a in (0.05, 0.3, 0.8):
        for nm, wi in (("Normal", WI_NORMAL), ("Grazing", WI_GRAZING)):
            out.append((f"Conductor{a}{nm}", dict(bsdf=CONDUCTOR, alpha=a, f0=(0.95, 0.64, 0.54), wi=wi),
                        [sample_pdf_match, sample_eval_consistent, reciprocal, energy_conserving] + A, None))
    # Not in the reference's suite: the sheen lobe as it is reachable from GltfBSDF (SheenLayer over the dielectric; the device cannot
    # isolate SheenBRDF) and the whole nest. The reference's property checks that hold for a layered BSDF; both arms build the same nest.
    for a in (0.09, 0.3, 0.5, 1.0):
        for nm, wi in (("Normal", WI_NORMAL), ("Grazing", WI_GRAZING)):
            out.append((f"SheenLayerAlpha{a}{nm}", dict(bsdf=SHEEN_LAYER, albedo=(0.8, 0.5, 0.2), alpha=a, wi=wi),
                        [sample_pdf_match, sample_eval_consistent, energy_conserving], None))
    for s in (0.0, 1.0):
        for nm, wi in (("Normal", wi_zenith(20, 30)), ("Grazing", wi_zenith(75, 30))):
            out.append((f"GltfFull{s}{nm}", dict(bsdf=GLTF_FULL, albedo=(0.8, 0.5, 0.2), alpha=0.3, alpha_bitangent=0.1, iridescence=s, wi=wi),
                        [sample_eval_consistent, energy_conserving], None))
    return out


def run_case(run, params, checks, extra):
    r = R(run(**params))
    for c in checks:
        c(r)
    if extra == "lambert_furnace":  # Lambert furnace == albedo, test_bsdf.cpp:361-363
        assert (np.abs(r.furnace - np.array(params["albedo"])) <= MC_SIGMA * r.furnace_se + 1e-6).all()
    elif extra == "lambert_furnace_x":
        assert abs(r.furnace[0] - params["albedo"][0]) <= MC_SIGMA * r.furnace_se[0] + 1e-6
    return r
