"""Chi-square tests of the sampling routines on the path -- the reference's own test strategy for them:
src/tests/test_chisquare.cpp:508-573 (phase functions of data/tests/test_phase.xml: isotropic, hg g=0.9, hg g=-0.3; 20 incident
directions each) and 575-623 (emitters, sampleDirect against pdfDirect), with the contingency tables, the pooling of cells with
low expected frequencies, the Sidak correction and the significance level of src/libcore/chisquare.cpp:56-262 /
test_chisquare.cpp:30 restated in numpy.  The routines under test are the oracle's (oracle_capi.cpp: samplePhase, Medium::phaseEval,
emitterSampleDirect, pdfEmitterDirect); the device tracer and the device ground truth are bit-identical to them
(tests/test_tracer.py, tests/test_volpath.py), so what these tests pin carries over to the device."""
import ctypes as C

import numpy as np
import pytest
from scipy import stats

SIGNIFICANCE_LEVEL = 0.0025          # test_chisquare.cpp:30
CHISQR_MIN_EXP_FREQUENCY = 5         # include/mitsuba/core/chisquare.h:30


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def to_spherical(d):
    """toSphericalCoordinates (include/mitsuba/core/util.h): theta = acos(z), phi = atan2(y, x) in [0, 2 pi)"""
    theta = np.arccos(np.clip(d[:, 2], -1.0, 1.0))
    phi = np.arctan2(d[:, 1], d[:, 0])
    return theta, np.where(phi < 0, phi + 2 * np.pi, phi)


def spherical_direction(theta, phi):
    st = np.sin(theta)
    return np.stack([st * np.cos(phi), st * np.sin(phi), np.cos(theta)], -1)


class ChiSquare:
    """src/libcore/chisquare.cpp: fill() (contingency table of the samples, reference table = the pdf integrated over every
    (theta, phi) cell, there by an adaptive cubature at 1e-6, here by a fixed Gauss-Legendre / midpoint product rule) and
    runTest()."""

    def __init__(self, theta_bins=10, phi_bins=0, num_tests=1, sample_count=0):
        self.tb = theta_bins
        self.pb = phi_bins or 2 * theta_bins
        self.num_tests = num_tests
        self.sample_count = sample_count or self.tb * self.pb * 1000        # chisquare.cpp:53-54
        self.tolerance = self.sample_count * 1e-4

    def _cells(self, cells, pdf_fn, x, w):
        """integral of pdf * sin(theta) over the listed (i, j) cells by the product rule (x, w) on [0, 1]"""
        dt, dp = np.pi / self.tb, 2 * np.pi / self.pb
        out = np.zeros(len(cells))
        for k, (i, j) in enumerate(cells):
            T, P = np.meshgrid((i + x) * dt, (j + x) * dp, indexing="ij")
            f = pdf_fn(spherical_direction(T.reshape(-1), P.reshape(-1))).reshape(T.shape) * np.sin(T)     # ChiSquare::integrand
            out[k] = np.einsum("ab,a,b->", f, w, w) * dt * dp
        return out

    def fill(self, dirs, weights, pdf_fn, quad=48, refine=0):
        """refine > 0: the density has jumps (the edge of an emitter) -- midpoint rule with `quad` points per axis on every
        cell, then `refine` points per axis on the cells where the density is not zero everywhere and on their neighbours."""
        theta, phi = to_spherical(dirs)
        ti = np.clip(np.floor(theta * (self.tb / np.pi)).astype(int), 0, self.tb - 1)
        pi_ = np.clip(np.floor(phi * (self.pb / (2 * np.pi))).astype(int), 0, self.pb - 1)
        self.table = np.zeros((self.tb, self.pb))
        np.add.at(self.table, (ti, pi_), weights)
        cells = [(i, j) for i in range(self.tb) for j in range(self.pb)]
        if refine:
            x, w = (np.arange(quad) + 0.5) / quad, np.full(quad, 1.0 / quad)
        else:
            x, w = np.polynomial.legendre.leggauss(quad)
            x, w = 0.5 * (x + 1), 0.5 * w
        tab = self._cells(cells, pdf_fn, x, w).reshape(self.tb, self.pb)
        if refine:
            nz = tab > 0
            padded = np.pad(nz, ((1, 1), (0, 0)))                       # theta does not wrap, phi does
            grow = np.zeros_like(nz)
            for di in (0, 1, 2):
                for dj in (-1, 0, 1):
                    grow |= np.roll(padded[di:di + self.tb], dj, axis=1)
            again = [(i, j) for i in range(self.tb) for j in range(self.pb) if grow[i, j]]
            x, w = (np.arange(refine) + 0.5) / refine, np.full(refine, 1.0 / refine)
            for (i, j), v in zip(again, self._cells(again, pdf_fn, x, w)):
                tab[i, j] = v
        self.integral_table = tab
        self.ref = tab * self.sample_count
        return tab.sum()

    def run_test(self, pval_thresh=SIGNIFICANCE_LEVEL):
        """-> (accepted, p-value, chi-square statistic, degrees of freedom); chisquare.cpp:179-262"""
        table, ref = self.table.reshape(-1), self.ref.reshape(-1)
        pooled_counts = pooled_ref = chsq = 0.0
        pooled_cells = df = 0
        for idx in np.argsort(ref, kind="stable"):
            if ref[idx] == 0:
                if table[idx] > self.tolerance:
                    return False, 0.0, np.inf, df
            elif ref[idx] < CHISQR_MIN_EXP_FREQUENCY or (0 < pooled_ref < CHISQR_MIN_EXP_FREQUENCY):
                pooled_counts += table[idx]
                pooled_ref += ref[idx]
                pooled_cells += 1
            else:
                chsq += (table[idx] - ref[idx]) ** 2 / ref[idx]
                df += 1
        if pooled_cells > 0:
            chsq += (pooled_counts - pooled_ref) ** 2 / pooled_ref
            df += 1
        df -= 1
        assert df > 0, "too few degrees of freedom (ELowDoF)"
        pval = float(stats.chi2.sf(chsq, df))
        alpha = 1 - (1 - pval_thresh) ** (1.0 / self.num_tests)        # Sidak correction
        return pval >= alpha, pval, chsq, df


# ---- the harness itself: accepts a correct sampler, rejects wrong ones -------------------------------------------------------

def test_chisquare_harness_accepts_uniform_sphere_and_rejects_a_skewed_one():
    rng = np.random.default_rng(1)
    cs = ChiSquare()
    u = rng.random((cs.sample_count, 2))
    z = 1 - 2 * u[:, 1]
    r = np.sqrt(np.maximum(0, 1 - z * z))
    d = np.stack([r * np.cos(2 * np.pi * u[:, 0]), r * np.sin(2 * np.pi * u[:, 0]), z], -1)
    integral = cs.fill(d, np.ones(len(d)), lambda w: np.full(len(w), 1 / (4 * np.pi)))
    assert abs(integral - 1) < 1e-9
    ok, pval, _, df = cs.run_test()
    assert ok and df == 199, (pval, df)
    z2 = 1 - 2 * u[:, 1] ** 1.1                                            # a warped cosine
    r2 = np.sqrt(np.maximum(0, 1 - z2 * z2))
    d2 = np.stack([r2 * np.cos(2 * np.pi * u[:, 0]), r2 * np.sin(2 * np.pi * u[:, 0]), z2], -1)
    cs.fill(d2, np.ones(len(d2)), lambda w: np.full(len(w), 1 / (4 * np.pi)))
    assert not cs.run_test()[0]


# ---- test02_PhaseFunction ------------------------------------------------------------------------------------------------

def _phase_sample(lib, ptype, g, wi, u):
    n = len(u)
    wo, pdf = np.zeros((n, 3), np.float32), np.zeros(n, np.float32)
    wi = np.ascontiguousarray(wi, np.float32)
    assert lib.orc_test_phase_sample(C.c_int32(ptype), C.c_float(g), _p(wi), _p(u), C.c_uint32(n), _p(wo), _p(pdf)) == 0
    return wo, pdf


def _phase_eval(lib, ptype, g, wi, wo):
    wo = np.ascontiguousarray(wo, np.float32)
    wi = np.ascontiguousarray(wi, np.float32)
    val = np.zeros(len(wo), np.float32)
    assert lib.orc_test_phase_eval(C.c_int32(ptype), C.c_float(g), _p(wi), _p(wo), C.c_uint32(len(wo)), _p(val)) == 0
    return val.astype(np.float64)


@pytest.mark.parametrize("ptype,g", [(0, 0.0), (1, 0.9), (1, -0.3), (1, 0.0)],
                         ids=["isotropic", "hg g=0.9", "hg g=-0.3", "hg g=0 (the |g| < Epsilon branch)"])
def test_phase_function_sampling_matches_its_density(orc, ptype, g):
    """test_chisquare.cpp:508-573 on the phase functions of test_phase.xml the path supports: 20 incident directions, a
    10 x 20 table, 200 000 samples per direction, Sidak-corrected significance."""
    lib = orc.api().lib
    rng = np.random.default_rng(107 + int(100 * g))
    wi_samples = 20
    largest_weight_dev = 0.0
    for j in range(wi_samples):
        a = rng.random(2)
        z = 1 - 2 * a[1]
        r = np.sqrt(max(0.0, 1 - z * z))
        wi = np.array([r * np.cos(2 * np.pi * a[0]), r * np.sin(2 * np.pi * a[0]), z], np.float32)   # squareToUniformSphere
        cs = ChiSquare(10, 20, wi_samples)
        u = rng.random((cs.sample_count, 2), dtype=np.float32)
        wo, pdf = _phase_sample(lib, ptype, g, wi, u)
        assert np.allclose(np.linalg.norm(wo.astype(np.float64), axis=1), 1.0, atol=2e-6)
        # the adapter's consistency check (test_chisquare.cpp:262-306): the density sample() reports is eval() of what it returned
        ev = _phase_eval(lib, ptype, g, wi, wo)
        assert np.allclose(pdf, ev, rtol=1e-6, atol=0)
        largest_weight_dev = max(largest_weight_dev, float(np.abs(ev / pdf - 1).max()))
        integral = cs.fill(wo.astype(np.float64), np.ones(len(wo)), lambda w: _phase_eval(lib, ptype, g, wi, w))
        assert abs(integral - 1) < 2e-4, integral                          # the density integrates to one (fp32 evaluation)
        ok, pval, chsq, df = cs.run_test()
        assert ok, f"chi-square rejects wi={wi}: p={pval:.3e} chi2={chsq:.1f} df={df}"
    assert largest_weight_dev < 1e-6          # the importance weight of sample() is 1 (isotropic.cpp:69-74, hg.cpp:99-104)


def test_phase_chisquare_rejects_the_mirrored_lobe(orc):
    """the test has teeth where it matters for the path: hg sampled around +wi instead of -wi (the Frame(-pRec.wi) of hg.cpp:95)
    is rejected"""
    lib = orc.api().lib
    rng = np.random.default_rng(3)
    wi = np.array([0.36, -0.48, 0.8], np.float32)
    cs = ChiSquare(10, 20, 1)
    u = rng.random((cs.sample_count, 2), dtype=np.float32)
    wo, _ = _phase_sample(lib, 1, -0.3, wi, u)
    cs.fill(-wo.astype(np.float64), np.ones(len(wo)), lambda w: _phase_eval(lib, 1, -0.3, wi, w))
    assert not cs.run_test()[0]


# ---- test03_EmitterDirect ------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("ref", [(0.5, 0.5, 0.5), (0.2, 0.8, 0.35), (0.85, 0.4, 0.15)], ids=["centre", "near", "oblique"])
def test_area_emitter_direct_sampling_matches_its_density(pkg, orc, ref):
    """test_chisquare.cpp:575-623 for the path's emitter (an area light on triangles, src/emitters/area.cpp:93-120 through
    src/shapes/triangle mesh sampling): directions of sampleDirect from a reference point in the medium against pdfDirect of the
    surface the query ray meets (as EmitterAdapter::pdf, with the intersection the volpath restatement does before it,
    volpath.cpp:521-524).  The density jumps at the edge of the light, so the cells are integrated by a midpoint rule on a
    fine grid."""
    scene, em, rad = pkg.scenes.tracer_scene(16, 16, glass=False)
    o = orc.Oracle(volVolSamples=2, volSurfSamples=2, targetNumSlices=4)
    o.set_scene(scene)
    o.set_area_emitter(em, rad)
    lib = orc.api().lib
    ref = np.array(ref, np.float32)
    rng = np.random.default_rng(11)
    cs = ChiSquare(10, 20, 1)
    n = cs.sample_count
    u = rng.random((n, 2), dtype=np.float32)
    d, pdf, val = np.zeros((n, 3), np.float32), np.zeros(n, np.float32), np.zeros((n, 3), np.float32)
    assert lib.orc_test_emitter_sample_direct(o.h, _p(ref), _p(u), C.c_uint32(n), _p(d), _p(pdf), _p(val)) == 0
    assert (pdf > 0).all() and np.allclose(np.linalg.norm(d.astype(np.float64), axis=1), 1.0, atol=2e-6)
    # value = radiance / pdf (emitter.cpp sampleDirect): the product gives the radiance back
    assert np.allclose(val * pdf[:, None], rad[None, :], rtol=2e-6)

    def pdf_fn(w):
        w = np.ascontiguousarray(w, np.float32)
        out = np.zeros(len(w), np.float32)
        assert lib.orc_test_emitter_pdf_direct(o.h, _p(ref), _p(w), C.c_uint32(len(w)), _p(out)) == 0
        return out.astype(np.float64)

    # the density reported with the sample is the density of the query from the reference point
    q = pdf_fn(d[:4000])
    hit = q > 0                                    # (a sampled point on the very edge can miss by rounding)
    assert hit.mean() > 0.995 and np.allclose(q[hit], pdf[:4000][hit], rtol=2e-3)
    integral = cs.fill(d.astype(np.float64), np.ones(n), pdf_fn, quad=32, refine=640)
    assert abs(integral - 1) < 2e-3, integral
    ok, pval, chsq, df = cs.run_test()
    assert ok, f"chi-square rejects: p={pval:.3e} chi2={chsq:.1f} df={df}"


# ---- test01_BSDF ---------------------------------------------------------------------------------------------------------
BSDF_SMOOTH, BSDF_DIELECTRIC, BSDF_CONDUCTOR = 1, 2, 4        # include/alvrl.h: ALVRL_BSDF_*


def _bsdf_sample(lib, bits, albedo, optics, wi, u, radiance=False):
    n = len(u)
    wo, pdf, wt, delta = np.zeros((n, 3), np.float32), np.zeros(n, np.float32), np.zeros((n, 3), np.float32), np.zeros(n, np.uint8)
    al, w = np.ascontiguousarray(albedo, np.float32), np.ascontiguousarray(wi, np.float32)
    op = np.ascontiguousarray(optics, np.float32) if optics is not None else None
    assert lib.orc_test_bsdf_sample(C.c_uint32(bits), _p(al), _p(op) if op is not None else None, _p(w), _p(u), C.c_uint32(n), int(radiance),
                                    _p(wo), _p(pdf), _p(wt), _p(delta)) == 0
    return wo, pdf, wt, delta


def test_bsdf_flag_values_match_the_header():
    import re
    from conftest import ROOT
    import os
    txt = open(os.path.join(ROOT, "include", "alvrl.h")).read()
    for name, val in (("ALVRL_BSDF_SMOOTH", BSDF_SMOOTH), ("ALVRL_BSDF_DIELECTRIC", BSDF_DIELECTRIC), ("ALVRL_BSDF_CONDUCTOR", BSDF_CONDUCTOR)):
        m = re.search(r"#define\s+" + name + r"\s+(\w+)", txt)
        assert m and int(m.group(1).rstrip("u"), 0) == val, name


def test_diffuse_bsdf_sampling_matches_its_density(orc):
    """test_chisquare.cpp:398-506 (test01_BSDF) for the `diffuse` model of test_bsdf.xml: directions of sample() against pdf() =
    cos(theta) / pi on the upper hemisphere (diffuse.cpp:120-148), for 20 incident directions; the weight is the reflectance"""
    lib = orc.api().lib
    rng = np.random.default_rng(21)
    albedo = np.array([0.2, 0.5, 0.7], np.float32)
    wi_samples = 20
    for j in range(wi_samples):
        a = rng.random(2)
        z = a[1]                                                                    # the upper hemisphere (BSDFAdapter: wi.z > 0 for one-sided models)
        r = np.sqrt(max(0.0, 1 - z * z))
        wi = np.array([r * np.cos(2 * np.pi * a[0]), r * np.sin(2 * np.pi * a[0]), max(z, 1e-3)], np.float32)
        cs = ChiSquare(10, 20, wi_samples)
        u = rng.random((cs.sample_count, 2), dtype=np.float32)
        wo, pdf, wt, delta = _bsdf_sample(lib, BSDF_SMOOTH, albedo, None, wi, u)
        assert not delta.any() and (wo[:, 2] > 0).all()
        assert np.allclose(pdf, wo[:, 2] / np.pi, rtol=1e-6) and np.array_equal(wt, np.broadcast_to(albedo, wt.shape))
        integral = cs.fill(wo.astype(np.float64), np.ones(len(wo)), lambda w: np.where(w[:, 2] > 0, w[:, 2] / np.pi, 0.0), quad=32)
        assert abs(integral - 1) < 1e-6
        ok, pval, chsq, df = cs.run_test()
        assert ok, f"chi-square rejects: p={pval:.3e} chi2={chsq:.1f} df={df}"
    # from below, a one-sided diffuse surface returns no sample (diffuse.cpp:131-133)
    wo, pdf, wt, delta = _bsdf_sample(lib, BSDF_SMOOTH, albedo, None, np.array([0.3, 0.1, -0.9], np.float32), rng.random((16, 2), dtype=np.float32))
    assert not wt.any()


def _fresnel_dielectric(cos_i, eta):
    """unpolarised Fresnel reflectance, the textbook form (independent of fresnelDielectricExt's arrangement, util.cpp:651-681)"""
    if cos_i < 0:
        eta, cos_i = 1 / eta, -cos_i
    sin2_t = (1 - cos_i * cos_i) / (eta * eta)
    if sin2_t >= 1:
        return 1.0, 0.0
    cos_t = np.sqrt(1 - sin2_t)
    rs = (cos_i - eta * cos_t) / (cos_i + eta * cos_t)
    rp = (eta * cos_i - cos_t) / (eta * cos_i + cos_t)
    return 0.5 * (rs * rs + rp * rp), cos_t


@pytest.mark.parametrize("eta", [1.5, 1 / 1.5, 1.33])
def test_dielectric_bsdf_discrete_components(orc, eta):
    """the smooth `dielectric` of test_bsdf.xml has two discrete components (dielectric.cpp:281-364): the chi-square test
    compares how often each direction is drawn with the probability pdf() reports (chisquare.cpp:124-141: discrete directions
    enter the reference table with pdf x sampleCount).  Here additionally: the probabilities are the Fresnel terms of an
    independent formula, the directions obey reflection and Snell's law, and radiance picks up the squared relative index
    when it is transmitted while importance does not (322-326)."""
    lib = orc.api().lib
    rng = np.random.default_rng(1000 + int(eta * 100))    # (seed 133 draws a 4-sigma sample of the uniforms themselves in one of its 20 runs)
    optics = np.zeros(12, np.float32); optics[0] = eta; optics[6:12] = 1.0
    one = np.ones(3, np.float32)
    for j in range(20):
        a = rng.random(2)
        z = 1 - 2 * a[1]
        if abs(z) < 0.02:
            z = 0.02
        r = np.sqrt(max(0.0, 1 - z * z))
        wi = np.array([r * np.cos(2 * np.pi * a[0]), r * np.sin(2 * np.pi * a[0]), z], np.float32)
        wi /= np.linalg.norm(wi)
        n = 200_000
        u = rng.random((n, 2), dtype=np.float32)
        wo, pdf, wt, delta = _bsdf_sample(lib, BSDF_DIELECTRIC, one, optics, wi, u, radiance=False)
        assert delta.all()
        F, cos_t = _fresnel_dielectric(float(wi[2]), eta)
        refl = np.all(wo == np.array([-wi[0], -wi[1], wi[2]], np.float32), axis=1)
        assert np.allclose(pdf[refl], F, atol=2e-6) and np.allclose(pdf[~refl], 1 - F, atol=2e-6)
        # two cells of a contingency table: reflected / transmitted counts against pdf x sampleCount
        cs = ChiSquare(1, 2, 20)
        cs.table = np.array([[refl.sum(), (~refl).sum()]], np.float64)
        cs.ref = np.array([[F * n, (1 - F) * n]])
        if min(cs.ref.reshape(-1)) >= CHISQR_MIN_EXP_FREQUENCY:
            ok, pval, chsq, df = cs.run_test()
            assert ok, (wi, F, pval)
        else:
            assert abs(refl.mean() - F) < 1e-3
        if (~refl).any():                                                          # Snell: sin_t = sin_i / eta (towards the other side)
            t = wo[~refl][0].astype(np.float64)
            assert abs(np.linalg.norm(t) - 1) < 1e-5 and np.sign(t[2]) == -np.sign(wi[2]) and abs(abs(t[2]) - cos_t) < 1e-5
            rel = eta if wi[2] > 0 else 1 / eta
            assert np.allclose(t[:2], -wi[:2].astype(np.float64) / rel, atol=1e-5)
            wo_r, _, wt_r, _ = _bsdf_sample(lib, BSDF_DIELECTRIC, one, optics, wi, u[:64], radiance=True)
            tr = ~np.all(wo_r == np.array([-wi[0], -wi[1], wi[2]], np.float32), axis=1)
            assert np.allclose(wt[~refl][0], 1.0) and (not tr.any() or np.allclose(wt_r[tr], 1 / (rel * rel), rtol=1e-5))


def test_conductor_bsdf_reflects_with_the_exact_fresnel_term(orc):
    """the smooth `conductor` (conductor.cpp:254-283): one discrete direction, weight = specularReflectance x fresnelConductorExact
    per channel (util.cpp:739-761) -- against the closed form written out independently here"""
    lib = orc.api().lib
    eta, k = np.array([0.27, 0.68, 1.22]), np.array([3.61, 2.63, 2.29])
    optics = np.zeros(12, np.float32); optics[0:3] = eta; optics[3:6] = k; optics[6:12] = 1.0
    u = np.random.default_rng(1).random((8, 2), dtype=np.float32)
    for cz in (1.0, 0.8, 0.3, 0.05):
        wi = np.array([np.sqrt(1 - cz * cz), 0.0, cz], np.float32)
        wo, pdf, wt, delta = _bsdf_sample(lib, BSDF_CONDUCTOR, np.ones(3, np.float32), optics, wi, u)
        assert delta.all() and (pdf == 1).all() and (wo == np.array([-wi[0], 0.0, wi[2]], np.float32)).all()
        c2 = float(wi[2]) ** 2; s2 = 1 - c2; s4 = s2 * s2
        t = eta * eta - k * k - s2
        a2b2 = np.sqrt(t * t + 4 * eta * eta * k * k)
        a = np.sqrt(0.5 * (a2b2 + t))
        rs = (a2b2 + c2 - 2 * a * np.sqrt(c2)) / (a2b2 + c2 + 2 * a * np.sqrt(c2))
        rp = rs * (c2 * a2b2 + s4 - 2 * a * np.sqrt(c2) * s2) / (c2 * a2b2 + s4 + 2 * a * np.sqrt(c2) * s2)
        assert np.allclose(wt[0], 0.5 * (rs + rp), rtol=2e-5), (cz, wt[0], 0.5 * (rs + rp))
    wo, pdf, wt, delta = _bsdf_sample(lib, BSDF_CONDUCTOR, np.ones(3, np.float32), optics, np.array([0.6, 0, -0.8], np.float32), u)
    assert not wt.any()                                                            # from behind: no sample (conductor.cpp:262-263)
