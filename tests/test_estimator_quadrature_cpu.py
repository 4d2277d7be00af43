"""The transport estimator against deterministic quadrature (CPU).

integrateVRL (src/integrators/vrl/vrlIntegrator.cpp:603-785) estimates, for one camera segment and one VRL, the radiance that
the VRL's flux sends to the eye after one more scattering event: a double line integral over the camera segment and the VRL
(volume to volume) plus a line integral over the VRL for the surface point the segment ends on (volume to surface).  It draws
the point on the VRL with the inverse-sinh warp of sampleVtoDistance (916-953) and the point on the camera segment
equi-angularly around it (889-914), and divides by the product of the two densities.

The reference has no test for any of this.  Here the integrand is written down independently in numpy from the
physics -- flux x sigma_s^2 x the three transmittances x the two phase functions / squared distance -- and integrated with
Gauss-Legendre rules; the mean of many estimator evaluations of the oracle (orc_integrate_pair, each on fresh uniforms) must
agree within its standard error.  This pins the sampling densities against the warps (an inconsistent pdf biases the mean) and
the integrand of rows a3 / a4 / a14 / a15 of SURVEY 8, per channel, for isotropic and HG phase functions, with the VRL's own
transmittance either carried in the integrand (shortVrls = false) or divided by the tracer's pdfFailure (shortVrls = true,
vrlIntegrator.cpp:655-656, homogeneous.cpp:354-396)."""
import numpy as np
import pytest

from conftest import small_case

SIGMA_S = np.array([1.0, 0.7, 0.4])
SIGMA_A = np.array([0.05, 0.1, 0.2])
VRLS = [((0.2, 0.3, 0.2), (0.8, 0.62, 0.5), (1.0, 0.8, 0.6)),           # crosses in front of the camera ray, 0.1 away at its closest
        ((0.9, 0.9, 0.9), (0.3, 0.8, -0.4), (0.5, 1.0, 2.0)),           # runs back towards the camera, beyond the medium's walls
        ((0.45, 0.1, 0.6), (0.47, 0.95, 0.62), (1.0, 1.0, 1.0))]         # almost perpendicular to the camera ray


def _phase(cos_theta, g):
    """isotropic (g = 0) / Henyey-Greenstein in the convention of hg.cpp:107-110: cos_theta = dot(wi, wo), forward = -1"""
    if g == 0:
        return np.full_like(cos_theta, 1 / (4 * np.pi))
    t = 1 + g * g + 2 * g * cos_theta
    return (1 - g * g) / (4 * np.pi * t * np.sqrt(t))


def _sampling_weight():
    """mediumSamplingWeight's default (homogeneous.cpp:165-176): the largest albedo, at least 0.5"""
    return max(float((SIGMA_S / (SIGMA_S + SIGMA_A)).max()), 0.5)


def _pdf_failure(dist):
    sw = _sampling_weight()
    return np.exp(-(SIGMA_S + SIGMA_A)[:, None] * dist[None, :]).mean(0) * sw + (1 - sw)


def _make(pkg, orc, nvv, nvs, g, short):
    scene, vrls, params = small_case(pkg, "C1", 16, 16, 8)
    scene = dict(scene, medium=dict(type="homogeneous", sigmaA=SIGMA_A.astype(np.float32), sigmaS=SIGMA_S.astype(np.float32),
                                    samplingWeight=-1.0, phase=1 if g else 0, g=g))
    params.update(volVolSamples=nvv, volSurfSamples=nvs, shortVrls=int(short), targetNumSlices=4)
    o = orc.Oracle(**params)
    o.set_scene(scene)
    s, e, p = (np.array([v[k] for v in VRLS], np.float32) for k in range(3))
    o.set_vrls(s, e, p, len(VRLS))
    o.set_no_visibility(True)                       # the medium's transmittance only: the quadrature has no occluders either
    prim, t, hitp, hitn = o.primary_hits()
    return o, scene, hitp, hitn, prim


def _centre_pixel(o, scene, prim):
    """a pixel whose camera segment ends on the back wall; returns (internal index, E, U_surf, n, albedo)"""
    W, H = scene["camera"]["width"], scene["camera"]["height"]
    idx = (H // 2) + H * (W // 2)                   # the oracle's pixel index is y + H * x (oracle_capi.cpp: tracePrimary)
    assert prim[idx] != 0xFFFFFFFF
    return idx


def _estimate(o, pixel, vrl, nvv, nvs, calls, seed):
    rng = np.random.default_rng(seed)
    out = np.zeros((calls, 3))
    for k in range(calls):
        out[k] = o.integrate_pair(pixel, vrl, rng.random(2 * nvv + nvs, dtype=np.float32))[2:5]
    return out.mean(0), out.std(0, ddof=1) / np.sqrt(calls)


def _gauss(n, a, b, panels):
    x, w = np.polynomial.legendre.leggauss(n)
    edges = np.linspace(a, b, panels + 1)
    mid, half = 0.5 * (edges[1:] + edges[:-1]), 0.5 * (edges[1:] - edges[:-1])
    return (mid[:, None] + half[:, None] * x[None, :]).reshape(-1), (half[:, None] * w[None, :]).reshape(-1)


def _vol_vol_quadrature(E, Us, S, End, power, g, short):
    Lc, Lv = np.linalg.norm(Us - E), np.linalg.norm(End - S)
    dc, dv = (Us - E) / Lc, (End - S) / Lv
    u, wu = _gauss(16, 0, Lc, 60)
    v, wv = _gauss(16, 0, Lv, 40)
    U = E[None, :] + u[:, None] * dc[None, :]                     # [nu, 3]
    V = S[None, :] + v[:, None] * dv[None, :]                     # [nv, 3]
    D = U[:, None, :] - V[None, :, :]                             # V -> U
    d = np.linalg.norm(D, axis=2)
    VU = D / d[:, :, None]
    cos_eye = VU @ dc                                             # light travels along VU, the eye looks along dc: wi = -VU, wo = -dc
    cos_vrl = -(VU @ dv)                                          # wi = -dv (where the flux comes from), wo = VU
    sig_t = SIGMA_S + SIGMA_A
    out = np.zeros(3)
    fail = _pdf_failure(v) if short else np.ones_like(v)
    for c in range(3):
        f = (power[c] * SIGMA_S[c] ** 2 * np.exp(-sig_t[c] * (u[:, None] + v[None, :] + d)) / fail[None, :]
             * _phase(cos_eye, g) * _phase(cos_vrl, g) / (d * d))
        out[c] = wu @ f @ wv
    return out


def _vol_surf_quadrature(E, Us, n, albedo, S, End, power, g, short):
    Lc, Lv = np.linalg.norm(Us - E), np.linalg.norm(End - S)
    dv = (End - S) / Lv
    v, wv = _gauss(16, 0, Lv, 60)
    V = S[None, :] + v[:, None] * dv[None, :]
    D = Us[None, :] - V
    d = np.linalg.norm(D, axis=1)
    VU = D / d[:, None]
    cos_vrl = -(VU @ dv)
    cos_wo = np.maximum(0.0, -(VU @ n))                            # the surface is lit from the side its normal points to
    sig_t = SIGMA_S + SIGMA_A
    fail = _pdf_failure(v) if short else np.ones_like(v)
    out = np.zeros(3)
    for c in range(3):
        f = power[c] * SIGMA_S[c] * np.exp(-sig_t[c] * (Lc + v + d)) / fail * _phase(cos_vrl, g) * albedo[c] / np.pi * cos_wo / (d * d)
        out[c] = wv @ f
    return out


@pytest.mark.parametrize("short", [False, True], ids=["long-vrls", "short-vrls"])
@pytest.mark.parametrize("g", [0.0, 0.8, -0.3], ids=["isotropic", "hg0.8", "hg-0.3"])
def test_volume_to_volume_estimator_is_unbiased(pkg, orc, g, short):
    nvv = 32
    o, scene, hitp, hitn, prim = _make(pkg, orc, nvv, 0, g, short)
    px = _centre_pixel(o, scene, prim)
    E, Us = scene["camera"]["origin"].astype(np.float64), hitp[px].astype(np.float64)
    for k, (s, e, p) in enumerate(VRLS):
        want = _vol_vol_quadrature(E, Us, np.array(s, np.float64), np.array(e, np.float64), np.array(p, np.float64), g, short)
        got, se = _estimate(o, px, k, nvv, 0, 1500 if g == 0 else 8000, seed=10 * k + int(short))
        assert (want > 0).all() and (se / want < 0.01).all(), (want, se)          # the test resolves 1 %
        assert (np.abs(got - want) < 4.5 * se + 2e-4 * want).all(), (k, got, want, se)


@pytest.mark.parametrize("short", [False, True], ids=["long-vrls", "short-vrls"])
@pytest.mark.parametrize("g", [0.0, 0.8], ids=["isotropic", "hg0.8"])
def test_volume_to_surface_estimator_is_unbiased(pkg, orc, g, short):
    nvs = 32
    o, scene, hitp, hitn, prim = _make(pkg, orc, 0, nvs, g, short)
    px = _centre_pixel(o, scene, prim)
    E, Us, n = scene["camera"]["origin"].astype(np.float64), hitp[px].astype(np.float64), hitn[px].astype(np.float64)
    albedo = scene["albedo"][scene["tri_material"][prim[px]]].astype(np.float64)
    for k, (s, e, p) in enumerate(VRLS):
        want = _vol_surf_quadrature(E, Us, n, albedo, np.array(s, np.float64), np.array(e, np.float64), np.array(p, np.float64), g, short)
        got, se = _estimate(o, px, k, 0, nvs, 1500 if g == 0 else 8000, seed=100 + 10 * k + int(short))
        assert (want > 0).all() and (se / want < 0.01).all(), (want, se)
        assert (np.abs(got - want) < 4.5 * se + 2e-4 * want).all(), (k, got, want, se)
