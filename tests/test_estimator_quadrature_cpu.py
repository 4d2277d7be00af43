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


# ---- grid medium: the Simpson march against brute-force integration of the trilinear density ------------------------------

def _trilinear(density, bmin, bmax, p):
    """GridDataSource::lookupFloat written independently (gridvolume.cpp:188-196, 337-373): grid point i sits at
    bmin + i * extent / (res - 1), x fastest; zero outside the cells"""
    res = np.array(density.shape[::-1])                                       # density is [z][y][x]
    gp = (p - bmin) / (bmax - bmin) * (res - 1)
    i0 = np.floor(gp).astype(np.int64)
    inside = ((i0 >= 0) & (i0 + 1 < res)).all(1)
    i0 = np.clip(i0, 0, res - 2)
    f = gp - i0
    out = np.zeros(len(p))
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                w = (f[:, 0] if dx else 1 - f[:, 0]) * (f[:, 1] if dy else 1 - f[:, 1]) * (f[:, 2] if dz else 1 - f[:, 2])
                out += w * density[i0[:, 2] + dz, i0[:, 1] + dy, i0[:, 0] + dx]
    return np.where(inside, out, 0.0)


def test_grid_medium_transmittance_against_brute_force_integration(pkg, orc):
    """HeterogeneousMedium::evalTransmittance, method = simpson (heterogeneous.cpp:301-376, 665-691): exp(-scale x the line
    integral of the trilinearly interpolated density).  The march steps at most half a voxel (gridvolume.cpp:197-199) with
    Simpson weights; 20 000 midpoint samples of an independently written trilinear lookup give the same optical depth to a few
    1e-4 -- on segments inside the grid, crossing its boundary, and missing it."""
    scene, vrls, params = small_case(pkg, "C3", 16, 16, 8, grid=20)
    m = scene["medium"]
    bmin, bmax = np.array([0.1, 0.55, 0.2]), np.array([0.9, 0.9, 1.0])            # the grid's box, inside the Cornell box
    o = orc.Oracle(**params); o.set_scene(scene)
    o.set_medium_grid(m["density"], bmin, bmax, m["scale"], m["albedo"], m["sigmaS_base"])
    rng = np.random.default_rng(12)
    n = 300
    lo, hi = np.array([0.02, 0.62, 0.02]), np.array([0.98, 0.98, 0.98])           # free space above the two boxes (no occluder between)
    p1 = rng.uniform(lo, hi, (n, 3)); p2 = rng.uniform(lo, hi, (n, 3))
    p1[:20, 0] = rng.uniform(0.02, 0.05, 20); p2[:20, 0] = rng.uniform(0.05, 0.09, 20)                  # beside the grid's box
    p1f, p2f = p1.astype(np.float32), p2.astype(np.float32)
    got = o.eval_transmittance(p1f, np.zeros(n, np.int32), p2f)[:, 0].astype(np.float64)
    steps = 20000
    t = (np.arange(steps) + 0.5) / steps
    dens = m["density"].astype(np.float64)
    want = np.zeros(n)
    for k in range(n):
        a, b = p1f[k].astype(np.float64), p2f[k].astype(np.float64)
        pts = a[None, :] + t[:, None] * (b - a)[None, :]
        want[k] = np.exp(-float(m["scale"]) * _trilinear(dens, bmin, bmax, pts).mean() * np.linalg.norm(b - a))
    assert (got[:20] == 1.0).all() and np.allclose(want[:20], 1.0)                                        # no density there
    assert 1e-4 < want.min() and (want < 0.9).sum() > 100                                                 # the segments see real optical depth
    depth_err = np.abs(np.log(got) - np.log(want))
    inside = ((p1f > bmin) & (p1f < bmax) & (p2f > bmin) & (p2f < bmax)).all(1)
    assert inside.sum() > 40 and (~inside).sum() > 100
    print("optical-depth error, segments inside the grid's box: max %.2e median %.2e; crossing its faces: max %.2e median %.2e"
          % (depth_err[inside].max(), np.median(depth_err[inside]), depth_err[~inside].max(), np.median(depth_err[~inside])))
    assert depth_err[inside].max() < 2e-3 and np.median(depth_err[inside]) < 3e-4
    # where the segment is clipped against the box the march's end sample sits ON a face, and lookupFloat reads zero there
    # (x2 >= res on the far faces, a floor of -1 by rounding on the near ones: gridvolume.cpp:344-346): the composite rule loses
    # the end sample's weight, density x scale x step / 3 per crossing.  The reference's behaviour, reproduced -- and bounded here.
    step = 0.5 * ((bmax - bmin) / (np.array(dens.shape[::-1]) - 1)).min()
    assert depth_err[~inside].max() < 2 * float(m["scale"]) * dens.max() * step / 3 + 2e-3
    assert (np.log(got[~inside]) >= np.log(want[~inside]) - 2e-3).all()                               # never denser than the truth


# ---- the walks' free-flight sampling (VRL tracer and ground truth) against the transmittance it has to follow ---------------

def _sample_distance(orc, o, origin, direction, its_t, u):
    import ctypes as C
    n = len(u)
    t, ok = np.zeros(n, np.float32), np.zeros(n, np.uint8)
    ps, pf, tr = np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros((n, 3), np.float32)
    po, pd = np.ascontiguousarray(origin, np.float32), np.ascontiguousarray(direction, np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert orc.api().lib.orc_test_sample_distance(o.h, p(po), p(pd), C.c_float(its_t), p(u), C.c_uint32(n), p(t), p(ok), p(ps), p(pf), p(tr)) == 0
    return t.astype(np.float64), ok.astype(bool), ps.astype(np.float64), pf.astype(np.float64), tr.astype(np.float64)


def test_homogeneous_free_flight_sampling_follows_its_density(pkg, orc):
    """HomogeneousMedium::sampleDistance, strategy balance (homogeneous.cpp:275-352): with probability mediumSamplingWeight a
    channel is picked uniformly and an exponential free flight drawn.  The distances of the sampled interactions follow
    sw / 3 * sum_c sigma_t,c exp(-sigma_t,c t) (Kolmogorov-Smirnov), the share that reaches the surface is pdfFailure, and
    pdfSuccess / pdfFailure / transmittance are what the record reports"""
    from scipy import stats
    o, scene, hitp, hitn, prim = _make(pkg, orc, 2, 2, 0.0, True)
    sig_t, sw = SIGMA_S + SIGMA_A, _sampling_weight()
    its_t = 1.3
    u = np.random.default_rng(5).random((200_000, 2), dtype=np.float32)
    t, ok, ps, pf, tr = _sample_distance(orc, o, (0.1, 0.2, 0.3), (0.6, 0.0, 0.8), its_t, u)
    p_fail = 1 - sw + sw * np.exp(-sig_t * its_t).mean()
    assert abs((~ok).mean() - p_fail) < 4.5 * np.sqrt(p_fail * (1 - p_fail) / len(u))
    assert np.allclose(pf[~ok], p_fail, rtol=1e-5) and (t[~ok] == np.float32(its_t)).all()
    cdf = lambda x: (1 - np.exp(-sig_t[None, :] * np.asarray(x)[:, None]).mean(1)) / (1 - np.exp(-sig_t * its_t).mean())
    assert stats.kstest(t[ok], cdf).pvalue > 1e-3
    assert np.allclose(ps[ok], sw * (sig_t[None, :] * np.exp(-sig_t[None, :] * t[ok, None])).mean(1), rtol=2e-5)
    assert np.allclose(tr[ok], np.exp(-sig_t[None, :] * t[ok, None]), rtol=2e-5)


def test_grid_medium_free_flight_sampling_follows_the_optical_depth(pkg, orc):
    """HeterogeneousMedium::sampleDistance, method simpson (heterogeneous.cpp:589-616 over invertDensityIntegral, 422-545): the
    march is inverted for the optical depth -log(1 - u).  Along three rays through the grid the sampled distances follow
    1 - exp(-tau(t)) with tau integrated by brute force from the independently written trilinear lookup, the share of rays
    that leave without an interaction is exp(-tau(end)), and pdfSuccess = sigma_t(t) exp(-tau(t))"""
    from scipy import stats
    scene, vrls, params = small_case(pkg, "C3", 16, 16, 8, grid=20)
    m = scene["medium"]
    bmin, bmax = np.array([0.1, 0.1, 0.1]), np.array([0.9, 0.9, 0.9])
    o = orc.Oracle(**params); o.set_scene(scene)
    o.set_medium_grid(m["density"], bmin, bmax, m["scale"], m["albedo"], m["sigmaS_base"])
    dens, scale = m["density"].astype(np.float64), float(m["scale"])
    rng = np.random.default_rng(6)
    for origin, target in (((0.3, 0.4, 0.2), (0.7, 0.5, 0.8)), ((0.02, 0.5, 0.5), (0.98, 0.45, 0.55)), ((0.5, 0.85, 0.15), (0.45, 0.2, 0.7))):
        origin, target = np.float32(origin), np.float32(target)
        d = (target - origin).astype(np.float64); its_t = float(np.linalg.norm(d)); d /= its_t
        d32 = d.astype(np.float32)
        steps = 40000
        edges = np.linspace(0, its_t, steps + 1)
        mid = 0.5 * (edges[1:] + edges[:-1])
        sigma = scale * _trilinear(dens, bmin, bmax, origin.astype(np.float64)[None, :] + mid[:, None] * d32.astype(np.float64)[None, :])
        tau = np.concatenate([[0.0], np.cumsum(sigma * (its_t / steps))])
        u = rng.random((60_000, 2), dtype=np.float32)
        t, ok, ps, pf, tr = _sample_distance(orc, o, origin, d32, np.float32(its_t), u)
        p_fail = np.exp(-tau[-1])
        assert 0.01 < p_fail < 0.9
        noise = 4.5 * np.sqrt(p_fail * (1 - p_fail) / len(u))
        crosses = not (((origin > bmin) & (origin < bmax)).all() and ((target > bmin) & (target < bmax)).all())
        if not crosses:
            assert abs((~ok).mean() - p_fail) < noise + 2e-3                                          # (+ the march's own error, see above)
            cdf = lambda x: (1 - np.exp(-np.interp(x, edges, tau))) / (1 - p_fail)
            ks = stats.kstest(t[ok], cdf)
            assert ks.statistic < 6e-3, ks                                                            # 60 000 samples: noise ~ 4e-3
            want_ps = np.interp(t[ok], mid, sigma) * np.exp(-np.interp(t[ok], edges, tau))
            assert np.median(np.abs(ps[ok] - want_ps) / np.maximum(want_ps, 1e-12)) < 2e-3
        else:
            # the end-sample-on-a-face quirk (see the transmittance test above) in the inverted march: its Simpson panels are two
            # steps wide, the samples on the two faces read zero, so up to scale x density x step / 3 of optical depth is lost per
            # face and more rays leave without an interaction than the medium's true transmittance allows -- bounded, one-sided
            step = 0.5 * ((bmax - bmin) / (np.array(dens.shape[::-1]) - 1)).min()
            lost = np.log((~ok).mean() / p_fail)
            assert -noise / p_fail - 2e-3 < lost < 2 * scale * dens.max() * step / 3 + noise / p_fail, lost
