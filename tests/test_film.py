"""Film: reconstruction-filter splat + pass accumulation + development.

CPU part: the oracle's restatement of ReconstructionFilter::configure / ImageBlock::put / the weight division against the
properties the reference's construction implies.  GPU part: alvrl_film_* against the oracle, bit for bit."""
import numpy as np
import pytest

BOX, TENT, GAUSS = 0, 1, 2


def _frames(H=37, W=53, n=2, seed=5):
    rng = np.random.default_rng(seed)
    fr = rng.random((n, H, W, 3), dtype=np.float32) * 3.0
    fr[0, 3, 4, 1] = np.nan          # invalid samples are rejected by ImageBlock::put (imageblock.h:147-151)
    fr[0, 10, 20, 0] = -1.0
    fr[-1, H - 1, W - 1, 2] = np.inf
    return fr


def test_oracle_film_box_is_identity_for_valid_samples(orc):
    """box filter, radius 0.5: a sample at a pixel centre reaches its own pixel only, weight cancels in the division"""
    fr = _frames(n=1)
    out = orc.film(fr, BOX)
    ok = np.isfinite(fr[0]).all(-1) & (fr[0] >= 0).all(-1)
    assert np.allclose(out[ok], fr[0][ok], rtol=1e-6, atol=0)
    assert (out[~ok] == 0).all()      # a rejected sample leaves weight 0 -> developed value 0 (bitmap.cpp:1620)


@pytest.mark.parametrize("filt", [TENT, GAUSS])
def test_oracle_film_constant_image_and_pass_average(orc, filt):
    const = np.full((1, 20, 30, 3), 0.7, np.float32)
    out = orc.film(const, filt)
    assert np.allclose(out, 0.7, rtol=2e-6)                               # normalised by the weight channel, edges included
    a, b = _frames(n=1, seed=1), _frames(n=1, seed=2)
    a, b = np.nan_to_num(np.abs(a), posinf=1.0), np.nan_to_num(np.abs(b), posinf=1.0)
    both = orc.film(np.concatenate([a, b]), filt)
    assert np.allclose(both, 0.5 * (orc.film(a, filt) + orc.film(b, filt)), rtol=1e-5, atol=1e-6)   # passes accumulate


def test_oracle_film_gaussian_footprint(orc):
    """gaussian, stddev 0.5: radius 2, but an offset of 2 pixels looks up the last table entry, which configure() sets to 0
    (rfilter.cpp:48; evalDiscretized, rfilter.h:76-77): a sample at a pixel centre reaches a 3 x 3 neighbourhood"""
    img = np.zeros((1, 15, 15, 3), np.float32)
    img[0, 7, 7] = 1.0
    out = orc.film(img, GAUSS)
    nz = np.argwhere(out[..., 0] > 0)
    assert nz.min() == 6 and nz.max() == 8
    assert out[7, 7, 0] > out[7, 8, 0] > out[8, 8, 0] > 0 and out[7, 9, 0] == 0
    assert np.allclose(out[7, 8], out[8, 7]) and np.allclose(out[6, 6], out[8, 8])


def _numpy_film(frames, radius, eval_fn):
    """ImageBlock::put for samples at pixel centres + the weight division, written independently: the offsets between a sample and
    the pixels it reaches are integers, so the pre-rasterised filter (31 entries over the radius, the 32nd zero; rfilter.cpp:37-55,
    rfilter.h:76-77) is looked up at |k| * 31 / radius only, and the film is a separable correlation of the valid samples with those
    weights divided by the same correlation of the valid mask (imageblock.h:144-185, bitmap.cpp:1617-1624)"""
    R = int(np.floor(radius))
    table = np.array([eval_fn(radius * i / 31.0) for i in range(31)] + [0.0])
    w = np.array([table[min(int(abs(k) * (31.0 / radius)), 31)] for k in range(-R, R + 1)])
    n, H, W, _ = frames.shape
    acc = np.zeros((H + 2 * R, W + 2 * R, 4))
    for f in frames.astype(np.float64):
        valid = (np.isfinite(f) & (f >= 0)).all(-1)
        val = np.where(valid[..., None], np.concatenate([np.nan_to_num(f, nan=0.0, posinf=0.0, neginf=0.0), np.ones((H, W, 1))], -1), 0.0)
        for dy in range(-R, R + 1):
            for dx in range(-R, R + 1):
                acc[R + dy:R + dy + H, R + dx:R + dx + W] += w[dy + R] * w[dx + R] * val
    inner = acc[R:R + H, R:R + W]
    return np.where(inner[..., 3:] > 0, inner[..., :3] / np.maximum(inner[..., 3:], 1e-300), 0.0)


@pytest.mark.parametrize("filt,param", [(GAUSS, 0.0), (GAUSS, 0.8), (TENT, 0.0)], ids=["gaussian-0.5", "gaussian-0.8", "tent"])
def test_oracle_film_equals_an_independent_numpy_film(orc, filt, param):
    fr = _frames(H=23, W=31, n=3, seed=9)
    if filt == GAUSS:
        sd = param or 0.5
        radius, ev = 4 * sd, (lambda x: max(0.0, np.exp(-x * x / (2 * sd * sd)) - np.exp(-(4 * sd) ** 2 / (2 * sd * sd))))   # gaussian.cpp:36-58
    else:
        radius, ev = 1.0, (lambda x: max(0.0, 1.0 - abs(x)))                                                                # tent.cpp:34-44
    want = _numpy_film(fr, radius, ev)
    got = orc.film(fr, filt, param)
    assert np.allclose(got, want, rtol=3e-6, atol=1e-7), float(np.abs(got - want).max())


@pytest.mark.gpu
@pytest.mark.parametrize("filt,param", [(BOX, 0.0), (TENT, 0.0), (GAUSS, 0.0), (GAUSS, 0.8)])
def test_film_matches_oracle_bit_exact(pkg, orc, filt, param):
    fr = _frames()
    n, H, W, _ = fr.shape
    scene, vrls, params = pkg.scenes.make_config("C1", width=W, height=H, n_vrls=8)
    g = pkg.integrator(0, **params)
    g.set_scene(scene)
    g.film_configure(filt, param)
    for k in range(n):
        g.film_put(fr[k])
    got = g.film_develop()
    want = orc.film(fr, filt, param)
    assert np.array_equal(got, want), float(np.abs(got - want).max())
    g.film_clear()
    g.film_put(fr[1])
    assert np.array_equal(g.film_develop(), orc.film(fr[1], filt, param))


@pytest.mark.gpu
def test_film_of_rendered_frame(pkg, orc):
    """alvrl_film_put(NULL) splats the frame alvrl_render left on the device: two passes of a progressive render"""
    scene, vrls, params = pkg.scenes.make_config("C1", width=48, height=40, n_vrls=64)
    params.update(targetNumSlices=6)
    g = pkg.integrator(0, **params)
    g.set_scene(scene); g.set_vrls(*vrls); g.build_slices(); g.prepass()
    g.film_configure(GAUSS)
    imgs = []
    for _ in range(2):
        imgs.append(g.render())
        g.film_put()
    assert np.array_equal(g.film_develop(), orc.film(np.stack(imgs), GAUSS))
