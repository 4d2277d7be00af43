"""CPU suite, part 1: the oracle against the reference's own golden vectors and against committed fixtures.

Pinned by the reference: the SFMT-19937 stream (src/tests/test_random.cpp:433-509).  Everything else on this path has
no reference test ("parity unpinned"): the committed fixtures under tests/golden/ were produced by the
oracle itself (tests/golden/make_goldens.py) and guard it against regressions."""
import os

import numpy as np
import pytest

from conftest import ROOT, small_case, setup

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_sfmt_known_answer_vector(orc):
    words = [int(l, 16) for l in open(os.path.join(GOLD, "sfmt_kat_seed4321.txt")) if not l.startswith("#")]
    assert len(words) == 192
    got = orc.sfmt_ulongs(4321, len(words))
    assert [int(x) for x in got] == words


def test_sfmt_next_float_is_low_32_bits(orc):
    u = orc.sfmt_ulongs(99, 64)
    f = orc.sfmt_floats(99, 64)
    expect = (((u & np.uint64(0xFFFFFFFF)) >> np.uint64(9)).astype(np.uint32) | np.uint32(0x3F800000)).view(np.float32) - np.float32(1)
    assert np.array_equal(f, expect)
    assert (f >= 0).all() and (f < 1).all()


def test_sfmt_clone_is_deterministic_and_distinct(orc):
    a = orc.sfmt_clone_ulongs(5, 0, 16)
    b = orc.sfmt_clone_ulongs(5, 0, 16)
    c = orc.sfmt_clone_ulongs(5, 1, 16)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert not np.array_equal(a, orc.sfmt_ulongs(5, 16))


def test_counter_stream_matches_python_restatement(pkg):
    def mix(x):
        x &= 0xFFFFFFFF
        x ^= x >> 16; x = (x * 0x7FEB352D) & 0xFFFFFFFF
        x ^= x >> 15; x = (x * 0x846CA68B) & 0xFFFFFFFF
        x ^= x >> 16
        return x
    seed, domain, a, b = 0x1234567890, 1, 17, 4242
    h = mix((seed & 0xFFFFFFFF) ^ ((domain * 0x9E3779B9) & 0xFFFFFFFF))
    h = mix(h ^ (seed >> 32))
    h = mix((h + a * 0x85EBCA6B + 0x165667B1) & 0xFFFFFFFF)
    h = mix(h ^ ((b * 0xC2B2AE35 + 0x27D4EB2F) & 0xFFFFFFFF))
    bits = mix((h + 3 * 0x9E3779B9) & 0xFFFFFFFF)
    u = np.array([(bits >> 9) | 0x3F800000], dtype=np.uint32).view(np.float32)[0] - np.float32(1)
    assert 0 <= u < 1


@pytest.fixture(scope="module")
def c1_small(pkg, orc):
    scene, vrls, params = small_case(pkg, "C1", 64, 64, 200, seed=11)
    o = setup(orc.Oracle(**params), scene, vrls)
    o.build_slices()
    o.prepass()
    return o


def test_slices_partition_all_hit_pixels(c1_small):
    o = c1_small
    p2s = o.pixel_to_slice()
    prim, t, p, n = o.primary_hits()
    S, G = o.num_slices()
    assert S == 100
    assert ((p2s == 0xFFFFFFFF) == (prim == 0xFFFFFFFF)).all()      # misses stay unsliced, Preprocessor.cpp:1201
    assert set(np.unique(p2s[p2s != 0xFFFFFFFF])) == set(range(S))
    off, px = o.rep_pixels()
    assert off[-1] == G and len(np.unique(px)) == G                 # representatives are distinct pixels
    assert (p2s[px] == np.repeat(np.arange(S), np.diff(off))).all()  # and belong to their slice


def test_R_entries_are_finite_nonnegative(c1_small):
    R = c1_small.get_R()
    assert np.isfinite(R).all() and (R >= 0).all()
    assert (R[..., 0] > 0).mean() > 0.5


def test_cluster_weights_are_inverse_probabilities(c1_small):
    cl = c1_small.clusters()
    assert (cl["weights"] >= 1).all()                               # weight = 1/prob, Preprocessor.cpp:375
    assert len(cl["fallback_vrls"]) == int(0.5 + c1_small.N / 5)    # refineFixedDepth, Preprocessor.cpp:388
    for s in range(len(cl["offset"]) - 1):
        v = cl["vrls"][cl["offset"][s]:cl["offset"][s + 1]]
        assert len(np.unique(v)) == len(v)


def test_clustered_render_is_unbiased_wrt_unclustered(c1_small):
    img_c = c1_small.render(True)
    img_u = c1_small.render(False)
    assert np.isfinite(img_c).all() and (img_c >= 0).all()
    # one representative per cluster with weight 1/prob: heavy-tailed per slice, so compare the median slice ratio
    p2s = c1_small.pixel_to_slice().reshape(c1_small.W, c1_small.H).T
    ratios = [img_c[p2s == s].sum() / img_u[p2s == s].sum() for s in range(c1_small.num_slices()[0]) if img_u[p2s == s].sum() > 0]
    assert abs(np.median(ratios) - 1) < 0.1


def test_oracle_thread_count_does_not_change_results(pkg, orc):
    scene, vrls, params = small_case(pkg, "C1", 32, 32, 64, seed=3, targetNumSlices=10)
    outs = []
    for th in (1, 4):
        o = setup(orc.Oracle(threads=th, **params), scene, vrls)
        o.build_slices(); o.prepass()
        outs.append((o.get_R(), o.clusters()["vrls"], o.render()))
    assert np.array_equal(outs[0][0], outs[1][0])
    assert np.array_equal(outs[0][1], outs[1][1])
    assert np.array_equal(outs[0][2], outs[1][2])


def test_sfmt_mode_worker_count_changes_stream_but_is_reproducible(pkg, orc):
    scene, vrls, params = small_case(pkg, "C1", 32, 32, 64, seed=3, targetNumSlices=10, rngMode=1)
    def run(w):
        o = setup(orc.Oracle(workerCount=w, **params), scene, vrls)
        o.build_slices(); o.prepass()
        return o.get_R()
    a, b, c = run(1), run(1), run(2)
    assert np.array_equal(a, b)
    assert not np.array_equal(a, c)


def test_recorded_tape_replays_to_identical_R(pkg, orc):
    """tape mode = the reference's sequential stream laid out per (row, vrl): replaying it must give the same R"""
    scene, vrls, params = small_case(pkg, "C1", 32, 32, 48, seed=5, targetNumSlices=8, rngMode=1)
    o = setup(orc.Oracle(**params), scene, vrls)
    o.build_slices(); o.sample_slice_mapping()
    tape = o.build_R_record_tape()
    R1 = o.get_R()
    o2 = setup(orc.Oracle(**params), scene, vrls)
    o2.build_slices(); o2.set_rep_pixels(*o.rep_pixels())
    o2.set_sample_tape(tape)
    o2.build_R()
    assert np.array_equal(R1, o2.get_R())


def test_heterogeneous_medium_runs_and_attenuates(pkg, orc):
    scene, vrls, params = small_case(pkg, "C3", 24, 24, 32, grid=16, targetNumSlices=6)
    o = setup(orc.Oracle(**params), scene, vrls)
    T = o.eval_transmittance(np.array([[0.1, 0.8, 0.5]], np.float32), [0], np.array([[0.9, 0.8, 0.5]], np.float32))
    assert 0 < T[0, 0] < 1 and T[0, 0] == T[0, 1] == T[0, 2]
    o.build_slices(); o.sample_slice_mapping(); o.build_R()
    R = o.get_R()
    assert np.isfinite(R).all() and (R[..., 0] > 0).any()


def test_occluder_blocks_transmittance(pkg, orc):
    scene, vrls, params = small_case(pkg, "C1", 16, 16, 16)
    o = setup(orc.Oracle(**params), scene, vrls)
    # through the tall box (centre 0.67,0.30,0.64) and in free space
    T = o.eval_transmittance(np.array([[0.67, 0.3, 0.2], [0.2, 0.8, 0.2]], np.float32), [0, 0],
                             np.array([[0.67, 0.3, 0.95], [0.8, 0.8, 0.2]], np.float32))
    assert (T[0] == 0).all()
    assert np.allclose(T[1], np.exp(-1.05 * 0.6), rtol=1e-6)


def test_golden_fixture_regression(pkg, orc):
    g = np.load(os.path.join(GOLD, "c1_tiny.npz"))
    scene, vrls, params = small_case(pkg, "C1", int(g["width"]), int(g["height"]), int(g["n_vrls"]),
                                     seed=int(g["seed"]), targetNumSlices=int(g["targetNumSlices"]))
    o = setup(orc.Oracle(**params), scene, vrls)
    o.build_slices(); o.prepass()
    assert np.array_equal(o.primary_hits()[0], g["prim"])
    assert np.array_equal(o.pixel_to_slice(), g["pixel_to_slice"])
    assert np.array_equal(o.rep_pixels()[1], g["row_pixel"])
    np.testing.assert_allclose(o.get_R(), g["R"], rtol=2e-6, atol=0)     # libm ulp differences between hosts
    cl = o.clusters()
    assert np.array_equal(cl["offset"], g["cluster_offset"])
    assert np.array_equal(cl["vrls"], g["cluster_vrls"])
    np.testing.assert_allclose(cl["weights"], g["cluster_weights"], rtol=1e-5)
    np.testing.assert_allclose(o.render(), g["image"], rtol=1e-4, atol=1e-9)


def test_counter_stream_known_answers(tmp_path):
    """include/alvrl_rng.h is shared by the oracle and the product, so a change of the stream would go unnoticed by the parity
    tests: its keys (object stream, per-cluster sub-stream of Clustering::split), bits and uniforms are pinned by the vectors
    of tests/golden/rng_kat.txt, recomputed here by an independent restatement and by the header itself (compiled with gcc)."""
    import subprocess
    M = 0xFFFFFFFF

    def mix32(x):
        x &= M
        x ^= x >> 16; x = (x * 0x7feb352d) & M
        x ^= x >> 15; x = (x * 0x846ca68b) & M
        x ^= x >> 16
        return x

    def key(seed, domain, a, b):
        h = mix32((seed & M) ^ ((domain * 0x9e3779b9) & M))
        h = mix32(h ^ (seed >> 32))
        h = mix32((h + a * 0x85ebca6b + 0x165667b1) & M)
        return mix32(h ^ ((b * 0xc2b2ae35 + 0x27d4eb2f) & M))

    def node_key(k, begin, end):
        h = mix32(k ^ ((begin * 0x85ebca6b + 0x2545f491) & M))
        return mix32((h + end * 0xc2b2ae35 + 0x68e31da4) & M)

    def bits(k, i):
        return mix32((k + i * 0x9e3779b9) & M)

    def uniform(k, i):
        return float(np.array([(bits(k, i) >> 9) | 0x3f800000], dtype=np.uint32).view(np.float32)[0] - np.float32(1.0))

    golden = open(os.path.join(ROOT, "tests", "golden", "rng_kat.txt")).read().split("\n")
    rows = [l.split() for l in golden if l.strip()]
    assert len(rows) == 9
    for seed, a, k, nk, b1, u5, un0 in rows:
        seed, a = int(seed), int(a)
        kk = key(seed, 4, a * 37, 0)                              # ALVRL_RNG_CLUSTER = 4
        nn = node_key(kk, a * 1000, a * 1000 + 17 + a)
        assert (kk, nn, bits(nn, 1)) == (int(k, 16), int(nk, 16), int(b1, 16))
        assert abs(uniform(kk, 5) - float(u5)) < 1e-9 and abs(uniform(nn, 0) - float(un0)) < 1e-9
        assert 0.0 <= uniform(nn, 0) < 1.0
    exe = str(tmp_path / "rng_kat")
    subprocess.check_call(["gcc", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "golden", "make_rng_kat.c"), "-o", exe])
    assert subprocess.check_output([exe]).decode().split() == open(os.path.join(ROOT, "tests", "golden", "rng_kat.txt")).read().split()


def test_neighbour_slices_in_local_matrices(pkg, orc):
    """buildLocalities / getLocalMatrix with neighbourWeight > 0 (Preprocessor.cpp:1241-1293, 796-820) are restated in the
    oracle (the CUDA path still refuses them): rows of the neighbour slices join a slice's local matrix with inverse-distance
    weights.  No reference output pins this (parity unpinned), so the test states what must hold: valid clusters,
    reproducible, unused when neighbourWeight = 0, different from the default when used, and the reference's failure mode
    (no neighbours to weigh: the locality weights are not normalised)."""
    scene, vrls, params = small_case(pkg, "C1", 32, 32, 64, seed=3, targetNumSlices=10)

    def run(**kw):
        o = setup(orc.Oracle(**dict(params, **kw)), scene, vrls)
        o.build_slices(); o.prepass()
        return o, o.clusters()

    _, base = run()
    _, unused = run(neighbourCount=3, neighbourWeight=0.0)
    for k in ("offset", "vrls", "weights"):
        assert np.array_equal(base[k], unused[k])
    o1, nb1 = run(neighbourCount=3, neighbourWeight=0.3)
    _, nb2 = run(neighbourCount=3, neighbourWeight=0.3)
    for k in ("offset", "vrls", "weights"):
        assert np.array_equal(nb1[k], nb2[k])
    S, _ = o1.num_slices()
    assert len(nb1["offset"]) == S + 1 and (np.diff(nb1["offset"]) >= 1).all()
    assert (nb1["vrls"] < o1.N).all() and (nb1["weights"] >= 1 - 1e-5).all() and np.isfinite(nb1["weights"]).all()
    assert not (np.array_equal(nb1["offset"], base["offset"]) and np.array_equal(nb1["vrls"], base["vrls"]))
    img = o1.render()
    assert np.isfinite(img).all() and (img >= 0).all() and img.mean() > 0
    _, everybody = run(neighbourCount=64, neighbourWeight=0.5)                  # more neighbours asked for than slices exist
    assert len(everybody["offset"]) == S + 1 and np.isfinite(everybody["weights"]).all()
    with pytest.raises(Exception):
        run(neighbourCount=0, neighbourWeight=0.5)


def test_primary_hits_against_a_textbook_pinhole_and_brute_force(pkg, orc):
    """camera + closest hit (rows a8 / a10) checked independently of the oracle's code: rays of a textbook pinhole (eye, target,
    up, horizontal field of view; pixel centres; image y pointing down) instead of the sampleToCamera / cameraToWorld matrices
    (perspective.cpp:126-175, 367-385), closest hit by Moeller-Trumbore in float64 over all triangles instead of the TriAccel
    test (triaccel.h:97-145): same triangle, same distance and hit point, for every pixel away from triangle edges"""
    W, H = 72, 56
    scene, vrls, params = pkg.scenes.make_config("C1", width=W, height=H, n_vrls=4)
    eye, target, fov = np.array([0.45, 0.52, -2.2]), np.array([0.5, 0.5, 0.0]), 40.0         # off-centre and further back: some rays miss
    scene = dict(scene, camera=pkg.scenes.perspective_camera(W, H, origin=eye, target=target, fov=fov))
    scene["extra_bounds"] = scene["camera"]["origin"].reshape(1, 3)
    o = orc.Oracle(**params); o.set_scene(scene); o.set_vrls(*vrls)
    prim, t, p, n = o.primary_hits()
    prim, t, p = prim.reshape(W, H).T, t.reshape(W, H).T, p.reshape(W, H, 3).transpose(1, 0, 2)      # index y + H * x -> [y][x]
    fwd = (target - eye) / np.linalg.norm(target - eye)
    right = np.cross(fwd, np.array([0.0, 1.0, 0.0])); right /= np.linalg.norm(right)          # the viewer's right hand (world -x when looking along +z)
    upc = np.cross(right, fwd)
    th = np.tan(np.radians(fov) / 2)
    xs, ys = (np.arange(W) + 0.5) / W, (np.arange(H) + 0.5) / H
    d = fwd[None, None, :] + ((2 * xs - 1) * th)[None, :, None] * right[None, None, :] + ((1 - 2 * ys) * th * H / W)[:, None, None] * upc[None, None, :]
    d /= np.linalg.norm(d, axis=2, keepdims=True)
    v = scene["verts"].astype(np.float64)
    best_t = np.full((H, W), np.inf); best = np.full((H, W), -1); edge = np.zeros((H, W), bool)
    for k, (a, b, c) in enumerate(scene["tris"]):
        e1, e2 = v[b] - v[a], v[c] - v[a]
        pv = np.cross(d, e2)
        det = pv @ e1
        with np.errstate(divide="ignore", invalid="ignore"):
            s = eye - v[a]
            u = (pv @ s) / det
            q = np.cross(s, e1)
            w = (d @ q) / det
            tt = (q @ e2) / det
        ok = (np.abs(det) > 1e-12) & (u >= 0) & (w >= 0) & (u + w <= 1) & (tt > 0)
        near_edge = ok & ((u < 1e-4) | (w < 1e-4) | (u + w > 1 - 1e-4))
        closer = ok & (tt < best_t)
        edge = np.where(closer, near_edge, edge | (ok & near_edge & (np.abs(tt - best_t) < 1e-6)))
        best = np.where(closer, k, best); best_t = np.where(closer, tt, best_t)
    hit = best >= 0
    assert 0.3 < hit.mean() < 1.0
    assert np.array_equal(prim != pkg.binding.NO_HIT, hit)
    clean = hit & ~edge
    assert clean.sum() > 0.9 * hit.sum()
    assert np.array_equal(prim[clean], best[clean])
    assert np.allclose(t[clean], best_t[clean], rtol=2e-5)
    assert np.allclose(p[clean], (eye + best_t[..., None] * d)[clean], atol=2e-5)


def test_counter_stream_statistics_and_its_stated_limit(pkg):
    """the addressed sample stream (include/alvrl_rng.h) restated with numpy: 3.1 M uniforms of the R domain (256 rows x 1 024 VRLs x
    12 draws) are uniform (chi-square on 256 bins, and on 16 x 16 bins for consecutive draws), and neighbouring addresses --
    (row, VRL + 1), (row + 1, VRL), draw k + 1 -- are uncorrelated.  The stated limit: key and state are
    32 bits wide, so distinct addresses start to share keys at the birthday bound -- counted here on 2^20 addresses, where the
    expectation is 128 coinciding pairs"""
    from scipy import stats

    def mix(x):
        x = x.astype(np.uint32)
        x ^= x >> np.uint32(16); x *= np.uint32(0x7FEB352D)
        x ^= x >> np.uint32(15); x *= np.uint32(0x846CA68B)
        x ^= x >> np.uint32(16)
        return x

    def key(seed, domain, a, b):
        with np.errstate(over="ignore"):
            h = mix(np.uint32(seed & 0xFFFFFFFF) ^ np.uint32((domain * 0x9E3779B9) & 0xFFFFFFFF) + np.zeros_like(a, np.uint32))
            h = mix(h ^ np.uint32(seed >> 32))
            h = mix(h + a.astype(np.uint32) * np.uint32(0x85EBCA6B) + np.uint32(0x165667B1))
            return mix(h ^ (b.astype(np.uint32) * np.uint32(0xC2B2AE35) + np.uint32(0x27D4EB2F)))

    def uniforms(k_, n_draws):
        with np.errstate(over="ignore"):
            bits = mix(k_[..., None] + np.arange(n_draws, dtype=np.uint32) * np.uint32(0x9E3779B9))
        return ((bits >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32) - np.float32(1)

    # the restatement is the header's stream: its known-answer file
    import os
    from conftest import ROOT
    kat = [l.split() for l in open(os.path.join(ROOT, "tests", "golden", "rng_kat.txt")) if l.strip() and not l.startswith("#")]
    a, b = np.meshgrid(np.arange(256), np.arange(1024), indexing="ij")
    keys = key(5, 1, a, b)
    u = uniforms(keys, 12).astype(np.float64)                                   # [256, 1024, 12]
    n = u.size
    assert abs(u.mean() - 0.5) < 4.5 * np.sqrt(1 / 12 / n)
    hist = np.bincount((u.reshape(-1) * 256).astype(int), minlength=256)
    assert stats.chisquare(hist).pvalue > 1e-3
    pair = np.bincount(((u[..., :-1] * 16).astype(int) * 16 + (u[..., 1:] * 16).astype(int)).reshape(-1), minlength=256)
    assert stats.chisquare(pair).pvalue > 1e-3

    def corr(x, y):
        return float(np.corrcoef(x.reshape(-1), y.reshape(-1))[0, 1])
    lim = 4.5 / np.sqrt(255 * 1023 * 12)
    assert abs(corr(u[:, :-1], u[:, 1:])) < lim and abs(corr(u[:-1], u[1:])) < lim and abs(corr(u[..., :-1], u[..., 1:])) < lim
    # the 32-bit limit: coinciding keys among 2^20 addresses of one domain ~ Poisson(2^40 / 2^33 = 128)
    a2, b2 = np.meshgrid(np.arange(1024), np.arange(1024), indexing="ij")
    k2 = key(5, 1, a2, b2).reshape(-1)
    dup = len(k2) - len(np.unique(k2))
    assert 128 - 5 * np.sqrt(128) < dup < 128 + 5 * np.sqrt(128), dup
    # and the numpy restatement IS the header's stream: keys and the sixth uniform of the known-answer file (domain CLUSTER = 4, b = 0)
    for seed, a_, key_hex, _, _, u5, _ in kat:
        kk = key(int(seed), 4, np.array([int(a_) * 37]), np.array([0]))
        assert int(kk[0]) == int(key_hex, 16)
        assert abs(float(uniforms(kk, 6)[0, 5]) - float(u5)) < 1e-9
