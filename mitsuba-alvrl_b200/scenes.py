"""Procedural scenes of the BASELINE.json configs.  numpy only; everything is seeded.

The reference ships no example scenes for the vrl integrator (README.md:33-36), so the named shapes are generated:
  C1  cornell 256x256,  1k VRLs, homogeneous isotropic medium, Nvv = Nvs = 2
  C2  cornell 1024x1024, 100k VRLs, Nvv = Nvs = 4
  C3  cornell + 512^3 procedural density grid (method=simpson), 1024x1024, 200k VRLs
  C4  ~1M-triangle icosphere occluders in fog inside the closed box, 1920x1080, 300k VRLs
  C5  3840x2160, 1M VRLs, HG g = 0.8, Nvv = 16, Nvs = 4, targetNumSlices = 512
Matrices follow Mitsuba's conventions (src/sensors/perspective.cpp:150-155, src/libcore/transform.cpp:99-123,
191-214); they are *inputs* of both the CUDA path and the oracle.
"""
import math
import numpy as np

BSDF_SMOOTH = 1
BSDF_DIELECTRIC, BSDF_CONDUCTOR = 2, 4          # delta BSDFs of the specular chains (include/alvrl.h)
MAT_TRANSITION, MAT_INTERIOR_MEDIUM, MAT_EXTERIOR_MEDIUM = 8, 16, 32


# ---- camera ------------------------------------------------------------------------------------
def _perspective(fov_deg, near, far):
    recip = 1.0 / (far - near)
    cot = 1.0 / math.tan(math.radians(fov_deg / 2.0))
    return np.array([[cot, 0, 0, 0], [0, cot, 0, 0], [0, 0, far * recip, -near * far * recip], [0, 0, 1, 0]], dtype=np.float64)


def _scale(x, y, z):
    return np.diag([x, y, z, 1.0])


def _translate(x, y, z):
    m = np.eye(4)
    m[:3, 3] = [x, y, z]
    return m


def look_at(p, t, up):
    p, t, up = (np.asarray(a, dtype=np.float64) for a in (p, t, up))
    d = (t - p) / np.linalg.norm(t - p)
    left = np.cross(up, d)
    left /= np.linalg.norm(left)
    new_up = np.cross(d, left)
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = left, new_up, d, p
    return m


def perspective_camera(width, height, origin, target, up=(0, 1, 0), fov=40.0, near=1e-2, far=1e4):
    aspect = width / float(height)
    cam_to_sample = (_scale(-0.5, -0.5 * aspect, 1.0) @ _translate(-1.0, -1.0 / aspect, 0.0)
                     @ _perspective(fov, near, far))
    return dict(sampleToCamera=np.linalg.inv(cam_to_sample).astype(np.float32),
                cameraToWorld=look_at(origin, target, up).astype(np.float32),
                width=width, height=height, near=np.float32(near), far=np.float32(far),
                origin=np.asarray(origin, dtype=np.float32))


# ---- geometry ----------------------------------------------------------------------------------
class _Mesh:
    def __init__(self):
        self.v, self.t, self.m = [], [], []

    def quad(self, a, b, c, d, mat):
        """two triangles (a,b,c),(a,c,d); the front face is the side from which a,b,c,d appear counter-clockwise"""
        i = len(self.v)
        self.v += [a, b, c, d]
        self.t += [(i, i + 1, i + 2), (i, i + 2, i + 3)]
        self.m += [mat, mat]

    def box(self, centre, half, angle_deg, mat):
        cx, cy, cz = centre
        hx, hy, hz = half
        ca, sa = math.cos(math.radians(angle_deg)), math.sin(math.radians(angle_deg))

        def P(x, y, z):
            return (cx + ca * x * hx + sa * z * hz, cy + y * hy, cz - sa * x * hx + ca * z * hz)
        c = {(x, y, z): P(x, y, z) for x in (-1, 1) for y in (-1, 1) for z in (-1, 1)}
        # outward-facing quads
        self.quad(c[(-1, -1, -1)], c[(-1, 1, -1)], c[(1, 1, -1)], c[(1, -1, -1)], mat)  # -z
        self.quad(c[(-1, -1, 1)], c[(1, -1, 1)], c[(1, 1, 1)], c[(-1, 1, 1)], mat)      # +z
        self.quad(c[(-1, -1, -1)], c[(-1, -1, 1)], c[(-1, 1, 1)], c[(-1, 1, -1)], mat)  # -x
        self.quad(c[(1, -1, -1)], c[(1, 1, -1)], c[(1, 1, 1)], c[(1, -1, 1)], mat)      # +x
        self.quad(c[(-1, 1, -1)], c[(-1, 1, 1)], c[(1, 1, 1)], c[(1, 1, -1)], mat)      # +y
        self.quad(c[(-1, -1, -1)], c[(1, -1, -1)], c[(1, -1, 1)], c[(-1, -1, 1)], mat)  # -y

    def arrays(self):
        return (np.asarray(self.v, dtype=np.float32), np.asarray(self.t, dtype=np.uint32),
                np.asarray(self.m, dtype=np.uint32))


WHITE, RED, GREEN = 0, 1, 2
_ALBEDO = np.array([[.73, .73, .73], [.65, .05, .05], [.12, .45, .15]], dtype=np.float32)


def _cornell_mesh(closed=False):
    m = _Mesh()
    # walls of the unit cube, normals pointing inside
    m.quad((0, 0, 0), (0, 0, 1), (1, 0, 1), (1, 0, 0), WHITE)          # floor  (+y)
    m.quad((0, 1, 0), (1, 1, 0), (1, 1, 1), (0, 1, 1), WHITE)          # ceiling (-y)
    m.quad((0, 0, 1), (0, 1, 1), (1, 1, 1), (1, 0, 1), WHITE)          # back   (-z)
    m.quad((0, 0, 0), (0, 1, 0), (0, 1, 1), (0, 0, 1), RED)            # left   (+x)
    m.quad((1, 0, 0), (1, 0, 1), (1, 1, 1), (1, 1, 0), GREEN)          # right  (-x)
    if closed:
        m.quad((0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), WHITE)      # front  (+z)
    m.box((0.33, 0.15, 0.33), (0.15, 0.15, 0.15), 17.0, WHITE)         # short box
    m.box((0.67, 0.30, 0.64), (0.15, 0.30, 0.15), -18.0, WHITE)        # tall box
    return m


def _icosphere(subdiv):
    t = (1.0 + math.sqrt(5.0)) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    v = [tuple(np.asarray(p) / np.linalg.norm(p)) for p in v]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
         (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10),
         (8, 6, 7), (9, 8, 1)]
    for _ in range(subdiv):
        cache, nf = {}, []

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                p = (np.asarray(v[a]) + np.asarray(v[b])) / 2.0
                v.append(tuple(p / np.linalg.norm(p)))
                cache[key] = len(v) - 1
            return cache[key]
        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    return np.asarray(v, dtype=np.float64), np.asarray(f, dtype=np.int64)


def occluder_mesh(n_spheres=780, subdiv=3, seed=4):
    """C4: icospheres inside the closed Cornell walls (998 400 triangles at the defaults)."""
    base = _cornell_mesh(closed=True)
    v0, t0, m0 = base.arrays()
    sv, sf = _icosphere(subdiv)
    rng = np.random.default_rng(seed)
    centres = rng.uniform(0.05, 0.95, size=(n_spheres, 3))
    radii = rng.uniform(0.01, 0.04, size=n_spheres)
    verts = (centres[:, None, :] + radii[:, None, None] * sv[None, :, :]).reshape(-1, 3)
    faces = (sf[None, :, :] + (np.arange(n_spheres) * len(sv))[:, None, None]).reshape(-1, 3) + len(v0)
    return (np.concatenate([v0, verts.astype(np.float32)]), np.concatenate([t0, faces.astype(np.uint32)]),
            np.concatenate([m0, np.zeros(len(faces), np.uint32)]))


# ---- media -------------------------------------------------------------------------------------
def homogeneous_medium(sigma_s=1.0, sigma_a=0.05, phase=0, g=0.0):
    """docstring example coefficients of src/medium/homogeneous.cpp:82-83"""
    return dict(type="homogeneous", sigmaA=np.full(3, sigma_a, np.float32), sigmaS=np.full(3, sigma_s, np.float32),
                samplingWeight=-1.0, phase=phase, g=g)


def fbm_density(res, seed=3, octaves=3):
    """clamp(fBm, 0, 1) on a res^3 grid: value noise summed over octaves, cheap and seeded."""
    rng = np.random.default_rng(seed)
    out = np.zeros((res, res, res), np.float32)
    amp, total = 1.0, 0.0
    lin = np.linspace(0.0, 1.0, res, dtype=np.float32)
    for o in range(octaves):
        n = 4 * (2 ** o) + 1
        lattice = rng.random((n, n, n), dtype=np.float32)
        x = lin * (n - 1)
        i0 = np.minimum(x.astype(np.int32), n - 2)
        f = (x - i0).astype(np.float32)
        f = f * f * (3 - 2 * f)
        # trilinear via separable lerps
        def lerp_axis(arr, axis):
            lo = np.take(arr, i0, axis=axis)
            hi = np.take(arr, i0 + 1, axis=axis)
            shape = [1, 1, 1]
            shape[axis] = res
            w = f.reshape(shape)
            return lo * (1 - w) + hi * w
        g = lerp_axis(lattice, 0)
        g = lerp_axis(g, 1)
        g = lerp_axis(g, 2)
        out += amp * g
        total += amp
        amp *= 0.5
    out /= total
    out = np.clip((out - 0.35) * 2.5, 0.0, 1.0)
    return out.astype(np.float32)


def grid_medium(res=64, scale=8.0, albedo=0.9, seed=3, phase=0, g=0.0):
    """HeterogeneousMedium method=simpson over a procedural grid in [0,1]^3.  sigmaS_base restates the
    base-class Medium::getSigmaS() used by the vol->surf term (quirk B2): the Skin1 preset x scale x (1 - g)
    (src/librender/medium.cpp:27-37, src/medium/materials.h:108-129)."""
    skin1_sigma_s = np.array([0.74, 0.88, 1.01], np.float32) * 100.0  # materials.h Skin1 reduced scattering, mm^-1 -> x100
    return dict(type="grid", density=fbm_density(res, seed), bbox_min=np.zeros(3, np.float32),
                bbox_max=np.ones(3, np.float32), scale=np.float32(scale), albedo=np.full(3, albedo, np.float32),
                sigmaS_base=(skin1_sigma_s * np.float32(scale)).astype(np.float32), phase=phase, g=g)


# ---- VRLs --------------------------------------------------------------------------------------
def synthetic_vrls(n, sigma_t=1.05, seed=1000):
    """start ~ U([0,1]^3), direction uniform on the sphere, length ~ Exp(sigma_t) clipped to the box,
    power = e^{-sigma_t len} U(0.5, 1.5); particleCount = ceil(n / 3)."""
    rng = np.random.default_rng(seed)
    start = rng.random((n, 3))
    z = rng.uniform(-1.0, 1.0, n)
    phi = rng.uniform(0.0, 2 * math.pi, n)
    r = np.sqrt(np.maximum(0.0, 1 - z * z))
    d = np.stack([r * np.cos(phi), r * np.sin(phi), z], axis=1)
    length = rng.exponential(1.0 / sigma_t, n)
    with np.errstate(divide="ignore", invalid="ignore"):
        tmax = np.where(d > 0, (1.0 - start) / d, np.where(d < 0, (0.0 - start) / d, np.inf)).min(axis=1)
    length = np.minimum(length, 0.999 * tmax)
    length = np.maximum(length, 1e-3)
    end = start + d * length[:, None]
    power = (np.exp(-sigma_t * length) * rng.uniform(0.5, 1.5, n))[:, None] * np.ones((1, 3))
    return (start.astype(np.float32), end.astype(np.float32), power.astype(np.float32), int(math.ceil(n / 3)))


def write_vrl_file(path, start, end, power):
    """ASCII format of VRL.h:43-54: sx sy sz ex ey ez r g b"""
    with open(path, "w") as f:
        for s, e, p in zip(start, end, power):
            f.write(" ".join(repr(float(x)) for x in (*s, *e, *p)) + "\n")


# ---- configs -----------------------------------------------------------------------------------
def cornell_scene(width, height, medium=None, closed=False, mesh=None):
    v, t, m = mesh if mesh is not None else _cornell_mesh(closed).arrays()
    if closed:   # the camera has to sit inside the closed walls (and inside the medium)
        cam = perspective_camera(width, height, origin=(0.5, 0.5, 0.02), target=(0.5, 0.5, 1.0), fov=80.0)
    else:
        cam = perspective_camera(width, height, origin=(0.5, 0.5, -1.4), target=(0.5, 0.5, 0.0), fov=40.0)
    return dict(verts=v, tris=t, tri_material=m, albedo=_ALBEDO.copy(),
                mat_bits=np.full(len(_ALBEDO), BSDF_SMOOTH, np.uint32),
                medium=medium or homogeneous_medium(), camera=cam, extra_bounds=cam["origin"].reshape(1, 3))


def tracer_scene(width, height, medium=None, glass=True):
    """Closed Cornell box (front wall added, camera inside) with a ceiling area light -- a quad just below the ceiling, facing
    down -- and, optionally, the glass sphere and the conductor box of chain_scene: what the VRL tracer (vrlTracer.h) walks.
    Returns (scene, emitter triangle indices, radiance)."""
    scene = dict(chain_scene(width, height, medium) if glass else cornell_scene(width, height, medium))
    v, t, m = scene["verts"], scene["tris"], scene["tri_material"]
    q = np.array([(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0),                                      # front wall (+z)
                  (0.35, 0.998, 0.35), (0.65, 0.998, 0.35), (0.65, 0.998, 0.65), (0.35, 0.998, 0.65)], np.float32)   # light (-y)
    i = len(v)
    scene["verts"] = np.concatenate([v, q])
    scene["tris"] = np.concatenate([t, np.array([(i, i + 1, i + 2), (i, i + 2, i + 3), (i + 4, i + 5, i + 6), (i + 4, i + 6, i + 7)], np.uint32)])
    scene["tri_material"] = np.concatenate([m, np.full(4, WHITE, np.uint32)])
    cam = perspective_camera(width, height, origin=(0.5, 0.5, 0.02), target=(0.5, 0.5, 1.0), fov=80.0)
    scene["camera"] = cam
    scene["extra_bounds"] = cam["origin"].reshape(1, 3)
    return scene, np.array([len(t) + 2, len(t) + 3], np.uint32), np.array([18.0, 15.0, 12.0], np.float32)


def chain_scene(width, height, medium=None, glass_eta=1.5):
    """Cornell box with a glass sphere (smooth dielectric, the fog outside, vacuum inside) in front of the boxes and the tall
    box turned into a copper-like mirror (smooth conductor): camera segments that end on them continue as specular chains
    (vrlIntegrator.cpp:445-511)."""
    m = _Mesh()
    m.quad((0, 0, 0), (0, 0, 1), (1, 0, 1), (1, 0, 0), WHITE)
    m.quad((0, 1, 0), (1, 1, 0), (1, 1, 1), (0, 1, 1), WHITE)
    m.quad((0, 0, 1), (0, 1, 1), (1, 1, 1), (1, 0, 1), WHITE)
    m.quad((0, 0, 0), (0, 1, 0), (0, 1, 1), (0, 0, 1), RED)
    m.quad((1, 0, 0), (1, 0, 1), (1, 1, 1), (1, 1, 0), GREEN)
    GLASS, MIRROR = 3, 4
    m.box((0.33, 0.15, 0.33), (0.15, 0.15, 0.15), 17.0, WHITE)
    m.box((0.67, 0.30, 0.64), (0.15, 0.30, 0.15), -18.0, MIRROR)
    v0, t0, m0 = m.arrays()
    sv, sf = _icosphere(2)
    verts = (np.array([0.36, 0.52, 0.30]) + 0.17 * sv).astype(np.float32)
    faces = (sf + len(v0)).astype(np.uint32)
    v = np.concatenate([v0, verts]); t = np.concatenate([t0, faces]); mm = np.concatenate([m0, np.full(len(faces), GLASS, np.uint32)])
    albedo = np.concatenate([_ALBEDO, np.zeros((2, 3), np.float32)])
    bits = np.array([BSDF_SMOOTH, BSDF_SMOOTH, BSDF_SMOOTH,
                     BSDF_DIELECTRIC | MAT_TRANSITION | MAT_EXTERIOR_MEDIUM, BSDF_CONDUCTOR], np.uint32)
    optics = np.zeros((5, 12), np.float32)
    optics[:, 6:12] = 1.0
    optics[GLASS, 0] = glass_eta
    optics[MIRROR, 0:3] = (0.27, 0.68, 1.22)          # eta, k of a copper-like conductor (rgb)
    optics[MIRROR, 3:6] = (3.61, 2.63, 2.29)
    cam = perspective_camera(width, height, origin=(0.5, 0.5, -1.4), target=(0.5, 0.5, 0.0), fov=40.0)
    return dict(verts=v, tris=t, tri_material=mm, albedo=albedo, mat_bits=bits, optics=optics,
                medium=medium or homogeneous_medium(), camera=cam, extra_bounds=cam["origin"].reshape(1, 3))


CONFIGS = {
    "C1": dict(width=256, height=256, n_vrls=1000, params=dict(volVolSamples=2, volSurfSamples=2)),
    "C2": dict(width=1024, height=1024, n_vrls=100_000, params=dict(volVolSamples=4, volSurfSamples=4)),
    "C3": dict(width=1024, height=1024, n_vrls=200_000, params=dict(volVolSamples=2, volSurfSamples=2), grid=512),
    "C4": dict(width=1920, height=1080, n_vrls=300_000, params=dict(volVolSamples=2, volSurfSamples=2), occluders=780),
    "C5": dict(width=3840, height=2160, n_vrls=1_000_000,
               params=dict(volVolSamples=16, volSurfSamples=4, targetNumSlices=512), hg=0.8),
}


def make_config(name, width=None, height=None, n_vrls=None, grid=None, occluders=None):
    """Returns (scene, (start, end, power, particleCount), params) for a BASELINE config, optionally shrunk."""
    c = dict(CONFIGS[name])
    width, height = width or c["width"], height or c["height"]
    n = n_vrls or c["n_vrls"]
    idx = int(name[1:])
    if name == "C3":
        med = grid_medium(res=grid or c["grid"])
        sigma_t = 4.0
    elif name == "C4":
        med = homogeneous_medium(sigma_s=0.25, sigma_a=0.05)
        sigma_t = 0.3
    elif name == "C5":
        med = homogeneous_medium(phase=1, g=c["hg"])
        sigma_t = 1.05
    else:
        med = homogeneous_medium()
        sigma_t = 1.05
    mesh = None
    if name == "C4":
        mesh = occluder_mesh(n_spheres=occluders or c["occluders"])
    scene = cornell_scene(width, height, medium=med, closed=(name == "C4"), mesh=mesh)
    vrls = synthetic_vrls(n, sigma_t=sigma_t, seed=1000 + idx)
    return scene, vrls, dict(c["params"])
