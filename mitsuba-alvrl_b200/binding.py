"""ctypes binding of the C ABI declared in include/alvrl.h.

The same class binds libalvrl.so (prefix ``alvrl_``, the CUDA product) and -- from tests/ and bench.py only --
the CPU oracle (prefix ``orc_``), because the oracle mirrors the ABI one to one.  Method names follow the
reference's own vocabulary (build_slices, sample_slice_mapping, build_R, build_clusters, prepass, render;
src/integrators/vrl/vrlIntegrator.cpp:237-356, Preprocessor.cpp:133,1130,1502).
"""
import ctypes as C
import os

import numpy as np

NO_SLICE = 0xFFFFFFFF
NO_HIT = 0xFFFFFFFF
BSDF_SMOOTH = 1
PHASE_ISOTROPIC, PHASE_HG = 0, 1
RNG_COUNTER, RNG_SFMT = 0, 1
RNG_R, RNG_RENDER, RNG_SLICEMAP, RNG_CLUSTER = 1, 2, 3, 4


class Params(C.Structure):
    """alvrl_params: XML parameters of integrator type="vrl" (vrlIntegrator.cpp:128-208)."""
    _fields_ = [
        ("shortVrls", C.c_int32), ("vrlTargetNum", C.c_int32), ("maxParticleDepth", C.c_int32),
        ("specularForcedRRdepth", C.c_int32), ("initialSpecularThroughput", C.c_float),
        ("volVolSamples", C.c_int32), ("volSurfSamples", C.c_int32), ("globalCluster", C.c_int32),
        ("globalUndersampling", C.c_float), ("localRefinement", C.c_int32), ("localUndersampling", C.c_float),
        ("fallBackUndersampling", C.c_float), ("targetNumSlices", C.c_int32),
        ("targetPixelUndersampling", C.c_float), ("sliceCurvatureFactor", C.c_float),
        ("neighbourCount", C.c_int32), ("neighbourWeight", C.c_float), ("Rsamples", C.c_int32),
        ("depthCorrection", C.c_float), ("numVrlFalseColor", C.c_int32), ("slicesFalseColor", C.c_int32),
        ("convergenceFalseColor", C.c_int32), ("maxPasses", C.c_int32),
        ("rngMode", C.c_int32), ("seed", C.c_uint64), ("anyHitShadowRays", C.c_int32),
        ("workerCount", C.c_int32), ("rrDepth", C.c_int32), ("reserved", C.c_int32 * 5),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("pairsPreprocess", C.c_uint64), ("pairsRender", C.c_uint64), ("shadowRays", C.c_uint64),
        ("msSlices", C.c_float), ("msSliceMapping", C.c_float), ("msBuildR", C.c_float),
        ("msClusters", C.c_float), ("msRender", C.c_float),
        ("msTransportKernelR", C.c_float), ("msTransportKernelRender", C.c_float),
        ("kernelLaunches", C.c_uint32), ("numSlices", C.c_uint32), ("numRows", C.c_uint32),
        ("numVrls", C.c_uint32), ("bvhNodes", C.c_uint32), ("visMode", C.c_uint32),
        ("msSceneBuild", C.c_float),
    ]


class AlvrlError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code = code


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Api:
    """Function table of one shared library."""

    def __init__(self, path, prefix):
        self.lib = C.CDLL(path)
        self.prefix = prefix
        self.path = path
        le = getattr(self.lib, prefix + "last_error")
        le.restype = C.c_char_p
        self._last_error = le

    def fn(self, name):
        f = getattr(self.lib, self.prefix + name)
        f.restype = C.c_int
        return f

    def has(self, name):
        return hasattr(self.lib, self.prefix + name)

    def default_params(self, **kw):
        p = Params()
        f = getattr(self.lib, self.prefix + "params_default")
        f.restype = None
        f(C.byref(p))
        for k, v in kw.items():
            if not hasattr(p, k):
                raise AttributeError(f"unknown vrl integrator parameter '{k}'")
            setattr(p, k, v)
        return p


class Integrator:
    """One handle of the C ABI = one `vrlIntegrator` instance bound to one device."""

    def __init__(self, api, device=0, **params):
        self.api = api
        self.params = api.default_params(**params)
        self.h = C.c_void_p()
        self._chk(api.fn("create")(C.c_int(device), C.byref(self.params), C.byref(self.h)))
        self.W = self.H = 0
        self.N = 0

    # -- plumbing --------------------------------------------------------------------------------
    def _chk(self, rc):
        if rc != 0:
            raise AlvrlError(rc, self.api._last_error().decode())

    def _call(self, name, *args):
        self._chk(self.api.fn(name)(self.h, *args))

    def close(self):
        if self.h:
            d = getattr(self.api.lib, self.api.prefix + "destroy")
            d.restype = None
            d(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- scene upload ----------------------------------------------------------------------------
    def set_scene(self, scene):
        """scene: dict from scenes.py (mesh, materials, medium, camera)."""
        self.set_mesh(scene["verts"], scene["tris"], scene["tri_material"])
        self.set_materials(scene["albedo"], scene["mat_bits"])
        if scene.get("optics") is not None:
            self.set_material_optics(scene["optics"])
        if scene.get("extra_bounds") is not None:
            eb = _f32(scene["extra_bounds"])
            self._call("set_extra_bounds", _p(eb), C.c_uint32(len(eb)))
        m = scene["medium"]
        if m["type"] == "homogeneous":
            self.set_medium_homogeneous(m["sigmaA"], m["sigmaS"], m.get("samplingWeight", -1.0),
                                        m.get("phase", PHASE_ISOTROPIC), m.get("g", 0.0))
        else:
            self.set_medium_grid(m["density"], m["bbox_min"], m["bbox_max"], m["scale"], m["albedo"],
                                 m["sigmaS_base"], m.get("phase", PHASE_ISOTROPIC), m.get("g", 0.0))
        cam = scene["camera"]
        self.set_camera(cam["sampleToCamera"], cam["cameraToWorld"], cam["width"], cam["height"],
                        cam["near"], cam["far"])

    def set_mesh(self, verts, tris, tri_material):
        v, t, m = _f32(verts), _u32(tris), _u32(tri_material)
        self._call("set_mesh", _p(v), C.c_uint32(len(v)), _p(t), C.c_uint32(len(t)), _p(m))

    def add_rectangle(self, to_world, material, flip_normals=False):
        """a `rectangle` shape (src/shapes/rectangle.cpp) as two triangles appended to the mesh -> index of the first one"""
        m = _f32(to_world).reshape(16)
        first = C.c_uint32()
        self._call("add_rectangle", _p(m), C.c_int(int(flip_normals)), C.c_uint32(material), C.byref(first))
        return first.value

    def add_sphere(self, center, radius, material, flip_normals=False, theta_steps=0):
        """a `sphere` shape (src/shapes/sphere.cpp) as triangles appended to the mesh -> (first triangle, count)"""
        c = _f32(center)
        first, count = C.c_uint32(), C.c_uint32()
        self._call("add_sphere", _p(c), C.c_float(radius), C.c_int(int(flip_normals)), C.c_uint32(theta_steps), C.c_uint32(material),
                   C.byref(first), C.byref(count))
        return first.value, count.value

    def set_materials(self, albedo, bits):
        a, b = _f32(albedo), _u32(bits)
        self._call("set_materials", _p(a), _p(b), C.c_uint32(len(b)))

    def set_seed(self, seed):
        """the sample stream of the next progressive pass (integrator.cpp:398-434)"""
        self._call("set_seed", C.c_uint64(seed))

    # -- VRL tracer (include/alvrl.h) ---------------------------------------------------------------
    def set_area_emitter(self, tris, radiance):
        t, r = _u32(tris), _f32(radiance)
        self._call("set_area_emitter", _p(t), C.c_uint32(len(t)), _p(r))

    def trace_vrls(self, target=0):
        self._call("trace_vrls", C.c_uint32(target))
        n = C.c_uint32()
        self._call("get_num_vrls", C.byref(n))
        self.N = n.value

    def get_vrls(self):
        n = C.c_uint32()
        self._call("get_num_vrls", C.byref(n))
        s, e, p = (np.zeros((n.value, 3), np.float32) for _ in range(3))
        pc = C.c_uint64()
        self._call("get_vrls", _p(s), _p(e), _p(p), C.byref(pc))
        return s, e, p, pc.value

    # -- ground truth: volpath restricted to VRL paths (include/alvrl.h) -------------------------------
    VOLPATH_ONLY_VRL_PATHS, VOLPATH_VOL_TO_VOL, VOLPATH_VOL_TO_SURF, VOLPATH_SINGLE_SCATTER = 1, 2, 4, 8
    VOLPATH_STRICT_NORMALS, VOLPATH_HIDE_EMITTERS, VOLPATH_CENTRE_SAMPLES, VOLPATH_DEFAULT = 16, 32, 64, 7

    def volpath_render(self, spp=1, internal_samples=1, flags=7, max_depth=-1):
        """VolumetricPathTracer with onlyVRLpaths (volpath.cpp:76-460): image [H, W, 3]"""
        out = np.zeros((self.H, self.W, 3), dtype=np.float32)
        self._call("volpath_render", C.c_uint32(spp), C.c_uint32(internal_samples), C.c_uint32(flags), C.c_int32(max_depth), _p(out))
        return out

    def set_material_optics(self, optics):
        o = _f32(optics).reshape(-1, 12)
        self._call("set_material_optics", _p(o), C.c_uint32(len(o)))

    def chain_segments(self):
        """(offset[P + 1], segs[total, 16]): the specular-chain segments below the camera segments, grouped by pixel"""
        off = np.zeros(self.W * self.H + 1, np.uint32)
        self._call("get_chain_segments", _p(off), None)
        segs = np.zeros((int(off[-1]), 16), np.float32)
        if len(segs):
            self._call("get_chain_segments", _p(off), _p(segs))
        return off, segs

    def set_medium_homogeneous(self, sigmaA, sigmaS, weight=-1.0, phase=PHASE_ISOTROPIC, g=0.0):
        a, s = _f32(sigmaA), _f32(sigmaS)
        self._call("set_medium_homogeneous", _p(a), _p(s), C.c_float(weight), C.c_int32(phase), C.c_float(g))

    def set_medium_grid(self, density, bmin, bmax, scale, albedo, sigmaS_base, phase=PHASE_ISOTROPIC, g=0.0):
        d = _f32(density)  # shape [z][y][x]
        res = np.array([d.shape[2], d.shape[1], d.shape[0]], dtype=np.int32)
        mn, mx, al, sb = _f32(bmin), _f32(bmax), _f32(albedo), _f32(sigmaS_base)
        self._call("set_medium_grid", _p(d), _p(res), _p(mn), _p(mx), C.c_float(scale), _p(al), _p(sb),
                   C.c_int32(phase), C.c_float(g))

    def set_medium_grid_file(self, path, scale, albedo, sigmaS_base, bmin=None, bmax=None, phase=PHASE_ISOTROPIC, g=0.0):
        """the density from a grid volume file ("VOL", version 3: src/volume/gridvolume.cpp:217-287); bmin / bmax override the
        AABB stored in the file"""
        al, sb = _f32(albedo), _f32(sigmaS_base)
        mn = _f32(bmin) if bmin is not None else None
        mx = _f32(bmax) if bmax is not None else None
        self._call("set_medium_grid_file", C.c_char_p(os.fsencode(path)), _p(mn), _p(mx), C.c_float(scale), _p(al), _p(sb),
                   C.c_int32(phase), C.c_float(g))

    def set_camera(self, s2c, c2w, W, H, near, far):
        a, b = _f32(s2c).reshape(16), _f32(c2w).reshape(16)
        self._call("set_camera", _p(a), _p(b), C.c_uint32(W), C.c_uint32(H), C.c_float(near), C.c_float(far))
        self.W, self.H = W, H

    def set_vrls(self, start, end, power, particle_count=0):
        s, e, p = _f32(start), _f32(end), _f32(power)
        self._call("set_vrls", _p(s), _p(e), _p(p), C.c_uint32(len(s)), C.c_uint64(particle_count))
        n = C.c_uint32()
        if self.api.has("get_num_vrls"):
            self._call("get_num_vrls", C.byref(n))
            self.N = n.value
        else:
            self.N = self.stats().numVrls

    def load_vrl_file(self, path):
        self._call("load_vrl_file", C.c_char_p(path.encode()))
        self.N = self.stats().numVrls

    def set_sample_tape(self, tape):
        if tape is None:
            self._call("set_sample_tape", None, C.c_uint64(0))
        else:
            t = _f32(tape).reshape(-1)
            self._call("set_sample_tape", _p(t), C.c_uint64(t.size))

    # -- the path --------------------------------------------------------------------------------
    def build_slices(self):
        self._call("build_slices")

    def build_slices_from_gather(self, pos, direc):
        a, b = _f32(pos), _f32(direc)
        self._call("build_slices_from_gather", _p(a), _p(b))

    def sample_slice_mapping(self):
        self._call("sample_slice_mapping")

    def build_R(self):
        self._call("build_R")

    def column_nonzero(self):
        out = np.zeros(self.N, np.uint8)
        self._call("get_column_nonzero", _p(out))
        return out

    def set_column_nonzero(self, flags):
        if flags is None:
            self._call("set_column_nonzero", None)
        else:
            f = np.ascontiguousarray(flags, dtype=np.uint8)
            self._call("set_column_nonzero", _p(f))

    def render_device(self, fb_ptr, stream_ptr=0):
        """fb_ptr: device pointer of a zero-initialised W*H*4 float32 framebuffer owned by the caller"""
        self._call("render_device", C.c_void_p(fb_ptr), C.c_void_p(stream_ptr))

    def build_clusters(self):
        self._call("build_clusters")

    def prepass(self):
        self._call("prepass")

    def render(self, clustered=True):
        out = np.zeros((self.H, self.W, 3), dtype=np.float32)
        self._call("render" if clustered else "render_unclustered", _p(out))
        return out

    # -- film (reconstruction-filter splat + pass accumulation, include/alvrl.h) -----------------------
    def film_configure(self, filter, param=0.0):
        self._call("film_configure", C.c_int(filter), C.c_float(param))

    def film_clear(self):
        self._call("film_clear")

    def film_put(self, rgb=None):
        if rgb is None:
            self._call("film_put", None)
        else:
            a = np.ascontiguousarray(rgb, dtype=np.float32)
            assert a.shape == (self.H, self.W, 3)
            self._call("film_put", _p(a))

    def film_develop(self):
        out = np.zeros((self.H, self.W, 3), dtype=np.float32)
        self._call("film_develop", _p(out))
        return out

    def film_write_npy(self, path):
        """the developed film as mfilm's NumPy output (src/films/mfilm.cpp:337-348): '<f4', shape (H, W, 3)"""
        self._call("film_write_npy", C.c_char_p(os.fsencode(path)))

    def set_slice_range(self, b, e):
        self._call("set_slice_range", C.c_uint32(b), C.c_uint32(e))

    # -- introspection ---------------------------------------------------------------------------
    def stats(self):
        s = Stats()
        self._call("get_stats", C.byref(s))
        return s

    def primary_hits(self):
        P = self.W * self.H
        prim = np.zeros(P, np.uint32)
        t = np.zeros(P, np.float32)
        p = np.zeros((P, 3), np.float32)
        n = np.zeros((P, 3), np.float32)
        self._call("get_primary_hits", _p(prim), _p(t), _p(p), _p(n))
        return prim, t, p, n

    def pixel_to_slice(self):
        out = np.zeros(self.W * self.H, np.uint32)
        self._call("get_pixel_to_slice", _p(out))
        return out

    def num_slices(self):
        a, b = C.c_uint32(), C.c_uint32()
        self._call("get_num_slices", C.byref(a), C.byref(b))
        return a.value, b.value

    def rep_pixels(self):
        S, G = self.num_slices()
        off = np.zeros(S + 1, np.uint32)
        px = np.zeros(G, np.uint32)
        self._call("get_rep_pixels", _p(off), _p(px))
        return off, px

    def set_rep_pixels(self, off, px):
        o, p = _u32(off), _u32(px)
        self._call("set_rep_pixels", _p(o), _p(p), C.c_uint32(len(o) - 1))

    def get_R(self, r0=0, r1=None):
        S, G = self.num_slices()
        r1 = G if r1 is None else r1
        out = np.zeros((r1 - r0, self.N, 2), np.float32)
        self._call("get_R", C.c_uint32(r0), C.c_uint32(r1), _p(out))
        return out

    def set_R(self, R):
        r = _f32(R)
        self._call("set_R", _p(r))

    def clusters(self):
        S, _ = self.num_slices()
        off = np.zeros(S + 1, np.uint32)
        ng, nf = C.c_uint32(), C.c_uint32()
        self._call("get_cluster_counts", _p(off), C.byref(ng), C.byref(nf))
        vr = np.zeros(int(off[-1]), np.uint32)
        wt = np.zeros(int(off[-1]), np.float32)
        gv, gw = np.zeros(ng.value, np.uint32), np.zeros(ng.value, np.float32)
        fv, fw = np.zeros(nf.value, np.uint32), np.zeros(nf.value, np.float32)
        self._call("get_clusters", _p(vr), _p(wt), _p(gv), _p(gw), _p(fv), _p(fw))
        return dict(offset=off, vrls=vr, weights=wt, global_vrls=gv, global_weights=gw,
                    fallback_vrls=fv, fallback_weights=fw)

    def set_clusters(self, cl):
        off, vr, wt = _u32(cl["offset"]), _u32(cl["vrls"]), _f32(cl["weights"])
        fv, fw = _u32(cl["fallback_vrls"]), _f32(cl["fallback_weights"])
        self._call("set_clusters", _p(off), C.c_uint32(len(off) - 1), _p(vr), _p(wt), _p(fv), _p(fw),
                   C.c_uint32(len(fv)))

    def trace_rays(self, o, d, mint, maxt):
        o, d, mint, maxt = _f32(o), _f32(d), _f32(mint), _f32(maxt)
        n = len(o)
        prim = np.zeros(n, np.uint32)
        t = np.zeros(n, np.float32)
        if self.api.prefix == "orc_":
            tie = np.zeros(n, np.uint8)
            self._call("trace_rays", _p(o), _p(d), _p(mint), _p(maxt), C.c_uint32(n), _p(prim), _p(t), _p(tie))
            return prim, t, tie
        self._call("trace_rays", _p(o), _p(d), _p(mint), _p(maxt), C.c_uint32(n), _p(prim), _p(t))
        return prim, t

    def eval_transmittance(self, p1, on_surface, p2):
        p1, p2 = _f32(p1), _f32(p2)
        s = np.ascontiguousarray(on_surface, dtype=np.int32)
        out = np.zeros((len(p1), 3), np.float32)
        self._call("eval_transmittance", _p(p1), _p(s), _p(p2), C.c_uint32(len(p1)), _p(out))
        return out


class Group:
    """alvrl_group_*: the ranks' handles + their NCCL communicator; one call renders a frame with the slices sharded over
    the ranks (include/alvrl.h, csrc/group.cu).  `Group.local(api, devices, **params)` drives several GPUs from this process;
    `Group.rank(integrator, rank, world, unique_id)` is the one-process-per-GPU form."""
    ID_BYTES = 128

    def __init__(self, api, handle, members):
        self.api, self.g, self.members = api, handle, members

    @staticmethod
    def _err(api, rc):
        f = api.lib.alvrl_group_last_error
        f.restype = C.c_char_p
        raise AlvrlError(rc, f().decode())

    @staticmethod
    def unique_id(api):
        buf = (C.c_uint8 * Group.ID_BYTES)()
        rc = api.lib.alvrl_group_unique_id(buf)
        if rc != 0:
            Group._err(api, rc)
        return bytes(buf)

    @classmethod
    def rank(cls, integrator, rank, world, unique_id=None):
        api = integrator.api
        g = C.c_void_p()
        idbuf = (C.c_uint8 * cls.ID_BYTES).from_buffer_copy(unique_id) if unique_id is not None else None
        rc = api.lib.alvrl_group_create_rank(integrator.h, C.c_int(rank), C.c_int(world), idbuf, C.byref(g))
        if rc != 0:
            cls._err(api, rc)
        return cls(api, g, [integrator])

    @classmethod
    def local(cls, api, devices, **params):
        p = api.default_params(**params)
        g = C.c_void_p()
        dev = (C.c_int * len(devices))(*devices)
        rc = api.lib.alvrl_group_create_local(C.c_int(len(devices)), dev, C.byref(p), C.byref(g))
        if rc != 0:
            cls._err(api, rc)
        members = []
        for i in range(len(devices)):
            it = Integrator.__new__(Integrator)
            it.api, it.params, it.h, it.W, it.H, it.N = api, p, C.c_void_p(), 0, 0, 0
            api.lib.alvrl_group_member(g, C.c_int(i), C.byref(it.h), None)
            it.close = lambda: None                       # the group owns these handles
            members.append(it)
        return cls(api, g, members)

    def comm_size(self):
        n = C.c_int()
        rc = self.api.lib.alvrl_group_comm_size(self.g, C.byref(n))
        if rc != 0:
            self._err(self.api, rc)
        return n.value

    def frame(self, want_image=True):
        """one frame over all ranks; returns the H x W x 3 image where rank 0 lives (None elsewhere / when not wanted)"""
        m = self.members[0]
        out = np.zeros((m.H, m.W, 3), dtype=np.float32) if want_image else None
        rc = self.api.lib.alvrl_group_frame(self.g, _p(out))
        if rc != 0:
            self._err(self.api, rc)
        return out

    def slice_range(self, i=0):
        b, e = C.c_uint32(), C.c_uint32()
        self.api.lib.alvrl_group_get_range(self.g, C.c_int(i), C.byref(b), C.byref(e))
        return b.value, e.value

    def close(self):
        if self.g:
            d = self.api.lib.alvrl_group_destroy
            d.restype = None
            d(self.g)
            self.g = C.c_void_p()
