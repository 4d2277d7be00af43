/*
 * alvrl.h -- C ABI of the B200-native Adaptive-LightSlice VRL hot path (libalvrl.so).
 *
 * This is the drop-in boundary a Mitsuba `integrator type="vrl"` plugin binds instead of
 * running the reference's CPU loops.  Every entry point names the reference code it replaces
 * (paths relative to the reference tree, neodyme06/mitsuba-ALVRL).  Plain pointers and sizes
 * only; no C++ types, no exceptions; every function returns ALVRL_OK (0) or a negative error
 * code and leaves a message in alvrl_last_error().  Host buffers belong to the caller and are
 * copied during the call; device memory belongs to the library.  One handle drives one GPU and
 * is not thread-safe.  There is no CPU fallback: without a usable CUDA device alvrl_create fails.
 *
 * Pixel indexing follows the reference's `m_slices[y + H*x]` (vrlIntegrator.cpp:560,
 * Preprocessor.cpp:1140-1146): "pixel index" below always means  y + H*x  (x-major).
 */
#ifndef ALVRL_H
#define ALVRL_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ALVRL_OK              0
#define ALVRL_ERR_ARG        -1   /* invalid argument / parameter (the reference Log(EError)s) */
#define ALVRL_ERR_STATE      -2   /* call order violated (e.g. build_R before build_slices)      */
#define ALVRL_ERR_CUDA       -3   /* CUDA runtime failure                                         */
#define ALVRL_ERR_IO         -4   /* file could not be read                                       */
#define ALVRL_ERR_UNSUPPORTED -5  /* scene feature outside the hot path          */

#define ALVRL_NO_SLICE 0xffffffffu      /* UINT32_T_MAX slice id, Preprocessor.cpp:1201 */
#define ALVRL_NO_HIT   0xffffffffu

/* BSDF type bits the path needs (include/mitsuba/render/bsdf.h:230-284) */
#define ALVRL_BSDF_SMOOTH 1u            /* diffuse reflectance: ESmooth */
/* delta BSDFs of the specular chains (vrlIntegrator.cpp:445-511) and the media on the two sides of a surface */
#define ALVRL_BSDF_DIELECTRIC 2u        /* smooth dielectric: EDeltaReflection | EDeltaTransmission (src/bsdfs/dielectric.cpp) */
#define ALVRL_BSDF_CONDUCTOR  4u        /* smooth conductor: EDeltaReflection (src/bsdfs/conductor.cpp) */
#define ALVRL_BSDF_DELTA (ALVRL_BSDF_DIELECTRIC | ALVRL_BSDF_CONDUCTOR)
#define ALVRL_MAT_TRANSITION      8u    /* the shape has an interior or exterior medium (Shape::isMediumTransition) */
#define ALVRL_MAT_INTERIOR_MEDIUM 16u   /* ... its interior medium is the scene's medium (else vacuum) */
#define ALVRL_MAT_EXTERIOR_MEDIUM 32u   /* ... its exterior medium is the scene's medium (else vacuum) */

#define ALVRL_PHASE_ISOTROPIC 0         /* src/phase/isotropic.cpp:76-78 */
#define ALVRL_PHASE_HG        1         /* src/phase/hg.cpp:107-110      */

#define ALVRL_RNG_MODE_COUNTER 0        /* include/alvrl_rng.h                                        */
#define ALVRL_RNG_MODE_SFMT    1        /* reference stream: one SFMT-19937 sampler, workerCount = 1  */

/*
 * XML parameters of the reference plugin (vrlIntegrator.cpp:128-208, same names and defaults),
 * the inherited ones the path reads, and a few implementation knobs that have no XML equivalent.
 */
typedef struct alvrl_params {
    int32_t shortVrls;                  /* true  */
    int32_t vrlTargetNum;               /* 500  (tracer only; kept for the plugin interface) */
    int32_t maxParticleDepth;           /* -1   (tracer only) */
    int32_t specularForcedRRdepth;      /* 100  (specular chains; host side) */
    float   initialSpecularThroughput;  /* 20 */
    int32_t volVolSamples;              /* 2, must be 0 or >= 2 */
    int32_t volSurfSamples;             /* 2, must be 0 or >= 2 */
    int32_t globalCluster;              /* false */
    float   globalUndersampling;        /* -1 */
    int32_t localRefinement;            /* true */
    float   localUndersampling;         /* -1 */
    float   fallBackUndersampling;      /* 5 */
    int32_t targetNumSlices;            /* 100 */
    float   targetPixelUndersampling;   /* 64 */
    float   sliceCurvatureFactor;       /* 0.5 */
    int32_t neighbourCount;             /* 0 */
    float   neighbourWeight;            /* 0 */
    int32_t Rsamples;                   /* 1 */
    float   depthCorrection;            /* 1 */
    int32_t numVrlFalseColor;           /* false */
    int32_t slicesFalseColor;           /* false */
    int32_t convergenceFalseColor;      /* false */
    int32_t maxPasses;                  /* 1 (src/librender/integrator.cpp:348) */
    /* ---- implementation knobs ---- */
    int32_t rngMode;                    /* ALVRL_RNG_MODE_COUNTER */
    uint64_t seed;                      /* 0 */
    int32_t anyHitShadowRays;           /* 1: shadow rays stop at the first occluder (same T as closest hit
                                           in scenes without ENull surfaces, scene.cpp:634-642) */
    int32_t workerCount;                /* 1: emulated Scheduler::getWorkerCount() -- decides how the SFMT
                                           sampler is cloned over contiguous slice ranges
                                           (vrlIntegrator.cpp:305-321,1048-1051; Preprocessor.cpp:738-741) */
    int32_t rrDepth;                    /* 5 (tracer only; src/librender/integrator.cpp:272-298) */
    int32_t reserved[5];
} alvrl_params;

typedef struct alvrl_ctx *alvrl_handle;

typedef struct alvrl_stats {
    uint64_t pairsPreprocess;   /* "Number of integrated VRLs during preprocessing", vrlIntegrator.cpp:119 */
    uint64_t pairsRender;       /* "Number of integrated VRLs during rendering",     vrlIntegrator.cpp:121 */
    uint64_t shadowRays;
    float msSlices, msSliceMapping, msBuildR, msClusters, msRender;   /* device/host phase timers */
    float msTransportKernelR, msTransportKernelRender;                /* CUDA-event time of the kernels */
    uint32_t kernelLaunches;    /* kernels launched by this handle since creation */
    uint32_t numSlices, numRows, numVrls, bvhNodes;
    uint32_t visMode;           /* shadow-ray strategy of the fast flavour: 0 BVH traversal, 1 flat leaf sweep, 2 compiled occluder set */
    float msSceneBuild;         /* host: BVH build + triangle records + occluder set + upload of the last scene change; outside the
                                   frame time by definition ("excluding scene upload and BVH build (reported separately)") */
} alvrl_stats;

/* ---- life cycle -------------------------------------------------------------------------- */
void alvrl_params_default(alvrl_params *p);                   /* vrlIntegrator.cpp:128-208 defaults */
/* CreateInstance(props) -> new vrlIntegrator(props), vrlIntegrator.cpp:128,1127 */
int  alvrl_create(int cuda_device, const alvrl_params *p, alvrl_handle *out);
void alvrl_destroy(alvrl_handle h);
const char *alvrl_last_error(void);
int  alvrl_get_params(alvrl_handle h, alvrl_params *out);
/* 0 (default): fast-math flavour of the transport kernels; 1: strict flavour (reference operation order,
 * no FMA contraction, exp through double).  Also selectable with the environment variable ALVRL_MATH=strict. */
int  alvrl_set_math_mode(alvrl_handle h, int strict);

/* ---- scene upload (what the plugin marshals out of `const Scene *`) -------------------- */
/* Triangle soup + per-triangle material; replaces ShapeKDTree's TriAccel array
 * (include/mitsuba/render/triaccel.h:61-95, src/librender/skdtree.cpp:60-104) with a device BVH. */
int alvrl_set_mesh(alvrl_handle h, const float *verts_xyz, uint32_t nverts,
                   const uint32_t *tris, uint32_t ntris, const uint32_t *tri_material);
/* The reference's analytic shapes, handed over as triangles appended to the mesh (call after alvrl_set_mesh, or instead of
 * it; alvrl_set_mesh starts over).  firstTriangle / triangleCount (may be NULL) tell where they went, e.g. for
 * alvrl_set_area_emitter.
 *   rectangle (src/shapes/rectangle.cpp:76-118,170-196): the square [-1, 1]^2 of the xy plane under toWorld (row-major 4x4,
 *     affine), flipNormals as there; two triangles cover exactly the same surface, wound so that their geometric normal is
 *     the shape's normal.
 *   sphere (src/shapes/sphere.cpp:108-131,245-251): centre, radius, flipNormals (what the constructor folds toWorld into);
 *     an approximation by 4 T (T - 2) triangles with the vertices on the sphere, T = thetaSteps polar steps (0 selects 64:
 *     the surface is within 6.5e-4 radius of the sphere). */
int alvrl_add_rectangle(alvrl_handle h, const float toWorld[16], int flipNormals, uint32_t material, uint32_t *firstTriangle);
int alvrl_add_sphere(alvrl_handle h, const float center[3], float radius, int flipNormals, uint32_t thetaSteps, uint32_t material,
                     uint32_t *firstTriangle, uint32_t *triangleCount);
/* Diffuse reflectance + type bits per material (src/bsdfs/diffuse.cpp:110-118). */
int alvrl_set_materials(alvrl_handle h, const float *albedo_rgb, const uint32_t *type_bits, uint32_t nmat);
/* Optics of the delta BSDFs (ALVRL_BSDF_DIELECTRIC / ALVRL_BSDF_CONDUCTOR), 12 floats per material: dielectric eta =
 * intIOR / extIOR in [0] (dielectric.cpp:172-189); conductor eta rgb in [0..2] and k rgb in [3..5] (conductor.cpp:199-215);
 * specularReflectance rgb in [6..8], specularTransmittance rgb in [9..11].  With such materials in the scene every camera
 * segment grows the specular chain of LiInternal (vrlIntegrator.cpp:445-511): one segment per delta component that survives
 * the Russian roulette, recursively, each with its weight, all looked up in the slice of the ORIGINAL camera ray.  The rows
 * of R (getLiLuminanceVrlContributions, 514-526) and the render pass sum over a pixel's chain.  Call after
 * alvrl_set_materials.  Chains are cut after 30 bounces (the reference only forces roulette from specularForcedRRdepth on). */
int alvrl_set_material_optics(alvrl_handle h, const float *optics12, uint32_t nmat);
/* Introspection (tests): the chain segments beyond the camera segment, grouped by pixel.  offset: P + 1 entries; segs: 16
 * floats per segment {o.xyz, d.xyz, p.xyz, dist, weight.rgb, in-medium flag, path code, material}.  Pass segs = NULL to
 * get the offsets (and through offset[P] the total) first. */
int alvrl_get_chain_segments(alvrl_handle h, uint32_t *offset, float *segs);
/* Extra points to union into Scene::getAABB() (sensor/emitter boxes, scene.cpp:387-413). */
int alvrl_set_extra_bounds(alvrl_handle h, const float *points_xyz, uint32_t npoints);
/* HomogeneousMedium (src/medium/homogeneous.cpp:156-184,354-396); samplingWeight < 0 => reference default. */
int alvrl_set_medium_homogeneous(alvrl_handle h, const float sigmaA[3], const float sigmaS[3],
                                 float mediumSamplingWeight, int32_t phaseType, float g);
/* HeterogeneousMedium, method=simpson, over a float32 grid (src/medium/heterogeneous.cpp:301-376,665-691;
 * src/volume/gridvolume.cpp:188-215,337-388).  sigmaS_base is Medium::getSigmaS() (quirk B2). */
int alvrl_set_medium_grid(alvrl_handle h, const float *density, const int32_t res[3],
                          const float bbox_min[3], const float bbox_max[3], float scale,
                          const float albedo[3], const float sigmaS_base[3],
                          int32_t phaseType, float g);
/* The same medium with the density read from the grid volume file a `gridvolume` plugin would map
 * (src/volume/gridvolume.cpp:217-287: "VOL", version 3; one channel, float32 or uint8 -- uint8 through the reference's
 * density map i / 255.0f, gridvolume.cpp:212-215,374-389).  bbox_min / bbox_max: NULL takes the AABB stored in the file,
 * non-NULL is the `min` / `max` override of gridvolume.cpp:112-117.  Errors keep the reference's messages (ALVRL_ERR_IO:
 * unreadable or truncated; ALVRL_ERR_ARG: bad identifier / version / type; ALVRL_ERR_UNSUPPORTED: float16, 3 channels). */
int alvrl_set_medium_grid_file(alvrl_handle h, const char *path, const float *bbox_min, const float *bbox_max, float scale,
                               const float albedo[3], const float sigmaS_base[3], int32_t phaseType, float g);
/* PerspectiveCamera (src/sensors/perspective.cpp:126-175,247-269): row-major 4x4 matrices. */
int alvrl_set_camera(alvrl_handle h, const float sampleToCamera[16], const float cameraToWorld[16],
                     uint32_t width, uint32_t height, float nearClip, float farClip);
/* vrlVector (src/integrators/vrl/VRL.h:105-194); zero-power / zero-length VRLs are dropped like put()
 * (VRL.h:148-158).  particleCount == 0 => number of kept VRLs (VRL.h:128). */
int alvrl_set_vrls(alvrl_handle h, const float *start_xyz, const float *end_xyz, const float *power_rgb,
                   uint32_t n, uint64_t particleCount);
/* ASCII VRL file, 9 floats per line (VRL.h:43-54,120-128; vrlIntegrator.cpp:243-252). */
int alvrl_load_vrl_file(alvrl_handle h, const char *path);
/* Parity mode for build_R: u(row, vrl, k) = tape[(row*N + vrl)*(2*Nvv+Nvs) + k]. NULL clears. */
int alvrl_set_sample_tape(alvrl_handle h, const float *tape, uint64_t n);

/* ---- the path ------------------------------------------------------------------------------- */
/* vrlIntegrator::preprocess -> Preprocessor::buildSlices (Preprocessor.cpp:1130-1227,1349-1418). */
int alvrl_build_slices(alvrl_handle h);
/* Same, from caller-supplied gather points (P x 3 position, P x 3 scaled normal; NaN = miss). */
int alvrl_build_slices_from_gather(alvrl_handle h, const float *pos_xyz, const float *dir_xyz);
/* Preprocessor::sampleSliceMapping (Preprocessor.cpp:66-121,1502-1525). */
int alvrl_sample_slice_mapping(alvrl_handle h);
/* "Building R": prepass loops + Rbuilder + getLiLuminanceVrlContributions + getVRLContributions
 * (vrlIntegrator.cpp:302-337,527-539,792-825,1038-1083). */
int alvrl_build_R(alvrl_handle h);
/* Multi-GPU exchange step of buildClusters: Preprocessor::cluster() splits the VRLs into zero / non-zero columns over
 * ALL rows of R (Preprocessor.cpp:846-855,936-945).  With rows sharded by slice each rank computes its local flags
 * (get), the ranks combine them with a logical OR (e.g. ncclAllReduce MAX on N bytes) and hand the result back (set)
 * before build_clusters.  Without a set call the local flags are used (single GPU). */
int alvrl_get_column_nonzero(alvrl_handle h, uint8_t *flags /*N*/);
int alvrl_set_column_nonzero(alvrl_handle h, const uint8_t *flags /*N, NULL clears*/);
/* Preprocessor::buildClusters (Preprocessor.cpp:133-283). */
int alvrl_build_clusters(alvrl_handle h);
/* vrlIntegrator::prepass = the three calls above (vrlIntegrator.cpp:270-356). */
int alvrl_prepass(alvrl_handle h);
/* Render pass: Li -> getClusteredVrlContributions for every pixel centre (vrlIntegrator.cpp:386-393,
 * 542-599; src/librender/integrator.cpp:232-264 with spp == 1, rfilter = box).
 * rgb: W*H*3 floats, row-major image order [y][x][c]. */
int alvrl_render(alvrl_handle h, float *rgb_host);
/* Unclustered render (globalCluster = localRefinement = false): getVRLContributions over all VRLs. */
int alvrl_render_unclustered(alvrl_handle h, float *rgb_host);
/* A new sample stream for the next progressive pass.  ProgressiveMonteCarloIntegrator::render (src/librender/integrator.cpp:
 * 398-434) runs prepass + render once per pass on a sampler that keeps advancing, so every pass traces a fresh VRL set and
 * draws fresh samples; in the counter stream a pass is addressed by its seed.  Everything drawn from the stream (traced VRLs
 * stay until alvrl_trace_vrls is called again; slice mapping, R, clusters, specular chains) is invalidated; the slices are
 * not (Preprocessor::buildSlices draws nothing).  Counter stream only. */
int alvrl_set_seed(alvrl_handle h, uint64_t seed);

/* ---- VRL tracer: the step before the path ----------------------------------------------------------------
 * vrlTracer::randomWalk (src/integrators/vrl/vrlTracer.h:14-58): light particles are traced from an area emitter
 * (Scene::sampleEmitterPosition, scene.cpp:958-974; AreaLight::samplePosition / sampleDirection, src/emitters/area.cpp:94-123;
 * TriMesh::samplePosition, trimesh.cpp:412-423) through the medium (HomogeneousMedium::sampleDistance, homogeneous.cpp:275-352;
 * phase function sampling, isotropic.cpp:62-67, hg.cpp:74-98) and off the surfaces (diffuse.cpp:129-138, dielectric.cpp:335-364,
 * conductor.cpp:254-268, medium transitions) with Russian roulette from rrDepth on (traceOneParticle, vrlTracer.h:91-230);
 * every path segment inside the scattering medium becomes a VRL (vrlVector::put, VRL.h:148-158) until vrlTargetNum VRLs
 * exist.  Particle i draws from the counter stream of (ALVRL_RNG_TRACER, i); the particles are independent, so the device
 * traces them in parallel and keeps the particles 0 .. n-1, n the first count that reaches the target -- what the
 * reference's sequential loop stops at.  Homogeneous media and grid media (HeterogeneousMedium::sampleDistance, method = simpson,
 * heterogeneous.cpp:422-545, 589-616: the Simpson march inverted by Newton-bisection).
 *   emitter_tris: indices (into the mesh of alvrl_set_mesh) of the triangles of the emitter's shape, in the shape's order.
 * alvrl_trace_vrls replaces the handle's VRL set (as alvrl_set_vrls would) and its particle count. */
int alvrl_set_area_emitter(alvrl_handle h, const uint32_t *emitter_tris, uint32_t ntris, const float radiance_rgb[3]);
int alvrl_trace_vrls(alvrl_handle h, uint32_t target_num /* 0: params.vrlTargetNum */);
int alvrl_get_vrls(alvrl_handle h, float *start_xyz, float *end_xyz, float *power_rgb, uint64_t *particle_count);

/* ---- ground truth: volpath restricted to VRL paths ------------------------------------------------------
 * VolumetricPathTracer::Li / Li_original with `onlyVRLpaths` (src/integrators/path/volpath.cpp:76-460: the unbiased estimate of
 * exactly the light paths the VRL integrator renders -- first vertex in the volume or on a diffuse surface inside the medium,
 * second vertex in the volume, initial specular vertices ignored) and rayIntersectAndLookForEmitter (484-535), driven like
 * SamplingIntegrator::renderBlock (src/librender/integrator.cpp:210-268): `spp` samples per pixel (pixel centre when spp == 1,
 * jittered otherwise), each the mean of `internalSamples` walks (volpath.cpp:111-120); box reconstruction filter, invalid
 * samples rejected (imageblock.h:147-151).  Emitter sampling + phase / BSDF sampling combined by the power heuristic, Russian
 * roulette from params.rrDepth on.  Needs alvrl_set_area_emitter; one homogeneous or grid medium (the sensor sits in it), diffuse /
 * smooth dielectric / smooth conductor surfaces, no index-matched (ENull) boundaries, no environment emitter.
 * Outer sample j of pixel p draws from the counter stream of (ALVRL_RNG_VOLPATH, p, j); thread = pixel, one launch per outer
 * sample, so the device image equals the oracle's bit for bit.  max_depth: -1 = unlimited (the `maxDepth` property).
 * rgb_host: W*H*3 floats [y][x][c].  The image metric that goes with it (src/utils/rms.cpp) is mitsuba-alvrl_b200/rms.py. */
#define ALVRL_VOLPATH_ONLY_VRL_PATHS  1u    /* onlyVRLpaths   (default true)  */
#define ALVRL_VOLPATH_VOL_TO_VOL      2u    /* vrlVolToVol    (default true)  */
#define ALVRL_VOLPATH_VOL_TO_SURF     4u    /* vrlVolToSurf   (default true)  */
#define ALVRL_VOLPATH_SINGLE_SCATTER  8u    /* onlySingleScatter (default false) */
#define ALVRL_VOLPATH_STRICT_NORMALS 16u    /* strictNormals  (default false) */
#define ALVRL_VOLPATH_HIDE_EMITTERS  32u    /* hideEmitters   (default false) */
#define ALVRL_VOLPATH_CENTRE_SAMPLES 64u    /* every sample through the pixel centre, as the VRL render pass does */
#define ALVRL_VOLPATH_DEFAULT (ALVRL_VOLPATH_ONLY_VRL_PATHS | ALVRL_VOLPATH_VOL_TO_VOL | ALVRL_VOLPATH_VOL_TO_SURF)
int alvrl_volpath_render(alvrl_handle h, uint32_t spp, uint32_t internalSamples, uint32_t flags, int32_t max_depth, float *rgb_host);

/* ---- film: the step after the path -------------------------------------------------------------------
 * Reconstruction-filter splat and pass accumulation: ImageBlock::put (include/mitsuba/render/imageblock.h:124-202) with
 * the pre-rasterised filter of ReconstructionFilter::configure / evalDiscretized (src/libcore/rfilter.cpp:37-55,
 * include/mitsuba/core/rfilter.h:76-77), channels {rgb, alpha, weight}, and the division by the weight channel when the
 * film is developed (src/libcore/bitmap.cpp:1617-1624).  One sample per pixel centre and pass.
 *   filter: ALVRL_FILTER_BOX (param = radius, default 0.5; src/rfilters/box.cpp:38,46-48), ALVRL_FILTER_TENT (radius 1,
 *   src/rfilters/tent.cpp:34,42-44), ALVRL_FILTER_GAUSSIAN (param = stddev, default 0.5, radius 4 stddev; the scene
 *   default, src/rfilters/gaussian.cpp:30-58).  param <= 0 selects the default.
 * alvrl_film_configure clears the film; alvrl_film_put adds one pass -- rgb_host (W*H*3, [y][x][c]) or, with NULL, the frame
 * the last alvrl_render left on the device; alvrl_film_develop returns the normalised image. */
#define ALVRL_FILTER_BOX 0
#define ALVRL_FILTER_TENT 1
#define ALVRL_FILTER_GAUSSIAN 2
int alvrl_film_configure(alvrl_handle h, int filter, float param);
int alvrl_film_clear(alvrl_handle h);
int alvrl_film_put(alvrl_handle h, const float *rgb_host);
int alvrl_film_develop(alvrl_handle h, float *rgb_host);
/* The developed film as the NumPy file the reference's `mfilm` writes with fileFormat = numpy (src/films/mfilm.cpp:337-348
 * through cnpy::npy_save, src/films/cnpy.h:207-236): format 1.0, '<f4', C order, shape (H, W, 3). */
int alvrl_film_write_npy(alvrl_handle h, const char *path);

/* Device variants for multi-GPU: only slices [sliceBegin, sliceEnd) are processed; the
 * framebuffer (W*H*4 floats, zero-initialised by the caller) lives in caller-owned device memory. */
int alvrl_set_slice_range(alvrl_handle h, uint32_t sliceBegin, uint32_t sliceEnd);
int alvrl_render_device(alvrl_handle h, void *rgba_device, void *cuda_stream);

/* ---- results / introspection (tests, multi-GPU exchange) -------------------------------- */
int alvrl_get_stats(alvrl_handle h, alvrl_stats *out);
/* sustained FP32 FFMA rate of the device measured by a register-resident microbenchmark (roofline denominator) */
int alvrl_measure_fp32_peak(int cuda_device, float *tflops);
int alvrl_get_num_vrls(alvrl_handle h, uint32_t *n);          /* vrlVector::size() after the put() filter */
int alvrl_get_primary_hits(alvrl_handle h, uint32_t *prim /*P*/, float *t /*P*/, float *p_xyz /*3P*/, float *n_xyz /*3P*/);
int alvrl_get_pixel_to_slice(alvrl_handle h, uint32_t *out /*P*/);
int alvrl_get_num_slices(alvrl_handle h, uint32_t *nslices, uint32_t *nrows);
int alvrl_get_rep_pixels(alvrl_handle h, uint32_t *sliceRowOffset /*S+1*/, uint32_t *rowPixel /*G*/);
int alvrl_set_rep_pixels(alvrl_handle h, const uint32_t *sliceRowOffset, const uint32_t *rowPixel, uint32_t nslices);
int alvrl_get_R(alvrl_handle h, uint32_t rowBegin, uint32_t rowEnd, float *mean_var /*(rows*N*2)*/);
int alvrl_set_R(alvrl_handle h, const float *mean_var /*(G*N*2)*/);
int alvrl_get_cluster_counts(alvrl_handle h, uint32_t *sliceOffset /*S+1*/, uint32_t *nGlobal, uint32_t *nFallback);
int alvrl_get_clusters(alvrl_handle h, uint32_t *vrls, float *weights,
                       uint32_t *globalVrls, float *globalWeights,
                       uint32_t *fallbackVrls, float *fallbackWeights);
int alvrl_set_clusters(alvrl_handle h, const uint32_t *sliceOffset, uint32_t nslices,
                       const uint32_t *vrls, const float *weights,
                       const uint32_t *fallbackVrls, const float *fallbackWeights, uint32_t nFallback);
/* Closest-hit query through the device BVH with ShapeKDTree::rayIntersect semantics
 * (skdtree.cpp:112-204): prim = triangle index or ALVRL_NO_HIT. */
int alvrl_trace_rays(alvrl_handle h, const float *o_xyz, const float *d_xyz, const float *mint,
                     const float *maxt, uint32_t n, uint32_t *prim, float *t);
/* Scene::evalTransmittance (scene.cpp:619-679) for n point pairs. */
int alvrl_eval_transmittance(alvrl_handle h, const float *p1_xyz, const int32_t *p1OnSurface,
                             const float *p2_xyz, uint32_t n, float *T_rgb);

/* ---- multi-GPU: slices sharded over the GPUs of one box (group.cu) ---------------------------------------------
 * Replaces the reference's own parallel split of the path: contiguous slice ranges over Rbuilder / ClusterRefiner threads
 * (vrlIntegrator.cpp:305-321,1048-1051; Preprocessor.cpp:212-228,738-741) and image blocks over render workers
 * (src/librender/integrator.cpp:181-198).  VRLs, mesh, medium and camera are replicated (upload them to every member with
 * the alvrl_set_* calls); a frame exchanges N bytes of column flags (all-reduce, MAX) and the framebuffer (reduce to rank 0)
 * over NCCL.  Slice ranges are cut by pixel count.  globalCluster and the fallback clustering need all rows of R on one
 * handle and are refused (ALVRL_ERR_UNSUPPORTED) while a handle owns a proper slice range. */
typedef struct alvrl_group *alvrl_group_handle;
#define ALVRL_GROUP_ID_BYTES 128
const char *alvrl_group_last_error(void);
/* one process, several GPUs: creates ndev handles (devices[i]) and their communicator */
int  alvrl_group_create_local(int ndev, const int *devices, const alvrl_params *p, alvrl_group_handle *out);
/* one process per GPU: rank 0 obtains an id, the host hands it to every rank, each rank wraps its own handle */
int  alvrl_group_unique_id(uint8_t id[ALVRL_GROUP_ID_BYTES]);
int  alvrl_group_create_rank(alvrl_handle h, int rank, int nranks, const uint8_t id[ALVRL_GROUP_ID_BYTES], alvrl_group_handle *out);
int  alvrl_group_size(alvrl_group_handle g, int *world, int *local);
int  alvrl_group_member(alvrl_group_handle g, int i, alvrl_handle *h, int *rank);
int  alvrl_group_comm_size(alvrl_group_handle g, int *nranks);            /* ncclCommCount of the communicator */
/* buildSlices -> range of each rank -> sampleSliceMapping -> "Building R" -> flag all-reduce -> buildClusters -> render ->
 * framebuffer reduce.  rgb_host: W*H*3 floats [y][x][c], filled where rank 0 lives (NULL: leave it on the device). */
int  alvrl_group_frame(alvrl_group_handle g, float *rgb_host);
int  alvrl_group_get_range(alvrl_group_handle g, int i, uint32_t *sliceBegin, uint32_t *sliceEnd);
int  alvrl_group_framebuffer(alvrl_group_handle g, int i, void **rgba_device);
void alvrl_group_destroy(alvrl_group_handle g);

#ifdef __cplusplus
}
#endif
#endif /* ALVRL_H */
