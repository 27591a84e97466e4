/* b2pt.h — C ABI of the B200-native path-tracing engine (libb2pt.so).
 *
 * This is the drop-in boundary for the reference's GPU renderer: each entry point replaces one
 * step of `OptixRenderer`'s public lifecycle (reference include/gpu/optix_renderer.hpp:11-42, driven
 * by src/main.cpp:74-96).  Plain C types, caller-owned buffers, no exceptions across the boundary:
 * every function returns B2PT_OK (0) or a negative status, and b2pt_last_error() returns the text
 * the reference would have thrown as std::runtime_error (include/gpu/cuda_utils.hpp:16-43).
 * There is NO CPU fallback (the reference's src/main.cpp:98-113 fallback is deliberately removed):
 * without a usable sm_100 device b2pt_create fails.
 *
 * Threading: like the reference (src/gpu/optix_renderer.cu:439-451) one context is driven by one
 * host thread; calls are synchronous unless the name ends in _async.
 */
#ifndef B2PT_H
#define B2PT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2PT_OK 0
#define B2PT_ERR_INVALID (-1)   /* bad argument / call order (e.g. render before upload) */
#define B2PT_ERR_CUDA (-2)      /* a CUDA runtime call failed; see b2pt_last_error */
#define B2PT_ERR_NO_DEVICE (-3) /* no CUDA device, or the device is not sm_100 */

typedef struct b2pt_ctx b2pt_ctx;

/* MaterialType, reference include/material.hpp:6-10 */
enum { B2PT_DIFFUSE = 0, B2PT_SPECULAR = 1, B2PT_DIELECTRIC = 2 };

/* Material, reference include/material.hpp:12-18 (GPUMaterial, include/gpu/optix_types.hpp:36-46). */
typedef struct b2pt_material {
    int32_t type;
    float albedo[3];
    float roughness;
    float metallic;
    float ior;
    float _pad;
} b2pt_material;

/* Light, reference include/scene.hpp:21-37 (GPULight, optix_types.hpp:49-57). */
typedef struct b2pt_light {
    float position[3];
    float color[3];
    float intensity;
} b2pt_light;

/* What OptixRenderer::render reads from Camera (include/camera.hpp:32-36, optix_renderer.cu:357-363):
 * position / forward / right / up exactly as Camera's ctor computed them, and the vertical fov in
 * degrees.  The viewport maths of Camera::getRay (camera.hpp:18-29, fixed 16:9) happens inside. */
typedef struct b2pt_camera {
    float position[3];
    float forward[3];
    float right[3];
    float up[3];
    float fov;
} b2pt_camera;

/* OptixRenderer::Settings, include/gpu/optix_renderer.hpp:14-24 (defaults 800/450/10/3/2.2). */
typedef struct b2pt_settings {
    int32_t width;
    int32_t height;
    int32_t samples_per_pixel;
    int32_t max_bounces;
    float gamma;
} b2pt_settings;

/* Which part of the frame this context renders (multi-GPU).  Pixels are dealt out in RUNS of tile_size*tile_size
 * consecutive pixels in row-major order (run k = pixels [k*A, (k+1)*A), A = tile_size^2 — scanline runs, not 2D tiles);
 * run k belongs to rank k % tile_world.  Samples [sample_begin, sample_begin+sample_count) of every owned pixel are
 * traced and the per-pixel sum is divided by settings.samples_per_pixel.  Pixels not owned are written as 0, so the
 * per-rank buffers combine with one sum-reduce (or a gather).  All-zero struct == whole frame, all samples.
 * A pass that does not cover all samples never writes the magenta "no valid sample" colour (renderer.hpp:78): that
 * decision needs the whole sample set (b2pt_progressive_pass makes it on the last pass). */
typedef struct b2pt_partition {
    int32_t tile_rank;
    int32_t tile_world;   /* 0 or 1 => no tile split */
    int32_t tile_size;    /* 0 => 32 */
    int32_t sample_begin;
    int32_t sample_count; /* 0 => settings.samples_per_pixel - sample_begin */
} b2pt_partition;

typedef struct b2pt_config {
    int32_t device;       /* CUDA device ordinal */
    int32_t flags;        /* B2PT_FLAG_* */
    int64_t max_paths_in_flight; /* wavefront batch size; 0 => default */
} b2pt_config;

#define B2PT_FLAG_COUNT_FETCHES 1  /* instrumented traversal: count node / triangle fetches (slower) */
#define B2PT_FLAG_EXACT_ONLY 2     /* closest-hit queries use only the exact reference-order DFS kernel */
#define B2PT_FLAG_LANE_KERNELS 4    /* occlusion queries of incoherent batches use the one-ray-per-lane kernels instead of the phase-split pool kernels */
#define B2PT_FLAG_POOL_EXTEND 16    /* closest-hit queries of unordered batches (b2pt_trace_closest; renderer with B2PT_FLAG_NO_SORT) use the pool kernels too (measured: no faster) */
#define B2PT_FLAG_NO_LEARN_ORDER 8  /* do not re-order wide-node children by measured occlusion rate after the first batch */
#define B2PT_FLAG_NO_SORT 32        /* renderer: do not put the paths of a bounce in the Morton order of their hit points (no effect on the image) */

/* Counters of the last trace / render call. */
typedef struct b2pt_stats {
    int64_t extend_rays;      /* closest-hit queries traced */
    int64_t shadow_rays;      /* any-hit queries traced */
    int64_t samples;          /* camera paths started */
    int64_t fallback_rays;    /* closest-hit queries re-run by the exact DFS kernel */
    int64_t node_fetches;     /* wide-node fetches (only with B2PT_FLAG_COUNT_FETCHES) */
    int64_t tri_fetches;      /* triangle fetches   (only with B2PT_FLAG_COUNT_FETCHES) */
    int64_t kernel_launches;  /* kernels launched by the call */
    double gpu_seconds;       /* CUDA-event time of the call's device work */
    double trace_seconds;     /* CUDA-event time spent in extend + shadow traversal kernels */
    double build_seconds;     /* (upload) acceleration-structure build, device time */
    double extend_seconds;    /* part of trace_seconds spent in closest-hit kernels (renderer: the k_extend_* launches alone) */
    double shadow_seconds;    /* part of trace_seconds spent in the direct-light / any-hit kernels */
    int64_t extend_launches;  /* closest-hit kernel launches (fast kernel only) */
    int64_t shadow_launches;  /* direct-light / any-hit kernel launches */
    double order_seconds;     /* renderer: between closest hit and shadow rays of every bounce — exact fallback, hit-point sort, k_hitinfo */
} b2pt_stats;

/* ---- lifecycle -------------------------------------------------------------------------------- */

/* OptixRenderer ctor + initialize() (optix_renderer.cu:85-101). */
int b2pt_create(const b2pt_config* cfg, b2pt_ctx** out);
/* ~OptixRenderer / cleanup() (optix_renderer.cu:482-515). */
void b2pt_destroy(b2pt_ctx* ctx);
/* Text of the last failure on this context (ctx == NULL: last b2pt_create failure). */
const char* b2pt_last_error(const b2pt_ctx* ctx);

/* ---- scene ------------------------------------------------------------------------------------ */

/* Host-side restatement of the reference's BVH::build ordering (include/bvh.hpp:27-72): writes
 * order[p] = input index of the triangle that ends up at position p.  Scene::loadFromObj runs this
 * on the host in the reference too (src/scene.cpp:290) — it is scene preparation, not the hot path.
 * pos: ntri*9 floats. */
int b2pt_reference_order(const float* pos, int64_t ntri, int32_t* order);

/* OptixRenderer::uploadScene (optix_renderer.cu:383-409) + buildAccelerationStructure (:233-353).
 * Arrays are in the reference's POST-build order (what scene.getTriangles() returns after
 * Scene::loadFromObj): that order *is* the reference tree, which bit-exact hit ids are defined
 * against.  pos, nrm: ntri*9 floats (v0 v1 v2 / n0 n1 n2); mat: ntri material ids.
 * Copies everything; the caller's buffers are not referenced after return. */
int b2pt_upload_scene(b2pt_ctx* ctx, const float* pos, const float* nrm, const int32_t* mat, int64_t ntri,
                      const b2pt_material* mats, int32_t nmat, const b2pt_light* lights, int32_t nlight);

/* ---- queries (replace Scene::intersect, include/scene.hpp:96-99 -> bvh.hpp:37-116) --------------- */

/* Closest hit for n rays given in HOST memory.  o, d: n*3 floats; d is normalised inside exactly as
 * the Ray ctor does (include/ray.hpp:11-12); tMin = 0.001; tmax: n floats or NULL (+inf).
 * tri[i] = position (post-build order) of the triangle the reference returns, or -1; t[i] = its
 * distance (+inf on miss); uv (n*2, may be NULL) = barycentrics.  Bit-exact vs the reference. */
int b2pt_trace_closest(b2pt_ctx* ctx, const float* o, const float* d, const float* tmax, int64_t n,
                       int32_t* tri, float* t, float* uv);
/* Boolean form used for shadow rays (renderer.hpp:274-278). occluded[i] in {0,1}. */
int b2pt_trace_any(b2pt_ctx* ctx, const float* o, const float* d, const float* tmax, int64_t n,
                   uint8_t* occluded);
/* Same, all pointers in DEVICE memory of the context's device (d_tmax / d_uv may be NULL).  The context works on its own
 * non-blocking stream (b2pt_stream): device buffers passed to any *_device entry point must be READY (their producers
 * finished — synchronise the producing stream first) and must not be in use on other streams until the call returns;
 * every entry point is synchronous, so results are complete on return. */
int b2pt_trace_closest_device(b2pt_ctx* ctx, const float* d_o, const float* d_d, const float* d_tmax, int64_t n,
                              int32_t* d_tri, float* d_t, float* d_uv);
int b2pt_trace_any_device(b2pt_ctx* ctx, const float* d_o, const float* d_d, const float* d_tmax, int64_t n,
                          uint8_t* d_occluded);

/* ---- render (replaces OptixRenderer::render, optix_renderer.cu:420-457) ------------------------- */

/* Renders into HOST memory: rgb = width*height*3 floats, row 0 = bottom of the view (v≈0), i.e. the
 * reference's frameBuffer layout (renderer.hpp:81).  part may be NULL (whole frame). */
int b2pt_render(b2pt_ctx* ctx, const b2pt_camera* cam, const b2pt_settings* settings, uint64_t seed,
                const b2pt_partition* part, float* rgb);
/* Same with the output in DEVICE memory (stays resident; used by multi-GPU reduce and bench). */
int b2pt_render_device(b2pt_ctx* ctx, const b2pt_camera* cam, const b2pt_settings* settings, uint64_t seed,
                       const b2pt_partition* part, float* d_rgb);

/* Progressive, resumable accumulation (generalises the single render call of src/main.cpp:87-92).  _begin fixes camera,
 * settings and seed and clears the accumulation; every _pass traces the next `sample_count` samples of every pixel
 * (clamped to what is left of settings.samples_per_pixel), adds them to per-pixel sums that stay on the device and
 * returns the running mean in rgb (host, may be NULL) and the samples done so far.  Samples are added in sample order,
 * so after the last pass the frame is BIT-IDENTICAL to one b2pt_render call, whatever the pass sizes.  A b2pt_render /
 * b2pt_upload_scene in between invalidates the accumulation (call _begin again). */
int b2pt_progressive_begin(b2pt_ctx* ctx, const b2pt_camera* cam, const b2pt_settings* settings, uint64_t seed);
int b2pt_progressive_pass(b2pt_ctx* ctx, int32_t sample_count, float* rgb, int32_t* samples_done);

/* Output stage of Renderer::saveImage (src/renderer.cpp:8-17): clamp -> pow(c, 1/gamma) -> (unsigned char)(c * 255), on
 * the device and BYTE-EXACT: the byte boundaries are found with the host's own powf (b2pt_tonemap_thresholds) and the
 * device only compares.  d_rgb: width*height*3 floats (device); rgb8: width*height*3 bytes (host).  flip = 0 keeps the
 * reference's row order (row 0 = bottom of the view: the reference's PNG is upside down), flip = 1 writes rows top-down. */
int b2pt_tonemap(b2pt_ctx* ctx, const float* d_rgb, int32_t width, int32_t height, float gamma, int32_t flip, uint8_t* rgb8);
/* Same for the frame of the last b2pt_render / b2pt_progressive_pass of this context, which stays resident on the
 * device (what OptixRenderer::saveImage downloads, optix_renderer.cu:459-462). */
int b2pt_tonemap_last(b2pt_ctx* ctx, float gamma, int32_t flip, uint8_t* rgb8);
/* thr256[k] = smallest float in [0,1] that the reference's tonemap maps to a byte >= k (host only, no GPU needed). */
int b2pt_tonemap_thresholds(float gamma, float* thr256);

/* ---- several GPUs behind one renderer object (one process; OptixRenderer is one object too) ----------
 * The scene is replicated, the frame is split into interleaved runs of 1024 pixels (b2pt_partition, tile_size 32), every
 * device renders its runs on its own host thread, and device 0 gathers the other devices' runs over NVLink peer loads
 * (staged copies without peer access).  The frame is bit-identical for every device count. */
typedef struct b2pt_multi b2pt_multi;
/* devices: ndev CUDA ordinals (ndev <= 0 or NULL: every visible device; an ordinal may repeat). */
int b2pt_multi_create(const int32_t* devices, int32_t ndev, int32_t flags, int64_t max_paths_in_flight, b2pt_multi** out);
void b2pt_multi_destroy(b2pt_multi* m);
const char* b2pt_multi_last_error(const b2pt_multi* m);   /* m == NULL: last b2pt_multi_create failure */
int32_t b2pt_multi_device_count(const b2pt_multi* m);
b2pt_ctx* b2pt_multi_ctx(const b2pt_multi* m, int32_t i);  /* the per-device context (stats, queries) */
int b2pt_multi_upload_scene(b2pt_multi* m, const float* pos, const float* nrm, const int32_t* mat, int64_t ntri,
                            const b2pt_material* mats, int32_t nmat, const b2pt_light* lights, int32_t nlight);
/* rgb: host, width*height*3 floats (may be NULL: the frame stays on device 0 for b2pt_multi_tonemap_last). */
int b2pt_multi_render(b2pt_multi* m, const b2pt_camera* cam, const b2pt_settings* settings, uint64_t seed, float* rgb);
int b2pt_multi_tonemap_last(b2pt_multi* m, float gamma, int32_t flip, uint8_t* rgb8);
/* Counters summed over the devices, times of the slowest device. */
int b2pt_multi_get_stats(const b2pt_multi* m, b2pt_stats* out);

/* ---- introspection ------------------------------------------------------------------------------ */
int b2pt_get_stats(const b2pt_ctx* ctx, b2pt_stats* out);
/* The acceleration structure built by the last upload: out[0]=wide nodes, out[1]=bytes per wide node,
 * out[2]=reference leaves, out[3]=reference binary nodes, out[4]=triangle bytes, out[5]=clustering (PLOC)
 * iterations run as separate launches, out[6]=levels of the wide tree, out[7]=reference leaves hoisted out of the tree (tested by every ray). */
int b2pt_get_accel_info(const b2pt_ctx* ctx, int64_t* out8);
/* The CUDA stream all of the context's kernels are launched on (cudaStream_t as void*). */
void* b2pt_stream(const b2pt_ctx* ctx);
/* Library version string. */
const char* b2pt_version(void);

#ifdef __cplusplus
}
#endif
#endif /* B2PT_H */
