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

/* Which part of the frame this context renders (multi-GPU, one process per GPU).  Pixels are
 * dealt out in tile_size x tile_size tiles, tile k belongs to rank k % tile_world; samples
 * [sample_begin, sample_begin+sample_count) of every owned pixel are traced and the per-pixel sum
 * is divided by settings.samples_per_pixel.  Pixels not owned are written as 0, so the per-rank
 * buffers combine with one sum-reduce.  All-zero struct == whole frame, all samples. */
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
#define B2PT_FLAG_POOL_EXTEND 16    /* closest-hit queries of incoherent batches use the pool kernels too (measured: no faster) */
#define B2PT_FLAG_NO_LEARN_ORDER 8  /* do not re-order wide-node children by measured occlusion rate after the first batch */

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
    double extend_seconds;    /* part of trace_seconds spent in closest-hit kernels (incl. exact fallback) */
    double shadow_seconds;    /* part of trace_seconds spent in the direct-light / any-hit kernels */
    int64_t extend_launches;  /* closest-hit kernel launches (fast kernel only) */
    int64_t shadow_launches;  /* direct-light / any-hit kernel launches */
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
/* Same, all pointers in DEVICE memory of the context's device (d_tmax / d_uv may be NULL). */
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

/* Output stage of Renderer::saveImage (src/renderer.cpp:8-17): clamp -> pow(1/gamma) -> truncate to
 * 8 bit, on the device.  d_rgb: n_pixels*3 floats (device); rgb8: n_pixels*3 bytes (host). */
int b2pt_tonemap(b2pt_ctx* ctx, const float* d_rgb, int64_t n_pixels, float gamma, uint8_t* rgb8);

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
