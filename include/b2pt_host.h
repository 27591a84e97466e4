/* b2pt_host.h — C entry points onto the host-side scene preparation that sits in front of the GPU
 * engine: the OBJ+MTL loader with the reference's scene normalisation (reference
 * src/scene.cpp:8-293), the Camera constructor (include/camera.hpp:9-16) and the PNG writer used
 * by saveImage (src/renderer.cpp:19).  Exported from libb2pt.so next to include/b2pt.h. */
#ifndef B2PT_HOST_H
#define B2PT_HOST_H

#include "b2pt.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b2pt_scene b2pt_scene;

/* Scene() + Scene::loadFromObj(path): room + normalised model + material rules, triangles left in
 * the reference's post-BVH-build order.  Fails (B2PT_ERR_INVALID) if the OBJ cannot be read. */
int b2pt_scene_load_obj(const char* path, b2pt_scene** out);
/* The same through a binary scene cache (cache_path NULL: <path>.b2ptscene): the finished scene — triangles in
 * post-build order, build order, materials — is reused while the OBJ's size and modification time are the ones
 * recorded in the cache, and rebuilt + rewritten otherwise (parsing + BVH::build ordering of a 10M-triangle file takes
 * seconds; reading the cache is a sequential read). */
int b2pt_scene_load_obj_cached(const char* path, const char* cache_path, b2pt_scene** out);
void b2pt_scene_free(b2pt_scene* scene);
int64_t b2pt_scene_num_triangles(const b2pt_scene* scene);
int32_t b2pt_scene_num_materials(const b2pt_scene* scene);
int32_t b2pt_scene_num_lights(const b2pt_scene* scene);
/* scene.getTriangles(): pos/nrm ntri*9 floats, mat ntri ints (any may be NULL); order (may be NULL)
 * = for each post-build position, the triangle's index in the loader's pre-build list. */
int b2pt_scene_get_triangles(const b2pt_scene* scene, float* pos, float* nrm, int32_t* mat, int32_t* order);
int b2pt_scene_get_materials(const b2pt_scene* scene, b2pt_material* mats);
int b2pt_scene_get_lights(const b2pt_scene* scene, b2pt_light* lights);

/* Loader self-check: parses the OBJ twice — with the chunked multi-threaded parser (nthreads <= 0: all host
 * threads; chunk_bytes: smallest chunk, a few bytes puts every construct on a chunk boundary) and line by line on
 * one thread — and compares every array.  0 = identical, > 0 = which array differs first, < 0 = cannot open. */
int b2pt_obj_parser_selfcheck(const char* path, int32_t nthreads, int64_t chunk_bytes);

/* Camera(position, target, up, fov) as the reference constructs it. */
int b2pt_camera_look_at(const float* position, const float* target, const float* up, float fov, b2pt_camera* out);

/* 8-bit RGB PNG, rows in the order given. */
int b2pt_write_png(const char* path, int32_t width, int32_t height, const uint8_t* rgb8);

/* Linear float framebuffer as a portable float map (.pfm, little-endian RGB); rows in the order given, which for the
 * reference's frameBuffer (row 0 = bottom of the view) is exactly the PFM convention. */
int b2pt_write_pfm(const char* path, int32_t width, int32_t height, const float* rgb);

#ifdef __cplusplus
}
#endif
#endif /* B2PT_HOST_H */
