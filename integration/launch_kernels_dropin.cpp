// Drop-in replacement for the reference's src/kernel.cu: same symbol, same signature (src/launch.h:6), same globals read
// (camera: camera.h:24; sun_position, sun_position_changed, sm_cores, kernel_stream: variables.h:37-41), forwarding to the
// C ABI of libbrickmap_b200.so. A maintainer of the reference removes kernel.cu + sunsky.cu from CMakeLists.txt:38,
// adds this file and links brickmap_b200. It is compiled INSIDE the reference's tree (it includes the reference's own
// stdafx.h), so it is not part of this repo's product build; oracle/Makefile builds it against the read-only reference
// sources to prove it links and to run the reference host (Scene.cpp, main-loop body) on top of the new kernels.
#include "stdafx.h"
#include "state.h"
#include "launch.h"
#include "sunsky.cuh"

#include "brickmap_b200.h"

cudaStream_t kernel_stream;  // kernel.cu:15 (created by Scene::Scene(), Scene.cpp:35)

glm::vec3 fromSpherical(glm::vec2 p) {  // sunsky.cu:28-30; only used by the reference's launch_kernels, kept for link compatibility
	return glm::vec3(cos(p.x) * sin(p.y), sin(p.x) * sin(p.y), cos(p.y));
}

static_assert(sizeof(RayQueue) == sizeof(bm_ray) && sizeof(ShadowQueue) == sizeof(bm_shadow), "queue records must match");
static_assert(sizeof(Brick) == sizeof(bm_brick) && sizeof(Scene::GPUScene) == sizeof(bm_gpu_scene), "scene records must match");

cudaError launch_kernels(State& state, cudaSurfaceObject_t surf, glm::vec4* blit_buffer, Scene::GPUScene gpuScene, RayQueue* queue, RayQueue* queue2,
                         ShadowQueue* shadowQueue) {
	static bm_context* ctx = nullptr;
	static Scene::GPUScene bound{};
	(void)surf;  // display blit (kernel.cu:428) is out of scope; bm_tonemap writes a plain float4 image instead
	int rc = 0;
	if (!ctx) {
		bm_config cfg;
		bm_default_config(&cfg);
		cudaGetDevice(&cfg.device);
		cfg.grid_size = grid_size;                             // variables.h:7
		cfg.grid_height = grid_height;                         // variables.h:8
		cfg.lod_distance_2x2x2 = lod_distance_2x2x2;           // variables.h:27
		cfg.lod_distance_8x8x8 = lod_distance_8x8x8;           // variables.h:25
		cfg.brick_load_queue_size = brick_load_queue_size;     // variables.h:35
		cfg.ray_queue_buffer_size = ray_queue_buffer_size;     // variables.h:61
		cfg.screen_width = (uint32_t)state.screen_width;
		cfg.screen_height = (uint32_t)state.screen_height;
		if ((rc = bm_create(&ctx, &cfg)) != 0) return rc > 0 ? (cudaError)rc : cudaErrorInvalidValue;
	}
	// GPUScene is passed by value every frame (launch.h:6); the pointer tables only change identity when the host re-creates them.
	// bm_scene_bind derives the emptiness bitmaps from the index words, so it is repeated only then. That is correct for the
	// reference host: Scene.cpp never turns an empty cell into a non-empty one after generate() (streaming rewrites non-zero
	// words only, Scene.cpp:158-164,224; growing a superchunk patches a brick-table ENTRY, Scene.cpp:242-246, which is read
	// through the table every frame). A host that edits voxels in place (empty <-> non-empty cells) under the same table pointers
	// must call bm_scene_bind itself after the edit.
	if (bound.indices != gpuScene.indices || bound.bricks != gpuScene.bricks) {
		bm_gpu_scene s;
		memcpy(&s, &gpuScene, sizeof(s));
		if ((rc = bm_scene_bind(ctx, s)) != 0) return rc > 0 ? (cudaError)rc : cudaErrorInvalidValue;
		bound = gpuScene;
	}
	bm_camera cam;
	memcpy(cam.position, &camera.position, 12);
	memcpy(cam.direction, &camera.direction, 12);
	memcpy(cam.up, &camera.up, 12);
	cam.focal_distance = camera.focalDistance;
	cam.lens_radius = camera.lensRadius;
	bm_set_camera(ctx, &cam);
	if (sun_position_changed) {  // kernel.cu:389-394
		sun_position_changed = false;
		bm_set_sun(ctx, sun_position.x, sun_position.y);
	}
	rc = bm_launch_frame(ctx, reinterpret_cast<float*>(blit_buffer), reinterpret_cast<bm_ray*>(queue), reinterpret_cast<bm_ray*>(queue2),
	                     reinterpret_cast<bm_shadow*>(shadowQueue), BM_FRAME_DEFAULT);
	if (rc != 0) return rc > 0 ? (cudaError)rc : cudaErrorInvalidValue;
	return cudaDeviceSynchronize();  // kernel.cu:431: the reference host reads the request queue with blocking copies on the default stream
}
