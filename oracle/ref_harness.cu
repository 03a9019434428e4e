// Headless harness around the UNMODIFIED reference sources (kernel.cu, sunsky.cu, Scene.cpp, ...).
//
// TEST INFRASTRUCTURE ONLY (oracle/): built by oracle/Makefile into oracle/_ref/libbrickmap_ref_<variant>.so
// from the sources where they lie under /root/reference/src (or, for variants that need other compile-time
// constants, from a sed-patched build-time copy of variables.h under oracle/_ref/, never committed).
// Only tests/, __graft_entry__.smoke() and bench.py --impl reference may load it.
//
// What it adds on top of the reference (all of it is glue, none of it is the algorithm):
//   * the symbols main.cpp/camera.cpp/interop.cpp would have defined (camera, cuda_interop stubs),
//   * a real cudaArray-backed surface, because launch_kernels always launches blit_onto_framebuffer
//     (kernel.cu:428),
//   * extern "C" entry points to drive launch_kernels (kernel.cu:366), to launch the reference's own
//     __global__ kernels one stage at a time (kernel.cu:412-420), and to read back queues/counters,
//   * ref_force_resident(): uploads every host brick in host order (Scene.cpp:104 slot numbering) so the
//     traversal can be compared on a fully resident scene without running the streaming loop first.
#include "stdafx.h"
#include "sunsky.cuh"
#include "state.h"
#include "launch.h"
#include <vector>
#include <cstring>

// ---- symbols the reference expects from translation units we do not build ------------------------
Camera camera;                                   // camera.cpp:56
cuda_interop::cuda_interop() : width(0), height(0), fb(0), rb(0), surf(0) {}
cuda_interop::~cuda_interop() {}
cudaError cuda_interop::set_size(const int w, const int h) { width = w; height = h; return cudaSuccess; }
void cuda_interop::blit() {}

// BM_DROPIN: the same harness linked against integration/launch_kernels_dropin.cpp + libbrickmap_b200.so INSTEAD of the
// reference's kernel.cu/sunsky.cu: the reference HOST (Scene.cpp, State, the main-loop body) drives the new kernels.
// Everything that touches the reference's own kernels or device counters is compiled out in that variant.
#ifndef BM_DROPIN
// ---- reference device symbols/kernels (external linkage under -rdc) ------------------------------
extern __device__ unsigned int primary_ray_cnt;   // kernel.cu:106
extern __device__ unsigned int start_position;    // kernel.cu:109
extern __device__ unsigned int raynr_primary;     // kernel.cu:111
extern __device__ unsigned int raynr_extend;      // kernel.cu:113
extern __device__ unsigned int raynr_shade;       // kernel.cu:115
extern __device__ unsigned int raynr_connect;     // kernel.cu:117
extern __device__ unsigned int shadow_ray_cnt;    // kernel.cu:119

__global__ void set_wavefront_globals(uint32_t render_width, uint32_t render_height);
__global__ void upload(Scene::GPUScene scene);
__global__ void primary_rays(RayQueue* ray_buffer, glm::vec3 camera_right, glm::vec3 camera_up, glm::vec3 camera_direction, glm::vec3 O, unsigned int frame, float focalDistance, float lens_radius, Scene::GPUScene scene, glm::vec4* blit_buffer, glm::ivec3 camera_position, uint32_t render_width, uint32_t render_height);
__global__ void extend(RayQueue* ray_buffer, Scene::GPUScene scene, glm::ivec3 camera_position);
__global__ void shade(RayQueue* ray_buffer, RayQueue* ray_buffer_next, ShadowQueue* shadowQueue, Scene::GPUScene scene, glm::vec4* blit_buffer, unsigned int frame);
__global__ void connect(ShadowQueue* queue, Scene::GPUScene scene, glm::vec4* blit_buffer, glm::ivec3 camera_position);

// sky evaluation through the reference's own device functions (sunsky.cu:32,76,116)
__global__ void harness_eval_sky(int n, const float* dirs, int mode, float* out) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const glm::vec3 d(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]);
	glm::vec3 c = mode == 0 ? sun(d) : (mode == 1 ? sky(d) : sunsky(d));
	out[3 * i] = c.x; out[3 * i + 1] = c.y; out[3 * i + 2] = c.z;
}
#endif  // !BM_DROPIN

// sum of the alpha channel of the accumulation buffer = number of finished paths (kernel.cu:301,322); used by the
// benchmark's reference arm to find how many frames make N samples per pixel. Harness glue, not the algorithm.
__global__ void harness_alpha_sum(const glm::vec4* buf, size_t n, double* out) {
	double acc = 0.0;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) acc += buf[i].a;
	for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xFFFFFFFFu, acc, o);
	if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

namespace {
State* g_state = nullptr;
Scene* g_scene = nullptr;
cudaArray_t g_array = nullptr;
cudaSurfaceObject_t g_surf = 0;
bool g_owns_storage = true; // false after ref_force_resident replaced the per-superchunk brick arrays

#define HCHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "ref_harness: %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return (int)e_; } } while (0)

void camera_basis(glm::vec3& right, glm::vec3& up) {
	// same expressions as kernel.cu:384-385
	right = glm::normalize(glm::cross(camera.direction, camera.up)) * 1.5f * ((float)g_state->screen_width / g_state->screen_height);
	up = glm::normalize(glm::cross(right, camera.direction)) * 1.5f;
}
} // namespace

extern "C" {

// out[0..8] = grid_size, grid_height, ray_queue_buffer_size, lod_2x2x2, lod_8x8x8, brick_load_queue_size,
//             supergrid_starting_size, sizeof(RayQueue), sizeof(ShadowQueue)
int ref_constants(int64_t* out) {
	out[0] = grid_size; out[1] = grid_height; out[2] = ray_queue_buffer_size; out[3] = lod_distance_2x2x2;
	out[4] = lod_distance_8x8x8; out[5] = brick_load_queue_size; out[6] = supergrid_starting_size;
	out[7] = sizeof(RayQueue); out[8] = sizeof(ShadowQueue);
	return 0;
}

int ref_init(int device, int width, int height) {
	HCHECK(cudaSetDevice(device));
	cudaDeviceProp props;
	HCHECK(cudaGetDeviceProperties(&props, device));
	sm_cores = props.multiProcessorCount; // main.cpp:97
	g_state = new State(width, height);
	cudaChannelFormatDesc desc = cudaCreateChannelDesc<float4>();
	HCHECK(cudaMallocArray(&g_array, &desc, width, height, cudaArraySurfaceLoadStore));
	cudaResourceDesc rd;
	memset(&rd, 0, sizeof(rd));
	rd.resType = cudaResourceTypeArray;
	rd.res.array.array = g_array;
	HCHECK(cudaCreateSurfaceObject(&g_surf, &rd));
	g_state->interop.surf = g_surf;
	g_scene = new Scene(); // creates load_stream and kernel_stream (Scene.cpp:34-35)
	HCHECK(cudaMemset(g_state->blit_buffer, 0, (size_t)width * height * sizeof(glm::vec4)));
	HCHECK(cudaMemset(g_state->ray_buffer_work, 0, (size_t)ray_queue_buffer_size * sizeof(RayQueue)));
	HCHECK(cudaMemset(g_state->ray_buffer_next, 0, (size_t)ray_queue_buffer_size * sizeof(RayQueue)));
	HCHECK(cudaMemset(g_state->shadow_queue_buffer, 0, (size_t)ray_queue_buffer_size * sizeof(ShadowQueue)));
	return 0;
}

int ref_generate() {
	g_scene->generate(); // Scene.cpp:118
	return (int)cudaDeviceSynchronize();
}

// Number of superchunks and, per superchunk, the host brick count (for sizing read-backs).
int ref_supergrid_count() { return (int)g_scene->supergrid.size(); }
int ref_host_brick_counts(int* out) {
	for (size_t i = 0; i < g_scene->supergrid.size(); i++) out[i] = (int)g_scene->supergrid[i]->bricks.size();
	return 0;
}
// Host-side scene as generated by the reference (Scene.cpp:44-116): index words and bricks of one superchunk.
int ref_host_supercell(int sc, uint32_t* indices_out, uint32_t* bricks_out) {
	const auto& s = g_scene->supergrid[sc];
	memcpy(indices_out, s->indices.data(), s->indices.size() * sizeof(uint32_t));
	if (bricks_out && !s->bricks.empty()) memcpy(bricks_out, s->bricks.data(), s->bricks.size() * sizeof(Brick));
	return 0;
}

// Make every brick resident in host order: slot == host slot, index word == host index word.
int ref_force_resident() {
	for (size_t i = 0; i < g_scene->supergrid.size(); i++) {
		auto& s = g_scene->supergrid[i];
		const size_t n = s->bricks.size();
		if (n > 0) {
			Brick* fresh = nullptr;
			size_t cap = supergrid_starting_size;
			while (cap < n + 1) cap *= 2;
			HCHECK(cudaMalloc(&fresh, cap * sizeof(Brick)));
			HCHECK(cudaMemcpy(fresh, s->bricks.data(), n * sizeof(Brick), cudaMemcpyHostToDevice));
			HCHECK(cudaFree(s->gpu_brick_location));
			s->gpu_brick_location = fresh;
			s->gpu_count = (int)cap;
			s->gpu_index_highest = (int)n;
			HCHECK(cudaMemcpy(g_scene->gpuScene.bricks + i, &fresh, sizeof(Brick*), cudaMemcpyHostToDevice));
		}
		HCHECK(cudaMemcpy(s->gpu_indices_location, s->indices.data(), s->indices.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
	}
	HCHECK(cudaMemset(g_scene->gpuScene.brick_load_queue_count, 0, 4));
	return (int)cudaDeviceSynchronize();
}

// The six device pointers of Scene::GPUScene (Scene.h:9-17), in declaration order.
int ref_get_scene(void** out) {
	const Scene::GPUScene& s = g_scene->gpuScene;
	out[0] = s.indices; out[1] = s.bricks; out[2] = s.brick_load_queue; out[3] = s.brick_load_queue_count;
	out[4] = s.bricks_queue; out[5] = s.indices_queue;
	return 0;
}
int ref_get_state(void** out) {
	out[0] = g_state->ray_buffer_work; out[1] = g_state->ray_buffer_next; out[2] = g_state->shadow_queue_buffer; out[3] = g_state->blit_buffer;
	return 0;
}

int ref_set_camera(const float* pos, const float* dir, const float* up, float focal, float lens) {
	camera.position = glm::vec3(pos[0], pos[1], pos[2]);
	camera.direction = glm::vec3(dir[0], dir[1], dir[2]);
	camera.up = glm::vec3(up[0], up[1], up[2]);
	camera.focalDistance = focal;
	camera.lensRadius = lens;
	return 0;
}
int ref_set_sun(float x, float y) {
	sun_position = glm::vec2(x, y);
	sun_position_changed = true;
	return 0;
}
int ref_mark_sun_changed() { sun_position_changed = true; return 0; }

// One iteration of the reference main loop body (main.cpp:142-146).
int ref_frame(int process_queue) {
	launch_kernels(*g_state, g_state->interop.surf, g_state->blit_buffer, g_scene->gpuScene, g_state->ray_buffer_work, g_state->ray_buffer_next, g_state->shadow_queue_buffer);
	if (process_queue) g_scene->process_load_queue();
	std::swap(g_state->ray_buffer_work, g_state->ray_buffer_next);
	return (int)cudaDeviceSynchronize();
}
int ref_process_load_queue() { g_scene->process_load_queue(); return (int)cudaDeviceSynchronize(); }

// Timed loop for the reference arm of the benchmark: `frames` iterations of the main-loop body, CUDA-event
// timed on the legacy default stream (which serialises with the reference's blocking streams).
// shadow_total receives the sum of shadow_ray_cnt over the frames (rays = frames*N + shadow_total).
int ref_run_frames(int frames, int process_queue, float* ms_out, uint64_t* shadow_total) {
	cudaEvent_t e0, e1;
	HCHECK(cudaEventCreate(&e0));
	HCHECK(cudaEventCreate(&e1));
	uint64_t total = 0;
	HCHECK(cudaDeviceSynchronize());
	HCHECK(cudaEventRecord(e0, 0));
	for (int f = 0; f < frames; f++) {
		launch_kernels(*g_state, g_state->interop.surf, g_state->blit_buffer, g_scene->gpuScene, g_state->ray_buffer_work, g_state->ray_buffer_next, g_state->shadow_queue_buffer);
#ifndef BM_DROPIN
		unsigned int sc = 0;
		HCHECK(cudaMemcpyFromSymbol(&sc, shadow_ray_cnt, 4)); // device is idle here (kernel.cu:431)
		total += sc;
#endif
		if (process_queue) g_scene->process_load_queue();
		std::swap(g_state->ray_buffer_work, g_state->ray_buffer_next);
	}
	HCHECK(cudaEventRecord(e1, 0));
	HCHECK(cudaEventSynchronize(e1));
	HCHECK(cudaEventElapsedTime(ms_out, e0, e1));
	*shadow_total = total;
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	return 0;
}

#ifndef BM_DROPIN
// counters out[0..6] = primary_ray_cnt, start_position, raynr_primary, raynr_extend, raynr_shade, raynr_connect, shadow_ray_cnt
int ref_read_counters(uint32_t* out) {
	HCHECK(cudaDeviceSynchronize());
	HCHECK(cudaMemcpyFromSymbol(out + 0, primary_ray_cnt, 4));
	HCHECK(cudaMemcpyFromSymbol(out + 1, start_position, 4));
	HCHECK(cudaMemcpyFromSymbol(out + 2, raynr_primary, 4));
	HCHECK(cudaMemcpyFromSymbol(out + 3, raynr_extend, 4));
	HCHECK(cudaMemcpyFromSymbol(out + 4, raynr_shade, 4));
	HCHECK(cudaMemcpyFromSymbol(out + 5, raynr_connect, 4));
	HCHECK(cudaMemcpyFromSymbol(out + 6, shadow_ray_cnt, 4));
	return 0;
}
int ref_write_counters(const uint32_t* in) {
	HCHECK(cudaMemcpyToSymbol(primary_ray_cnt, in + 0, 4));
	HCHECK(cudaMemcpyToSymbol(start_position, in + 1, 4));
	HCHECK(cudaMemcpyToSymbol(raynr_primary, in + 2, 4));
	HCHECK(cudaMemcpyToSymbol(raynr_extend, in + 3, 4));
	HCHECK(cudaMemcpyToSymbol(raynr_shade, in + 4, 4));
	HCHECK(cudaMemcpyToSymbol(raynr_connect, in + 5, 4));
	HCHECK(cudaMemcpyToSymbol(shadow_ray_cnt, in + 6, 4));
	return 0;
}
// Sun globals normally set inside launch_kernels (kernel.cu:374-375, 392-394); needed before stage-wise runs.
int ref_upload_sun() {
	float sun_angular = cos(sunSize * pi / 180.f);
	HCHECK(cudaMemcpyToSymbol(sunAngularDiameterCos, &sun_angular, sizeof(float)));
	HCHECK(cudaMemcpyToSymbol(SunPos, &sun_position, sizeof(glm::vec2)));
	glm::vec3 sun_direction = glm::normalize(fromSpherical((sun_position - glm::vec2(0.0, 0.5)) * glm::vec2(6.28f, 3.14f)));
	HCHECK(cudaMemcpyToSymbol(sunDirection, &sun_direction, sizeof(glm::vec3)));
	return 0;
}
#endif  // !BM_DROPIN
int ref_sun_direction(float* out) {
	glm::vec3 d = glm::normalize(fromSpherical((sun_position - glm::vec2(0.0, 0.5)) * glm::vec2(6.28f, 3.14f)));
	out[0] = d.x; out[1] = d.y; out[2] = d.z;
	return 0;
}

// which: 0 = work buffer, 1 = next buffer. Records are the reference's 64-byte RayQueue (variables.h:43-52).
int ref_read_rays(int which, void* out, size_t first, size_t n) {
	HCHECK(cudaDeviceSynchronize());
	const RayQueue* src = which == 0 ? g_state->ray_buffer_work : g_state->ray_buffer_next;
	HCHECK(cudaMemcpy(out, src + first, n * sizeof(RayQueue), cudaMemcpyDeviceToHost));
	return 0;
}
int ref_write_rays(int which, const void* in, size_t first, size_t n) {
	RayQueue* dst = which == 0 ? g_state->ray_buffer_work : g_state->ray_buffer_next;
	HCHECK(cudaMemcpy(dst + first, in, n * sizeof(RayQueue), cudaMemcpyHostToDevice));
	return 0;
}
int ref_read_shadow(void* out, size_t first, size_t n) {
	HCHECK(cudaDeviceSynchronize());
	HCHECK(cudaMemcpy(out, g_state->shadow_queue_buffer + first, n * sizeof(ShadowQueue), cudaMemcpyDeviceToHost));
	return 0;
}
int ref_write_shadow(const void* in, size_t first, size_t n) {
	HCHECK(cudaMemcpy(g_state->shadow_queue_buffer + first, in, n * sizeof(ShadowQueue), cudaMemcpyHostToDevice));
	return 0;
}
int ref_read_accum(float* out) {
	HCHECK(cudaDeviceSynchronize());
	HCHECK(cudaMemcpy(out, g_state->blit_buffer, g_state->screen_width * g_state->screen_height * sizeof(glm::vec4), cudaMemcpyDeviceToHost));
	return 0;
}
int ref_alpha_sum(double* out_host) {
	double* d = nullptr;
	HCHECK(cudaMalloc(&d, sizeof(double)));
	HCHECK(cudaMemset(d, 0, sizeof(double)));
	harness_alpha_sum<<<592, 256>>>(g_state->blit_buffer, g_state->screen_width * g_state->screen_height, d);
	HCHECK(cudaGetLastError());
	HCHECK(cudaMemcpy(out_host, d, sizeof(double), cudaMemcpyDeviceToHost));
	cudaFree(d);
	return 0;
}
int ref_clear_accum() {
	HCHECK(cudaMemset(g_state->blit_buffer, 0, g_state->screen_width * g_state->screen_height * sizeof(glm::vec4)));
	return 0;
}
int ref_swap_buffers() { std::swap(g_state->ray_buffer_work, g_state->ray_buffer_next); return 0; }

#ifndef BM_DROPIN
// Launch ONE of the reference's own kernels with the arguments launch_kernels would pass (kernel.cu:416-420).
// stage: 0 primary_rays, 1 set_wavefront_globals, 2 extend, 3 shade, 4 connect, 5 upload(count)
// serial != 0 launches <<<1,1>>> so that the atomic slot assignment becomes slot-index ordered (canonical).
int ref_run_stage(int stage, int serial, unsigned int frame, int upload_count) {
	glm::vec3 right, up;
	camera_basis(right, up);
	const int blocks = serial ? 1 : sm_cores * 8;
	const int threads = serial ? 1 : 128;
	const uint32_t w = (uint32_t)g_state->screen_width, h = (uint32_t)g_state->screen_height;
	switch (stage) {
	case 0: primary_rays<<<blocks, threads>>>(g_state->ray_buffer_work, right, up, camera.direction, camera.position, frame, camera.focalDistance, camera.lensRadius, g_scene->gpuScene, g_state->blit_buffer, camera.position, w, h); break;
	case 1: set_wavefront_globals<<<1, 1>>>(w, h); break;
	case 2: extend<<<blocks, threads>>>(g_state->ray_buffer_work, g_scene->gpuScene, camera.position / 8.f); break;
	case 3: shade<<<blocks, threads>>>(g_state->ray_buffer_work, g_state->ray_buffer_next, g_state->shadow_queue_buffer, g_scene->gpuScene, g_state->blit_buffer, frame); break;
	case 4: connect<<<blocks, threads>>>(g_state->shadow_queue_buffer, g_scene->gpuScene, g_state->blit_buffer, camera.position / 8.f); break;
	case 5: if (upload_count > 0) { upload<<<1, upload_count>>>(g_scene->gpuScene); HCHECK(cudaMemset(g_scene->gpuScene.brick_load_queue_count, 0, 4)); } break;
	default: return -1;
	}
	HCHECK(cudaGetLastError());
	HCHECK(cudaDeviceSynchronize());
	return 0;
}

// Reference arm of BASELINE config 2 (primary rays only): `frames` times primary_rays -> set_wavefront_globals -> extend
// (kernel.cu:416-418) with the stock launch shapes, CUDA-event timed, no host synchronisation in between. The survivor count
// is zeroed before every frame, so that each frame generates and extends ray_queue_buffer_size fresh primaries.
int ref_run_primary_extend(int frames, unsigned int first_frame, float* ms_out) {
	glm::vec3 right, up;
	camera_basis(right, up);
	const int blocks = sm_cores * 8, threads = 128;
	const uint32_t w = (uint32_t)g_state->screen_width, h = (uint32_t)g_state->screen_height;
	cudaEvent_t e0, e1;
	HCHECK(cudaEventCreate(&e0));
	HCHECK(cudaEventCreate(&e1));
	const unsigned int zero = 0;
	HCHECK(cudaDeviceSynchronize());
	HCHECK(cudaEventRecord(e0, 0));
	for (int f = 0; f < frames; f++) {
		HCHECK(cudaMemcpyToSymbolAsync(primary_ray_cnt, &zero, 4, 0, cudaMemcpyHostToDevice, 0));
		primary_rays<<<blocks, threads>>>(g_state->ray_buffer_work, right, up, camera.direction, camera.position, first_frame + f, camera.focalDistance, camera.lensRadius, g_scene->gpuScene, g_state->blit_buffer, camera.position, w, h);
		set_wavefront_globals<<<1, 1>>>(w, h);
		extend<<<blocks, threads>>>(g_state->ray_buffer_work, g_scene->gpuScene, camera.position / 8.f);
	}
	HCHECK(cudaEventRecord(e1, 0));
	HCHECK(cudaEventSynchronize(e1));
	HCHECK(cudaEventElapsedTime(ms_out, e0, e1));
	HCHECK(cudaGetLastError());
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	return 0;
}

#endif  // !BM_DROPIN

int ref_read_load_queue(uint32_t* count, int* positions /* 3*brick_load_queue_size */) {
	HCHECK(cudaDeviceSynchronize());
	HCHECK(cudaMemcpy(count, g_scene->gpuScene.brick_load_queue_count, 4, cudaMemcpyDeviceToHost));
	HCHECK(cudaMemcpy(positions, g_scene->gpuScene.brick_load_queue, brick_load_queue_size * sizeof(glm::ivec3), cudaMemcpyDeviceToHost));
	return 0;
}

// All device index words, superchunk-major (sc * 4096 + local), read through the reference's pointer table.
int ref_read_indices(uint32_t* out) {
	HCHECK(cudaDeviceSynchronize());
	const size_t per = (size_t)supergrid_cell_size * supergrid_cell_size * supergrid_cell_size;
	for (size_t i = 0; i < g_scene->supergrid.size(); i++)
		HCHECK(cudaMemcpy(out + i * per, g_scene->supergrid[i]->gpu_indices_location, per * sizeof(uint32_t), cudaMemcpyDeviceToHost));
	return 0;
}

#ifndef BM_DROPIN
int ref_eval_sky(int n, const float* dirs_host, int mode, float* out_host) {
	float *d = nullptr, *o = nullptr;
	HCHECK(cudaMalloc(&d, (size_t)n * 3 * sizeof(float)));
	HCHECK(cudaMalloc(&o, (size_t)n * 3 * sizeof(float)));
	HCHECK(cudaMemcpy(d, dirs_host, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice));
	harness_eval_sky<<<(n + 127) / 128, 128>>>(n, d, mode, o);
	HCHECK(cudaGetLastError());
	HCHECK(cudaMemcpy(out_host, o, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost));
	cudaFree(d);
	cudaFree(o);
	return 0;
}

#endif  // !BM_DROPIN

} // extern "C"
