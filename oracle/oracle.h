/* CPU oracle for the BrickMap path-tracing hot path -- C ABI (loaded with ctypes by tests/ and bench.py).
 *
 * TEST INFRASTRUCTURE ONLY. This is a scalar restatement of the reference algorithm
 * (kernel.cu, voxel.cuh, sunsky.cu, Scene.cpp, SimplexNoise.cpp), used as the checker for the CUDA
 * product and as the "port" CPU baseline of bench.py. Nothing under brickmap_b200/ may include, link
 * or call it. See oracle.cpp for the per-function reference citations and the arithmetic contract.
 */
#ifndef BRICKMAP_ORACLE_H
#define BRICKMAP_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_scene orc_scene;

/* 64-byte ray record, layout of RayQueue (variables.h:43-52). */
typedef struct orc_ray {
	float origin[3];
	float direction[3];
	float throughput[3];
	float normal[3];
	float distance;
	int32_t identifier;
	int32_t bounces;
	uint32_t pixel_index;
} orc_ray;

/* 40-byte shadow-ray record, layout of ShadowQueue (variables.h:54-59). */
typedef struct orc_shadow {
	float origin[3];
	float direction[3];
	float color[3];
	uint32_t pixel_index;
} orc_shadow;

typedef struct orc_camera {
	float position[3];
	float direction[3];
	float up[3];
	float focal_distance;
	float lens_radius;
} orc_camera;

/* Per-frame wavefront state, the seven device counters of kernel.cu:106-119 plus the host frame counter. */
typedef struct orc_frame_state {
	uint32_t primary_ray_cnt;
	uint32_t start_position;
	uint32_t shadow_ray_cnt;
	uint32_t frame; /* kernel.cu:369, starts at 1 */
} orc_frame_state;

/* Traversal work counters: summed over the rays of a call. */
typedef struct orc_stats {
	uint64_t rays;          /* intersect_voxel calls */
	uint64_t index_reads;   /* S: cell-grid DDA steps that read an index word (voxel.cuh:198) */
	uint64_t bricks;        /* K: bricks entered (voxel.cuh:223-224) */
	uint64_t lod_bytes;     /* 2x2x2 LoD DDAs entered (voxel.cuh:217) */
	uint64_t voxel_steps;   /* iterations of the 8^3 DDA loop (voxel.cuh:109) */
	uint64_t requests;      /* Q: requests enqueued (voxel.cuh:236) */
	uint64_t hits;
	uint64_t terminations;  /* P: alpha increments */
	uint64_t unoccluded;    /* V: shadow rays that added light */
	uint64_t unique_index_sectors; /* filled by orc_footprint_report */
	uint64_t unique_brick_sectors;
} orc_stats;

/* ---- scene ------------------------------------------------------------------------------------ */
/* grid_x == grid_y == grid_size of variables.h:7; grid_z == grid_height (variables.h:8). Multiples of 128. */
orc_scene* orc_scene_create(int grid_xy, int grid_z, int lod_2x2x2, int lod_8x8x8, int load_queue_size);
void orc_scene_destroy(orc_scene*);
/* Scene::generate_supercell (Scene.cpp:44-116) for every superchunk; threads <= 0 -> hardware concurrency. */
int orc_scene_generate_terrain(orc_scene*, int threads);
/* New (not in the reference): sparse cave world from an integer lattice noise, for BASELINE config 4. */
int orc_scene_generate_caves(orc_scene*, uint32_t seed, int threads);
/* Build the scene from a dense voxel occupancy array vox[x + gx*(y + gy*z)] (small worlds, tests). */
int orc_scene_from_voxels(orc_scene*, const uint8_t* vox);
/* Device-side view: 0 = nothing resident (Scene.cpp:157-164), 1 = every brick resident in host slot order. */
int orc_scene_set_residency(orc_scene*, int all_resident);
int orc_scene_supergrid_count(const orc_scene*);
int orc_scene_brick_count(const orc_scene*, int sc);            /* host bricks of one superchunk */
const uint32_t* orc_scene_host_indices(const orc_scene*, int sc); /* 4096 words */
const uint32_t* orc_scene_host_bricks(const orc_scene*, int sc);  /* brick_count*16 words */
const uint32_t* orc_scene_gpu_indices(const orc_scene*, int sc);  /* 4096 words, device-side view */
int orc_scene_gpu_brick(const orc_scene*, int sc, int slot, uint32_t* out16);
/* Request queue (voxel.cuh:228-241): count may exceed the queue size, positions hold min(count,size) entries. */
uint32_t orc_scene_queue_count(const orc_scene*);
const int32_t* orc_scene_queue_positions(const orc_scene*);
/* Scene::process_load_queue (Scene.cpp:200-229) followed by the upload kernel (kernel.cu:141-151) and the
 * count reset (kernel.cu:413). Returns the number of bricks uploaded. */
int orc_scene_stream(orc_scene*);

/* ---- traversal -------------------------------------------------------------------------------- */
/* intersect_voxel (voxel.cuh:135-261) for n rays. normal_io/distance_io are read-modify-write exactly like the
 * references the reference passes in (kernel.cu:236, 340). hit_out[i] = return value. threads>1 only when no
 * request can be emitted (fully resident scene); otherwise rays run in index order on one thread. */
int orc_trace(orc_scene*, size_t n, const float* origins, const float* directions, const int32_t cam_cell[3],
              float* normal_io, float* distance_io, uint8_t* hit_out, orc_stats* stats, int threads);

/* ---- wavefront stages (canonical = slot-index order) -------------------------------------------- */
/* Host math of launch_kernels: camera basis (kernel.cu:384-385) and sun direction (kernel.cu:393). */
void orc_camera_basis(const orc_camera*, uint32_t width, uint32_t height, float right[3], float up[3]);
void orc_sun_direction(float sun_x, float sun_y, float out[3]);
/* primary_rays (kernel.cu:154-223): fills slots [state->primary_ray_cnt, n_slots). */
void orc_primary_rays(orc_ray* rays, uint32_t n_slots, const orc_frame_state* state, const orc_camera*, uint32_t width, uint32_t height);
/* set_wavefront_globals (kernel.cu:122-139). */
void orc_set_wavefront_globals(orc_frame_state* state, uint32_t n_slots, uint32_t width, uint32_t height);
/* extend (kernel.cu:226-238) over all slots. */
void orc_extend(orc_scene*, orc_ray* rays, uint32_t n_slots, const orc_camera*, orc_stats* stats, int threads);
/* shade (kernel.cu:242-325) in slot order with stable compaction; accum is width*height*4 floats. */
void orc_shade(const orc_ray* rays, orc_ray* next, orc_shadow* shadows, uint32_t n_slots, orc_frame_state* state,
               const float sun_dir[3], float* accum, orc_stats* stats);
/* connect (kernel.cu:328-346) over shadows[0, state->shadow_ray_cnt). */
void orc_connect(orc_scene*, const orc_shadow* shadows, const orc_frame_state* state, const orc_camera*, float* accum, orc_stats* stats, int threads);
/* One whole frame = the five stages above in the order of kernel.cu:416-420, then frame++ (kernel.cu:423).
 * The caller swaps rays/next afterwards (main.cpp:146). */
void orc_frame(orc_scene*, orc_ray* rays, orc_ray* next, orc_shadow* shadows, uint32_t n_slots, orc_frame_state* state,
               const orc_camera*, float sun_x, float sun_y, uint32_t width, uint32_t height, float* accum, orc_stats* stats, int threads);

/* Image partition of the multi-GPU mode (not in the reference; include/brickmap_b200.h bm_config.tile_* / strip_*).
 * tile = {row0, rows, strip_rows, strip_count, strip_index}: the instance renders `rows` rows of the full image; kernel.cu:170-171
 * runs inside the tile (pixel_index = row_in_tile * width + x), the camera mapping kernel.cu:183-184 uses the full-image row
 * (row0 + r, or ((r / strip_rows) * strip_count + strip_index) * strip_rows + r % strip_rows). accum is rows * width * 4 floats. */
void orc_primary_rays_tiled(orc_ray* rays, uint32_t n_slots, const orc_frame_state* state, const orc_camera*, uint32_t width, uint32_t height,
                            const uint32_t tile[5]);
void orc_frame_tiled(orc_scene*, orc_ray* rays, orc_ray* next, orc_shadow* shadows, uint32_t n_slots, orc_frame_state* state, const orc_camera*,
                     float sun_x, float sun_y, uint32_t width, uint32_t height, const uint32_t tile[5], float* accum, orc_stats* stats, int threads);

/* ---- sky (sunsky.cu) ----------------------------------------------------------------------------- */
/* mode 0 = sun(), 1 = sky(), 2 = sunsky(); dirs/out are n*3 floats. */
void orc_sky_eval(size_t n, const float* dirs, int mode, const float sun_dir[3], float* out);
/* getConeSample (sunsky.cu:170-184) with the xorshift state passed in/out. */
void orc_cone_sample(const float dir[3], float extent, uint32_t* seed, float out[3]);

/* ---- footprint instrumentation -------------------------------------------------------------------- */
void orc_footprint_begin(orc_scene*);
void orc_footprint_report(orc_scene*, orc_stats* stats);

/* blit_onto_framebuffer tone map (kernel.cu:348-364): rgb/alpha then pow(1/2.2); out is width*height*4. */
void orc_tonemap(const float* accum, size_t pixels, float* out);

int orc_hardware_threads(void);

#ifdef __cplusplus
}
#endif
#endif
