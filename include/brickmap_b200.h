/* brickmap_b200 -- C ABI of the B200-native BrickMap path-tracing hot path.
 *
 * This is the drop-in boundary for ONE path of stijnherfst/BrickMap: what src/launch.h:6 (`launch_kernels`)
 * does once per frame -- primary-ray generation, brick-grid DDA traversal with 2x2x2 LoD and 8x8x8 bitmask
 * bricks, hit shading / NEE / cosine bounce, sun-sky evaluation, brick request emission and staged-brick
 * upload (reference: src/kernel.cu, src/voxel.cuh, src/sunsky.cu). All citations are relative to the
 * reference tree. Plain pointers and sizes only; every pointer named "device" is a CUDA device pointer
 * owned by the caller, exactly as in the reference where State (state.h:16-22) and Scene (Scene.cpp:153-190)
 * own all storage and launch_kernels borrows it.
 *
 * Error convention: every function returns 0 on success, a positive value = cudaError_t of the failing CUDA
 * call, or a negative BM_E_* value for API misuse. Nothing calls exit() (the reference's cuda() macro does,
 * assert_cuda.cpp:3-14). bm_last_error_string() describes the last failure of the calling thread.
 */
#ifndef BRICKMAP_B200_H
#define BRICKMAP_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BM_E_INVALID (-1)   /* bad argument */
#define BM_E_STATE (-2)     /* call order violated (e.g. frame before scene bind) */
#define BM_E_NOMEM (-3)

/* ---- record layouts shared with the reference host ------------------------------------------------- */
/* RayQueue, variables.h:43-52 (64 bytes). */
typedef struct bm_ray {
	float origin[3];
	float direction[3];
	float throughput[3];
	float normal[3];
	float distance;
	int32_t identifier;
	int32_t bounces;
	uint32_t pixel_index;
} bm_ray;

/* ShadowQueue, variables.h:54-59 (40 bytes). */
typedef struct bm_shadow {
	float origin[3];
	float direction[3];
	float color[3];
	uint32_t pixel_index;
} bm_shadow;

/* Brick, Scene.h:3-5 (64 bytes): bit x + 8y + 64z of an 8x8x8 voxel block. Brick arrays must be at least 8-byte aligned (the
 * kernels read a z-slice as one 64-bit word); cudaMalloc'ed arrays of 64-byte bricks (Scene.cpp:170-176) are 64-byte aligned. */
typedef struct bm_brick {
	uint32_t data[16];
} bm_brick;

/* Scene::GPUScene, Scene.h:9-17 (48 bytes, six device pointers, passed by value). */
typedef struct bm_gpu_scene {
	uint32_t** indices;               /* [superchunks] -> 4096 index words, local order x + 16y + 256z */
	bm_brick** bricks;                /* [superchunks] -> brick array of that superchunk */
	int32_t* brick_load_queue;        /* [queue_size][3] cell coordinates (glm::ivec3) */
	uint32_t* brick_load_queue_count; /* may exceed queue_size; consumers clamp (kernel.cu:409) */
	bm_brick* bricks_queue;           /* [queue_size] staged bricks */
	uint32_t* indices_queue;          /* [queue_size] staged index words */
} bm_gpu_scene;

/* Index-word bits, variables.h:29-33. */
#define BM_BRICK_INDEX_BITS 0xFFFu
#define BM_BRICK_LOD_BITS 0xFF000u
#define BM_BRICK_LOADED_BIT 0x80000000u
#define BM_BRICK_UNLOADED_BIT 0x40000000u
#define BM_BRICK_REQUESTED_BIT 0x20000000u

/* The compile-time constants of variables.h:7-35,61 as run-time values. bm_default_config() fills in the
 * reference's: 4096 x 4096 x 512 voxels, LoD thresholds 100 000 / 600 000, queue 1024, 2 097 152 slots. */
typedef struct bm_config {
	int32_t device;            /* CUDA device ordinal */
	int32_t grid_size;         /* voxels in x and in y (variables.h:7); multiple of 128 */
	int32_t grid_height;       /* voxels in z (variables.h:8); multiple of 128 */
	int32_t lod_distance_2x2x2; /* variables.h:27, squared cell distance */
	int32_t lod_distance_8x8x8; /* variables.h:25 */
	int32_t brick_load_queue_size; /* variables.h:35 */
	uint32_t ray_queue_buffer_size; /* variables.h:61: segment slots per frame */
	uint32_t screen_width, screen_height; /* full image (state.h:13-14) */
	/* Image tile rendered by this context (multi-GPU partition, not in the reference): rows
	 * [tile_row0, tile_row0 + tile_rows) of the full image. tile_rows == 0 means the whole image. The
	 * accumulation buffer handed to bm_* then covers only the tile (tile_rows * screen_width pixels).
	 * A tile is an INDEPENDENT renderer of its rows: the pixel cursor wraps inside the tile and RayQueue.pixel_index is the index
	 * into the tile's own buffer (row_in_tile * width + x), which also feeds shade's seed (kernel.cu:252). Only the camera mapping
	 * uses the full image (kernel.cu:183-184 with the full-image row). A tiled render is therefore a different random sequence
	 * from the single-context render of the same image (same estimator, other samples); it equals, bit for bit, the reference
	 * algorithm run with the same row mapping (the CPU oracle's tile mode, tests/test_gpu_parity.py). */
	uint32_t tile_row0, tile_rows;
	/* Interleaved partition (better load balance than one contiguous band: sky rows finish their paths after one segment,
	 * terrain rows need up to four). strip_rows != 0: the image is cut into strips of strip_rows rows and this context owns
	 * strips strip_index, strip_index + strip_count, ...; tile_rows is then the number of rows it owns (tile_row0 is ignored)
	 * and row r of its accumulation buffer is image row ((r / strip_rows) * strip_count + strip_index) * strip_rows + r % strip_rows. */
	uint32_t strip_rows, strip_count, strip_index;
} bm_config;

/* Camera fields read by launch_kernels (camera.h:4-9; kernel.cu:384-387,416). */
typedef struct bm_camera {
	float position[3];
	float direction[3];
	float up[3];
	float focal_distance;
	float lens_radius;
} bm_camera;

/* The seven device counters of kernel.cu:106-119 plus the host statics of kernel.cu:367-369. */
typedef struct bm_counters {
	uint32_t primary_ray_cnt;
	uint32_t start_position;
	uint32_t shadow_ray_cnt;
	uint32_t frame; /* next frame number, starts at 1 (kernel.cu:369) */
} bm_counters;

/* Work done so far (since bm_create or bm_reset_stats), for Mrays/s and the roofline. */
typedef struct bm_stats {
	uint64_t frames;
	uint64_t extend_rays;   /* intersect_voxel calls from extend (kernel.cu:236) */
	uint64_t shadow_rays;   /* intersect_voxel calls from connect (kernel.cu:340) */
	uint64_t terminations;  /* alpha increments = finished paths (kernel.cu:301,322) */
	uint64_t unoccluded;    /* shadow rays that added light (kernel.cu:341-343) */
	uint64_t cell_steps;    /* traversal counters, only filled with BM_FRAME_COUNT_WORK: cell-grid DDA iterations, */
	uint64_t index_reads;   /*   index words actually loaded (<= cell_steps thanks to the emptiness bitmap), */
	uint64_t bricks_entered;
	uint64_t requests;
	uint64_t kernel_launches; /* kernels of this library launched */
} bm_stats;

typedef struct bm_context bm_context;

void bm_default_config(bm_config* cfg);
int bm_create(bm_context** out, const bm_config* cfg);
void bm_destroy(bm_context* ctx);
const char* bm_last_error_string(void);

/* Bind the scene handle (the GPUScene that launch_kernels receives by value, launch.h:6). Builds the derived,
 * library-private traversal aids from it: emptiness bitmaps over blocks of cells and per cell (a cell is empty iff its
 * index word is 0, which streaming never changes: Scene.cpp:158-164, kernel.cu:150), the per-column tops of the non-empty
 * cells (rays that cannot meet anything any more are ended as the misses they are) and a check whether the
 * per-superchunk index arrays form one flat arena. Must be called again only if the SET of empty cells or
 * the pointer tables' index-array entries change; brick arrays may be re-allocated freely (Scene.cpp:242-246),
 * they are always reached through the table. */
int bm_scene_bind(bm_context* ctx, bm_gpu_scene scene);

/* Host inputs of launch_kernels that the reference reads from globals (camera.h:24, variables.h:37-38).
 * Changing camera or sun marks the accumulation buffer for reset on the next frame (kernel.cu:387-403). */
int bm_set_camera(bm_context* ctx, const bm_camera* cam);
int bm_set_sun(bm_context* ctx, float sun_x, float sun_y);

/* Wavefront cursors (kernel.cu:106-119). */
int bm_get_counters(bm_context* ctx, bm_counters* out);
int bm_set_counters(bm_context* ctx, const bm_counters* in);
int bm_get_stats(bm_context* ctx, bm_stats* out);
int bm_reset_stats(bm_context* ctx);

#define BM_FRAME_DEFAULT 0u
#define BM_FRAME_NO_UPLOAD 1u   /* skip the staged-brick upload step (kernel.cu:407-414) */
#define BM_FRAME_NO_RESET 2u    /* never reset the accumulation buffer on camera/sun change */
#define BM_FRAME_COUNT_WORK 4u  /* fill the traversal work counters of bm_stats (slower) */
#define BM_FRAME_EXACT_PATHS 8u /* bm_render with target_paths: start exactly target_paths paths since the last reset -- a frame only
                                   takes as many fresh primaries as are still missing (the last frames shrink to the survivors), so
                                   with target = spp * tile pixels every pixel gets exactly spp paths and no ray is traced beyond them.
                                   Per-slot results are those of the reference algorithm run with the same per-frame slot counts. */
#define BM_FRAME_EXTEND_ONLY 16u /* bm_launch_frame: stop after extend (primary_rays, set_wavefront_globals, extend: kernel.cu:416-418).
                                   `queue` holds the post-extend record of every slot; nothing is shaded, no survivors, no shadow
                                   rays, the accumulation buffer is untouched; cursor and frame number advance as usual. */

/* One frame with the reference's buffer contract (launch_kernels, kernel.cu:366-439, minus the display blit):
 *   - applies the pending staged bricks and zeroes *brick_load_queue_count (kernel.cu:407-414),
 *   - fills slots [primary_ray_cnt, N) of `queue` with primaries (kernel.cu:154-223), advances the cursor,
 *   - extends every slot, leaving normal/distance in `queue` (kernel.cu:226-238),
 *   - shades: survivors to queue2[0, primary_ray_cnt) and shadow rays to shadow_queue[0, shadow_ray_cnt), both
 *     in slot order (= the reference scheduled one thread at a time), sky/alpha into blit_buffer,
 *   - connects the shadow rays (kernel.cu:328-346), emits brick requests (voxel.cuh:228-241),
 *   - frame++ and the device is idle on return (kernel.cu:423,431).
 * The caller swaps queue/queue2 between calls like main.cpp:146. blit_buffer is float4[w*h] (state.h:22). */
int bm_launch_frame(bm_context* ctx, float* blit_buffer_device, bm_ray* queue_device, bm_ray* queue2_device,
                    bm_shadow* shadow_queue_device, uint32_t flags);

/* Throughput path: `frames` consecutive frames with exactly the per-frame semantics above (same slots, seeds
 * and stable compaction order), but the ray and shadow state stay in library-private buffers and no
 * per-stage records are written. Stops early once `target_paths` paths have finished since the last
 * accumulation reset (0 = no target); frames the device skips that way leave every piece of state (survivors,
 * cursor, frame number) where the last executed frame put it. Asynchronous on the context's stream unless `sync` is set.
 * Streaming contract: the staged-brick upload (kernel.cu:407-414) runs ONCE, before the first frame of the call -- the
 * reference pairs one process_load_queue with one launch_kernels (main.cpp:142-143), so at most one staged batch exists
 * per call; requests of all `frames` frames accumulate in the queue for the host to stage afterwards.
 * After bm_set_counters with primary_ray_cnt > 0 the survivor records must be supplied with bm_import_rays first
 * (BM_E_STATE otherwise); bm_launch_frame takes them from `queue` instead. */
int bm_render(bm_context* ctx, float* blit_buffer_device, uint32_t frames, uint64_t target_paths, uint32_t flags, int sync);

/* Primary visibility (BASELINE config 2, "primary rays only"): `frames` times primary_rays -> set_wavefront_globals -> extend
 * (kernel.cu:416-418) over all ray_queue_buffer_size slots, every frame from an empty survivor set; nothing is shaded, the
 * accumulation buffer is not touched. queue_device receives the post-extend record of every slot (of the last frame): origin,
 * direction, hit normal and distance (1e20 = miss). Cursor and frame number advance as in bm_launch_frame. */
int bm_extend_primaries(bm_context* ctx, bm_ray* queue_device, uint32_t frames, int sync);

/* Hand the throughput path an explicit survivor set: `count` dense records (the layout bm_launch_frame leaves in queue2) become
 * the survivors of the previous frame, primary_ray_cnt = count. bm_export_rays is the inverse: the current private survivor set,
 * densely, in slot order (what the reference would hold in ray_buffer_next[0, primary_ray_cnt)). */
int bm_import_rays(bm_context* ctx, const bm_ray* queue_device, uint32_t count);
int bm_export_rays(bm_context* ctx, bm_ray* queue2_device);

/* bm_render followed by device->host copies of the results into HOST buffers: accumulation tile
 * (tile pixels * 4 floats), request count and request positions (queue_size * 3 ints; may be NULL). */
int bm_render_to_host(bm_context* ctx, float* blit_buffer_device, uint32_t frames, uint64_t target_paths, uint32_t flags,
                      float* accum_host, uint32_t* request_count_host, int32_t* request_positions_host);

/* Per-launch device timing of the frame kernel (the dominant kernel) with CUDA events on the context's stream:
 * enable, run, then read the summed duration and the number of launches it covers (the read synchronises
 * and restarts the sum). */
int bm_kernel_timing(bm_context* ctx, int enable);
int bm_kernel_time(bm_context* ctx, double* ms_sum, uint64_t* launches);

/* Request queue read-back (what Scene::process_load_queue copies to the host, Scene.cpp:202-209): count (may
 * exceed queue_size) and queue_size * 3 ints of cell coordinates, into HOST memory. Synchronises the stream. */
int bm_read_requests(bm_context* ctx, uint32_t* count_host, int32_t* positions_host);

/* ---- multi-GPU request exchange (not in the reference, which is single-GPU) ------------------
 * Every GPU renders its own image tile against a full replica of the brick store; the only state that must
 * agree between replicas is which bricks get streamed in, in which order (slot numbers are handed out in queue
 * order, Scene.cpp:224-225). Per frame each rank packs its request block {count, positions[queue_size][3]}
 * (1 + 3*queue_size int32), the blocks are all-gathered (NCCL), and every rank applies the same merge:
 * blocks in rank order, entries in queue order, duplicates dropped, at most queue_size kept. The merged list
 * replaces the local queue; the requested bit (variables.h:33) is set on cells requested by other ranks and
 * cleared on local requests that did not make the cut (the overflow path of voxel.cuh:237-240). */
int bm_requests_pack(bm_context* ctx, int32_t* block_device);
int bm_requests_merge(bm_context* ctx, const int32_t* gathered_blocks_device, int world_size);

/* Stream the context launches on (cudaStream_t as void*). */
void* bm_stream(bm_context* ctx);
int bm_synchronize(bm_context* ctx);

/* intersect_voxel (voxel.cuh:135-261) for n independent rays: the traversal alone, for parity tests and
 * microbenchmarks. origins/directions/normals are n*3 floats, distances n floats, hits n bytes (device). */
int bm_trace(bm_context* ctx, size_t n, const float* origins_device, const float* directions_device, float* normals_io_device,
             float* distances_io_device, uint8_t* hits_device);

/* sun() / sky() / sunsky() of sunsky.cu:32,76,116 for n directions (mode 0/1/2), device buffers. */
int bm_eval_sky(bm_context* ctx, size_t n, const float* dirs_device, int mode, float* out_device);

/* blit_onto_framebuffer's tone map (kernel.cu:348-364) into a plain float4 image instead of a GL surface. */
int bm_tonemap(bm_context* ctx, const float* blit_buffer_device, float* out_device);

/* ---- device-resident scene store (the GPU-side counterpart of Scene::generate, Scene.cpp:118-194) ------ */
typedef struct bm_scene_store bm_scene_store;

#define BM_SCENE_TERRAIN 0 /* Scene::generate_supercell's simplex heightfield (Scene.cpp:44-116) */
#define BM_SCENE_CAVES 1   /* sparse 3-D lattice-noise caves (not in the reference; BASELINE config 4) */
#define BM_SCENE_NONFLAT 0x100 /* OR into `kind`: lay the index arrays out NOT as one affine arena (the reference's
                                  per-superchunk cudaMalloc, Scene.cpp:170), to exercise the pointer-table path */

/* Generates the whole world on the device and lays it out exactly like the reference host does: per
 * superchunk 4096 index words and a brick array in z,y,x cell order (Scene.cpp:75-108), reachable through
 * GPUScene pointer tables. resident != 0: every brick loaded, slot == host slot (index = slot | loaded |
 * lod << 12, Scene.cpp:104). resident == 0: nothing loaded (Scene.cpp:158-164); bricks arrive through
 * bm_scene_store_stream. */
int bm_scene_store_create(bm_scene_store** out, const bm_config* cfg, int kind, uint32_t seed, int resident);
/* The same store from a world built elsewhere (a host application's Scene::supergrid, Scene.h:21-31, or a test's
 * voxel array): indices_host = superchunks * 4096 host-view index words (Scene.cpp:104), brick_counts_host one
 * entry per superchunk, bricks_host all bricks superchunk-major in host slot order. */
int bm_scene_store_create_from_host(bm_scene_store** out, const bm_config* cfg, const uint32_t* indices_host, const uint32_t* brick_counts_host,
                                    const bm_brick* bricks_host, int resident, int nonflat);
void bm_scene_store_destroy(bm_scene_store* s);
int bm_scene_store_gpu_scene(bm_scene_store* s, bm_gpu_scene* out);
/* Scene::process_load_queue (Scene.cpp:200-252): drain the request queue, stage bricks + index words for the
 * next frame's upload step. Returns the number staged through *staged (may be NULL). */
int bm_scene_store_stream(bm_scene_store* s, void* cuda_stream, uint32_t* staged);
/* Host copies for tests: index words of one superchunk as the device sees them / as generated. */
int bm_scene_store_read_indices(bm_scene_store* s, int superchunk, int host_view, uint32_t* out4096);
int bm_scene_store_brick_count(bm_scene_store* s, int superchunk, uint32_t* out);
int bm_scene_store_read_bricks(bm_scene_store* s, int superchunk, uint32_t first, uint32_t n, bm_brick* out_host);
int bm_scene_store_read_gpu_bricks(bm_scene_store* s, int superchunk, uint32_t first, uint32_t n, bm_brick* out_host);
uint64_t bm_scene_store_total_bricks(bm_scene_store* s);
const char* bm_scene_store_last_error(void);
int bm_scene_store_superchunks(bm_scene_store* s);

#ifdef __cplusplus
}
#endif
#endif
