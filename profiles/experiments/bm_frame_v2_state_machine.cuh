// frame_kernel_v2: the throughput form of one frame (same results as frame_kernel, bit for bit on everything geometric).
//
// SIMT-efficiency design. ncu on the simple one-thread-per-slot kernel: 11.75 of 32 lanes active per instruction; a
// first state-machine version that handled every event inline: stepping at 27 lanes but 60 % of all instructions in
// rare sections executing with 2-6 lanes. Hence:
//   * a WARP is a persistent worker; each lane is a small state machine over its current ray
//        IDLE -> STEP (tracing) <-> CELL (a non-empty cell awaits its index word / LoD decision / brick set-up)
//             -> DONE (trace ended: shade, or finish the shadow ray) -> IDLE | STEP (shadow ray)
//   * the warp runs straight-line SECTIONS and every lane only takes part in the section its state asks for:
//        STEP   one DDA iteration (advance + occupancy test) for every tracing lane, whatever level of the brickmap it is
//               at: the cell grid, an 8x8x8 brick or 2x2x2 LoD octants use the same instructions, only the tested word
//               and the bounds differ. Lanes inside a brick and lanes crossing empty cells no longer serialise
//               (voxel.cuh:222-227 nested inside :192-259);
//        CELL   index-word load, normal / entry distance, LoD selection, brick request, set-up of the finer DDA;
//        SHADE  shading + shadow-ray set-up, and REFILL: idle lanes take the next slots and set up their rays.
//     Heavy sections are DEFERRED until enough lanes wait for them (or nothing else can run), so they execute with many
//     lanes instead of one or two. A lane whose ray ends early is refilled instead of idling until the longest ray of its
//     warp ends.
//   * two emptiness bitmaps: blocks of cells in shared memory (1 LDS per step), and one bit per cell in global memory
//     (64 bits per 4x4x4 block, touched only inside non-empty blocks), so an index word is loaded only for cells that are
//     really non-empty (2.2 per ray instead of 8.5, and instead of the reference's 88).
//   * everything a lane needs only at events (world origin, direction, throughput, normal, parked cell-level DDA, ...)
//     lives in a per-lane shared-memory stash (structure of arrays, conflict-free); the stepping loop keeps 17 registers
//     of ray state.
//   * slots are handed to warps in runs of 256 by one global atomic per run; survivors are written sparse at their slot and
//     flagged in a bitmask (stable order without any block barrier).
#pragma once

namespace bm {

#ifndef BM_V2_BLOCK
#define BM_V2_BLOCK 512
#endif
#ifndef BM_V2_MIN_BLOCKS
#define BM_V2_MIN_BLOCKS 2
#endif
constexpr int kV2Block = BM_V2_BLOCK;  // upper bound of the block size (launch bounds); the launch may use fewer threads
// scheduling knobs (FrameIO.tune_*): run SHADE/REFILL once tune_shade lanes wait for it, CELL once tune_cell lanes wait for
// it, and do tune_burst TRACE iterations between two looks at the lane states

enum : int { M_IDLE = 0, M_TRACE = 1, M_CELL = 2, M_DONE = 3 };

// per-lane shared-memory stash, structure-of-arrays over the block's threads
enum : int {
	ST_OX = 0, ST_OY, ST_OZ,    // world-space ray origin (shading needs it; tracing works in cell units)
	ST_TX, ST_TY, ST_TZ,        // throughput until shade, then the shadow ray's colour
	ST_NX, ST_NY, ST_NZ,        // normal (read-modify-write target of intersect_voxel, kernel.cu:236)
	ST_PX, ST_PY, ST_PZ,        // parked cell-level DDA position while inside a brick / LoD block
	ST_MX, ST_MY, ST_MZ,        // parked cell-level tmax
	ST_CX, ST_CY, ST_CZ,        // ray origin in cell units (voxel.cuh:157)
	ST_DX, ST_DY, ST_DZ,        // ray direction
	ST_TMIN,                    // tminn (voxel.cuh:136)
	ST_ND,                      // new_distance of the cell being refined (voxel.cuh:200-206)
	ST_SLOT, ST_PIXEL, ST_BOUNCES,
	ST_WORDS
};

#define STASH(k) s_stash[(k) * stash_stride + threadIdx.x]
#define STASHI(k) reinterpret_cast<int*>(s_stash)[(k) * stash_stride + threadIdx.x]
#define STASHU(k) reinterpret_cast<uint32_t*>(s_stash)[(k) * stash_stride + threadIdx.x]

// registers a lane keeps while stepping
struct LaneDda {
	int px, py, pz;
	float tx, ty, tz;
	float ex, ey, ez;    // tdelta (identical at every level: step * (1/d), voxel.cuh:48/101/186)
	int sx, sy, sz;      // step as integers (-1, 0, +1)
	int limxy, limz;     // positions stay in [0, lim) until the exit step (voxel.cuh:171-173 / 86-88 / 33-35)
	int level;           // 0 cell grid, 1 brick, 2 LoD octants
	int axis;            // last stepped axis at the current level, -1 = none yet
	int bm;              // the lane stands in an EMPTY aligned box of (bm + 1)^3 cells: 0 (just its cell), 1 or 2^SHIFT - 1
	const uint32_t* brick;  // level 1: brick words; level 2: the LoD byte in the low bits
};

__device__ __forceinline__ float axis_sel(int axis, float x, float y, float z) { return axis == 0 ? x : (axis == 1 ? y : z); }

// What does the cell-level DDA do next in the cell the lane stands in? Shared-memory block bitmap first: an empty block
// can be crossed in one exact multi-step (bm = block size - 1). Otherwise the per-cell bitmap of the 4x4x4 block (64 bits,
// global, L1/L2): non-empty cell -> M_CELL; empty cell whose 2x2x2 sub-block is empty too -> bm = 1; else single step.
template <int SHIFT>
__device__ __forceinline__ int classify_cell(LaneDda& t, const SceneView& sv, const uint32_t* s_coarse) {
	const int cb = (t.px >> SHIFT) + (t.py >> SHIFT) * sv.coarse_nx + (t.pz >> SHIFT) * sv.coarse_nxy;
	t.bm = (1 << SHIFT) - 1;
	if (!((s_coarse[cb >> 5] >> (cb & 31)) & 1u)) return M_TRACE;
	const int fb = SHIFT == 2 ? cb : (t.px >> 2) + (t.py >> 2) * sv.fine_nx + (t.pz >> 2) * sv.fine_nxy;
	const uint2 w = __ldg(reinterpret_cast<const uint2*>(sv.fine) + fb);
	const uint32_t half = (t.pz & 2) ? w.y : w.x;                     // z = 0,1 in the low word, z = 2,3 in the high word
	const int bit = (t.px & 3) | ((t.py & 3) << 2) | ((t.pz & 1) << 4);
	if ((half >> bit) & 1u) return M_CELL;
	// the 2x2x2 sub-block: x in {x&2, x&2+1}, y likewise, z in {z&2, z&2+1} -> bits 0,1,4,5,16,17,20,21 shifted
	const uint32_t sub = 0x00330033u << ((t.px & 2) | ((t.py & 2) << 2));
	t.bm = (half & sub) ? 0 : 1;
	return M_TRACE;
}

// Exact multi-step advance through an EMPTY aligned block of B^3 cells: performs, in one go, every DDA step the reference
// takes inside the block plus the step that leaves it (voxel.cuh:249-258 repeated), and lands in exactly the reference's
// state (pos, tmax, last axis). tmax along an axis only ever changes by `tmax += tdelta` on that axis, so the j-th crossing
// time of axis a is the j-fold sequential float sum t_a(j) = fl(t_a(j-1) + tdelta_a), independent of the other axes; the
// interleaving only decides how many crossings each axis gets to make before the first axis reaches the block boundary.
// Steps are taken in increasing (time, axis) order where on equal times z goes before y goes before x (the reference's
// tie rules: x only if strictly smallest, y if ty <= tx && ty < tz, else z).
// Returns false when the exit step leaves the world (voxel.cuh:256-257).
// The box is per lane: (t.bm + 1)^3 cells, t.bm + 1 <= B. With t.bm == 0 this is exactly one ordinary DDA step, so lanes
// that must test every cell (non-empty blocks, bricks, LoD octants) run the same instructions as lanes crossing empty space.
template <int B>
__device__ __forceinline__ bool jump_advance(LaneDda& t) {
	const int ax = t.px & t.bm, ay = t.py & t.bm, az = t.pz & t.bm;
	const int nx = t.sx > 0 ? t.bm + 1 - ax : ax + 1;  // steps along x until the box boundary is crossed
	const int ny = t.sy > 0 ? t.bm + 1 - ay : ay + 1;
	const int nz = t.sz > 0 ? t.bm + 1 - az : az + 1;
	float bx = t.tx, by = t.ty, bz = t.tz;  // time of the boundary crossing per axis = iterate n-1
#pragma unroll
	for (int j = 1; j < B; j++) {
		bx = j < nx ? bx + t.ex : bx;
		by = j < ny ? by + t.ey : by;
		bz = j < nz ? bz + t.ez : bz;
	}
	const bool xy = bx < by, xz = bx < bz, yz = by < bz;
	const bool exit_x = xy && xz, exit_y = !xy && yz;
	const float E = exit_x ? bx : (exit_y ? by : bz);
	// crossings ordered before the exit crossing: strictly earlier, or equal with a higher-priority axis
	const bool y_ties = exit_x, z_ties = exit_x || exit_y;
	int kx = 0, ky = 0, kz = 0;
	float vx = t.tx, vy = t.ty, vz = t.tz;
#pragma unroll
	for (int j = 1; j < B; j++) {
		const bool fx = vx < E;
		const bool fy = (vy < E) || (y_ties && vy == E);
		const bool fz = (vz < E) || (z_ties && vz == E);
		vx = fx ? vx + t.ex : vx; kx += fx ? 1 : 0;
		vy = fy ? vy + t.ey : vy; ky += fy ? 1 : 0;
		vz = fz ? vz + t.ez : vz; kz += fz ? 1 : 0;
	}
	// the exit step itself
	kx += exit_x ? 1 : 0;
	ky += exit_y ? 1 : 0;
	kz += (exit_x || exit_y) ? 0 : 1;
	t.px += kx * t.sx;
	t.py += ky * t.sy;
	t.pz += kz * t.sz;
	t.axis = exit_x ? 0 : (exit_y ? 1 : 2);
	if (!((unsigned)t.px < (unsigned)t.limxy && (unsigned)t.py < (unsigned)t.limxy && (unsigned)t.pz < (unsigned)t.limz)) return false;
	t.tx = exit_x ? vx + t.ex : vx;
	t.ty = exit_y ? vy + t.ey : vy;
	t.tz = (exit_x || exit_y) ? vz : vz + t.ez;
	return true;
}

__device__ __forceinline__ bool sub_occupied(const LaneDda& t) {  // voxel.cuh:110-113 / 57
	if (t.level == 1) {
		const int lin = t.px + t.py * 8 + t.pz * 64;
		return lin >= 0 && lin < 512 && ((__ldg(t.brick + (lin >> 5)) >> (lin & 31)) & 1u);
	}
	const int bit = t.px + t.py * 2 + t.pz * 4;
	return bit >= 0 && bit < 8 && (((uint32_t)(size_t)t.brick >> bit) & 1u);
}

// a hit inside a brick / LoD block: voxel.cuh:114-120 / 58-64, then :218 / :225
__device__ __forceinline__ float sub_hit_distance(const LaneDda& t, float* s_stash, const int stash_stride) {
	float sub = 0.f;
	if (t.axis > -1) {
		STASH(ST_NX) = t.axis == 0 ? -(float)t.sx : 0.f;
		STASH(ST_NY) = t.axis == 1 ? -(float)t.sy : 0.f;
		STASH(ST_NZ) = t.axis == 2 ? -(float)t.sz : 0.f;
		sub = axis_sel(t.axis, t.tx, t.ty, t.tz) - axis_sel(t.axis, t.ex, t.ey, t.ez);
	}
	const float nd8 = STASH(ST_ND) * 8.f;
	return t.level == 1 ? (nd8 + sub) + STASH(ST_TMIN) : (nd8 + sub * 4.f) + STASH(ST_TMIN);
}

// intersect_voxel up to the DDA loop (voxel.cuh:136-190) plus the test of the start cell. Returns the lane's next mode:
// M_DONE (missed outright, hit = false), M_CELL (start cell is non-empty) or M_TRACE.
template <int SHIFT>
__device__ __forceinline__ int trace_begin(LaneDda& t, const SceneView& sv, const uint32_t* s_coarse, float* s_stash, const int stash_stride, F3 origin, const F3 direction) {
	float tminn;
	STASH(ST_DX) = direction.x; STASH(ST_DY) = direction.y; STASH(ST_DZ) = direction.z;
	if (!intersect_aabb(sv, origin, direction, tminn)) return M_DONE;
	if (tminn > 0) {  // voxel.cuh:142-155
		origin = F3{ fmaf(direction.x, tminn, origin.x), fmaf(direction.y, tminn, origin.y), fmaf(direction.z, tminn, origin.z) };
		const float ratio = sv.grid_size_f / sv.grid_height_f;
		const float sxy = 1.f / ratio;
		const F3 center{ sv.grid_size_f / 2.f, sv.grid_size_f / 2.f, sv.grid_height_f / 2.f };
		F3 to_center{ fabsf(center.x - origin.x) * sxy, fabsf(center.y - origin.y) * sxy, fabsf(center.z - origin.z) * 1.f };
		const F3 signs{ gsign(origin.x - center.x), gsign(origin.y - center.y), gsign(origin.z - center.z) };
		const float m = gmax(to_center.x, gmax(to_center.y, to_center.z));
		to_center = F3{ to_center.x / m, to_center.y / m, to_center.z / m };
		const F3 n{ signs.x * truncf(to_center.x + 0.000001f), signs.y * truncf(to_center.y + 0.000001f), signs.z * truncf(to_center.z + 0.000001f) };
		STASH(ST_NX) = n.x; STASH(ST_NY) = n.y; STASH(ST_NZ) = n.z;
		origin = F3{ origin.x - n.x * kEpsilon, origin.y - n.y * kEpsilon, origin.z - n.z * kEpsilon };
	}
	origin = F3{ origin.x * 0.125f, origin.y * 0.125f, origin.z * 0.125f };
	Dda a;
	dda_setup(origin, direction, a);
	if (a.pos.x < 0 || a.pos.x >= sv.cells || a.pos.y < 0 || a.pos.y >= sv.cells || a.pos.z < 0 || a.pos.z >= sv.cells_height) return M_DONE;
	STASH(ST_CX) = origin.x; STASH(ST_CY) = origin.y; STASH(ST_CZ) = origin.z;
	STASH(ST_TMIN) = tminn;
	t.px = a.pos.x; t.py = a.pos.y; t.pz = a.pos.z;
	t.tx = a.tmax.x; t.ty = a.tmax.y; t.tz = a.tmax.z;
	t.ex = a.tdelta.x; t.ey = a.tdelta.y; t.ez = a.tdelta.z;
	t.sx = a.stepi.x; t.sy = a.stepi.y; t.sz = a.stepi.z;
	t.limxy = sv.cells;
	t.limz = sv.cells_height;
	t.level = 0;
	t.axis = -1;
	t.bm = 0;
	t.brick = nullptr;
	return classify_cell<SHIFT>(t, sv, s_coarse);
}

// Descend from a non-empty cell into its brick (level 1, voxel.cuh:222-227 + 79-107) or its LoD octants (level 2,
// voxel.cuh:215-220 + 26-54): park the cell-level DDA in the stash, set up the finer DDA along the same ray.
__device__ __forceinline__ void trace_descend(LaneDda& t, float* s_stash, const int stash_stride, int level, const uint32_t* payload, float nd) {
	const float nx = STASH(ST_NX), ny = STASH(ST_NY), nz = STASH(ST_NZ);
	const float dx = STASH(ST_DX), dy = STASH(ST_DY), dz = STASH(ST_DZ);
	const float x = fmaf(dx, nd, STASH(ST_CX)), y = fmaf(dy, nd, STASH(ST_CY)), z = fmaf(dz, nd, STASH(ST_CZ));
	float sox, soy, soz;
	if (level == 1) {
		sox = x * 8.f - nx * kEpsilon; soy = y * 8.f - ny * kEpsilon; soz = z * 8.f - nz * kEpsilon;
	} else {
		sox = fmaf(nx * 0.2f, -kEpsilon, x + x); soy = fmaf(ny * 0.2f, -kEpsilon, y + y); soz = fmaf(nz * 0.2f, -kEpsilon, z + z);
	}
	STASHI(ST_PX) = t.px; STASHI(ST_PY) = t.py; STASHI(ST_PZ) = t.pz;
	STASH(ST_MX) = t.tx; STASH(ST_MY) = t.ty; STASH(ST_MZ) = t.tz;
	// dda_setup with rdinv recovered exactly from tdelta = step * rdinv (step is -1, 0 or +1)
	const int ipx = (int)sox, ipy = (int)soy, ipz = (int)soz;
	const float cbx = dx > 0.f ? (float)(ipx + 1) : (float)ipx, cby = dy > 0.f ? (float)(ipy + 1) : (float)ipy, cbz = dz > 0.f ? (float)(ipz + 1) : (float)ipz;
	const float rx = (float)t.sx * t.ex, ry = (float)t.sy * t.ey, rz = (float)t.sz * t.ez;
	t.tx = dx != 0.f ? (cbx - sox) * rx : 1000000.f;
	t.ty = dy != 0.f ? (cby - soy) * ry : 1000000.f;
	t.tz = dz != 0.f ? (cbz - soz) * rz : 1000000.f;
	const int side = level == 1 ? 8 : 2;
	t.px = ipx % side; t.py = ipy % side; t.pz = ipz % side;
	t.limxy = side;
	t.limz = side;
	t.level = level;
	t.axis = -1;
	t.bm = 0;
	t.brick = payload;
}

// SHIFT = sv.coarse_shift (block size of the shared-memory bitmap = size of an exact jump)
template <int SHIFT>
__global__ void __launch_bounds__(kV2Block, BM_V2_MIN_BLOCKS) frame_kernel_v2(const FrameParams fp, const SceneView sv, const FrameIO io) {
	extern __shared__ uint32_t s_dyn[];
	uint32_t* s_coarse = s_dyn;                                              // sv.coarse_words
	float* s_stash = reinterpret_cast<float*>(s_dyn + sv.coarse_words);      // ST_WORDS * blockDim.x
	const int stash_stride = (int)blockDim.x;
	const int kShadeThreshold = io.tune_shade, kCellThreshold = io.tune_cell, kStepBurst = io.tune_burst;
	__shared__ unsigned long long s_stats[4];

	DeviceState* st = io.st;
	if (st->done) return;
	for (uint32_t i = threadIdx.x; i < sv.coarse_words; i += blockDim.x) s_coarse[i] = __ldg(sv.coarse + i);
	if (threadIdx.x < 4) s_stats[threadIdx.x] = 0;
	__syncthreads();

	const uint32_t c = st->primary_ray_cnt;
	const uint32_t start = st->start_position;
	const uint32_t frame = st->frame;
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t lt_mask = (1u << lane) - 1u;
	uint32_t n_shadow = 0, n_term = 0, n_unocc = 0;

	uint32_t pool_next = 0, pool_end = 0;  // warp-uniform run of slots [pool_next, pool_end)
	bool pool_dry = false;

	int mode = M_IDLE;
	bool shadow = false;  // which trace the lane is on: extend (kernel.cu:226-238) or connect (kernel.cu:328-346)
	bool hit = false;
	float distance = 0.f;
	LaneDda t;
	t.px = t.py = t.pz = 0;
	t.tx = t.ty = t.tz = 0.f;
	t.ex = t.ey = t.ez = 0.f;
	t.sx = t.sy = t.sz = 0;
	t.limxy = t.limz = 0;
	t.level = 0;
	t.axis = -1;
	t.bm = 0;
	t.brick = nullptr;

	for (;;) {
		uint32_t b_step = __ballot_sync(0xFFFFFFFFu, mode == M_TRACE);
		uint32_t b_cell = __ballot_sync(0xFFFFFFFFu, mode == M_CELL);
		const uint32_t b_done = __ballot_sync(0xFFFFFFFFu, mode == M_DONE);
		const uint32_t b_idle = ~(b_step | b_cell | b_done);
		const uint32_t refillable = pool_dry ? 0u : b_idle;

		// ---------------- SHADE + REFILL ----------------------------------------------------------------------------------
		if (__popc(b_done | refillable) >= kShadeThreshold || ((b_step | b_cell) == 0)) {
			if ((b_done | refillable) == 0) break;  // nothing in flight and nothing left to take
			if (mode == M_DONE) {
				const uint32_t pixel = STASHU(ST_PIXEL);
				if (!shadow) {
					const uint32_t slot = STASHU(ST_SLOT);
					Ray ray;
					ray.origin = F3{ STASH(ST_OX), STASH(ST_OY), STASH(ST_OZ) };
					ray.direction = F3{ STASH(ST_DX), STASH(ST_DY), STASH(ST_DZ) };
					ray.throughput = F3{ STASH(ST_TX), STASH(ST_TY), STASH(ST_TZ) };
					ray.normal = F3{ STASH(ST_NX), STASH(ST_NY), STASH(ST_NZ) };
					ray.distance = hit ? distance : kVeryFar;
					ray.identifier = 0;
					ray.bounces = STASHI(ST_BOUNCES);
					ray.pixel_index = pixel;
					const ShadeResult s = shade_vertex(fp, frame, slot, ray);  // kernel.cu:242-325
					if (s.add_radiance) accum_add(io.accum, pixel, s.radiance.x, s.radiance.y, s.radiance.z, 1.f);
					else if (s.terminated) accum_add(io.accum, pixel, 0.f, 0.f, 0.f, 1.f);
					n_term += s.terminated;
					if (s.survives) {
						store_ray(io.out + slot, ray);
						atomicOr(io.out_mask + (slot >> 5), 1u << (slot & 31));
					}
					mode = M_IDLE;
					if (s.has_shadow) {  // connect, kernel.cu:328-346: origin = shaded hit point, normal y{} = 0
						n_shadow++;
						STASH(ST_TX) = s.shadow_color.x; STASH(ST_TY) = s.shadow_color.y; STASH(ST_TZ) = s.shadow_color.z;
						STASH(ST_NX) = 0.f; STASH(ST_NY) = 0.f; STASH(ST_NZ) = 0.f;
						shadow = true;
						hit = false;
						mode = trace_begin<SHIFT>(t, sv, s_coarse, s_stash, stash_stride, ray.origin, s.shadow_dir);
					}
				} else {
					if (!hit) {
						accum_add(io.accum, pixel, STASH(ST_TX), STASH(ST_TY), STASH(ST_TZ), 0.f);
						n_unocc++;
					}
					mode = M_IDLE;
				}
			}
			// REFILL: idle lanes take the next slots in order
			const uint32_t idle = __ballot_sync(0xFFFFFFFFu, mode == M_IDLE);
			const uint32_t need = __popc(idle);
			if (need && !pool_dry) {
				const uint32_t my_rank = __popc(idle & lt_mask);
				uint32_t given = 0, my_slot = 0xFFFFFFFFu;
				while (given < need) {
					if (pool_next == pool_end) {
						uint32_t tile = 0;
						if (lane == 0) tile = atomicAdd(&st->tile_ticket, 1u);
						tile = __shfl_sync(0xFFFFFFFFu, tile, 0);
						if (tile >= io.ntiles) {
							pool_dry = true;
							break;
						}
						pool_next = tile * kTile;
						pool_end = min(pool_next + kTile, fp.n_slots);
					}
					const uint32_t take = min(need - given, pool_end - pool_next);
					if (mode == M_IDLE && my_rank >= given && my_rank < given + take) my_slot = pool_next + (my_rank - given);
					pool_next += take;
					given += take;
				}
				if (mode == M_IDLE && my_slot != 0xFFFFFFFFu) {
					Ray ray;
					if (my_slot < c) ray = load_ray(survivor_ptr(io, my_slot));
					else ray = generate_primary(fp, frame, start, my_slot - c);  // primary_rays, kernel.cu:154-223
					STASHU(ST_SLOT) = my_slot;
					STASHU(ST_PIXEL) = ray.pixel_index;
					STASHI(ST_BOUNCES) = ray.bounces;
					STASH(ST_OX) = ray.origin.x; STASH(ST_OY) = ray.origin.y; STASH(ST_OZ) = ray.origin.z;
					STASH(ST_TX) = ray.throughput.x; STASH(ST_TY) = ray.throughput.y; STASH(ST_TZ) = ray.throughput.z;
					STASH(ST_NX) = ray.normal.x; STASH(ST_NY) = ray.normal.y; STASH(ST_NZ) = ray.normal.z;
					shadow = false;
					hit = false;
					mode = trace_begin<SHIFT>(t, sv, s_coarse, s_stash, stash_stride, ray.origin, ray.direction);  // extend, kernel.cu:226-238
				}
			}
			b_step = __ballot_sync(0xFFFFFFFFu, mode == M_TRACE);
			b_cell = __ballot_sync(0xFFFFFFFFu, mode == M_CELL);
		}

		// ---------------- CELL: a non-empty cell (voxel.cuh:197-246) ---------------------------------------------------------
		if (__popc(b_cell) >= kCellThreshold || (b_cell && !b_step)) {
			if (mode == M_CELL) {
				const int sc = (t.px >> 4) + (t.py >> 4) * sv.supergrid_xy + (t.pz >> 4) * sv.supergrid_xy * sv.supergrid_xy;  // voxel.cuh:197
				const int local = (t.px & 15) + (t.py & 15) * 16 + (t.pz & 15) * 256;                                        // voxel.cuh:198
				uint32_t* word = sv.flat_indices ? sv.flat_indices + (((size_t)sc << 12) + local) : sv.indices[sc] + local;
				const uint32_t index = __ldg(word);
				mode = M_TRACE;  // whatever happens next at the cell level is a single step: the cell's block is not empty
				t.bm = 0;
				if (index) {
					float nd = 0.f;
					if (t.axis != -1) {  // voxel.cuh:201-206
						STASH(ST_NX) = t.axis == 0 ? -(float)t.sx : 0.f;
						STASH(ST_NY) = t.axis == 1 ? -(float)t.sy : 0.f;
						STASH(ST_NZ) = t.axis == 2 ? -(float)t.sz : 0.f;
						nd = axis_sel(t.axis, t.tx, t.ty, t.tz) - axis_sel(t.axis, t.ex, t.ey, t.ez);
					}
					STASH(ST_ND) = nd;
					const int ddx = fp.cam_cell.x - t.px, ddy = fp.cam_cell.y - t.py, ddz = fp.cam_cell.z - t.pz;
					const int lod_distance_squared = ddx * ddx + ddy * ddy + ddz * ddz;
					if (lod_distance_squared > sv.lod8) {  // voxel.cuh:212-214
						hit = true;
						distance = nd * 8.f + STASH(ST_TMIN);
						mode = M_DONE;
					} else if (lod_distance_squared > sv.lod2) {  // voxel.cuh:215-220
						trace_descend(t, s_stash, stash_stride, 2, reinterpret_cast<const uint32_t*>((size_t)((index & BM_BRICK_LOD_BITS) >> 12)), nd);
					} else if (index & BM_BRICK_LOADED_BIT) {  // voxel.cuh:222-227
						trace_descend(t, s_stash, stash_stride, 1, sv.bricks[sc][index & BM_BRICK_INDEX_BITS].data, nd);
					} else if (index & BM_BRICK_UNLOADED_BIT) {  // voxel.cuh:228-244
						const uint32_t old = atomicOr(word, BM_BRICK_REQUESTED_BIT);
						if (!(old & BM_BRICK_REQUESTED_BIT)) {
							const uint32_t load_index = atomicAdd(sv.load_queue_count, 1u);
							if (load_index < sv.queue_size) {
								sv.load_queue[3 * load_index + 0] = t.px;
								sv.load_queue[3 * load_index + 1] = t.py;
								sv.load_queue[3 * load_index + 2] = t.pz;
							} else {
								atomicAnd(word, ~BM_BRICK_REQUESTED_BIT);
							}
						}
						hit = true;
						distance = nd * 8.f + STASH(ST_TMIN);
						mode = M_DONE;
					}
					// first voxel / octant of the finer level (the DDA loops test before they step, voxel.cuh:109-120)
					if (mode == M_TRACE && t.level != 0 && sub_occupied(t)) {
						hit = true;
						distance = sub_hit_distance(t, s_stash, stash_stride);
						mode = M_DONE;
					}
				}
			}
			b_step = __ballot_sync(0xFFFFFFFFu, mode == M_TRACE);
		}

		// ---------------- TRACE: advance + occupancy test, a few iterations -------------------------------------------------
		if (b_step) {
#pragma unroll 1
			for (int k = 0; k < kStepBurst; k++) {
				if (mode == M_TRACE) {
					// voxel.cuh:66-74 / 122-130 / 249-258, one step or one exact multi-step through the lane's empty box
					if (jump_advance<(1 << SHIFT)>(t)) {
						if (t.level == 0) {
							mode = classify_cell<SHIFT>(t, sv, s_coarse);
						} else if (sub_occupied(t)) {
							hit = true;
							distance = sub_hit_distance(t, s_stash, stash_stride);
							mode = M_DONE;
						}
					} else if (t.level == 0) {  // left the world: miss
						hit = false;
						mode = M_DONE;
					} else {
						// left the brick / LoD block without a hit: resume the parked cell-level DDA (it stands in a non-empty
						// block: single step); its advance comes next
						t.px = STASHI(ST_PX); t.py = STASHI(ST_PY); t.pz = STASHI(ST_PZ);
						t.tx = STASH(ST_MX); t.ty = STASH(ST_MY); t.tz = STASH(ST_MZ);
						t.limxy = sv.cells;
						t.limz = sv.cells_height;
						t.level = 0;
						t.bm = 0;
					}
				}
			}
		}
	}

	// per-block statistics -> a handful of atomics per block
	unsigned long long a = warp_sum((unsigned long long)n_shadow), b = warp_sum((unsigned long long)n_term), u = warp_sum((unsigned long long)n_unocc);
	if (lane == 0) {
		atomicAdd(&s_stats[0], a);
		atomicAdd(&s_stats[1], b);
		atomicAdd(&s_stats[2], u);
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		if (s_stats[0]) atomicAdd(&st->shadow_rays, s_stats[0]);
		if (s_stats[1]) atomicAdd(&st->terminations, s_stats[1]);
		if (s_stats[1]) atomicAdd(&st->paths_since_reset, s_stats[1]);
		if (s_stats[2]) atomicAdd(&st->unoccluded, s_stats[2]);
	}
}

#undef STASH
#undef STASHI
#undef STASHU

}  // namespace bm
