// Kernels and C ABI of the B200 path tracer (include/brickmap_b200.h).
//
// Execution model (B200-first, not the reference's five atomically-fed wavefront kernels, kernel.cu:412-420):
//   * ONE kernel per frame. A slot of the frame goes from ray generation (or survivor fetch) through extend, shade and its
//     shadow ray without leaving the SM; the 64-byte RayQueue record is touched once on the way in (survivors only) and once
//     on the way out (survivors only), as four 128-bit transactions. The reference moves ~310 B per slot through L2/HBM.
//   * Two frame kernels with identical results. frame_kernel_q (bm_frame_quantum.cuh) is the throughput path of bm_render:
//     a warp traces 32 rays at a time and suspends / regroups them through a work queue in shared memory, so that long rays
//     do not hold 31 idle lanes; its RECORD instantiation serves bm_launch_frame (every buffer the reference's kernels
//     leave). frame_kernel (below) keeps one thread on one slot from start to end; it serves the work counters (COUNT) and
//     worlds whose bitmap or queue entries do not fit the throughput kernel.
//   * Warps pull runs of 128 consecutive slots with one atomic per run (the reference: one same-address atomic per ray per
//     kernel, kernel.cu:158,228,245,330). Survivors are written sparse at their slot and flagged in a bitmask; a 1-block scan
//     kernel turns the mask popcounts into the next frame's slot numbering (search + select in survivor_ptr), advances the
//     pixel cursor and the frame counter on the device (set_wavefront_globals, kernel.cu:122-139). No block barrier inside a
//     frame, no host round trip between frames.
//   * Two derived emptiness bitmaps: 1 bit per 4x4x4 cells with a one-block border in shared memory (46 KiB at reference dims,
//     one LDS per DDA step, no bounds test in empty space) and 1 bit per cell in global memory (2 MiB, touched only inside
//     non-empty blocks). The DDA performs the reference's exact float step sequence (hand-scheduled in PTX) but loads an
//     index word only for cells that are really non-empty (2.2 per ray instead of the reference's 88). Index words of a flat
//     arena are addressed directly (no pointer-table dependent load, voxel.cuh:197-198).
//   * Slot order == "the reference scheduled one thread at a time", so seeds (kernel.cu:165,252) and results are
//     reproducible and comparable with the reference launched <<<1,1>>>.
// Variants that were built, verified and measured slower are listed in DESIGN.md section 4 (sources: profiles/experiments/).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "bm_device.cuh"

namespace bm {

constexpr int kTile = 256;  // slots per tile == threads per block

struct DeviceState {
	uint32_t primary_ray_cnt;  // kernel.cu:106
	uint32_t start_position;   // kernel.cu:109
	uint32_t shadow_ray_cnt;   // kernel.cu:119 (of the last frame)
	uint32_t frame;            // kernel.cu:369
	uint32_t tile_ticket;      // replaces raynr_primary/extend/shade/connect (kernel.cu:111-117)
	uint32_t done;             // bm_render target reached
	uint32_t cur;              // which private survivor set (FrameIO::rays/mask/prefix) holds the survivors of the last frame that RAN;
	                           // toggled by scan_kernel, so frames the device skipped (done) do not move it
	uint32_t n_active;         // slots of the NEXT frame: ray_queue_buffer_size, fewer only in the last frames of BM_FRAME_EXACT_PATHS
	unsigned long long frames, extend_rays, shadow_rays, terminations, unoccluded, cell_steps, index_reads, bricks, requests;
	unsigned long long paths_since_reset, target_paths;
	uint32_t exact;            // BM_FRAME_EXACT_PATHS
	uint32_t pad0;
};

// Slots of the next frame. Normally all N (the reference: every frame fills the queue, kernel.cu:160-163). BM_FRAME_EXACT_PATHS: only
// as many fresh primaries as are still needed for `target` paths since the last reset -- every path started so far has either
// finished (paths_since_reset) or is one of the c survivors, so target - paths_since_reset - c are still to be started.
__device__ __forceinline__ uint32_t next_frame_slots(const DeviceState* st, uint32_t n_slots, uint32_t c) {
	if (!st->exact || !st->target_paths) return n_slots;
	const unsigned long long started = st->paths_since_reset + c;
	const unsigned long long want = st->target_paths > started ? st->target_paths - started : 0ull;
	const unsigned long long room = n_slots - c;
	return c + (uint32_t)(want < room ? want : room);
}

struct FrameIO {
	DeviceState* st;
	// Two private survivor sets; set st->cur holds the survivors of the previous frame, the other one receives this frame's.
	// A set is SPARSE: the survivor that came out of slot s sits at rays[s], bit s of mask is set, and prefix[t] counts the
	// survivors of slots < 256 t, so that survivor number k (= the next frame's slot k) is found by search + select.
	bm_ray* rays[2];
	uint32_t* mask[2];       // 1 bit per slot (the output set's is zero on entry)
	uint32_t* prefix[2];
	const bm_ray* dense_in;  // != nullptr: the previous frame's survivors are the dense records [0,c) of a caller's queue instead
	bm_ray* record;          // RECORD: the reference's work queue, post-extend record of every slot
	bm_shadow* shadow_out;   // RECORD: shadow rays, sparse by slot
	uint32_t* shadow_mask;   // RECORD
	float4* accum;           // blit_buffer (state.h:22)
	uint32_t ntiles;
	uint32_t extend_only;    // BM_FRAME_EXTEND_ONLY: stop after extend (kernel.cu:416-418), nothing is shaded
};

__device__ __forceinline__ void accum_add(float4* accum, uint32_t pixel, float r, float g, float b, float a) {
	// one 128-bit reduction instead of the reference's 3-4 scalar float atomics (kernel.cu:319-322,341-343)
#if BM_SURV_HINTS & 128
	// MEASUREMENT BUILD ONLY (wrong images): no accumulation at all -- the rays are the same ones (nothing reads the image back), so the
	// difference in DRAM traffic to the normal build is the accumulation buffer's share (profiles/r2_w_*)
	(void)accum; (void)pixel; (void)r; (void)g; (void)b; (void)a;
#elif BM_SURV_HINTS & 16
	// ... carrying an L2 evict_last policy: the image (33 MB at 1080p) is touched all over by every frame and should outlive the survivor streams
	asm volatile(
	    "{\n\t"
	    ".reg .b64 pol;\n\t"
	    "createpolicy.fractional.L2::evict_last.b64 pol, 1.0;\n\t"
	    "red.global.add.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, pol;\n\t"
	    "}" ::"l"(accum + pixel), "f"(r), "f"(g), "f"(b), "f"(a)
	    : "memory");
#else
	atomicAdd(accum + pixel, make_float4(r, g, b, a));
#endif
}

// Where is survivor number `k` of the previous frame? Stable compaction = the reference's atomicAdd(&primary_ray_cnt, 1)
// (kernel.cu:298-299) with the schedule fixed to slot order; the records themselves are never moved.
// Prefix layout of a survivor set (two levels, so that the scan needs no pass over all tiles to add block bases): prefix[t], t < ntiles =
// survivors of the tiles before t INSIDE t's scan block (kScanTiles tiles); prefix[ntiles] = total; prefix[ntiles + 1 + b] = survivors
// before scan block b (b <= nblocks; the last entry is the total again).
constexpr int kScanTiles = 256;
__host__ __device__ __forceinline__ uint32_t scan_block_count(uint32_t ntiles) { return (ntiles + kScanTiles - 1) / kScanTiles; }
__host__ __device__ __forceinline__ size_t prefix_words(uint32_t ntiles) { return (size_t)ntiles + 2 + scan_block_count(ntiles); }
struct SurvivorSet {
	const bm_ray* in;
	const uint32_t* in_mask;
	const uint32_t* in_prefix;  // nullptr: dense
	uint32_t ntiles;
};
// (selects between kernel parameters at the point of use instead of indexing them: the pointers stay in the constant bank and
// only `cur` lives in a register across the frame)
__device__ __forceinline__ SurvivorSet input_set(const FrameIO& io, uint32_t cur) {
	if (io.dense_in) return SurvivorSet{ io.dense_in, nullptr, nullptr, io.ntiles };
	return cur ? SurvivorSet{ io.rays[1], io.mask[1], io.prefix[1], io.ntiles } : SurvivorSet{ io.rays[0], io.mask[0], io.prefix[0], io.ntiles };
}
__device__ __forceinline__ bm_ray* output_rays(const FrameIO& io, uint32_t cur) { return cur ? io.rays[0] : io.rays[1]; }
__device__ __forceinline__ uint32_t* output_mask(const FrameIO& io, uint32_t cur) { return cur ? io.mask[0] : io.mask[1]; }
__device__ __forceinline__ const bm_ray* survivor_ptr(const SurvivorSet& io, uint32_t k) {
	if (!io.in_prefix) return io.in + k;
	const uint32_t* base = io.in_prefix + io.ntiles + 1;
	uint32_t blo = 0, bhi = scan_block_count(io.ntiles);  // scan block b = max{ b : base[b] <= k }
	while (bhi - blo > 1) {
		const uint32_t mid = (blo + bhi) >> 1;
		if (__ldg(base + mid) <= k) blo = mid; else bhi = mid;
	}
	uint32_t r = k - __ldg(base + blo);
	uint32_t lo = blo * kScanTiles, hi = min(io.ntiles, lo + kScanTiles);  // tile t = max{ t in the block : prefix[t] <= r }
	while (hi - lo > 1) {
		const uint32_t mid = (lo + hi) >> 1;
		if (__ldg(io.in_prefix + mid) <= r) lo = mid; else hi = mid;
	}
	r -= __ldg(io.in_prefix + lo);
	const uint32_t* m = io.in_mask + (size_t)lo * (kTile / 32);
	uint32_t pos = 0;
#pragma unroll
	for (int w = 0; w < kTile / 32; w++) {
		const uint32_t word = __ldg(m + w);
		const uint32_t cnt = __popc(word);
		if (r < cnt) {
			pos = w * 32 + __fns(word, 0, r + 1);
			break;
		}
		r -= cnt;
	}
	return io.in + (size_t)lo * kTile + pos;
}

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, o);
	return v;
}

__device__ __forceinline__ void store_shadow(bm_shadow* q, const F3& o, const F3& d, const F3& c, uint32_t pixel) {  // kernel.cu:277-278
	q->origin[0] = o.x; q->origin[1] = o.y; q->origin[2] = o.z;
	q->direction[0] = d.x; q->direction[1] = d.y; q->direction[2] = d.z;
	q->color[0] = c.x; q->color[1] = c.y; q->color[2] = c.z;
	q->pixel_index = pixel;
}

constexpr int kShadowQueue = 64;  // per-warp shadow-ray queue entries (structure of arrays, 10 words per entry)

template <bool COUNT>
__device__ __forceinline__ void trace_shadow(const SceneView& sv, const uint32_t* coarse, const FrameParams& fp, const FrameIO& io, const float* q, uint32_t e,
                                             WorkCounters& wc, unsigned long long& n_unocc) {
	const F3 o{ q[0 * kShadowQueue + e], q[1 * kShadowQueue + e], q[2 * kShadowQueue + e] };
	const F3 d{ q[3 * kShadowQueue + e], q[4 * kShadowQueue + e], q[5 * kShadowQueue + e] };
	F3 y{ 0.f, 0.f, 0.f };  // kernel.cu:337-338
	float t = 0.f;
	if (!intersect_voxel<COUNT>(sv, coarse, o, d, y, t, fp.cam_cell, &wc)) {
		accum_add(io.accum, __float_as_uint(q[9 * kShadowQueue + e]), q[6 * kShadowQueue + e], q[7 * kShadowQueue + e], q[8 * kShadowQueue + e], 0.f);
		n_unocc++;
	}
}

// ---- the frame kernel: one thread owns one slot from ray generation / survivor fetch through extend, shade and its shadow
// ray. RECORD (bm_launch_frame) additionally leaves every buffer the reference's five kernels leave; COUNT fills the work
// counters. Warps pull runs of 128 consecutive slots; there is no block-level synchronisation inside a frame.
template <bool RECORD, bool COUNT>
__global__ void __launch_bounds__(kTile, 4) frame_kernel(const FrameParams fp, const SceneView sv, const FrameIO io) {
	extern __shared__ uint32_t s_coarse[];
	DeviceState* st = io.st;
	if (st->done) return;
	for (uint32_t i = threadIdx.x; i < sv.coarse_words; i += blockDim.x) s_coarse[i] = __ldg(sv.coarse + i);
	const uint32_t* coarse = s_coarse;
	__syncthreads();

	const uint32_t c = st->primary_ray_cnt;
	const uint32_t start = st->start_position;
	const uint32_t frame = st->frame;
	const uint32_t cur = st->cur;
	const uint32_t n_slots = st->n_active;
	unsigned long long n_shadow = 0, n_term = 0, n_unocc = 0;
	WorkCounters wc{ 0, 0, 0, 0 };

	// each warp pulls kRun * 32 consecutive slots with one atomic (the reference: one same-address atomic per ray per
	// kernel, kernel.cu:158,228,245,330); no block-level synchronisation inside the frame
	constexpr uint32_t kRun = 1;  // one ticket per 32 slots: what a warp still holds when the pool runs dry is the frame's tail
	const uint32_t lane = threadIdx.x & 31;
	float* q = reinterpret_cast<float*>(s_coarse + sv.coarse_words) + (threadIdx.x >> 5) * (10 * kShadowQueue);  // this warp's shadow-ray queue
	uint32_t qn = 0;                                                                                               // entries queued (warp-uniform)
	const uint32_t nruns = (n_slots + kRun * 32 - 1) / (kRun * 32);
	uint32_t run = 0, round = kRun;
	for (;;) {
		if (round == kRun) {
			if (lane == 0) run = atomicAdd(&st->tile_ticket, 1u);
			run = __shfl_sync(0xFFFFFFFFu, run, 0);
			if (run >= nruns) break;
			round = 0;
		}
		const uint32_t slot = (run * kRun + round) * 32 + lane;
		round++;
		if (slot - lane >= n_slots) continue;
		const bool valid = slot < n_slots;
		Ray ray;
		bool survives = false, has_shadow = false;
		F3 shadow_dir{ 0, 0, 0 }, shadow_color{ 0, 0, 0 };
		if (valid) {
			if (slot < c) ray = load_ray(survivor_ptr(input_set(io, cur), slot));
			else ray = generate_primary(fp, frame, start, slot - c);  // primary_rays, kernel.cu:154-223
			// extend, kernel.cu:226-238
			ray.distance = kVeryFar;
			intersect_voxel<COUNT>(sv, coarse, ray.origin, ray.direction, ray.normal, ray.distance, fp.cam_cell, &wc);
			if (RECORD) store_ray(io.record + slot, ray);
			if (!io.extend_only) {
				// shade, kernel.cu:242-325
				const ShadeResult s = shade_vertex(fp, frame, slot, ray);
				survives = s.survives;
				has_shadow = s.has_shadow;
				shadow_dir = s.shadow_dir;
				shadow_color = s.shadow_color;
				if (s.add_radiance) accum_add(io.accum, ray.pixel_index, s.radiance.x, s.radiance.y, s.radiance.z, 1.f);
				else if (s.terminated) accum_add(io.accum, ray.pixel_index, 0.f, 0.f, 0.f, 1.f);
				n_term += s.terminated;
			}
		}
		// connect, kernel.cu:328-346. Only about half of the lanes come out of shade with a shadow ray, so the rays are queued
		// per warp in shared memory and traced 32 at a time: the shadow traversal always runs with a full warp. (Which thread
		// traces a shadow ray does not matter: it only adds to the accumulation buffer, kernel.cu:341-343.)
		{
			const uint32_t hm = __ballot_sync(0xFFFFFFFFu, has_shadow);
			if (has_shadow) {
				const uint32_t e = qn + __popc(hm & ((1u << lane) - 1u));
				q[0 * kShadowQueue + e] = ray.origin.x; q[1 * kShadowQueue + e] = ray.origin.y; q[2 * kShadowQueue + e] = ray.origin.z;
				q[3 * kShadowQueue + e] = shadow_dir.x; q[4 * kShadowQueue + e] = shadow_dir.y; q[5 * kShadowQueue + e] = shadow_dir.z;
				q[6 * kShadowQueue + e] = shadow_color.x; q[7 * kShadowQueue + e] = shadow_color.y; q[8 * kShadowQueue + e] = shadow_color.z;
				q[9 * kShadowQueue + e] = __uint_as_float(ray.pixel_index);
			}
			qn += __popc(hm);
			__syncwarp();
			if (qn >= 32) {
				qn -= 32;
				trace_shadow<COUNT>(sv, coarse, fp, io, q, qn + lane, wc, n_unocc);
				n_shadow++;
				__syncwarp();
			}
		}
		if (survives) store_ray(output_rays(io, cur) + slot, ray);
		const uint32_t smask = __ballot_sync(0xFFFFFFFFu, survives);
		if (lane == 0) output_mask(io, cur)[slot >> 5] = smask;
		if (RECORD) {
			if (has_shadow) store_shadow(io.shadow_out + slot, ray.origin, shadow_dir, shadow_color, ray.pixel_index);
			const uint32_t hmask = __ballot_sync(0xFFFFFFFFu, has_shadow);
			if (lane == 0) io.shadow_mask[slot >> 5] = hmask;
		}
	}

	if (lane < qn) {  // the last, partial batch of shadow rays
		trace_shadow<COUNT>(sv, coarse, fp, io, q, lane, wc, n_unocc);
		n_shadow++;
	}

	// per-warp statistics -> a handful of atomics per warp
	n_shadow = warp_sum(n_shadow);
	n_term = warp_sum(n_term);
	n_unocc = warp_sum(n_unocc);
	if (lane == 0) {
		if (n_shadow) atomicAdd(&st->shadow_rays, n_shadow);
		if (n_term) atomicAdd(&st->terminations, n_term);
		if (n_term) atomicAdd(&st->paths_since_reset, n_term);
		if (n_unocc) atomicAdd(&st->unoccluded, n_unocc);
	}
	if (COUNT) {
		const unsigned long long a = warp_sum(wc.index_reads), b = warp_sum(wc.bricks), q = warp_sum(wc.requests), t = warp_sum(wc.steps);
		if (lane == 0) {
			atomicAdd(&st->index_reads, a);
			atomicAdd(&st->bricks, b);
			atomicAdd(&st->requests, q);
			atomicAdd(&st->cell_steps, t);
		}
	}
}

}  // namespace bm
#include "bm_frame_quantum.cuh"
namespace bm {

// set_wavefront_globals (kernel.cu:122-139) + per-tile survivor counts (popcount of the slot masks) -> exclusive prefix.
// mask[ntiles * 8] -> prefix[ntiles + 1]; optional second pair for the shadow queue (RECORD). Also zeroes the mask buffer the NEXT
// frame will write with atomicOr. Multi-block: block b scans kScanTiles tiles (one per thread) and leaves its total; the block that
// finishes last (ticket counter, no waiting anywhere) scans the block totals into the block bases and advances the frame state.
__device__ __forceinline__ uint32_t tile_count(const uint32_t* mask, uint32_t tile) {
	const uint4 lo = *reinterpret_cast<const uint4*>(mask + (size_t)tile * 8), hi = *reinterpret_cast<const uint4*>(mask + (size_t)tile * 8 + 4);
	return __popc(lo.x) + __popc(lo.y) + __popc(lo.z) + __popc(lo.w) + __popc(hi.x) + __popc(hi.y) + __popc(hi.z) + __popc(hi.w);
}
// exclusive scan over the block's kScanTiles threads; returns the block total through `total`
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* s_warp /*kScanTiles / 32 + 1*/, uint32_t& total) {
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t incl = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, o);
		if (lane >= (uint32_t)o) incl += u;
	}
	__syncthreads();  // s_warp may still be read from a previous call
	if (lane == 31) s_warp[warp] = incl;
	__syncthreads();
	uint32_t base = 0, all = 0;
#pragma unroll
	for (int w = 0; w < kScanTiles / 32; w++) {
		const uint32_t t = s_warp[w];
		if (w < (int)warp) base += t;
		all += t;
	}
	total = all;
	return base + incl - v;
}
__global__ void __launch_bounds__(kScanTiles) scan_kernel(const FrameIO io, const uint32_t* shadow_mask, uint32_t* shadow_prefix, uint32_t ntiles, uint32_t n_slots,
                                                           uint32_t pixels, uint32_t* block_totals /*2 * gridDim.x*/, uint32_t* ticket) {
	__shared__ uint32_t s_warp[kScanTiles / 32 + 1];
	__shared__ uint32_t s_last, s_carry;
	DeviceState* st = io.st;
	if (st->done) return;          // (only the last block changes the state, after every other block has taken its ticket)
	const uint32_t cur = st->cur;  // the set the frame kernel read; it wrote set cur ^ 1
	const uint32_t* mask = cur ? io.mask[0] : io.mask[1];
	uint32_t* prefix = cur ? io.prefix[0] : io.prefix[1];
	uint32_t* clear_mask = cur ? io.mask[1] : io.mask[0];  // this frame's input becomes the next frame's output
	const uint32_t tile = blockIdx.x * kScanTiles + threadIdx.x;
	for (int pass = 0; pass < 2; pass++) {
		const uint32_t* in = pass == 0 ? mask : shadow_mask;
		uint32_t* out = pass == 0 ? prefix : shadow_prefix;
		if (!in) continue;
		const uint32_t n = tile < ntiles ? tile_count(in, tile) : 0u;
		uint32_t total;
		const uint32_t excl = block_exclusive_scan(n, s_warp, total);
		if (tile < ntiles) out[tile] = excl;  // relative to the scan block (see the prefix layout above)
		if (threadIdx.x == 0) block_totals[pass * gridDim.x + blockIdx.x] = total;
	}
	if (tile < ntiles) {
		reinterpret_cast<uint4*>(clear_mask)[(size_t)tile * 2] = make_uint4(0, 0, 0, 0);
		reinterpret_cast<uint4*>(clear_mask)[(size_t)tile * 2 + 1] = make_uint4(0, 0, 0, 0);
	}
	__threadfence();
	__syncthreads();
	if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
	__syncthreads();
	if (!s_last) return;
	__threadfence();
	for (int pass = 0; pass < 2; pass++) {
		uint32_t* out = pass == 0 ? prefix : shadow_prefix;
		if (!(pass == 0 ? mask : shadow_mask)) continue;
		uint32_t* totals = block_totals + pass * gridDim.x;
		if (threadIdx.x == 0) s_carry = 0;
		__syncthreads();
		for (uint32_t b0 = 0; b0 < gridDim.x; b0 += kScanTiles) {  // block totals -> block bases, kScanTiles at a time
			const uint32_t b = b0 + threadIdx.x;
			const uint32_t v = b < gridDim.x ? totals[b] : 0u;
			uint32_t chunk_total;
			const uint32_t excl = block_exclusive_scan(v, s_warp, chunk_total);
			const uint32_t carry = s_carry;
			if (b < gridDim.x) out[ntiles + 1 + b] = carry + excl;  // survivors before scan block b
			__syncthreads();
			if (threadIdx.x == 0) s_carry = carry + chunk_total;
			__syncthreads();
		}
		const uint32_t total = s_carry;
		if (threadIdx.x == 0) {
			out[ntiles] = total;
			out[ntiles + 1 + gridDim.x] = total;
			if (pass == 0) {
				const uint32_t ran = st->n_active;                    // slots of the frame that just ran (n_slots unless BM_FRAME_EXACT_PATHS)
				const uint32_t progress = ran - st->primary_ray_cnt;  // kernel.cu:125
				st->start_position = (st->start_position + progress) % pixels;  // kernel.cu:130-131
				st->primary_ray_cnt = total;  // survivors written by shade (kernel.cu:298)
				st->frame += 1;               // kernel.cu:423
				st->tile_ticket = 0;
				st->cur = cur ^ 1;
				st->frames += 1;
				st->extend_rays += ran;
				if (st->target_paths && st->paths_since_reset >= st->target_paths) st->done = 1;
				st->n_active = next_frame_slots(st, n_slots, total);
			} else {
				st->shadow_ray_cnt = total;
			}
		}
		__syncthreads();
	}
	if (threadIdx.x == 0) *ticket = 0;
}

// start of a bm_render / bm_launch_frame call: install the stop target; a target that is already met stops at once
__global__ void begin_kernel(DeviceState* st, unsigned long long target_paths, uint32_t exact, uint32_t n_slots) {
	st->target_paths = target_paths;
	st->exact = exact;
	st->done = (target_paths && st->paths_since_reset >= target_paths) ? 1u : 0u;
	st->tile_ticket = 0;
	st->n_active = next_frame_slots(st, n_slots, st->primary_ray_cnt);
}

// sparse-by-slot -> dense (the layout the reference's queues have): dst[prefix[t] + rank of slot within tile t] = src[slot]
template <typename T>
__device__ __forceinline__ void export_tile(const T* src, const uint32_t* mask, const uint32_t* prefix, T* dst, uint32_t ntiles);
// the private survivor set the last executed frame wrote (st->cur, see scan_kernel)
__global__ void __launch_bounds__(kTile) export_rays_kernel(const FrameIO io, bm_ray* dst) {
	const bool one = io.st->cur != 0;  // (never io.dense_in: that is the INPUT of a frame whose survivors are exported here)
	export_tile<bm_ray>(one ? io.rays[1] : io.rays[0], one ? io.mask[1] : io.mask[0], one ? io.prefix[1] : io.prefix[0], dst, io.ntiles);
}
__global__ void __launch_bounds__(kTile) export_shadow_kernel(const bm_shadow* src, const uint32_t* mask, const uint32_t* prefix, bm_shadow* dst, uint32_t ntiles) {
	export_tile<bm_shadow>(src, mask, prefix, dst, ntiles);
}
template <typename T>
__device__ __forceinline__ void export_tile(const T* src, const uint32_t* mask, const uint32_t* prefix, T* dst, uint32_t ntiles) {
	const uint32_t tile = blockIdx.x;
	if (tile >= ntiles) return;
	const uint32_t* m = mask + (size_t)tile * (kTile / 32);
	const uint32_t w = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t word = m[w];
	if (!((word >> lane) & 1u)) return;
	uint32_t rank = __popc(word & ((1u << lane) - 1u));
	for (uint32_t k = 0; k < w; k++) rank += __popc(m[k]);
	dst[prefix[ntiles + 1 + tile / kScanTiles] + prefix[tile] + rank] = src[(size_t)tile * kTile + threadIdx.x];
}


// dense survivor set -> the private sparse representation: slots [0, count) all present
__global__ void import_masks_kernel(uint32_t* mask, uint32_t* prefix, uint32_t ntiles, uint32_t count) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < ntiles * 8) {
		const uint32_t first = i * 32;
		mask[i] = count >= first + 32 ? 0xFFFFFFFFu : (count > first ? ((1u << (count - first)) - 1u) : 0u);
	}
	const uint32_t per_block = (uint32_t)kScanTiles * kTile;  // slots per scan block
	if (i < ntiles) prefix[i] = min(i * (uint32_t)kTile, count) - min((i / kScanTiles) * per_block, count);
	if (i == ntiles) prefix[i] = count;
	if (i <= scan_block_count(ntiles)) prefix[ntiles + 1 + i] = (uint32_t)min((unsigned long long)i * per_block, (unsigned long long)count);
}

// upload, kernel.cu:141-151, with the count read on the device (the reference copies it to the host first,
// kernel.cu:408-409)
__global__ void upload_kernel(const SceneView sv, const bm_brick* bricks_queue, const uint32_t* indices_queue) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t count = min(*sv.load_queue_count, sv.queue_size);
	if (i >= count) return;
	const int px = sv.load_queue[3 * i], py = sv.load_queue[3 * i + 1], pz = sv.load_queue[3 * i + 2];
	const int sc = (px >> 4) + (py >> 4) * sv.supergrid_xy + (pz >> 4) * sv.supergrid_xy * sv.supergrid_xy;
	const int local = (px & 15) + (py & 15) * 16 + (pz & 15) * 256;
	const uint32_t word = indices_queue[i];
	const uint4* s = reinterpret_cast<const uint4*>(bricks_queue + i);
	uint4* d = reinterpret_cast<uint4*>(sv.bricks[sc] + (word & BM_BRICK_INDEX_BITS));
	d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; d[3] = s[3];
	sv.indices[sc][local] = word;
}

// Merge of the all-gathered request blocks (see include/brickmap_b200.h), for any world size and queue size.
// gathered: world blocks of (1 + 3q) int32 = {count, positions}. Entry order: rank-major, queue order inside; flat index i = r q + k.
// First occurrences are found with a STABLE sort by cell number (cub radix sort): the head of every run of equal keys is the
// earliest entry of that cell. An exclusive scan of the head flags in flat order gives each surviving entry its place.
__global__ void merge_keys_kernel(const SceneView sv, const int32_t* gathered, int world, uint32_t* keys, uint32_t* vals) {
	const uint32_t q = sv.queue_size;
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= (uint32_t)world * q) return;
	const uint32_t r = i / q, k = i % q;
	const int32_t* block = gathered + (size_t)r * (1 + 3 * (size_t)q);
	uint32_t key = 0xFFFFFFFFu;  // entries past a block's count sort to the end
	if (k < min((uint32_t)block[0], q)) {
		const int32_t* p = block + 1 + 3 * k;
		key = (uint32_t)p[0] + (uint32_t)sv.cells * ((uint32_t)p[1] + (uint32_t)sv.cells * (uint32_t)p[2]);
	}
	keys[i] = key;
	vals[i] = i;
}
__global__ void merge_heads_kernel(const uint32_t* sorted_keys, const uint32_t* sorted_vals, uint32_t n, uint32_t* first) {
	const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= n) return;
	const uint32_t key = sorted_keys[p];
	first[sorted_vals[p]] = (key != 0xFFFFFFFFu && (p == 0 || sorted_keys[p - 1] != key)) ? 1u : 0u;
}
__global__ void merge_write_kernel(const SceneView sv, const int32_t* gathered, int world, const uint32_t* first, const uint32_t* place) {
	const uint32_t q = sv.queue_size;
	const uint32_t n = (uint32_t)world * q;
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	if (i == n - 1) *sv.load_queue_count = place[i] + first[i];  // all unique requests; may exceed q, consumers clamp (kernel.cu:409)
	if (!first[i]) return;
	const int32_t* p = gathered + (size_t)(i / q) * (1 + 3 * (size_t)q) + 1 + 3 * (i % q);
	const int px = p[0], py = p[1], pz = p[2];
	const int sc = (px >> 4) + (py >> 4) * sv.supergrid_xy + (pz >> 4) * sv.supergrid_xy * sv.supergrid_xy;
	const int local = (px & 15) + (py & 15) * 16 + (pz & 15) * 256;
	uint32_t* word = sv.indices[sc] + local;
	const uint32_t out = place[i];
	if (out < q) {
		sv.load_queue[3 * out] = px; sv.load_queue[3 * out + 1] = py; sv.load_queue[3 * out + 2] = pz;
		atomicOr(word, BM_BRICK_REQUESTED_BIT);
	} else {
		atomicAnd(word, ~BM_BRICK_REQUESTED_BIT);  // did not make the cut: released for a later frame (voxel.cuh:237-240)
	}
}

// emptiness bitmap: one warp per word (32 consecutive blocks in x), bit set iff any index word of the block is non-zero;
// block coordinates count from the one-block border, whose bits are all set (SceneView::coarse)
__global__ void coarse_build_kernel(const SceneView sv, uint32_t* coarse) {
	const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t word = gid >> 5, lane = gid & 31;
	if (word >= sv.coarse_words) return;
	const int side = 1 << sv.coarse_shift;
	const int row = word / sv.coarse_roww, bx = (int)((word % sv.coarse_roww) << 5) + (int)lane - 1;
	const int by = row % sv.coarse_nby - 1, bz = row / sv.coarse_nby - 1;
	bool any = false;
	if (bx < 0 || by < 0 || bz < 0 || bx * side >= sv.cells || by * side >= sv.cells || bz * side >= sv.cells_height) any = true;  // border (and row padding)
	for (int z = 0; z < side && !any; z++)
		for (int y = 0; y < side && !any; y++)
			for (int x = 0; x < side; x++) {
				const int px = bx * side + x, py = by * side + y, pz = bz * side + z;
				if (px >= sv.cells || py >= sv.cells || pz >= sv.cells_height) continue;
				const int sc = (px >> 4) + (py >> 4) * sv.supergrid_xy + (pz >> 4) * sv.supergrid_xy * sv.supergrid_xy;
				const int local = (px & 15) + (py & 15) * 16 + (pz & 15) * 256;
				if (sv.indices[sc][local]) { any = true; break; }
			}
	const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, any);
	if (lane == 0) coarse[word] = ballot;
}

// per-cell emptiness, 64 bits per 4x4x4 cells of the biased position space, ones outside the world (SceneView::fine)
__global__ void fine_build_kernel(const SceneView sv, uint32_t* fine, uint32_t nblocks) {
	const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= nblocks) return;
	const int bias = 1 << sv.coarse_shift;
	const int bx = b % sv.fine_nx, by = (b / sv.fine_nx) % (sv.fine_nxy / sv.fine_nx), bz = b / sv.fine_nxy;
	uint32_t lo = 0, hi = 0;
	for (int z = 0; z < 4; z++)
		for (int y = 0; y < 4; y++)
			for (int x = 0; x < 4; x++) {
				const int px = bx * 4 + x - bias, py = by * 4 + y - bias, pz = bz * 4 + z - bias;
				bool set = true;
				if (px >= 0 && py >= 0 && pz >= 0 && px < sv.cells && py < sv.cells && pz < sv.cells_height) {
					const int sc = (px >> 4) + (py >> 4) * sv.supergrid_xy + (pz >> 4) * sv.supergrid_xy * sv.supergrid_xy;
					const int local = (px & 15) + (py & 15) * 16 + (pz & 15) * 256;
					set = sv.indices[sc][local] != 0;
				}
				if (set) {
					const int bit = x | (y << 2) | (z << 4);
					if (bit < 32) lo |= 1u << bit; else hi |= 1u << (bit - 32);
				}
			}
	fine[(size_t)b * 2] = lo;
	fine[(size_t)b * 2 + 1] = hi;
}

// SceneView::sky, step 1: per cell column (x, y) the highest non-empty cell z, -1 if the column is empty
__global__ void sky_cells_kernel(const SceneView sv, int16_t* cellmax) {
	const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
	if (x >= sv.cells) return;
	int top = -1;
	for (int z = sv.cells_height - 1; z >= 0; z--) {
		const int sc = (x >> 4) + (y >> 4) * sv.supergrid_xy + (z >> 4) * sv.supergrid_xy * sv.supergrid_xy;
		if (sv.indices[sc][(x & 15) + (y & 15) * 16 + (z & 15) * 256]) { top = z; break; }
	}
	cellmax[(size_t)y * sv.cells + x] = (int16_t)top;
}
// step 2: per column of (1 << shift)^2 cells the maximum over the column grown by two cells on every side
__global__ void sky_columns_kernel(const int16_t* cellmax, int cells, int shift, int n, int16_t* sky) {
	const int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= n * n) return;
	const int cx = c % n, cy = c / n, side = 1 << shift;
	int top = -1;
	for (int y = max(0, cy * side - 2); y < min(cells, (cy + 1) * side + 2); y++)
		for (int x = max(0, cx * side - 2); x < min(cells, (cx + 1) * side + 2); x++) top = max(top, (int)cellmax[(size_t)y * cells + x]);
	sky[c] = (int16_t)top;
}

// is indices[sc] == indices[0] + sc * 4096 for every superchunk?
__global__ void flat_check_kernel(uint32_t* const* indices, uint32_t n, uint32_t* not_flat) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n && indices[i] != indices[0] + (size_t)i * 4096) atomicExch(not_flat, 1u);
}

__global__ void trace_kernel(const SceneView sv, I3 cam, size_t n, const float* origins, const float* directions, float* normals, float* distances, uint8_t* hits) {
	extern __shared__ uint32_t s_coarse[];
	for (uint32_t i = threadIdx.x; i < sv.coarse_words; i += blockDim.x) s_coarse[i] = __ldg(sv.coarse + i);
	const uint32_t* coarse = s_coarse;
	__syncthreads();
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		F3 nrm{ normals[3 * i], normals[3 * i + 1], normals[3 * i + 2] };
		float dist = distances[i];
		WorkCounters wc{ 0, 0, 0, 0 };
		const bool h = intersect_voxel<false>(sv, coarse, F3{ origins[3 * i], origins[3 * i + 1], origins[3 * i + 2] },
		                                      F3{ directions[3 * i], directions[3 * i + 1], directions[3 * i + 2] }, nrm, dist, cam, &wc);
		normals[3 * i] = nrm.x; normals[3 * i + 1] = nrm.y; normals[3 * i + 2] = nrm.z;
		distances[i] = dist;
		hits[i] = h ? 1 : 0;
	}
}

__global__ void sky_kernel(const FrameParams fp, size_t n, const float* dirs, int mode, float* out) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const F3 d{ dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2] };
	const F3 c = mode == 0 ? sun_radiance(fp, d) : (mode == 1 ? sky_radiance(fp, d) : sunsky_radiance(fp, d));
	out[3 * i] = c.x; out[3 * i + 1] = c.y; out[3 * i + 2] = c.z;
}

// blit_onto_framebuffer, kernel.cu:348-364 (rgb / alpha, gamma 1/2.2, alpha = 1)
__global__ void tonemap_kernel(const float4* accum, float4* out, size_t pixels) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= pixels) return;
	const float4 c = accum[i];
	const float g = 1.f / 2.2f;
	out[i] = make_float4(powf(c.x / c.w, g), powf(c.y / c.w, g), powf(c.z / c.w, g), powf(1.f, g));
}

}  // namespace bm

// ================================================================================================================
// host side
// ================================================================================================================
using namespace bm;

static thread_local char g_error[256] = "";
static int fail_cuda(cudaError_t e, const char* what, int line) {
	snprintf(g_error, sizeof(g_error), "%s: %s (bm_kernels.cu:%d)", what, cudaGetErrorString(e), line);
	return (int)e;
}
static int fail_api(int code, const char* what) {
	snprintf(g_error, sizeof(g_error), "%s", what);
	return code;
}
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail_cuda(e_, #x, __LINE__); } while (0)

struct bm_context {
	bm_config cfg;
	cudaStream_t stream = nullptr;
	int sm_count = 0;
	uint32_t ntiles = 0;
	uint32_t tile_pixels = 0;
	DeviceState* d_state = nullptr;
	// private survivor buffers (tile-local layout), tile counts and prefixes, double-buffered
	bm_ray* d_rays[2] = { nullptr, nullptr };
	uint32_t* d_mask[2] = { nullptr, nullptr };    // 1 bit per slot: survivor present
	uint32_t* d_prefix[2] = { nullptr, nullptr };  // survivors before each 256-slot tile
	bm_shadow* d_shadow = nullptr;  // RECORD scratch (sparse by slot)
	uint32_t* d_shadow_mask = nullptr;
	uint32_t* d_shadow_prefix = nullptr;
	uint32_t* d_scan_totals = nullptr;  // scan_kernel: 2 * scan_blocks block totals
	uint32_t* d_scan_ticket = nullptr;
	uint32_t scan_blocks = 0;
	bool private_valid = false;  // survivors of the last frame are in the private set DeviceState::cur (tile-local)
	uint32_t caller_survivors = 0;  // primary_ray_cnt installed by bm_set_counters and not yet backed by records
	// scene
	bool bound = false;
	bm_gpu_scene scene{};
	SceneView sv{};
	uint32_t* d_coarse = nullptr;
	uint32_t* d_fine = nullptr;
	int16_t* d_sky = nullptr;
	uint32_t* d_flag = nullptr;
	// host statics of launch_kernels (kernel.cu:367-382)
	bm_camera cam{};
	bm_camera last_cam{};
	bool have_last = false;
	float sun_x = 0.05f, sun_y = 0.1f;  // variables.cpp:3
	bool sun_changed = true;            // variables.cpp:4
	FrameParams fp{};
	uint64_t launches = 0;
	// optional per-launch timing of frame_kernel
	bool timing = false;
	std::vector<cudaEvent_t> events;  // pairs
	size_t events_used = 0;
	double timed_ms = 0;
	uint64_t timed_launches = 0;
	int frame_blocks = 0;
	size_t frame_smem = 0;
	// throughput kernel (frame_kernel_q)
	bool use_quantum = true;  // BRICKMAP_B200_SIMPLE_KERNEL=1 switches the throughput path back to frame_kernel
	int quantum = 256;        // BRICKMAP_B200_QUANTUM: cell tests per lane and batch at most (with min_share 16: 64 -> 2440, 128 -> 2478, 256 -> 2494 Mrays/s)
	int min_share = 10;       // BRICKMAP_B200_MIN_SHARE: a batch is given up when fewer than min_share / 32 of its tracing lanes are left (profiles/r2_o_sweep_chunk32.txt, r2_u_sweep_sky_columns.txt)
	int run_len = 1;          // BRICKMAP_B200_RUN_LEN (1: 2723, 2: 2666, 4: 2503, 8: 2215, 16: 1683 Mrays/s -- the runs a warp still holds when the pool runs dry are the frame's tail)
	int resume_at = 28;       // BRICKMAP_B200_RESUME_AT (profiles/r2_j_sweep_resume_at.txt, r2_o_sweep_chunk32.txt)
	int brick_lanes = 3;      // BRICKMAP_B200_BRICK_LANES (0: 3089, 2: 3184, 3: 3200, 4: 3192, 6: 3157, 10: 3031, 16: 2886 Mrays/s): bricks reached by fewer lanes than this suspend the ray in front of the brick (0: never)
	int descending = 0;       // BRICKMAP_B200_DESCENDING: hand out slot runs from the end of the frame
	// scratch of bm_requests_merge (sized for world * queue entries on first use)
	uint32_t *m_keys = nullptr, *m_vals = nullptr, *m_keys_sorted = nullptr, *m_vals_sorted = nullptr, *m_first = nullptr, *m_place = nullptr;
	void* m_cub = nullptr;
	size_t m_cub_bytes = 0, m_entries = 0;
	int q_blocks = 0;
	bool q_stock = false;
	size_t q_smem = 0, q_smem_record = 0;
};

static int props_smem_optin(int device) {
	int v = 0;
	cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
	return v;
}

static inline F3 h_cross(const F3& a, const F3& b) { return F3{ a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y }; }
static inline F3 h_normalize(const F3& v) {
	const float r = 1.0f / sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
	return F3{ v.x * r, v.y * r, v.z * r };
}

// Per-frame constants: camera basis (kernel.cu:384-385), sun direction (kernel.cu:393, sunsky.cu:28-30) and the
// view-independent factors of the sky model (sunsky.cu:10-26,36-66) evaluated once in double precision.
static void update_frame_params(bm_context* c) {
	FrameParams& fp = c->fp;
	const bm_camera& cam = c->cam;
	const F3 dir{ cam.direction[0], cam.direction[1], cam.direction[2] }, up{ cam.up[0], cam.up[1], cam.up[2] };
	const float aspect = (float)(size_t)c->cfg.screen_width / (size_t)c->cfg.screen_height;
	const F3 rn = h_normalize(h_cross(dir, up));
	fp.cam_right = F3{ rn.x * 1.5f * aspect, rn.y * 1.5f * aspect, rn.z * 1.5f * aspect };
	const F3 un = h_normalize(h_cross(fp.cam_right, dir));
	fp.cam_up = F3{ un.x * 1.5f, un.y * 1.5f, un.z * 1.5f };
	fp.cam_dir = dir;
	fp.cam_pos = F3{ cam.position[0], cam.position[1], cam.position[2] };
	fp.focal3 = cam.focal_distance * 3.0f;
	fp.lens_radius = cam.lens_radius;
	fp.cam_cell = I3{ (int)(cam.position[0] / 8.f), (int)(cam.position[1] / 8.f), (int)(cam.position[2] / 8.f) };
	fp.width = c->cfg.screen_width;
	fp.height = c->cfg.screen_height;
	fp.tile_row0 = c->cfg.tile_row0;
	fp.tile_rows = c->cfg.tile_rows;
	fp.strip_rows = c->cfg.strip_rows;
	fp.strip_count = c->cfg.strip_count;
	fp.strip_index = c->cfg.strip_index;
	fp.n_slots = c->cfg.ray_queue_buffer_size;

	const float px = (c->sun_x - 0.0f) * 6.28f, py = (c->sun_y - 0.5f) * 3.14f;
	const F3 p{ cosf(px) * sinf(py), sinf(px) * sinf(py), cosf(py) };
	fp.sun_dir = h_normalize(p);
	fp.sun_angular_cos = cosf(1.5f * kPi / 180.f);  // sunSize (sunsky.cuh:25), kernel.cu:374
	fp.cone_extent = 1.0f - fp.sun_angular_cos;
	// SunIntensity(cosSunUp), sunsky.cu:24-26: cutoffAngle = pi/1.95, steepness 1.5, sunIntensity 1000
	const float cos_sun_up = fp.sun_dir.z;
	const float cutoff = kPi / 1.95f;
	const double e = 1.0 - (double)expf(-((cutoff - acosf(cos_sun_up)) / 1.5f));
	fp.sun_e = (float)(1000.0 * (e > 0.0 ? e : 0.0));
	fp.rayleigh = F3{ 5.176821E-6f, 1.2785348E-5f, 2.8530756E-5f };
	// totalMie(primaryWavelengths, K, turbidity=1) * mieCoefficient, sunsky.cu:14-18,44
	const float cc = (float)((0.2 * 1.0) * 10E-18);
	const float k = 0.434f * cc * kPi;
	const float lambda[3] = { 680E-9f, 550E-9f, 450E-9f }, K[3] = { 0.686f, 0.678f, 0.666f };
	float mie[3];
	for (int i = 0; i < 3; i++) mie[i] = (k * powf((2.0f * kPi) / lambda[i], 2.0f) * K[i]) * 0.005f;
	fp.mie = F3{ mie[0], mie[1], mie[2] };
	fp.total = F3{ fp.rayleigh.x + mie[0], fp.rayleigh.y + mie[1], fp.rayleigh.z + mie[2] };
	const float a = powf(1.0f - fp.sun_dir.z, 5.0f);
	fp.mix_a = a < 0.f ? 0.f : (a > 1.f ? 1.f : a);
	fp.rayleigh_k = (float)(3.0 / (16.0 * (double)kPi));
	fp.hg_g = 0.80f;
	fp.hg_k = (float)((1.0 / (4.0 * (double)kPi)) * (1.0 - (double)(0.80f * 0.80f)));
}

extern "C" {

const char* bm_last_error_string(void) { return g_error; }

void bm_default_config(bm_config* cfg) {
	memset(cfg, 0, sizeof(*cfg));
	cfg->device = 0;
	cfg->grid_size = 4096;                     // variables.h:7
	cfg->grid_height = 512;                    // variables.h:8
	cfg->lod_distance_2x2x2 = 100000;          // variables.h:27
	cfg->lod_distance_8x8x8 = 600000;          // variables.h:25
	cfg->brick_load_queue_size = 1024;         // variables.h:35
	cfg->ray_queue_buffer_size = 2 * 1048576;  // variables.h:61
	cfg->screen_width = 1920;
	cfg->screen_height = 1080;
	cfg->tile_row0 = 0;
	cfg->tile_rows = 0;
}

void bm_destroy(bm_context* c) {
	if (!c) return;
	cudaSetDevice(c->cfg.device);
	for (int i = 0; i < 2; i++) {
		cudaFree(c->d_rays[i]);
		cudaFree(c->d_mask[i]);
		cudaFree(c->d_prefix[i]);
	}
	cudaFree(c->d_shadow);
	cudaFree(c->d_shadow_mask);
	cudaFree(c->d_shadow_prefix);
	cudaFree(c->d_state);
	cudaFree(c->d_coarse);
	cudaFree(c->d_fine);
	cudaFree(c->d_sky);
	cudaFree(c->d_flag);
	cudaFree(c->d_scan_totals);
	cudaFree(c->d_scan_ticket);
	cudaFree(c->m_keys); cudaFree(c->m_vals); cudaFree(c->m_keys_sorted); cudaFree(c->m_vals_sorted); cudaFree(c->m_first); cudaFree(c->m_place); cudaFree(c->m_cub);
	for (cudaEvent_t e : c->events) cudaEventDestroy(e);
	if (c->stream) cudaStreamDestroy(c->stream);
	delete c;
}

int bm_create(bm_context** out, const bm_config* cfg) {
	if (!out || !cfg) return fail_api(BM_E_INVALID, "bm_create: null argument");
	if (cfg->grid_size <= 0 || cfg->grid_height <= 0 || cfg->grid_size % 128 || cfg->grid_height % 128)
		return fail_api(BM_E_INVALID, "bm_create: grid dimensions must be positive multiples of 128");
	if (!cfg->screen_width || !cfg->screen_height || !cfg->ray_queue_buffer_size || cfg->brick_load_queue_size <= 0)
		return fail_api(BM_E_INVALID, "bm_create: zero-sized image or queue");
	if (cfg->tile_rows && !cfg->strip_rows && cfg->tile_row0 + cfg->tile_rows > cfg->screen_height) return fail_api(BM_E_INVALID, "bm_create: tile exceeds the image");
	if (cfg->strip_rows) {
		if (!cfg->strip_count || cfg->strip_index >= cfg->strip_count || !cfg->tile_rows) return fail_api(BM_E_INVALID, "bm_create: bad strip partition");
		const uint32_t last = cfg->tile_rows - 1;
		const uint32_t y = ((last / cfg->strip_rows) * cfg->strip_count + cfg->strip_index) * cfg->strip_rows + last % cfg->strip_rows;
		if (y >= cfg->screen_height) return fail_api(BM_E_INVALID, "bm_create: strip partition exceeds the image");
	}
	bm_context* c = new (std::nothrow) bm_context();
	if (!c) return fail_api(BM_E_NOMEM, "bm_create: out of host memory");
	c->cfg = *cfg;
	if (!c->cfg.tile_rows) {
		c->cfg.tile_row0 = 0;
		c->cfg.tile_rows = cfg->screen_height;
	}
	c->tile_pixels = c->cfg.tile_rows * c->cfg.screen_width;
	const uint32_t n = c->cfg.ray_queue_buffer_size;
	c->ntiles = (n + kTile - 1) / kTile;
#define CKC(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { int rc_ = fail_cuda(e_, #x, __LINE__); bm_destroy(c); return rc_; } } while (0)
	CKC(cudaSetDevice(cfg->device));
	cudaDeviceProp props;
	CKC(cudaGetDeviceProperties(&props, cfg->device));
	c->sm_count = props.multiProcessorCount;
	// a BLOCKING stream, like the reference's kernel_stream (cudaStreamCreate, Scene.cpp:35): work a host enqueues on the
	// legacy default stream (the reference's staging copies, Scene.cpp:228-229; torch's default stream) is ordered with ours
	CKC(cudaStreamCreate(&c->stream));
	CKC(cudaMalloc(&c->d_state, sizeof(DeviceState)));
	CKC(cudaMemset(c->d_state, 0, sizeof(DeviceState)));
	const uint32_t one = 1;
	CKC(cudaMemcpy(&c->d_state->frame, &one, 4, cudaMemcpyHostToDevice));
	for (int i = 0; i < 2; i++) {
		CKC(cudaMalloc(&c->d_rays[i], (size_t)c->ntiles * kTile * sizeof(bm_ray)));
		CKC(cudaMalloc(&c->d_mask[i], (size_t)c->ntiles * 32));
		CKC(cudaMalloc(&c->d_prefix[i], prefix_words(c->ntiles) * 4));
		CKC(cudaMemset(c->d_mask[i], 0, (size_t)c->ntiles * 32));
		CKC(cudaMemset(c->d_prefix[i], 0, prefix_words(c->ntiles) * 4));
	}
	CKC(cudaMalloc(&c->d_shadow_mask, (size_t)c->ntiles * 32));
	CKC(cudaMemset(c->d_shadow_mask, 0, (size_t)c->ntiles * 32));
	CKC(cudaMalloc(&c->d_shadow_prefix, prefix_words(c->ntiles) * 4));
	CKC(cudaMalloc(&c->d_flag, 4));
	c->scan_blocks = scan_block_count(c->ntiles);
	CKC(cudaMalloc(&c->d_scan_totals, (size_t)c->scan_blocks * 8));
	CKC(cudaMalloc(&c->d_scan_ticket, 4));
	CKC(cudaMemset(c->d_scan_ticket, 0, 4));
#undef CKC
	// default camera (camera.h:4-9) and sun (variables.cpp:3)
	const bm_camera def = { { 512, 512, 300 }, { 1, 0, 0 }, { 0, 0, 1 }, 1.f, 0.f };
	c->cam = def;
	update_frame_params(c);
	*out = c;
	return 0;
}

int bm_scene_bind(bm_context* c, bm_gpu_scene scene) {
	if (!c) return fail_api(BM_E_INVALID, "bm_scene_bind: null context");
	if (!scene.indices || !scene.bricks || !scene.brick_load_queue || !scene.brick_load_queue_count)
		return fail_api(BM_E_INVALID, "bm_scene_bind: null scene pointer");
	CK(cudaSetDevice(c->cfg.device));
	c->scene = scene;
	SceneView& sv = c->sv;
	sv.indices = scene.indices;
	sv.bricks = scene.bricks;
	sv.load_queue = scene.brick_load_queue;
	sv.load_queue_count = scene.brick_load_queue_count;
	sv.cells = c->cfg.grid_size / 8;
	sv.cells_height = c->cfg.grid_height / 8;
	sv.supergrid_xy = sv.cells / 16;
	sv.grid_size_f = (float)c->cfg.grid_size;
	sv.grid_height_f = (float)c->cfg.grid_height;
	sv.lod2 = c->cfg.lod_distance_2x2x2;
	sv.lod8 = c->cfg.lod_distance_8x8x8;
	sv.queue_size = (uint32_t)c->cfg.brick_load_queue_size;
	// emptiness bitmap: the finest block size whose bitmap (with its one-block border) fits in 64 KiB of shared memory
	int shift = 0;
	uint64_t words, roww, nby;
	for (;; shift++) {
		const uint64_t nb = ((uint64_t)(sv.cells + (1 << shift) - 1) >> shift) + 2, nz = ((uint64_t)(sv.cells_height + (1 << shift) - 1) >> shift) + 2;
		roww = (nb + 31) / 32;
		nby = nb;
		words = (nz * nb * roww + 3) & ~(uint64_t)3;  // a multiple of 16 bytes: the bulk-copy granularity (stage_bitmap); the pad words are never addressed
		if (words * 4 <= 64u * 1024u) break;
	}
	sv.coarse_shift = shift;
	sv.coarse_nby = (int)nby;
	sv.coarse_roww = (int)roww;
	sv.coarse_words = (uint32_t)words;
	cudaFree(c->d_coarse);
	c->d_coarse = nullptr;
	CK(cudaMalloc(&c->d_coarse, (size_t)sv.coarse_words * 4));
	sv.coarse = nullptr;
	sv.flat_indices = nullptr;
	coarse_build_kernel<<<(sv.coarse_words * 32 + 255) / 256, 256, 0, c->stream>>>(sv, c->d_coarse);
	CK(cudaGetLastError());

	sv.fine = nullptr;
	uint32_t nfine;
	if (shift == 2) {  // the coarse bitmap's own grid: a pair's index is its block's bit index
		sv.fine_nx = sv.coarse_roww * 32;
		sv.fine_nxy = sv.fine_nx * sv.coarse_nby;
		nfine = sv.coarse_words * 32;
	} else {
		const int bias = 1 << shift;
		sv.fine_nx = (sv.cells + 2 * bias + 3) >> 2;
		sv.fine_nxy = sv.fine_nx * sv.fine_nx;
		nfine = (uint32_t)sv.fine_nxy * (uint32_t)((sv.cells_height + 2 * bias + 3) >> 2);
	}
	cudaFree(c->d_fine);
	c->d_fine = nullptr;
	CK(cudaMalloc(&c->d_fine, (size_t)nfine * 8));
	fine_build_kernel<<<(nfine + 255) / 256, 256, 0, c->stream>>>(sv, c->d_fine, nfine);
	CK(cudaGetLastError());
	c->launches += 1;
	// open-sky table (SceneView::sky): columns of 16 x 16 cells
	sv.sky = nullptr;
	sv.sky_shift = 4;
	if (const char* e = getenv("BRICKMAP_B200_SKY_SHIFT")) sv.sky_shift = atoi(e) >= 1 && atoi(e) <= 6 ? atoi(e) : sv.sky_shift;  // columns of 2^shift cells (tuning)
	sv.sky_n = sv.cells >> sv.sky_shift;
	sv.sky_top = sv.cells_height;
	cudaFree(c->d_sky);
	c->d_sky = nullptr;
	{
		int16_t* cellmax = nullptr;
		CK(cudaMalloc(&cellmax, (size_t)sv.cells * sv.cells * sizeof(int16_t)));
		CK(cudaMalloc(&c->d_sky, (size_t)sv.sky_n * sv.sky_n * sizeof(int16_t)));
		sky_cells_kernel<<<dim3((sv.cells + 127) / 128, sv.cells), 128, 0, c->stream>>>(sv, cellmax);
		CK(cudaGetLastError());
		sky_columns_kernel<<<(sv.sky_n * sv.sky_n + 127) / 128, 128, 0, c->stream>>>(cellmax, sv.cells, sv.sky_shift, sv.sky_n, c->d_sky);
		CK(cudaGetLastError());
		std::vector<int16_t> h((size_t)sv.sky_n * sv.sky_n);
		CK(cudaMemcpyAsync(h.data(), c->d_sky, h.size() * sizeof(int16_t), cudaMemcpyDeviceToHost, c->stream));
		CK(cudaStreamSynchronize(c->stream));
		CK(cudaFree(cellmax));
		c->launches += 2;
		int top = -1;
		size_t open = 0;  // columns with at least two free cell layers above everything
		for (int16_t v : h) {
			top = v > top ? v : top;
			open += v < sv.cells_height - 2;
		}
		sv.sky_top = top;
		// worth testing only where there is sky: a world filled to the top (caves) would pay for tests that never succeed
		bool enable = open * 8 >= h.size();
		if (const char* e = getenv("BRICKMAP_B200_NO_SKY")) enable = enable && e[0] != '1';
		if (enable) sv.sky = c->d_sky;
	}
	const uint32_t nsc = (uint32_t)(sv.supergrid_xy * sv.supergrid_xy * (sv.cells_height / 16));
	CK(cudaMemsetAsync(c->d_flag, 0, 4, c->stream));
	flat_check_kernel<<<(nsc + 255) / 256, 256, 0, c->stream>>>(scene.indices, nsc, c->d_flag);
	CK(cudaGetLastError());
	c->launches += 2;
	uint32_t not_flat = 1;
	uint32_t* first = nullptr;
	CK(cudaMemcpyAsync(&not_flat, c->d_flag, 4, cudaMemcpyDeviceToHost, c->stream));
	CK(cudaMemcpyAsync(&first, scene.indices, sizeof(uint32_t*), cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	sv.coarse = c->d_coarse;
	sv.fine = c->d_fine;
	sv.flat_indices = not_flat ? nullptr : first;
	// launch geometry of the persistent frame kernel
	c->frame_smem = (size_t)sv.coarse_words * 4 + (size_t)(kTile / 32) * 10 * kShadowQueue * 4;  // emptiness bitmap + per-warp shadow queues
	CK(cudaFuncSetAttribute(frame_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->frame_smem));
	CK(cudaFuncSetAttribute(frame_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->frame_smem));
	CK(cudaFuncSetAttribute(frame_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->frame_smem));
	CK(cudaFuncSetAttribute(frame_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->frame_smem));
	CK(cudaFuncSetAttribute(trace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->frame_smem));
	int per_sm = 0;
	CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, frame_kernel<false, false>, kTile, c->frame_smem));
	if (per_sm < 1) per_sm = 1;
	c->frame_blocks = c->sm_count * per_sm;
	if (const char* e = getenv("BRICKMAP_B200_SIMPLE_KERNEL")) c->use_quantum = e[0] != '1';
	if (const char* e = getenv("BRICKMAP_B200_QUANTUM")) c->quantum = atoi(e) > 0 ? atoi(e) : c->quantum;
	c->quantum = (c->quantum + kTraceChunk - 1) / kTraceChunk * kTraceChunk;
	if (const char* e = getenv("BRICKMAP_B200_RUN_LEN")) c->run_len = atoi(e) >= 1 && atoi(e) <= 64 ? atoi(e) : c->run_len;
	if (const char* e = getenv("BRICKMAP_B200_RESUME_AT")) c->resume_at = atoi(e) >= 1 && atoi(e) <= 32 ? atoi(e) : c->resume_at;
	if (const char* e = getenv("BRICKMAP_B200_BRICK_LANES")) c->brick_lanes = atoi(e) >= 0 && atoi(e) <= 32 ? atoi(e) : c->brick_lanes;
	if (const char* e = getenv("BRICKMAP_B200_DESCENDING")) c->descending = e[0] == '1';
	if (const char* e = getenv("BRICKMAP_B200_MIN_SHARE")) c->min_share = atoi(e) >= 0 && atoi(e) <= 32 ? atoi(e) : c->min_share;
	if (sv.cells + (2 << shift) > 65535 || sv.cells_height + (2 << shift) > 4095) c->use_quantum = false;  // queue entries pack the biased cell position into 16 + 16 + 12 bits
	c->q_stock = sv.coarse_shift == 2 && sv.coarse_nby == 130 && sv.coarse_roww == 5;  // the stock world's bitmap geometry is compiled in
	if (const char* e = getenv("BRICKMAP_B200_NO_STOCK")) c->q_stock = c->q_stock && e[0] != '1';  // A/B switch for profiling
	c->q_smem = (size_t)sv.coarse_words * 4 + (size_t)(kQBlock / 32) * E_WORDS * kQueueEntries * 4;
	c->q_smem_record = (size_t)sv.coarse_words * 4 + (size_t)(kQBlock / 32) * E_WORDS_RECORD * kQueueEntries * 4;
	if (c->q_smem_record > (size_t)props_smem_optin(c->cfg.device)) c->use_quantum = false;
	if (c->use_quantum) {
		CK(cudaFuncSetAttribute(frame_kernel_q<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->q_smem));
		CK(cudaFuncSetAttribute(frame_kernel_q<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->q_smem));
		CK(cudaFuncSetAttribute(frame_kernel_q<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->q_smem_record));
		CK(cudaFuncSetAttribute(frame_kernel_q<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->q_smem_record));
		CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, frame_kernel_q<false, true>, kQBlock, c->q_smem_record));
	}
	if (per_sm < 1) c->use_quantum = false;
	c->q_blocks = c->sm_count * (per_sm < 1 ? 1 : per_sm);
	c->bound = true;
	return 0;
}

int bm_set_camera(bm_context* c, const bm_camera* cam) {
	if (!c || !cam) return fail_api(BM_E_INVALID, "bm_set_camera: null argument");
	c->cam = *cam;
	update_frame_params(c);
	return 0;
}

int bm_set_sun(bm_context* c, float x, float y) {
	if (!c) return fail_api(BM_E_INVALID, "bm_set_sun: null context");
	c->sun_x = x;
	c->sun_y = y;
	c->sun_changed = true;  // kernel.cu:389
	update_frame_params(c);
	return 0;
}

int bm_get_counters(bm_context* c, bm_counters* out) {
	if (!c || !out) return fail_api(BM_E_INVALID, "bm_get_counters: null argument");
	CK(cudaSetDevice(c->cfg.device));
	CK(cudaStreamSynchronize(c->stream));
	DeviceState s;
	CK(cudaMemcpy(&s, c->d_state, sizeof(s), cudaMemcpyDeviceToHost));
	out->primary_ray_cnt = s.primary_ray_cnt;
	out->start_position = s.start_position;
	out->shadow_ray_cnt = s.shadow_ray_cnt;
	out->frame = s.frame;
	return 0;
}

int bm_set_counters(bm_context* c, const bm_counters* in) {
	if (!c || !in) return fail_api(BM_E_INVALID, "bm_set_counters: null argument");
	CK(cudaSetDevice(c->cfg.device));
	CK(cudaStreamSynchronize(c->stream));
	const uint32_t v[4] = { in->primary_ray_cnt, in->start_position, in->shadow_ray_cnt, in->frame };
	CK(cudaMemcpy(c->d_state, v, sizeof(v), cudaMemcpyHostToDevice));
	c->private_valid = false;  // the caller now owns the meaning of the survivor set (dense `queue`)
	c->caller_survivors = in->primary_ray_cnt;
	return 0;
}

int bm_get_stats(bm_context* c, bm_stats* out) {
	if (!c || !out) return fail_api(BM_E_INVALID, "bm_get_stats: null argument");
	CK(cudaSetDevice(c->cfg.device));
	CK(cudaStreamSynchronize(c->stream));
	DeviceState s;
	CK(cudaMemcpy(&s, c->d_state, sizeof(s), cudaMemcpyDeviceToHost));
	out->frames = s.frames;
	out->extend_rays = s.extend_rays;
	out->shadow_rays = s.shadow_rays;
	out->terminations = s.terminations;
	out->unoccluded = s.unoccluded;
	out->cell_steps = s.cell_steps;
	out->index_reads = s.index_reads;
	out->bricks_entered = s.bricks;
	out->requests = s.requests;
	out->kernel_launches = c->launches;
	return 0;
}

int bm_reset_stats(bm_context* c) {
	if (!c) return fail_api(BM_E_INVALID, "bm_reset_stats: null context");
	CK(cudaSetDevice(c->cfg.device));
	CK(cudaStreamSynchronize(c->stream));
	CK(cudaMemset(&c->d_state->frames, 0, 9 * sizeof(unsigned long long)));
	c->launches = 0;
	return 0;
}

void* bm_stream(bm_context* c) { return c ? (void*)c->stream : nullptr; }

int bm_synchronize(bm_context* c) {
	if (!c) return fail_api(BM_E_INVALID, "bm_synchronize: null context");
	CK(cudaStreamSynchronize(c->stream));
	return 0;
}

}  // extern "C"

// reset logic of launch_kernels (kernel.cu:387-403): camera / lens / sun change -> zero the accumulation buffer and
// primary_ray_cnt (NOT start_position, NOT frame)
static int maybe_reset(bm_context* c, float* blit, uint32_t flags) {
	bool reset = !c->have_last || memcmp(c->last_cam.position, c->cam.position, 12) != 0 || memcmp(c->last_cam.direction, c->cam.direction, 12) != 0 ||
	             c->last_cam.focal_distance != c->cam.focal_distance || c->last_cam.lens_radius != c->cam.lens_radius;
	if (!c->have_last) {
		// first call: last_pos/last_dir are zero-initialised statics, last_focaldistance = 1, last_lensradius = 0.02f
		// (kernel.cu:379-382) -> the comparison is true for any usable camera
		reset = true;
	}
	if (c->sun_changed) {
		c->sun_changed = false;
		reset = true;
	}
	c->last_cam = c->cam;
	c->have_last = true;
	if (reset && !(flags & BM_FRAME_NO_RESET)) {
		CK(cudaMemsetAsync(blit, 0, (size_t)c->tile_pixels * 16, c->stream));
		CK(cudaMemsetAsync(&c->d_state->primary_ray_cnt, 0, 4, c->stream));
		CK(cudaMemsetAsync(&c->d_state->paths_since_reset, 0, 8, c->stream));
		CK(cudaMemsetAsync(&c->d_state->done, 0, 4, c->stream));
	}
	return 0;
}

static int launch_upload(bm_context* c) {
	const uint32_t q = c->sv.queue_size;
	upload_kernel<<<(q + 255) / 256, 256, 0, c->stream>>>(c->sv, c->scene.bricks_queue, c->scene.indices_queue);
	CK(cudaGetLastError());
	CK(cudaMemsetAsync(c->scene.brick_load_queue_count, 0, 4, c->stream));  // kernel.cu:413
	c->launches += 1;
	return 0;
}

static int drain_events(bm_context* c) {
	if (c->events_used) {
		CK(cudaEventSynchronize(c->events[c->events_used - 1]));
		for (size_t i = 0; i + 1 < c->events_used; i += 2) {
			float ms = 0.f;
			CK(cudaEventElapsedTime(&ms, c->events[i], c->events[i + 1]));
			c->timed_ms += ms;
			c->timed_launches += 1;
		}
		c->events_used = 0;
	}
	return 0;
}

static FrameIO private_io(bm_context* c, float* blit) {
	FrameIO io{};
	io.st = c->d_state;
	for (int i = 0; i < 2; i++) {
		io.rays[i] = c->d_rays[i];
		io.mask[i] = c->d_mask[i];
		io.prefix[i] = c->d_prefix[i];
	}
	io.accum = reinterpret_cast<float4*>(blit);
	io.ntiles = c->ntiles;
	return io;
}

// no private survivor set (first frame, or after bm_set_counters without records): both sets empty, set 0 current
static int clear_private_sets(bm_context* c) {
	for (int i = 0; i < 2; i++) {
		CK(cudaMemsetAsync(c->d_prefix[i], 0, prefix_words(c->ntiles) * 4, c->stream));
		CK(cudaMemsetAsync(c->d_mask[i], 0, (size_t)c->ntiles * 32, c->stream));
	}
	CK(cudaMemsetAsync(&c->d_state->cur, 0, 4, c->stream));
	return 0;
}

template <bool RECORD>
static int launch_frame_kernels(bm_context* c, const FrameIO& io, bool count) {
	const int blocks = (int)(c->ntiles < (uint32_t)c->frame_blocks ? c->ntiles : (uint32_t)c->frame_blocks);
	cudaEvent_t e0 = nullptr, e1 = nullptr;
	if (c->timing) {
		if (c->events_used + 2 > 8192) {
			const int rc = drain_events(c);
			if (rc) return rc;
		}
		while (c->events.size() < c->events_used + 2) {
			cudaEvent_t e;
			CK(cudaEventCreate(&e));
			c->events.push_back(e);
		}
		e0 = c->events[c->events_used];
		e1 = c->events[c->events_used + 1];
		c->events_used += 2;
		CK(cudaEventRecord(e0, c->stream));
	}
	if (count) frame_kernel<RECORD, true><<<blocks, kTile, c->frame_smem, c->stream>>>(c->fp, c->sv, io);
	else if (!c->use_quantum) frame_kernel<RECORD, false><<<blocks, kTile, c->frame_smem, c->stream>>>(c->fp, c->sv, io);
	else {
		const size_t smem = RECORD ? c->q_smem_record : c->q_smem;
		const QSched sch{ c->quantum, c->min_share, c->descending, c->run_len, c->resume_at, c->brick_lanes };
		if (RECORD) CK(cudaMemsetAsync(io.shadow_mask, 0, (size_t)c->ntiles * 32, c->stream));  // the kernel sets bits with atomicOr
		if (c->q_stock) frame_kernel_q<true, RECORD><<<c->q_blocks, kQBlock, smem, c->stream>>>(c->fp, c->sv, io, sch);
		else frame_kernel_q<false, RECORD><<<c->q_blocks, kQBlock, smem, c->stream>>>(c->fp, c->sv, io, sch);
	}
	CK(cudaGetLastError());
	if (c->timing) CK(cudaEventRecord(e1, c->stream));
	// counts the survivors, advances the cursor, toggles DeviceState::cur; the mask that was this frame's input becomes the next
	// frame's output: the scan clears it
	scan_kernel<<<c->scan_blocks, kScanTiles, 0, c->stream>>>(io, RECORD ? io.shadow_mask : nullptr, RECORD ? c->d_shadow_prefix : nullptr, c->ntiles,
	                                                          c->cfg.ray_queue_buffer_size, c->tile_pixels, c->d_scan_totals, c->d_scan_ticket);
	CK(cudaGetLastError());
	c->launches += 2;
	return 0;
}

extern "C" {

int bm_launch_frame(bm_context* c, float* blit, bm_ray* queue, bm_ray* queue2, bm_shadow* shadow_queue, uint32_t flags) {
	if (!c || !blit || !queue || !queue2 || !shadow_queue) return fail_api(BM_E_INVALID, "bm_launch_frame: null argument");
	if (!c->bound) return fail_api(BM_E_STATE, "bm_launch_frame: no scene bound (bm_scene_bind)");
	CK(cudaSetDevice(c->cfg.device));
	if (!c->d_shadow) CK(cudaMalloc(&c->d_shadow, (size_t)c->ntiles * kTile * sizeof(bm_shadow)));
	int rc = maybe_reset(c, blit, flags);
	if (rc) return rc;
	if (!(flags & BM_FRAME_NO_UPLOAD) && c->scene.bricks_queue && c->scene.indices_queue) {
		rc = launch_upload(c);
		if (rc) return rc;
	}
	begin_kernel<<<1, 1, 0, c->stream>>>(c->d_state, 0ull, 0u, c->cfg.ray_queue_buffer_size);
	CK(cudaGetLastError());
	c->launches += 1;
	FrameIO io = private_io(c, blit);
	if (!c->private_valid) {
		// dense survivors [0, primary_ray_cnt) supplied by the caller in `queue` (after bm_set_counters); otherwise the survivors
		// of the previous frame are still in the private set and the caller's swapped `queue` (main.cpp:146) is not needed
		rc = clear_private_sets(c);
		if (rc) return rc;
		io.dense_in = queue;
	}
	io.record = queue;
	io.shadow_out = c->d_shadow;
	io.shadow_mask = c->d_shadow_mask;
	io.extend_only = (flags & BM_FRAME_EXTEND_ONLY) ? 1u : 0u;
	rc = launch_frame_kernels<true>(c, io, (flags & BM_FRAME_COUNT_WORK) != 0);
	if (rc) return rc;
	export_rays_kernel<<<c->ntiles, kTile, 0, c->stream>>>(io, queue2);
	CK(cudaGetLastError());
	export_shadow_kernel<<<c->ntiles, kTile, 0, c->stream>>>(c->d_shadow, c->d_shadow_mask, c->d_shadow_prefix, shadow_queue, c->ntiles);
	CK(cudaGetLastError());
	c->launches += 2;
	c->caller_survivors = 0;
	c->private_valid = true;
	CK(cudaStreamSynchronize(c->stream));  // kernel.cu:431
	return 0;
}

int bm_render(bm_context* c, float* blit, uint32_t frames, uint64_t target_paths, uint32_t flags, int sync) {
	if (!c || !blit) return fail_api(BM_E_INVALID, "bm_render: null argument");
	if (!c->bound) return fail_api(BM_E_STATE, "bm_render: no scene bound (bm_scene_bind)");
	if (!c->private_valid && c->caller_survivors)
		return fail_api(BM_E_STATE, "bm_render: bm_set_counters installed primary_ray_cnt > 0 without records (bm_import_rays, or use bm_launch_frame)");
	CK(cudaSetDevice(c->cfg.device));
	int rc = maybe_reset(c, blit, flags);
	if (rc) return rc;
	if ((flags & BM_FRAME_EXACT_PATHS) && !target_paths) return fail_api(BM_E_INVALID, "bm_render: BM_FRAME_EXACT_PATHS needs target_paths");
	begin_kernel<<<1, 1, 0, c->stream>>>(c->d_state, (unsigned long long)target_paths, (flags & BM_FRAME_EXACT_PATHS) ? 1u : 0u, c->cfg.ray_queue_buffer_size);
	CK(cudaGetLastError());
	c->launches += 1;
	if (!c->private_valid) {
		// no private survivor set (first frame, or the counters were set by the caller): start from an empty one
		CK(cudaMemsetAsync(&c->d_state->primary_ray_cnt, 0, 4, c->stream));
		rc = clear_private_sets(c);
		if (rc) return rc;
		c->private_valid = true;
	}
	const FrameIO io = private_io(c, blit);
	for (uint32_t f = 0; f < frames; f++) {
		// at most one batch of staged bricks exists per call (the host stages between calls, Scene.cpp:200-229): the upload step
		// (kernel.cu:407-414) runs before the first frame only; later frames of the call append new requests to the emptied queue
		if (f == 0 && !(flags & BM_FRAME_NO_UPLOAD) && c->scene.bricks_queue && c->scene.indices_queue) {
			rc = launch_upload(c);
			if (rc) return rc;
		}
		rc = launch_frame_kernels<false>(c, io, (flags & BM_FRAME_COUNT_WORK) != 0);
		if (rc) return rc;
	}
	if (sync) CK(cudaStreamSynchronize(c->stream));
	return 0;
}

int bm_extend_primaries(bm_context* c, bm_ray* queue, uint32_t frames, int sync) {
	if (!c || !queue) return fail_api(BM_E_INVALID, "bm_extend_primaries: null argument");
	if (!c->bound) return fail_api(BM_E_STATE, "bm_extend_primaries: no scene bound (bm_scene_bind)");
	CK(cudaSetDevice(c->cfg.device));
	// every frame starts from an empty survivor set and leaves one (nothing is shaded)
	CK(cudaMemsetAsync(&c->d_state->primary_ray_cnt, 0, 4, c->stream));
	int rc = clear_private_sets(c);
	if (rc) return rc;
	c->private_valid = true;
	c->caller_survivors = 0;
	begin_kernel<<<1, 1, 0, c->stream>>>(c->d_state, 0ull, 0u, c->cfg.ray_queue_buffer_size);
	CK(cudaGetLastError());
	c->launches += 1;
	FrameIO io = private_io(c, nullptr);
	io.record = queue;
	io.shadow_mask = c->d_shadow_mask;
	io.extend_only = 1u;
	for (uint32_t f = 0; f < frames; f++) {
		rc = launch_frame_kernels<true>(c, io, false);
		if (rc) return rc;
	}
	if (sync) CK(cudaStreamSynchronize(c->stream));
	return 0;
}

int bm_import_rays(bm_context* c, const bm_ray* queue, uint32_t count) {
	if (!c || (count && !queue) || count > c->cfg.ray_queue_buffer_size) return fail_api(BM_E_INVALID, "bm_import_rays: bad argument");
	CK(cudaSetDevice(c->cfg.device));
	if (count) CK(cudaMemcpyAsync(c->d_rays[0], queue, (size_t)count * sizeof(bm_ray), cudaMemcpyDeviceToDevice, c->stream));
	import_masks_kernel<<<(c->ntiles * 8 + 256) / 256, 256, 0, c->stream>>>(c->d_mask[0], c->d_prefix[0], c->ntiles, count);
	CK(cudaGetLastError());
	CK(cudaMemsetAsync(c->d_mask[1], 0, (size_t)c->ntiles * 32, c->stream));
	CK(cudaMemsetAsync(&c->d_state->cur, 0, 4, c->stream));
	CK(cudaMemcpyAsync(&c->d_state->primary_ray_cnt, &c->d_prefix[0][c->ntiles], 4, cudaMemcpyDeviceToDevice, c->stream));
	c->caller_survivors = 0;
	c->launches += 1;
	c->private_valid = true;
	return 0;
}

int bm_export_rays(bm_context* c, bm_ray* queue2) {
	if (!c || !queue2) return fail_api(BM_E_INVALID, "bm_export_rays: null argument");
	if (!c->private_valid) return fail_api(BM_E_STATE, "bm_export_rays: no private survivor set");
	CK(cudaSetDevice(c->cfg.device));
	export_rays_kernel<<<c->ntiles, kTile, 0, c->stream>>>(private_io(c, nullptr), queue2);
	CK(cudaGetLastError());
	c->launches += 1;
	CK(cudaStreamSynchronize(c->stream));
	return 0;
}

int bm_render_to_host(bm_context* c, float* blit, uint32_t frames, uint64_t target_paths, uint32_t flags, float* accum_host, uint32_t* request_count_host,
                      int32_t* request_positions_host) {
	int rc = bm_render(c, blit, frames, target_paths, flags, 0);
	if (rc) return rc;
	if (accum_host) CK(cudaMemcpyAsync(accum_host, blit, (size_t)c->tile_pixels * 16, cudaMemcpyDeviceToHost, c->stream));
	if (request_count_host) CK(cudaMemcpyAsync(request_count_host, c->scene.brick_load_queue_count, 4, cudaMemcpyDeviceToHost, c->stream));
	if (request_positions_host)
		CK(cudaMemcpyAsync(request_positions_host, c->scene.brick_load_queue, (size_t)c->sv.queue_size * 12, cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	return 0;
}

int bm_requests_pack(bm_context* c, int32_t* block) {
	if (!c || !block) return fail_api(BM_E_INVALID, "bm_requests_pack: null argument");
	if (!c->bound) return fail_api(BM_E_STATE, "bm_requests_pack: no scene bound (bm_scene_bind)");
	CK(cudaSetDevice(c->cfg.device));
	CK(cudaMemcpyAsync(block, c->scene.brick_load_queue_count, 4, cudaMemcpyDeviceToDevice, c->stream));
	CK(cudaMemcpyAsync(block + 1, c->scene.brick_load_queue, (size_t)c->sv.queue_size * 12, cudaMemcpyDeviceToDevice, c->stream));
	return 0;
}

int bm_requests_merge(bm_context* c, const int32_t* gathered, int world) {
	if (!c || !gathered || world < 1) return fail_api(BM_E_INVALID, "bm_requests_merge: bad argument");
	if (!c->bound) return fail_api(BM_E_STATE, "bm_requests_merge: no scene bound (bm_scene_bind)");
	CK(cudaSetDevice(c->cfg.device));
	const size_t n = (size_t)world * c->sv.queue_size;
	if (n > (size_t)1 << 30) return fail_api(BM_E_INVALID, "bm_requests_merge: world_size * queue_size too large");
	if ((uint64_t)c->sv.cells * c->sv.cells * c->sv.cells_height >= 0xFFFFFFFFull) return fail_api(BM_E_INVALID, "bm_requests_merge: world too large for 32-bit cell keys");
	if (n > c->m_entries) {
		cudaFree(c->m_keys); cudaFree(c->m_vals); cudaFree(c->m_keys_sorted); cudaFree(c->m_vals_sorted); cudaFree(c->m_first); cudaFree(c->m_place); cudaFree(c->m_cub);
		c->m_keys = c->m_vals = c->m_keys_sorted = c->m_vals_sorted = c->m_first = c->m_place = nullptr;
		c->m_cub = nullptr;
		c->m_entries = 0;
		CK(cudaMalloc(&c->m_keys, n * 4)); CK(cudaMalloc(&c->m_vals, n * 4)); CK(cudaMalloc(&c->m_keys_sorted, n * 4));
		CK(cudaMalloc(&c->m_vals_sorted, n * 4)); CK(cudaMalloc(&c->m_first, n * 4)); CK(cudaMalloc(&c->m_place, n * 4));
		size_t a = 0, b = 0;
		CK(cub::DeviceRadixSort::SortPairs(nullptr, a, c->m_keys, c->m_keys_sorted, c->m_vals, c->m_vals_sorted, (int)n));
		CK(cub::DeviceScan::ExclusiveSum(nullptr, b, c->m_first, c->m_place, (int)n));
		c->m_cub_bytes = a > b ? a : b;
		CK(cudaMalloc(&c->m_cub, c->m_cub_bytes ? c->m_cub_bytes : 16));
		c->m_entries = n;
	}
	const unsigned blocks = (unsigned)((n + 255) / 256);
	size_t bytes = c->m_cub_bytes;
	merge_keys_kernel<<<blocks, 256, 0, c->stream>>>(c->sv, gathered, world, c->m_keys, c->m_vals);
	CK(cudaGetLastError());
	CK(cub::DeviceRadixSort::SortPairs(c->m_cub, bytes, c->m_keys, c->m_keys_sorted, c->m_vals, c->m_vals_sorted, (int)n, 0, 32, c->stream));
	merge_heads_kernel<<<blocks, 256, 0, c->stream>>>(c->m_keys_sorted, c->m_vals_sorted, (uint32_t)n, c->m_first);
	CK(cudaGetLastError());
	bytes = c->m_cub_bytes;
	CK(cub::DeviceScan::ExclusiveSum(c->m_cub, bytes, c->m_first, c->m_place, (int)n, c->stream));
	merge_write_kernel<<<blocks, 256, 0, c->stream>>>(c->sv, gathered, world, c->m_first, c->m_place);
	CK(cudaGetLastError());
	c->launches += 3;
	return 0;
}

int bm_kernel_timing(bm_context* c, int enable) {
	if (!c) return fail_api(BM_E_INVALID, "bm_kernel_timing: null context");
	CK(cudaSetDevice(c->cfg.device));
	const int rc = drain_events(c);
	if (rc) return rc;
	c->timing = enable != 0;
	c->timed_ms = 0;
	c->timed_launches = 0;
	return 0;
}

int bm_kernel_time(bm_context* c, double* ms_sum, uint64_t* launches) {
	if (!c || !ms_sum || !launches) return fail_api(BM_E_INVALID, "bm_kernel_time: null argument");
	CK(cudaSetDevice(c->cfg.device));
	const int rc = drain_events(c);
	if (rc) return rc;
	*ms_sum = c->timed_ms;
	*launches = c->timed_launches;
	c->timed_ms = 0;
	c->timed_launches = 0;
	return 0;
}

int bm_read_requests(bm_context* c, uint32_t* count_host, int32_t* positions_host) {
	if (!c) return fail_api(BM_E_INVALID, "bm_read_requests: null context");
	if (!c->bound) return fail_api(BM_E_STATE, "bm_read_requests: no scene bound (bm_scene_bind)");
	CK(cudaSetDevice(c->cfg.device));
	if (count_host) CK(cudaMemcpyAsync(count_host, c->scene.brick_load_queue_count, 4, cudaMemcpyDeviceToHost, c->stream));
	if (positions_host) CK(cudaMemcpyAsync(positions_host, c->scene.brick_load_queue, (size_t)c->sv.queue_size * 12, cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	return 0;
}

int bm_trace(bm_context* c, size_t n, const float* origins, const float* directions, float* normals, float* distances, uint8_t* hits) {
	if (!c || (n && (!origins || !directions || !normals || !distances || !hits))) return fail_api(BM_E_INVALID, "bm_trace: null argument");
	if (!c->bound) return fail_api(BM_E_STATE, "bm_trace: no scene bound (bm_scene_bind)");
	if (!n) return 0;
	CK(cudaSetDevice(c->cfg.device));
	size_t blocks = (n + kTile - 1) / kTile;
	if (blocks > (size_t)c->frame_blocks) blocks = (size_t)c->frame_blocks;
	trace_kernel<<<(unsigned)blocks, kTile, c->frame_smem, c->stream>>>(c->sv, c->fp.cam_cell, n, origins, directions, normals, distances, hits);
	CK(cudaGetLastError());
	c->launches += 1;
	CK(cudaStreamSynchronize(c->stream));
	return 0;
}

int bm_eval_sky(bm_context* c, size_t n, const float* dirs, int mode, float* out) {
	if (!c || mode < 0 || mode > 2 || (n && (!dirs || !out))) return fail_api(BM_E_INVALID, "bm_eval_sky: bad argument");
	if (!n) return 0;
	CK(cudaSetDevice(c->cfg.device));
	sky_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->fp, n, dirs, mode, out);
	CK(cudaGetLastError());
	c->launches += 1;
	CK(cudaStreamSynchronize(c->stream));
	return 0;
}

int bm_tonemap(bm_context* c, const float* blit, float* out) {
	if (!c || !blit || !out) return fail_api(BM_E_INVALID, "bm_tonemap: null argument");
	CK(cudaSetDevice(c->cfg.device));
	tonemap_kernel<<<(c->tile_pixels + 255) / 256, 256, 0, c->stream>>>(reinterpret_cast<const float4*>(blit), reinterpret_cast<float4*>(out), c->tile_pixels);
	CK(cudaGetLastError());
	c->launches += 1;
	CK(cudaStreamSynchronize(c->stream));
	return 0;
}

}  // extern "C"
