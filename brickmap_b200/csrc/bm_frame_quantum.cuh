// frame_kernel_q: the throughput form of one frame. Same results as frame_kernel (bit for bit on everything geometric).
//
// Why: in frame_kernel a warp traces 32 rays until the LONGEST of them ends; on the benchmark workload that leaves 43 % of
// the lanes busy (ray lengths: median 79 cell steps, 90th percentile 215, maximum 830). Here a warp traces a BATCH of 32 rays,
// and a ray that has not ended is SUSPENDED when it has used its budget of `quantum` cell tests or -- the rule that fires in
// practice -- as soon as fewer than min_share / 32 of the lanes that started the batch are still tracing. Its traversal state
// (position, tmax, axis -- tdelta and the step signs are recomputed from the direction) and its payload go into the warp's
// work queue in shared memory; whenever `resume_at` (32) suspended rays have piled up they are resumed together, otherwise
// the warp takes a ticket for the next 32 fresh slots of the frame. Shadow rays go through the same queue (they are born
// into it by shade), so they too are traced 32 at a time. Suspending never changes a result: the DDA state is saved and
// restored exactly. Measured on the benchmark view: lanes active in the cell loop 13 -> 24 of 32.
//
// Only rays that start inside the world (tminn == 0, voxel.cuh:136-141) are ever suspended: for them the trace-space origin is
// the world-space origin times 1/8 exactly, so the entry needs neither tminn nor the world-space origin that shading wants. A
// ray that enters from outside gets an unbounded quantum (all primaries of an outside camera: the kernel then behaves like
// frame_kernel for them).
// One queue entry = 16 words, structure of arrays per warp, 64 entries (a batch pops at most 32 and every lane pushes at most
// one entry per batch: the ray it had to suspend, or the shadow ray its shaded vertex produced).
#pragma once

namespace bm {

#ifndef BM_QBLOCK
#define BM_QBLOCK 1024
#endif
constexpr int kQBlock = BM_QBLOCK;  // 32 warps, one block per SM (shared memory: bitmap 46 KiB + 32 queues 128 KiB).
                                    // Measured on the benchmark view (128-slot tickets): 512 threads 2.06, 768 2.35, 1024 2.49 Grays/s
constexpr int kQueueEntries = 64;
enum : int {
	E_OX = 0, E_OY, E_OZ,   // trace-space origin in cell units (continuations) / world-space origin (new shadow rays)
	E_DX, E_DY, E_DZ,       // direction
	E_POS,                  // (biased) cell position x | y << 16; z is in E_FLAGS
	E_TX, E_TY, E_TZ,       // tmax
	E_FLAGS,                // kind | (last stepped axis + 1) << 2 | cell position z << 4 (12 bits) | bounces << 16
	E_CX, E_CY, E_CZ,       // throughput of an extend ray / colour of a shadow ray
	E_PIXEL, E_SLOT,
	E_WORDS,                // entry size of the throughput instantiation
	E_NX = E_WORDS, E_NY, E_NZ,  // RECORD only: the normal as extend has left it so far (part of the post-extend record of every slot;
	                             // without RECORD a suspended ray's normal is dead: any later hit overwrites it, a miss does not use it)
	E_WORDS_RECORD
};
enum : int { K_NONE = -1, K_EXTEND = 0, K_SHADOW = 1, K_SHADOW_NEW = 2 };

#define QF(field, e) q[(field) * kQueueEntries + (e)]
#define QU(field, e) reinterpret_cast<uint32_t*>(q)[(field) * kQueueEntries + (e)]

// Scheduling parameters of frame_kernel_q (none of them can change a result; defaults = measured best)
struct QSched {
	int quantum;       // cell tests a ray may take in one batch before it is suspended
	int min_share;     // a batch is given up once fewer than min_share / 32 of the lanes that started tracing are left
	int descending;    // hand out slot runs from the end of the frame
	int run_len;       // a warp pulls run_len * 32 consecutive slots per ticket
	int resume_at;     // queued rays are resumed as soon as this many have piled up (<= 32)
	int brick_lanes;   // a brick reached by fewer lanes than this (in the same iteration) suspends the ray in front of it (0: never)
};

// The block's copy of the emptiness bitmap (46 KiB at reference dims). BM_BULK_PROLOGUE=1: ONE bulk asynchronous copy global -> shared
// (cp.async.bulk, the 1-D form of TMA; SASS UBLKCP) issued by thread 0 and awaited by all threads on an mbarrier, instead of the
// cooperative __ldg / st.shared loop. `words` is a multiple of 4 (16-byte granularity of the bulk copy), both addresses are 16-byte aligned.
#ifndef BM_BULK_PROLOGUE
#define BM_BULK_PROLOGUE 1  // same speed as the loop (2747 vs 2748 Mrays/s, profiles/r2_c_ab_bulk_fastrad.txt): kept as the B200-idiomatic form
#endif
__device__ __forceinline__ void stage_bitmap(uint32_t* s_coarse, const uint32_t* g_coarse, uint32_t words) {
#if BM_BULK_PROLOGUE
	__shared__ __align__(8) unsigned long long s_bar;
	const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar), dst = (uint32_t)__cvta_generic_to_shared(s_coarse), bytes = words * 4u;
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
		asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(g_coarse), "r"(bytes), "r"(bar) : "memory");
	}
	asm volatile(
	    "{\n\t"
	    ".reg .pred p;\n\t"
	    "BM_WAIT_BITMAP:\n\t"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
	    "@p bra BM_WAIT_DONE;\n\t"
	    "bra BM_WAIT_BITMAP;\n\t"
	    "BM_WAIT_DONE:\n\t"
	    "}" ::"r"(bar) : "memory");
#else
	for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) s_coarse[i] = __ldg(g_coarse + i);
	__syncthreads();
#endif
}

// RECORD (bm_launch_frame): additionally leaves the post-extend record of every slot and the shadow-ray records, like frame_kernel<RECORD>
template <bool STOCK, bool RECORD>
__global__ void __launch_bounds__(kQBlock, 1) frame_kernel_q(const FrameParams fp, const SceneView sv, const FrameIO io, const QSched sch) {
	const int quantum = sch.quantum, min_share = sch.min_share, descending = sch.descending;
	extern __shared__ __align__(16) uint32_t s_coarse[];
	DeviceState* st = io.st;
	if (st->done) return;
	stage_bitmap(s_coarse, sv.coarse, sv.coarse_words);
	const uint32_t* coarse = s_coarse;

	const uint32_t c = st->primary_ray_cnt;
	const uint32_t start = st->start_position;
	const uint32_t frame = st->frame;
	const uint32_t cur = st->cur;
	const uint32_t n_slots = st->n_active;
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t lt_mask = (1u << lane) - 1u;
	float* q = reinterpret_cast<float*>(s_coarse + sv.coarse_words) + (threadIdx.x >> 5) * ((RECORD ? E_WORDS_RECORD : E_WORDS) * kQueueEntries);
	uint32_t qn = 0;  // entries in the warp's queue (warp-uniform)
	uint32_t n_shadow = 0, n_term = 0, n_unocc = 0;

	const uint32_t kRun = (uint32_t)sch.run_len, resume_at = (uint32_t)sch.resume_at;  // warps pull runs of kRun * 32 consecutive slots with one atomic
	const uint32_t nruns = (n_slots + kRun * 32 - 1) / (kRun * 32);
	uint32_t run = 0, round = kRun;
	bool pool_dry = false;

	for (;;) {
		int kind = K_NONE;
		int status = TRACE_MISS;
		bool tracing = false;
		TraceState ts;
		F3 direction{ 0.f, 0.f, 1.f }, normal{ 0.f, 0.f, 0.f }, world{ 0.f, 0.f, 0.f }, payload{ 0.f, 0.f, 0.f };
		float distance = kVeryFar;
		uint32_t pixel = 0, slot = 0;
		int bounces = 0;

		if (qn >= resume_at || (pool_dry && qn > 0)) {
			// ---- resume a batch of queued rays --------------------------------------------------------------------------
			const uint32_t take = min(qn, 32u);
			qn -= take;
			if (lane < take) {
				const uint32_t e = qn + lane;
				const uint32_t flags = QU(E_FLAGS, e);
				kind = (int)(flags & 3u);
				bounces = (int)(flags >> 16);
				direction = F3{ QF(E_DX, e), QF(E_DY, e), QF(E_DZ, e) };
				payload = F3{ QF(E_CX, e), QF(E_CY, e), QF(E_CZ, e) };
				pixel = QU(E_PIXEL, e);
				if (kind == K_SHADOW_NEW) {
					// connect, kernel.cu:328-346: origin = shaded hit point, normal y{} = 0, t = 0
					kind = K_SHADOW;
					tracing = trace_setup(sv, F3{ QF(E_OX, e), QF(E_OY, e), QF(E_OZ, e) }, direction, normal, ts);
				} else {
					if (kind == K_EXTEND) {
						slot = QU(E_SLOT, e);
						if (RECORD) normal = F3{ QF(E_NX, e), QF(E_NY, e), QF(E_NZ, e) };
					}
					// the traversal state exactly as it was saved; tdelta and the integer steps as dda_setup derives them
					ts.origin = F3{ QF(E_OX, e), QF(E_OY, e), QF(E_OZ, e) };
					ts.tminn = 0.f;
					world = F3{ ts.origin.x * 8.f, ts.origin.y * 8.f, ts.origin.z * 8.f };  // exact inverse of trace_setup's origin / 8
					const uint32_t packed = QU(E_POS, e);
					ts.a.pos = I3{ (int)(packed & 0xFFFFu), (int)(packed >> 16), (int)((flags >> 4) & 0xFFFu) };
					ts.step_axis = (int)((flags >> 2) & 3u) - 1;
					ts.a.tmax = F3{ QF(E_TX, e), QF(E_TY, e), QF(E_TZ, e) };
					const F3 step{ gsign(direction.x), gsign(direction.y), gsign(direction.z) };
					ts.a.stepi = I3{ (int)step.x, (int)step.y, (int)step.z };
					const F3 rdinv{ direction.x == 0.0f ? 0.0f : 1.f / direction.x, direction.y == 0.0f ? 0.0f : 1.f / direction.y,
						            direction.z == 0.0f ? 0.0f : 1.f / direction.z };
					ts.a.tdelta = F3{ step.x * rdinv.x, step.y * rdinv.y, step.z * rdinv.z };
					tracing = true;
				}
			}
		} else if (!pool_dry) {
			// ---- start 32 fresh slots ---------------------------------------------------------------------------------------
			if (round == kRun) {
				if (lane == 0) run = atomicAdd(&st->tile_ticket, 1u);
				run = __shfl_sync(0xFFFFFFFFu, run, 0);
				round = 0;
				if (run >= nruns) {
					pool_dry = true;
					continue;
				}
				// Which slots a warp takes when is free (results are per slot). Descending = the frame's fresh primary rays before
				// the survivors of the previous frame: the longest rays (primaries near the horizon) start first, the tail shrinks.
				if (descending) run = nruns - 1 - run;
			}
			slot = (run * kRun + round) * 32 + lane;
			round++;
			const bm_ray* mine = nullptr;  // the private survivor record this lane reads (BM_SURV_HINTS)
			uint32_t witness = 0;
			if (slot < n_slots) {
				Ray ray;
				if (slot < c) {
					const SurvivorSet in = input_set(io, cur);
					const bm_ray* p = survivor_ptr(in, slot);
					if (BM_SURV_HINTS && !RECORD && in.in_prefix) {
						ray = load_ray_private(p, witness);
						mine = p;
					} else {
						ray = load_ray(p);
					}
				} else {
					ray = generate_primary(fp, frame, start, slot - c);  // primary_rays, kernel.cu:154-223
				}
				kind = K_EXTEND;
				world = ray.origin;
				direction = ray.direction;
				payload = ray.throughput;
				normal = ray.normal;
				pixel = ray.pixel_index;
				bounces = ray.bounces;
				tracing = trace_setup(sv, ray.origin, ray.direction, normal, ts);  // extend, kernel.cu:226-238
			}
#if BM_SURV_HINTS & 4
			if (!RECORD) {
				// Drop the lines just read from L2 without a write-back: the records are dead (a set is read once, by the frame after the one
				// that wrote it). A 128-byte line holds the records of slots 2j and 2j + 1 of the sparse set. Two survivors that share a line
				// are consecutive survivors, i.e. in neighbouring lanes -- unless a warp boundary falls between them. The lane holding the
				// even slot drops the line when the odd one was read by the next lane; either lane drops it when the other half holds no
				// survivor (mask bit clear); a line split between two warps is left to the normal eviction. The shuffle of `witness` makes
				// the drop wait for the loads of BOTH lanes (the shuffle cannot issue before every lane's loads have returned).
				const unsigned long long a = (unsigned long long)mine;
				const unsigned long long up = __shfl_down_sync(0xFFFFFFFFu, a, 1);
				const uint32_t w_up = __shfl_down_sync(0xFFFFFFFFu, witness, 1);
				bool drop = false;
				if (mine) {
					const SurvivorSet in = input_set(io, cur);
					const size_t s = (size_t)(mine - in.in), other = s ^ 1;
					if (!(s & 1) && lane < 31 && up == a + sizeof(bm_ray)) drop = true;
					else drop = !((__ldg(in.in_mask + (other >> 5)) >> (other & 31)) & 1u);
				}
				// (a real use of both witnesses: the drop is predicated on their xor not being one particular value -- skipping a drop is harmless)
				if (drop)
					asm volatile(
					    "{\n\t"
					    ".reg .pred p;\n\t"
					    ".reg .b32 t;\n\t"
					    "xor.b32 t, %1, %2;\n\t"
					    "setp.ne.u32 p, t, 0x9E3779B9;\n\t"
					    "@p discard.global.L2 [%0], 128;\n\t"
					    "}" ::"l"(a & ~127ull), "r"(witness), "r"(w_up)
					    : "memory");
			}
#endif
		} else {
			break;
		}
		__syncwarp();  // every popped entry has been read before anything is pushed

		// ---- one quantum of traversal -------------------------------------------------------------------------------------
		WorkCounters wc;
		// give the batch up (suspend what is left of it) once fewer than min_share / 32 of the lanes that started tracing are
		// still at it; rays that entered the world from outside cannot be suspended (see above) and run to their end
		const int min_lanes = (__popc(__ballot_sync(0xFFFFFFFFu, tracing)) * min_share) >> 5;
		if (tracing) {
			// no suspending where it cannot regroup anything: rays that entered from outside (see above), and the last batch of a
			// warp whose queue and slot pool are both empty
			const bool pinned = ts.tminn > 0.f || (pool_dry && qn == 0);
			status = trace_run<false, true, STOCK>(sv, coarse, direction, normal, distance, fp.cam_cell, ts, pinned ? 0x3FFFFFF8 : quantum, &wc, pinned ? 0 : min_lanes, pinned ? 0 : sch.brick_lanes);
		}
		__syncwarp();  // lanes whose ray ended early wait here: they are shaded together, not interleaved with the tracing lanes

		// ---- outcomes -----------------------------------------------------------------------------------------------------
		int push = K_NONE;  // what this lane appends to the queue
		F3 push_o{ 0.f, 0.f, 0.f }, push_d{ 0.f, 0.f, 0.f }, push_c{ 0.f, 0.f, 0.f };
		if (kind != K_NONE) {
			if (status == TRACE_SUSPENDED || status == TRACE_AT_BRICK) {
				push = kind;
			} else if (kind == K_SHADOW) {
				if (status == TRACE_MISS) {  // kernel.cu:340-344
					accum_add(io.accum, pixel, payload.x, payload.y, payload.z, 0.f);
					n_unocc++;
				}
			} else {
				// shade, kernel.cu:242-325
				Ray ray;
				ray.origin = world;
				ray.direction = direction;
				ray.throughput = payload;
				ray.normal = normal;
				ray.distance = status == TRACE_HIT ? distance : kVeryFar;
				ray.identifier = 0;
				ray.bounces = bounces;
				ray.pixel_index = pixel;
				if (RECORD) store_ray(io.record + slot, ray);  // what extend leaves in the work queue (kernel.cu:235-236)
				ShadeResult s;
				if (RECORD && io.extend_only) s.has_shadow = s.survives = s.terminated = s.add_radiance = false;  // BM_FRAME_EXTEND_ONLY
				else s = shade_vertex(fp, frame, slot, ray);
				if (s.add_radiance) accum_add(io.accum, pixel, s.radiance.x, s.radiance.y, s.radiance.z, 1.f);
				else if (s.terminated) accum_add(io.accum, pixel, 0.f, 0.f, 0.f, 1.f);
				n_term += s.terminated;
				if (s.survives) {
					if (BM_SURV_HINTS && !RECORD) store_ray_private(output_rays(io, cur) + slot, ray);
					else store_ray(output_rays(io, cur) + slot, ray);
					atomicOr(output_mask(io, cur) + (slot >> 5), 1u << (slot & 31));
				}
				if (s.has_shadow) {
					n_shadow++;
					push = K_SHADOW_NEW;
					push_o = ray.origin;
					push_d = s.shadow_dir;
					push_c = s.shadow_color;
					if (RECORD) {  // kernel.cu:277-278, sparse by slot like the survivors
						store_shadow(io.shadow_out + slot, ray.origin, s.shadow_dir, s.shadow_color, pixel);
						atomicOr(io.shadow_mask + (slot >> 5), 1u << (slot & 31));
					}
				}
			}
		}
#ifdef BM_QDEBUG
		{
			const uint32_t tm = __ballot_sync(0xFFFFFFFFu, tracing), sm = __ballot_sync(0xFFFFFFFFu, tracing && status == TRACE_SUSPENDED);
			const uint32_t km = __ballot_sync(0xFFFFFFFFu, kind != K_NONE);
			if (lane == 0) {
				atomicAdd(&st->index_reads, 1ull);
				atomicAdd(&st->cell_steps, (unsigned long long)__popc(tm));
				atomicAdd(&st->bricks, (unsigned long long)__popc(sm));
				atomicAdd(&st->requests, (unsigned long long)__popc(km));
			}
		}
#endif
		const uint32_t pm = __ballot_sync(0xFFFFFFFFu, push != K_NONE);
		if (push != K_NONE) {
			const uint32_t e = qn + __popc(pm & lt_mask);
			if (push == K_SHADOW_NEW) {
				QF(E_OX, e) = push_o.x; QF(E_OY, e) = push_o.y; QF(E_OZ, e) = push_o.z;
				QF(E_DX, e) = push_d.x; QF(E_DY, e) = push_d.y; QF(E_DZ, e) = push_d.z;
				QF(E_CX, e) = push_c.x; QF(E_CY, e) = push_c.y; QF(E_CZ, e) = push_c.z;
				QU(E_FLAGS, e) = (uint32_t)K_SHADOW_NEW;
				QU(E_PIXEL, e) = pixel;
			} else {
				QF(E_OX, e) = ts.origin.x; QF(E_OY, e) = ts.origin.y; QF(E_OZ, e) = ts.origin.z;
				QF(E_DX, e) = direction.x; QF(E_DY, e) = direction.y; QF(E_DZ, e) = direction.z;
				QU(E_POS, e) = (uint32_t)ts.a.pos.x | ((uint32_t)ts.a.pos.y << 16);
				QF(E_TX, e) = ts.a.tmax.x; QF(E_TY, e) = ts.a.tmax.y; QF(E_TZ, e) = ts.a.tmax.z;
				QU(E_FLAGS, e) = (uint32_t)push | ((uint32_t)(ts.step_axis + 1) << 2) | ((uint32_t)ts.a.pos.z << 4) | ((uint32_t)bounces << 16);
				QF(E_CX, e) = payload.x; QF(E_CY, e) = payload.y; QF(E_CZ, e) = payload.z;
				QU(E_PIXEL, e) = pixel;
				if (push == K_EXTEND) {
					QU(E_SLOT, e) = slot;
					if (RECORD) { QF(E_NX, e) = normal.x; QF(E_NY, e) = normal.y; QF(E_NZ, e) = normal.z; }
				}
			}
		}
		qn += __popc(pm);
		__syncwarp();
	}

	// per-warp statistics -> a handful of atomics per warp
	const unsigned long long a = warp_sum((unsigned long long)n_shadow), b = warp_sum((unsigned long long)n_term), u = warp_sum((unsigned long long)n_unocc);
	if (lane == 0) {
		if (a) atomicAdd(&st->shadow_rays, a);
		if (b) atomicAdd(&st->terminations, b);
		if (b) atomicAdd(&st->paths_since_reset, b);
		if (u) atomicAdd(&st->unoccluded, u);
	}
}

#undef QF
#undef QU

}  // namespace bm
