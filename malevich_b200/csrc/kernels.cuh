// kernels.cuh -- the draw pipeline as hand-written sm_100a kernels.
//
//   k_clear      clear_render_target_view + clear_depth_stencil_view      (reference main.c:1191-1217)
//   k_front<VS> / k_front_clip<VS>  input assembler + vertex shader + primitive assembly (main.c:662-913), the half
//                that does not depend on the render target: one thread per input triangle, no inter-thread dependency
//                (triangles are named by order-preserving keys instead of compacted ids, see mlv_internal.cuh);
//                runs ahead of the draw-to-draw chain on a stream of its own
//   k_back<VS>   Hi-Z + binner pass 1 per triangle, setup records for the survivors
//   k_bin_scan   binner exclusive scan + compaction of non-empty bins     (main.c:937-974), multi-CTA
//                single pass with decoupled look-back
//   k_fill       binner pass 2                                            (main.c:950-962)
//   k_tile<PS>   rasterizer + Hi-Z + early-Z + pixel shader + output merger (main.c:983-1189)
//                one warp per non-empty bin: lanes over triangles for coverage (64-bit masks), a 32x32 bit
//                transpose turns them into per-pixel cover sets, lanes over pixels for depth, pixel shader
//                run once per pixel on the last fragment that passed (bit-identical to in-order shading)
//   k_resolve    tiled -> row-major, 128-bit stores                       (replaces GDI blit main.c:314)
#pragma once
#include "mlv_internal.cuh"

namespace mlv {

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// Programmatic dependent launch (sm_90+): every kernel is launched with programmaticStreamSerialization, lets the
// NEXT kernel of the stream start launching at once (its CTAs become resident as SM resources free up) and then waits
// here until the PREVIOUS kernel has completed and its writes are visible. Semantics are those of plain stream order;
// what is saved is the launch latency between the ~7 dependent kernels of every draw.
__device__ __forceinline__ void pdl_prologue() {
	asm volatile("griddepcontrol.launch_dependents;");
	asm volatile("griddepcontrol.wait;" ::: "memory");
}

// Device-side timeline (mlv_timeline_*): a launch that carries a slot stamps %globaltimer into it -- [0] the first CTA
// resident, [1] the first CTA past griddepcontrol.wait (the previous kernel of its stream has completed), [2] the last CTA
// done, stored inverted so that all three are atomicMin over words preset to ~0. Unlike event brackets this works inside a
// recorded command list and across streams: it shows the frame as it really overlaps.
__device__ __forceinline__ unsigned long long globaltimer_ns() {
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}
struct TimelineScope {
	unsigned long long *slot;
	__device__ __forceinline__ explicit TimelineScope(unsigned long long *s) : slot(s) {
		if(slot && threadIdx.x == 0) atomicMin(slot, globaltimer_ns());
	}
	__device__ __forceinline__ void started() const {
		if(slot && threadIdx.x == 0) atomicMin(slot + 1, globaltimer_ns());
	}
	__device__ __forceinline__ ~TimelineScope() {
		if(slot && threadIdx.x == 0) atomicMin(slot + 2, ~globaltimer_ns());
	}
};
#define MLV_KERNEL_PROLOGUE(slot_ptr) \
	TimelineScope timeline_scope_(slot_ptr); \
	pdl_prologue(); \
	timeline_scope_.started()

// =================================================================================================
// clear
// =================================================================================================
// mode bit 0: colour, bit 1: depth (+ tile minima := 0, main.c:1212-1214)
__global__ void __launch_bounds__(256) k_clear(uint4 *__restrict__ fb, float *__restrict__ tile_min, uint32_t bin_begin, uint32_t bin_end, uint32_t color, float depth, int mode) {
	pdl_prologue();
	const uint32_t i = bin_begin * 32u + blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= bin_end * 32u) return;
	const uint32_t d = __float_as_uint(depth);
	if(mode == 3) {
		fb[i] = make_uint4(color, color, d, d);
	} else if(mode == 1) {
		uint2 *p = reinterpret_cast<uint2 *>(fb + i);
		p[0] = make_uint2(color, color);
	} else {
		uint2 *p = reinterpret_cast<uint2 *>(fb + i);
		p[1] = make_uint2(d, d);
	}
	if((mode & 2) && (i & 31u) == 0) tile_min[i >> 5] = __uint_as_float(MLV_TILE_MIN_CLEARED); // == 0.0 (main.c:1212-1214), tagged "not refreshed yet"
}

// =================================================================================================
// geometry: IA + VS + primitive assembly
// =================================================================================================

struct TriSetup {
	int e[9];     // a0,b0,c0,a1,b1,c1,a2,b2,c2
	float rw[3];
	float ooa;
	float max_depth;
	float4 p[3];  // screen-space positions written over register 0 (main.c:881-883)
	int minx, miny, maxx, maxy;
	int sx[3], sy[3]; // snapped 28.4 vertex coordinates
	int signed_area;
	bool nowrap;  // every edge-function evaluation on this render target is free of i32 wrap-around
};

// set_edge_function (main.c:563-575), wrapping i32 arithmetic
__device__ __forceinline__ void set_edge(int *e, int signed_area, int x0, int y0, int x1, int y1) {
	uint32_t a = (uint32_t)y0 - (uint32_t)y1;
	uint32_t b = (uint32_t)x1 - (uint32_t)x0;
	if(signed_area < 0) {
		a = 0u - a;
		b = 0u - b;
	}
	const uint32_t c = (0u - a) * (uint32_t)x0 - b * (uint32_t)y0;
	e[0] = (int)a;
	e[1] = (int)b;
	e[2] = (int)c;
}

// main.c:843-848: floor(x * (1 << 4) + 0.5) -- float multiply, double add, double floor, C conversion to i32
// The C conversion is x86 cvttsd2si: NaN and everything outside [-2^31, 2^31) give the "integer indefinite" 0x80000000
// (CUDA's conversion would saturate, and map NaN to 0).
__device__ __forceinline__ int snap(float v) {
	const double d = floor((double)(v * 16.0f) + 0.5);
	if(!(d >= -2147483648.0 && d < 2147483648.0)) return (int)0x80000000;
	return __double2int_rz(d);
}

// Per-vertex part of primitive assembly (main.c:803-848): 1/w, projection, viewport transform, snap to 28.4.
// A pure function of the clip-space position, so it can be evaluated once per vertex (k_vertex) or once per
// triangle corner (as the reference does) with identical results.
struct ProjVertex {
	float4 s; // screen-space position (x, y, z, w ~ 1), what the reference stores back into register 0 (main.c:881-883)
	float rw; // a_reciprocal_ws
	int sx, sy;
};
__device__ __forceinline__ ProjVertex project_vertex(float4 p, const GeomParams &P) {
	ProjVertex o;
	// a_reciprocal_ws[i] = 1.0 / w  (double divide rounded to f32 == correctly rounded f32 divide)
	const float rw = 1.0f / p.w;
	o.rw = rw;
	p.x *= rw;
	p.y *= rw;
	p.z *= rw;
	p.w *= rw;
	// m4x4f32_mul_v4f32(&screen_from_ndc, v): serial dot products with the literal zero entries
	o.s.x = P.vp_m00 * p.x + 0.0f * p.y + 0.0f * p.z + P.vp_m03 * p.w;
	o.s.y = 0.0f * p.x + P.vp_m11 * p.y + 0.0f * p.z + P.vp_m13 * p.w;
	o.s.z = 0.0f * p.x + 0.0f * p.y + P.vp_m22 * p.z + P.vp_m23 * p.w;
	o.s.w = 0.0f * p.x + 0.0f * p.y + 0.0f * p.z + 1.0f * p.w;
	o.sx = snap(o.s.x);
	o.sy = snap(o.s.y);
	return o;
}

// Per-triangle part: signed area, face culling, max depth and bounds (main.c:851-857, 871, 888-898). Returns false
// when the triangle is back-face culled (signed_area > 0, main.c:856). Everything binning and Hi-Z need; the edge
// functions follow in setup_edges().
__device__ __forceinline__ bool setup_from_projected(const ProjVertex &v0, const ProjVertex &v1, const ProjVertex &v2, const GeomParams &P, TriSetup &S) {
	S.rw[0] = v0.rw, S.rw[1] = v1.rw, S.rw[2] = v2.rw;
	S.p[0] = v0.s, S.p[1] = v1.s, S.p[2] = v2.s;
	const int x0 = v0.sx, x1 = v1.sx, x2 = v2.sx, y0 = v0.sy, y1 = v1.sy, y2 = v2.sy;
	S.sx[0] = x0, S.sx[1] = x1, S.sx[2] = x2;
	S.sy[0] = y0, S.sy[1] = y1, S.sy[2] = y2;
	S.signed_area = (int)(((uint32_t)x1 - (uint32_t)x0) * ((uint32_t)y2 - (uint32_t)y0) - ((uint32_t)x2 - (uint32_t)x0) * ((uint32_t)y1 - (uint32_t)y0));
	if(S.signed_area > 0) return false;
	S.max_depth = ref_max_macro(S.p[0].z, ref_max_macro(S.p[1].z, S.p[2].z)); // MAX3 math.h:32
	const int mnx = min(x0, min(x1, x2)) >> 4, mny = min(y0, min(y1, y2)) >> 4;
	const int mxx = max(x0, max(x1, x2)) >> 4, mxy = max(y0, max(y1, y2)) >> 4;
	S.minx = min(max(mnx, 0), P.vp_w - 1);
	S.miny = min(max(mny, 0), P.vp_h - 1);
	S.maxx = min(mxx + 1, P.vp_w - 1);
	S.maxy = min(mxy + 1, P.vp_h - 1);
	return true;
}

__device__ __forceinline__ bool setup_project(const float4 &c0, const float4 &c1, const float4 &c2, const GeomParams &P, TriSetup &S) {
	return setup_from_projected(project_vertex(c0, P), project_vertex(c1, P), project_vertex(c2, P), P, S);
}

// Edge functions, 1/area (main.c:858-866) and the no-wrap flag.
__device__ __forceinline__ void setup_edges(const GeomParams &P, TriSetup &S) {
	const int x0 = S.sx[0], x1 = S.sx[1], x2 = S.sx[2], y0 = S.sy[0], y1 = S.sy[1], y2 = S.sy[2];
	set_edge(S.e + 6, S.signed_area, x0, y0, x1, y1);
	set_edge(S.e + 0, S.signed_area, x1, y1, x2, y2);
	set_edge(S.e + 3, S.signed_area, x2, y2, x0, y0);
	float area_f = (float)(S.signed_area >> 8);
	if(area_f == 0.0f) area_f = 1.0f;
	S.ooa = fabsf(1.0f / area_f);
	// No-wrap proof: with |x_i|,|y_i| <= M and sample coordinates in [0, 16W] x [0, 16H], every term of
	// a*(16x) + b*(16y) + c (c = -a*x_i - b*y_i) is bounded by (|a|+|b|) * max(M, 16W, 16H); if twice that stays
	// below 2^31 no intermediate wraps and the integer edge functions equal the exact ones, so coverage is
	// confined to the (closed) triangle and hence to [min_bounds, max_bounds].
	// Evaluated in fp32 with a 1 % safety margin (both factors are < 2^26, compared far from their rounding error).
	const int mi = max(max(max(abs(x0), abs(x1)), abs(x2)), max(max(abs(y0), abs(y1)), abs(y2)));
	const float m = (float)max(mi, max(P.vp_w, P.vp_h) * 16 + 16);
	float ab = 0.0f;
#pragma unroll
	for(int k = 0; k < 3; ++k) ab = fmaxf(ab, fabsf((float)S.e[k * 3]) + fabsf((float)S.e[k * 3 + 1]));
	S.nowrap = (mi < (1 << 24)) && (2.0f * ab * m < 2126008811.0f); // 0.99 * 2^31
}

__device__ __forceinline__ bool setup_triangle(const float4 &c0, const float4 &c1, const float4 &c2, const GeomParams &P, TriSetup &S) {
	if(!setup_project(c0, c1, c2, P, S)) return false;
	setup_edges(P, S);
	return true;
}

__device__ __forceinline__ VsOut lerp_vertex(const VsOut &a, const VsOut &b, float t) {
	const float g = 1.0f - t;
	VsOut o;
	o.r0 = make_float4(a.r0.x * g + b.r0.x * t, a.r0.y * g + b.r0.y * t, a.r0.z * g + b.r0.z * t, a.r0.w * g + b.r0.w * t);
	o.r1 = make_float4(a.r1.x * g + b.r1.x * t, a.r1.y * g + b.r1.y * t, a.r1.z * g + b.r1.z * t, a.r1.w * g + b.r1.w * t);
	o.r2x = a.r2x * g + b.r2x * t;
	return o;
}

// Polygons of the clipper live in SHARED memory, one column per thread: element (vertex v, component c) of thread `tid`
// is at p[(v * 9 + c) * MLV_CLIP_THREADS], p already offset by tid -- conflict-free, and a fixed ~30-cycle access.
// (As per-thread local arrays the two 16-vertex polygons of every thread overflowed L1 and each of the hundreds of
// dependent accesses of the clipping loop became an L2 round trip: k_geom_clip was a 20 us latency chain.)
#define MLV_CLIP_THREADS 64
#define MLV_CLIP_MAXV 10 /* a triangle clipped by six planes has at most 9 vertices (MAX_NUM_CLIP_VERTICES is 16, main.c:38) */
__device__ __forceinline__ VsOut poly_load(const float *p, int v) {
	const float *q = p + v * 9 * MLV_CLIP_THREADS;
	VsOut o;
	o.r0 = make_float4(q[0], q[MLV_CLIP_THREADS], q[2 * MLV_CLIP_THREADS], q[3 * MLV_CLIP_THREADS]);
	o.r1 = make_float4(q[4 * MLV_CLIP_THREADS], q[5 * MLV_CLIP_THREADS], q[6 * MLV_CLIP_THREADS], q[7 * MLV_CLIP_THREADS]);
	o.r2x = q[8 * MLV_CLIP_THREADS];
	return o;
}
__device__ __forceinline__ float4 poly_load_pos(const float *p, int v) {
	const float *q = p + v * 9 * MLV_CLIP_THREADS;
	return make_float4(q[0], q[MLV_CLIP_THREADS], q[2 * MLV_CLIP_THREADS], q[3 * MLV_CLIP_THREADS]);
}
__device__ __forceinline__ void poly_store(float *p, int v, const VsOut &o) {
	float *q = p + v * 9 * MLV_CLIP_THREADS;
	q[0] = o.r0.x, q[MLV_CLIP_THREADS] = o.r0.y, q[2 * MLV_CLIP_THREADS] = o.r0.z, q[3 * MLV_CLIP_THREADS] = o.r0.w;
	q[4 * MLV_CLIP_THREADS] = o.r1.x, q[5 * MLV_CLIP_THREADS] = o.r1.y, q[6 * MLV_CLIP_THREADS] = o.r1.z, q[7 * MLV_CLIP_THREADS] = o.r1.w;
	q[8 * MLV_CLIP_THREADS] = o.r2x;
}

// clip_by_plane (main.c:609-647), plane_d = 0: src -> dst. Returns the new vertex count, or -1 when the whole polygon
// is inside (the pass would reproduce it vertex for vertex; the caller keeps src).
__device__ __forceinline__ int clip_by_plane(const float *src, float *dst, int n, float4 pn) {
	{
		bool all_in = n > 0;
		for(int i = 0; i < n; ++i) all_in = all_in && (dot4_serial(pn, poly_load_pos(src, i)) > -0.0f);
		if(all_in) return -1;
	}
	int nout = 0;
	float cur = dot4_serial(pn, poly_load_pos(src, 0));
	bool cur_in = cur > -0.0f;
	for(int i = 0; i < n; ++i) {
		const int next = (i + 1 == n) ? 0 : i + 1;
		const VsOut vi = poly_load(src, i);
		if(cur_in && nout < MLV_CLIP_MAXV) poly_store(dst, nout++, vi);
		const float nd = dot4_serial(pn, poly_load_pos(src, next));
		const bool nin = nd > -0.0f;
		if(cur_in != nin && nout < MLV_CLIP_MAXV) {
			const float t = (0.0f + cur) / (cur - nd);
			poly_store(dst, nout++, lerp_vertex(vi, poly_load(src, next), t));
		}
		cur = nd;
		cur_in = nin;
	}
	return nout;
}

__device__ __forceinline__ uint2 pack_bounds(const TriSetup &S) {
	return make_uint2((uint32_t)(S.minx & 0xffff) | ((uint32_t)(S.miny & 0x7fff) << 16) | (S.nowrap ? MLV_NOWRAP_BIT : 0u),
	                  (uint32_t)(S.maxx & 0xffff) | ((uint32_t)(S.maxy & 0xffff) << 16));
}

#define MLV_HUGE_TILES 2048 /* tile rectangles larger than this are expanded by the whole grid, not by one warp */

// Tile rectangle exactly as the binner derives it from the pixel bounds (main.c:927-928), C division.
struct TileRect {
	int tx0, ty0, tx1, ty1;
	__device__ __forceinline__ int w() const { return max(tx1 - tx0 + 1, 0); }
	__device__ __forceinline__ int h() const { return max(ty1 - ty0 + 1, 0); }
};
__device__ __forceinline__ TileRect tile_rect(int minx, int miny, int maxx, int maxy, int wt, int ht) {
	TileRect r;
	r.tx0 = max(minx / 8, 0);
	r.ty0 = max(miny / 8, 0);
	r.tx1 = min(maxx / 8, wt - 1);
	r.ty1 = min(maxy / 8, ht - 1);
	return r;
}

// Hi-Z (main.c:1003-1010) decides per (triangle, tile) pair from values that are final before the draw starts
// (tile minima come from previous draws only, N3), so the test can run at binning time: a rejected pair is
// counted for Stats and marks its bin as touched, but is never stored, sorted or rasterized. keep_all (debug
// capture) disables this so the lists hold every pair like the reference's.
__device__ __forceinline__ bool hiz_rejects(float max_depth, const float *__restrict__ tile_min, uint32_t bin, bool keep_all) {
	return !keep_all && (max_depth < __ldg(tile_min + bin));
}

// Pass 1 of the binner (main.c:924-936) for one triangle, fused into the back half of geometry.
//   live  = at least one pair survives Hi-Z (big rectangles: decided by the warp-cooperative expansion; huge ones: owns a tile)
//   first = bins this thread touched FIRST in this draw (Stats: active_bin_count, main.c:1245)
//   wake  = a bin without surviving pairs still needs its k_tail visit (write_tile refreshes the tile minimum of every
//           non-empty bin, main.c:589-603; only a bin whose minimum still carries the clear tag would change)
struct BinTally {
	bool live, big, huge, wake;
	uint32_t survivors;
};

// "Touched in this draw" (Stats: active_bin_count counts the bins that received a pair, Hi-Z-rejected or not, main.c:1245)
// is one BYTE per bin in the DRAW CONTEXT's touch map: it depends on nothing an earlier draw leaves behind, so the FRONT
// half raises the bytes (for rectangles of at most 8 tiles; larger ones are expanded by the back half, which raises theirs)
// and a triangle hidden by Hi-Z costs the back half nothing for it. Plain idempotent byte stores: no read-modify-write,
// the lanes of a warp that store to the same bin merge in the load/store unit (one bit per bin with atomicOr serialised
// on the 32 bins sharing a word: 100 us per draw). The draw's k_tile counts the bytes and clears the map.
// `seen` is the value loaded from touch[bin] (L2: a stale 0 only repeats the store).
__device__ __forceinline__ void touch_bin(uint8_t *touch, uint32_t bin, uint32_t seen) {
	if(!seen) touch[bin] = 1;
}
template <int N>
__device__ __forceinline__ void touch_small_rect(uint8_t *touch_bits, const Partition &part, int wt, const TileRect &tr, int cnt) {
	uint32_t bins[N], word[N];
	bool ok[N];
	int tx = tr.tx0, ty = tr.ty0;
#pragma unroll
	for(int k = 0; k < N; ++k) {
		ok[k] = k < cnt && part.owns_row(ty);
		bins[k] = (uint32_t)(ty * wt + tx);
		if(++tx > tr.tx1) {
			tx = tr.tx0;
			++ty;
		}
	}
#pragma unroll
	for(int k = 0; k < N; ++k) word[k] = ok[k] ? (uint32_t)__ldcg(touch_bits + bins[k]) : 1u;
#pragma unroll
	for(int k = 0; k < N; ++k)
		if(ok[k]) touch_bin(touch_bits, bins[k], word[k]);
}
// (front half) the touch bits of a triangle whose tile rectangle holds 1..8 tiles
__device__ __forceinline__ void touch_rect(uint8_t *touch_bits, const Partition &part, int wt, const TileRect &tr, int cnt) {
	if(cnt <= 0 || cnt > 8) return;
	if(cnt <= 4) touch_small_rect<4>(touch_bits, part, wt, tr, cnt);
	else touch_small_rect<8>(touch_bits, part, wt, tr, cnt);
}

// One (triangle, tile) pair of the back half. tm is the value loaded from tile_min[bin].
template <typename Params>
__device__ __forceinline__ void count_pair(const Params &P, uint32_t bin, float tm, float max_depth, BinTally &r) {
	const bool rejected = !P.keep_all && (max_depth < tm); // Hi-Z (main.c:1006)
	if(!rejected) {
		atomicAdd(P.bin_count + bin, 1u);
		r.live = true;
		++r.survivors;
	} else if(__float_as_uint(tm) == MLV_TILE_MIN_CLEARED) {
		r.wake = true; // (see k_bin_scan)
	}
}

// Rectangles of at most N tiles: walk them with fully unrolled, predicated steps so that the tile-minimum and touch
// loads of all tiles are independent and in flight together (they were a chain of dependent L2 round trips), then
// issue the atomics.
template <int N>
__device__ __forceinline__ void count_small_rect(const GeomParams &P, const TileRect &tr, int cnt, float max_depth, BinTally &r) {
	uint32_t bins[N];
	bool ok[N];
	{
		int tx = tr.tx0, ty = tr.ty0;
#pragma unroll
		for(int k = 0; k < N; ++k) {
			ok[k] = k < cnt && P.part.owns_row(ty);
			bins[k] = (uint32_t)(ty * P.wt + tx);
			if(++tx > tr.tx1) {
				tx = tr.tx0;
				++ty;
			}
		}
	}
	float tm[N];
#pragma unroll
	for(int k = 0; k < N; ++k) tm[k] = ok[k] ? __ldg(P.tile_min + bins[k]) : 0.0f;
#pragma unroll
	for(int k = 0; k < N; ++k)
		if(ok[k]) count_pair(P, bins[k], tm[k], max_depth, r);
}

__device__ __forceinline__ BinTally count_bins(const GeomParams &P, const TileRect &tr, float max_depth) {
	BinTally r = { false, false, false, false, 0u };
	const int cnt = tr.w() * tr.h();
	if(cnt <= 0) return r;
	if(cnt > 8) { // counted by k_bin_big: a warp per rectangle, the whole grid for a huge one
		r.live = P.part.owned_rows(tr.ty0, tr.ty1) > 0;
		r.big = r.live && cnt <= MLV_HUGE_TILES;
		r.huge = r.live && cnt > MLV_HUGE_TILES;
		return r;
	}
	// dense meshes of pixel-sized triangles (BASELINE config 5: 2.2 tiles per triangle) almost never need more than
	// four steps; the eight-step walk costs a third of the kernel's instructions when every lane pays for it
	if(cnt <= 4) count_small_rect<4>(P, tr, cnt, max_depth, r);
	else count_small_rect<8>(P, tr, cnt, max_depth, r);
	return r;
}

// The 144-byte record of one assembled triangle (layouts in mlv_internal.cuh).
struct TriRecord {
	uint4 cov[MLV_TRI_COV_U4];
	float4 shade[MLV_TRI_SHADE_U4];
};

__device__ __forceinline__ void make_record(TriRecord &R, const TriSetup &S, uint2 pb, const float4 &r1a, const float4 &r1b, const float4 &r1c, float r2a, float r2b, float r2c) {
	R.cov[0] = make_uint4(S.e[0], S.e[1], S.e[2], S.e[3]);
	R.cov[1] = make_uint4(S.e[4], S.e[5], S.e[6], S.e[7]);
	R.cov[2] = make_uint4(S.e[8], __float_as_uint(S.max_depth), pb.x, pb.y);
	R.shade[0] = make_float4(S.ooa, S.p[0].z, S.p[1].z, S.p[2].z);
	R.shade[1] = make_float4(S.rw[0], S.rw[1], S.rw[2], r2a);
	R.shade[2] = r1a;
	R.shade[3] = r1b;
	R.shade[4] = r1c;
	R.shade[5] = make_float4(r2b, r2c, 0.0f, 0.0f);
}

// reference-layout copies for mlv_debug_read_triangles (debug capture only)
__device__ __forceinline__ void emit_debug(const GeomParams &P, uint32_t slot, uint32_t key, const TriSetup &S, const float4 &r1a, const float4 &r1b, const float4 &r1c, float r2a,
                                           float r2b, float r2c) {
	mlv_ref_triangle t;
	t.p_attributes = 0;
	t.min_bounds[0] = S.minx;
	t.min_bounds[1] = S.miny;
	t.max_bounds[0] = S.maxx;
	t.max_bounds[1] = S.maxy;
#pragma unroll
	for(int k = 0; k < 3; ++k) {
		t.edges[k][0] = S.e[k * 3 + 0];
		t.edges[k][1] = S.e[k * 3 + 1];
		t.edges[k][2] = S.e[k * 3 + 2];
	}
	t.reciprocal_ws[0] = S.rw[0];
	t.reciprocal_ws[1] = S.rw[1];
	t.reciprocal_ws[2] = S.rw[2];
	t.one_over_area = S.ooa;
	t.max_depth = S.max_depth;
	P.dbg.tris[slot] = t;
	float4 *a = reinterpret_cast<float4 *>(P.dbg.attrs + (size_t)slot * 36);
	a[0] = S.p[0];
	a[1] = r1a;
	a[2] = make_float4(r2a, 0.0f, 0.0f, 0.0f);
	a[3] = S.p[1];
	a[4] = r1b;
	a[5] = make_float4(r2b, 0.0f, 0.0f, 0.0f);
	a[6] = S.p[2];
	a[7] = r1c;
	a[8] = make_float4(r2c, 0.0f, 0.0f, 0.0f);
	P.dbg.slot_key[slot] = key;
}

// Clipper (main.c:649-660): returns the vertex count of the clipped polygon left in poly[].
__device__ __forceinline__ int clip_polygon(const GeomParams &P, const VsOut &v0, const VsOut &v1, const VsOut &v2, float *&cur, float *&other) {
	poly_store(cur, 0, v0);
	poly_store(cur, 1, v1);
	poly_store(cur, 2, v2);
	const float k = P.clip_k;
	// A plane pass reproduces its input vertex for vertex when the whole polygon is strictly inside (see
	// clip_by_plane), and every polygon vertex is a convex combination of v0,v1,v2. So a plane that has the three
	// ORIGINAL vertices inside by a margin that dwarfs the rounding error of the interpolated vertices
	// (<= ~30 ulp of the largest coordinate, margin = 1e-3 of it) cannot clip anything and is skipped without
	// being evaluated; planes closer than that go through the literal pass.
	const float big = fmaxf(fmaxf(fabsf(v0.r0.x) + fabsf(v0.r0.y) + fabsf(v0.r0.z) + fabsf(v0.r0.w), fabsf(v1.r0.x) + fabsf(v1.r0.y) + fabsf(v1.r0.z) + fabsf(v1.r0.w)),
	                        fabsf(v2.r0.x) + fabsf(v2.r0.y) + fabsf(v2.r0.z) + fabsf(v2.r0.w));
	const float margin = 1e-3f * big;
	int n = 3;
	const float4 planes[6] = { make_float4(k, 0.0f, 0.0f, k),  make_float4(-k, 0.0f, 0.0f, k), make_float4(0.0f, k, 0.0f, k),
		                       make_float4(0.0f, -k, 0.0f, k), make_float4(0.0f, 0.0f, k, k),  make_float4(0.0f, 0.0f, -k, k) };
#pragma unroll
	for(int pl = 0; pl < 6; ++pl) {
		const float d0 = dot4_serial(planes[pl], v0.r0), d1 = dot4_serial(planes[pl], v1.r0), d2 = dot4_serial(planes[pl], v2.r0);
		if(fminf(fminf(d0, d1), d2) > margin) continue; // safely inside (false for NaN: literal pass)
		const int m = clip_by_plane(cur, other, n, planes[pl]);
		if(m >= 0) { // ping-pong
			n = m;
			float *sw = cur;
			cur = other;
			other = sw;
		}
	}
	return n;
}

// Per-draw Stats contributions (main.c:1228-1246): one atomic per warp and counter.
// Stats go to one of MLV_STAT_STRIPES 64-bit accumulators, 128 bytes apart (L2 atomics serialise per address: with
// one shared counter the ~40 k warps of a 1.25 M-triangle draw queued up behind each other). k_tile folds them.
// High word: assembled triangles, low word: (triangle, tile) pairs.
__device__ __forceinline__ void tally_stats(unsigned long long *stripes, uint32_t emitted, uint32_t pairs) {
#pragma unroll
	for(int d = 16; d > 0; d >>= 1) {
		emitted += __shfl_xor_sync(0xffffffffu, emitted, d);
		pairs += __shfl_xor_sync(0xffffffffu, pairs, d);
	}
	if(lane_id() == 0 && (emitted | pairs)) {
		const uint32_t stripe = (blockIdx.x * 8u + (threadIdx.x >> 5)) % MLV_STAT_STRIPES;
		atomicAdd(stripes + stripe * 16u, ((unsigned long long)emitted << 32) | pairs);
	}
}

// Work counter (mlv_work_counters.records_written): second word of the warp's stripe; warp-uniform count, nothing to do
// for a draw that wrote no records.
__device__ __forceinline__ void tally_records(unsigned long long *stripes, uint32_t records) {
	if(lane_id() == 0 && records) {
		const uint32_t stripe = (blockIdx.x * 8u + (threadIdx.x >> 5)) % MLV_STAT_STRIPES;
		atomicAdd(stripes + stripe * 16u + 1u, (unsigned long long)records);
	}
}

#define MLV_GEOM_THREADS 256
#define MLV_TILE_THREADS 256

// ---- sort-first chunk culling (multi-GPU only) -------------------------------------------
// A chunk = the MLV_GEOM_THREADS input triangles of one k_geom CTA. k_chunk_bounds computes the object-space AABB of
// each chunk once per (vertex buffer, index buffer) pair; it is cached with the index buffer until either buffer is
// updated. Per draw, each k_geom CTA projects the 8 corners of its chunk's box and skips the whole chunk -- no index
// or vertex fetch -- when no tile row in the box's conservative screen-y range belongs to this rank. With contiguous
// bands per rank a rank then fetches and shades ~1/N of the geometry instead of all of it.
// Only for the vertex shaders whose SV_POSITION is clip_from_world * POSITION.xyz (basic_vs, vertex_lighting_vs).
template <bool INDEXED>
__global__ void __launch_bounds__(MLV_GEOM_THREADS) k_chunk_bounds(const IndexStream ix, const float4 *__restrict__ vb, uint32_t tri_count, float4 *__restrict__ chunk_bounds) {
	pdl_prologue();
	__shared__ float s_min[3][MLV_GEOM_THREADS / 32], s_max[3][MLV_GEOM_THREADS / 32];
	__shared__ int s_bad[MLV_GEOM_THREADS / 32];
	const uint32_t t = blockIdx.x * MLV_GEOM_THREADS + threadIdx.x;
	float mn[3] = { INFINITY, INFINITY, INFINITY }, mx[3] = { -INFINITY, -INFINITY, -INFINITY };
	int bad = 0;
	if(t < tri_count) {
#pragma unroll
		for(int c = 0; c < 3; ++c) {
			const uint32_t vi = INDEXED ? ix.fetch(3u * t + c) : 3u * t + c;
			const float4 p = __ldg(vb + 2 * (size_t)vi);
			const float q[3] = { p.x, p.y, p.z };
#pragma unroll
			for(int k = 0; k < 3; ++k) {
				if(!(fabsf(q[k]) <= 3.0e38f)) bad = 1; // NaN / Inf: never cull this chunk
				mn[k] = fminf(mn[k], q[k]);
				mx[k] = fmaxf(mx[k], q[k]);
			}
		}
	}
#pragma unroll
	for(int d = 16; d > 0; d >>= 1) {
#pragma unroll
		for(int k = 0; k < 3; ++k) {
			mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], d));
			mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], d));
		}
		bad |= __shfl_xor_sync(0xffffffffu, bad, d);
	}
	const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
	if(lane == 0) {
		for(int k = 0; k < 3; ++k) s_min[k][warp] = mn[k], s_max[k][warp] = mx[k];
		s_bad[warp] = bad;
	}
	__syncthreads();
	if(threadIdx.x == 0) {
		for(int w = 1; w < MLV_GEOM_THREADS / 32; ++w) {
			for(int k = 0; k < 3; ++k) mn[k] = fminf(mn[k], s_min[k][w]), mx[k] = fmaxf(mx[k], s_max[k][w]);
			bad |= s_bad[w];
		}
		chunk_bounds[2 * (size_t)blockIdx.x] = make_float4(mn[0], mn[1], mn[2], bad ? 1.0f : 0.0f);
		chunk_bounds[2 * (size_t)blockIdx.x + 1] = make_float4(mx[0], mx[1], mx[2], 0.0f);
	}
}

// The vertex-shader constant buffer of the draw (PerFrameCB, main.c:169-173) in shared memory: from the kernel parameters
// (immediate draws) or from the device copy a recorded command list keeps per draw (so that the constants of a replay can
// be replaced without touching the recording, mlv_command_list_set_constants).
__device__ __forceinline__ const float *stage_constants(const GeomParams &P, float *s_cb) {
	if(threadIdx.x < 48u) s_cb[threadIdx.x] = P.cbp ? __ldg(P.cbp + threadIdx.x) : P.cb[threadIdx.x];
	__syncthreads();
	return s_cb;
}

// true when the chunk certainly bins nothing on this rank
__device__ __forceinline__ bool chunk_is_foreign(const GeomParams &P, const float *cb, uint32_t chunk) {
	const float4 lo = __ldg(P.chunk_bounds + 2 * (size_t)chunk), hi = __ldg(P.chunk_bounds + 2 * (size_t)chunk + 1);
	if(lo.w != 0.0f) return false; // the chunk holds a NaN / Inf position
	float ymin = INFINITY, ymax = -INFINITY;
#pragma unroll
	for(int k = 0; k < 8; ++k) {
		const float4 corner = make_float4((k & 1) ? hi.x : lo.x, (k & 2) ? hi.y : lo.y, (k & 4) ? hi.z : lo.z, 1.0f);
		const float4 cs = mul_m4_v4_pairwise(cb, corner);
		// behind or near the eye plane the projection of the box is unbounded: keep the chunk
		if(!(cs.w > 1e-6f && fabsf(cs.y) <= 3.0e38f)) return false;
		const float ys = P.vp_m11 * (cs.y / cs.w) + P.vp_m13; // screen y of the corner (the exact path adds rounding of a few ulp)
		ymin = fminf(ymin, ys);
		ymax = fmaxf(ymax, ys);
	}
	// conservative tile-row range: 2 pixels of slack for rounding + the +1 of max_bounds (main.c:897-898), clamped like
	// min_bounds/max_bounds are (main.c:892-898)
	const float h = (float)P.vp_h;
	const int py0 = (int)floorf(fminf(fmaxf(ymin - 2.0f, 0.0f), h - 1.0f));
	const int py1 = (int)floorf(fminf(fmaxf(ymax + 3.0f, 0.0f), h - 1.0f));
	const int s0 = (py0 >> 3) / P.part.stripe_h, s1 = (py1 >> 3) / P.part.stripe_h;
	if(s1 - s0 + 1 >= P.part.num_ranks) return false;
	for(int st = s0; st <= s1; ++st)
		if(st % P.part.num_ranks == P.part.rank) return false;
	return true;
}

// Post-transform vertex cache (the reference's TODO at main.c:672; it re-shades every index, vertex_count =
// index_count main.c:673). For indexed meshes that reuse vertices, k_vertex evaluates the position part of the vertex
// shader and the per-vertex part of primitive assembly once per UNIQUE vertex; k_geom<VCACHE> then gathers 32 bytes
// per corner from an L2-resident table instead of fetching, transforming, dividing and snapping it again. Both are
// pure functions of the vertex, so the values are the ones the reference computes per corner.
// The per-vertex half of the cull / trivial-reject / trivial-accept tests (main.c:759-781), one bit per comparison so that
// the per-triangle tests become three bitwise operations. NaN makes every comparison false, as in the reference:
// bits 0-5  "outside":  x < -w, x > w, y < -w, y > w, z < 0, z > w   (rejected: some plane has all three vertices outside)
// bits 6-11 "inside":   x >= -w, x <= w, y >= -w, y <= w, z >= 0, z <= w   (inside: every plane has all three inside)
// bit 12    w == 0 (degenerate, main.c:759)
#define MLV_CODE_OUT 0x3fu
#define MLV_CODE_IN 0xfc0u
#define MLV_CODE_W0 0x1000u
__device__ __forceinline__ uint32_t clip_code(const float4 p) {
	return (p.x < -p.w ? 1u : 0u) | (p.x > p.w ? 2u : 0u) | (p.y < -p.w ? 4u : 0u) | (p.y > p.w ? 8u : 0u) | (p.z < 0.0f ? 16u : 0u) | (p.z > p.w ? 32u : 0u) |
	       (p.x >= -p.w ? 64u : 0u) | (p.x <= p.w ? 128u : 0u) | (p.y >= -p.w ? 256u : 0u) | (p.y <= p.w ? 512u : 0u) | (p.z >= 0.0f ? 1024u : 0u) | (p.z <= p.w ? 2048u : 0u) |
	       (p.w == 0.0f ? MLV_CODE_W0 : 0u);
}

template <int VS>
__global__ void __launch_bounds__(256) k_vertex(const __grid_constant__ GeomParams P, uint32_t vertex_count) {
	MLV_KERNEL_PROLOGUE(P.timeline);
	__shared__ float s_cb[48];
	const float *cb = stage_constants(P, s_cb);
	const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
	if(v >= vertex_count) return;
	const float4 pos = vs_position<VS>(__ldg(P.vb + 2 * (size_t)v), cb);
	const ProjVertex pv = project_vertex(pos, P);
	P.vcache[2 * (size_t)v] = pos;
	P.vcache[2 * (size_t)v + 1] = make_float4(__int_as_float(pv.sx), __int_as_float(pv.sy), pv.s.z, __uint_as_float(clip_code(pos)));
}

template <int VS, bool INDEXED, bool DEBUG, bool VCACHE>
__device__ __forceinline__ void fetch_indices(const GeomParams &P, uint32_t t, uint32_t &vi0, uint32_t &vi1, uint32_t &vi2) {
	if(INDEXED) {
		P.ix.fetch3(t, vi0, vi1, vi2);
	} else {
		vi0 = 3u * t;
		vi1 = vi0 + 1u;
		vi2 = vi0 + 2u;
	}
}

// =================================================================================================
// geometry, FRONT half: everything of input assembly + vertex shader + primitive assembly that does not depend on the
// render target's state (main.c:662-898 minus the Hi-Z test): index and vertex fetch, position part of the vertex shader,
// cull / trivial reject / clip test, projection, snap, signed area, face cull, bounds, Stats. Output per input triangle:
// 16 bytes of bounds {min, max, max_depth, key} (MLV_BOUNDS_EMPTY when it bins nothing on this rank) and a queue of the
// triangles that need the clipper. It reads nothing a previous draw writes, so the front halves of the draws of a frame run
// ahead on a stream of their own, underneath the tile kernels of earlier draws; only the BACK half (Hi-Z, binning, records
// for the survivors) sits on the draw-to-draw dependency chain.
// =================================================================================================
// (warp-collective) queues the slots whose tile rectangle holds more than 8 tiles: big_class 1 -> big queue, 2 -> huge queue
__device__ __forceinline__ void queue_big(const GeomParams &P, int big_class, uint32_t slot) {
	const uint32_t lane = lane_id();
	const uint32_t bmask = __ballot_sync(0xffffffffu, big_class == 1);
	if(bmask) {
		uint32_t base = 0;
		if(lane == 0) base = atomicAdd(&P.dctr->big_count, (uint32_t)__popc(bmask));
		base = __shfl_sync(0xffffffffu, base, 0);
		if(big_class == 1) P.big_queue[base + __popc(bmask & ((1u << lane) - 1u))] = slot;
	}
	if(big_class == 2) P.huge_queue[atomicAdd(&P.dctr->huge_count, 1u)] = slot; // rare: sky domes, full-screen quads
}

// WARP SUMMARY of 32 consecutive direct slots, 16 bytes {tx0 | ty0 << 16, tx1 | ty1 << 16, max of max_depth, flags}: the
// union of the tile rectangles and the nearest depth bound of the slots that bin something on this rank through a
// rectangle of at most 8 tiles. The back half tests the summary against the tile minima first: when the nearest depth of
// the 32 triangles lies behind every tile of the union, each of their (triangle, tile) pairs fails the Hi-Z test
// (max_depth <= summary depth < tile_min, main.c:1006) and none of them needs its bounds, its tile look-ups or a record
// -- a draw hidden behind earlier ones costs the draw-to-draw chain one 16-byte load and a few tile minima per 32 triangles.
// flags == 0: nothing to bin in these slots.
#define MLV_WS_SMALL 1u /* the rectangle / depth words are valid */
#define MLV_WS_FINE 2u  /* a slot needs the per-triangle path whatever the coarse test says */
__device__ __forceinline__ void write_warp_summary(const GeomParams &P, uint32_t t, uint32_t flags, uint32_t tx0, uint32_t ty0, uint32_t tx1, uint32_t ty1, uint32_t depth_bits) {
	const uint32_t f = __reduce_or_sync(0xffffffffu, flags);
	uint4 ws = make_uint4(0u, 0u, 0u, f);
	if(f & MLV_WS_SMALL) {
		const uint32_t wx0 = __reduce_min_sync(0xffffffffu, tx0), wy0 = __reduce_min_sync(0xffffffffu, ty0);
		const uint32_t wx1 = __reduce_max_sync(0xffffffffu, tx1), wy1 = __reduce_max_sync(0xffffffffu, ty1);
		ws.x = wx0 | (wy0 << 16);
		ws.y = wx1 | (wy1 << 16);
		ws.z = __reduce_max_sync(0xffffffffu, depth_bits); // non-negative floats order like their bit patterns
		// ---- the touch bytes of these triangles (touch_bin): neighbours in mesh order share their bins, so the warp merges the
		// rectangles of its lanes into one mask over the union and raises every bin once -- one predicated byte store per warp
		// instead of a look-up and a store per (triangle, tile) pair (22 us per 1.25 M-triangle draw)
		const uint32_t w = wx1 - wx0 + 1u, cnt = w * (wy1 - wy0 + 1u);
		if(cnt <= 32u) {
			uint32_t mask = 0u;
			if(flags & MLV_WS_SMALL) {
				const uint32_t row = (1u << (tx1 - tx0 + 1u)) - 1u; // at most 8 tiles wide
				for(uint32_t ty = ty0; ty <= ty1; ++ty)
					if(P.part.owns_row((int)ty)) mask |= row << ((ty - wy0) * w + (tx0 - wx0));
			}
			mask = __reduce_or_sync(0xffffffffu, mask);
			// lane -> (row, column) of the union without an integer division: lane < 32 and w <= 32, so the quotient of
			// (lane + 0.5) / w is at least 1 / 64 away from an integer and the float product cannot land on the wrong side
			const uint32_t lane = lane_id();
			const uint32_t row_of_lane = (uint32_t)(((float)lane + 0.5f) * __frcp_rn((float)w));
			if((mask >> lane) & 1u) P.touch_bits[(wy0 + row_of_lane) * (uint32_t)P.wt + wx0 + (lane - row_of_lane * w)] = 1;
		} else if(flags & MLV_WS_SMALL) { // a scattered mesh order
			const TileRect tr = { (int)tx0, (int)ty0, (int)tx1, (int)ty1 };
			touch_rect(P.touch_bits, P.part, P.wt, tr, tr.w() * tr.h());
		}
	}
	if(lane_id() == 0 && t < P.tri_count) P.warp_sum[t >> 5] = ws;
}

template <int VS, bool INDEXED, bool DEBUG, bool VCACHE>
__global__ void __launch_bounds__(MLV_GEOM_THREADS, 4) k_front(const __grid_constant__ GeomParams P) {
	MLV_KERNEL_PROLOGUE(P.timeline);
	__shared__ float s_cb[48];
	const float *cb = stage_constants(P, s_cb);
	const uint32_t lane = lane_id();
	uint32_t emitted = 0, pairs = 0;
	// Single GPU: CTA b processes chunk b, b + grid, ... Sort-first: each round, thread k of the CTA tests the cached
	// object-space bounds of the CTA's k-th chunk against this rank's tile rows, so a foreign chunk costs one thread's cull
	// test instead of a CTA (or a kernel of its own).
	__shared__ uint8_t s_live[MLV_GEOM_THREADS];
	const uint32_t num_chunks = (P.tri_count + MLV_GEOM_THREADS - 1u) / MLV_GEOM_THREADS;
	uint32_t pf0 = 0, pf1 = 0, pf2 = 0; // prefetched vertex indices of the next chunk
	bool pf_valid = false;
	for(uint32_t first = blockIdx.x; first < num_chunks; first += gridDim.x * MLV_GEOM_THREADS) {
	if(P.chunk_bounds) {
		const uint32_t c = first + threadIdx.x * gridDim.x;
		bool live = false;
		if(c < num_chunks) {
			live = !chunk_is_foreign(P, cb, c);
			P.chunk_live[c] = live ? 1 : 0; // the fill phase skips the stale bounds of the chunks nobody rewrote
			if(!live) { // ... and the back half their warp summaries
				const uint32_t n_ws = (P.tri_count + 31u) / 32u;
				for(uint32_t i = c * (MLV_GEOM_THREADS / 32u); i < min((c + 1u) * (MLV_GEOM_THREADS / 32u), n_ws); ++i) reinterpret_cast<uint32_t *>(P.warp_sum + i)[3] = 0u;
			}
		}
		s_live[threadIdx.x] = live ? 1 : 0;
		__syncthreads();
	}
	for(uint32_t k = 0; k < MLV_GEOM_THREADS; ++k) {
	const uint32_t chunk = first + k * gridDim.x;
	if(chunk >= num_chunks) break;
	if(P.chunk_bounds && !s_live[k]) continue;
	const uint32_t t = chunk * MLV_GEOM_THREADS + threadIdx.x;
	bool needs_clip = false;
	int big_class = 0; // 1: tile rectangle of 9 .. MLV_HUGE_TILES tiles, 2: larger
	uint32_t ws_flags = 0u, ws_tx0 = 0xffffu, ws_ty0 = 0xffffu, ws_tx1 = 0u, ws_ty1 = 0u, ws_depth = 0u; // this lane's share of the warp summary
	if(t < P.tri_count) {
		uint4 bounds = make_uint4(MLV_BOUNDS_EMPTY, 0u, 0u, t << 3);
		// ---- input assembler (main.c:662-696): index fetch + first half of each vertex (or its cache entry)
		uint32_t vi0, vi1, vi2;
		if(INDEXED && pf_valid) {
			vi0 = pf0, vi1 = pf1, vi2 = pf2;
		} else {
			fetch_indices<VS, INDEXED, DEBUG, VCACHE>(P, t, vi0, vi1, vi2);
		}
		if(INDEXED) {
			// persistent grid: the indices of this thread's triangle in the CTA's NEXT chunk are requested now, so the
			// first of the dependent round trips (index -> vertex) of that chunk is already over
			const uint32_t tn = t + gridDim.x * MLV_GEOM_THREADS;
			pf_valid = !P.chunk_bounds && tn < P.tri_count;
			if(pf_valid) fetch_indices<VS, INDEXED, DEBUG, VCACHE>(P, tn, pf0, pf1, pf2);
		}
		float4 a0, b0, c0, a, b, c, qa, qb, qc;
		if(VCACHE) {
			qa = __ldg(P.vcache + 2 * (size_t)vi0 + 1);
			qb = __ldg(P.vcache + 2 * (size_t)vi1 + 1);
			qc = __ldg(P.vcache + 2 * (size_t)vi2 + 1);
		} else {
			a0 = __ldg(P.vb + 2 * (size_t)vi0), b0 = __ldg(P.vb + 2 * (size_t)vi1), c0 = __ldg(P.vb + 2 * (size_t)vi2);
			// ---- vertex shader, position part (main.c:698-734)
			a = vs_position<VS>(a0, cb), b = vs_position<VS>(b0, cb), c = vs_position<VS>(c0, cb);
		}
		if(DEBUG) {
			const VsOut v0 = run_vs<VS>(a0, __ldg(P.vb + 2 * (size_t)vi0 + 1), cb, P.vs_tex, P.rsqrt_lut);
			const VsOut v1 = run_vs<VS>(b0, __ldg(P.vb + 2 * (size_t)vi1 + 1), cb, P.vs_tex, P.rsqrt_lut);
			const VsOut v2 = run_vs<VS>(c0, __ldg(P.vb + 2 * (size_t)vi2 + 1), cb, P.vs_tex, P.rsqrt_lut);
			float4 *o = reinterpret_cast<float4 *>(P.dbg.vs_out + (size_t)(3u * t) * 12);
			o[0] = v0.r0, o[1] = v0.r1, o[2] = make_float4(v0.r2x, 0.0f, 0.0f, 0.0f);
			o[3] = v1.r0, o[4] = v1.r1, o[5] = make_float4(v1.r2x, 0.0f, 0.0f, 0.0f);
			o[6] = v2.r0, o[7] = v2.r1, o[8] = make_float4(v2.r2x, 0.0f, 0.0f, 0.0f);
		}
		// ---- primitive assembly (main.c:750-898)
		bool degenerate, rejected, inside;
		if(VCACHE) { // the comparisons were made per vertex by k_vertex (clip_code)
			const uint32_t ca = __float_as_uint(qa.w), cb = __float_as_uint(qb.w), cc = __float_as_uint(qc.w);
			degenerate = ((ca | cb | cc) & MLV_CODE_W0) != 0u;
			rejected = (ca & cb & cc & MLV_CODE_OUT) != 0u;
			inside = (ca & cb & cc & MLV_CODE_IN) == MLV_CODE_IN;
		} else {
			degenerate = (a.w == 0.0f || b.w == 0.0f || c.w == 0.0f); // main.c:759
			rejected =                                                // main.c:764-772
			    (a.x < -a.w && b.x < -b.w && c.x < -c.w) || (a.x > a.w && b.x > b.w && c.x > c.w) || (a.y < -a.w && b.y < -b.w && c.y < -c.w) ||
			    (a.y > a.w && b.y > b.w && c.y > c.w) || (a.z < 0.0f && b.z < 0.0f && c.z < 0.0f) || (a.z > a.w && b.z > b.w && c.z > c.w);
			inside = // main.c:775-781
			    (a.x >= -a.w && b.x >= -b.w && c.x >= -c.w) && (a.x <= a.w && b.x <= b.w && c.x <= c.w) && (a.y >= -a.w && b.y >= -b.w && c.y >= -c.w) &&
			    (a.y <= a.w && b.y <= b.w && c.y <= c.w) && (a.z >= 0.0f && b.z >= 0.0f && c.z >= 0.0f) && (a.z <= a.w && b.z <= b.w && c.z <= c.w);
		}
		bool direct = false;
		if(!degenerate && !rejected) {
			if(inside) {
				TriSetup S;
				if(VCACHE) {
					ProjVertex pa, pb, pc; // s.x, s.y, s.w, rw are not needed for the bounds
					pa.s = make_float4(0.0f, 0.0f, qa.z, 0.0f), pa.rw = 0.0f, pa.sx = __float_as_int(qa.x), pa.sy = __float_as_int(qa.y);
					pb.s = make_float4(0.0f, 0.0f, qb.z, 0.0f), pb.rw = 0.0f, pb.sx = __float_as_int(qb.x), pb.sy = __float_as_int(qb.y);
					pc.s = make_float4(0.0f, 0.0f, qc.z, 0.0f), pc.rw = 0.0f, pc.sx = __float_as_int(qc.x), pc.sy = __float_as_int(qc.y);
					direct = setup_from_projected(pa, pb, pc, P, S);
				} else {
					direct = setup_project(a, b, c, P, S);
				}
				if(direct) {
					emitted += P.part.owns_row(S.miny / 8) ? 1u : 0u; // counted once across ranks: by the owner of its first tile row
					// (triangle, tile) pairs on this rank (Stats, main.c:1246): every tile of the bounds rectangle (main.c:927-936)
					const TileRect tr = tile_rect(S.minx, S.miny, S.maxx, S.maxy, P.wt, P.ht);
					const uint32_t np = (uint32_t)(tr.w() * P.part.owned_rows(tr.ty0, tr.ty1));
					pairs += np;
					if(np || DEBUG) {
						S.nowrap = false;
						const uint2 pb = pack_bounds(S);
						bounds = make_uint4(pb.x, pb.y, __float_as_uint(S.max_depth), t << 3);
					}
					const int cnt = tr.w() * tr.h();
					if(np && cnt > 8) big_class = cnt > MLV_HUGE_TILES ? 2 : 1;
					if(np || DEBUG) {
						// rectangles of more than 8 tiles get their record whatever Hi-Z says (count_bins); a negative or NaN depth
						// bound does not order like its bit pattern: both make the back half look at the triangles of this warp
						if(DEBUG || cnt > 8 || !(S.max_depth >= 0.0f)) ws_flags = MLV_WS_FINE;
						else ws_flags = MLV_WS_SMALL, ws_tx0 = (uint32_t)tr.tx0, ws_ty0 = (uint32_t)tr.ty0, ws_tx1 = (uint32_t)tr.tx1, ws_ty1 = (uint32_t)tr.ty1, ws_depth = __float_as_uint(S.max_depth);
					}
					if(np && ws_flags != MLV_WS_SMALL) touch_rect(P.touch_bits, P.part, P.wt, tr, cnt); // (the warp raises the touch bytes of its MLV_WS_SMALL lanes together, write_warp_summary)
				}
			} else {
				needs_clip = true;
			}
		}
		if(DEBUG && !direct) P.dbg.slot_key[t] = 0xffffffffu;
		P.tri_bounds[t] = bounds;
	}
	// ---- triangles that need the clipper are queued for k_front_clip (one warp-aggregated atomic): the slow path would
	// otherwise stall the 31 other lanes. Triangles with large tile rectangles are queued for the back half, which expands
	// them with a warp each, spread over its whole grid.
	{
		const uint32_t cmask = __ballot_sync(0xffffffffu, needs_clip);
		if(cmask) {
			uint32_t base = 0;
			if(lane == 0) base = atomicAdd(&P.dctr->clip_count, (uint32_t)__popc(cmask));
			base = __shfl_sync(0xffffffffu, base, 0);
			if(needs_clip) P.clip_queue[base + __popc(cmask & ((1u << lane) - 1u))] = t;
		}
		queue_big(P, big_class, t);
	}
	write_warp_summary(P, t, ws_flags, ws_tx0, ws_ty0, ws_tx1, ws_ty1, ws_depth);
	}
	if(P.chunk_bounds) __syncthreads(); // s_live is rewritten by the next round
	}
	tally_stats(P.stat_stripes, emitted, pairs);
}

// Fan triangulation (main.c:797) of a clipped polygon into the consecutive overflow slots T + base + j: bounds for the back
// half and -- always, the clipped attributes exist only here -- the 144-byte record in the draw's overflow arena.
// Returns the number of assembled triangles, adds the pairs to `pairs`.
__device__ __forceinline__ uint32_t emit_fan(const GeomParams &P, uint32_t t, const float *poly, int fan, uint32_t base, uint32_t &pairs, int j_first, int j_step) {
	uint32_t emitted = 0;
	for(int j = j_first; j < fan; j += j_step) {
		const uint32_t ovf = base + (uint32_t)j, slot = P.tri_count + ovf;
		const uint32_t key = (t << 3) | (uint32_t)j;
		TriSetup S;
		const VsOut p0 = poly_load(poly, 0), p1 = poly_load(poly, j + 1), p2 = poly_load(poly, j + 2);
		uint4 bounds = make_uint4(MLV_BOUNDS_EMPTY, 0u, 0u, key);
		if(setup_triangle(p0.r0, p1.r0, p2.r0, P, S)) {
			if(P.part.owns_row(S.miny / 8)) ++emitted;
			const TileRect tr = tile_rect(S.minx, S.miny, S.maxx, S.maxy, P.wt, P.ht);
			const uint32_t np = (uint32_t)(tr.w() * P.part.owned_rows(tr.ty0, tr.ty1));
			pairs += np;
			{
				const int cnt = tr.w() * tr.h();
				if(np) touch_rect(P.touch_bits, P.part, P.wt, tr, cnt);
				if(np && cnt > 8) { // (lanes diverge here: plain atomics)
					if(cnt > MLV_HUGE_TILES) P.huge_queue[atomicAdd(&P.dctr->huge_count, 1u)] = slot;
					else P.big_queue[atomicAdd(&P.dctr->big_count, 1u)] = slot;
				}
			}
			if(np || P.dbg.tris) {
				const uint2 pb = pack_bounds(S);
				bounds = make_uint4(pb.x & ~MLV_NOWRAP_BIT, pb.y, __float_as_uint(S.max_depth), key);
				TriRecord R;
				make_record(R, S, pb, p0.r1, p1.r1, p2.r1, p0.r2x, p1.r2x, p2.r2x);
				uint4 *cov = P.ovf_cov + (size_t)ovf * MLV_TRI_COV_U4;
				float4 *sh = reinterpret_cast<float4 *>(P.ovf_shade + (size_t)ovf * MLV_TRI_SHADE_U4);
#pragma unroll
				for(int i = 0; i < MLV_TRI_COV_U4; ++i) cov[i] = R.cov[i];
#pragma unroll
				for(int i = 0; i < MLV_TRI_SHADE_U4; ++i) sh[i] = R.shade[i];
			}
			if(P.dbg.tris) emit_debug(P, slot, key, S, p0.r1, p1.r1, p2.r1, p0.r2x, p1.r2x, p2.r2x);
		} else if(P.dbg.slot_key) {
			P.dbg.slot_key[slot] = 0xffffffffu;
		}
		P.tri_bounds[slot] = bounds;
	}
	return emitted;
}

#define MLV_CLIP_SPLIT 4u
// Clipping pass of the front half over the queued input triangles (dense, unlike the sparse occurrences inside k_front's
// warps). Re-runs input assembly + vertex shader for the triangle, clips, sets up the fan and writes bounds + records into
// the draw's overflow slots.
template <int VS, bool INDEXED>
__global__ void __launch_bounds__(MLV_CLIP_THREADS) k_front_clip(const __grid_constant__ GeomParams P) {
	__shared__ float s_poly[2][MLV_CLIP_MAXV * 9 * MLV_CLIP_THREADS];
	__shared__ float s_cb[48];
	MLV_KERNEL_PROLOGUE(P.timeline);
	const float *cb = stage_constants(P, s_cb);
	const uint32_t n = P.dctr->clip_count;
	const uint32_t lane = lane_id();
	uint32_t emitted = 0, pairs = 0;
	// MLV_CLIP_SPLIT consecutive lanes work on the same queued triangle: each repeats the (cheap, deterministic) vertex
	// shading and clipping and then sets up and emits every MLV_CLIP_SPLIT-th fan triangle. The kernel is one long
	// dependent chain per thread with far too few threads to hide it, so what counts is the length of that chain,
	// not the redundant work. Warp-uniform trip count: the overflow slots of a whole warp are taken with ONE atomic.
	const uint32_t sub = lane % MLV_CLIP_SPLIT;
	for(uint32_t i0 = ((blockIdx.x * blockDim.x + threadIdx.x) & ~31u) / MLV_CLIP_SPLIT; i0 < n; i0 += (gridDim.x * blockDim.x) / MLV_CLIP_SPLIT) {
		const uint32_t i = i0 + lane / MLV_CLIP_SPLIT;
		float *poly = s_poly[0] + threadIdx.x, *scratch = s_poly[1] + threadIdx.x;
		int fan = 0;
		uint32_t t = 0;
		if(i < n) {
			t = P.clip_queue[i];
			uint32_t vi0, vi1, vi2;
			fetch_indices<VS, INDEXED, false, false>(P, t, vi0, vi1, vi2);
			const VsOut v0 = run_vs<VS>(__ldg(P.vb + 2 * (size_t)vi0), __ldg(P.vb + 2 * (size_t)vi0 + 1), cb, P.vs_tex, P.rsqrt_lut);
			const VsOut v1 = run_vs<VS>(__ldg(P.vb + 2 * (size_t)vi1), __ldg(P.vb + 2 * (size_t)vi1 + 1), cb, P.vs_tex, P.rsqrt_lut);
			const VsOut v2 = run_vs<VS>(__ldg(P.vb + 2 * (size_t)vi2), __ldg(P.vb + 2 * (size_t)vi2 + 1), cb, P.vs_tex, P.rsqrt_lut);
			fan = clip_polygon(P, v0, v1, v2, poly, scratch) - 2;
			if(fan > 8) { // cannot happen for a convex clip of a triangle by six planes (<= 9 vertices)
				atomicOr(&P.ctr->error_flags, MLV_FLAG_TRI_OVERFLOW);
				fan = 0;
			}
			if(fan < 0) fan = 0;
		}
		uint32_t incl = (sub == 0u) ? (uint32_t)fan : 0u; // one allocation per triangle
#pragma unroll
		for(int d = 1; d < 32; d <<= 1) {
			const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
			if(lane >= (uint32_t)d) incl += o;
		}
		const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
		uint32_t base = 0;
		if(lane == 0 && total) base = atomicAdd(&P.dctr->ovf_count, total);
		base = __shfl_sync(0xffffffffu, base, 0) + incl - ((sub == 0u) ? (uint32_t)fan : 0u);
		base = __shfl_sync(0xffffffffu, base, (int)(lane - sub)); // the group shares the base of its first lane
		if(fan > 0) {
			if(base + (uint32_t)fan > P.ovf_capacity) atomicOr(&P.ctr->error_flags, MLV_FLAG_TRI_OVERFLOW);
			else emitted += emit_fan(P, t, poly, fan, base, pairs, (int)sub, MLV_CLIP_SPLIT);
		}
	}
	tally_stats(P.stat_stripes, emitted, pairs);
}

// =================================================================================================
// geometry, BACK half: the part of a draw that depends on what earlier draws left in the render target. Per slot with
// non-empty bounds (direct slots of k_front, overflow slots of k_front_clip): binner pass 1 with the Hi-Z test
// (main.c:924-936, 1003-1010); a DIRECT slot that survives in at least one tile fetches its vertices again, runs the
// attribute part of the vertex shader, builds the edge functions and writes its 144-byte record -- staged per warp in
// shared memory and written as contiguous 512-byte rows. A triangle hidden in every tile it touches costs 16 bytes of
// bounds and its tile-minimum look-ups, and writes nothing. The chunk loop is software-pipelined (the bounds of the CTA's
// next chunk are in flight while the current one is processed): at one chunk per round trip chain the kernel was
// latency-bound at every size.
// =================================================================================================
struct BackState {
	uint32_t records, survivors;
	bool wake;
};

struct SlotBounds {
	TileRect tr;
	float max_depth;
	uint32_t key;
	bool empty;
};
__device__ __forceinline__ SlotBounds load_bounds(const uint4 *__restrict__ tri_bounds, uint32_t slot, int wt, int ht) {
	const uint4 b = __ldg(tri_bounds + slot);
	SlotBounds s;
	s.empty = (b.x & 0xffffu) == MLV_BOUNDS_EMPTY;
	s.tr = tile_rect((int)(b.x & 0xffffu), (int)((b.x >> 16) & 0x7fffu), (int)(short)(b.y & 0xffffu), (int)(short)(b.y >> 16), wt, ht);
	s.max_depth = __uint_as_float(b.z);
	s.key = b.w;
	return s;
}

// Pass 1 of the binner (main.c:924-936) for the triangles whose tile rectangle holds more than 8 tiles (queued by the
// front half). Rectangles of up to MLV_HUGE_TILES tiles: one warp per triangle, lanes stride over the rectangle with four
// tiles in flight. Larger ones (sky domes, full-screen quads): the whole grid strides over the rectangle.
__device__ __forceinline__ void count_big_rects(const GeomParams &P, BackState &st, uint32_t n, uint32_t nhuge) {
	if((n | nhuge) == 0u) return;
	const uint32_t lane = lane_id();
	const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
	BinTally r = { false, false, false, false, 0u };
	auto one = [&](int k, int cnt, int w, const SlotBounds &s, float &tm, uint32_t &tc, uint32_t &bin, bool &ok) {
		ok = k < cnt;
		if(ok) {
			const int ty = s.tr.ty0 + k / w, tx = s.tr.tx0 + k % w;
			ok = P.part.owns_row(ty);
			bin = (uint32_t)(ty * P.wt + tx);
			if(ok) tm = __ldg(P.tile_min + bin), tc = (uint32_t)__ldcg(P.touch_bits + bin);
		}
	};
	auto expand = [&](const SlotBounds &s, int first, int stride) {
		if(s.empty) return;
		const int w = s.tr.w(), cnt = w * s.tr.h();
		for(int k = first; k < cnt; k += 4 * stride) {
			float tm[4];
			uint32_t tc[4], bin[4];
			bool ok[4];
#pragma unroll
			for(int u = 0; u < 4; ++u) one(k + u * stride, cnt, w, s, tm[u], tc[u], bin[u], ok[u]);
#pragma unroll
			for(int u = 0; u < 4; ++u)
				if(ok[u]) {
					touch_bin(P.touch_bits, bin[u], tc[u]);
					count_pair(P, bin[u], tm[u], s.max_depth, r);
				}
		}
	};
	for(uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) expand(load_bounds(P.tri_bounds, P.big_queue[i], P.wt, P.ht), (int)lane, 32);
	for(uint32_t i = 0; i < nhuge; ++i) expand(load_bounds(P.tri_bounds, P.huge_queue[i], P.wt, P.ht), (int)(blockIdx.x * blockDim.x + threadIdx.x), (int)(gridDim.x * blockDim.x));
	st.survivors += r.survivors;
	st.wake = st.wake || r.wake;
}

template <int VS, bool INDEXED, bool DEBUG, bool VCACHE>
__device__ __forceinline__ void back_process(const GeomParams &P, uint4 (*s_stage)[32 * (MLV_TRI_COV_U4 + MLV_TRI_SHADE_U4)], uint32_t slot, uint4 b, const float *cb, BackState &st) {
	const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
	bool staged = false, live = false, present = false;
	TileRect tr = { 0, 0, -1, -1 };
	float max_depth = 0.0f;
	if((b.x & 0xffffu) != MLV_BOUNDS_EMPTY) {
		present = true;
		const int minx = (int)(b.x & 0xffffu), miny = (int)((b.x >> 16) & 0x7fffu), maxx = (int)(short)(b.y & 0xffffu), maxy = (int)(short)(b.y >> 16);
		max_depth = __uint_as_float(b.z);
		tr = tile_rect(minx, miny, maxx, maxy, P.wt, P.ht);
		// ---- binner pass 1 + Hi-Z for this triangle
		const BinTally tally = count_bins(P, tr, max_depth);
		live = tally.live; // (rectangles of more than 8 tiles: counted below by a warp each, from the front half's queue)
		st.survivors += tally.survivors;
		st.wake = st.wake || tally.wake;
	}
	if(present) {
		if(!live && !DEBUG) reinterpret_cast<uint32_t *>(P.tri_bounds + slot)[0] = MLV_BOUNDS_EMPTY; // the fill phase skips it
		if(live) ++st.records;
		if(slot < P.tri_count && (live || DEBUG)) {
			// ---- the record of a surviving direct triangle: input assembler + vertex shader + setup again (pure functions
			// of the same inputs: the values k_front computed), now including attributes and edge functions
			const uint32_t t = slot;
			uint32_t vi0, vi1, vi2;
			fetch_indices<VS, INDEXED, DEBUG, VCACHE>(P, t, vi0, vi1, vi2);
			float4 a0 = __ldg(P.vb + 2 * (size_t)vi0), b0 = __ldg(P.vb + 2 * (size_t)vi1), c0 = __ldg(P.vb + 2 * (size_t)vi2);
			float4 a, bq, c;
			TriSetup S;
			if(VCACHE) {
				const float4 qa = __ldg(P.vcache + 2 * (size_t)vi0 + 1), qb = __ldg(P.vcache + 2 * (size_t)vi1 + 1), qc = __ldg(P.vcache + 2 * (size_t)vi2 + 1);
				a = __ldg(P.vcache + 2 * (size_t)vi0), bq = __ldg(P.vcache + 2 * (size_t)vi1), c = __ldg(P.vcache + 2 * (size_t)vi2);
				ProjVertex pa, pb, pc; // s.x, s.y, s.w are not needed outside debug capture (which never uses the cache)
				pa.s = make_float4(0.0f, 0.0f, qa.z, 0.0f), pa.sx = __float_as_int(qa.x), pa.sy = __float_as_int(qa.y);
				pb.s = make_float4(0.0f, 0.0f, qb.z, 0.0f), pb.sx = __float_as_int(qb.x), pb.sy = __float_as_int(qb.y);
				pc.s = make_float4(0.0f, 0.0f, qc.z, 0.0f), pc.sx = __float_as_int(qc.x), pc.sy = __float_as_int(qc.y);
				pa.rw = 1.0f / a.w, pb.rw = 1.0f / bq.w, pc.rw = 1.0f / c.w; // a_reciprocal_ws (project_vertex): the same correctly rounded divide
				setup_from_projected(pa, pb, pc, P, S);
			} else {
				a = vs_position<VS>(a0, cb), bq = vs_position<VS>(b0, cb), c = vs_position<VS>(c0, cb);
				setup_project(a, bq, c, P, S);
			}
			setup_edges(P, S);
			// ---- vertex shader, attribute part
			float4 r1a, r1b, r1c;
			float r2a, r2b, r2c;
			vs_attributes<VS>(a0, __ldg(P.vb + 2 * (size_t)vi0 + 1), a, cb, P.vs_tex, P.rsqrt_lut, r1a, r2a);
			vs_attributes<VS>(b0, __ldg(P.vb + 2 * (size_t)vi1 + 1), bq, cb, P.vs_tex, P.rsqrt_lut, r1b, r2b);
			vs_attributes<VS>(c0, __ldg(P.vb + 2 * (size_t)vi2 + 1), c, cb, P.vs_tex, P.rsqrt_lut, r1c, r2c);
			if(live) {
				staged = true;
				TriRecord R;
				make_record(R, S, pack_bounds(S), r1a, r1b, r1c, r2a, r2b, r2c);
				uint4 *stg = s_stage[warp];
#pragma unroll
				for(int i = 0; i < MLV_TRI_COV_U4; ++i) stg[lane * MLV_TRI_COV_U4 + i] = R.cov[i];
#pragma unroll
				for(int i = 0; i < MLV_TRI_SHADE_U4; ++i)
					stg[32 * MLV_TRI_COV_U4 + lane * MLV_TRI_SHADE_U4 + i] = make_uint4(__float_as_uint(R.shade[i].x), __float_as_uint(R.shade[i].y), __float_as_uint(R.shade[i].z), __float_as_uint(R.shade[i].w));
			}
			if(DEBUG) emit_debug(P, t, t << 3, S, r1a, r1b, r1c, r2a, r2b, r2c);
		}
	}
	// ---- coalesced write-out of the staged records: the direct slots of a warp are adjacent in HBM, so the warp stores
	// 1536 B of TriCov and 3072 B of TriShade with 128-bit stores instead of 32 scattered 16-byte pieces per instruction
	__syncwarp();
	const uint32_t valid = __ballot_sync(0xffffffffu, staged);
	if(valid) {
		const uint4 *stg = s_stage[warp];
		const size_t slot0 = (size_t)(slot - lane);
		uint4 *cov = P.tri_cov + slot0 * MLV_TRI_COV_U4;
		uint4 *sh = P.tri_shade + slot0 * MLV_TRI_SHADE_U4;
#pragma unroll
		for(int i = 0; i < MLV_TRI_COV_U4; ++i) {
			const uint32_t piece = i * 32 + lane;
			if((valid >> (piece / MLV_TRI_COV_U4)) & 1u) cov[piece] = stg[piece];
		}
#pragma unroll
		for(int i = 0; i < MLV_TRI_SHADE_U4; ++i) {
			const uint32_t piece = i * 32 + lane;
			if((valid >> (piece / MLV_TRI_SHADE_U4)) & 1u) sh[piece] = stg[32 * MLV_TRI_COV_U4 + piece];
		}
	}
	__syncwarp(); // the staging rows are reused by the next chunk
}

// Coarse Hi-Z of one warp summary (see write_warp_summary) by TWO adjacent lanes, each looking at up to 16 tiles of the
// union rectangle, so that all tile minima of a summary are requested at once: true when the 32 slots need the
// per-triangle path (the same answer in both lanes).
#define MLV_BACK_GROUP 128u /* warp summaries a CTA tests per round (4096 triangles) */
__device__ __forceinline__ bool summary_needs_triangles(const GeomParams &P, uint32_t wsi, bool in_range, uint32_t sub) {
	uint4 ws = make_uint4(0u, 0u, 0u, 0u);
	if(in_range) ws = __ldcg(P.warp_sum + wsi);
	// ws.w == 0: nothing to bin in these slots (or: a chunk the front half culled on this rank)
	bool need = ws.w != 0u, test = false;
	int hidden = 1;
	if(need && !(ws.w & MLV_WS_FINE) && !P.keep_all) {
		const int tx0 = (int)(ws.x & 0xffffu), ty0 = (int)(ws.x >> 16), tx1 = (int)(ws.y & 0xffffu), ty1 = (int)(ws.y >> 16);
		const int w = tx1 - tx0 + 1, cnt = w * (ty1 - ty0 + 1);
		if(cnt <= 32) { // (more: a scattered mesh order, the union says little)
			test = true;
			const float nearest = __uint_as_float(ws.z);
			const int k0 = (int)sub * 16;
			int tx = tx0 + k0 % w, ty = ty0 + k0 / w;
			float tm[16];
			bool ok[16];
#pragma unroll
			for(int u = 0; u < 16; ++u) {
				// (sort-first: rows of other ranks inside the union are tested like owned ones -- their minima still carry the
				// clear tag here, so a warp that straddles a band border takes the per-triangle path; asking owns_row per tile
				// cost 32 integer divisions per lane, 3 us of the kernel)
				ok[u] = k0 + u < cnt;
				tm[u] = ok[u] ? __ldg(P.tile_min + (uint32_t)(ty * P.wt + tx)) : 0.0f;
				if(++tx > tx1) {
					tx = tx0;
					++ty;
				}
			}
#pragma unroll
			for(int u = 0; u < 16; ++u) // count_pair's test with the nearest depth of the 32 triangles; a tile that still carries the clear tag hides nothing (wake logic)
				if(ok[u] && !((nearest < tm[u]) && __float_as_uint(tm[u]) != MLV_TILE_MIN_CLEARED)) hidden = 0;
		}
	}
	// the two lanes of a summary agree on everything but `hidden`
	hidden &= __shfl_xor_sync(0xffffffffu, hidden, 1);
	if(test) {
		need = hidden == 0;
		if(hidden && sub == 0u) reinterpret_cast<uint32_t *>(P.warp_sum + wsi)[3] = 0u; // the fill pass skips these slots as well
	}
	return need;
}

template <int VS, bool INDEXED, bool DEBUG, bool VCACHE>
__global__ void __launch_bounds__(MLV_GEOM_THREADS, 4) k_back(const __grid_constant__ GeomParams P) {
	MLV_KERNEL_PROLOGUE(P.timeline);
	__shared__ uint4 s_stage[MLV_GEOM_THREADS / 32][32 * (MLV_TRI_COV_U4 + MLV_TRI_SHADE_U4)];
	__shared__ uint32_t s_list[MLV_BACK_GROUP]; // the warp summaries of this round that need their triangles, compacted
	__shared__ uint32_t s_warp_count[MLV_GEOM_THREADS / 32];
	__shared__ float s_cb[48];
	const float *cb = stage_constants(P, s_cb);
	const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
	const uint32_t ovf_raw = P.dctr->ovf_count, n_big = P.dctr->big_count, n_huge = P.dctr->huge_count; // (in flight with the first summaries; only needed after the direct slots)
	// CTA g takes the summaries j * gridDim.x + g: every CTA samples the whole draw (sort-first: a rank's live triangles are
	// one contiguous run of the mesh order; contiguous groups would leave most CTAs without work), MLV_BACK_GROUP of them
	// per round -- one round for draws of up to gridDim.x * 4096 triangles (2.4 M on a B200).
	const uint32_t n_ws = (P.tri_count + 31u) / 32u, per_cta = (n_ws + gridDim.x - 1u) / gridDim.x;
	BackState st = { 0u, 0u, false };
	for(uint32_t j0 = 0; j0 < per_cta; j0 += MLV_BACK_GROUP) {
		// ---- phase 1, two lanes per warp summary: coarse Hi-Z. The front half cleared the summaries of the chunks it culled.
		const uint32_t wsi = (j0 + (threadIdx.x >> 1)) * gridDim.x + blockIdx.x, sub = threadIdx.x & 1u;
		const bool fine = summary_needs_triangles(P, wsi, wsi < n_ws, sub) && sub == 0u;
		const uint32_t m = __ballot_sync(0xffffffffu, fine);
		if(lane == 0) s_warp_count[warp] = (uint32_t)__popc(m);
		__syncthreads();
		uint32_t before = 0, total = 0;
#pragma unroll
		for(uint32_t w = 0; w < MLV_GEOM_THREADS / 32u; ++w) {
			const uint32_t cnt = s_warp_count[w];
			before += (w < warp) ? cnt : 0u;
			total += cnt;
		}
		if(fine) s_list[before + (uint32_t)__popc(m & ((1u << lane) - 1u))] = wsi;
		__syncthreads();
		// ---- phase 2, one warp per listed summary: the per-triangle path, the bounds of the warp's next item in flight
		const uint4 empty = make_uint4(MLV_BOUNDS_EMPTY, 0u, 0u, 0u);
		uint4 b_next = empty;
		if(warp < total) {
			const uint32_t s0 = s_list[warp] * 32u + lane;
			if(s0 < P.tri_count) b_next = P.tri_bounds[s0];
		}
		for(uint32_t i = warp; i < total; i += MLV_GEOM_THREADS / 32u) {
			const uint32_t slot = s_list[i] * 32u + lane;
			const uint4 b = b_next;
			b_next = empty;
			if(i + MLV_GEOM_THREADS / 32u < total) {
				const uint32_t sn = s_list[i + MLV_GEOM_THREADS / 32u] * 32u + lane;
				if(sn < P.tri_count) b_next = P.tri_bounds[sn];
			}
			back_process<VS, INDEXED, DEBUG, VCACHE>(P, s_stage, slot, b, cb, st);
		}
		__syncthreads(); // s_list is rewritten by the next round
	}
	// overflow slots: the fan triangles of clipped input triangles (their records exist already)
	const uint32_t n_ovf = min(ovf_raw, P.ovf_capacity);
	for(uint32_t base = blockIdx.x * MLV_GEOM_THREADS; base < n_ovf; base += gridDim.x * MLV_GEOM_THREADS) {
		const uint32_t o = base + threadIdx.x;
		const uint4 b = (o < n_ovf) ? P.tri_bounds[P.tri_count + o] : make_uint4(MLV_BOUNDS_EMPTY, 0u, 0u, 0u);
		back_process<VS, INDEXED, DEBUG, VCACHE>(P, s_stage, P.tri_count + o, b, cb, st);
	}
	// ---- rectangles of more than 8 tiles (the front half queued them): a warp each, spread over the whole grid
	count_big_rects(P, st, n_big, n_huge);
	// ---- per-draw tallies: records written (work counter, one atomic per warp); "a pair survived Hi-Z" (the scan, the fill
	// pass and the tile kernel return at once otherwise)
#pragma unroll
	for(int d = 16; d > 0; d >>= 1) {
		st.records += __shfl_xor_sync(0xffffffffu, st.records, d);
		st.survivors += __shfl_xor_sync(0xffffffffu, st.survivors, d);
	}
	const bool wake = __any_sync(0xffffffffu, st.wake);
	if(lane == 0) {
		const uint32_t stripe = (blockIdx.x * 8u + warp) % MLV_STAT_STRIPES;
		if(st.records) atomicAdd(P.stat_stripes + stripe * 16u + 1u, (unsigned long long)st.records);
		if(st.survivors) P.ctr->draw_alive = 1u; // (the exact total is the scan's)
		if(wake) P.ctr->n_wake = 1u;
	}
}

// =================================================================================================
// binner: k_bin_scan (exclusive scan of the per-bin counts pass 1 left + work list), k_fill (pass 2). All of them, and
// k_tile, return at once for a draw whose pairs were all rejected by Hi-Z (nothing is left behind by such a draw: bins
// are marked as touched by the draw's epoch, not by flags that would need resetting).
// =================================================================================================

// Pass 2 of the binner (main.c:950-962): every surviving (triangle, tile) pair takes the next position of its
// bin's list (atomic on the running offset the scan left in bin_offset). One lane per triangle slot; rectangles
// of more than 8 tiles are expanded cooperatively by the warp (four independent atomics in flight per lane),
// huge ones by the whole grid. The per-bin order this leaves is arbitrary; k_tile does not depend on it.
__device__ __forceinline__ void fill_one(const TailParams &P, int k, int cnt, int w, int tx0, int ty0, float max_depth, uint32_t slot) {
	if(k >= cnt) return;
	const int ty = ty0 + k / w, tx = tx0 + k % w;
	if(!P.part.owns_row(ty)) return;
	const uint32_t bin = (uint32_t)(ty * P.wt + tx);
	if(!hiz_rejects(max_depth, P.tile_min, bin, P.keep_all)) P.pair_ids[atomicAdd(P.bin_offset + bin, 1u)] = slot;
}

// unrolled, predicated walk over a rectangle of at most N tiles: all tile-minimum loads, then all atomics, then all stores
template <int N>
__device__ __forceinline__ void fill_small_rect(const TailParams &P, const SlotBounds &s, int cnt, uint32_t slot) {
	uint32_t bins[N], pos[N];
	bool ok[N];
	int tx = s.tr.tx0, ty = s.tr.ty0;
#pragma unroll
	for(int k = 0; k < N; ++k) {
		ok[k] = k < cnt && P.part.owns_row(ty);
		bins[k] = (uint32_t)(ty * P.wt + tx);
		if(++tx > s.tr.tx1) {
			tx = s.tr.tx0;
			++ty;
		}
	}
#pragma unroll
	for(int k = 0; k < N; ++k) ok[k] = ok[k] && !hiz_rejects(s.max_depth, P.tile_min, bins[k], P.keep_all);
#pragma unroll
	for(int k = 0; k < N; ++k) pos[k] = ok[k] ? atomicAdd(P.bin_offset + bins[k], 1u) : 0u;
#pragma unroll
	for(int k = 0; k < N; ++k)
		if(ok[k]) P.pair_ids[pos[k]] = slot;
}

__global__ void __launch_bounds__(256) k_fill(const __grid_constant__ TailParams P) {
	MLV_KERNEL_PROLOGUE(P.timeline ? P.timeline + 4 : nullptr);
	const uint32_t total = P.ctr->pair_total;
	// draw skipped (pair arena or overflow slots exhausted: k_tile raises the flag) / nothing survived Hi-Z
	if(total == 0u || total > P.pair_capacity || P.dctr->ovf_count > P.ovf_capacity) return;
	const uint32_t n = P.direct_slots + min(P.dctr->ovf_count, P.ovf_capacity);
	const uint32_t lane = lane_id();
	const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
	for(uint32_t base = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32u; base < n; base += warps * 32u) {
		const uint32_t slot = base + lane;
		SlotBounds s;
		s.empty = true;
		// slots of chunks the front half skipped on this rank hold stale bounds from an earlier draw
		const bool live_chunk = !P.chunk_live || slot >= P.direct_slots || __ldg(P.chunk_live + slot / MLV_GEOM_THREADS) != 0;
		// 32 direct slots whose warp summary says "nothing to bin" (the front half) or "hidden by Hi-Z" (the back half's coarse test)
		const bool summarised_out = base + 32u <= P.direct_slots && live_chunk && __ldg(P.warp_sum + (base >> 5) * 4u + 3u) == 0u;
		if(slot < n && live_chunk && !summarised_out) s = load_bounds(P.tri_bounds, slot, P.wt, P.ht);
		const int w = s.empty ? 0 : s.tr.w(), cnt = s.empty ? 0 : w * s.tr.h();
		if(cnt > 0 && cnt <= 4) fill_small_rect<4>(P, s, cnt, slot);
		else if(cnt > 0 && cnt <= 8) fill_small_rect<8>(P, s, cnt, slot);
	}
	const uint32_t nbig = P.dctr->big_count;
	for(uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < nbig; i += warps) { // 9..MLV_HUGE_TILES tiles: one warp per triangle
		const uint32_t slot = P.big_queue[i];
		const SlotBounds s = load_bounds(P.tri_bounds, slot, P.wt, P.ht);
		if(s.empty) continue;
		const int w = s.tr.w(), cnt = w * s.tr.h();
		for(int k = (int)lane; k < cnt; k += 128) {
			fill_one(P, k, cnt, w, s.tr.tx0, s.tr.ty0, s.max_depth, slot);
			fill_one(P, k + 32, cnt, w, s.tr.tx0, s.tr.ty0, s.max_depth, slot);
			fill_one(P, k + 64, cnt, w, s.tr.tx0, s.tr.ty0, s.max_depth, slot);
			fill_one(P, k + 96, cnt, w, s.tr.tx0, s.tr.ty0, s.max_depth, slot);
		}
	}
	const uint32_t nhuge = P.dctr->huge_count;
	for(uint32_t i = 0; i < nhuge; ++i) { // huge rectangles, grid-cooperative
		const uint32_t slot = P.huge_queue[i];
		const SlotBounds s = load_bounds(P.tri_bounds, slot, P.wt, P.ht);
		if(s.empty) continue;
		const int w = s.tr.w(), cnt = w * s.tr.h();
		for(int k = (int)(blockIdx.x * blockDim.x + threadIdx.x); k < cnt; k += (int)(gridDim.x * blockDim.x)) fill_one(P, k, cnt, w, s.tr.tx0, s.tr.ty0, s.max_depth, slot);
	}
}

// ---- decoupled look-back over CTAs (single-pass prefix sum) ---------------------------------------
#define MLV_SCAN_INVALID 0ull
#define MLV_SCAN_AGGREGATE 1ull
#define MLV_SCAN_PREFIX 2ull

__device__ __forceinline__ unsigned long long scan_pack(uint32_t epoch, unsigned long long flag, uint32_t value) {
	return ((unsigned long long)epoch << 34) | (flag << 32) | (unsigned long long)value;
}

// Executed by one full warp of CTA `tile`: publishes the CTA aggregate, returns the exclusive prefix over all
// earlier CTAs. State words carry the draw epoch, so the array never needs clearing between draws.
__device__ __forceinline__ uint32_t lookback_exclusive(volatile unsigned long long *state, uint32_t tile, uint32_t epoch, uint32_t block_total) {
	const uint32_t lane = lane_id();
	if(lane == 0) state[tile] = scan_pack(epoch, tile == 0 ? MLV_SCAN_PREFIX : MLV_SCAN_AGGREGATE, block_total);
	uint32_t exclusive = 0;
	if(tile > 0) {
		int look = (int)tile - 1;
		while(true) {
			const int idx = look - (int)lane;
			unsigned long long s;
			bool ready;
			do { // spin until the whole window has been published for this draw
				s = (idx >= 0) ? state[idx] : scan_pack(epoch, MLV_SCAN_PREFIX, 0u);
				ready = ((uint32_t)(s >> 34) == epoch) && (((s >> 32) & 3ull) != MLV_SCAN_INVALID);
			} while(!__all_sync(0xffffffffu, ready));
			const bool is_prefix = ((s >> 32) & 3ull) == MLV_SCAN_PREFIX;
			const uint32_t pmask = __ballot_sync(0xffffffffu, is_prefix);
			uint32_t val = (uint32_t)s;
			if(pmask) {
				const int first = __ffs(pmask) - 1;
				if((int)lane > first) val = 0u;
			}
#pragma unroll
			for(int d = 16; d > 0; d >>= 1) val += __shfl_xor_sync(0xffffffffu, val, d);
			exclusive += val;
			if(pmask) break;
			look -= 32;
		}
		if(lane == 0) state[tile] = scan_pack(epoch, MLV_SCAN_PREFIX, exclusive + block_total);
	}
	return exclusive;
}

// Exclusive scan of the per-bin counts + compaction of the bins k_tile has to visit, in ascending bin index
// (main.c:937-974), Stats (main.c:1245-1246). One bin per thread, 1024 bins per CTA; warp 0 resolves the
// pair-count prefix and warp 1 the work-list prefix concurrently. A bin is non-empty (counts for Stats, would
// be in the reference's CompactedBin list) when it was touched at all; it is on the work list when it holds
// surviving pairs, or when it was touched and its tile minimum still is the clear marker (write_tile refreshes
// the minimum of every non-empty bin, main.c:589-603 -- for any other fully Hi-Z-rejected bin that refresh
// would rewrite the value already stored).
#define MLV_SCAN_THREADS 1024
#define MLV_SCAN_ITEMS 4
__global__ void __launch_bounds__(MLV_SCAN_THREADS) k_bin_scan(const __grid_constant__ TailParams P) {
	MLV_KERNEL_PROLOGUE(P.timeline);
	if(P.ctr->draw_alive == 0u && P.ctr->n_wake == 0u) return; // nothing survived Hi-Z in this draw: there is nothing to lay out (k_fill and k_tile return as well)
	__shared__ uint32_t s_tile;
	__shared__ uint32_t s_sum[32], s_nz[32];
	__shared__ uint32_t s_excl[2];
	// CTAs take their logical position from a ticket so that every predecessor a CTA may wait on has started
	// (the ticket counter is re-armed and the epoch advanced by k_tile at the end of every draw: nothing per draw comes from
	// the host, so a recorded frame -- a CUDA graph, mlv_execute_command_list -- replays unchanged)
	if(threadIdx.x == 0) s_tile = atomicAdd(&P.ctr->ticket, 1u);
	__syncthreads();
	const uint32_t tile = s_tile;
	const uint32_t epoch = P.ctr->epoch, n_wake = P.ctr->n_wake;
	const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
	const uint32_t base = P.bin_begin + (tile * MLV_SCAN_THREADS + threadIdx.x) * MLV_SCAN_ITEMS; // 4 consecutive bins per thread, 16-byte accesses
	uint32_t raw[MLV_SCAN_ITEMS], c[MLV_SCAN_ITEMS];
	bool work[MLV_SCAN_ITEMS];
	const bool full = base + MLV_SCAN_ITEMS <= P.bin_end;
	if(full) {
		const uint4 r = *reinterpret_cast<const uint4 *>(P.bin_count + base);
		raw[0] = r.x, raw[1] = r.y, raw[2] = r.z, raw[3] = r.w;
	} else {
#pragma unroll
		for(int k = 0; k < MLV_SCAN_ITEMS; ++k) raw[k] = (base + k < P.bin_end) ? P.bin_count[base + k] : 0u;
	}
	uint32_t tsum = 0, twork = 0;
#pragma unroll
	for(int k = 0; k < MLV_SCAN_ITEMS; ++k) {
		c[k] = raw[k];
		// (a touched bin without surviving pairs whose tile minimum still carries the clear tag is visited for write_tile's
		// refresh, main.c:589-603; it takes a pair with negative depth, the back half counts such bins in n_wake)
		work[k] = c[k] != 0u || (n_wake != 0u && base + k < P.bin_end && P.touch_bits[base + k] != 0 && __float_as_uint(P.tile_min[base + k]) == MLV_TILE_MIN_CLEARED);
		tsum += c[k];
		twork += work[k] ? 1u : 0u;
	}
	if(twork) { // counters back to zero for the next draw
		if(full) *reinterpret_cast<uint4 *>(P.bin_count + base) = make_uint4(0u, 0u, 0u, 0u);
		else
			for(int k = 0; k < MLV_SCAN_ITEMS; ++k)
				if(base + k < P.bin_end) P.bin_count[base + k] = 0u;
	}
	uint32_t incl = tsum, wincl_work = twork;
#pragma unroll
	for(int d = 1; d < 32; d <<= 1) {
		const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d), o2 = __shfl_up_sync(0xffffffffu, wincl_work, d);
		if(lane >= (uint32_t)d) {
			incl += o;
			wincl_work += o2;
		}
	}
	if(lane == 31) {
		s_sum[warp] = incl;
		s_nz[warp] = wincl_work;
	}
	__syncthreads();
	if(warp < 2) {
		const uint32_t v = (warp == 0) ? s_sum[lane] : s_nz[lane];
		uint32_t wincl = v;
#pragma unroll
		for(int d = 1; d < 32; d <<= 1) {
			const uint32_t o = __shfl_up_sync(0xffffffffu, wincl, d);
			if(lane >= (uint32_t)d) wincl += o;
		}
		const uint32_t block_total = __shfl_sync(0xffffffffu, wincl, 31);
		const uint32_t excl = lookback_exclusive(warp == 0 ? P.state_sum : P.state_nz, tile, epoch, block_total);
		if(warp == 0) s_sum[lane] = wincl - v;
		else s_nz[lane] = wincl - v;
		if(lane == 0) s_excl[warp] = excl;
		if(tile == P.scan_blocks - 1 && lane == 0) {
			const uint32_t total = excl + block_total;
			if(warp == 0) P.ctr->pair_total = total; // (k_fill and k_tile skip a draw whose pairs or overflow slots do not fit)
			else P.ctr->n_cbins = total;
		}
	}
	__syncthreads();
	uint32_t upto = s_excl[0] + s_sum[warp] + incl - tsum;
	uint32_t wpos = s_excl[1] + s_nz[warp] + wincl_work - twork;
	uint32_t offs[MLV_SCAN_ITEMS];
#pragma unroll
	for(int k = 0; k < MLV_SCAN_ITEMS; ++k) {
		offs[k] = upto;
		if(work[k]) {
			mlv_ref_compacted_bin cb;
			cb.num_triangles_self = c[k];
			cb.num_triangles_upto = upto;
			cb.bin_index = base + k;
			P.cbins[wpos++] = cb;
		}
		upto += c[k];
	}
	if(full) *reinterpret_cast<uint4 *>(P.bin_offset + base) = make_uint4(offs[0], offs[1], offs[2], offs[3]);
	else
		for(int k = 0; k < MLV_SCAN_ITEMS; ++k)
			if(base + k < P.bin_end) P.bin_offset[base + k] = offs[k];
}

// =================================================================================================
// tile: rasterizer + Hi-Z + early-Z + pixel shader + output merger
// =================================================================================================

// record / key of a slot (mlv_internal.cuh "Triangle identity")
__device__ __forceinline__ const uint4 *cov_of(const TailParams &P, uint32_t slot) {
	return slot < P.direct_slots ? P.tri_cov + (size_t)slot * MLV_TRI_COV_U4 : P.ovf_cov + (size_t)(slot - P.direct_slots) * MLV_TRI_COV_U4;
}
__device__ __forceinline__ const float4 *shade_of(const TailParams &P, uint32_t slot) {
	return reinterpret_cast<const float4 *>(slot < P.direct_slots ? P.tri_shade + (size_t)slot * MLV_TRI_SHADE_U4 : P.ovf_shade + (size_t)(slot - P.direct_slots) * MLV_TRI_SHADE_U4);
}
__device__ __forceinline__ uint32_t key_of(const TailParams &P, uint32_t slot) {
	return slot < P.direct_slots ? (slot << 3) : __ldg(reinterpret_cast<const uint32_t *>(P.tri_bounds + slot) + 3);
}

// Debug capture only: restores ascending-KEY order inside one bin list of slots (the order the reference's serial fill
// produces, main.c:950-962). n <= 32: bitonic network in registers. Larger lists: stable LSD radix split on the key, one
// bit per pass, ping-ponging between the list and a scratch segment of the same extent.
__device__ __forceinline__ void sort_bin_ids(const TailParams &P, uint32_t *ids, uint32_t *tmp, uint32_t n, uint32_t key_bits) {
	const uint32_t lane = lane_id();
	if(n <= 32u) {
		const uint32_t slot = (lane < n) ? ids[lane] : 0xffffffffu;
		unsigned long long v = (lane < n) ? (((unsigned long long)key_of(P, slot) << 32) | slot) : ~0ull;
#pragma unroll
		for(int k = 2; k <= 32; k <<= 1) {
#pragma unroll
			for(int j = k >> 1; j > 0; j >>= 1) {
				const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, j);
				const bool up = ((lane & k) == 0);
				const bool lower = ((lane & j) == 0);
				v = (lower == up) ? min(v, o) : max(v, o);
			}
		}
		if(lane < n) ids[lane] = (uint32_t)v;
		__syncwarp();
		return;
	}
	volatile uint32_t *src = ids;
	volatile uint32_t *dst = tmp;
	const uint32_t lt = (1u << lane) - 1u;
	for(uint32_t bit = 0; bit < key_bits; ++bit) {
		uint32_t zeros = 0;
		for(uint32_t i0 = 0; i0 < n; i0 += 32u) {
			const uint32_t i = i0 + lane;
			const bool isz = (i < n) && !((key_of(P, src[i]) >> bit) & 1u);
			zeros += __popc(__ballot_sync(0xffffffffu, isz));
		}
		if(zeros == 0u || zeros == n) continue; // this bit does not discriminate
		uint32_t zpos = 0, opos = zeros;
		for(uint32_t i0 = 0; i0 < n; i0 += 32u) {
			const uint32_t i = i0 + lane;
			const bool valid = i < n;
			const uint32_t val = valid ? src[i] : 0u;
			const bool isz = valid && !((key_of(P, val) >> bit) & 1u);
			const uint32_t bz = __ballot_sync(0xffffffffu, isz);
			const uint32_t bo = __ballot_sync(0xffffffffu, valid && !isz);
			if(isz) dst[zpos + __popc(bz & lt)] = val;
			else if(valid) dst[opos + __popc(bo & lt)] = val;
			zpos += __popc(bz);
			opos += __popc(bo);
		}
		__syncwarp();
		volatile uint32_t *sw = src;
		src = dst;
		dst = sw;
	}
	if(src != ids) {
		for(uint32_t i = lane; i < n; i += 32u) ids[i] = src[i];
	}
	__syncwarp();
}

// Coverage tests of one triangle against one tile (main.c:1012-1038): E_k = ((a_k*x)<<4) + ((b_k*y)<<4) + c_k in
// wrapping i32 == a_k*(16x) + b_k*(16y) + c_k (mod 2^32), stepped incrementally; inside iff (E0|E1|E2) > 0.
// lo = rows 0-3, hi = rows 4-7 of the reference's 64-bit fragment mask (bit 8*row + col).
// Rows/columns [y0,y1] x [x0,x1] (tile-relative) are evaluated: the full tile, or -- for triangles whose
// arithmetic provably never wraps (TriSetup::nowrap) -- the part of the tile inside the triangle's bounds,
// outside of which the full evaluation yields 0 anyway.
__device__ __forceinline__ void coverage(const uint4 c0, const uint4 c1, const uint32_t c2x, uint32_t X0, uint32_t Y0, int x0, int x1, int y0, int y1, uint32_t &lo,
                                         uint32_t &hi) {
	const uint32_t a0 = c0.x, b0 = c0.y, a1 = c0.w, b1 = c1.x, a2 = c1.z, b2 = c1.w;
	const uint32_t Xs = X0 + ((uint32_t)x0 << 4), Ys = Y0 + ((uint32_t)y0 << 4);
	uint32_t e0 = a0 * Xs + b0 * Ys + c0.z;
	uint32_t e1 = a1 * Xs + b1 * Ys + c1.y;
	uint32_t e2 = a2 * Xs + b2 * Ys + c2x;
	const uint32_t sx0 = a0 << 4, sx1 = a1 << 4, sx2 = a2 << 4;
	const uint32_t sy0 = b0 << 4, sy1 = b1 << 4, sy2 = b2 << 4;
	unsigned long long m = 0ull;
	for(int y = y0; y <= y1; ++y) {
		uint32_t r0 = e0, r1 = e1, r2 = e2;
		uint32_t row = 0u;
		for(int x = x0; x <= x1; ++x) {
			row |= ((int)(r0 | r1 | r2) > 0) ? (1u << x) : 0u;
			r0 += sx0;
			r1 += sx1;
			r2 += sx2;
		}
		m |= (unsigned long long)row << (8 * y);
		e0 += sy0;
		e1 += sy1;
		e2 += sy2;
	}
	lo = (uint32_t)m;
	hi = (uint32_t)(m >> 32);
}

// 32x32 bit-matrix transpose across the warp: in: lane r holds row r (bit c = M[r][c]); out: lane r holds column r.
__device__ __forceinline__ uint32_t warp_transpose_bits(uint32_t x) {
	const uint32_t lane = lane_id();
#pragma unroll
	for(int k = 16; k >= 1; k >>= 1) {
		const uint32_t mask = (k == 16) ? 0x0000ffffu : (k == 8) ? 0x00ff00ffu : (k == 4) ? 0x0f0f0f0fu : (k == 2) ? 0x33333333u : 0x55555555u;
		const uint32_t y = __shfl_xor_sync(0xffffffffu, x, k);
		x = (lane & k) ? ((x & ~mask) | ((y >> k) & mask)) : ((x & mask) | ((y << k) & ~mask));
	}
	return x;
}

// barycentrics of one pixel (main.c:1089-1102)
__device__ __forceinline__ void barycentrics(uint32_t E1, uint32_t E2, float ooa, float &bx, float &by) {
	bx = (float)((int)E1 >> 8) * ooa;
	by = (float)((int)E2 >> 8) * ooa;
}

// attribute interpolation of one component (main.c:1131-1132)
__device__ __forceinline__ float interp(float v0, float v1, float v2, float u, float v) {
	float t = v0 + (v1 - v0) * u;
	t = t + (v2 - v0) * v;
	return t;
}

// perspective-correct barycentrics of the pixel whose edge functions are E1, E2 (main.c:1089-1115)
__device__ __forceinline__ void perspective_barycentrics(uint32_t E1, uint32_t E2, float ooa, const float4 &s1, float &pbx, float &pby) {
	float bx, by;
	barycentrics(E1, E2, ooa, bx, by);
	float denom = 1.0f - (bx + by);
	denom = denom * s1.x;
	denom = denom + bx * s1.y;
	denom = denom + by * s1.z;
	denom = 1.0f / denom;
	pbx = (bx * s1.y) * denom;
	pby = (by * s1.z) * denom;
}

#define MLV_PS_ID_BASIC_TRILINEAR 3

template <int PS>
__device__ __forceinline__ uint32_t shade_pixel(const TailParams &P, uint32_t slot, uint32_t X, uint32_t Y) {
	const uint4 *cov = cov_of(P, slot);
	const uint4 c0 = __ldg(cov), c1 = __ldg(cov + 1);
	const uint32_t c2x = __ldg(reinterpret_cast<const uint32_t *>(cov + 2));
	const uint32_t E1 = c0.w * X + c1.x * Y + c1.y;
	const uint32_t E2 = c1.z * X + c1.w * Y + c2x;
	const float4 *sh = shade_of(P, slot);
	const float4 s0 = __ldg(sh), s1 = __ldg(sh + 1), r1a = __ldg(sh + 2), r1b = __ldg(sh + 3), r1c = __ldg(sh + 4), s5 = __ldg(sh + 5);
	float pbx, pby;
	perspective_barycentrics(E1, E2, s0.x, s1, pbx, pby);
	if(PS == MLV_PS_ID_BASIC_TRILINEAR) {
		// Screen-space derivatives of UV, analytically: the same triangle's interpolation one pixel to the right
		// (E + 16a) and one pixel down (E + 16b) -- exact plane extrapolation, also outside the triangle.
		float qx, qy, rx, ry;
		perspective_barycentrics(E1 + (c0.w << 4), E2 + (c1.z << 4), s0.x, s1, qx, qy);
		perspective_barycentrics(E1 + (c1.x << 4), E2 + (c1.w << 4), s0.x, s1, rx, ry);
		const float u = interp(r1a.w, r1b.w, r1c.w, pbx, pby), v = interp(s1.w, s5.x, s5.y, pbx, pby);
		return encode_color(run_ps_basic_trilinear(u, v, interp(r1a.w, r1b.w, r1c.w, qx, qy), interp(s1.w, s5.x, s5.y, qx, qy), interp(r1a.w, r1b.w, r1c.w, rx, ry),
		                                           interp(s1.w, s5.x, s5.y, rx, ry), P.ps_tex));
	}
	const float4 r1 = make_float4(interp(r1a.x, r1b.x, r1c.x, pbx, pby), interp(r1a.y, r1b.y, r1c.y, pbx, pby), interp(r1a.z, r1b.z, r1c.z, pbx, pby),
	                              interp(r1a.w, r1b.w, r1c.w, pbx, pby));
	const float r2x = interp(s1.w, s5.x, s5.y, pbx, pby);
	return encode_color(run_ps<(PS == MLV_PS_ID_BASIC_TRILINEAR ? 1 : PS)>(r1, r2x, P.ps_tex, P.rsqrt_lut));
}


// Coverage by (triangle, row) items. With lanes over triangles (coverage() above) a batch costs what its TALLEST bounding
// box costs and the lanes beyond the list's length idle -- at BASELINE config 5 (21 pixel-sized triangles per tile, boxes of
// ~4 x 5 pixels) that loop was a third of k_tile's instructions. Here every lane publishes its triangle's edge values at the
// top-left pixel of (bounds & tile) and the steps, a warp scan numbers the rows of all boxes, and lane i evaluates row item
// i: same wrapping arithmetic, same bits, ~3 rounds of one row each instead of ~8 rows per lane. Row masks are assembled as
// bytes of the reference's 64-bit fragment mask (bit 8*row + col) in shared memory. Batches of tall boxes (more than
// MLV_ROWCOV_MAX_ITEMS rows in all) keep the per-triangle loop, which has less overhead per row.
#define MLV_ROWCOV_MAX_ITEMS 160
struct RowCovWarp {
	uint32_t e[3][32];      // edge functions at the box's top-left pixel
	uint32_t sx[3][32];     // step per pixel to the right (a << 4)
	uint32_t sy[3][32];     // step per row down (b << 4)
	uint32_t box[32];       // x0 | x1 << 4 | y0 << 8 (tile-relative)
	uint32_t first[32];     // number of the triangle's first row item
	uint2 mask[32];         // fragment mask under construction: .x rows 0-3, .y rows 4-7
	uint8_t item_tri[256];  // row item -> lane of its triangle
};

template <int PS>
__device__ __forceinline__ void tile_phase(const TailParams &P, RowCovWarp *s_rowcov, const uint32_t n_cbins) {
	const uint32_t lane = lane_id();
	const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
	const uint32_t px = lane & 7u, py = lane >> 3; // this lane's pixels: (px, py) and (px, py + 4)

	for(uint32_t cb = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; cb < n_cbins; cb += warps) {
		const mlv_ref_compacted_bin bin = P.cbins[cb];
		const uint32_t n = bin.num_triangles_self, off = bin.num_triangles_upto, b = bin.bin_index;
		const int tile_x = (int)(b % (uint32_t)P.wt) * 8, tile_y = (int)(b / (uint32_t)P.wt) * 8;
		const uint32_t X0 = (uint32_t)tile_x << 4, Y0 = (uint32_t)tile_y << 4;
		const uint32_t X = X0 + (px << 4), Y = Y0 + (py << 4);

		// read_tile (main.c:577-587)
		uint4 pix = P.fb[(size_t)b * 32u + lane];
		float d0 = __uint_as_float(pix.z), d1 = __uint_as_float(pix.w);
		uint32_t win0 = MLV_NO_WINNER, win1 = MLV_NO_WINNER; // keys of the fragments that hold the pixels (depth ties go to the greater key)
		uint32_t wslot0 = 0u, wslot1 = 0u;
		const float tile_min_old = P.tile_min[b]; // get_tile_minimum_depth: previous draws only (N3)

		if(n == 0u) { // touched bin whose pairs were all Hi-Z-rejected at binning time: write_tile's tile-minimum refresh only
			float m = ref_min_macro(ref_min_macro(1.0f, d0), d1);
#pragma unroll
			for(int d = 16; d > 0; d >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, d));
			if(lane == 0) P.tile_min[b] = m;
			continue;
		}
		// The reference walks a bin's list in ascending triangle id and lets a fragment through when z >= depth
		// (main.c:1166), so the surviving fragment of a pixel is the one with the greatest z and, among equal z, the
		// greatest id -- a property of the SET of fragments. Taking the maximum over (z, key) therefore gives the
		// reference's result in any visiting order, and the list the fill phase left in arrival order needs no sorting.
		// Debug capture sorts it anyway so that mlv_debug_read_bins shows the reference's per-tile order.
		uint32_t *ids = P.pair_ids + off;
		if(P.sort_lists) sort_bin_ids(P, ids, P.pair_tmp + off, n, P.key_bits);

		for(uint32_t base = 0; base < n; base += 32u) {
			// ---- rasterizer, lanes over triangles (main.c:996-1041)
			const uint32_t k = base + lane;
			uint32_t key = 0, lo = 0, hi = 0;
			uint32_t a1 = 0, b1 = 0, e1 = 0, a2 = 0, b2 = 0, e2 = 0; // E1/E2 of this lane's triangle: coefficients and value at the tile origin
			float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f);
			uint4 c0 = make_uint4(0u, 0u, 0u, 0u), c1 = c0;
			uint32_t c2x = 0u, slot = 0u;
			int x0 = 0, x1 = -1, y0 = 0, y1 = -1; // (bounds & tile) of a triangle that passes Hi-Z; empty otherwise
			if(k < n) {
				slot = ids[k];
				key = key_of(P, slot);
				const uint4 *cov = cov_of(P, slot);
				const uint4 c2 = __ldg(cov + 2);
				const float max_depth = __uint_as_float(c2.y);
				if(!(max_depth < tile_min_old)) { // Hi-Z (main.c:1005-1010)
					c0 = __ldg(cov), c1 = __ldg(cov + 1), c2x = c2.x;
					x0 = 0, x1 = 7, y0 = 0, y1 = 7;
					if(c2.z & MLV_NOWRAP_BIT) {
						x0 = max((int)(c2.z & 0xffffu) - tile_x, 0);
						y0 = max((int)((c2.z >> 16) & 0x7fffu) - tile_y, 0);
						x1 = min((int)(short)(c2.w & 0xffffu) - tile_x, 7);
						y1 = min((int)(short)(c2.w >> 16) - tile_y, 7);
					}
				}
			}
			{
				const uint32_t rows = (x1 >= x0 && y1 >= y0) ? (uint32_t)(y1 - y0 + 1) : 0u;
				uint32_t incl = rows;
#pragma unroll
				for(int d = 1; d < 32; d <<= 1) {
					const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
					if(lane >= (uint32_t)d) incl += o;
				}
				const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
				if(total > MLV_ROWCOV_MAX_ITEMS) { // tall boxes: per-triangle loop
					if(rows) coverage(c0, c1, c2x, X0, Y0, x0, x1, y0, y1, lo, hi);
				} else if(total) {
					RowCovWarp &R = s_rowcov[threadIdx.x >> 5];
					const uint32_t first = incl - rows;
					if(rows) {
						const uint32_t Xs = X0 + ((uint32_t)x0 << 4), Ys = Y0 + ((uint32_t)y0 << 4);
						R.e[0][lane] = c0.x * Xs + c0.y * Ys + c0.z, R.sx[0][lane] = c0.x << 4, R.sy[0][lane] = c0.y << 4;
						R.e[1][lane] = c0.w * Xs + c1.x * Ys + c1.y, R.sx[1][lane] = c0.w << 4, R.sy[1][lane] = c1.x << 4;
						R.e[2][lane] = c1.z * Xs + c1.w * Ys + c2x, R.sx[2][lane] = c1.z << 4, R.sy[2][lane] = c1.w << 4;
						R.box[lane] = (uint32_t)x0 | ((uint32_t)x1 << 4) | ((uint32_t)y0 << 8);
						for(uint32_t r = 0; r < rows; ++r) R.item_tri[first + r] = (uint8_t)lane;
					}
					R.first[lane] = first;
					R.mask[lane] = make_uint2(0u, 0u);
					__syncwarp();
					for(uint32_t it = 0; it < total; it += 32u) {
						const uint32_t i = it + lane;
						if(i < total) {
							const uint32_t j = R.item_tri[i];
							const uint32_t r = i - R.first[j], box = R.box[j];
							const uint32_t bx0 = box & 15u, bx1 = (box >> 4) & 15u, by = ((box >> 8) & 15u) + r;
							const uint32_t sx0 = R.sx[0][j], sx1 = R.sx[1][j], sx2 = R.sx[2][j];
							uint32_t r0 = R.e[0][j] + r * R.sy[0][j], r1 = R.e[1][j] + r * R.sy[1][j], r2 = R.e[2][j] + r * R.sy[2][j];
							uint32_t row = 0u;
							for(uint32_t x = bx0; x <= bx1; ++x) {
								row |= ((int)(r0 | r1 | r2) > 0) ? (1u << x) : 0u;
								r0 += sx0;
								r1 += sx1;
								r2 += sx2;
							}
							reinterpret_cast<uint8_t *>(&R.mask[j])[by] = (uint8_t)row;
						}
					}
					__syncwarp();
					const uint2 m = R.mask[lane];
					lo = m.x, hi = m.y;
					__syncwarp(); // the next batch rewrites the warp's rows
				}
			}
			if(k < n) {
				if(lo | hi) {
					a1 = c0.w, b1 = c1.x, e1 = c0.w * X0 + c1.x * Y0 + c1.y;
					a2 = c1.z, b2 = c1.w, e2 = c1.z * X0 + c1.w * Y0 + c2x;
					s0 = __ldg(shade_of(P, slot));
				}
				if(P.dbg.infos) {
					mlv_ref_tile_info ti;
					ti.triangle_id = key;
					ti._pad = 0;
					ti.fragment_mask = ((unsigned long long)hi << 32) | lo;
					P.dbg.infos[off + k] = ti;
				}
			}
			if(!__any_sync(0xffffffffu, (lo | hi) != 0u)) continue;
			// ---- per-pixel cover sets: bit j of set0/set1 <=> triangle (base + j) covers this lane's pixel 0/1
			const uint32_t set0 = warp_transpose_bits(lo), set1 = warp_transpose_bits(hi);
			// ---- early-Z in list order, lanes over pixels (main.c:1060-1168)
			uint32_t todo = set0 | set1;
			while(__any_sync(0xffffffffu, todo != 0u)) {
				const int j = todo ? (__ffs(todo) - 1) : 0;
				const uint32_t bitj = todo ? (1u << j) : 0u; // lanes that are done only take part in the shuffles
				todo &= ~bitj;
				const uint32_t ta1 = __shfl_sync(0xffffffffu, a1, j), tb1 = __shfl_sync(0xffffffffu, b1, j), te1 = __shfl_sync(0xffffffffu, e1, j);
				const uint32_t ta2 = __shfl_sync(0xffffffffu, a2, j), tb2 = __shfl_sync(0xffffffffu, b2, j), te2 = __shfl_sync(0xffffffffu, e2, j);
				const float ooa = __shfl_sync(0xffffffffu, s0.x, j), z0 = __shfl_sync(0xffffffffu, s0.y, j);
				const float z1 = __shfl_sync(0xffffffffu, s0.z, j), z2 = __shfl_sync(0xffffffffu, s0.w, j);
				const uint32_t tkey = __shfl_sync(0xffffffffu, key, j), tslot = __shfl_sync(0xffffffffu, slot, j);
				const uint32_t E1 = ta1 * (px << 4) + tb1 * (py << 4) + te1;
				const uint32_t E2 = ta2 * (px << 4) + tb2 * (py << 4) + te2;
				if(set0 & bitj) {
					float bx, by;
					barycentrics(E1, E2, ooa, bx, by);
					const float z = interp(z0, z1, z2, bx, by);
					if(z > d0 || (z == d0 && (win0 == MLV_NO_WINNER || tkey > win0))) { // z >= depth in id order (main.c:1166), see above; false on NaN
						d0 = z;
						win0 = tkey;
						wslot0 = tslot;
					}
				}
				if(set1 & bitj) {
					float bx, by;
					barycentrics(E1 + (tb1 << 6), E2 + (tb2 << 6), ooa, bx, by); // four rows down: + b*(4*16)
					const float z = interp(z0, z1, z2, bx, by);
					if(z > d1 || (z == d1 && (win1 == MLV_NO_WINNER || tkey > win1))) {
						d1 = z;
						win1 = tkey;
						wslot1 = tslot;
					}
				}
			}
		}

		// ---- pixel shader + output merger, once per pixel on the last fragment that passed (main.c:1170-1181)
		if(win0 != MLV_NO_WINNER) pix.x = shade_pixel<PS>(P, wslot0, X, Y);
		if(win1 != MLV_NO_WINNER) pix.y = shade_pixel<PS>(P, wslot1, X, Y + 64u);
		pix.z = __float_as_uint(d0);
		pix.w = __float_as_uint(d1);

		// write_tile (main.c:589-603): every non-empty bin is written and refreshes its tile minimum (N4)
		P.fb[(size_t)b * 32u + lane] = pix;
		float m = ref_min_macro(ref_min_macro(1.0f, d0), d1);
#pragma unroll
		for(int d = 16; d > 0; d >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, d));
		if(lane == 0) P.tile_min[b] = m;
	}
}

// One warp of the last CTA of a draw's last kernel: fold the draw's Stats contributions (main.c:1228-1246), re-arm the
// counters and the draw context for the next draw.
__device__ __forceinline__ void finish_draw(const TailParams &P, bool skipped, bool pair_overflow) {
	const uint32_t lane = lane_id();
	uint32_t tris = 0, pairs = 0, records = 0;
#pragma unroll
	for(uint32_t i = lane; i < MLV_STAT_STRIPES; i += 32u) {
		const unsigned long long v = P.stat_stripes[i * 16u];
		P.stat_stripes[i * 16u] = 0ull;
		tris += (uint32_t)(v >> 32);
		pairs += (uint32_t)v;
		records += (uint32_t)P.stat_stripes[i * 16u + 1u];
		P.stat_stripes[i * 16u + 1u] = 0ull;
	}
#pragma unroll
	for(int d = 16; d > 0; d >>= 1) {
		tris += __shfl_xor_sync(0xffffffffu, tris, d);
		pairs += __shfl_xor_sync(0xffffffffu, pairs, d);
		records += __shfl_xor_sync(0xffffffffu, records, d);
	}
	// the epoch tags this draw's touches in bin_touch and the look-back words of its scan; when its 30 bits wrap (after 2^30
	// draws) every tag is invalidated
	uint32_t next_epoch = 0u;
	if(lane == 0) next_epoch = (P.ctr->epoch + 1u) & 0x3fffffffu;
	next_epoch = __shfl_sync(0xffffffffu, next_epoch, 0);
	if(next_epoch == 0u) {
		for(uint32_t i = lane; i < 2u * P.scan_blocks; i += 32u) P.state_sum[i] = 0ull; // (state_nz follows state_sum)
		next_epoch = 1u;
	}
	if(lane == 0) {
		Counters *c = P.ctr;
		c->work.records_written += records;
		if(!skipped) {
			c->work.pairs_listed += c->pair_total;
			c->work.tiles_visited += c->n_cbins;
		}
		if(pair_overflow) atomicOr(&c->error_flags, MLV_FLAG_PAIR_OVERFLOW);
		// what the read-backs of the last draw's lists see (mlv_read_bin_lists, mlv_debug_read_bins)
		c->last_pair_total = skipped ? 0xffffffffu : c->pair_total;
		c->last_n_cbins = skipped ? 0u : c->n_cbins;
		c->stats.vertex_count += P.index_count;             // main.c:1228-1232
		c->stats.input_triangle_count += P.direct_slots;
		c->stats.assembled_triangle_count += tris;
		c->stats.total_triangle_count_in_bins += pairs;
		c->stats.active_bin_count += c->draw_active_bins;   // non-empty bins in the reference's sense: Hi-Z-rejected pairs count (main.c:1245)
		c->draw_active_bins = 0u;
		c->last_ovf_count = min(P.dctr->ovf_count, P.ovf_capacity);
		c->pair_total = c->n_cbins = c->ticket = c->n_wake = c->draw_alive = 0u;
		c->epoch = next_epoch;
		P.dctr->ovf_count = P.dctr->clip_count = P.dctr->big_count = P.dctr->huge_count = 0u; // the draw context is free for the front half of a later draw
	}
}

template <int PS>
__global__ void __launch_bounds__(MLV_TILE_THREADS, 4) k_tile(const __grid_constant__ TailParams P) {
	__shared__ RowCovWarp s_rowcov[MLV_TILE_THREADS / 32];
	MLV_KERNEL_PROLOGUE(P.timeline ? P.timeline + 8 : nullptr);
	Counters *c = P.ctr;
	const uint32_t total = c->pair_total, n_cbins = c->n_cbins;
	const bool skipped = total > P.pair_capacity || P.dctr->ovf_count > P.ovf_capacity; // MLV_FLAG_PAIR_OVERFLOW / MLV_FLAG_TRI_OVERFLOW: the draw is skipped as a whole
	// Stats: the bins this draw touched (every thread of the grid looks at its share of this rank's band of bins: one load,
	// in flight with the ones above). Tried: letting CTA 0 alone finish a draw without surviving pairs (no 592 arrival atomics):
	// 3.5 us instead of 4 for one rank of 8, but 12.6 us on one GPU, where the map is 130 KB -- the distributed form stays.
	const uint32_t word_begin = P.bin_begin / 4u, word_end = (min(P.bin_end, P.num_bins) + 3u) / 4u; // (the map is padded to a multiple of 4 bytes)
	uint32_t touched = 0;
	for(uint32_t w = word_begin + blockIdx.x * blockDim.x + threadIdx.x; w < word_end; w += gridDim.x * blockDim.x) {
		uint32_t *word = reinterpret_cast<uint32_t *>(P.touch_bits) + w;
		const uint32_t bytes = __ldcg(word);
		if(bytes) {
			touched += (uint32_t)__popc(bytes & 0x01010101u);
			*word = 0u; // the context's map is clean again for the front half of a later draw
		}
	}
	if(!skipped && n_cbins) tile_phase<PS>(P, s_rowcov, n_cbins);
	// The LAST CTA to get here folds the draw's Stats and re-arms the per-draw counters: CTAs of this grid that start late
	// (the grid may exceed what is resident at once) must still find pair_total / n_cbins as the scan left them. One 64-bit
	// atomic per CTA: arrivals in the low word, touched bins in the high word.
	__shared__ uint32_t s_touched[MLV_TILE_THREADS / 32];
	__shared__ bool s_last;
#pragma unroll
	for(int d = 16; d > 0; d >>= 1) touched += __shfl_xor_sync(0xffffffffu, touched, d);
	if(lane_id() == 0) s_touched[threadIdx.x >> 5] = touched;
	__syncthreads();
	if(threadIdx.x == 0) {
		touched = 0;
		for(uint32_t w = 0; w < blockDim.x / 32u; ++w) touched += s_touched[w];
		__threadfence();
		const unsigned long long before = atomicAdd(&c->tile_done, ((unsigned long long)touched << 32) | 1ull);
		s_last = (uint32_t)before == gridDim.x - 1u;
		if(s_last) {
			c->draw_active_bins = (uint32_t)(before >> 32) + touched;
			c->tile_done = 0ull;
		}
	}
	__syncthreads();
	if(s_last && threadIdx.x < 32) finish_draw(P, skipped, total > P.pair_capacity);
}

// =================================================================================================
// resolve / composite
// =================================================================================================

// load_texture's sRGB branch (main.c:546-558) as a byte transform over the texels. The curve has 256 possible inputs per
// channel; the table is evaluated on the host with the reference's own double arithmetic (libm pow) and travels as a
// kernel parameter, so the kernel is pure HBM work: 16 texels per thread-iteration through 128-bit accesses.
struct SrgbTable {
	uint8_t lin[256];
};
__global__ void __launch_bounds__(256) k_texture_srgb_to_linear(uint4 *__restrict__ texels, size_t count_u4, uint32_t *__restrict__ tail, uint32_t tail_count, const __grid_constant__ SrgbTable T) {
	pdl_prologue();
	__shared__ uint8_t s_lin[256];
	s_lin[threadIdx.x] = T.lin[threadIdx.x];
	__syncthreads();
	auto conv = [&](uint32_t t) { return (uint32_t)s_lin[t & 0xffu] | ((uint32_t)s_lin[(t >> 8) & 0xffu] << 8) | ((uint32_t)s_lin[(t >> 16) & 0xffu] << 16) | ((uint32_t)s_lin[t >> 24] << 24); };
	for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count_u4; i += (size_t)gridDim.x * blockDim.x) {
		uint4 v = texels[i];
		v.x = conv(v.x), v.y = conv(v.y), v.z = conv(v.z), v.w = conv(v.w);
		texels[i] = v;
	}
	if(blockIdx.x == 0 && threadIdx.x < tail_count) tail[threadIdx.x] = conv(tail[threadIdx.x]);
}

// One level of the mip chain of an R8G8B8A8 texture (mlv_texture_generate_mips): 2x2 box filter per channel with
// round-to-nearest ((a+b+c+d+2)>>2); an odd source extent repeats its last row / column. One thread per destination texel.
__global__ void __launch_bounds__(256) k_mip_downsample(const uint32_t *__restrict__ src, int sw, int sh, uint32_t *__restrict__ dst, int dw, int dh) {
	pdl_prologue();
	const size_t n = (size_t)dw * (size_t)dh;
	for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		const int x = (int)(i % (size_t)dw), y = (int)(i / (size_t)dw);
		const int x0 = min(2 * x, sw - 1), x1 = min(2 * x + 1, sw - 1), y0 = min(2 * y, sh - 1), y1 = min(2 * y + 1, sh - 1);
		const uint32_t a = __ldg(src + (size_t)y0 * sw + x0), b = __ldg(src + (size_t)y0 * sw + x1), c = __ldg(src + (size_t)y1 * sw + x0), d = __ldg(src + (size_t)y1 * sw + x1);
		uint32_t out = 0;
#pragma unroll
		for(int sft = 0; sft < 32; sft += 8) out |= ((((a >> sft) & 0xffu) + ((b >> sft) & 0xffu) + ((c >> sft) & 0xffu) + ((d >> sft) & 0xffu) + 2u) >> 2) << sft;
		dst[i] = out;
	}
}

// Tiled -> row-major. One thread per 4 horizontally adjacent pixels of rows y and y+4 of a tile: four 128-bit
// loads (64 contiguous bytes, all of them used), two 128-bit colour stores (+ two 128-bit depth stores).
struct Quad8 {
	uint4 c_top, c_bot;
	float4 d_top, d_bot;
};
__device__ __forceinline__ Quad8 load_quad8(const uint4 *__restrict__ fb, uint32_t bin, uint32_t row, uint32_t half) {
	const uint4 *p = fb + (size_t)bin * 32u + row * 8u + half * 4u; // lanes (x = 4*half .. 4*half+3, y = row)
	const uint4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
	Quad8 q;
	q.c_top = make_uint4(a.x, b.x, c.x, d.x);
	q.c_bot = make_uint4(a.y, b.y, c.y, d.y);
	q.d_top = make_float4(__uint_as_float(a.z), __uint_as_float(b.z), __uint_as_float(c.z), __uint_as_float(d.z));
	q.d_bot = make_float4(__uint_as_float(a.w), __uint_as_float(b.w), __uint_as_float(c.w), __uint_as_float(d.w));
	return q;
}

// work item w -> (bin, row 0..3, half 0..1); 8 items per tile
__global__ void __launch_bounds__(256) k_resolve(const uint4 *__restrict__ fb, uint4 *__restrict__ colors, float4 *__restrict__ depths, int width, int height) {
	pdl_prologue();
	const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t wt = (uint32_t)width >> 3, ht = (uint32_t)height >> 3;
	// order work so that consecutive threads write consecutive 16-byte quads of one image row
	const uint32_t quads_per_row = wt * 2u;
	const uint32_t rowgroup = w / quads_per_row; // = ty * 4 + row
	if(rowgroup >= ht * 4u) return;
	const uint32_t xq = w % quads_per_row, ty = rowgroup >> 2, row = rowgroup & 3u;
	const uint32_t bin = ty * wt + (xq >> 1);
	const Quad8 q = load_quad8(fb, bin, row, xq & 1u);
	const size_t top = ((size_t)(ty * 8u + row) * (uint32_t)width) / 4u + xq, bot = top + (size_t)width; // +4 rows = 4*width/4 quads
	colors[top] = q.c_top;
	colors[bot] = q.c_bot;
	if(depths) {
		depths[top] = q.d_top;
		depths[bot] = q.d_bot;
	}
}

// Sort-first compositing helpers. Chunk r of the gather buffer = row-major colour of rank r's
// stripes in ascending stripe order. PACK: tiled framebuffer of this rank -> its own chunk.
// UNPACK: every chunk -> the final row-major image.
__device__ __forceinline__ uint32_t chunk_row(uint32_t ty, uint32_t y_in_tile, int stripe_h, int num_ranks) {
	const uint32_t stripe = ty / (uint32_t)stripe_h;
	const uint32_t local_stripe = stripe / (uint32_t)num_ranks;
	return (local_stripe * (uint32_t)stripe_h + (ty % (uint32_t)stripe_h)) * 8u + y_in_tile;
}

__global__ void __launch_bounds__(256) k_composite_pack(const uint4 *__restrict__ fb, uint4 *__restrict__ chunk, int width, int height, Partition part) {
	pdl_prologue();
	const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t wt = (uint32_t)width >> 3, ht = (uint32_t)height >> 3;
	const uint32_t quads_per_row = wt * 2u;
	const uint32_t rowgroup = w / quads_per_row;
	if(rowgroup >= ht * 4u) return;
	const uint32_t xq = w % quads_per_row, ty = rowgroup >> 2, row = rowgroup & 3u;
	if(!part.owns_row((int)ty)) return;
	const Quad8 q = load_quad8(fb, ty * wt + (xq >> 1), row, xq & 1u);
	const size_t top = (size_t)chunk_row(ty, row, part.stripe_h, part.num_ranks) * quads_per_row + xq;
	chunk[top] = q.c_top;
	chunk[top + 4u * quads_per_row] = q.c_bot;
}

__global__ void __launch_bounds__(256) k_composite_unpack(const uint4 *__restrict__ gather, uint4 *__restrict__ colors, int width, int height, int num_ranks, int stripe_h,
                                                          size_t chunk_u4) {
	pdl_prologue();
	const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t qw = (uint32_t)width >> 2;
	if(q >= qw * (uint32_t)height) return;
	const uint32_t y = q / qw, xq = q % qw;
	const uint32_t ty = y >> 3;
	const uint32_t owner = (ty / (uint32_t)stripe_h) % (uint32_t)num_ranks;
	colors[q] = __ldg(gather + (size_t)owner * chunk_u4 + (size_t)chunk_row(ty, y & 7u, stripe_h, num_ranks) * qw + xq);
}

// Fused resolve + all-gather over peer memory: this rank's tiles go from the tiled framebuffer straight into the
// row-major image of EVERY rank (its own included) with 128-bit stores -- local HBM for itself, NVLink for the
// peers -- so there is no staging chunk, no collective call and no un-swizzle pass afterwards. When the last CTA
// has fenced its stores, one thread publishes the frame's sequence number in every rank's arrival word for this rank.
// Targets are the ranks [target_lo, target_hi); `publish` = 0 leaves the arrival words alone (the copy-engine form of
// the exchange resolves into this rank's own image only and publishes after its peer copies, k_composite_publish).
__global__ void __launch_bounds__(256) k_composite_broadcast(const uint4 *__restrict__ fb, const __grid_constant__ PeerTargets peers, int width, int height, Partition part, uint32_t seq,
                                                             Counters *__restrict__ ctr, int target_lo, int target_hi, int publish) {
	pdl_prologue();
	const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t wt = (uint32_t)width >> 3, ht = (uint32_t)height >> 3;
	const uint32_t quads_per_row = wt * 2u;
	const uint32_t rowgroup = w / quads_per_row;
	if(rowgroup < ht * 4u) {
		const uint32_t xq = w % quads_per_row, ty = rowgroup >> 2, row = rowgroup & 3u;
		if(part.owns_row((int)ty)) {
			const Quad8 q = load_quad8(fb, ty * wt + (xq >> 1), row, xq & 1u);
			const size_t top = (size_t)(ty * 8u + row) * quads_per_row + xq;
			for(int p = target_lo; p < target_hi; ++p) {
				uint4 *dst = peers.color[p];
				dst[top] = q.c_top;
				dst[top + 4u * quads_per_row] = q.c_bot;
			}
		}
	}
	if(!publish) return;
	__threadfence_system(); // my stores are visible to every GPU before my CTA is counted as done
	__syncthreads();
	if(threadIdx.x == 0) {
		const uint32_t done = atomicAdd(&ctr->bcast_done, 1u);
		if(done == gridDim.x - 1u) {
			ctr->bcast_done = 0u;
			__threadfence_system();
			for(int p = 0; p < part.num_ranks; ++p) *reinterpret_cast<volatile uint32_t *>(peers.flags[p] + part.rank) = seq;
		}
	}
}

// Copy-engine form of the exchange: after this rank's band has been copied into every peer's image (cudaMemcpyAsync on
// the exchange stream, ordered before this launch), tell every rank that frame `seq` of this rank has arrived.
__global__ void __launch_bounds__(32) k_composite_publish(const __grid_constant__ PeerTargets peers, int num_ranks, int rank, uint32_t seq) {
	pdl_prologue();
	if((int)threadIdx.x < num_ranks) {
		__threadfence_system();
		*reinterpret_cast<volatile uint32_t *>(peers.flags[threadIdx.x] + rank) = seq;
	}
}

// Waits until every rank's stripes of frame `seq` have arrived in this rank's image. Bounded: a peer that never
// arrives sets MLV_FLAG_COMPOSITE_TIMEOUT instead of hanging the GPU.
__global__ void __launch_bounds__(32) k_composite_wait(const uint32_t *flags, int num_ranks, uint32_t seq, Counters *__restrict__ ctr, unsigned long long timeout_ns) {
	pdl_prologue();
	if((int)threadIdx.x < num_ranks) {
		const volatile uint32_t *f = flags + threadIdx.x;
		unsigned long long t0;
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
		while((int)(*f - seq) < 0) {
			__nanosleep(100);
			unsigned long long t1;
			asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
			if(t1 - t0 > timeout_ns) {
				atomicOr(&ctr->error_flags, MLV_FLAG_COMPOSITE_TIMEOUT);
				break;
			}
		}
	}
	__threadfence_system();
}

} // namespace mlv
