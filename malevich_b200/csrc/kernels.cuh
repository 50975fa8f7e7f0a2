// kernels.cuh -- the draw pipeline as hand-written sm_100a kernels.
//
//   k_clear      clear_render_target_view + clear_depth_stencil_view      (reference main.c:1191-1217)
//   k_geom<VS>   input assembler + vertex shader + primitive assembly     (main.c:662-913)
//                fused, one thread per input triangle, ordered single-pass emission
//                (decoupled look-back scan) so that assembled-triangle ids equal the
//                reference's single-thread order
//   k_bin<FILL>  binner passes 1 and 2                                    (main.c:924-962)
//   k_bin_scan   binner exclusive scan + compaction of non-empty bins     (main.c:937-974)
//   k_tile<PS>   rasterizer + Hi-Z + early-Z + pixel shader + output merger (main.c:983-1189)
//                one warp per non-empty bin, lanes over triangles for coverage (64-bit masks),
//                lanes over pixels for depth, pixel shader run once per pixel on the winning
//                fragment (bit-identical to in-order shading: see DESIGN.md "deferred shading")
//   k_resolve    tiled -> row-major, 128-bit stores                       (replaces GDI blit main.c:314)
#pragma once
#include "mlv_internal.cuh"

namespace mlv {

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// =================================================================================================
// clear
// =================================================================================================
// mode bit 0: colour, bit 1: depth (+ tile minima := 0, main.c:1212-1214)
__global__ void __launch_bounds__(256) k_clear(uint4 *__restrict__ fb, float *__restrict__ tile_min, uint32_t num_bins, uint32_t color, float depth, int mode) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= num_bins * 32u) return;
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
	if((mode & 2) && (i & 31u) == 0) tile_min[i >> 5] = 0.0f;
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
__device__ __forceinline__ int snap(float v) {
	const double d = floor((double)(v * 16.0f) + 0.5);
	return __double2int_rz(d);
}

// Projection, viewport transform, snapping, triangle setup for one (sub-)triangle (main.c:797-898).
// Returns false when the triangle is back-face culled (signed_area > 0, main.c:856).
__device__ __forceinline__ bool setup_triangle(const float4 &c0, const float4 &c1, const float4 &c2, const GeomParams &P, TriSetup &S) {
	float4 p[3] = { c0, c1, c2 };
#pragma unroll
	for(int i = 0; i < 3; ++i) {
		// a_reciprocal_ws[i] = 1.0 / w  (double divide rounded to f32 == correctly rounded f32 divide)
		const float rw = 1.0f / p[i].w;
		S.rw[i] = rw;
		p[i].x *= rw;
		p[i].y *= rw;
		p[i].z *= rw;
		p[i].w *= rw;
		// m4x4f32_mul_v4f32(&screen_from_ndc, v): serial dot products with the literal zero entries
		const float4 v = p[i];
		float4 s;
		s.x = P.vp_m00 * v.x + 0.0f * v.y + 0.0f * v.z + P.vp_m03 * v.w;
		s.y = 0.0f * v.x + P.vp_m11 * v.y + 0.0f * v.z + P.vp_m13 * v.w;
		s.z = 0.0f * v.x + 0.0f * v.y + P.vp_m22 * v.z + P.vp_m23 * v.w;
		s.w = 0.0f * v.x + 0.0f * v.y + 0.0f * v.z + 1.0f * v.w;
		S.p[i] = s;
	}
	const int x0 = snap(S.p[0].x), x1 = snap(S.p[1].x), x2 = snap(S.p[2].x);
	const int y0 = snap(S.p[0].y), y1 = snap(S.p[1].y), y2 = snap(S.p[2].y);
	const int signed_area =
	    (int)(((uint32_t)x1 - (uint32_t)x0) * ((uint32_t)y2 - (uint32_t)y0) - ((uint32_t)x2 - (uint32_t)x0) * ((uint32_t)y1 - (uint32_t)y0));
	if(signed_area > 0) return false;
	set_edge(S.e + 6, signed_area, x0, y0, x1, y1);
	set_edge(S.e + 0, signed_area, x1, y1, x2, y2);
	set_edge(S.e + 3, signed_area, x2, y2, x0, y0);
	float area_f = (float)(signed_area >> 8);
	if(area_f == 0.0f) area_f = 1.0f;
	S.ooa = fabsf(1.0f / area_f);
	S.max_depth = ref_max_macro(S.p[0].z, ref_max_macro(S.p[1].z, S.p[2].z)); // MAX3 math.h:32
	const int mnx = min(x0, min(x1, x2)) >> 4, mny = min(y0, min(y1, y2)) >> 4;
	const int mxx = max(x0, max(x1, x2)) >> 4, mxy = max(y0, max(y1, y2)) >> 4;
	S.minx = min(max(mnx, 0), P.vp_w - 1);
	S.miny = min(max(mny, 0), P.vp_h - 1);
	S.maxx = min(mxx + 1, P.vp_w - 1);
	S.maxy = min(mxy + 1, P.vp_h - 1);
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

// clip_by_plane (main.c:609-647), plane_d = 0
__device__ __noinline__ int clip_by_plane(VsOut *v, int n, float4 pn) {
	VsOut res[16];
	int nout = 0;
	float cur = dot4_serial(pn, v[0].r0);
	bool cur_in = cur > -0.0f;
	for(int i = 0; i < n; ++i) {
		const int next = (i + 1) % n;
		if(cur_in && nout < 16) res[nout++] = v[i];
		const float nd = dot4_serial(pn, v[next].r0);
		const bool nin = nd > -0.0f;
		if(cur_in != nin && nout < 16) {
			const float t = (0.0f + cur) / (cur - nd);
			res[nout++] = lerp_vertex(v[i], v[next], t);
		}
		cur = nd;
		cur_in = nin;
	}
	for(int i = 0; i < nout; ++i) v[i] = res[i];
	return nout;
}

// clipper (main.c:649-660): six planes with host-normalised normals
__device__ __noinline__ int clip_polygon(VsOut *v, float k) {
	int n = 3;
	n = clip_by_plane(v, n, make_float4(k, 0.0f, 0.0f, k));
	n = clip_by_plane(v, n, make_float4(-k, 0.0f, 0.0f, k));
	n = clip_by_plane(v, n, make_float4(0.0f, k, 0.0f, k));
	n = clip_by_plane(v, n, make_float4(0.0f, -k, 0.0f, k));
	n = clip_by_plane(v, n, make_float4(0.0f, 0.0f, k, k));
	n = clip_by_plane(v, n, make_float4(0.0f, 0.0f, -k, k));
	return n;
}

__device__ __forceinline__ void emit_triangle(const GeomParams &P, uint32_t id, const TriSetup &S, const VsOut &v0, const VsOut &v1, const VsOut &v2) {
	if(id >= P.tri_capacity) return;
	// tile rectangle exactly as the binner derives it (main.c:927-928), C division truncating toward zero
	int tx0 = max(S.minx / 8, 0), ty0 = max(S.miny / 8, 0);
	int tx1 = min(S.maxx / 8, P.wt - 1), ty1 = min(S.maxy / 8, P.ht - 1);
	bool owned = false;
	if(tx0 <= tx1)
		for(int ty = ty0; ty <= ty1 && !owned; ++ty) owned = P.part.owns_row(ty);
	if(!owned) { // bins nothing on this rank
		tx0 = 1;
		tx1 = 0;
	}
	P.tri_bounds[id] = make_uint2((uint32_t)(tx0 & 0xffff) | ((uint32_t)(ty0 & 0xffff) << 16), (uint32_t)(tx1 & 0xffff) | ((uint32_t)(ty1 & 0xffff) << 16));
	if(owned) {
		uint4 *cov = P.tri_cov + (size_t)id * MLV_TRI_COV_U4;
		cov[0] = make_uint4(S.e[0], S.e[1], S.e[2], S.e[3]);
		cov[1] = make_uint4(S.e[4], S.e[5], S.e[6], S.e[7]);
		cov[2] = make_uint4(S.e[8], __float_as_uint(S.max_depth), (uint32_t)(tx0 & 0xffff) | ((uint32_t)(ty0 & 0xffff) << 16),
		                    (uint32_t)(tx1 & 0xffff) | ((uint32_t)(ty1 & 0xffff) << 16));
		float4 *sh = reinterpret_cast<float4 *>(P.tri_shade + (size_t)id * MLV_TRI_SHADE_U4);
		sh[0] = make_float4(S.ooa, S.p[0].z, S.p[1].z, S.p[2].z);
		sh[1] = make_float4(S.rw[0], S.rw[1], S.rw[2], v0.r2x);
		sh[2] = v0.r1;
		sh[3] = v1.r1;
		sh[4] = v2.r1;
		sh[5] = make_float4(v1.r2x, v2.r2x, 0.0f, 0.0f);
	}
	if(P.dbg.tris) {
		mlv_ref_triangle t;
		t.p_attributes = (uint64_t)id * 144ull;
		t.min_bounds[0] = S.minx;
		t.min_bounds[1] = S.miny;
		t.max_bounds[0] = S.maxx;
		t.max_bounds[1] = S.maxy;
		for(int k = 0; k < 3; ++k)
			for(int j = 0; j < 3; ++j) t.edges[k][j] = S.e[k * 3 + j];
		t.reciprocal_ws[0] = S.rw[0];
		t.reciprocal_ws[1] = S.rw[1];
		t.reciprocal_ws[2] = S.rw[2];
		t.one_over_area = S.ooa;
		t.max_depth = S.max_depth;
		P.dbg.tris[id] = t;
		float4 *a = reinterpret_cast<float4 *>(P.dbg.attrs + (size_t)id * 36);
		const VsOut *vv[3] = { &v0, &v1, &v2 };
		for(int k = 0; k < 3; ++k) {
			a[k * 3 + 0] = S.p[k];
			a[k * 3 + 1] = vv[k]->r1;
			a[k * 3 + 2] = make_float4(vv[k]->r2x, 0.0f, 0.0f, 0.0f);
		}
	}
}

#define MLV_GEOM_THREADS 256
#define MLV_SCAN_INVALID 0ull
#define MLV_SCAN_AGGREGATE 1ull
#define MLV_SCAN_PREFIX 2ull

__device__ __forceinline__ unsigned long long scan_pack(uint32_t epoch, unsigned long long flag, uint32_t value) {
	return ((unsigned long long)epoch << 34) | (flag << 32) | (unsigned long long)value;
}

template <int VS, bool INDEXED>
__global__ void __launch_bounds__(MLV_GEOM_THREADS) k_geom(const GeomParams P) {
	__shared__ uint32_t s_tile;
	__shared__ uint32_t s_warp_sums[MLV_GEOM_THREADS / 32];
	__shared__ uint32_t s_block_exclusive;

	// Blocks take their logical position from a ticket so that every predecessor a block may wait on in
	// the look-back below has already started (forward progress without relying on blockIdx order).
	if(threadIdx.x == 0) s_tile = atomicAdd(&P.ctr->ticket, 1u) - P.ticket_base;
	__syncthreads();
	const uint32_t tile = s_tile;
	const uint32_t t = tile * MLV_GEOM_THREADS + threadIdx.x;

	VsOut v[3];
	TriSetup S;
	VsOut poly[16];
	int n_poly = 0;
	uint32_t survivors = 0; // bit k: fan triangle (0,k+1,k+2) survives; bit 0 for the unclipped case
	bool clipped = false;

	if(t < P.tri_count) {
		// ---- input assembler (main.c:662-696): index fetch + 32-byte vertex fetch as two 128-bit loads
#pragma unroll
		for(int c = 0; c < 3; ++c) {
			const uint32_t vi = INDEXED ? __ldg(P.ib + 3u * t + c) : (3u * t + c);
			const float4 in0 = __ldg(P.vb + 2 * (size_t)vi);
			const float4 in1 = __ldg(P.vb + 2 * (size_t)vi + 1);
			// ---- vertex shader (main.c:698-734)
			v[c] = run_vs<VS>(in0, in1, P.cb, P.vs_tex, P.rsqrt_lut);
			if(P.dbg.vs_out) {
				float4 *o = reinterpret_cast<float4 *>(P.dbg.vs_out + (size_t)(3u * t + c) * 12);
				o[0] = v[c].r0;
				o[1] = v[c].r1;
				o[2] = make_float4(v[c].r2x, 0.0f, 0.0f, 0.0f);
			}
		}
		// ---- primitive assembly (main.c:750-908)
		const float4 a = v[0].r0, b = v[1].r0, c = v[2].r0;
		const bool degenerate = (a.w == 0.0f || b.w == 0.0f || c.w == 0.0f); // main.c:759
		const bool rejected =                                                // main.c:764-772
		    (a.x < -a.w && b.x < -b.w && c.x < -c.w) || (a.x > a.w && b.x > b.w && c.x > c.w) || (a.y < -a.w && b.y < -b.w && c.y < -c.w) ||
		    (a.y > a.w && b.y > b.w && c.y > c.w) || (a.z < 0.0f && b.z < 0.0f && c.z < 0.0f) || (a.z > a.w && b.z > b.w && c.z > c.w);
		if(!degenerate && !rejected) {
			const bool inside = // main.c:775-781
			    (a.x >= -a.w && b.x >= -b.w && c.x >= -c.w) && (a.x <= a.w && b.x <= b.w && c.x <= c.w) && (a.y >= -a.w && b.y >= -b.w && c.y >= -c.w) &&
			    (a.y <= a.w && b.y <= b.w && c.y <= c.w) && (a.z >= 0.0f && b.z >= 0.0f && c.z >= 0.0f) && (a.z <= a.w && b.z <= b.w && c.z <= c.w);
			if(inside) {
				survivors = setup_triangle(a, b, c, P, S) ? 1u : 0u;
			} else {
				clipped = true;
				poly[0] = v[0];
				poly[1] = v[1];
				poly[2] = v[2];
				n_poly = clip_polygon(poly, P.clip_k);
				for(int k = 1; k < n_poly - 1; ++k) { // fan (main.c:797-800); count the survivors now, emit them below
					TriSetup tmp;
					if(setup_triangle(poly[0].r0, poly[k].r0, poly[k + 1].r0, P, tmp)) survivors |= 1u << (k - 1);
				}
			}
		}
	}

	// ---- ordered output slots: block scan + decoupled look-back over blocks
	const uint32_t n_out = __popc(survivors);
	uint32_t incl = n_out;
#pragma unroll
	for(int d = 1; d < 32; d <<= 1) {
		const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
		if(lane_id() >= (uint32_t)d) incl += o;
	}
	const uint32_t warp = threadIdx.x >> 5;
	if(lane_id() == 31) s_warp_sums[warp] = incl;
	__syncthreads();
	if(warp == 0) {
		uint32_t ws = (lane_id() < MLV_GEOM_THREADS / 32) ? s_warp_sums[lane_id()] : 0u;
		uint32_t wincl = ws;
#pragma unroll
		for(int d = 1; d < 32; d <<= 1) {
			const uint32_t o = __shfl_up_sync(0xffffffffu, wincl, d);
			if(lane_id() >= (uint32_t)d) wincl += o;
		}
		const uint32_t block_total = __shfl_sync(0xffffffffu, wincl, MLV_GEOM_THREADS / 32 - 1);
		if(lane_id() < MLV_GEOM_THREADS / 32) s_warp_sums[lane_id()] = wincl - ws; // exclusive per warp

		volatile unsigned long long *state = P.scan_state;
		if(lane_id() == 0) state[tile] = scan_pack(P.epoch, tile == 0 ? MLV_SCAN_PREFIX : MLV_SCAN_AGGREGATE, block_total);
		uint32_t exclusive = 0;
		if(tile > 0) {
			int look = (int)tile - 1;
			while(true) {
				const int idx = look - (int)lane_id();
				unsigned long long s;
				bool ready;
				do { // spin until the whole window has been published for this draw
					s = (idx >= 0) ? state[idx] : scan_pack(P.epoch, MLV_SCAN_PREFIX, 0u);
					ready = ((uint32_t)(s >> 34) == P.epoch) && (((s >> 32) & 3ull) != MLV_SCAN_INVALID);
				} while(!__all_sync(0xffffffffu, ready));
				const bool is_prefix = ((s >> 32) & 3ull) == MLV_SCAN_PREFIX;
				const uint32_t pmask = __ballot_sync(0xffffffffu, is_prefix);
				uint32_t val = (uint32_t)s;
				if(pmask) {
					const int first = __ffs(pmask) - 1;
					if((int)lane_id() > first) val = 0u;
				}
#pragma unroll
				for(int d = 16; d > 0; d >>= 1) val += __shfl_xor_sync(0xffffffffu, val, d);
				exclusive += val;
				if(pmask) break;
				look -= 32;
			}
			if(lane_id() == 0) state[tile] = scan_pack(P.epoch, MLV_SCAN_PREFIX, exclusive + block_total);
		}
		if(lane_id() == 0) {
			s_block_exclusive = exclusive;
			if(tile == P.num_blocks - 1) { // stats (main.c:1228-1238) and the triangle count the later stages read
				const uint32_t total = exclusive + block_total;
				P.ctr->tri_count = min(total, P.tri_capacity);
				if(total > P.tri_capacity) atomicOr(&P.ctr->error_flags, MLV_FLAG_TRI_OVERFLOW);
				P.ctr->stats.vertex_count += P.index_count;
				P.ctr->stats.input_triangle_count += P.tri_count;
				P.ctr->stats.assembled_triangle_count += total;
			}
		}
	}
	__syncthreads();

	// ---- emission in reference order: input order, fan order inside a clipped triangle
	if(n_out) {
		uint32_t id = s_block_exclusive + s_warp_sums[warp] + (incl - n_out);
		if(!clipped) {
			emit_triangle(P, id, S, v[0], v[1], v[2]);
		} else {
			for(int k = 1; k < n_poly - 1; ++k) {
				if(!(survivors & (1u << (k - 1)))) continue;
				TriSetup tmp;
				setup_triangle(poly[0].r0, poly[k].r0, poly[k + 1].r0, P, tmp);
				emit_triangle(P, id++, tmp, poly[0], poly[k], poly[k + 1]);
			}
		}
	}
}

// =================================================================================================
// binner
// =================================================================================================

// Passes 1 (count) and 2 (fill) of the reference binner (main.c:924-962). One lane per assembled
// triangle; triangles overlapping more than 8 tiles are expanded cooperatively by the whole warp.
// The fill pass takes slots by decrementing the counters the count pass built (they are back to zero
// for the next draw when it finishes); the per-bin order this leaves is arbitrary and is restored to
// ascending triangle id by k_tile before use.
template <bool FILL>
__global__ void __launch_bounds__(256) k_bin(const BinParams P, uint32_t pair_capacity) {
	const uint32_t n = P.ctr->tri_count;
	if(FILL && P.ctr->pair_total > pair_capacity) return;
	const uint32_t lane = lane_id();
	const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
	for(uint32_t base = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32u; base < n; base += warps * 32u) {
		const uint32_t id = base + lane;
		int tx0 = 1, ty0 = 0, tx1 = 0, ty1 = -1;
		if(id < n) {
			const uint2 b = __ldg(P.tri_bounds + id);
			tx0 = (int)(short)(b.x & 0xffffu);
			ty0 = (int)(short)(b.x >> 16);
			tx1 = (int)(short)(b.y & 0xffffu);
			ty1 = (int)(short)(b.y >> 16);
		}
		const int w = max(tx1 - tx0 + 1, 0), h = max(ty1 - ty0 + 1, 0);
		const int cnt = w * h;
		const bool big = cnt > 8;
		if(cnt > 0 && !big) {
			for(int ty = ty0; ty <= ty1; ++ty) {
				if(!P.part.owns_row(ty)) continue;
				for(int tx = tx0; tx <= tx1; ++tx) {
					const uint32_t bin = (uint32_t)(ty * P.wt + tx);
					if(FILL) {
						const uint32_t old = atomicSub(P.bin_count + bin, 1u);
						P.pair_ids[P.bin_offset[bin] + old - 1u] = id;
					} else {
						atomicAdd(P.bin_count + bin, 1u);
					}
				}
			}
		}
		uint32_t bigmask = __ballot_sync(0xffffffffu, big);
		while(bigmask) {
			const int src = __ffs(bigmask) - 1;
			bigmask &= bigmask - 1;
			const int bx0 = __shfl_sync(0xffffffffu, tx0, src), by0 = __shfl_sync(0xffffffffu, ty0, src);
			const int bw = __shfl_sync(0xffffffffu, w, src), bc = __shfl_sync(0xffffffffu, cnt, src);
			const uint32_t bid = base + (uint32_t)src;
			for(int k = (int)lane; k < bc; k += 32) {
				const int ty = by0 + k / bw, tx = bx0 + k % bw;
				if(!P.part.owns_row(ty)) continue;
				const uint32_t bin = (uint32_t)(ty * P.wt + tx);
				if(FILL) {
					const uint32_t old = atomicSub(P.bin_count + bin, 1u);
					P.pair_ids[P.bin_offset[bin] + old - 1u] = bid;
				} else {
					atomicAdd(P.bin_count + bin, 1u);
				}
			}
		}
	}
}

// Exclusive scan of the per-bin counts + compaction of non-empty bins in ascending bin index
// (main.c:937-974), stats (main.c:1245-1246). Single CTA of 1024 threads; warp w owns a contiguous
// range of bins and walks it with coalesced 32-wide reads.
__global__ void __launch_bounds__(1024) k_bin_scan(const ScanParams P) {
	__shared__ uint32_t s_sum[32], s_nz[32];
	const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
	const uint32_t per_warp = ((P.num_bins + 32u * 32u - 1u) / (32u * 32u)) * 32u;
	const uint32_t begin = warp * per_warp, end = min(begin + per_warp, P.num_bins);
	uint32_t sum = 0, nz = 0;
	for(uint32_t i = begin + lane; i < end; i += 32u) {
		const uint32_t c = P.bin_count[i];
		sum += c;
		nz += (c != 0u);
	}
#pragma unroll
	for(int d = 16; d > 0; d >>= 1) {
		sum += __shfl_xor_sync(0xffffffffu, sum, d);
		nz += __shfl_xor_sync(0xffffffffu, nz, d);
	}
	if(lane == 0) {
		s_sum[warp] = sum;
		s_nz[warp] = nz;
	}
	__syncthreads();
	uint32_t run_sum = 0, run_nz = 0, tot_sum = 0, tot_nz = 0;
	for(uint32_t w2 = 0; w2 < 32u; ++w2) {
		if(w2 < warp) {
			run_sum += s_sum[w2];
			run_nz += s_nz[w2];
		}
		tot_sum += s_sum[w2];
		tot_nz += s_nz[w2];
	}
	const bool overflow = tot_sum > P.pair_capacity;
	for(uint32_t i0 = begin; i0 < end; i0 += 32u) {
		const uint32_t i = i0 + lane;
		const uint32_t c = (i < end) ? P.bin_count[i] : 0u;
		uint32_t incl = c;
#pragma unroll
		for(int d = 1; d < 32; d <<= 1) {
			const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
			if(lane >= (uint32_t)d) incl += o;
		}
		const uint32_t nzmask = __ballot_sync(0xffffffffu, c != 0u);
		if(i < end) {
			const uint32_t upto = run_sum + incl - c;
			P.bin_offset[i] = upto;
			if(c != 0u && !overflow) {
				mlv_ref_compacted_bin cb;
				cb.num_triangles_self = c;
				cb.num_triangles_upto = upto;
				cb.bin_index = i;
				P.cbins[run_nz + __popc(nzmask & ((1u << lane) - 1u))] = cb;
			}
			if(overflow) P.bin_count[i] = 0u; // the fill pass that would have drained the counters is skipped
		}
		run_sum += __shfl_sync(0xffffffffu, incl, 31);
		run_nz += __popc(nzmask);
	}
	if(threadIdx.x == 0) {
		P.ctr->pair_total = tot_sum;
		P.ctr->n_cbins_raw = tot_nz;
		P.ctr->n_cbins = overflow ? 0u : tot_nz;
		if(overflow) atomicOr(&P.ctr->error_flags, MLV_FLAG_PAIR_OVERFLOW);
		P.ctr->stats.active_bin_count += tot_nz;
		P.ctr->stats.total_triangle_count_in_bins += tot_sum;
	}
}

// =================================================================================================
// tile: rasterizer + Hi-Z + early-Z + pixel shader + output merger
// =================================================================================================

// Restores ascending-id order inside one bin list (the order the reference's serial fill produces,
// main.c:950-962). n <= 32: bitonic network in registers. Larger lists: stable LSD radix split, one bit
// per pass, ping-ponging between the list and a scratch segment of the same extent.
__device__ __forceinline__ void sort_bin_ids(uint32_t *ids, uint32_t *tmp, uint32_t n, uint32_t id_bits) {
	const uint32_t lane = lane_id();
	if(n <= 32u) {
		uint32_t v = (lane < n) ? ids[lane] : 0xffffffffu;
#pragma unroll
		for(int k = 2; k <= 32; k <<= 1) {
#pragma unroll
			for(int j = k >> 1; j > 0; j >>= 1) {
				const uint32_t o = __shfl_xor_sync(0xffffffffu, v, j);
				const bool up = ((lane & k) == 0);
				const bool lower = ((lane & j) == 0);
				v = (lower == up) ? min(v, o) : max(v, o);
			}
		}
		if(lane < n) ids[lane] = v;
		__syncwarp();
		return;
	}
	volatile uint32_t *src = ids;
	volatile uint32_t *dst = tmp;
	const uint32_t lt = (1u << lane) - 1u;
	for(uint32_t bit = 0; bit < id_bits; ++bit) {
		uint32_t zeros = 0;
		for(uint32_t i0 = 0; i0 < n; i0 += 32u) {
			const uint32_t i = i0 + lane;
			const bool isz = (i < n) && !((src[i] >> bit) & 1u);
			zeros += __popc(__ballot_sync(0xffffffffu, isz));
		}
		if(zeros == 0u || zeros == n) continue; // this bit does not discriminate
		uint32_t zpos = 0, opos = zeros;
		for(uint32_t i0 = 0; i0 < n; i0 += 32u) {
			const uint32_t i = i0 + lane;
			const bool valid = i < n;
			const uint32_t val = valid ? src[i] : 0u;
			const bool isz = valid && !((val >> bit) & 1u);
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

// 64 coverage tests of one triangle against one tile (main.c:1012-1038): E_k = ((a_k*x)<<4) + ((b_k*y)<<4) + c_k
// in wrapping i32 == a_k*(16x) + b_k*(16y) + c_k (mod 2^32), stepped incrementally; inside iff (E0|E1|E2) > 0.
__device__ __forceinline__ void coverage_64(const uint4 c0, const uint4 c1, const uint32_t c2x, uint32_t X0, uint32_t Y0, uint32_t &lo, uint32_t &hi) {
	const uint32_t a0 = c0.x, b0 = c0.y, a1 = c0.w, b1 = c1.x, a2 = c1.z, b2 = c1.w;
	uint32_t e0 = a0 * X0 + b0 * Y0 + c0.z;
	uint32_t e1 = a1 * X0 + b1 * Y0 + c1.y;
	uint32_t e2 = a2 * X0 + b2 * Y0 + c2x;
	const uint32_t sx0 = a0 << 4, sx1 = a1 << 4, sx2 = a2 << 4;
	const uint32_t sy0 = b0 << 4, sy1 = b1 << 4, sy2 = b2 << 4;
	lo = 0u;
	hi = 0u;
#pragma unroll
	for(int y = 0; y < 8; ++y) {
		uint32_t r0 = e0, r1 = e1, r2 = e2;
#pragma unroll
		for(int x = 0; x < 8; ++x) {
			const bool in = (int)(r0 | r1 | r2) > 0;
			if(y < 4) lo |= in ? (1u << (y * 8 + x)) : 0u;
			else hi |= in ? (1u << ((y - 4) * 8 + x)) : 0u;
			r0 += sx0;
			r1 += sx1;
			r2 += sx2;
		}
		e0 += sy0;
		e1 += sy1;
		e2 += sy2;
	}
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

template <int PS>
__device__ __forceinline__ uint32_t shade_pixel(const TileParams &P, uint32_t id, uint32_t X, uint32_t Y) {
	const uint4 *cov = P.tri_cov + (size_t)id * MLV_TRI_COV_U4;
	const uint4 c0 = __ldg(cov), c1 = __ldg(cov + 1);
	const uint32_t c2x = __ldg(reinterpret_cast<const uint32_t *>(cov + 2));
	const uint32_t E1 = c0.w * X + c1.x * Y + c1.y;
	const uint32_t E2 = c1.z * X + c1.w * Y + c2x;
	const float4 *sh = reinterpret_cast<const float4 *>(P.tri_shade + (size_t)id * MLV_TRI_SHADE_U4);
	const float4 s0 = __ldg(sh), s1 = __ldg(sh + 1), r1a = __ldg(sh + 2), r1b = __ldg(sh + 3), r1c = __ldg(sh + 4), s5 = __ldg(sh + 5);
	float bx, by;
	barycentrics(E1, E2, s0.x, bx, by);
	// perspective correction (main.c:1106-1115)
	float denom = 1.0f - (bx + by);
	denom = denom * s1.x;
	denom = denom + bx * s1.y;
	denom = denom + by * s1.z;
	denom = 1.0f / denom;
	const float pbx = (bx * s1.y) * denom;
	const float pby = (by * s1.z) * denom;
	const float4 r1 = make_float4(interp(r1a.x, r1b.x, r1c.x, pbx, pby), interp(r1a.y, r1b.y, r1c.y, pbx, pby), interp(r1a.z, r1b.z, r1c.z, pbx, pby),
	                              interp(r1a.w, r1b.w, r1c.w, pbx, pby));
	const float r2x = interp(s1.w, s5.x, s5.y, pbx, pby);
	return encode_color(run_ps<PS>(r1, r2x, P.ps_tex, P.rsqrt_lut));
}

#define MLV_TILE_THREADS 256

template <int PS>
__global__ void __launch_bounds__(MLV_TILE_THREADS) k_tile(const TileParams P) {
	const uint32_t lane = lane_id();
	const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
	const uint32_t n_cbins = P.ctr->n_cbins;
	const uint32_t tri_count = P.ctr->tri_count;
	const uint32_t id_bits = 32u - __clz(max(tri_count, 2u) - 1u);
	const uint32_t px = (lane & 3u) * 2u, py = lane >> 2;

	for(uint32_t cb = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; cb < n_cbins; cb += warps) {
		const mlv_ref_compacted_bin bin = P.cbins[cb];
		const uint32_t n = bin.num_triangles_self, off = bin.num_triangles_upto, b = bin.bin_index;
		const uint32_t X0 = ((b % (uint32_t)P.wt) * 8u) << 4, Y0 = ((b / (uint32_t)P.wt) * 8u) << 4;
		const uint32_t X = X0 + (px << 4), Y = Y0 + (py << 4);

		// read_tile (main.c:577-587)
		uint4 pix = P.fb[(size_t)b * 32u + lane];
		float d0 = __uint_as_float(pix.z), d1 = __uint_as_float(pix.w);
		uint32_t win0 = MLV_NO_WINNER, win1 = MLV_NO_WINNER;
		const float tile_min_old = P.tile_min[b]; // get_tile_minimum_depth: previous draws only (N3)

		uint32_t *ids = P.pair_ids + off;
		sort_bin_ids(ids, P.pair_tmp + off, n, id_bits);

		for(uint32_t base = 0; base < n; base += 32u) {
			// ---- rasterizer, lanes over triangles (main.c:996-1041)
			const uint32_t k = base + lane;
			uint32_t id = 0, lo = 0, hi = 0;
			uint4 c0 = make_uint4(0, 0, 0, 0), c1 = c0;
			uint32_t c2x = 0;
			if(k < n) {
				id = ids[k];
				const uint4 *cov = P.tri_cov + (size_t)id * MLV_TRI_COV_U4;
				c0 = __ldg(cov);
				c1 = __ldg(cov + 1);
				const uint2 c2 = __ldg(reinterpret_cast<const uint2 *>(cov + 2));
				c2x = c2.x;
				const float max_depth = __uint_as_float(c2.y);
				if(!(max_depth < tile_min_old)) coverage_64(c0, c1, c2x, X0, Y0, lo, hi); // Hi-Z (main.c:1005-1010)
				if(P.dbg.infos) {
					mlv_ref_tile_info ti;
					ti.triangle_id = id;
					ti._pad = 0;
					ti.fragment_mask = ((unsigned long long)hi << 32) | lo;
					P.dbg.infos[off + k] = ti;
				}
			}
			// ---- early-Z in list order, lanes over pixels (main.c:1060-1168)
			uint32_t live = __ballot_sync(0xffffffffu, (lo | hi) != 0u);
			while(live) {
				const int src = __ffs(live) - 1;
				live &= live - 1;
				const uint32_t sid = __shfl_sync(0xffffffffu, id, src);
				const uint32_t slo = __shfl_sync(0xffffffffu, lo, src), shi = __shfl_sync(0xffffffffu, hi, src);
				const uint32_t bits = (((py < 4u) ? slo : shi) >> ((py & 3u) * 8u + px)) & 3u;
				if(bits) {
					const uint4 *cov = P.tri_cov + (size_t)sid * MLV_TRI_COV_U4;
					const uint4 q0 = __ldg(cov), q1 = __ldg(cov + 1);
					const uint32_t q2x = __ldg(reinterpret_cast<const uint32_t *>(cov + 2));
					const float4 s0 = __ldg(reinterpret_cast<const float4 *>(P.tri_shade + (size_t)sid * MLV_TRI_SHADE_U4));
					const uint32_t E1 = q0.w * X + q1.x * Y + q1.y;
					const uint32_t E2 = q1.z * X + q1.w * Y + q2x;
					if(bits & 1u) {
						float bx, by;
						barycentrics(E1, E2, s0.x, bx, by);
						const float z = interp(s0.y, s0.z, s0.w, bx, by);
						if(z >= d0) { // _CMP_GE_OQ, reversed Z (main.c:1166)
							d0 = z;
							win0 = sid;
						}
					}
					if(bits & 2u) {
						float bx, by;
						barycentrics(E1 + (q0.w << 4), E2 + (q1.z << 4), s0.x, bx, by);
						const float z = interp(s0.y, s0.z, s0.w, bx, by);
						if(z >= d1) {
							d1 = z;
							win1 = sid;
						}
					}
				}
			}
		}

		// ---- pixel shader + output merger, once per pixel on the last fragment that passed (main.c:1170-1181)
		if(win0 != MLV_NO_WINNER) pix.x = shade_pixel<PS>(P, win0, X, Y);
		if(win1 != MLV_NO_WINNER) pix.y = shade_pixel<PS>(P, win1, X + 16u, Y);
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

// =================================================================================================
// resolve / composite
// =================================================================================================

// Tiled -> row-major. One thread per 4 horizontally adjacent pixels: two 128-bit loads, one 128-bit colour
// store (+ one 128-bit depth store).
__global__ void __launch_bounds__(256) k_resolve(const uint4 *__restrict__ fb, uint4 *__restrict__ colors, float4 *__restrict__ depths, int width, int height) {
	const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; // quad index, row-major over (width/4) x height
	const uint32_t qw = (uint32_t)width >> 2;
	if(q >= qw * (uint32_t)height) return;
	const uint32_t y = q / qw, xq = q % qw;
	const uint32_t tx = xq >> 1, ty = y >> 3;
	const uint32_t bin = ty * ((uint32_t)width >> 3) + tx;
	const uint32_t lane = (y & 7u) * 4u + (xq & 1u) * 2u;
	const uint4 a = __ldg(fb + (size_t)bin * 32u + lane), b = __ldg(fb + (size_t)bin * 32u + lane + 1u);
	colors[q] = make_uint4(a.x, a.y, b.x, b.y);
	if(depths) depths[q] = make_float4(__uint_as_float(a.z), __uint_as_float(a.w), __uint_as_float(b.z), __uint_as_float(b.w));
}

// Sort-first compositing helpers. Chunk r of the gather buffer = row-major colour of rank r's
// stripes in ascending stripe order. PACK: tiled framebuffer of this rank -> its own chunk.
// UNPACK: every chunk -> the final row-major image.
__global__ void __launch_bounds__(256) k_composite_pack(const uint4 *__restrict__ fb, uint4 *__restrict__ chunk, int width, int height, Partition part) {
	const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t qw = (uint32_t)width >> 2;
	if(q >= qw * (uint32_t)height) return;
	const uint32_t y = q / qw, xq = q % qw;
	const uint32_t ty = y >> 3;
	if(!part.owns_row((int)ty)) return;
	const uint32_t stripe = ty / (uint32_t)part.stripe_h;
	const uint32_t local_stripe = stripe / (uint32_t)part.num_ranks;
	const uint32_t local_y = (local_stripe * (uint32_t)part.stripe_h + (ty % (uint32_t)part.stripe_h)) * 8u + (y & 7u);
	const uint32_t bin = ty * ((uint32_t)width >> 3) + (xq >> 1);
	const uint32_t lane = (y & 7u) * 4u + (xq & 1u) * 2u;
	const uint4 a = __ldg(fb + (size_t)bin * 32u + lane), b = __ldg(fb + (size_t)bin * 32u + lane + 1u);
	chunk[(size_t)local_y * qw + xq] = make_uint4(a.x, a.y, b.x, b.y);
}

__global__ void __launch_bounds__(256) k_composite_unpack(const uint4 *__restrict__ gather, uint4 *__restrict__ colors, int width, int height, int num_ranks, int stripe_h,
                                                          size_t chunk_u4) {
	const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t qw = (uint32_t)width >> 2;
	if(q >= qw * (uint32_t)height) return;
	const uint32_t y = q / qw, xq = q % qw;
	const uint32_t ty = y >> 3;
	const uint32_t stripe = ty / (uint32_t)stripe_h;
	const uint32_t owner = stripe % (uint32_t)num_ranks;
	const uint32_t local_stripe = stripe / (uint32_t)num_ranks;
	const uint32_t local_y = (local_stripe * (uint32_t)stripe_h + (ty % (uint32_t)stripe_h)) * 8u + (y & 7u);
	colors[q] = __ldg(gather + (size_t)owner * chunk_u4 + (size_t)local_y * qw + xq);
}

} // namespace mlv
