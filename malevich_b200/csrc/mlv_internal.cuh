// mlv_internal.cuh -- device-side record layouts and kernel parameter blocks.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/malevich_b200.h"
#include "shaders.cuh"

namespace mlv {

// ---- HBM layouts (DESIGN.md "Data layout") ----------------------------------------------------
//
// Framebuffer: TILED. Tile (bin) b owns 32 consecutive uint4 (512 B). Lane l of the warp that owns the
// tile holds pixels p0 = (x = l&7, y = l>>3) and p1 = (x, y+4), i.e. bits l and 32+l of the reference's
// 64-bit fragment mask:  uint4 = { colour(p0), colour(p1), depth_bits(p0), depth_bits(p1) }.
// One LDG.128 + one STG.128 per lane per tile-draw.
//
// Triangle identity. The reference numbers assembled triangles in single-thread output order (input order,
// fan order inside a clipped triangle). Here a triangle is named by the order-preserving
// KEY = (input_triangle << 3) | fan_index, so no ordered compaction (scan) is needed: ascending key ==
// ascending reference id. Records live in SLOTS: an unclipped input triangle t uses direct slot t (key = t << 3); the
// fan triangles of a clipped input triangle use consecutive overflow slots T + base + fan_index, whose key is word 3 of
// their bounds entry. The per-tile lists hold SLOTS; keys are only needed to break depth ties (main.c:1166).
//   TriCov   48 B  { a0,b0,c0,a1 | b1,c1,a2,b2 | c2, max_depth, minx|flags_miny<<16, maxx|maxy<<16 }  coverage + Hi-Z
//   TriShade 96 B  { ooa,z0,z1,z2 | rw0,rw1,rw2,r2x_v0 | r1_v0 | r1_v1 | r1_v2 | r2x_v1,r2x_v2,0,0 }  depth + attributes
//   bounds   16 B  { minx | (miny|nowrap<<15)<<16, maxx | maxy<<16, max_depth, key } int16 pixel bounds (main.c:888-898);
//                  minx == 0x7fff: bins nothing (culled, redirected, not on this rank, or hidden by Hi-Z in every tile)
#define MLV_TRI_COV_U4 3
#define MLV_TRI_SHADE_U4 6
#define MLV_BOUNDS_EMPTY 0x00007fffu
#define MLV_NOWRAP_BIT 0x80000000u /* bit 15 of miny inside word 10 */

#define MLV_NO_WINNER 0xffffffffu
#define MLV_STAT_STRIPES 64u
#define MLV_TILE_MIN_CLEARED 0x80000000u /* tile_min bit pattern (-0.0f) the depth clear writes: equals 0.0 in every comparison, marks "never refreshed" */

enum { MLV_FLAG_TRI_OVERFLOW = 1u, MLV_FLAG_PAIR_OVERFLOW = 2u, MLV_FLAG_COMPOSITE_TIMEOUT = 4u };

// Per-draw counters the FRONT half of a draw hands to its BACK half (one set per draw context, reset by the draw's k_tile).
struct DrawCounters {
	uint32_t clip_count; // input triangles queued for k_front_clip
	uint32_t ovf_count;  // overflow slots taken by the fan triangles of clipped input triangles (may exceed the capacity: the draw is skipped)
	uint32_t big_count;  // slots whose tile rectangle holds 9 .. MLV_HUGE_TILES tiles (counted by a warp each in the back half)
	uint32_t huge_count; // ... more than MLV_HUGE_TILES tiles (counted by the whole grid)
};

struct Counters {
	uint32_t reserved0;
	uint32_t pair_total;  // (triangle,tile) pairs of the current draw
	uint32_t n_cbins;     // non-empty bins of the current draw (0 if the pair arena overflowed)
	uint32_t error_flags; // sticky MLV_FLAG_*
	uint32_t ticket;      // block ticket of the current draw's look-back scan (reset at the end of the draw)
	uint32_t draw_tris;   // assembled triangles of the current draw
	uint32_t last_ovf_count; // ovf_count of the last finished draw (debug read-back)
	uint32_t reserved1;
	uint32_t draw_alive;  // a pair of the current draw survived Hi-Z (set by the back half): the scan, the fill pass and k_tile have work
	uint32_t draw_pairs_all;   // (triangle,tile) pairs of the current draw including Hi-Z-rejected ones (Stats)
	uint32_t draw_active_bins; // bins the current draw touched (counted by its k_tile)
	uint32_t reserved2;
	uint32_t bcast_done;  // CTAs of k_composite_broadcast that have finished their stores (reset by the last one)
	uint32_t epoch;       // draw epoch tagging the look-back words of this draw's scan (advanced at the end of every draw; never 0)
	uint32_t n_wake;      // entries of wake_bins in the current draw
	uint32_t last_pair_total, last_n_cbins; // pair_total / n_cbins of the last finished draw (read-backs); pair_total 0xffffffff: it was skipped
	uint32_t pad[3];
	unsigned long long tile_done; // low word: CTAs of the current k_tile that have finished (the last one folds the draw's Stats); high word: bins they counted as touched
	mlv_stats stats;      // accumulated like reference main.c:1228-1246
	mlv_work_counters work; // what the kernels really processed (Hi-Z at binning time removes work the reference's Stats still count)
};

struct Partition { // sort-first ownership
	int num_ranks, rank, stripe_h;
	__host__ __device__ __forceinline__ bool owns_row(int ty) const { return num_ranks <= 1 || ((ty / stripe_h) % num_ranks) == rank; }
	// number of tile rows in [0, x) this rank owns: full periods of num_ranks stripes + the part of the last period inside this rank's stripe
	__host__ __device__ __forceinline__ int owned_below(int x) const {
		const int period = num_ranks * stripe_h, q = x / period, rem = x % period;
		int in = rem - rank * stripe_h;
		in = in < 0 ? 0 : (in > stripe_h ? stripe_h : in);
		return q * stripe_h + in;
	}
	// number of tile rows in [ty0, ty1] this rank owns
	__host__ __device__ __forceinline__ int owned_rows(int ty0, int ty1) const {
		if(ty1 < ty0) return 0;
		if(num_ranks <= 1) return ty1 - ty0 + 1;
		return owned_below(ty1 + 1) - owned_below(ty0);
	}
};

// Peer-memory compositing (fused form): every rank keeps two row-major images and one arrival word
// per source rank; peers write into them over NVLink (cudaIpc mappings, or plain pointers inside one process).
#ifndef MLV_MAX_PEERS
#define MLV_MAX_PEERS 16
#endif
struct PeerTargets {
	uint4 *color[MLV_MAX_PEERS];    // the image of the current frame parity on rank p
	uint32_t *flags[MLV_MAX_PEERS]; // rank p's arrival words, one per source rank
};

struct DebugOut {
	mlv_ref_triangle *tris; // 80 B per slot
	float *attrs;           // 36 floats per slot (9 x float4)
	uint32_t *slot_key;     // key per slot, 0xffffffff = slot unused
	float *vs_out;          // 12 floats per vertex
	mlv_ref_tile_info *infos;
};

// Index fetch of the input assembler (main.c:681-683) with the parts of ID3D11DeviceContext::DrawIndexed the reference
// leaves as TODOs: 16-bit index buffers (main.c:72), StartIndexLocation and BaseVertexLocation (main.c:1219).
struct IndexStream {
	const void *ib;
	uint32_t start_index;
	int32_t base_vertex;
	uint32_t index16; // 0: u32 indices, 1: u16 indices
	__device__ __forceinline__ uint32_t fetch(uint32_t i) const {
		const uint32_t raw = index16 ? (uint32_t)__ldg(reinterpret_cast<const unsigned short *>(ib) + start_index + i) : __ldg(reinterpret_cast<const uint32_t *>(ib) + start_index + i);
		return raw + (uint32_t)base_vertex;
	}
	// the three indices of triangle t: one decision about the index width and one address computation per triangle
	__device__ __forceinline__ void fetch3(uint32_t t, uint32_t &i0, uint32_t &i1, uint32_t &i2) const {
		const size_t first = (size_t)start_index + 3u * (size_t)t;
		if(index16) {
			const unsigned short *p = reinterpret_cast<const unsigned short *>(ib) + first;
			i0 = __ldg(p), i1 = __ldg(p + 1), i2 = __ldg(p + 2);
		} else {
			const uint32_t *p = reinterpret_cast<const uint32_t *>(ib) + first;
			i0 = __ldg(p), i1 = __ldg(p + 1), i2 = __ldg(p + 2);
		}
		i0 += (uint32_t)base_vertex, i1 += (uint32_t)base_vertex, i2 += (uint32_t)base_vertex;
	}
};

struct GeomParams {
	IndexStream ix;
	const float4 *vb;
	uint32_t tri_count;    // T: input triangles == number of direct slots
	uint32_t ovf_capacity; // overflow slots available
	float cb[48];
	const float *cbp; // non-null: the constants live here (a recorded command list's per-draw device copy), cb[] is ignored
	TexDesc vs_tex;
	const uint32_t *rsqrt_lut;
	// screen_from_ndc (main.c:825-830), computed on the host in double like the reference's initialiser
	float vp_m00, vp_m03, vp_m11, vp_m13, vp_m22, vp_m23;
	int vp_w, vp_h; // (i32)viewport.width / height
	int wt, ht;     // WIDTH_IN_TILES / HEIGHT_IN_TILES
	float clip_k;   // component of the host-normalised clip-plane normals (main.c:652-657)
	Partition part;
	uint4 *tri_cov;   // records of the direct slots (shared by all draws: written by k_back, read by the same draw's k_tile)
	uint4 *tri_shade;
	// ---- the draw context: what the front half of THIS draw produces for its back half
	uint4 *tri_bounds;          // 16 B per slot (direct + overflow)
	uint4 *ovf_cov;             // records of the overflow slots (fan triangles of clipped input triangles)
	uint4 *ovf_shade;
	DrawCounters *dctr;
	const float4 *chunk_bounds; // sort-first chunk culling: 2 x float4 per 256-triangle chunk, or null
	uint8_t *chunk_live;        // written by k_front's cull test: 1 = this rank processed the chunk
	float4 *vcache; // 2 x float4 per unique vertex: clip-space position, {snapped x, snapped y, screen z, clip_code bits}
	uint32_t *clip_queue;
	uint32_t *big_queue;
	uint32_t *huge_queue;
	uint32_t *bin_count;
	uint8_t *touch_bits;        // the draw context's "bin received a pair" map (Stats: active_bin_count), one byte per bin
	uint4 *warp_sum;            // the draw context's warp summaries, 16 B per 32 direct slots (kernels.cuh, write_warp_summary)
	const float *tile_min;
	bool keep_all; // debug capture: no Hi-Z at binning time, lists hold every pair like the reference's
	DebugOut dbg;
	Counters *ctr;
	unsigned long long *stat_stripes;
	uint32_t index_count;
	uint32_t draw_ordinal; // position of the draw inside a recorded command list (identifies the kernel nodes whose constants are updated)
	unsigned long long *timeline; // this launch's slot of the device-side timeline (4 words), or null
};

// Everything the tail kernel of a draw needs (scan | fill | tile).
struct TailParams {
	// binning
	const uint4 *tri_bounds;   // the draw context's bounds (word 3 of an overflow slot's entry is its key)
	DrawCounters *dctr;        // re-armed at the end of the draw
	const uint32_t *big_queue;
	const uint32_t *huge_queue;
	const uint8_t *chunk_live;
	float *tile_min;
	uint32_t *bin_count;       // surviving pairs per bin (zero outside a draw's back half .. scan)
	uint8_t *touch_bits;       // the draw context's touch map: counted and cleared by k_tile
	const uint32_t *warp_sum;  // the draw context's warp summaries as words (word 3 of an entry == 0: the fill pass skips the 32 slots)
	uint32_t *bin_offset;      // start of the bin's list (the scan), then the running fill position (k_fill)
	unsigned long long *state_sum, *state_nz; // look-back words of the scan, tagged with the draw epoch
	uint32_t scan_blocks;
	uint32_t *pair_ids;
	uint32_t *pair_tmp;
	mlv_ref_compacted_bin *cbins;
	Counters *ctr;
	unsigned long long *stat_stripes;
	uint32_t direct_slots;     // T
	uint32_t ovf_capacity;
	uint32_t pair_capacity;
	uint32_t num_bins;
	uint32_t bin_begin, bin_end; // this rank's bins (the whole render target unless it owns one contiguous band)
	int wt, ht;
	Partition part;
	bool keep_all;
	bool sort_lists;           // debug capture: restore ascending-key order inside every bin list
	// tile phase
	const uint4 *tri_cov;
	const uint4 *tri_shade;
	const uint4 *ovf_cov;      // records of the slots >= direct_slots
	const uint4 *ovf_shade;
	uint4 *fb;
	TexDesc ps_tex;
	const uint32_t *rsqrt_lut;
	DebugOut dbg;
	uint32_t index_count;      // Stats (main.c:1228-1232)
	uint32_t key_bits;
	unsigned long long *timeline; // three consecutive slots of the device-side timeline (scan, fill, tile), or null
};

} // namespace mlv
