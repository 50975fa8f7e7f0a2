// mlv_internal.cuh -- device-side record layouts and kernel parameter blocks.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/malevich_b200.h"
#include "shaders.cuh"

namespace mlv {

// ---- HBM layouts (DESIGN.md "Data layout") ----------------------------------------------------
//
// Framebuffer: TILED. Tile (bin) b owns 32 consecutive uint4 (512 B). Lane l of the warp that owns
// the tile holds pixels (x0 = 2*(l&3), y = l>>2) and (x0+1, y):  uint4 = { colour(x0), colour(x0+1),
// depth_bits(x0), depth_bits(x0+1) }.  One LDG.128 + one STG.128 per lane per tile-draw.
//
// Per assembled triangle (id = position in the reference's single-thread output order):
//   TriCov   48 B  { a0,b0,c0,a1 | b1,c1,a2,b2 | c2, max_depth, tile_bounds_lo, tile_bounds_hi }   coverage + Hi-Z
//   TriShade 96 B  { ooa,z0,z1,z2 | rw0,rw1,rw2,r2x_v0 | r1_v0 | r1_v1 | r1_v2 | r2x_v1,r2x_v2,0,0 }  depth + attributes
//   bounds    8 B  int16 { tx0, ty0, tx1, ty1 } inclusive tile rectangle (tx0 > tx1 => bins nothing)
#define MLV_TRI_COV_U4 3
#define MLV_TRI_SHADE_U4 6

#define MLV_NO_WINNER 0xffffffffu

enum { MLV_FLAG_TRI_OVERFLOW = 1u, MLV_FLAG_PAIR_OVERFLOW = 2u };

struct Counters {
	uint32_t tri_count;   // assembled triangles of the current draw (written by the last geometry block)
	uint32_t pair_total;  // (triangle,tile) pairs of the current draw
	uint32_t n_cbins;     // non-empty bins of the current draw (0 if the pair arena overflowed)
	uint32_t error_flags; // sticky MLV_FLAG_*
	uint32_t ticket;      // monotonically increasing block ticket of the geometry kernel
	uint32_t n_cbins_raw; // non-empty bins even when overflowed
	uint32_t pad[2];
	mlv_stats stats;      // accumulated like reference main.c:1228-1246
};

struct Partition { // sort-first ownership
	int num_ranks, rank, stripe_h;
	__host__ __device__ __forceinline__ bool owns_row(int ty) const { return num_ranks <= 1 || ((ty / stripe_h) % num_ranks) == rank; }
};

struct DebugOut {
	mlv_ref_triangle *tris; // 80 B each
	float *attrs;           // 36 floats each (9 x float4)
	float *vs_out;          // 12 floats per vertex
	mlv_ref_tile_info *infos;
};

struct GeomParams {
	const uint32_t *ib;
	const float4 *vb;
	uint32_t tri_count;
	uint32_t tri_capacity;
	float cb[48];
	TexDesc vs_tex;
	const uint32_t *rsqrt_lut;
	// screen_from_ndc (main.c:825-830), computed on the host in double like the reference's initialiser
	float vp_m00, vp_m03, vp_m11, vp_m13, vp_m22, vp_m23;
	int vp_w, vp_h; // (i32)viewport.width / height
	int wt, ht;     // WIDTH_IN_TILES / HEIGHT_IN_TILES
	float clip_k;   // component of the host-normalised clip-plane normals (main.c:652-657)
	Partition part;
	uint4 *tri_cov;
	uint4 *tri_shade;
	uint2 *tri_bounds;
	DebugOut dbg;
	unsigned long long *scan_state;
	Counters *ctr;
	uint32_t ticket_base;
	uint32_t epoch;
	uint32_t num_blocks;
	uint32_t index_count;
};

struct BinParams {
	const uint2 *tri_bounds;
	uint32_t *bin_count;
	uint32_t *bin_cursor;
	const uint32_t *bin_offset;
	uint32_t *pair_ids;
	Counters *ctr;
	int wt, ht;
	Partition part;
};

struct ScanParams {
	uint32_t *bin_count;
	uint32_t *bin_cursor;
	uint32_t *bin_offset;
	mlv_ref_compacted_bin *cbins;
	Counters *ctr;
	uint32_t num_bins;
	uint32_t pair_capacity;
};

struct TileParams {
	const mlv_ref_compacted_bin *cbins;
	uint32_t *pair_ids;
	uint32_t *pair_tmp;
	const uint4 *tri_cov;
	const uint4 *tri_shade;
	uint4 *fb;
	float *tile_min;
	uint32_t *bin_count;
	Counters *ctr;
	TexDesc ps_tex;
	const uint32_t *rsqrt_lut;
	DebugOut dbg;
	int wt;
};

} // namespace mlv
