// shaders.cuh -- shader core + the seven shader programs as __device__ functions.
//
// Arithmetic contract: every fp32 operation below is a separately rounded IEEE
// operation in the reference's order. This TU is compiled with -fmad=false -prec-div=true, so `a*b+c`
// here is a multiply followed by an add exactly like the reference's _mm256_mul_ps/_mm256_add_ps
// pairs under -ffp-contract=off; the only fused operations are the explicit __fmaf_rn calls that
// correspond to the reference's _mm256_fmadd_ps (passthrough_vs.c:20-21, fullscreen_vs.c:28-29).
// Multiplications by literal 0/1 are kept where the reference performs them (NaN/Inf/-0 propagate
// identically).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mlv {

struct TexDesc {
	const void *data;
	int width;
	int height;
	int format;
	const void *mips; // levels 1 .. mip_levels-1, tightly packed one after the other (mlv_texture_generate_mips); null = none
	int mip_levels;   // including level 0; 1 = no mip chain
};

// VS output registers actually consumed downstream: r0 = SV_POSITION, r1 = (NORMAL|COLOR|VIEW_DIR).xyz + UV.x,
// r2.x = UV.y (Vs_Output structs, e.g. basic_vs.c:9-14). r2.yzw are uninitialised stack in the reference
// (f256 vertex_output[12] main.c:711) and never read by any pixel shader; they are 0 here.
struct VsOut {
	float4 r0, r1;
	float r2x;
};

// ---- x86 emulation helpers -------------------------------------------------------------------

// vrsqrtps (math.h:278) -- piecewise-constant hardware approximation, reproduced from the table
// dumped by tools/gen_rsqrt_lut.py.
__device__ __forceinline__ float x86_rsqrt(float x, const uint32_t *__restrict__ lut) {
	const uint32_t u = __float_as_uint(x);
	const uint32_t e = (u >> 23) & 0xffu;
	if(e == 255u) {
		if(u & 0x7fffffu) return __uint_as_float(u | 0x00400000u); // NaN in, quiet NaN out
		return (u >> 31) ? __uint_as_float(0xffc00000u) : 0.0f;      // -inf -> NaN, +inf -> 0
	}
	if(e == 0u) return __uint_as_float((u & 0x80000000u) | 0x7f800000u); // +-0 / denormal -> +-inf
	if(u >> 31) return __uint_as_float(0xffc00000u);                        // negative -> NaN
	const int parity = ((int)e - 127) & 1;
	const int half = ((int)e - 127 - parity) / 2;
	const uint32_t l = __ldg(lut + parity * 1024 + ((u >> 13) & 1023u));
	return __uint_as_float(l - ((uint32_t)half << 23));
}

// _mm256_cvtps_epi32 (main.c:1176-1178, common_shader_core.h:106-107): round-to-nearest-even, and the
// x86 "integer indefinite" 0x80000000 for NaN / out-of-range instead of CUDA's saturation (App. C 11).
__device__ __forceinline__ int x86_cvt_rne(float f) {
	if(!(fabsf(f) < 2147483648.0f)) return (int)0x80000000; // (one compare: -2^31 itself converts to the same bit pattern)
	return __float2int_rn(f);
}

__device__ __forceinline__ float ref_max_macro(float x, float y) { return (x > y) ? x : y; } // MAX math.h:29
__device__ __forceinline__ float ref_min_macro(float x, float y) { return (x < y) ? x : y; } // MIN math.h:26

// v4f256_dot (math.h:142-146): (x*x' + y*y') + (z*z' + w*w')
__device__ __forceinline__ float dot4_pairwise(float4 a, float4 b) {
	const float xy = a.x * b.x + a.y * b.y;
	const float zw = a.z * b.z + a.w * b.w;
	return xy + zw;
}
// v4f32_dot (math.h:137-140): ((x*x' + y*y') + z*z') + w*w'
__device__ __forceinline__ float dot4_serial(float4 a, float4 b) {
	return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
}
// v3f256_dot (math.h:153-157): (x*x' + y*y') + z*z'
__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {
	float r = ax * bx + ay * by;
	r = r + az * bz;
	return r;
}
// m4x4f32_mul_v4f256 (math.h:174-177)
__device__ __forceinline__ float4 mul_m4_v4_pairwise(const float *__restrict__ m, float4 v) {
	return make_float4(dot4_pairwise(make_float4(m[0], m[1], m[2], m[3]), v), dot4_pairwise(make_float4(m[4], m[5], m[6], m[7]), v),
	                   dot4_pairwise(make_float4(m[8], m[9], m[10], m[11]), v), dot4_pairwise(make_float4(m[12], m[13], m[14], m[15]), v));
}
// v3f256_normalize (math.h:277-280)
__device__ __forceinline__ void normalize3(float &x, float &y, float &z, const uint32_t *__restrict__ lut) {
	const float ool = x86_rsqrt(dot3(x, y, z, x, y, z), lut);
	x = x * ool;
	y = y * ool;
	z = z * ool;
}

// ---- texture core (common_shader_core.h) -------------------------------------------------------

__device__ __forceinline__ int clamp_texel(int s, int hi) { // _mm256_max_epi32(_mm256_min_epi32(s, hi), 0) :31-32
	s = (s < hi) ? s : hi;
	return (s > 0) ? s : 0;
}

// decode_u32_as_color_x8 (math.h:336-344): (f32)byte * (f32)(1.0/255.0)
__device__ __forceinline__ float4 decode_u32(uint32_t c) {
	const float n = (float)(1.0 / 255.0);
	return make_float4((float)(int)(c & 0xffu) * n, (float)(int)((c >> 8) & 0xffu) * n, (float)(int)((c >> 16) & 0xffu) * n, (float)(int)(c >> 24) * n);
}

// v4f256_lerp (math.h:381-384): a*(1-f) + b*f
__device__ __forceinline__ float4 lerp4(float4 a, float4 b, float f) {
	const float g = 1.0f - f;
	return make_float4(a.x * g + b.x * f, a.y * g + b.y * f, a.z * g + b.z * f, a.w * g + b.w * f);
}

struct BilinearSetup {
	int s, t;
	float frac_s, frac_t;
};
// common prologue of bilinear_u_x8 / bilinear_f_x8 (common_shader_core.h:104-109, 144-149)
__device__ __forceinline__ BilinearSetup bilinear_setup(const TexDesc &tex, float u, float v) {
	BilinearSetup b;
	const float s_f = (float)tex.width * u + (-0.5f);
	const float t_f = (float)tex.height * (1.0f - v) + (-0.5f);
	b.s = x86_cvt_rne(floorf(s_f));
	b.t = x86_cvt_rne(floorf(t_f));
	b.frac_s = s_f - (float)b.s;
	b.frac_t = t_f - (float)b.t;
	return b;
}

// bilinear_u_x8 (common_shader_core.h:103-121) over get_texel_u_x8 (:30-36)
__device__ __forceinline__ float4 bilinear_u(const TexDesc &tex, float u, float v) {
	const BilinearSetup b = bilinear_setup(tex, u, v);
	const uint32_t *__restrict__ p = (const uint32_t *)tex.data;
	const int s0 = clamp_texel(b.s, tex.width - 1), s1 = clamp_texel(b.s + 1, tex.width - 1);
	const int t0 = clamp_texel(b.t, tex.height - 1), t1 = clamp_texel(b.t + 1, tex.height - 1);
	const float4 t00 = decode_u32(__ldg(p + t0 * tex.width + s0));
	const float4 t10 = decode_u32(__ldg(p + t0 * tex.width + s1));
	const float4 t0010 = lerp4(t00, t10, b.frac_s);
	const float4 t01 = decode_u32(__ldg(p + t1 * tex.width + s0));
	const float4 t11 = decode_u32(__ldg(p + t1 * tex.width + s1));
	const float4 t0111 = lerp4(t01, t11, b.frac_s);
	return lerp4(t0010, t0111, b.frac_t);
}

// bilinear_f_x8 (common_shader_core.h:143-161) over get_texel_f_x8 (:42-53)
__device__ __forceinline__ float4 bilinear_f(const TexDesc &tex, float u, float v) {
	const BilinearSetup b = bilinear_setup(tex, u, v);
	const float4 *__restrict__ p = (const float4 *)tex.data;
	const int s0 = clamp_texel(b.s, tex.width - 1), s1 = clamp_texel(b.s + 1, tex.width - 1);
	const int t0 = clamp_texel(b.t, tex.height - 1), t1 = clamp_texel(b.t + 1, tex.height - 1);
	const float4 t00 = __ldg(p + t0 * tex.width + s0);
	const float4 t10 = __ldg(p + t0 * tex.width + s1);
	const float4 t0010 = lerp4(t00, t10, b.frac_s);
	const float4 t01 = __ldg(p + t1 * tex.width + s0);
	const float4 t11 = __ldg(p + t1 * tex.width + s1);
	const float4 t0111 = lerp4(t01, t11, b.frac_s);
	return lerp4(t0010, t0111, b.frac_t);
}

// ---- mip-mapped trilinear sampling: an EXTENSION. The reference samples level 0 bilinearly and nothing
// else (sample_2D_u_x8 common_shader_core.h:195-199). Built from the reference's own bilinear_u_x8 per level, so that a
// texture that is not minified anywhere renders bit-identically to basic_ps.
__device__ __forceinline__ int mip_extent(int e, int level) {
	const int v = e >> level;
	return v > 0 ? v : 1;
}
__device__ __forceinline__ TexDesc mip_level_desc(const TexDesc &tex, int level) {
	if(level <= 0) return tex;
	size_t texels = 0;
	for(int l = 1; l < level; ++l) texels += (size_t)mip_extent(tex.width, l) * (size_t)mip_extent(tex.height, l);
	TexDesc d = tex;
	d.data = (const uint32_t *)tex.mips + texels;
	d.width = mip_extent(tex.width, level);
	d.height = mip_extent(tex.height, level);
	return d;
}
// (u,v) at the pixel, at its right neighbour and at the pixel below -> level of detail the way D3D11 defines it:
// rho = the longer of the two screen-space derivatives measured in texels, lod = log2(rho) clamped to the chain.
__device__ __forceinline__ float4 trilinear_u(const TexDesc &tex, float u, float v, float u_dx, float v_dx, float u_dy, float v_dy) {
	const float w = (float)tex.width, h = (float)tex.height;
	const float ax = (u_dx - u) * w, ay = (v_dx - v) * h, bx = (u_dy - u) * w, by = (v_dy - v) * h;
	const float rho2 = fmaxf(ax * ax + ay * ay, bx * bx + by * by);
	float lod = 0.5f * log2f(rho2);                               // log2(sqrt(rho2)); rho2 == 0 -> -inf -> clamped to 0
	const float max_lod = (float)(tex.mip_levels - 1);
	lod = (lod > 0.0f) ? lod : 0.0f;                              // also maps NaN to 0
	lod = (lod < max_lod) ? lod : max_lod;
	const int l0 = (int)lod;
	const float f = lod - (float)l0;
	const float4 c0 = bilinear_u(mip_level_desc(tex, l0), u, v);
	if(f == 0.0f) return c0;                                      // magnification: exactly the reference's bilinear sample
	const float4 c1 = bilinear_u(mip_level_desc(tex, l0 + 1), u, v); // f > 0 implies l0 + 1 <= mip_levels - 1
	return lerp4(c0, c1, f);
}

// sample_2D_latlon_x8 (common_shader_core.h:226-244). The three axes go through v3f256_normalize, so
// each is scaled by vrsqrtps(1) = 0x1.ffep-1 (App. A); zero components are multiplied out literally.
__device__ __forceinline__ float4 sample_latlon(const TexDesc &tex, float dx, float dy, float dz, const uint32_t *__restrict__ lut) {
	const float r1 = x86_rsqrt(1.0f, lut);
	const float zx = 0.0f * r1, zy = 0.0f * r1, zz = 1.0f * r1; // normalize((0,0,1))
	const float cos_theta = dot3(zx, zy, zz, dx, dy, dz);
	float cx = dx, cy = dy, cz = 0.0f;
	normalize3(cx, cy, cz, lut);                                  // normalize((dir.x, dir.y, 0))
	const float cos_x = dot3(1.0f * r1, 0.0f * r1, 0.0f * r1, cx, cy, cz);
	const float cos_y = dot3(0.0f * r1, 1.0f * r1, 0.0f * r1, cx, cy, cz);

	const float acos_x_over_tau = acosf(cos_x) * (float)(1.0 / (double)6.283185307f);
	float uv_x = (cos_y >= 0.0f) ? acos_x_over_tau : (1.0f - acos_x_over_tau);
	float uv_y = acosf(cos_theta) * (float)(1.0 / (double)3.141592654f);
	const bool pos_cond = cos_theta > (float)0.999;
	const bool neg_cond = cos_theta < (float)-0.999;
	uv_x = (pos_cond || neg_cond) ? 0.5f : uv_x;
	uv_y = pos_cond ? 0.0f : uv_y;
	uv_y = neg_cond ? 1.0f : uv_y;
	uv_y = 1.0f - uv_y;
	return bilinear_f(tex, uv_x, uv_y);
}

// tone map shared by env_lighting_ps.c:21 and vertex_lighting_vs.c:35: pow(1 - exp(c * -exposure), 1/2.2), exposure = 1
__device__ __forceinline__ float tone_map(float c) {
	return powf(1.0f - expf(c * (-1.0f)), (float)(1.0 / 2.2));
}

// f256_srgb_from_linear_approx (math.h:419-422): max(1.055*pow(c, 0.416666667) + (-0.055), 0); _mm256_max_ps
// returns its second operand when the first is NaN.
//
// The reference evaluates pow through Intel SVML (un-vendored, "parity unpinned"); the oracle substitutes glibc's
// powf. Here pow(c, y) = ex2(y * lg2(c)) on the SFU (MUFU.LG2 / MUFU.EX2): relative error ~3e-7, i.e. < 1e-4 of
// one 8-bit colour step, far inside the colour tolerance, and ~25x fewer instructions than libdevice powf --
// pow is the hottest arithmetic of k_tile (3 per shaded pixel). c >= 0 always (bilinear mix of u8 texels);
// c == 0 gives lg2 = -inf, ex2(-inf) = 0 like powf(0, y).
__device__ __forceinline__ float srgb_from_linear_approx(float c) {
	const float v = (float)1.055 * __powf(c, (float)0.416666667) + (float)-0.055;
	return (v > 0.0f) ? v : 0.0f;
}

// ---- vertex shaders ------------------------------------------------------------------------------
// Input vertex = 8 floats (ia.input_layout == 32 bytes): in0 = floats 0..3, in1 = floats 4..7.

// Each program is split into its SV_POSITION part and its attribute part: the geometry kernel needs the
// position of every triangle but the attributes only of the triangles that survive culling and Hi-Z.

template <int VS>
__device__ __forceinline__ float4 vs_position(float4 in0, const float *__restrict__ cb);
template <int VS>
__device__ __forceinline__ void vs_attributes(float4 in0, float4 in1, float4 pos, const float *__restrict__ cb, const TexDesc &tex, const uint32_t *__restrict__ lut, float4 &r1,
                                              float &r2x);

// passthrough_vs.c:16-25 -- Vs_Input {POSITION xyzw, COLOR xyz, pad}
template <>
__device__ __forceinline__ float4 vs_position<0>(float4 in0, const float *__restrict__) {
	return make_float4(__fmaf_rn(in0.x, 2.0f, -1.0f), __fmaf_rn(in0.y, 2.0f, -1.0f), in0.z, in0.w);
}
template <>
__device__ __forceinline__ void vs_attributes<0>(float4, float4 in1, float4, const float *__restrict__, const TexDesc &, const uint32_t *__restrict__, float4 &r1, float &r2x) {
	r1 = make_float4(in1.x, in1.y, in1.z, 0.0f);
	r2x = 0.0f;
}

// basic_vs.c:22-33 -- Vs_Input {POSITION xyz, NORMAL xyz, UV xy}
template <>
__device__ __forceinline__ float4 vs_position<1>(float4 in0, const float *__restrict__ cb) {
	return mul_m4_v4_pairwise(cb, make_float4(in0.x, in0.y, in0.z, 1.0f));
}
template <>
__device__ __forceinline__ void vs_attributes<1>(float4 in0, float4 in1, float4, const float *__restrict__, const TexDesc &, const uint32_t *__restrict__ lut, float4 &r1, float &r2x) {
	float nx = in0.w, ny = in1.x, nz = in1.y;
	normalize3(nx, ny, nz, lut);
	r1 = make_float4(nx, ny, nz, in1.z);
	r2x = in1.w;
}

// vertex_lighting_vs.c:22-39
template <>
__device__ __forceinline__ float4 vs_position<2>(float4 in0, const float *__restrict__ cb) {
	return mul_m4_v4_pairwise(cb, make_float4(in0.x, in0.y, in0.z, 1.0f));
}
template <>
__device__ __forceinline__ void vs_attributes<2>(float4 in0, float4 in1, float4, const float *__restrict__, const TexDesc &tex, const uint32_t *__restrict__ lut, float4 &r1, float &r2x) {
	float nx = in0.w, ny = in1.x, nz = in1.y;
	normalize3(nx, ny, nz, lut);
	const float4 c = sample_latlon(tex, nx, ny, nz, lut);
	r1 = make_float4(tone_map(c.x), tone_map(c.y), tone_map(c.z), in1.z);
	r2x = in1.w;
}

// fullscreen_vs.c:22-39 -- Vs_Input {POSITION xyzw, VIEW_DIR xyz, pad}; cb+16 = view_from_clip, cb+32 = world_from_view
template <>
__device__ __forceinline__ float4 vs_position<3>(float4 in0, const float *__restrict__) {
	return make_float4(__fmaf_rn(in0.x, 2.0f, -1.0f), __fmaf_rn(in0.y, 2.0f, -1.0f), in0.z, in0.w);
}
template <>
__device__ __forceinline__ void vs_attributes<3>(float4, float4, float4 pos_cs, const float *__restrict__ cb, const TexDesc &, const uint32_t *__restrict__, float4 &r1, float &r2x) {
	const float4 dir_vs = mul_m4_v4_pairwise(cb + 16, pos_cs);
	const float4 dir_ws = mul_m4_v4_pairwise(cb + 32, make_float4(dir_vs.x, dir_vs.y, dir_vs.z, 0.0f));
	r1 = make_float4(dir_ws.x, dir_ws.y, dir_ws.z, 0.0f);
	r2x = 0.0f;
}

template <int VS>
__device__ __forceinline__ VsOut run_vs(float4 in0, float4 in1, const float *__restrict__ cb, const TexDesc &tex, const uint32_t *__restrict__ lut) {
	VsOut o;
	o.r0 = vs_position<VS>(in0, cb);
	vs_attributes<VS>(in0, in1, o.r0, cb, tex, lut, o.r1, o.r2x);
	return o;
}

// ---- pixel shaders -------------------------------------------------------------------------------
// Inputs: interpolated r1 (xyzw) and r2.x. Output: SV_TARGET.xyz.

template <int PS>
__device__ __forceinline__ float3 run_ps(float4 r1, float r2x, const TexDesc &tex, const uint32_t *__restrict__ lut);

// passthrough_ps.c:13-20
template <>
__device__ __forceinline__ float3 run_ps<0>(float4 r1, float, const TexDesc &, const uint32_t *__restrict__) {
	return make_float3(r1.x, r1.y, r1.z);
}
// basic_ps.c:16-27 -- UV = (r1.w, r2.x)
template <>
__device__ __forceinline__ float3 run_ps<1>(float4 r1, float r2x, const TexDesc &tex, const uint32_t *__restrict__) {
	const float4 c = bilinear_u(tex, r1.w, r2x);
	return make_float3(srgb_from_linear_approx(c.x), srgb_from_linear_approx(c.y), srgb_from_linear_approx(c.z));
}
// env_lighting_ps.c:13-24
template <>
__device__ __forceinline__ float3 run_ps<2>(float4 r1, float, const TexDesc &tex, const uint32_t *__restrict__ lut) {
	float nx = r1.x, ny = r1.y, nz = r1.z;
	normalize3(nx, ny, nz, lut);
	const float4 c = sample_latlon(tex, nx, ny, nz, lut);
	return make_float3(tone_map(c.x), tone_map(c.y), tone_map(c.z));
}

// basic_ps with mip-mapped trilinear filtering of SRV0 (extension, see trilinear_u): UV = (r1.w, r2.x) at the pixel, at the
// pixel to its right and at the pixel below it
__device__ __forceinline__ float3 run_ps_basic_trilinear(float u, float v, float u_dx, float v_dx, float u_dy, float v_dy, const TexDesc &tex) {
	const float4 c = trilinear_u(tex, u, v, u_dx, v_dx, u_dy, v_dy);
	return make_float3(srgb_from_linear_approx(c.x), srgb_from_linear_approx(c.y), srgb_from_linear_approx(c.z));
}

// Output merger encode (main.c:1176-1178): i32 adds of shifted RNE conversions, no clamp, alpha 0.
__device__ __forceinline__ uint32_t encode_color(float3 c) {
	const uint32_t r = (uint32_t)x86_cvt_rne(c.x * 255.0f);
	const uint32_t g = (uint32_t)x86_cvt_rne(c.y * 255.0f);
	const uint32_t b = (uint32_t)x86_cvt_rne(c.z * 255.0f);
	return (r << 16) + (g << 8) + b;
}

} // namespace mlv
