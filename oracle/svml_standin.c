/* oracle/svml_standin.c -- TEST INFRASTRUCTURE ONLY.
 *
 * The reference links Intel SVML (svml_disp.lib, reference main.c:19) for the three
 * transcendental entry points it declares at reference math.h:21-23.  SVML is a
 * proprietary, un-vendored dependency (Intel C++ 18.0, no version pin, no source),
 * so these stand-ins evaluate the same functions lane-wise with glibc's libm.
 * They only influence COLOUR (sRGB curve, tone map, lat-long angles), never
 * coverage / depth / ordering.  => "parity unpinned" for SVML; the project colour
 * tolerance (<=1/255 on >=99.9% of pixels, none >2/255) absorbs libm-vs-SVML ulps.
 */
#include <immintrin.h>
#include <math.h>

__m256 _mm256_acos_ps(__m256 a) {
	float v[8] __attribute__((aligned(32)));
	_mm256_store_ps(v, a);
	for(int i = 0; i < 8; ++i) v[i] = acosf(v[i]);
	return _mm256_load_ps(v);
}

__m256 _mm256_exp_ps(__m256 a) {
	float v[8] __attribute__((aligned(32)));
	_mm256_store_ps(v, a);
	for(int i = 0; i < 8; ++i) v[i] = expf(v[i]);
	return _mm256_load_ps(v);
}

__m256 _mm256_pow_ps(__m256 a, __m256 b) {
	float v[8] __attribute__((aligned(32)));
	float w[8] __attribute__((aligned(32)));
	_mm256_store_ps(v, a);
	_mm256_store_ps(w, b);
	for(int i = 0; i < 8; ++i) v[i] = powf(v[i], w[i]);
	return _mm256_load_ps(v);
}
