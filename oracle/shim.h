/* oracle/shim.h -- TEST INFRASTRUCTURE ONLY (never part of the product path).
 *
 * Forced-include shim (gcc -include) that lets the reference's MSVC/ICC-dialect
 * pipeline sources compile unmodified-in-arithmetic under gcc on Linux.
 * Nothing in here touches arithmetic:
 *   - __cdecl            : MSVC calling-convention keyword (reference math.h:21-23)
 *   - min/max            : lower-case macros that <windows.h> supplies in the original
 *                          (used at reference main.c:739,1504-1505, math.h:415)
 *   - malloc             : the reference dereferences malloc'ed stage buffers as __m256
 *                          (main.c:676,680,709); gcc emits aligned vmovaps for those,
 *                          so the buffers must be 32-byte aligned like MSVC's x64
 *                          allocator + ICC unaligned moves happened to tolerate.
 */
#ifndef MLV_ORACLE_SHIM_H
#define MLV_ORACLE_SHIM_H
#include <assert.h>
#include <stdlib.h>
#include <string.h>
#define __cdecl
#ifndef max
#define max(a, b) (((a) > (b)) ? (a) : (b))
#define min(a, b) (((a) < (b)) ? (a) : (b))
#endif
#define malloc(n) aligned_alloc(32, (((size_t)(n)) + 31) & ~(size_t)31)
#endif
