/* oracle/prefix.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Emitted (by oracle/build_ref.sh) in front of the pipeline section that is extracted at
 * build time from /root/reference/source/main.c.  It replaces the reference's
 * <windows.h>/Remotery/octarine includes (reference main.c:1-19):
 *   - UINT/HWND typedefs (from <windows.h>)
 *   - rmt_BeginCPUSample/rmt_EndCPUSample (reference main.c:663,699,737,916,984,1047,
 *     1192,1205,1220,1266,1481) become a tiny nested stage timer so that the CPU
 *     baseline can report the same per-stage split the Remotery timeline showed.
 */
#include <omp.h>
#include <stdbool.h>
#include <stdio.h>
#include <time.h>
#include "math.h"
#include "common_shader_core.h"

typedef unsigned int UINT;
typedef void *HWND;

#define REF_PROF_MAX 16
static const char *ref_prof_names[REF_PROF_MAX];
static double ref_prof_ms[REF_PROF_MAX];
static int ref_prof_count = 0;
static int ref_prof_stack[32];
static double ref_prof_t0[32];
static int ref_prof_depth = 0;

static double ref_now_ms(void) {
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static void ref_prof_begin(const char *name) {
	int slot = -1;
	for(int i = 0; i < ref_prof_count; ++i) if(ref_prof_names[i] == name || !strcmp(ref_prof_names[i], name)) { slot = i; break; }
	if(slot < 0 && ref_prof_count < REF_PROF_MAX) { slot = ref_prof_count++; ref_prof_names[slot] = name; ref_prof_ms[slot] = 0; }
	ref_prof_stack[ref_prof_depth] = slot;
	ref_prof_t0[ref_prof_depth] = ref_now_ms();
	ref_prof_depth++;
}

static void ref_prof_end(void) {
	ref_prof_depth--;
	int slot = ref_prof_stack[ref_prof_depth];
	if(slot >= 0) ref_prof_ms[slot] += ref_now_ms() - ref_prof_t0[ref_prof_depth];
}

#define rmt_BeginCPUSample(name, flags) ref_prof_begin(#name)
#define rmt_EndCPUSample() ref_prof_end()
