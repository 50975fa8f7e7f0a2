#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE ONLY.
#
# Compiles the REFERENCE's own pipeline (not a restatement) from the sources where they lie
# under /root/reference/source, into oracle/_ref/libmalevich_ref_<W>x<H>.so -- one library per
# resolution because the reference fixes WIDTH/HEIGHT at compile time (main.c:21-22).
# No reference source is copied into the repo: the translation unit is assembled in a
# temporary directory and deleted after compilation.
#
# Recipe = SURVEY.md 8c / App. D:
#   main.c:21-283    constants, types, globals, embedded scenes
#   main.c:561-1299  pipeline stages, clears, draw_indexed, render
#   main.c:1422-1477 camera init block, wrapped as oracle_init_camera()
#   main.c:1480-1562 update()
#   main.c:547-558   sRGB->linear re-quantisation of 8-bit textures (body of load_texture)
# MSVC/ICC-isms patched (no arithmetic touched): `.m256_f32[i]` -> `[i]` (main.c:715-726),
# the 3-parameter PS function-pointer type (main.c:99) -> the 4-parameter form it is called
# with (main.c:1172), `mask.m256i_i32[i]` (common_shader_core.h:174, unused function).
# Flags: -ffp-contract=off (gcc must not fuse the reference's separate mul/add intrinsics),
# -fwrapv (edge functions rely on i32 wrap-around at 4K), -std=gnu2x (digit separators).
set -euo pipefail
REF=${MLV_REFERENCE_ROOT:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
SRC=$REF/source
if [ ! -f "$SRC/main.c" ]; then
	echo "build_ref.sh: $SRC/main.c not found (reference not mounted) -- keeping prebuilt oracle/_ref" >&2
	exit 0
fi
RES_LIST=${*:-"320x200 1200x720 1280x720 1920x1080 3840x2160"}
mkdir -p "$OUT"
TMP=$(mktemp -d)
trap 'rm -rf "$TMP"' EXIT

CFLAGS=(-std=gnu2x -O2 -mavx2 -mfma -fopenmp -ffp-contract=off -fwrapv -fPIC -w
	-fvisibility=hidden "-Dinline=static inline" -include "$HERE/shim.h" -I"$TMP")

# headers + shaders: compiled from where they lie; only common_shader_core.h needs a patched copy
sed 's/mask\.m256i_i32\[i\]/((const int*)\&mask)[i]/' "$SRC/common_shader_core.h" > "$TMP/common_shader_core.h"
cp "$SRC/math.h" "$TMP/math.h"
SHADER_OBJS=()
for s in basic_vs basic_ps passthrough_vs passthrough_ps fullscreen_vs vertex_lighting_vs env_lighting_ps; do
	# -iquote so that `#include "common_shader_core.h"` resolves to the patched copy, not the sibling file
	cp "$SRC/$s.c" "$TMP/$s.c"
	gcc "${CFLAGS[@]}" -c "$TMP/$s.c" -o "$TMP/$s.o"
	SHADER_OBJS+=("$TMP/$s.o")
done
gcc -O2 -mavx2 -fPIC -fvisibility=hidden -c "$HERE/svml_standin.c" -o "$TMP/svml_standin.o"

for RES in $RES_LIST; do
	W=${RES%x*}; H=${RES#*x}
	TU=$TMP/pipeline_${W}x${H}.c
	{
		cat "$HERE/prefix.h"
		sed -n '21,283p' "$SRC/main.c" | sed -E "s/^#define WIDTH\s.*/#define WIDTH $W/; s/^#define HEIGHT\s.*/#define HEIGHT $H/"
		sed -n '561,1299p' "$SRC/main.c"
		echo 'void oracle_init_camera(void)'
		sed -n '1422,1477p' "$SRC/main.c"
		sed -n '1480,1562p' "$SRC/main.c"
		echo '__attribute__((visibility("default"))) void ref_texture_srgb_to_linear(void *p_data, unsigned w, unsigned h) {'
		echo '	Texture2D tex = { p_data, w, h }; Texture2D *p_tex = &tex;'
		echo '	struct { i32 width, height; } header = { (i32)w, (i32)h }; bool is_in_srgb = true;'
		sed -n '547,558p' "$SRC/main.c"
		echo '}'
		cat "$HERE/harness.c"
	} | sed -E 's/\.m256_f32\[i\]/[i]/g; s/void\(\*shader\)\(void \*p_pixel_input_data.*$/void(*shader)(const void*, void*, const void*, __m256i);/' > "$TU"
	gcc "${CFLAGS[@]}" -shared "$TU" "${SHADER_OBJS[@]}" "$TMP/svml_standin.o" -o "$OUT/libmalevich_ref_${W}x${H}.so" -lm -fopenmp
	echo "built $OUT/libmalevich_ref_${W}x${H}.so"
done
