/* host/malevich_compat.h -- the reference's L4 boundary, re-declared for a C host that links the B200 library.
 *
 * A host written against the reference (its `render()`, main.c:1265-1299) touches exactly this surface: the global
 * `Pipeline graphics_pipeline` whose fields it writes directly (main.c:71-115, 222), the shader descriptors it
 * copies function pointers from (main.c:46-52, common_shader_core.h:10-18), `Texture2D` (common_shader_core.h:20-24)
 * and the three entry points (main.c:1191, 1204, 1219). The declarations below keep the reference's names, field
 * order and meaning so such a host compiles unchanged; host/malevich_compat.c implements them on top of
 * include/malevich_b200.h. Types are restated here (not included from the reference) so the host builds without it.
 */
#ifndef MALEVICH_COMPAT_H
#define MALEVICH_COMPAT_H

#include <stdint.h>

typedef uint8_t u8;
typedef uint32_t u32;
typedef int32_t i32;
typedef float f32;
typedef unsigned int UINT;

#define COMMONSHADER_CONSTANT_BUFFER_HW_SLOT_COUNT 16 /* main.c:41 */
#define COMMONSHADER_INPUT_RESOURCE_REGISTER_COUNT 16 /* main.c:42 */
#define VECTOR_WIDTH 8                                /* main.c:26 */
#define MAX_OBJECT_COUNT_PER_SCENE 8                  /* main.c:44 */

typedef struct VertexShader { /* common_shader_core.h:10-14 */
	unsigned int in_vertex_size;
	unsigned int out_vertex_size;
	void (*vs_main)();
} VertexShader;

typedef struct PixelShader { /* common_shader_core.h:16-18 */
	void (*ps_main)();
} PixelShader;

typedef struct Texture2D { /* common_shader_core.h:20-24 */
	void *p_data;
	unsigned int width;
	unsigned int height;
} Texture2D;

typedef enum PrimitiveTopology { PRIMITIVE_TOPOLOGY_UNDEFINED = 0, PRIMITIVE_TOPOLOGY_TRIANGLELIST = 1 } PrimitiveTopology; /* main.c:66-69 */

typedef struct IA { /* main.c:71-76 */
	u32 *p_index_buffer;
	void *p_vertex_buffer;
	u32 input_layout;
	PrimitiveTopology primitive_topology;
} IA;

typedef struct VS { /* main.c:78-83 */
	void (*shader)();
	u8 output_register_count;
	void *p_constant_buffers[COMMONSHADER_CONSTANT_BUFFER_HW_SLOT_COUNT];
	void *p_shader_resource_views[COMMONSHADER_INPUT_RESOURCE_REGISTER_COUNT];
} VS;

typedef struct Viewport { /* main.c:85-92 */
	f32 top_left_x, top_left_y, width, height, min_depth, max_depth;
} Viewport;

typedef struct RS { Viewport viewport; } RS; /* main.c:94-96 */

typedef struct PS { /* main.c:98-101 */
	void (*shader)();
	void *p_shader_resource_views[COMMONSHADER_INPUT_RESOURCE_REGISTER_COUNT];
} PS;

typedef struct OM { /* main.c:103-107: filled by malevich_gpu_present(), not rendered into directly */
	u32 *p_colors;
	f32 *p_depth;
} OM;

typedef struct Pipeline { IA ia; VS vs; RS rs; PS ps; OM om; } Pipeline; /* main.c:109-115 */

typedef struct Stats { /* main.c:199-206 */
	f32 frame_time;
	u32 vertex_count, input_triangle_count, assembled_triangle_count, active_bin_count, total_triangle_count_in_bins;
} Stats;

extern Pipeline graphics_pipeline; /* main.c:222 */
extern Stats stats;                /* main.c:231 */
extern VertexShader passthrough_vs, basic_vs, vertex_lighting_vs, fullscreen_vs; /* main.c:46-52 */
extern PixelShader passthrough_ps, basic_ps, env_lighting_ps;

/* The reference fixes the render-target size at compile time (WIDTH/HEIGHT main.c:21-22); here it is chosen once,
 * before the first clear or draw. Returns 0 on success. */
int malevich_gpu_init(unsigned width, unsigned height);
void malevich_gpu_shutdown(void);
/* The device copies of vertex / index buffers and textures are cached by host pointer (the reference never frees or
 * rewrites an asset). A host that rewrites one in place, or frees it and reuses the address, says so here; the next draw
 * that binds the pointer uploads it again. */
void malevich_gpu_invalidate(const void *host_pointer);

void clear_render_target_view(const f32 *p_clear_color); /* main.c:1191 */
void clear_depth_stencil_view(const f32 depth);          /* main.c:1204 */
void draw_indexed(UINT index_count);                     /* main.c:1219 */

/* Replaces reading the static frame_buffer/depth_buffer (main.c:35-36, 314): copies the frame into om.p_colors /
 * om.p_depth (row-major y*W+x) and refreshes `stats`. */
void malevich_gpu_present(void);

/* For pixel shaders that take an RGBA32F texture (env_lighting_ps) the host must say so, because Texture2D carries no
 * format (the reference infers it from the sampler used, common_shader_core.h:30,42). Default: inferred from the PS. */

#endif
