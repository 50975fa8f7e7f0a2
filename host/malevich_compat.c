/* host/malevich_compat.c -- the reference's entry points and global pipeline state on top of the C-ABI.
 *
 * clear_render_target_view / clear_depth_stencil_view / draw_indexed keep the reference's signatures
 * (main.c:1191, 1204, 1219). draw_indexed snapshots the global `graphics_pipeline` exactly when the reference
 * would read it, maps the bound shader function pointers to device shader ids by identity, and uploads vertex /
 * index buffers and textures on first use, cached by host pointer (the reference never frees or rewrites its
 * assets, main.c:526-559). Errors abort like the reference's asserts / error() (main.c:451-465).
 */
#include "malevich_compat.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/malevich_b200.h"

Pipeline graphics_pipeline;
Stats stats;

/* The descriptors only need distinct addresses: the device selects the __device__ shader by id. The sizes are the
 * reference's sizeof(Vs_Input) / sizeof(Vs_Output) of its 8-wide SoA blocks (e.g. basic_vs.c:3-14, 35). */
static void tag_passthrough_vs(void) {}
static void tag_basic_vs(void) {}
static void tag_vertex_lighting_vs(void) {}
static void tag_fullscreen_vs(void) {}
static void tag_passthrough_ps(void) {}
static void tag_basic_ps(void) {}
static void tag_env_lighting_ps(void) {}
VertexShader passthrough_vs = { 256, 384, (void (*)())tag_passthrough_vs };
VertexShader basic_vs = { 256, 384, (void (*)())tag_basic_vs };
VertexShader vertex_lighting_vs = { 256, 384, (void (*)())tag_vertex_lighting_vs };
VertexShader fullscreen_vs = { 256, 384, (void (*)())tag_fullscreen_vs };
PixelShader passthrough_ps = { (void (*)())tag_passthrough_ps };
PixelShader basic_ps = { (void (*)())tag_basic_ps };
PixelShader env_lighting_ps = { (void (*)())tag_env_lighting_ps };

/* Device copies of the host's buffers and textures, keyed by host pointer; the tables grow on demand. The reference never
 * frees or rewrites its assets (main.c:526-559); a host that does calls malevich_gpu_invalidate(pointer). An index buffer's
 * entry also keeps the largest index of its first `index_count` entries, so that the scan that sizes the vertex upload
 * runs once per (buffer, count) and not once per draw (30 M indices per frame on BASELINE config 5). */
static mlv_device *g_dev;
static unsigned g_width, g_height;
typedef struct BufferEntry { const void *host; size_t bytes; mlv_buffer *dev; UINT scanned_count; u32 max_index; } BufferEntry;
typedef struct TextureEntry { const void *host; mlv_texture *dev; } TextureEntry;
static BufferEntry *g_buffers;
static TextureEntry *g_textures;
static size_t g_buffer_slots, g_texture_slots;

static void die(const char *what) {
	fprintf(stderr, "malevich_compat: %s failed with error: %s\n", what, mlv_last_error_string());
	abort();
}
#define CHECK(call)                        \
	do {                                   \
		if((call) != MLV_OK) die(#call);   \
	} while(0)

int malevich_gpu_init(unsigned width, unsigned height) {
	if(g_dev) return 0;
	mlv_device_desc d;
	memset(&d, 0, sizeof(d));
	d.width = width;
	d.height = height;
	d.cuda_device = -1;
	/* MLV_NUM_GPUS=N: a device group -- the reference's render() drives N GPUs through the same three entry points */
	const char *n = getenv("MLV_NUM_GPUS");
	if(n && atoi(n) > 1) {
		d.num_gpus = (uint32_t)atoi(n);
		d.cuda_device = 0;
		if(getenv("MLV_GROUP_SAME_GPU") && atoi(getenv("MLV_GROUP_SAME_GPU"))) d.flags |= MLV_DEVICE_GROUP_SAME_GPU;
		if(getenv("MLV_GROUP_NCCL") && atoi(getenv("MLV_GROUP_NCCL"))) d.flags |= MLV_DEVICE_GROUP_NCCL;
		if(getenv("MLV_GROUP_STRIPE")) d.stripe_height_tiles = (uint32_t)atoi(getenv("MLV_GROUP_STRIPE"));
	}
	if(mlv_create_device(&d, &g_dev) != MLV_OK) {
		fprintf(stderr, "malevich_compat: %s\n", mlv_last_error_string());
		return 1;
	}
	g_width = width;
	g_height = height;
	return 0;
}

void malevich_gpu_shutdown(void) {
	if(!g_dev) return;
	for(size_t i = 0; i < g_buffer_slots; ++i)
		if(g_buffers[i].dev) mlv_release_buffer(g_dev, g_buffers[i].dev);
	for(size_t i = 0; i < g_texture_slots; ++i)
		if(g_textures[i].dev) mlv_release_texture(g_dev, g_textures[i].dev);
	free(g_buffers);
	free(g_textures);
	g_buffers = NULL, g_textures = NULL;
	g_buffer_slots = g_texture_slots = 0;
	mlv_destroy_device(g_dev);
	g_dev = NULL;
}

static void *grow(void *table, size_t *slots, size_t entry_bytes) {
	const size_t n = *slots ? *slots * 2 : 64;
	void *p = realloc(table, n * entry_bytes);
	if(!p) die("host allocation");
	memset((char *)p + *slots * entry_bytes, 0, (n - *slots) * entry_bytes);
	*slots = n;
	return p;
}

static BufferEntry *buffer_for(const void *host, size_t bytes, int kind) {
	BufferEntry *free_slot = NULL;
	for(size_t i = 0; i < g_buffer_slots; ++i) {
		if(g_buffers[i].host == host) {
			if(g_buffers[i].bytes >= bytes) return &g_buffers[i];
			mlv_release_buffer(g_dev, g_buffers[i].dev); /* the same pointer now addresses more data: upload again */
			memset(&g_buffers[i], 0, sizeof(BufferEntry));
		}
		if(!g_buffers[i].host && !free_slot) free_slot = &g_buffers[i];
	}
	if(!free_slot) {
		const size_t old = g_buffer_slots;
		g_buffers = (BufferEntry *)grow(g_buffers, &g_buffer_slots, sizeof(BufferEntry));
		free_slot = &g_buffers[old];
	}
	CHECK(mlv_create_buffer(g_dev, host, bytes, kind, &free_slot->dev));
	free_slot->host = host;
	free_slot->bytes = bytes;
	return free_slot;
}

static mlv_texture *texture_for(const Texture2D *t, int format) {
	if(!t || !t->p_data) return NULL;
	TextureEntry *free_slot = NULL;
	for(size_t i = 0; i < g_texture_slots; ++i) {
		if(g_textures[i].host == t->p_data) return g_textures[i].dev;
		if(!g_textures[i].host && !free_slot) free_slot = &g_textures[i];
	}
	if(!free_slot) {
		const size_t old = g_texture_slots;
		g_textures = (TextureEntry *)grow(g_textures, &g_texture_slots, sizeof(TextureEntry));
		free_slot = &g_textures[old];
	}
	CHECK(mlv_create_texture2d(g_dev, t->p_data, t->width, t->height, format, &free_slot->dev));
	free_slot->host = t->p_data;
	return free_slot->dev;
}

void malevich_gpu_invalidate(const void *host_pointer) {
	if(!g_dev) return;
	for(size_t i = 0; i < g_buffer_slots; ++i)
		if(g_buffers[i].host == host_pointer) {
			mlv_release_buffer(g_dev, g_buffers[i].dev);
			memset(&g_buffers[i], 0, sizeof(BufferEntry));
		}
	for(size_t i = 0; i < g_texture_slots; ++i)
		if(g_textures[i].host == host_pointer) {
			mlv_release_texture(g_dev, g_textures[i].dev);
			memset(&g_textures[i], 0, sizeof(TextureEntry));
		}
}

static int vs_id(void (*f)()) {
	if(f == passthrough_vs.vs_main) return MLV_VS_PASSTHROUGH;
	if(f == basic_vs.vs_main) return MLV_VS_BASIC;
	if(f == vertex_lighting_vs.vs_main) return MLV_VS_VERTEX_LIGHTING;
	if(f == fullscreen_vs.vs_main) return MLV_VS_FULLSCREEN;
	die("unknown vertex shader");
	return -1;
}
static int ps_id(void (*f)()) {
	if(f == passthrough_ps.ps_main) return MLV_PS_PASSTHROUGH;
	if(f == basic_ps.ps_main) return MLV_PS_BASIC;
	if(f == env_lighting_ps.ps_main) return MLV_PS_ENV_LIGHTING;
	die("unknown pixel shader");
	return -1;
}

void clear_render_target_view(const f32 *p_clear_color) {
	if(!g_dev) die("malevich_gpu_init not called");
	CHECK(mlv_clear_render_target_view(g_dev, p_clear_color));
}

void clear_depth_stencil_view(const f32 depth) {
	if(!g_dev) die("malevich_gpu_init not called");
	CHECK(mlv_clear_depth_stencil_view(g_dev, depth));
}

void draw_indexed(UINT index_count) {
	if(!g_dev) die("malevich_gpu_init not called");
	const Pipeline *gp = &graphics_pipeline;
	const int vs = vs_id(gp->vs.shader), ps = ps_id(gp->ps.shader);
	/* the reference has no vertex count (vertex_count = index_count, main.c:673): the largest index sizes the upload;
	 * scanned once per (index buffer, count) and kept with the buffer's entry */
	BufferEntry *ib = buffer_for(gp->ia.p_index_buffer, (size_t)index_count * 4, MLV_BUFFER_INDEX);
	if(ib->scanned_count != index_count) {
		u32 max_index = 0;
		for(UINT i = 0; i < index_count; ++i)
			if(gp->ia.p_index_buffer[i] > max_index) max_index = gp->ia.p_index_buffer[i];
		ib->max_index = max_index;
		ib->scanned_count = index_count;
	}
	const u32 max_index = ib->max_index;
	mlv_buffer *ib_dev = ib->dev; /* (the table may move when the vertex buffer takes a new slot) */
	CHECK(mlv_ia_set_primitive_topology(g_dev, gp->ia.primitive_topology));
	CHECK(mlv_ia_set_input_layout(g_dev, gp->ia.input_layout));
	CHECK(mlv_ia_set_vertex_buffer(g_dev, buffer_for(gp->ia.p_vertex_buffer, (size_t)(max_index + 1) * gp->ia.input_layout, MLV_BUFFER_VERTEX)->dev));
	CHECK(mlv_ia_set_index_buffer(g_dev, ib_dev));
	CHECK(mlv_vs_set_shader(g_dev, vs));
	CHECK(mlv_ps_set_shader(g_dev, ps));
	if(gp->vs.p_constant_buffers[0]) CHECK(mlv_vs_set_constant_buffer(g_dev, 0, gp->vs.p_constant_buffers[0], 192)); /* PerFrameCB main.c:169-173 */
	CHECK(mlv_vs_set_shader_resource(g_dev, 0, vs == MLV_VS_VERTEX_LIGHTING ? texture_for((const Texture2D *)gp->vs.p_shader_resource_views[0], MLV_FORMAT_R32G32B32A32_FLOAT) : NULL));
	CHECK(mlv_ps_set_shader_resource(g_dev, 0, ps == MLV_PS_PASSTHROUGH ? NULL
	                                               : texture_for((const Texture2D *)gp->ps.p_shader_resource_views[0],
	                                                             ps == MLV_PS_BASIC ? MLV_FORMAT_R8G8B8A8_UNORM : MLV_FORMAT_R32G32B32A32_FLOAT)));
	mlv_viewport vp = { gp->rs.viewport.top_left_x, gp->rs.viewport.top_left_y, gp->rs.viewport.width, gp->rs.viewport.height, gp->rs.viewport.min_depth, gp->rs.viewport.max_depth };
	CHECK(mlv_rs_set_viewport(g_dev, &vp));
	CHECK(mlv_draw_indexed(g_dev, index_count));
}

void malevich_gpu_present(void) {
	if(!g_dev) die("malevich_gpu_init not called");
	CHECK(mlv_present_readback(g_dev, graphics_pipeline.om.p_colors, graphics_pipeline.om.p_depth));
	mlv_stats s;
	CHECK(mlv_get_stats(g_dev, &s));
	CHECK(mlv_reset_stats(g_dev)); /* memset(&stats, 0) at the top of render(), main.c:1268 */
	const f32 frame_time = stats.frame_time;
	memcpy(&stats, &s, sizeof(stats));
	stats.frame_time = frame_time;
}
