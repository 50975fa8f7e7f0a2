/* host/render_host.c -- headless C host: the reference's scene table + render() driving the B200 pipeline.
 *
 * This is the caller of the hot path, written in the reference's own language against the
 * reference's own interface: `render()` below is the reference's render() (main.c:1265-1299) statement for
 * statement -- it writes the fields of the global `graphics_pipeline` and calls clear_render_target_view,
 * clear_depth_stencil_view and draw_indexed -- and knows nothing about CUDA. host/malevich_compat.c forwards those
 * three entry points to the C-ABI of include/malevich_b200.h. What replaces WinMain / the message loop / the GDI blit
 * (main.c:286-447, 1568-1607): read a scene file, render `frames` frames, write the frame buffer.
 *
 *   render_host <scene.bin> <out.bin> [frames]
 *
 * scene.bin (little-endian, written by tests/test_c_host.py from malevich_b200.scenes):
 *   "MLVSCENE", u32 width, height, num_objects, 0, f32 per_frame_cb[48],
 *   per object: u32 vs, ps, vertex_count, index_count, tex_kind (0 none, 1 R8G8B8A8, 2 RGBA32F), tex_w, tex_h, 0,
 *               vertices (32 B each), indices (u32), texels
 * out.bin: u32 width, height, colours (u32 x W*H), depths (f32 x W*H), Stats (6 x 4 B), f64 milliseconds per frame
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "malevich_compat.h"

typedef struct MeshHeader { /* main.c:54-58 */
	uint32_t size, vertex_count, index_count;
} MeshHeader;

typedef struct Mesh { /* main.c:60-64 */
	MeshHeader header;
	void *p_vertex_buffer;
	u32 *p_index_buffer;
} Mesh;

typedef struct Scene { /* main.c:214-220 */
	Mesh a_meshes[MAX_OBJECT_COUNT_PER_SCENE];
	Texture2D a_textures[MAX_OBJECT_COUNT_PER_SCENE];
	VertexShader a_vertex_shaders[MAX_OBJECT_COUNT_PER_SCENE];
	PixelShader a_pixel_shaders[MAX_OBJECT_COUNT_PER_SCENE];
	u32 num_objects;
} Scene;

static Scene scene;
static float per_frame_cb[48]; /* PerFrameCB main.c:169-173 */
static int frame_width, frame_height;
static u32 *frame_buffer;
static f32 *depth_buffer;

/* render() main.c:1265-1299 */
static void render(f32 delta_t_ms) {
	memset(&stats, 0, sizeof(Stats));
	stats.frame_time = delta_t_ms;

	const f32 clear_color[4] = { (f32)227 / 255, (f32)223 / 255, (f32)216 / 255, 0.f };
	clear_render_target_view(clear_color);
	clear_depth_stencil_view(0.0);

	/* Set the common part of the pipeline */
	graphics_pipeline.ia.primitive_topology = PRIMITIVE_TOPOLOGY_TRIANGLELIST;
	Viewport viewport = { 0.f, 0.f, (f32)frame_width, (f32)frame_height, 0.f, 1.f };
	graphics_pipeline.rs.viewport = viewport;
	graphics_pipeline.om.p_colors = frame_buffer;
	graphics_pipeline.om.p_depth = depth_buffer;
	graphics_pipeline.vs.p_constant_buffers[0] = per_frame_cb;

	Scene *p_scene = &scene;
	for(i32 object_index = 0; object_index < (i32)p_scene->num_objects; ++object_index) {
		/* Set the draw call specific part of the pipeline */
		graphics_pipeline.ia.input_layout = p_scene->a_vertex_shaders[object_index].in_vertex_size / VECTOR_WIDTH;
		graphics_pipeline.vs.output_register_count = p_scene->a_vertex_shaders[object_index].out_vertex_size / (16 * VECTOR_WIDTH);
		graphics_pipeline.vs.shader = p_scene->a_vertex_shaders[object_index].vs_main;
		graphics_pipeline.ps.shader = p_scene->a_pixel_shaders[object_index].ps_main;

		graphics_pipeline.ia.p_index_buffer = p_scene->a_meshes[object_index].p_index_buffer;
		graphics_pipeline.ia.p_vertex_buffer = p_scene->a_meshes[object_index].p_vertex_buffer;
		graphics_pipeline.vs.p_shader_resource_views[0] = &p_scene->a_textures[object_index];
		graphics_pipeline.ps.p_shader_resource_views[0] = &p_scene->a_textures[object_index];
		draw_indexed(p_scene->a_meshes[object_index].header.index_count);
	}
}

static void *read_exact(FILE *f, size_t bytes) {
	void *p = malloc(bytes ? bytes : 1);
	if(!p || fread(p, 1, bytes, f) != bytes) {
		fprintf(stderr, "render_host: short read\n");
		exit(2);
	}
	return p;
}

int main(int argc, char **argv) {
	if(argc < 3) {
		fprintf(stderr, "usage: render_host <scene.bin> <out.bin> [frames]\n");
		return 2;
	}
	const int frames = argc > 3 ? atoi(argv[3]) : 1;
	FILE *f = fopen(argv[1], "rb");
	if(!f) {
		perror(argv[1]);
		return 2;
	}
	char magic[8];
	uint32_t head[4];
	if(fread(magic, 1, 8, f) != 8 || memcmp(magic, "MLVSCENE", 8) || fread(head, 4, 4, f) != 4 || fread(per_frame_cb, 4, 48, f) != 48 || head[2] > MAX_OBJECT_COUNT_PER_SCENE) {
		fprintf(stderr, "render_host: bad scene file\n");
		return 2;
	}
	frame_width = (int)head[0];
	frame_height = (int)head[1];
	scene.num_objects = head[2];
	static const VertexShader *vs_table[4] = { &passthrough_vs, &basic_vs, &vertex_lighting_vs, &fullscreen_vs };
	static const PixelShader *ps_table[3] = { &passthrough_ps, &basic_ps, &env_lighting_ps };
	for(u32 i = 0; i < scene.num_objects; ++i) {
		uint32_t o[8];
		if(fread(o, 4, 8, f) != 8 || o[0] > 3 || o[1] > 2) {
			fprintf(stderr, "render_host: bad object header\n");
			return 2;
		}
		scene.a_vertex_shaders[i] = *vs_table[o[0]];
		scene.a_pixel_shaders[i] = *ps_table[o[1]];
		scene.a_meshes[i].header.vertex_count = o[2];
		scene.a_meshes[i].header.index_count = o[3];
		scene.a_meshes[i].p_vertex_buffer = read_exact(f, (size_t)o[2] * 32);
		scene.a_meshes[i].p_index_buffer = (u32 *)read_exact(f, (size_t)o[3] * 4);
		scene.a_textures[i].width = o[5];
		scene.a_textures[i].height = o[6];
		scene.a_textures[i].p_data = o[4] ? read_exact(f, (size_t)o[5] * o[6] * (o[4] == 1 ? 4 : 16)) : NULL;
	}
	fclose(f);

	frame_buffer = (u32 *)malloc((size_t)frame_width * frame_height * 4);
	depth_buffer = (f32 *)malloc((size_t)frame_width * frame_height * 4);
	if(malevich_gpu_init((unsigned)frame_width, (unsigned)frame_height)) return 1;

	struct timespec t0, t1;
	double ms = 0.0;
	for(int i = 0; i < frames; ++i) {
		if(i == frames - 1) clock_gettime(CLOCK_MONOTONIC, &t0);
		render(16.f);
		malevich_gpu_present(); /* present() main.c:1301-1306 */
		if(i == frames - 1) {
			clock_gettime(CLOCK_MONOTONIC, &t1);
			ms = (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6;
		}
	}

	f = fopen(argv[2], "wb");
	if(!f) {
		perror(argv[2]);
		return 2;
	}
	uint32_t wh[2] = { (uint32_t)frame_width, (uint32_t)frame_height };
	fwrite(wh, 4, 2, f);
	fwrite(frame_buffer, 4, (size_t)frame_width * frame_height, f);
	fwrite(depth_buffer, 4, (size_t)frame_width * frame_height, f);
	fwrite(&stats, sizeof(Stats), 1, f);
	fwrite(&ms, 8, 1, f);
	fclose(f);
	printf("render_host: %dx%d, %u objects, %d frame(s), last frame %.3f ms, %u triangles in, %u assembled\n", frame_width, frame_height, scene.num_objects, frames, ms,
	       stats.input_triangle_count, stats.assembled_triangle_count);
	malevich_gpu_shutdown();
	return 0;
}
