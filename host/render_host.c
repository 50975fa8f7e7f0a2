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
 *   render_host --assets <dir> --scene toon|ftm|suprematism --size <W>x<H> [--pose x,y,z,yaw,pitch] [--cb <192-byte file>]
 *               [--ppm <frame.ppm>] <out.bin> [frames]
 *   render_host --list-assets <dir>        (no GPU: every *_mesh.octrn / image of the directory through the C reader)
 *
 * The second form is the reference's own start-up: init()'s scene table (main.c:1317-1420) filled by load_mesh /
 * load_texture (main.c:526-559) from `.octrn` files through host/octrn.c, and update()'s camera (main.c:1422-1562,
 * no input devices) unless --cb supplies the PerFrameCB bytes.
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

#include <math.h>

#include "malevich_compat.h"
#include "octrn.h"

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

/* ---- asset loaders (main.c:526-559) on top of host/octrn.c ---------------------------------------------------------- */
static char assets_dir[1024] = "../assets";

static void asset_path(char *out, size_t cap, const char *name) { snprintf(out, cap, "%s/%s.octrn", assets_dir, name); }

static void load_mesh(const char *p_mesh_name, Mesh *p_mesh) { /* main.c:526-536 */
	char path[1200];
	asset_path(path, sizeof(path), p_mesh_name);
	OctrnMeshHeader h;
	void *p_data = NULL;
	if(octrn_read_mesh(path, &h, &p_data)) {
		fprintf(stderr, "render_host: %s\n", octrn_last_error());
		exit(2);
	}
	p_mesh->header.size = h.size, p_mesh->header.vertex_count = h.vertex_count, p_mesh->header.index_count = h.index_count;
	const u32 vertex_size = sizeof(float) * 8;
	p_mesh->p_vertex_buffer = p_data;
	p_mesh->p_index_buffer = (u32 *)(((uint8_t *)p_data) + (size_t)h.vertex_count * vertex_size);
}

/* srgb_to_linear (math.h:386-395): f32 in, double arithmetic, f32 out */
static f32 srgb_to_linear(f32 v) {
	f32 result;
	if(v <= 0.04045) result = (f32)(v / 12.92);
	else result = (f32)pow(((v + 0.055) / 1.055), 2.4);
	return result;
}

static void load_texture(const char *p_tex_name, Texture2D *p_tex, int is_in_srgb) { /* main.c:538-559 */
	char path[1200];
	asset_path(path, sizeof(path), p_tex_name);
	OctrnImageHeader h;
	if(octrn_read_image(path, &h, &p_tex->p_data)) {
		fprintf(stderr, "render_host: %s\n", octrn_last_error());
		exit(2);
	}
	p_tex->width = h.width;
	p_tex->height = h.height;
	if(is_in_srgb) { /* get rid of the gamma mapping: decode (math.h:326-334), curve, encode by truncation (math.h:322-324); alpha too */
		u32 lut[256];
		const f32 normalizer = (f32)(1.0 / 255.0);
		for(u32 b = 0; b < 256; ++b) lut[b] = (u32)(srgb_to_linear((f32)b * normalizer) * 255.f);
		u32 *t = (u32 *)p_tex->p_data;
		for(size_t i = 0; i < (size_t)h.width * h.height; ++i)
			t[i] = lut[t[i] & 0xff] | (lut[(t[i] >> 8) & 0xff] << 8) | (lut[(t[i] >> 16) & 0xff] << 16) | (lut[t[i] >> 24] << 24);
	}
}

/* SUPREMATISM's embedded geometry (main.c:232-254): 11 vertices of (pos4, colour3, pad), 24 indices of which the last
 * 9 are (0,0,0) padding triangles. These are the reference's INPUT DATA (tests compare them with the compiled arrays). */
static float suprematist_vertex_buffer[11][8] = {
	{ 0.34107f, 0.12215f, 0.5f, 1.0f, 0.07500f, 0.08200f, 0.06300f, 0.0f }, { 0.95357f, 0.12500f, 0.5f, 1.0f, 0.07500f, 0.08200f, 0.06300f, 0.0f },
	{ 0.96250f, 0.86931f, 0.5f, 1.0f, 0.07500f, 0.08200f, 0.06300f, 0.0f }, { 0.33928f, 0.86505f, 0.5f, 1.0f, 0.07500f, 0.08200f, 0.06300f, 0.0f },
	{ 0.09464f, 0.12500f, 0.75f, 1.0f, 0.14100f, 0.29000f, 0.60800f, 0.0f }, { 0.69285f, 0.39772f, 0.75f, 1.0f, 0.14100f, 0.29000f, 0.60800f, 0.0f },
	{ 0.09107f, 0.60937f, 0.75f, 1.0f, 0.14100f, 0.29000f, 0.60800f, 0.0f }, { 0.00000f, 0.00000f, 0.25f, 1.0f, 0.96100f, 0.96100f, 0.92900f, 0.0f },
	{ 1.00000f, 0.00000f, 0.25f, 1.0f, 0.96100f, 0.96100f, 0.92900f, 0.0f }, { 1.00000f, 1.00000f, 0.25f, 1.0f, 0.96100f, 0.96100f, 0.92900f, 0.0f },
	{ 0.00000f, 1.00000f, 0.25f, 1.0f, 0.96100f, 0.96100f, 0.92900f, 0.0f },
};
static u32 suprematist_index_buffer[24] = { 0, 1, 2, 2, 3, 0, 4, 5, 6, 7, 8, 9, 9, 10, 7, 0, 0, 0, 0, 0, 0, 0, 0, 0 };

/* init()'s scene table (main.c:1317-1420) for the scenes whose assets can exist here */
static int init_scene(const char *name) {
	u32 n = 0;
	if(!strcmp(name, "ftm")) { /* main.c:1317-1362 */
		static const char *parts[7] = { "piedras", "madera", "leaves", "dec", "roof", "ground", "sky" };
		for(; n < 7; ++n) {
			char mesh[64], tex[64];
			snprintf(mesh, sizeof(mesh), "ftm_%s_mesh", parts[n]);
			snprintf(tex, sizeof(tex), "ftm_%s_tex", parts[n]);
			load_mesh(mesh, scene.a_meshes + n);
			load_texture(tex, scene.a_textures + n, 1);
			scene.a_vertex_shaders[n] = basic_vs;
			scene.a_pixel_shaders[n] = basic_ps;
		}
	} else if(!strcmp(name, "toon")) { /* main.c:1364-1379 */
		static const char *parts[2] = { "house", "sky" };
		for(; n < 2; ++n) {
			char mesh[64], tex[64];
			snprintf(mesh, sizeof(mesh), "toon_%s_mesh", parts[n]);
			snprintf(tex, sizeof(tex), "toon_%s_tex", parts[n]);
			load_mesh(mesh, scene.a_meshes + n);
			load_texture(tex, scene.a_textures + n, 1);
			scene.a_vertex_shaders[n] = basic_vs;
			scene.a_pixel_shaders[n] = basic_ps;
		}
	} else if(!strcmp(name, "suprematism")) { /* main.c:1381-1389 */
		scene.a_meshes[0].p_vertex_buffer = suprematist_vertex_buffer;
		scene.a_meshes[0].p_index_buffer = suprematist_index_buffer;
		scene.a_meshes[0].header.index_count = 24;
		scene.a_vertex_shaders[0] = passthrough_vs;
		scene.a_pixel_shaders[0] = passthrough_ps;
		n = 1;
	} else {
		return 1;
	}
	scene.num_objects = n;
	return 0;
}

/* ---- camera: init() + update() without input devices (main.c:1422-1562) -> PerFrameCB ------------------------------- */
typedef struct M4 { float m[4][4]; } M4;
static M4 m4_mul(const M4 *a, const M4 *b) { /* m4x4f32_mul_m4x4f32 math.h:188-197: serial fp32 dot products */
	M4 r;
	for(int i = 0; i < 4; ++i)
		for(int j = 0; j < 4; ++j) r.m[i][j] = ((a->m[i][0] * b->m[0][j] + a->m[i][1] * b->m[1][j]) + a->m[i][2] * b->m[2][j]) + a->m[i][3] * b->m[3][j];
	return r;
}
static M4 m4_inverse(const M4 *a) { /* cofactor expansion in double (the reference spells out a closed form, math.h:282-320; agreement to a few ulp) */
	double m[4][4], c[4][4];
	for(int i = 0; i < 4; ++i)
		for(int j = 0; j < 4; ++j) m[i][j] = a->m[i][j];
	for(int i = 0; i < 4; ++i)
		for(int j = 0; j < 4; ++j) {
			double s[3][3];
			for(int r = 0, rr = 0; r < 4; ++r) {
				if(r == i) continue;
				for(int q = 0, qq = 0; q < 4; ++q) {
					if(q == j) continue;
					s[rr][qq++] = m[r][q];
				}
				++rr;
			}
			const double det3 = s[0][0] * (s[1][1] * s[2][2] - s[1][2] * s[2][1]) - s[0][1] * (s[1][0] * s[2][2] - s[1][2] * s[2][0]) + s[0][2] * (s[1][0] * s[2][1] - s[1][1] * s[2][0]);
			c[i][j] = ((i + j) & 1) ? -det3 : det3;
		}
	const double det = m[0][0] * c[0][0] + m[0][1] * c[0][1] + m[0][2] * c[0][2] + m[0][3] * c[0][3];
	M4 r;
	for(int i = 0; i < 4; ++i)
		for(int j = 0; j < 4; ++j) r.m[i][j] = (float)(c[j][i] / det);
	return r;
}
static void update_camera(float px, float py, float pz, float yaw_rad, float pitch_rad) {
	const float PI_F = 3.141592654f, TAU_F = 6.283185307f, PI_OVER_TWO_F = 1.570796326f;
	const float fov_y_angle_rad = 75.f * (PI_F / 180.f); /* TO_RADIANS(camera.fov_y_angle_deg) */
	const float aspect_ratio = (float)frame_width / frame_height;
	const float scale_y = (float)(1.0 / tan(fov_y_angle_rad / 2.0));
	const float scale_x = scale_y / aspect_ratio;
	const M4 clip_from_view = { { { scale_x, 0, 0, 0 }, { 0, scale_y, 0, 0 }, { 0, 0, 0, 0.01f }, { 0, 0, 1, 0 } } }; /* left-handed reversed-z infinite projection */
	if(yaw_rad > PI_F) yaw_rad -= TAU_F;
	else if(yaw_rad <= -PI_F) yaw_rad += TAU_F;
	pitch_rad = fminf(PI_OVER_TWO_F, pitch_rad);
	pitch_rad = fmaxf(-PI_OVER_TWO_F, pitch_rad);
	const float cos_pitch = (float)cos(-pitch_rad), sin_pitch = (float)sin(-pitch_rad);
	const float cos_yaw = (float)cos(-yaw_rad), sin_yaw = (float)sin(-yaw_rad);
	const M4 rotation_pitch = { { { 1, 0, 0, 0 }, { 0, cos_pitch, sin_pitch, 0 }, { 0, -sin_pitch, cos_pitch, 0 }, { 0, 0, 0, 1 } } };
	const M4 rotation_yaw = { { { cos_yaw, 0, -sin_yaw, 0 }, { 0, 1, 0, 0 }, { sin_yaw, 0, cos_yaw, 0 }, { 0, 0, 0, 1 } } };
	const M4 change_of_basis = { { { 0, 0, -1, 0 }, { 1, 0, 0, 0 }, { 0, 1, 0, 0 }, { 0, 0, 0, 1 } } };
	M4 world_from_view = m4_mul(&rotation_yaw, &rotation_pitch);
	world_from_view = m4_mul(&change_of_basis, &world_from_view);
	world_from_view.m[0][3] = px, world_from_view.m[1][3] = py, world_from_view.m[2][3] = pz;
	const M4 view_from_world = m4_inverse(&world_from_view);
	const M4 clip_from_world = m4_mul(&clip_from_view, &view_from_world);
	const M4 view_from_clip = m4_inverse(&clip_from_view);
	memcpy(per_frame_cb, &clip_from_world, 64); /* PerFrameCB main.c:169-173 */
	memcpy(per_frame_cb + 16, &view_from_clip, 64);
	memcpy(per_frame_cb + 32, &world_from_view, 64);
}

/* what replaces the GDI blit (StretchDIBits, main.c:286-357): the frame buffer as a binary PPM. A frame-buffer word is
 * 0x00RRGGBB... as the reference stores it for GDI: blue in the low byte (encode of the output merger, main.c:1176-1180) */
static int write_ppm(const char *path) {
	FILE *f = fopen(path, "wb");
	if(!f) return 1;
	fprintf(f, "P6\n%d %d\n255\n", frame_width, frame_height);
	for(size_t i = 0; i < (size_t)frame_width * frame_height; ++i) {
		const u32 c = frame_buffer[i];
		const unsigned char rgb[3] = { (unsigned char)(c >> 16), (unsigned char)(c >> 8), (unsigned char)c };
		fwrite(rgb, 1, 3, f);
	}
	fclose(f);
	return 0;
}

static uint64_t fnv64(const void *p, size_t n) {
	uint64_t h = 0xcbf29ce484222325ull;
	for(size_t i = 0; i < n; ++i) h = (h ^ ((const unsigned char *)p)[i]) * 0x100000001b3ull;
	return h;
}

/* every asset of a directory through the C reader: name, extents and a hash of the payload (CPU-only: tests compare
 * them with the Python reader's) */
static int list_assets(const char *dir, int argc, char **argv) {
	snprintf(assets_dir, sizeof(assets_dir), "%s", dir);
	for(int i = 0; i < argc; ++i) {
		char path[1200];
		asset_path(path, sizeof(path), argv[i]);
		OctrnMeshHeader mh;
		OctrnImageHeader ih;
		void *data = NULL;
		if(octrn_read_mesh(path, &mh, &data) == 0) {
			printf("mesh %s %u %u %016llx\n", argv[i], mh.vertex_count, mh.index_count, (unsigned long long)fnv64(data, mh.size));
		} else if(octrn_read_image(path, &ih, &data) == 0) {
			printf("image %s %u %u %u %016llx\n", argv[i], ih.width, ih.height, ih.format, (unsigned long long)fnv64(data, (size_t)ih.size));
		} else {
			fprintf(stderr, "render_host: %s\n", octrn_last_error());
			return 2;
		}
		free(data);
	}
	return 0;
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
	if(argc >= 3 && !strcmp(argv[1], "--list-assets")) return list_assets(argv[2], argc - 3, argv + 3);
	const char *scene_name = NULL, *cb_path = NULL, *ppm_path = NULL;
	float pose[5] = { 3.5f, 1.0f, 1.0f, 0.0f, 0.0f }; /* init()'s camera (main.c:1423-1425) */
	int a = 1;
	for(; a + 1 < argc && !strncmp(argv[a], "--", 2); a += 2) {
		if(!strcmp(argv[a], "--assets")) snprintf(assets_dir, sizeof(assets_dir), "%s", argv[a + 1]);
		else if(!strcmp(argv[a], "--scene")) scene_name = argv[a + 1];
		else if(!strcmp(argv[a], "--size")) sscanf(argv[a + 1], "%dx%d", &frame_width, &frame_height);
		else if(!strcmp(argv[a], "--pose")) sscanf(argv[a + 1], "%f,%f,%f,%f,%f", pose, pose + 1, pose + 2, pose + 3, pose + 4);
		else if(!strcmp(argv[a], "--cb")) cb_path = argv[a + 1];
		else if(!strcmp(argv[a], "--ppm")) ppm_path = argv[a + 1];
		else {
			fprintf(stderr, "render_host: unknown option %s\n", argv[a]);
			return 2;
		}
	}
	if(scene_name ? argc - a < 1 : argc - a < 2) {
		fprintf(stderr, "usage: render_host <scene.bin> <out.bin> [frames]\n"
		                "       render_host --assets <dir> --scene toon|ftm|suprematism --size <W>x<H> [--pose x,y,z,yaw,pitch] [--cb file] [--ppm frame.ppm] <out.bin> [frames]\n"
		                "       render_host --list-assets <dir> <name> ...\n");
		return 2;
	}
	const char *out_path = scene_name ? argv[a] : argv[a + 1];
	const int frames = argc > a + (scene_name ? 1 : 2) ? atoi(argv[a + (scene_name ? 1 : 2)]) : 1;
	FILE *f = NULL;
	if(scene_name) { /* the reference's own start-up: init() + update() */
		if(frame_width <= 0 || frame_height <= 0 || (frame_width & 7) || (frame_height & 7)) {
			fprintf(stderr, "render_host: --size must be given in multiples of the 8-pixel tile\n");
			return 2;
		}
		if(init_scene(scene_name)) {
			fprintf(stderr, "render_host: unknown scene %s\n", scene_name);
			return 2;
		}
		update_camera(pose[0], pose[1], pose[2], pose[3], pose[4]);
		if(cb_path) {
			FILE *c = fopen(cb_path, "rb");
			if(!c || fread(per_frame_cb, 4, 48, c) != 48) {
				fprintf(stderr, "render_host: cannot read 192 bytes from %s\n", cb_path);
				return 2;
			}
			fclose(c);
		}
		goto scene_ready;
	}
	f = fopen(argv[a], "rb");
	if(!f) {
		perror(argv[a]);
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

scene_ready:
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

	if(ppm_path && write_ppm(ppm_path)) {
		perror(ppm_path);
		return 2;
	}
	f = fopen(out_path, "wb");
	if(!f) {
		perror(out_path);
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
