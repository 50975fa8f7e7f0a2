/* oracle/harness.c -- TEST INFRASTRUCTURE ONLY (never on the product path).
 *
 * Appended by oracle/build_ref.sh to the pipeline section extracted from the reference's
 * main.c, inside the SAME translation unit, so it can drive the reference's own globals
 * and stage functions.  It contains no rendering arithmetic of its own: every pixel,
 * triangle and bin it returns was produced by the reference's code
 * (clear_render_target_view main.c:1191, clear_depth_stencil_view :1204, draw_indexed :1219,
 * stage functions :662-1189, camera :1422-1477, update :1480-1562).
 *
 * Exposed as a C API (ref_*) and loaded with ctypes by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs only.
 */

#define REF_API __attribute__((visibility("default")))

/* shader ids shared with include/malevich_b200.h (MLV_VS_* / MLV_PS_*) */
enum { REF_VS_PASSTHROUGH = 0, REF_VS_BASIC = 1, REF_VS_VERTEX_LIGHTING = 2, REF_VS_FULLSCREEN = 3 };
enum { REF_PS_PASSTHROUGH = 0, REF_PS_BASIC = 1, REF_PS_ENV_LIGHTING = 2 };

static Texture2D ref_bound_texture;

/* intermediates of the last staged draw (owned here, freed on the next staged draw) */
static void *ref_last_vs_in = NULL, *ref_last_vs_out = NULL;
static Triangle *ref_last_tris = NULL;
static v4f32 *ref_last_attrs = NULL;
static u32 *ref_last_ids = NULL;
static CompactedBin *ref_last_cbins = NULL;
static TileInfo *ref_last_infos = NULL;
static u32 ref_last_tri_count = 0, ref_last_pair_count = 0, ref_last_cbin_count = 0, ref_last_index_count = 0;

REF_API int ref_width(void) { return WIDTH; }
REF_API int ref_height(void) { return HEIGHT; }
REF_API void ref_set_threads(int n) { omp_set_num_threads(n); }
REF_API int ref_max_threads(void) { return omp_get_max_threads(); }

/* Runs the reference's camera init (main.c:1422-1477) and update() (main.c:1480-1562) for a given pose
 * and returns the resulting PerFrameCB (3 row-major 4x4 matrices, 192 bytes). */
REF_API const void *ref_camera(float px, float py, float pz, float yaw_rad, float pitch_rad) {
	memset(&input, 0, sizeof(input));
	oracle_init_camera();
	camera.pos = (v3f32){ px, py, pz };
	camera.yaw_rad = yaw_rad;
	camera.pitch_rad = pitch_rad;
	update(16.f);
	return &per_frame_cb;
}

REF_API void ref_set_cb0(const void *p_cb) { memcpy(&per_frame_cb, p_cb, sizeof(per_frame_cb)); }

/* Same state setup as render() main.c:1268-1281, minus the clears (issued separately). */
REF_API void ref_begin_frame(void) {
	memset(&stats, 0, sizeof(Stats));
	graphics_pipeline.ia.primitive_topology = PRIMITIVE_TOPOLOGY_TRIANGLELIST;
	Viewport viewport = { 0.f, 0.f, (f32)frame_width, (f32)frame_height, 0.f, 1.f };
	graphics_pipeline.rs.viewport = viewport;
	graphics_pipeline.om.p_colors = &frame_buffer[0][0];
	graphics_pipeline.om.p_depth = &depth_buffer[0][0];
	graphics_pipeline.vs.p_constant_buffers[0] = &per_frame_cb;
}

REF_API void ref_clear(const float *p_rgba, float depth) {
	clear_render_target_view(p_rgba);
	clear_depth_stencil_view(depth);
}

static void ref_bind_draw(const void *p_vb, const u32 *p_ib, int vs_id, int ps_id, const void *p_tex, u32 tex_w, u32 tex_h) {
	VertexShader vs;
	PixelShader ps;
	switch(vs_id) {
		case REF_VS_PASSTHROUGH: vs = passthrough_vs; break;
		case REF_VS_BASIC: vs = basic_vs; break;
		case REF_VS_VERTEX_LIGHTING: vs = vertex_lighting_vs; break;
		default: vs = fullscreen_vs; break;
	}
	switch(ps_id) {
		case REF_PS_PASSTHROUGH: ps = passthrough_ps; break;
		case REF_PS_BASIC: ps = basic_ps; break;
		default: ps = env_lighting_ps; break;
	}
	/* mirrors render() main.c:1286-1294 */
	graphics_pipeline.ia.input_layout = vs.in_vertex_size / VECTOR_WIDTH;
	graphics_pipeline.vs.output_register_count = vs.out_vertex_size / (sizeof(v4f32) * VECTOR_WIDTH);
	graphics_pipeline.vs.shader = (void *)vs.vs_main;
	graphics_pipeline.ps.shader = (void *)ps.ps_main;
	graphics_pipeline.ia.p_index_buffer = (u32 *)p_ib;
	graphics_pipeline.ia.p_vertex_buffer = (void *)p_vb;
	ref_bound_texture.p_data = (void *)p_tex;
	ref_bound_texture.width = tex_w;
	ref_bound_texture.height = tex_h;
	graphics_pipeline.vs.p_shader_resource_views[0] = &ref_bound_texture;
	graphics_pipeline.ps.p_shader_resource_views[0] = &ref_bound_texture;
}

REF_API void ref_draw(const void *p_vb, const u32 *p_ib, u32 index_count, int vs_id, int ps_id, const void *p_tex, u32 tex_w, u32 tex_h) {
	ref_bind_draw(p_vb, p_ib, vs_id, ps_id, p_tex, tex_w, tex_h);
	draw_indexed(index_count);
}

static void ref_free_staged(void) {
	free(ref_last_vs_in); free(ref_last_vs_out); free(ref_last_attrs); free(ref_last_tris);
	free(ref_last_ids); free(ref_last_infos); free(ref_last_cbins);
	ref_last_vs_in = ref_last_vs_out = NULL; ref_last_attrs = NULL; ref_last_tris = NULL;
	ref_last_ids = NULL; ref_last_infos = NULL; ref_last_cbins = NULL;
}

/* Calls the reference's stage functions in draw_indexed order (main.c:1222-1251) but keeps the
 * intermediates alive so tests can compare them one by one. */
REF_API void ref_draw_staged(const void *p_vb, const u32 *p_ib, u32 index_count, int vs_id, int ps_id, const void *p_tex, u32 tex_w, u32 tex_h) {
	ref_bind_draw(p_vb, p_ib, vs_id, ps_id, p_tex, tex_w, tex_h);
	ref_free_staged();
	ref_last_index_count = index_count;

	input_assembler_stage(index_count, &ref_last_vs_in);
	u32 per_vertex_output_data_size = 0;
	vertex_shader_stage(index_count, ref_last_vs_in, &per_vertex_output_data_size, &ref_last_vs_out);
	stats.vertex_count += index_count;
	u32 triangle_count = index_count / 3;
	stats.input_triangle_count += triangle_count;
	primitive_assembly_stage(triangle_count, ref_last_vs_out, &ref_last_tri_count, &ref_last_tris, &ref_last_attrs);
	stats.assembled_triangle_count += ref_last_tri_count;
	binner(ref_last_tri_count, ref_last_tris, &ref_last_ids, &ref_last_cbins, &ref_last_cbin_count, &ref_last_pair_count);
	stats.active_bin_count += ref_last_cbin_count;
	stats.total_triangle_count_in_bins += ref_last_pair_count;
	rasterizer(ref_last_pair_count, ref_last_cbin_count, ref_last_tris, ref_last_ids, ref_last_cbins, &ref_last_infos);
	pixel_shader_stage(ref_last_infos, ref_last_tris, ref_last_cbins, ref_last_cbin_count);
}

REF_API const void *ref_staged_vs_out(u32 *p_vertex_count) { *p_vertex_count = ref_last_index_count; return ref_last_vs_out; }
REF_API const void *ref_staged_triangles(u32 *p_count) { *p_count = ref_last_tri_count; return ref_last_tris; }
REF_API const void *ref_staged_attributes(void) { return ref_last_attrs; }
REF_API const void *ref_staged_triangle_ids(u32 *p_count) { *p_count = ref_last_pair_count; return ref_last_ids; }
REF_API const void *ref_staged_compacted_bins(u32 *p_count) { *p_count = ref_last_cbin_count; return ref_last_cbins; }
REF_API const void *ref_staged_tile_infos(void) { return ref_last_infos; }
REF_API unsigned ref_sizeof_triangle(void) { return sizeof(Triangle); }
REF_API unsigned ref_sizeof_tile_info(void) { return sizeof(TileInfo); }

REF_API const u32 *ref_colors(void) { return &frame_buffer[0][0]; }
REF_API const f32 *ref_depths(void) { return &depth_buffer[0][0]; }
REF_API const f32 *ref_tile_min_depths(void) { return a_tile_min_depths; }
REF_API const void *ref_stats(void) { return &stats; }

/* Embedded scenes of the reference (main.c:232-270): SUPREMATISM and the fullscreen quad. */
REF_API const void *ref_suprematist_vb(u32 *p_bytes) { *p_bytes = sizeof(suprematist_vertex_buffer); return suprematist_vertex_buffer; }
REF_API const void *ref_suprematist_ib(u32 *p_count) { *p_count = sizeof(suprematist_index_buffer) / 4; return suprematist_index_buffer; }
REF_API const void *ref_fullscreen_vb(u32 *p_bytes) { *p_bytes = sizeof(fullscreen_vertex_buffer); return fullscreen_vertex_buffer; }
REF_API const void *ref_fullscreen_ib(u32 *p_count) { *p_count = sizeof(fullscreen_index_buffer) / 4; return fullscreen_index_buffer; }

/* stage timer (see prefix.h) */
REF_API void ref_prof_reset(void) { for(int i = 0; i < ref_prof_count; ++i) ref_prof_ms[i] = 0; }
REF_API int ref_prof_get(int i, const char **pp_name, double *p_ms) {
	if(i >= ref_prof_count) return 0;
	*pp_name = ref_prof_names[i]; *p_ms = ref_prof_ms[i];
	return 1;
}

/* The reference's vrsqrtps (math.h:277-280) as executed by THIS host, for N6. */
REF_API float ref_rsqrt(float x) { return _mm_cvtss_f32(_mm_rsqrt_ss(_mm_set_ss(x))); }
