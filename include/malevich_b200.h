/* include/malevich_b200.h -- C-ABI of the B200-native Malevich draw pipeline.
 *
 * This is the drop-in boundary: plain C, pointers and sizes only, every entry
 * point returns an int status (MLV_OK == 0) and mlv_last_error_string() explains a failure.
 * Each function cites the reference interface it replaces (paths relative to the reference
 * checkout, `source/main.c` unless stated).
 *
 * Threading: one host thread per mlv_device. Calls are stream-ordered and asynchronous until
 * mlv_present_readback / mlv_finish / mlv_get_stats / any mlv_debug_read_* (these synchronise).
 * There is NO CPU fallback: every entry point that touches pixels fails with MLV_ERR_CUDA when no
 * sm_100 device is usable.
 */
#ifndef MALEVICH_B200_H
#define MALEVICH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MLV_API __attribute__((visibility("default")))
#else
#define MLV_API
#endif

/* ---- status codes -------------------------------------------------------------------------- */
enum {
	MLV_OK = 0,
	MLV_ERR_INVALID_ARGUMENT = 1, /* reference: assert() main.c:666,670,1230 */
	MLV_ERR_CUDA = 2,             /* CUDA runtime failure / no device (reference: none, CPU only) */
	MLV_ERR_OUT_OF_MEMORY = 3,
	MLV_ERR_CAPACITY = 4,         /* a per-draw arena overflowed (reference mallocs per draw, main.c:745-746,947,986) */
	MLV_ERR_STATE = 5             /* draw issued with incomplete pipeline state */
};

/* ---- enumerations mirroring the reference's state ------------------------------------------ */
/* PrimitiveTopology, main.c:66-69 */
enum { MLV_PRIMITIVE_TOPOLOGY_UNDEFINED = 0, MLV_PRIMITIVE_TOPOLOGY_TRIANGLELIST = 1 };
/* The reference binds shaders as host function pointers (VS.shader main.c:79, PS.shader main.c:99,
 * descriptors main.c:46-52). On the device they are __device__ functions selected by id. */
enum { MLV_VS_PASSTHROUGH = 0,     /* passthrough_vs.c:16-25 */
       MLV_VS_BASIC = 1,           /* basic_vs.c:22-33 */
       MLV_VS_VERTEX_LIGHTING = 2, /* vertex_lighting_vs.c:22-39 */
       MLV_VS_FULLSCREEN = 3,      /* fullscreen_vs.c:22-39 */
       MLV_VS_COUNT = 4 };
enum { MLV_PS_PASSTHROUGH = 0,  /* passthrough_ps.c:13-20 */
       MLV_PS_BASIC = 1,        /* basic_ps.c:16-27 */
       MLV_PS_ENV_LIGHTING = 2, /* env_lighting_ps.c:13-24 */
       MLV_PS_BASIC_TRILINEAR = 3, /* EXTENSION: basic_ps.c:16-27 with SRV0 sampled trilinearly from the mip chain
                                    * mlv_texture_generate_mips built; the reference samples level 0 only (common_shader_core.h:195-199) */
       MLV_PS_COUNT = 4 };
/* Texture2D, common_shader_core.h:20-24: the reference infers the texel type from the sampler used
 * (get_texel_u_x8 :30 vs get_texel_f_x8 :42); here it is explicit. */
enum { MLV_FORMAT_R8G8B8A8_UNORM = 0, MLV_FORMAT_R32G32B32A32_FLOAT = 1 };
enum { MLV_BUFFER_VERTEX = 0, MLV_BUFFER_INDEX = 1 };
enum { MLV_INDEX_U32 = 0, MLV_INDEX_U16 = 1 }; /* the reference has u32 only ("TODO: 16-bit index buffers", main.c:72) */

#define MLV_CONSTANT_BUFFER_SLOT_COUNT 16 /* COMMONSHADER_CONSTANT_BUFFER_HW_SLOT_COUNT main.c:41 */
#define MLV_SHADER_RESOURCE_SLOT_COUNT 16 /* COMMONSHADER_INPUT_RESOURCE_REGISTER_COUNT main.c:42 */
#define MLV_TILE_SIZE 8                   /* TILE_WIDTH/TILE_HEIGHT main.c:24-25 */

/* ---- plain-data structs --------------------------------------------------------------------- */
/* Viewport, main.c:85-92 (same field order). The viewport must equal the render-target size, as in
 * the reference (tile pitch uses viewport.width, bins use WIDTH_IN_TILES: main.c:582,590). */
typedef struct mlv_viewport {
	float top_left_x, top_left_y, width, height, min_depth, max_depth;
} mlv_viewport;

/* Stats, main.c:199-206 (same field order, same accumulation points main.c:1228-1246). */
typedef struct mlv_stats {
	float frame_time;
	uint32_t vertex_count;
	uint32_t input_triangle_count;
	uint32_t assembled_triangle_count;
	uint32_t active_bin_count;
	uint32_t total_triangle_count_in_bins;
} mlv_stats;

/* What the device really processed since the last mlv_reset_stats -- unlike Stats, which count every assembled triangle and
 * every (triangle, tile) pair like the reference does, these exclude what Hi-Z at binning time removed before it cost memory
 * traffic. bench.py derives the per-kernel algorithmic bytes (SURVEY.md 8d per-unit figures) from them. */
typedef struct mlv_work_counters {
	uint64_t records_written; /* assembled triangles that survived Hi-Z in at least one tile: setup records written, later read by k_tile */
	uint64_t pairs_listed;    /* (triangle, tile) pairs stored in the per-tile lists */
	uint64_t tiles_visited;   /* tile-draws k_tile loaded and stored (work-list bins) */
} mlv_work_counters;

/* Replaces the compile-time WIDTH/HEIGHT (main.c:21-22) and adds the sort-first partition
 * (SURVEY.md 8e): rank r of num_ranks owns the 8-pixel tile rows ty with (ty / stripe_height_tiles)
 * % num_ranks == r. width and height must be multiples of 8 (WIDTH_IN_TILES main.c:28-29). */
typedef struct mlv_device_desc {
	uint32_t width, height;
	int32_t cuda_device;          /* ordinal; -1 = current device */
	uint32_t num_ranks;           /* 0 or 1 = single GPU */
	uint32_t rank;
	uint32_t stripe_height_tiles; /* 0 = default (1) */
	uint64_t max_pairs_per_draw;  /* capacity of the (triangle,tile) list arena; 0 = default */
	uint32_t flags;               /* MLV_DEVICE_* */
	uint32_t num_gpus;            /* 0 or 1 = this device only. N > 1 = a device GROUP: one host thread, CUDA devices cuda_device ..
	                               * cuda_device + N - 1 in this process, one sort-first rank each (num_ranks / rank must be 0) */
} mlv_device_desc;
enum { MLV_DEVICE_DEBUG_CAPTURE = 1, /* keep reference-layout intermediates of the last draw for mlv_debug_read_* */
       MLV_DEVICE_GROUP_SAME_GPU = 2, /* device group: every rank on cuda_device (tests on a single-GPU box) */
       MLV_DEVICE_GROUP_NCCL = 4,     /* device group: compose the frame on the device with pack + in-place ncclAllGather + unpack
                                       * (libnccl.so.2, loaded on demand), then read rank 0's image back */
       MLV_DEVICE_GROUP_PEER_EXCHANGE = 8 }; /* device group: compose it on the device with the asynchronous peer-memory exchange over
                                       * NVLink, then read rank 0's image back. Default (neither flag): the frame is composed IN HOST
                                       * MEMORY -- every GPU copies the rows it owns over its own PCIe link (mlv_present_owned_rows_async) */
/* DEVICE GROUPS. With num_gpus = N the multi-GPU fan-out lives behind the same calls: the handle returned by
 * mlv_create_device stands for N per-GPU devices -- geometry, textures and pipeline state replicated, rank i rasterising the
 * tile rows (ty / stripe_height_tiles) % N == i (default: one contiguous band per GPU) -- and the entry points a renderer
 * needs fan out to them: buffers and textures (create / update / release, sRGB, mips), every state setter, both clears, the
 * three draws, command lists (one recording per GPU, one graph launch per GPU per frame), mlv_present_readback[_async] /
 * mlv_present_wait (every GPU delivers the rows it owns straight to the host frame; or, by flag, exchange of the stripes over
 * peer memory / NCCL and then rank 0's image; depth rows are collected from their owners on the host), mlv_get_stats / mlv_get_work_counters (sums of the ranks' shares = the reference's Stats),
 * mlv_reset_stats, mlv_finish, mlv_kernel_launch_count, mlv_destroy_device. Every other entry point returns
 * MLV_ERR_STATE for a group. host/malevich_compat.c creates a group when MLV_NUM_GPUS is set, so the reference's
 * render() runs on N GPUs unchanged. */

typedef struct mlv_device mlv_device;
typedef struct mlv_buffer mlv_buffer;
typedef struct mlv_texture mlv_texture;

/* Debug read-back records, byte-compatible with the reference's x86-64 layouts. */
typedef struct mlv_ref_triangle { /* Triangle main.c:134-139 + Setup :127-132 + EdgeFunction :121-125, 80 bytes */
	uint64_t p_attributes;        /* host pointer in the reference; here: triangle index * 144 */
	int32_t min_bounds[2], max_bounds[2];
	int32_t edges[3][3];          /* a_edge_functions[k] = {a,b,c} */
	float reciprocal_ws[3];
	float one_over_area;
	float max_depth;
} mlv_ref_triangle;
typedef struct mlv_ref_compacted_bin { uint32_t num_triangles_self, num_triangles_upto, bin_index; } mlv_ref_compacted_bin; /* main.c:146-150 */
typedef struct mlv_ref_tile_info { uint32_t triangle_id; uint32_t _pad; uint64_t fragment_mask; } mlv_ref_tile_info;         /* main.c:159-162 */

/* ---- device ---------------------------------------------------------------------------------- */
MLV_API const char *mlv_last_error_string(void);                                    /* error() main.c:451-465 */
MLV_API int mlv_create_device(const mlv_device_desc *desc, mlv_device **out_device); /* static frame_buffer/depth_buffer/a_bins/a_tile_min_depths main.c:35-36,229-230 */
MLV_API void mlv_destroy_device(mlv_device *dev);
MLV_API int mlv_finish(mlv_device *dev);
MLV_API void *mlv_get_stream(mlv_device *dev); /* the cudaStream_t all work is launched on (for CUDA-event timing) */

/* ---- resources (load_mesh main.c:526-536, load_texture main.c:538-559 hand the pipeline host pointers) */
MLV_API int mlv_create_buffer(mlv_device *dev, const void *data, size_t bytes, int kind, mlv_buffer **out);
/* Uploads are asynchronous and run on the device's copy stream: the copy starts after every draw issued so far and
 * each later draw waits only for the buffers it binds, so uploading mesh k+1 overlaps drawing mesh k. `data` must stay
 * valid until mlv_finish / mlv_present_readback returns when it points to page-locked memory. */
MLV_API int mlv_update_buffer(mlv_device *dev, mlv_buffer *buf, const void *data, size_t bytes);
MLV_API void mlv_release_buffer(mlv_device *dev, mlv_buffer *buf);
/* Sharded uploads for replicated geometry (sort-first): every rank holds the same host buffers, so rank r
 * uploads only bytes [offset, offset + bytes) over its own PCIe link (mlv_update_buffer_range), the ranks exchange the
 * shards in place over NVLink (the caller's collective, e.g. ncclAllGather on mlv_buffer_device_ptr, ordered after the
 * upload on mlv_get_copy_stream) and mlv_buffer_mark_updated tells the library on which stream the buffer becomes
 * complete: later draws that bind it wait for that point. Create the buffer with mlv_create_buffer(dev, NULL, capacity,
 * kind) when the collective wants a padded size. */
MLV_API int mlv_update_buffer_range(mlv_device *dev, mlv_buffer *buf, size_t offset, const void *data, size_t bytes);
MLV_API void *mlv_buffer_device_ptr(mlv_buffer *buf);
MLV_API void *mlv_get_copy_stream(mlv_device *dev);
MLV_API int mlv_buffer_mark_updated(mlv_device *dev, mlv_buffer *buf, void *cuda_stream /* NULL = the copy stream */);
MLV_API int mlv_create_texture2d(mlv_device *dev, const void *texels, uint32_t width, uint32_t height, int format, mlv_texture **out);
MLV_API void mlv_release_texture(mlv_device *dev, mlv_texture *tex);
/* load_texture's `is_in_srgb` branch (main.c:546-558): every channel of every texel of an R8G8B8A8 texture -- alpha
 * included -- goes through decode_u32_as_color (math.h:326-334), srgb_to_linear (math.h:386-395, double arithmetic)
 * and encode_color_as_u32 (math.h:322-324, truncation), in place, on the device. */
MLV_API int mlv_texture_srgb_to_linear(mlv_device *dev, mlv_texture *tex);
/* EXTENSION (the reference has no mips): builds the full mip chain of an R8G8B8A8 texture on the device, level
 * l = max(1, extent >> l), each level a 2x2 box filter of the one above per channel ((a+b+c+d+2)>>2; an odd extent repeats its
 * last row / column). Call it after mlv_texture_srgb_to_linear so that the filter works on linear values. Without a chain
 * MLV_PS_BASIC_TRILINEAR samples level 0 only and equals MLV_PS_BASIC. */
MLV_API int mlv_texture_generate_mips(mlv_device *dev, mlv_texture *tex);
MLV_API int mlv_texture_mip_levels(const mlv_texture *tex, uint32_t *out_levels); /* 1 = level 0 only */
MLV_API int mlv_read_texture_mip(mlv_device *dev, const mlv_texture *tex, uint32_t level, void *out_texels); /* synchronises */
/* copies the texels back (R8G8B8A8: 4 bytes, R32G32B32A32_FLOAT: 16 bytes per texel); synchronises */
MLV_API int mlv_read_texture(mlv_device *dev, const mlv_texture *tex, void *out_texels);

/* ---- pipeline state: graphics_pipeline.{ia,vs,rs,ps} main.c:71-115, written by render() main.c:1276-1294 */
MLV_API int mlv_ia_set_vertex_buffer(mlv_device *dev, mlv_buffer *vb);       /* ia.p_vertex_buffer main.c:1292 */
MLV_API int mlv_ia_set_index_buffer(mlv_device *dev, mlv_buffer *ib);        /* ia.p_index_buffer main.c:1291 (u32 indices) */
MLV_API int mlv_ia_set_index_format(mlv_device *dev, int format);            /* MLV_INDEX_U32 (default, what the reference has) or MLV_INDEX_U16 (its TODO at main.c:72) */
MLV_API int mlv_ia_set_input_layout(mlv_device *dev, uint32_t bytes_per_vertex); /* ia.input_layout main.c:1286 (must be 32) */
MLV_API int mlv_ia_set_primitive_topology(mlv_device *dev, int topology);    /* ia.primitive_topology main.c:1276 */
MLV_API int mlv_vs_set_shader(mlv_device *dev, int vs_id);                   /* vs.shader + vs.output_register_count main.c:1287-1288 */
MLV_API int mlv_vs_set_constant_buffer(mlv_device *dev, uint32_t slot, const void *data, size_t bytes); /* vs.p_constant_buffers[slot] main.c:1281 (PerFrameCB main.c:169-173, 192 bytes) */
MLV_API int mlv_vs_set_shader_resource(mlv_device *dev, uint32_t slot, mlv_texture *tex); /* vs.p_shader_resource_views[slot] main.c:1293 */
MLV_API int mlv_rs_set_viewport(mlv_device *dev, const mlv_viewport *vp);    /* rs.viewport main.c:1277-1278 */
MLV_API int mlv_ps_set_shader(mlv_device *dev, int ps_id);                   /* ps.shader main.c:1289 */
MLV_API int mlv_ps_set_shader_resource(mlv_device *dev, uint32_t slot, mlv_texture *tex); /* ps.p_shader_resource_views[slot] main.c:1294 */

/* ---- the hot path ---------------------------------------------------------------------------- */
MLV_API int mlv_clear_render_target_view(mlv_device *dev, const float rgba[4]); /* clear_render_target_view main.c:1191-1202 */
MLV_API int mlv_clear_depth_stencil_view(mlv_device *dev, float depth);         /* clear_depth_stencil_view main.c:1204-1217 */
MLV_API int mlv_draw_indexed(mlv_device *dev, uint32_t index_count);            /* draw_indexed main.c:1219-1261 */
MLV_API int mlv_draw_indexed_ex(mlv_device *dev, uint32_t index_count, uint32_t start_index_location, int32_t base_vertex_location); /* ID3D11DeviceContext::DrawIndexed in full: the reference's TODO at main.c:1219 */
MLV_API int mlv_draw(mlv_device *dev, uint32_t vertex_count);                   /* draw_indexed with identity indices (all shipped meshes) */

/* ---- command lists: the frame recorded once, replayed with one launch ------------------------------------------------
 * D3D11's deferred-context pattern (ID3D11DeviceContext::FinishCommandList / ExecuteCommandList), which the reference --
 * an immediate-mode CPU renderer whose render() (main.c:1265-1299) re-issues every call every frame -- has no analogue of.
 * Between mlv_begin_command_list and mlv_finish_command_list the state setters, clears, draws, mlv_resolve,
 * mlv_reset_stats and mlv_composite_pack are RECORDED (captured into a CUDA graph) instead of executed; everything that
 * synchronises, reads back or exchanges with peers fails with MLV_ERR_STATE. mlv_execute_command_list replays the recording:
 * one launch instead of ~7 per draw. The recording binds buffer and texture OBJECTS, not their contents: mlv_update_buffer
 * between executions is honoured (the execution waits for it). The vertex-shader constant buffer is recorded by value;
 * mlv_command_list_set_constants replaces it (for one draw or MLV_ALL_DRAWS) without re-recording -- what update()
 * (main.c:1480-1562) changes per frame. Buffers, textures and the list must outlive its executions.
 * The device keeps a PAIR of tiled framebuffers. In immediate mode a frame that opens with a full clear (colour + depth) is
 * drawn into the one nothing still reads -- the exchange of the previous frame, or its asynchronous present, which resolves /
 * packs and copies on a stream of its own -- so the next frame starts at once. A recorded list addresses ONE of the two; a
 * recording that opens with a full clear takes the one the previous such recording did not, so a host that records its frame
 * twice and replays the two lists in turn gets the same overlap. Either way the images are the same: nothing of the old
 * contents survives a full clear. A list that does NOT open with a full clear draws over what its framebuffer holds; when
 * those contents live in the other one (only possible after asynchronous presents or exchanges made a frame move over)
 * mlv_execute_command_list fails with MLV_ERR_STATE instead of drawing over the wrong image. */
typedef struct mlv_command_list mlv_command_list;
#define MLV_ALL_DRAWS 0xffffffffu
MLV_API int mlv_begin_command_list(mlv_device *dev);
MLV_API int mlv_finish_command_list(mlv_device *dev, mlv_command_list **out_list);
MLV_API int mlv_execute_command_list(mlv_device *dev, mlv_command_list *list);
MLV_API int mlv_command_list_set_constants(mlv_device *dev, mlv_command_list *list, uint32_t draw_index, const void *data, size_t bytes);
MLV_API int mlv_command_list_info(const mlv_command_list *list, uint32_t *out_draws, uint64_t *out_kernel_launches);
MLV_API void mlv_release_command_list(mlv_device *dev, mlv_command_list *list);

/* ---- results ---------------------------------------------------------------------------------- */
/* Replaces the GDI blit of frame_buffer (paint_window main.c:286-357): row-major y*W+x, colour
 * 0x00RRGGBB from the PS path, depth f32 reversed-Z. Either pointer may be NULL. Synchronises.
 * With num_ranks > 1 only the tiles this rank owns are meaningful unless the caller composited
 * the ranks first (mlv_composite_*). */
MLV_API int mlv_present_readback(mlv_device *dev, uint32_t *colors, float *depths);
/* The same without blocking the host OR the device stream: the resolve and the device-to-host copies run on a read-back
 * stream of their own (highest priority, ordered after everything issued so far), so frame f leaves while frame f+1 is
 * uploaded and rendered -- a frame that opens with a colour + depth clear is drawn into the other tiled framebuffer of the
 * device's pair and does not wait for the resolve that still reads the first; anything else that writes the framebuffer
 * waits for it. `colors` / `depths` should be page-locked (mlv_register_host_memory) and must stay untouched until
 * mlv_present_wait (or mlv_finish) returns; presents follow each other on the read-back stream, which protects the
 * resolved device images (they belong to the present until then: use mlv_resolve for a device-resident image). Capacity
 * errors of the frame are reported by the next synchronising call (mlv_get_stats, mlv_present_readback). */
MLV_API int mlv_present_readback_async(mlv_device *dev, uint32_t *colors, float *depths);
MLV_API int mlv_present_wait(mlv_device *dev);
MLV_API int mlv_get_stats(mlv_device *dev, mlv_stats *out);  /* stats main.c:231,1268 */
MLV_API int mlv_reset_stats(mlv_device *dev);                /* memset(&stats,0) main.c:1268 */
MLV_API int mlv_get_work_counters(mlv_device *dev, mlv_work_counters *out); /* synchronises; reset by mlv_reset_stats */

/* Device-resident resolve: tiled colour/depth -> row-major device buffers with 128-bit stores. */
MLV_API int mlv_resolve(mlv_device *dev);
MLV_API void *mlv_resolved_color_device_ptr(mlv_device *dev); /* W*H u32, valid after mlv_resolve / mlv_composite_unpack / mlv_composite_wait (call it after those: the buffer alternates) */
MLV_API void *mlv_resolved_depth_device_ptr(mlv_device *dev); /* W*H f32, valid after mlv_resolve */

/* Sort-first compositing. The gather buffer holds num_ranks equal chunks; rank r's
 * chunk is the row-major colour of its owned stripes in ascending stripe order (padded to the
 * largest rank's stripe count). mlv_composite_pack fills this rank's chunk; the caller runs
 * ncclAllGather in place over the whole buffer on mlv_get_stream(); mlv_composite_unpack scatters all
 * chunks into the resolved row-major colour image. */
MLV_API int mlv_composite_layout(mlv_device *dev, void **out_gather_device_ptr, size_t *out_chunk_bytes);
MLV_API int mlv_composite_pack(mlv_device *dev);
MLV_API int mlv_composite_unpack(mlv_device *dev);

/* Peer-memory compositing: the fused form of the exchange above for GPUs that can address each other (NVLink /
 * NVSwitch; cudaIpc between processes). mlv_composite_broadcast resolves this rank's tiles straight into the row-major
 * image of EVERY rank and publishes a per-frame arrival word on each; mlv_composite_wait blocks the device stream (not
 * the host) until the stripes of all ranks have arrived, after which mlv_resolved_color_device_ptr points at the complete
 * image. No staging chunk, no collective call, no un-swizzle pass. Setup, once per device:
 *   1. every rank: mlv_composite_peer_export(dev, &info)        (allocates two images + the arrival words)
 *   2. the caller exchanges the mlv_peer_info structs (any transport: ncclAllGather, MPI, a pipe)
 *   3. every rank: mlv_composite_peer_attach(dev, infos, same_process)
 * Every mlv_composite_broadcast must be followed by mlv_composite_wait before the next one (a rank may run at most one
 * frame ahead of the slowest; the images are double-buffered by frame parity). A peer that never arrives makes the
 * wait give up after 10 s and the next read-back / mlv_get_stats return MLV_ERR_STATE. */
#define MLV_MAX_PEERS 16
typedef struct mlv_peer_info {
	unsigned char ipc_color[2][64]; /* cudaIpcMemHandle_t of the two images */
	unsigned char ipc_flags[64];    /* cudaIpcMemHandle_t of the arrival words */
	void *color[2];                 /* the same allocations as plain device pointers (same_process != 0) */
	void *flags;
	int cuda_device;
	int reserved;
} mlv_peer_info;
MLV_API int mlv_composite_peer_export(mlv_device *dev, mlv_peer_info *out);
MLV_API int mlv_composite_peer_attach(mlv_device *dev, const mlv_peer_info *infos /* [num_ranks], in rank order */, int same_process);
MLV_API int mlv_composite_broadcast(mlv_device *dev);
MLV_API int mlv_composite_wait(mlv_device *dev);
/* The same exchange off the critical path. mlv_composite_broadcast_async puts the broadcast and the wait of frame f on
 * an exchange stream of their own (high priority, ordered after everything issued so far) and returns; the device
 * goes on with frame f+1 at once -- a frame that starts with a colour + depth clear is drawn into the second tiled
 * framebuffer of a pair, so it does not wait for the broadcast that still reads the first. mlv_composite_join makes the
 * device stream wait for that exchange; afterwards mlv_resolved_color_device_ptr is the complete image of frame f. Rules:
 * every _async is joined before the next broadcast of either kind, and whatever reads the joined image (work on
 * mlv_get_stream(), or mlv_present_*) is issued before the next broadcast -- peers reuse the image two frames later,
 * after this rank's next broadcast has run. The usual frame loop is
 *     draw frame f+1;  mlv_composite_join (frame f);  consume frame f;  mlv_composite_broadcast_async (frame f+1)
 * whose frame time is a rank's own rendering time: the NVLink transfer and the wait for the slowest rank overlap the next
 * frame's geometry. */
MLV_API int mlv_composite_broadcast_async(mlv_device *dev);
/* Device-to-host copy of the composited image on the read-back stream (like mlv_present_readback_async, which reads this
 * rank's own tiles); completed by mlv_present_wait / mlv_finish. Call it after mlv_composite_wait / mlv_composite_join. */
MLV_API int mlv_composite_readback_async(mlv_device *dev, uint32_t *colors);
MLV_API int mlv_composite_join(mlv_device *dev);
/* Composite in HOST memory: this rank packs the stripes it owns and copies them to their rows of frame_colors -- the FULL
 * row-major frame (width * height words) in pinned host memory shared by all ranks -- on the read-back stream, over its own
 * PCIe link; nothing is exchanged between the GPUs. The frame is complete when every rank's mlv_present_wait has returned.
 * A single-rank device reads its whole frame back; a device group fans the call out. */
MLV_API int mlv_present_owned_rows_async(mlv_device *dev, uint32_t *frame_colors);
/* Page-lock host memory the caller owns (malloc, or a MAP_SHARED mapping that several rank processes open: the shared frame
 * of mlv_present_owned_rows_async) for this device's CUDA context, so that the asynchronous read-backs into it run at full
 * PCIe rate and really are asynchronous; undo it before the memory is freed or unmapped. A host that does not link CUDA
 * has no other way to do this (the reference's frame buffer is a plain static array, main.c:35). */
MLV_API int mlv_register_host_memory(mlv_device *dev, void *ptr, size_t bytes);
MLV_API int mlv_unregister_host_memory(mlv_device *dev, void *ptr);

/* ---- debug read-back of the last draw (needs MLV_DEVICE_DEBUG_CAPTURE). Each synchronises.
 * Pass NULL data pointers to query the counts only. */
MLV_API int mlv_debug_read_vs_out(mlv_device *dev, float *out12_per_vertex, uint32_t *out_vertex_count);        /* vertex_shader_stage output main.c:714-727, 48 B/vertex */
MLV_API int mlv_debug_read_triangles(mlv_device *dev, mlv_ref_triangle *tris, float *attributes36, uint32_t *out_count); /* primitive_assembly_stage outputs main.c:877-906 */
MLV_API int mlv_debug_read_bins(mlv_device *dev, uint32_t *triangle_ids, uint32_t *out_pair_count, mlv_ref_compacted_bin *bins, uint32_t *out_bin_count); /* binner outputs main.c:947-978 */
MLV_API int mlv_debug_read_masks(mlv_device *dev, mlv_ref_tile_info *infos, uint32_t *out_pair_count);          /* rasterizer output main.c:986-1041 */
MLV_API int mlv_debug_read_tile_min_depths(mlv_device *dev, float *out_bins);                                    /* a_tile_min_depths main.c:230 (works without debug capture) */
/* keys[i] = the device's order-preserving key of the reference's assembled triangle id i of the last draw (ascending) */
MLV_API int mlv_debug_read_keys(mlv_device *dev, uint32_t *keys, uint32_t *out_count);
/* The per-tile lists of the last draw as the PRODUCTION path built them (no debug capture needed): device keys in arrival
 * order with the pairs Hi-Z rejects at binning time (main.c:1003-1010) already removed, and the work list of bins
 * (num_triangles_self may be 0: a touched bin visited only for write_tile's tile-minimum refresh, main.c:589-603).
 * Sorting each list and adding the rejected pairs back gives the reference's triangle_ids (main.c:950-962). Synchronises. */
MLV_API int mlv_read_bin_lists(mlv_device *dev, uint32_t *keys, uint32_t *out_pair_count, mlv_ref_compacted_bin *bins, uint32_t *out_bin_count);

/* Per-stage device timing (replaces the reference's Remotery scopes, rmt_BeginCPUSample main.c:663,699,737,916,
 * 984,1047,1192,1205): between mlv_profile_begin and mlv_profile_end every kernel launch is bracketed by CUDA events
 * on the device's stream. mlv_profile_end synchronises and returns, per stage, the summed kernel time in
 * milliseconds and the number of launches. Arrays have MLV_STAGE_COUNT entries. */
enum { MLV_STAGE_CLEAR = 0,      /* k_clear */
       MLV_STAGE_GEOMETRY = 1,   /* k_front: the front half of geometry (+ k_chunk_bounds when the sort-first chunk bounds are (re)built) */
       MLV_STAGE_BIN_COUNT = 2,  /* k_bin_big */
       MLV_STAGE_BIN_SCAN = 3,   /* k_bin_scan */
       MLV_STAGE_BIN_FILL = 4,   /* k_bin_fill */
       MLV_STAGE_TILE = 5,       /* k_tile */
       MLV_STAGE_RESOLVE = 6,    /* k_resolve */
       MLV_STAGE_COMPOSITE = 7,  /* k_composite_pack / k_composite_unpack */
       MLV_STAGE_VERTEX = 8,     /* k_vertex (post-transform vertex cache) */
       MLV_STAGE_CLIP = 9,       /* k_front_clip */
       MLV_STAGE_BACK = 10,      /* k_back: Hi-Z + binner pass 1 + setup records of the survivors */
       MLV_STAGE_COUNT = 11 };
MLV_API int mlv_profile_begin(mlv_device *dev);
MLV_API int mlv_profile_end(mlv_device *dev, double *out_ms, uint32_t *out_launches);
/* The same region launch by launch -- the timeline Remotery's viewer draws (external/Remotery/vis): stage, start relative
 * to the first launch of the region, duration. Valid after mlv_profile_end; out == NULL queries the count.
 * tools/trace_frame.py turns it into a Chrome/Perfetto trace with the reference's scope names. */
typedef struct mlv_profile_event { int32_t stage; float start_ms; float duration_ms; } mlv_profile_event;
MLV_API int mlv_profile_read_events(mlv_device *dev, mlv_profile_event *out, uint32_t capacity, uint32_t *out_count);

/* Device-side timeline: the frame as it really overlaps, INSIDE a recorded command list and across the library's streams
 * (the event brackets above serialise the launches and cannot be recorded). Every geometry / binning / tile launch issued
 * or recorded between mlv_timeline_begin and mlv_timeline_end carries a slot into which its CTAs stamp %globaltimer:
 * first CTA resident, first CTA past the wait for the previous kernel of its stream, last CTA done. mlv_timeline_reset
 * (stream-ordered on the device stream, not recordable) re-arms the slots before the execution to be looked at;
 * mlv_timeline_read synchronises and returns one event per slot in issue order, microseconds relative to the earliest
 * stamp (-1: the launch has not run since the reset). The stamps cost two atomics per CTA; launches issued while the
 * timeline is off carry no slot. tools/timeline_frame.py writes a Chrome/Perfetto trace under the reference's Remotery
 * scope names (main.c:663,699,737,916,984,1047). */
typedef struct mlv_timeline_event { int32_t stage; int32_t draw; double resident_us, start_us, end_us; } mlv_timeline_event;
MLV_API int mlv_timeline_begin(mlv_device *dev);
MLV_API int mlv_timeline_end(mlv_device *dev);
MLV_API int mlv_timeline_reset(mlv_device *dev);
MLV_API int mlv_timeline_read(mlv_device *dev, mlv_timeline_event *out, uint32_t capacity, uint32_t *out_count);

/* how many kernels this device has launched since creation (bench.py's gpu_launches) */
MLV_API uint64_t mlv_kernel_launch_count(mlv_device *dev);
/* word-wise 64-bit FNV-1a over u32 words: the frame hash used by tests/golden/golden.json and bench.py (host-side helper) */
MLV_API uint64_t mlv_fnv64_words(const uint32_t *words, size_t count);

#ifdef __cplusplus
}
#endif
#endif /* MALEVICH_B200_H */
