// malevich_b200.cu -- C-ABI host layer over the sm_100a kernels (include/malevich_b200.h).
//
// Mirrors the reference's L4 boundary: a bound-state object that the host mutates
// (graphics_pipeline main.c:71-115,222) and three entry points (clear_render_target_view :1191,
// clear_depth_stencil_view :1204, draw_indexed :1219). Everything a draw needs stays on the device:
// one draw = six or seven stream-ordered kernel launches, no host synchronisation, no per-draw allocation
// (the reference mallocs/frees seven intermediates per draw, main.c:1222-1259).
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <new>
#include <algorithm>
#include <utility>
#include <vector>

#include <cuda.h> // driver-API types only: the entry points are fetched through cudaGetDriverEntryPoint, libcuda is not linked

#include <dlfcn.h>
#include <nvtx3/nvToolsExt.h>

#include "kernels.cuh"

using namespace mlv;

static const uint32_t k_rsqrt_lut_host[2048] = {
#include "rsqrt_lut.inc"
};

static thread_local char g_last_error[512] = "";

// NVTX ranges under the names of the reference's Remotery scopes (rmt_BeginCPUSample, main.c:663, 699, 737, 916, 984,
// 1047, 1192, 1205, 1220, 1302): a timeline tool (Nsight Systems) shows the host side of every entry point and correlates
// the kernels launched inside each range. Header-only NVTX 3: a no-op costing one pointer test when no tool is attached.
struct NvtxScope {
	explicit NvtxScope(const char *name) { nvtxRangePushA(name); }
	~NvtxScope() { nvtxRangePop(); }
};

static int fail(int code, const char *fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
	va_end(ap);
	return code;
}

#define CUDA_TRY(expr)                                                                                               \
	do {                                                                                                             \
		cudaError_t _e = (expr);                                                                                     \
		if(_e != cudaSuccess) return fail(_e == cudaErrorMemoryAllocation ? MLV_ERR_OUT_OF_MEMORY : MLV_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
	} while(0)

struct mlv_buffer {
	std::vector<mlv_buffer *> *children; // buffer of a device group (num_gpus > 1): one replica per GPU, nothing else is used
	void *d;
	size_t bytes;
	int kind;
	uint64_t uid;              // unique per buffer object for the life of the process (a freed buffer's address can come back)
	uint64_t version;          // bumped by every update
	cudaEvent_t ready;         // recorded on the copy stream after the last upload
	bool ready_pending;        // no draw has waited on `ready` yet
	// sort-first chunk bounds cached with the buffer that defines the triangle list (index buffer, or vertex buffer for mlv_draw)
	float4 *chunk_bounds;
	uint32_t chunk_capacity, chunk_count;
	uint64_t chunk_vb_uid;
	uint64_t chunk_vb_version, chunk_self_version;
	int chunk_indexed;
	uint32_t chunk_start_index, chunk_index16;
	int32_t chunk_base_vertex;
};
struct mlv_texture {
	std::vector<mlv_texture *> *children; // texture of a device group: one replica per GPU
	void *d;
	uint32_t width, height;
	int format;
	void *mips;          // levels 1 .. mip_levels-1 (mlv_texture_generate_mips), null = level 0 only
	uint32_t mip_levels; // including level 0
};

struct GeomNode { // a geometry kernel node of a recorded command list (its constants can be replaced without re-recording)
	cudaGraphNode_t node;
	cudaKernelNodeParams params; // kernelParams point at gp / extra below
	GeomParams gp;
	uint32_t extra;              // k_vertex's second parameter (vertex count)
	void *argv[2];
};

struct mlv_command_list {
	std::vector<mlv_command_list *> *children; // command list of a device group: one recording per GPU
	mlv_device *owner;
	cudaGraph_t graph;
	cudaGraphExec_t exec;
	uint64_t launches;               // kernels per execution
	uint32_t draws;
	int fb_sel;                      // the tiled framebuffer of the pair the recorded kernels address
	bool opens_with_full_clear;      // nothing of the framebuffer's old contents survives the recording's first operation
	bool has_resolve;
	uint32_t last_index_count, last_direct_slots;
	std::vector<mlv_buffer *> *buffers;    // every buffer a recorded draw binds (an execution waits for uploads still in flight)
	std::vector<const void *> *geom_funcs; // kernels whose first parameter is a GeomParams
	std::vector<const void *> *vertex_funcs; // ... and that take a second u32 (k_vertex)
	std::vector<GeomNode *> *geom_nodes;
	// The PerFrameCB of the first MLV_LIST_CONSTANT_DRAWS recorded draws lives in a device arena (64 floats per draw) the
	// geometry kernels read through GeomParams::cbp: mlv_command_list_set_constants rewrites the host shadow and the next
	// execution uploads it with ONE small copy ahead of the graph launch -- no graph node is touched. (Replacing the
	// by-value constants of every geometry node cost ~35 cudaGraphExecKernelNodeSetParams per config-5 frame and made the
	// driver upload the graph again.) Draws beyond the arena keep by-value constants and the node path.
	float *d_constants, *h_constants;
	bool constants_dirty;
};
#define MLV_LIST_CONSTANT_DRAWS 256u

#define MLV_DRAW_CONTEXTS 8 /* draws whose front half may run ahead; MAX_OBJECT_COUNT_PER_SCENE of the reference is 8 (main.c:44) */
struct DrawCtx {
	uint4 *tri_bounds;
	uint32_t *big_queue, *huge_queue; // slots with large tile rectangles (capacity: every slot)
	uint32_t slot_capacity;
	uint32_t *clip_queue;
	uint32_t queue_capacity;
	uint4 *ovf_cov, *ovf_shade;
	uint32_t ovf_capacity;
	float4 *vcache; // post-transform vertex cache (k_vertex): read by the front half and, for the survivors, by the back half
	uint32_t vcache_capacity;
	uint8_t *chunk_live;
	uint32_t chunk_capacity;
	uint8_t *touch_bits;  // one byte per bin: raised by the front half (and the back half's large rectangles), counted + cleared by k_tile
	uint4 *warp_sum;      // 16 B per 32 direct slots (slot_capacity / 32 + 1 entries)
	DrawCounters *dctr;
	unsigned long long *stat_stripes; // Stats contributions of the draw, folded into Stats by its k_tile (main-stream order)
	cudaEvent_t front_done, free_ev;
	bool free_recorded;
};

struct mlv_device {
	// A device GROUP (mlv_device_desc.num_gpus > 1) is only a list of per-GPU devices, one sort-first rank each, driven by one
	// host thread: the entry points a renderer needs fan out to them (group_* below), every other one refuses a group.
	std::vector<mlv_device *> *children;
	uint32_t *group_scratch_color; // host staging for the depth images of the ranks (mlv_present_readback with depths)
	void *group_nccl;              // GroupNccl: communicators of the ncclAllGather exchange (MLV_DEVICE_GROUP_NCCL), loaded on demand
	mlv_device_desc desc;
	int cuda_dev;
	cudaStream_t stream;
	int W, H, wt, ht;
	uint32_t num_bins;
	uint32_t sm_count; // cudaDevAttrMultiProcessorCount: persistent grids are sized in multiples of it
	Partition part;

	// persistent render state (reference: static frame_buffer/depth_buffer/a_tile_min_depths)
	uint4 *fb;                 // the tiled framebuffer draws go to (= fb_pair[fb_sel])
	// Asynchronous exchange (mlv_composite_broadcast_async): the broadcast of frame f reads the tiled framebuffer on the
	// exchange stream while frame f+1 is already being drawn, so a frame that starts with a full clear is drawn into
	// the other framebuffer of the pair instead of waiting for the broadcast.
	uint4 *fb_pair[2];
	int last_recorded_fb;      // the framebuffer the last recording that opened with a full clear addresses (-1: none yet)
	int fb_sel;
	bool fb_busy[2];           // a broadcast on the exchange stream reads fb_pair[i]; ev_fb_free[i] fires when it is done
	cudaEvent_t ev_fb_free[2];
	bool fb_reading[2];        // an asynchronous present on the read-back stream (resolve / pack) reads fb_pair[i]; ev_fb_read[i] fires when it is done
	cudaEvent_t ev_fb_read[2];
	cudaEvent_t ev_present_src; // everything the present reads has been written (main stream)
	cudaStream_t xchg_stream;
	cudaEvent_t ev_frame_done, ev_xchg_done;
	bool xchg_pending;         // mlv_composite_join has not yet been called for the last mlv_composite_broadcast_async
	float *tile_min;
	// per-draw arenas
	uint32_t *bin_count, *bin_offset;
	unsigned long long *scan_state; // 2 x scan_blocks look-back words
	uint32_t scan_blocks;
	mlv_ref_compacted_bin *cbins;
	uint32_t *pair_ids, *pair_tmp;
	uint64_t pair_capacity;
	uint4 *tri_cov, *tri_shade;      // records of the direct slots

	// Draw contexts: what the FRONT half of a draw (its own stream, running ahead) hands to its BACK half (main stream).
	DrawCtx ctxs[MLV_DRAW_CONTEXTS];
	uint32_t num_ctx;                // 1 under debug capture (the reference-layout capture arrays exist once)
	uint64_t draw_seq;
	DrawCtx *last_ctx;
	cudaStream_t front_streams[MLV_DRAW_CONTEXTS]; // context i runs its front half on stream i % num_front_streams: independent latency chains overlap
	uint32_t num_front_streams;
	bool front_needs_sync[MLV_DRAW_CONTEXTS];      // per front stream: not yet ordered after what the main stream held when reset_front_order ran (creation, a graph execution)
	bool front_touched[MLV_DRAW_CONTEXTS];         // per front stream: has joined the capture of the command list being recorded
	cudaEvent_t ev_front_join[MLV_DRAW_CONTEXTS];
	int prio_chain, prio_front, knob_explicit_priority; // launch priorities of the draw-to-draw chain and of the front halves
	int knob_no_pdl_tile, knob_no_pdl_back;           // the front stream has not yet been ordered after what the main stream holds (creation, a graph execution)
	std::vector<void *> *graveyard;  // arenas replaced while command lists that address them exist
	uint32_t lists_alive;
	// Buffer uploads run on a copy stream: a draw waits only for the buffers it binds, so the upload of mesh k+1
	// overlaps the draw of mesh k (a host that streams its geometry every frame is otherwise PCIe-then-render serial).
	cudaStream_t copy_stream;
	cudaEvent_t ev_last_draw; // recorded after every draw: the next upload may overwrite a buffer only after the draws that read it
	bool last_draw_recorded;
	// Asynchronous present: the device-to-host copies run on their own stream, so the read-back of frame f overlaps the
	// uploads (other PCIe direction) and the rendering of frame f+1.
	cudaStream_t readback_stream;
	cudaEvent_t ev_resolved, ev_readback_done;
	bool readback_in_flight;
	uint32_t owned_parity;
	cudaEvent_t ev_main_sync;
	uint32_t bin_begin, bin_end; // bins this rank can touch: everything, or one contiguous band
	uint32_t tri_capacity; // slots (direct + overflow)
	Counters *ctr;
	uint32_t *rsqrt_lut;
	// debug capture
	DebugOut dbg;
	uint32_t dbg_tri_capacity, dbg_vertex_capacity;
	uint32_t last_index_count, last_direct_slots;
	// present / composite
	uint4 *resolved_color;
	float4 *resolved_depth;
	uint4 *gather;
	size_t chunk_bytes;
	// peer-memory compositing (mlv_composite_peer_*)
	uint4 *p2p_color[2];       // my two images (peers write their stripes into them)
	uint32_t *p2p_flags;       // my arrival words, one per source rank
	uint4 *peer_color[2][MLV_MAX_PEERS]; // every rank's images as mapped here ([.][rank] = my own)
	uint32_t *peer_flags[MLV_MAX_PEERS];
	void *ipc_opened[3 * MLV_MAX_PEERS];
	int ipc_opened_count;
	bool peers_attached, bcast_pending;
	uint32_t p2p_seq;          // frame sequence number of the last broadcast
	uint4 *present_color;      // what mlv_resolved_color_device_ptr hands out

	// bound pipeline state (graphics_pipeline)
	mlv_buffer *vb, *ib;
	int index_format; // MLV_INDEX_*
	uint32_t input_layout;
	int topology;
	int vs_id, ps_id;
	float cb[MLV_CONSTANT_BUFFER_SLOT_COUNT][64];
	size_t cb_bytes[MLV_CONSTANT_BUFFER_SLOT_COUNT];
	mlv_texture *vs_srv[MLV_SHADER_RESOURCE_SLOT_COUNT], *ps_srv[MLV_SHADER_RESOURCE_SLOT_COUNT];
	mlv_viewport viewport;
	bool viewport_set;

	bool pend_color, pend_depth;
	uint32_t clear_color;
	float clear_depth;

	uint64_t launches;
	// command-list recording (mlv_begin_command_list .. mlv_finish_command_list): the device stream is in CUDA stream capture
	mlv_command_list *recording;

	// device-side timeline (mlv_timeline_*): 4 words per launch issued or recorded while it is on
	bool timeline_on;
	unsigned long long *timeline;
	uint32_t timeline_used;
	std::vector<int> *timeline_stages; // stage << 16 | draw ordinal of the frame
	uint32_t timeline_draw;
	// per-stage profiling (mlv_profile_begin/end)
	bool prof_on;
	std::vector<cudaEvent_t> *prof_events; // pairs (start, end)
	std::vector<int> *prof_stages;
	size_t prof_used;
};

// Events recorded inside / outside a capture cannot order work on the other side: the next draw's front half starts from
// a fork of the main stream instead of the per-context events.
static void reset_front_order(mlv_device *dev) {
	cudaEventRecord(dev->ev_main_sync, dev->stream); // the fork point: front halves issued from now on come after everything the main stream holds now
	for(bool &b : dev->front_needs_sync) b = true;
	for(bool &b : dev->front_touched) b = false;
	for(DrawCtx &c : dev->ctxs) c.free_recorded = false;
}

// Every kernel goes through here: launched with programmatic stream serialisation (see pdl_prologue in kernels.cuh).
static thread_local int g_pdl = 1; // 0: the next launches are plain stream-ordered launches
static thread_local int g_launch_priority = 0;
static thread_local bool g_launch_priority_set = false;
template <typename... KArgs, typename... Args>
static void launch_pdl(void (*kernel)(KArgs...), uint32_t grid, uint32_t block, cudaStream_t stream, Args &&...args) {
	{ // One shared-memory carve-out for every kernel of the library. An SM changes its L1 / shared split only when it is
		// idle, so a kernel whose CTAs need another split than the resident ones waits for whole SMs to drain: with the
		// driver's per-kernel choice (k_front ~1 KB, k_tile 62 KB, k_back 150 KB of shared memory per SM) the back half of
		// the visible draw became resident 50 us after its inputs were ready -- when the front halves running ahead had
		// drained -- whatever the stream priorities said (device timeline, profiles/).
		static std::vector<const void *> seen;
		static const int carveout = getenv("MLV_SMEM_CARVEOUT") ? atoi(getenv("MLV_SMEM_CARVEOUT")) : 72; // percent of the maximum: 164 KB
		const void *fn = (const void *)kernel;
		if(carveout > 0 && std::find(seen.begin(), seen.end(), fn) == seen.end()) {
			cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, carveout);
			seen.push_back(fn);
		}
	}
	cudaLaunchConfig_t cfg;
	memset(&cfg, 0, sizeof(cfg));
	cfg.gridDim = dim3(grid, 1, 1);
	cfg.blockDim = dim3(block, 1, 1);
	cfg.stream = stream;
	cudaLaunchAttribute attr[2];
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[0].val.programmaticStreamSerializationAllowed = g_pdl;
	cfg.attrs = attr;
	cfg.numAttrs = 1;
	if(g_launch_priority_set) { // explicit per launch: recorded with the kernel node of a command list
		attr[1].id = cudaLaunchAttributePriority;
		attr[1].val.priority = g_launch_priority;
		cfg.numAttrs = 2;
	}
	cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

static int use_device(mlv_device *dev) {
	if(!dev) return fail(MLV_ERR_INVALID_ARGUMENT, "null device");
	if(dev->children) return fail(MLV_ERR_STATE, "this entry point is not available on a device group (mlv_device_desc.num_gpus > 1)");
	CUDA_TRY(cudaSetDevice(dev->cuda_dev));
	return MLV_OK;
}

// Waits on `st` until every rank's arrival word has reached `seq` (cyclic comparison, like k_composite_wait). A stream
// memory operation (cuStreamWaitValue32) when the driver offers it: the wait then occupies no SM -- a spinning kernel would
// sit resident next to the cooperative k_tail launches of the next frame, which need every CTA slot they were sized for.
// Fallback (MLV_COMPOSITE_WAIT_KERNEL=1 forces it): the bounded spin kernel.
typedef CUresult (*mlv_pfn_wait32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
static bool wait_flags_on_stream(mlv_device *dev, cudaStream_t st, uint32_t seq) {
	static mlv_pfn_wait32 fn = nullptr;
	static bool resolved = false;
	if(!resolved) {
		resolved = true;
		const char *force = getenv("MLV_COMPOSITE_WAIT_KERNEL");
		if(!(force && atoi(force))) {
			void *p = nullptr;
			cudaDriverEntryPointQueryResult q;
			if(cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (mlv_pfn_wait32)p;
			cudaGetLastError();
		}
	}
	if(fn) {
		bool ok = true;
		for(int p = 0; p < dev->part.num_ranks && ok; ++p)
			ok = fn((CUstream)st, (CUdeviceptr)(uintptr_t)(dev->p2p_flags + p), (cuuint32_t)seq, CU_STREAM_WAIT_VALUE_GEQ) == CUDA_SUCCESS;
		if(ok) return true;
		fn = nullptr; // e.g. not supported on this device: use the kernel from now on
	}
	launch_pdl(k_composite_wait, 1, 32, st, (const uint32_t *)dev->p2p_flags, dev->part.num_ranks, seq, dev->ctr, 10000000000ull);
	return false;
}

// Entry points that synchronise, read back or use other streams cannot be part of a recorded command list.
static int immediate_only(mlv_device *dev, const char *what) {
	if(dev && dev->recording) return fail(MLV_ERR_STATE, "%s cannot be recorded into a command list (call mlv_finish_command_list first)", what);
	return MLV_OK;
}

// Brackets one kernel launch with CUDA events when profiling is on: prof_pre before the <<<>>>, check_launch after.
static void prof_pre(mlv_device *dev, int stage) {
	if(!dev->prof_on) return;
	if(dev->prof_used + 2 > dev->prof_events->size()) {
		for(int i = 0; i < 2; ++i) {
			cudaEvent_t e;
			cudaEventCreate(&e);
			dev->prof_events->push_back(e);
		}
	}
	dev->prof_stages->push_back(stage);
	cudaEventRecord((*dev->prof_events)[dev->prof_used], dev->stream);
}

// The next launch's slot of the device-side timeline (null when it is off or full).
#define MLV_TIMELINE_CAPACITY 4096u
static unsigned long long *timeline_slot(mlv_device *dev, int stage) {
	if(!dev->timeline_on || dev->timeline_used >= MLV_TIMELINE_CAPACITY) return nullptr;
	dev->timeline_stages->push_back((stage << 16) | (int)(dev->timeline_draw & 0xffffu));
	return dev->timeline + 4 * (size_t)dev->timeline_used++;
}

static int check_launch(mlv_device *dev, const char *what) {
	cudaError_t e = cudaGetLastError();
	if(e != cudaSuccess) return fail(MLV_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
	if(dev->recording) dev->recording->launches++;
	else dev->launches++;
	if(dev->prof_on) {
		cudaEventRecord((*dev->prof_events)[dev->prof_used + 1], dev->stream);
		dev->prof_used += 2;
	}
	return MLV_OK;
}

static int check_launch_only(mlv_device *, const char *what) {
	cudaError_t e = cudaGetLastError();
	if(e != cudaSuccess) return fail(MLV_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
	return MLV_OK;
}

extern "C" {

const char *mlv_last_error_string(void) { return g_last_error; }

// ---- device groups: fan-out helpers ------------------------------------------------------------------
#define GROUP_EACH(dev, expr)                                                      \
	if((dev) && (dev)->children) {                                                 \
		for(size_t gi = 0; gi < (dev)->children->size(); ++gi) {                   \
			mlv_device *c = (*(dev)->children)[gi];                                \
			(void)c;                                                               \
			if(int rc = (expr)) return rc;                                         \
		}                                                                          \
		return MLV_OK;                                                             \
	}
#define GROUP_CHILD(obj) ((obj) ? (*(obj)->children)[gi] : nullptr)
// a resource handed to a device group must be one the SAME group created (one replica per GPU)
#define GROUP_OWNS(dev, obj)                                                                                                        \
	if((dev) && (dev)->children && (obj) && (!(obj)->children || (obj)->children->size() != (dev)->children->size()))                 \
		return fail(MLV_ERR_INVALID_ARGUMENT, "the resource does not belong to this device group");                                     \
	if((dev) && !(dev)->children && (obj) && (obj)->children) return fail(MLV_ERR_INVALID_ARGUMENT, "the resource belongs to a device group, not to this device");
static int group_create(const mlv_device_desc *desc, mlv_device **out_device);
static void group_destroy(mlv_device *dev);
static int group_present(mlv_device *dev, uint32_t *colors, float *depths, bool wait);
static int group_present_wait(mlv_device *dev);
static int mlv_present_copy_async(mlv_device *dev, uint32_t *colors);

int mlv_create_device(const mlv_device_desc *desc, mlv_device **out_device) {
	if(desc && out_device && desc->num_gpus > 1) return group_create(desc, out_device);
	if(!desc || !out_device) return fail(MLV_ERR_INVALID_ARGUMENT, "null argument");
	*out_device = nullptr;
	if(desc->width == 0 || desc->height == 0 || (desc->width % 8) || (desc->height % 8))
		return fail(MLV_ERR_INVALID_ARGUMENT, "render target %ux%u: width and height must be non-zero multiples of 8 (reference WIDTH_IN_TILES main.c:28-29)", desc->width,
		            desc->height);
	if(desc->width > 32768 || desc->height > 32768) return fail(MLV_ERR_INVALID_ARGUMENT, "render target too large");
	const uint32_t num_ranks = desc->num_ranks ? desc->num_ranks : 1;
	if(desc->rank >= num_ranks) return fail(MLV_ERR_INVALID_ARGUMENT, "rank %u out of range for %u ranks", desc->rank, num_ranks);

	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if(e != cudaSuccess || count == 0) return fail(MLV_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
	int cuda_dev = desc->cuda_device;
	if(cuda_dev < 0) CUDA_TRY(cudaGetDevice(&cuda_dev));
	if(cuda_dev >= count) return fail(MLV_ERR_INVALID_ARGUMENT, "cuda_device %d out of range (%d devices)", cuda_dev, count);
	cudaDeviceProp prop;
	CUDA_TRY(cudaGetDeviceProperties(&prop, cuda_dev));
	if(prop.major != 10) return fail(MLV_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", cuda_dev, prop.major, prop.minor);
	CUDA_TRY(cudaSetDevice(cuda_dev));

	mlv_device *dev = new(std::nothrow) mlv_device();
	if(!dev) return fail(MLV_ERR_OUT_OF_MEMORY, "host allocation failed");
	memset(dev, 0, sizeof(*dev));
	dev->desc = *desc;
	dev->cuda_dev = cuda_dev;
	dev->sm_count = prop.multiProcessorCount > 0 ? (uint32_t)prop.multiProcessorCount : 148u;
	dev->W = (int)desc->width;
	dev->H = (int)desc->height;
	dev->wt = dev->W / 8;
	dev->ht = dev->H / 8;
	dev->num_bins = (uint32_t)(dev->wt * dev->ht);
	dev->part.num_ranks = (int)num_ranks;
	dev->part.rank = (int)desc->rank;
	dev->part.stripe_h = desc->stripe_height_tiles ? (int)desc->stripe_height_tiles : 1;
	dev->bin_begin = 0;
	dev->bin_end = dev->num_bins;
	if(num_ranks > 1 && (uint64_t)dev->part.stripe_h * num_ranks >= (uint64_t)dev->ht) { // one contiguous band of tile rows per rank
		const uint32_t row_lo = (uint32_t)dev->part.stripe_h * desc->rank;
		const uint32_t row_hi = row_lo + (uint32_t)dev->part.stripe_h;
		dev->bin_begin = (row_lo < (uint32_t)dev->ht ? row_lo : (uint32_t)dev->ht) * (uint32_t)dev->wt;
		dev->bin_end = (row_hi < (uint32_t)dev->ht ? row_hi : (uint32_t)dev->ht) * (uint32_t)dev->wt;
		dev->bin_begin &= ~3u; // k_bin_scan uses 16-byte accesses
	}
	dev->pair_capacity = desc->max_pairs_per_draw ? desc->max_pairs_per_draw : (16ull << 20);
	if(dev->pair_capacity > 0xfffffff0ull) dev->pair_capacity = 0xfffffff0ull;
	dev->vs_id = -1;
	dev->ps_id = -1;
	dev->prof_events = new std::vector<cudaEvent_t>();
	dev->prof_stages = new std::vector<int>();

#define CREATE_TRY(expr)                          \
	do {                                          \
		cudaError_t _e2 = (expr);                 \
		if(_e2 != cudaSuccess) {                  \
			int rc = fail(_e2 == cudaErrorMemoryAllocation ? MLV_ERR_OUT_OF_MEMORY : MLV_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e2)); \
			mlv_destroy_device(dev);              \
			return rc;                            \
		}                                         \
	} while(0)

	{ // The device stream carries the draw-to-draw dependency chain (back half, binning, tile kernel): its CTAs go first when
		// SM slots free up. At equal priority the front halves running ahead on their own streams kept the chain's next kernel
		// waiting for slots (device timeline, config 5: the visible draw's back half became resident 50 us after its inputs
		// were ready, the first hidden draw's 37 us after the tile kernel before it had finished).
		int prio_lo = 0, prio_hi = 0;
		CREATE_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
		const char *p = getenv("MLV_MAIN_PRIORITY");
		CREATE_TRY(cudaStreamCreateWithPriority(&dev->stream, cudaStreamNonBlocking, (p && !atoi(p)) ? prio_lo : prio_hi));
		dev->prio_chain = prio_hi;
		dev->prio_front = prio_lo;
		dev->knob_explicit_priority = getenv("MLV_EXPLICIT_PRIORITY") ? atoi(getenv("MLV_EXPLICIT_PRIORITY")) : 1;
	}
	dev->graveyard = new std::vector<void *>();
	dev->num_ctx = (desc->flags & MLV_DEVICE_DEBUG_CAPTURE) ? 1u : MLV_DRAW_CONTEXTS;
	{
		const char *e = getenv("MLV_FRONT_STREAMS");
		uint32_t n = e ? (uint32_t)atoi(e) : 2u; // measured: two independent front chains overlap best with the draw-to-draw chain (1: 0.727, 2: 0.721, 4: 0.749 ms per config-5 frame)
		if(n < 1) n = 1;
		if(n > dev->num_ctx) n = dev->num_ctx;
		dev->num_front_streams = n;
		int prio_lo = 0, prio_hi = 0;
		CREATE_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
		const char *p = getenv("MLV_FRONT_PRIORITY");
		const int prio = (p && atoi(p)) ? prio_hi : prio_lo;
		for(uint32_t i = 0; i < n; ++i) CREATE_TRY(cudaStreamCreateWithPriority(&dev->front_streams[i], cudaStreamNonBlocking, prio));
		dev->knob_no_pdl_tile = getenv("MLV_NO_PDL_TILE") ? atoi(getenv("MLV_NO_PDL_TILE")) : 0;
		dev->knob_no_pdl_back = getenv("MLV_NO_PDL_BACK") ? atoi(getenv("MLV_NO_PDL_BACK")) : 0;
	}
	CREATE_TRY(cudaStreamCreateWithFlags(&dev->copy_stream, cudaStreamNonBlocking));
	CREATE_TRY(cudaEventCreateWithFlags(&dev->ev_last_draw, cudaEventDisableTiming));
	{ // the resolve / pack of an asynchronous present is short and the frame's way out: its CTAs go first when SM slots free up
		// (at the lowest priority the next frame's kernels starved it: the copy, and with it the host's next frame, came a frame late)
		int prio_lo = 0, prio_hi = 0;
		CREATE_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
		CREATE_TRY(cudaStreamCreateWithPriority(&dev->readback_stream, cudaStreamNonBlocking, prio_hi));
	}
	CREATE_TRY(cudaEventCreateWithFlags(&dev->ev_resolved, cudaEventDisableTiming));
	CREATE_TRY(cudaEventCreateWithFlags(&dev->ev_readback_done, cudaEventDisableTiming));
	CREATE_TRY(cudaEventCreateWithFlags(&dev->ev_main_sync, cudaEventDisableTiming));
	for(uint32_t i = 0; i < dev->num_ctx; ++i) {
		DrawCtx &c = dev->ctxs[i];
		CREATE_TRY(cudaEventCreateWithFlags(&c.front_done, cudaEventDisableTiming));
		CREATE_TRY(cudaEventCreateWithFlags(&c.free_ev, cudaEventDisableTiming));
		CREATE_TRY(cudaMalloc(&c.dctr, sizeof(DrawCounters)));
		CREATE_TRY(cudaMalloc(&c.stat_stripes, MLV_STAT_STRIPES * 128));
		CREATE_TRY(cudaMemsetAsync(c.dctr, 0, sizeof(DrawCounters), dev->stream));
		CREATE_TRY(cudaMemsetAsync(c.stat_stripes, 0, MLV_STAT_STRIPES * 128, dev->stream));
		const size_t touch_bytes = ((size_t)dev->num_bins + 15) / 16 * 16;
		CREATE_TRY(cudaMalloc(&c.touch_bits, touch_bytes));
		CREATE_TRY(cudaMemsetAsync(c.touch_bits, 0, touch_bytes, dev->stream));
	}
	CREATE_TRY(cudaEventCreateWithFlags(&dev->ev_fb_read[0], cudaEventDisableTiming));
	CREATE_TRY(cudaEventCreateWithFlags(&dev->ev_fb_read[1], cudaEventDisableTiming));
	CREATE_TRY(cudaEventCreateWithFlags(&dev->ev_present_src, cudaEventDisableTiming));
	if(num_ranks > 1) {
		int prio_lo = 0, prio_hi = 0; // the exchange is short and latency-critical: its CTAs go first when SM slots free up
		CREATE_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
		CREATE_TRY(cudaStreamCreateWithPriority(&dev->xchg_stream, cudaStreamNonBlocking, prio_hi));
		CREATE_TRY(cudaEventCreateWithFlags(&dev->ev_frame_done, cudaEventDisableTiming));
		CREATE_TRY(cudaEventCreateWithFlags(&dev->ev_xchg_done, cudaEventDisableTiming));
		CREATE_TRY(cudaEventCreateWithFlags(&dev->ev_fb_free[0], cudaEventDisableTiming));
		CREATE_TRY(cudaEventCreateWithFlags(&dev->ev_fb_free[1], cudaEventDisableTiming));
	}
	const size_t nb = dev->num_bins;
	CREATE_TRY(cudaMalloc(&dev->fb_pair[0], nb * 32 * sizeof(uint4)));
	dev->fb = dev->fb_pair[0];
	dev->last_recorded_fb = -1;
	CREATE_TRY(cudaMalloc(&dev->tile_min, nb * sizeof(float)));
	CREATE_TRY(cudaMalloc(&dev->bin_count, nb * sizeof(uint32_t)));
	CREATE_TRY(cudaMalloc(&dev->bin_offset, nb * sizeof(uint32_t)));
	CREATE_TRY(cudaMalloc(&dev->cbins, nb * sizeof(mlv_ref_compacted_bin)));
	CREATE_TRY(cudaMalloc(&dev->pair_ids, dev->pair_capacity * sizeof(uint32_t)));
	// (pair_tmp, the scratch of the debug-capture list sort, is allocated with the first debug draw)
	CREATE_TRY(cudaMalloc(&dev->ctr, sizeof(Counters)));
	dev->scan_blocks = (dev->bin_end - dev->bin_begin + MLV_SCAN_THREADS * MLV_SCAN_ITEMS - 1) / (MLV_SCAN_THREADS * MLV_SCAN_ITEMS);
	if(dev->scan_blocks == 0) dev->scan_blocks = 1;
	CREATE_TRY(cudaMalloc(&dev->scan_state, (size_t)dev->scan_blocks * 2 * sizeof(unsigned long long)));
	CREATE_TRY(cudaMemsetAsync(dev->scan_state, 0, (size_t)dev->scan_blocks * 2 * sizeof(unsigned long long), dev->stream));
	CREATE_TRY(cudaMalloc(&dev->rsqrt_lut, sizeof(k_rsqrt_lut_host)));
	CREATE_TRY(cudaMalloc(&dev->resolved_color, (size_t)dev->W * dev->H * 4));
	dev->present_color = dev->resolved_color;
	CREATE_TRY(cudaMalloc(&dev->resolved_depth, (size_t)dev->W * dev->H * 4));
	CREATE_TRY(cudaMemsetAsync(dev->fb, 0, nb * 32 * sizeof(uint4), dev->stream));
	// the second framebuffer of the pair: a frame that opens with a full clear is drawn into the one that nothing still reads
	// (the exchange of the previous frame, or its asynchronous present on the read-back stream)
	CREATE_TRY(cudaMalloc(&dev->fb_pair[1], nb * 32 * sizeof(uint4)));
	CREATE_TRY(cudaMemsetAsync(dev->fb_pair[1], 0, nb * 32 * sizeof(uint4), dev->stream));
	CREATE_TRY(cudaMemsetAsync(dev->tile_min, 0, nb * sizeof(float), dev->stream));
	CREATE_TRY(cudaMemsetAsync(dev->bin_count, 0, nb * sizeof(uint32_t), dev->stream));
	CREATE_TRY(cudaMemsetAsync(dev->ctr, 0, sizeof(Counters), dev->stream));
	{
		static const uint32_t first_epoch = 1u; // epoch 0 marks a look-back word as never published
		CREATE_TRY(cudaMemcpyAsync((char *)dev->ctr + offsetof(Counters, epoch), &first_epoch, sizeof(uint32_t), cudaMemcpyHostToDevice, dev->stream));
	}
	CREATE_TRY(cudaMemcpyAsync(dev->rsqrt_lut, k_rsqrt_lut_host, sizeof(k_rsqrt_lut_host), cudaMemcpyHostToDevice, dev->stream));
	if(num_ranks > 1) {
		const uint32_t sh = (uint32_t)dev->part.stripe_h;
		const uint32_t num_stripes = ((uint32_t)dev->ht + sh - 1) / sh;
		const uint32_t local_stripes = (num_stripes + num_ranks - 1) / num_ranks;
		dev->chunk_bytes = (size_t)local_stripes * sh * 8 * dev->W * 4;
		CREATE_TRY(cudaMalloc(&dev->gather, dev->chunk_bytes * num_ranks));
		CREATE_TRY(cudaMemsetAsync(dev->gather, 0, dev->chunk_bytes * num_ranks, dev->stream));
	}
	CREATE_TRY(cudaStreamSynchronize(dev->stream));
	for(uint32_t i = 0; i < dev->num_front_streams; ++i) CREATE_TRY(cudaEventCreateWithFlags(&dev->ev_front_join[i], cudaEventDisableTiming));
	reset_front_order(dev);
#undef CREATE_TRY
	*out_device = dev;
	return MLV_OK;
}

static void free_command_list(mlv_command_list *list);

void mlv_destroy_device(mlv_device *dev) {
	if(dev && dev->children) return group_destroy(dev);
	if(!dev) return;
	cudaSetDevice(dev->cuda_dev);
	if(dev->recording) { // abandon a recording in progress
		cudaGraph_t g = nullptr;
		cudaStreamEndCapture(dev->stream, &g);
		dev->recording->graph = g;
		free_command_list(dev->recording);
		dev->recording = nullptr;
		cudaGetLastError();
	}
	if(dev->stream) cudaStreamSynchronize(dev->stream);
	for(cudaStream_t fs : dev->front_streams)
		if(fs) cudaStreamSynchronize(fs);
	if(dev->copy_stream) cudaStreamSynchronize(dev->copy_stream);
	if(dev->xchg_stream) cudaStreamSynchronize(dev->xchg_stream);
	for(int i = 0; i < dev->ipc_opened_count; ++i) cudaIpcCloseMemHandle(dev->ipc_opened[i]);
	void *ptrs[] = { dev->fb_pair[0], dev->fb_pair[1], dev->tile_min, dev->bin_count, dev->bin_offset, dev->scan_state, dev->cbins, dev->pair_ids, dev->pair_tmp, dev->tri_cov, dev->tri_shade,
		             dev->ctr, dev->rsqrt_lut, dev->dbg.tris, dev->dbg.attrs, dev->dbg.slot_key, dev->dbg.vs_out, dev->dbg.infos, dev->resolved_color, dev->resolved_depth, dev->gather, dev->p2p_color[0], dev->p2p_color[1], dev->p2p_flags };
	for(void *p : ptrs)
		if(p) cudaFree(p);
	for(cudaEvent_t e : dev->ev_front_join)
		if(e) cudaEventDestroy(e);
	for(DrawCtx &c : dev->ctxs) {
		for(void *p : { (void *)c.tri_bounds, (void *)c.big_queue, (void *)c.huge_queue, (void *)c.clip_queue, (void *)c.ovf_cov, (void *)c.ovf_shade, (void *)c.vcache, (void *)c.chunk_live, (void *)c.touch_bits, (void *)c.warp_sum, (void *)c.dctr, (void *)c.stat_stripes })
			if(p) cudaFree(p);
		if(c.front_done) cudaEventDestroy(c.front_done);
		if(c.free_ev) cudaEventDestroy(c.free_ev);
	}
	if(dev->graveyard) {
		for(void *p : *dev->graveyard) cudaFree(p);
		delete dev->graveyard;
	}
	if(dev->stream) cudaStreamDestroy(dev->stream);
	for(cudaStream_t fs : dev->front_streams)
		if(fs) cudaStreamDestroy(fs);
	if(dev->copy_stream) cudaStreamDestroy(dev->copy_stream);
	if(dev->xchg_stream) cudaStreamDestroy(dev->xchg_stream);
	if(dev->readback_stream) {
		cudaStreamSynchronize(dev->readback_stream);
		cudaStreamDestroy(dev->readback_stream);
	}
	for(cudaEvent_t e : { dev->ev_last_draw, dev->ev_resolved, dev->ev_readback_done, dev->ev_main_sync, dev->ev_frame_done, dev->ev_xchg_done, dev->ev_fb_free[0], dev->ev_fb_free[1], dev->ev_fb_read[0], dev->ev_fb_read[1], dev->ev_present_src })
		if(e) cudaEventDestroy(e);
	if(dev->prof_events) {
		for(cudaEvent_t e : *dev->prof_events) cudaEventDestroy(e);
		delete dev->prof_events;
	}
	delete dev->prof_stages;
	if(dev->timeline) cudaFree(dev->timeline);
	delete dev->timeline_stages;
	delete dev;
}

int mlv_finish(mlv_device *dev) {
	GROUP_EACH(dev, mlv_finish(c));
	if(int rc = immediate_only(dev, "mlv_finish")) return rc;
	if(int rc = use_device(dev)) return rc;
	CUDA_TRY(cudaStreamSynchronize(dev->stream));
	for(uint32_t i = 0; i < dev->num_front_streams; ++i) CUDA_TRY(cudaStreamSynchronize(dev->front_streams[i]));
	CUDA_TRY(cudaStreamSynchronize(dev->copy_stream)); // uploads no draw has consumed yet
	CUDA_TRY(cudaStreamSynchronize(dev->readback_stream));
	dev->readback_in_flight = false;
	dev->fb_reading[0] = dev->fb_reading[1] = false;
	if(dev->xchg_stream) {
		CUDA_TRY(cudaStreamSynchronize(dev->xchg_stream));
		dev->fb_busy[0] = dev->fb_busy[1] = false;
	}
	return MLV_OK;
}

void *mlv_get_stream(mlv_device *dev) { return dev ? (void *)dev->stream : nullptr; }

// ---- resources -----------------------------------------------------------------------------------

int mlv_create_buffer(mlv_device *dev, const void *data, size_t bytes, int kind, mlv_buffer **out) {
	if(dev && dev->children) {
		if(!out) return fail(MLV_ERR_INVALID_ARGUMENT, "bad buffer arguments");
		mlv_buffer *g = new mlv_buffer();
		memset(g, 0, sizeof(*g));
		g->children = new std::vector<mlv_buffer *>();
		g->bytes = bytes, g->kind = kind;
		for(mlv_device *c : *dev->children) {
			mlv_buffer *b = nullptr;
			if(int rc = mlv_create_buffer(c, data, bytes, kind, &b)) {
				mlv_release_buffer(dev, g);
				return rc;
			}
			g->children->push_back(b);
		}
		*out = g;
		return MLV_OK;
	}
	if(int rc = use_device(dev)) return rc;
	if(!out || bytes == 0 || (kind != MLV_BUFFER_VERTEX && kind != MLV_BUFFER_INDEX)) return fail(MLV_ERR_INVALID_ARGUMENT, "bad buffer arguments");
	mlv_buffer *b = new(std::nothrow) mlv_buffer();
	if(!b) return fail(MLV_ERR_OUT_OF_MEMORY, "host allocation failed");
	memset(b, 0, sizeof(*b));
	b->bytes = bytes;
	b->kind = kind;
	static uint64_t next_uid = 0;
	b->uid = __atomic_add_fetch(&next_uid, 1, __ATOMIC_RELAXED);
	cudaError_t e = cudaMalloc(&b->d, (bytes + 15) & ~(size_t)15);
	if(e == cudaSuccess && (e = cudaEventCreateWithFlags(&b->ready, cudaEventDisableTiming)) != cudaSuccess) cudaFree(b->d);
	if(e != cudaSuccess) {
		delete b;
		return fail(MLV_ERR_OUT_OF_MEMORY, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
	}
	*out = b;
	if(data) return mlv_update_buffer(dev, b, data, bytes);
	return MLV_OK;
}

int mlv_update_buffer(mlv_device *dev, mlv_buffer *buf, const void *data, size_t bytes) {
	GROUP_OWNS(dev, buf);
	GROUP_EACH(dev, mlv_update_buffer(c, GROUP_CHILD(buf), data, bytes));
	if(int rc = use_device(dev)) return rc;
	if(!buf || !data || bytes > buf->bytes) return fail(MLV_ERR_INVALID_ARGUMENT, "bad buffer update");
	// after every draw issued so far (any of them may read this buffer; each k_vertex on the side stream has already
	// been joined into the main stream by its k_geom), before the first draw that binds it afterwards. Only draws
	// read buffers: a resolve or read-back issued since the last draw does not hold the upload back.
	if(dev->last_draw_recorded) CUDA_TRY(cudaStreamWaitEvent(dev->copy_stream, dev->ev_last_draw, 0));
	CUDA_TRY(cudaMemcpyAsync(buf->d, data, bytes, cudaMemcpyHostToDevice, dev->copy_stream));
	CUDA_TRY(cudaEventRecord(buf->ready, dev->copy_stream));
	buf->ready_pending = true;
	buf->version++;
	return MLV_OK;
}

int mlv_update_buffer_range(mlv_device *dev, mlv_buffer *buf, size_t offset, const void *data, size_t bytes) {
	GROUP_OWNS(dev, buf);
	GROUP_EACH(dev, mlv_update_buffer_range(c, GROUP_CHILD(buf), offset, data, bytes));
	if(int rc = use_device(dev)) return rc;
	if(!buf || !data || offset > buf->bytes || bytes > buf->bytes - offset) return fail(MLV_ERR_INVALID_ARGUMENT, "bad buffer range update");
	if(dev->last_draw_recorded) CUDA_TRY(cudaStreamWaitEvent(dev->copy_stream, dev->ev_last_draw, 0));
	if(bytes) CUDA_TRY(cudaMemcpyAsync((char *)buf->d + offset, data, bytes, cudaMemcpyHostToDevice, dev->copy_stream));
	CUDA_TRY(cudaEventRecord(buf->ready, dev->copy_stream));
	buf->ready_pending = true;
	buf->version++;
	return MLV_OK;
}

void *mlv_buffer_device_ptr(mlv_buffer *buf) { return buf ? buf->d : nullptr; }
void *mlv_get_copy_stream(mlv_device *dev) { return dev ? (void *)dev->copy_stream : nullptr; }

int mlv_buffer_mark_updated(mlv_device *dev, mlv_buffer *buf, void *stream) {
	if(int rc = use_device(dev)) return rc;
	if(!buf) return fail(MLV_ERR_INVALID_ARGUMENT, "null buffer");
	CUDA_TRY(cudaEventRecord(buf->ready, stream ? (cudaStream_t)stream : dev->copy_stream));
	buf->ready_pending = true;
	buf->version++;
	return MLV_OK;
}

void mlv_release_buffer(mlv_device *dev, mlv_buffer *buf) {
	if(dev && dev->children) {
		if(!buf) return;
		if(buf->children) {
			for(size_t gi = 0; gi < buf->children->size(); ++gi) mlv_release_buffer((*dev->children)[gi], (*buf->children)[gi]);
			delete buf->children;
		}
		delete buf;
		return;
	}
	if(immediate_only(dev, "mlv_release_buffer")) return;
	if(!dev || !buf) return;
	cudaSetDevice(dev->cuda_dev);
	cudaStreamSynchronize(dev->stream);
	for(uint32_t i = 0; i < dev->num_front_streams; ++i) cudaStreamSynchronize(dev->front_streams[i]);
	cudaStreamSynchronize(dev->copy_stream);
	if(dev->vb == buf) dev->vb = nullptr;
	if(dev->ib == buf) dev->ib = nullptr;
	if(buf->ready) cudaEventDestroy(buf->ready);
	if(buf->chunk_bounds) cudaFree(buf->chunk_bounds);
	cudaFree(buf->d);
	delete buf;
}

int mlv_create_texture2d(mlv_device *dev, const void *texels, uint32_t width, uint32_t height, int format, mlv_texture **out) {
	if(dev && dev->children) {
		if(!out) return fail(MLV_ERR_INVALID_ARGUMENT, "bad texture arguments");
		mlv_texture *g = new mlv_texture();
		memset(g, 0, sizeof(*g));
		g->children = new std::vector<mlv_texture *>();
		g->width = width, g->height = height, g->format = format;
		for(mlv_device *c : *dev->children) {
			mlv_texture *t = nullptr;
			if(int rc = mlv_create_texture2d(c, texels, width, height, format, &t)) {
				mlv_release_texture(dev, g);
				return rc;
			}
			g->children->push_back(t);
		}
		*out = g;
		return MLV_OK;
	}
	if(int rc = use_device(dev)) return rc;
	if(!out || !texels || width == 0 || height == 0 || (format != MLV_FORMAT_R8G8B8A8_UNORM && format != MLV_FORMAT_R32G32B32A32_FLOAT))
		return fail(MLV_ERR_INVALID_ARGUMENT, "bad texture arguments");
	if((uint64_t)width * height > 0x7fffffffull / 4) return fail(MLV_ERR_INVALID_ARGUMENT, "texture too large for the reference's i32 texel addressing (common_shader_core.h:33,46)");
	mlv_texture *t = new(std::nothrow) mlv_texture();
	if(!t) return fail(MLV_ERR_OUT_OF_MEMORY, "host allocation failed");
	t->width = width;
	t->height = height;
	t->format = format;
	t->mips = nullptr;
	t->mip_levels = 1;
	const size_t bytes = (size_t)width * height * (format == MLV_FORMAT_R8G8B8A8_UNORM ? 4 : 16);
	cudaError_t e = cudaMalloc(&t->d, bytes);
	if(e == cudaSuccess) {
		if(dev->recording) { // the device stream is capturing: the upload must not become a node of the list
			e = cudaMemcpyAsync(t->d, texels, bytes, cudaMemcpyHostToDevice, dev->copy_stream);
			if(e == cudaSuccess) e = cudaStreamSynchronize(dev->copy_stream);
		} else {
			e = cudaMemcpyAsync(t->d, texels, bytes, cudaMemcpyHostToDevice, dev->stream);
		}
	}
	if(e != cudaSuccess) {
		if(t->d) cudaFree(t->d);
		delete t;
		return fail(MLV_ERR_CUDA, "texture upload: %s", cudaGetErrorString(e));
	}
	*out = t;
	return MLV_OK;
}

int mlv_texture_srgb_to_linear(mlv_device *dev, mlv_texture *tex) {
	GROUP_OWNS(dev, tex);
	GROUP_EACH(dev, mlv_texture_srgb_to_linear(c, GROUP_CHILD(tex)));
	if(int rc = immediate_only(dev, "mlv_texture_srgb_to_linear")) return rc;
	if(int rc = use_device(dev)) return rc;
	if(!tex || tex->format != MLV_FORMAT_R8G8B8A8_UNORM) return fail(MLV_ERR_INVALID_ARGUMENT, "sRGB conversion wants an R8G8B8A8 texture (main.c:546-558)");
	SrgbTable table;
	for(int b = 0; b < 256; ++b) {
		const float normalizer = (float)(1.0 / 255.0);            // math.h:328
		const float v = (float)(uint32_t)b * normalizer;          // math.h:329
		float lin;                                                 // srgb_to_linear math.h:386-395: double arithmetic, f32 result
		if((double)v <= 0.04045) lin = (float)((double)v / 12.92);
		else lin = (float)pow(((double)v + 0.055) / 1.055, 2.4);
		table.lin[b] = (uint8_t)(uint32_t)(lin * 255.f);          // encode_color_as_u32 math.h:323 (truncation)
	}
	const size_t texel_count = (size_t)tex->width * tex->height;
	const size_t count_u4 = texel_count / 4;
	const uint32_t tail = (uint32_t)(texel_count % 4);
	size_t blocks = (count_u4 + 255) / 256;
	if(blocks > dev->sm_count * 8u) blocks = dev->sm_count * 8u;
	if(blocks == 0) blocks = 1;
	launch_pdl(k_texture_srgb_to_linear, (uint32_t)blocks, 256, dev->stream, (uint4 *)tex->d, count_u4, (uint32_t *)tex->d + count_u4 * 4, tail, table);
	return check_launch(dev, "k_texture_srgb_to_linear");
}

static uint32_t mip_extent_host(uint32_t e, uint32_t level) {
	const uint32_t v = e >> level;
	return v ? v : 1u;
}

int mlv_texture_generate_mips(mlv_device *dev, mlv_texture *tex) {
	GROUP_OWNS(dev, tex);
	GROUP_EACH(dev, mlv_texture_generate_mips(c, GROUP_CHILD(tex)));
	if(int rc = immediate_only(dev, "mlv_texture_generate_mips")) return rc;
	if(int rc = use_device(dev)) return rc;
	if(!tex || tex->format != MLV_FORMAT_R8G8B8A8_UNORM) return fail(MLV_ERR_INVALID_ARGUMENT, "mip chains are built for R8G8B8A8 textures");
	uint32_t levels = 1;
	size_t texels = 0;
	while(mip_extent_host(tex->width, levels - 1) > 1u || mip_extent_host(tex->height, levels - 1) > 1u) {
		texels += (size_t)mip_extent_host(tex->width, levels) * mip_extent_host(tex->height, levels);
		++levels;
	}
	if(levels == 1) return MLV_OK; // a 1x1 texture is its own chain
	if(!tex->mips) {
		cudaError_t e = cudaMalloc(&tex->mips, texels * 4);
		if(e != cudaSuccess) {
			tex->mips = nullptr;
			return fail(MLV_ERR_OUT_OF_MEMORY, "cudaMalloc(%zu): %s", texels * 4, cudaGetErrorString(e));
		}
	}
	tex->mip_levels = levels;
	const uint32_t *src = (const uint32_t *)tex->d;
	uint32_t *dst = (uint32_t *)tex->mips;
	for(uint32_t l = 1; l < levels; ++l) {
		const uint32_t sw = mip_extent_host(tex->width, l - 1), sh = mip_extent_host(tex->height, l - 1), dw = mip_extent_host(tex->width, l), dh = mip_extent_host(tex->height, l);
		size_t blocks = ((size_t)dw * dh + 255) / 256;
		if(blocks > dev->sm_count * 8u) blocks = dev->sm_count * 8u;
		launch_pdl(k_mip_downsample, (uint32_t)blocks, 256, dev->stream, src, (int)sw, (int)sh, dst, (int)dw, (int)dh);
		if(int rc = check_launch(dev, "k_mip_downsample")) return rc;
		src = dst;
		dst += (size_t)dw * dh;
	}
	return MLV_OK;
}

int mlv_texture_mip_levels(const mlv_texture *tex, uint32_t *out_levels) {
	if(!tex || !out_levels) return fail(MLV_ERR_INVALID_ARGUMENT, "null argument");
	*out_levels = tex->mips ? tex->mip_levels : 1u;
	return MLV_OK;
}

int mlv_read_texture_mip(mlv_device *dev, const mlv_texture *tex, uint32_t level, void *out_texels) {
	if(int rc = immediate_only(dev, "mlv_read_texture_mip")) return rc;
	if(int rc = use_device(dev)) return rc;
	if(!tex || !out_texels) return fail(MLV_ERR_INVALID_ARGUMENT, "null argument");
	if(level == 0) return mlv_read_texture(dev, tex, out_texels);
	if(!tex->mips || level >= tex->mip_levels) return fail(MLV_ERR_INVALID_ARGUMENT, "texture has no mip level %u", level);
	size_t offset = 0;
	for(uint32_t l = 1; l < level; ++l) offset += (size_t)mip_extent_host(tex->width, l) * mip_extent_host(tex->height, l);
	const size_t bytes = (size_t)mip_extent_host(tex->width, level) * mip_extent_host(tex->height, level) * 4;
	CUDA_TRY(cudaMemcpyAsync(out_texels, (const uint32_t *)tex->mips + offset, bytes, cudaMemcpyDeviceToHost, dev->stream));
	CUDA_TRY(cudaStreamSynchronize(dev->stream));
	return MLV_OK;
}

int mlv_read_texture(mlv_device *dev, const mlv_texture *tex, void *out_texels) {
	if(int rc = immediate_only(dev, "mlv_read_texture")) return rc;
	if(int rc = use_device(dev)) return rc;
	if(!tex || !out_texels) return fail(MLV_ERR_INVALID_ARGUMENT, "null argument");
	const size_t bytes = (size_t)tex->width * tex->height * (tex->format == MLV_FORMAT_R8G8B8A8_UNORM ? 4 : 16);
	CUDA_TRY(cudaMemcpyAsync(out_texels, tex->d, bytes, cudaMemcpyDeviceToHost, dev->stream));
	CUDA_TRY(cudaStreamSynchronize(dev->stream));
	return MLV_OK;
}

void mlv_release_texture(mlv_device *dev, mlv_texture *tex) {
	if(dev && dev->children) {
		if(!tex) return;
		if(tex->children) {
			for(size_t gi = 0; gi < tex->children->size(); ++gi) mlv_release_texture((*dev->children)[gi], (*tex->children)[gi]);
			delete tex->children;
		}
		delete tex;
		return;
	}
	if(immediate_only(dev, "mlv_release_texture")) return;
	if(!dev || !tex) return;
	cudaSetDevice(dev->cuda_dev);
	cudaStreamSynchronize(dev->stream);
	for(int i = 0; i < MLV_SHADER_RESOURCE_SLOT_COUNT; ++i) {
		if(dev->vs_srv[i] == tex) dev->vs_srv[i] = nullptr;
		if(dev->ps_srv[i] == tex) dev->ps_srv[i] = nullptr;
	}
	cudaFree(tex->d);
	if(tex->mips) cudaFree(tex->mips);
	delete tex;
}

// ---- pipeline state ------------------------------------------------------------------------------

int mlv_ia_set_vertex_buffer(mlv_device *dev, mlv_buffer *vb) {
	GROUP_OWNS(dev, vb);
	GROUP_EACH(dev, mlv_ia_set_vertex_buffer(c, GROUP_CHILD(vb)));
	if(!dev) return fail(MLV_ERR_INVALID_ARGUMENT, "null device");
	if(vb && vb->kind != MLV_BUFFER_VERTEX) return fail(MLV_ERR_INVALID_ARGUMENT, "buffer is not a vertex buffer");
	dev->vb = vb;
	return MLV_OK;
}
int mlv_ia_set_index_buffer(mlv_device *dev, mlv_buffer *ib) {
	GROUP_OWNS(dev, ib);
	GROUP_EACH(dev, mlv_ia_set_index_buffer(c, GROUP_CHILD(ib)));
	if(!dev) return fail(MLV_ERR_INVALID_ARGUMENT, "null device");
	if(ib && ib->kind != MLV_BUFFER_INDEX) return fail(MLV_ERR_INVALID_ARGUMENT, "buffer is not an index buffer");
	dev->ib = ib;
	return MLV_OK;
}
int mlv_ia_set_index_format(mlv_device *dev, int format) {
	GROUP_EACH(dev, mlv_ia_set_index_format(c, format));
	if(!dev) return fail(MLV_ERR_INVALID_ARGUMENT, "null device");
	if(format != MLV_INDEX_U32 && format != MLV_INDEX_U16) return fail(MLV_ERR_INVALID_ARGUMENT, "unknown index format %d", format);
	dev->index_format = format;
	return MLV_OK;
}
int mlv_ia_set_input_layout(mlv_device *dev, uint32_t bytes_per_vertex) {
	GROUP_EACH(dev, mlv_ia_set_input_layout(c, bytes_per_vertex));
	if(!dev) return fail(MLV_ERR_INVALID_ARGUMENT, "null device");
	if(bytes_per_vertex != 32) return fail(MLV_ERR_INVALID_ARGUMENT, "input layout %u: every reference shader consumes 32-byte vertices (in_vertex_size/VECTOR_WIDTH main.c:1286)", bytes_per_vertex);
	dev->input_layout = bytes_per_vertex;
	return MLV_OK;
}
int mlv_ia_set_primitive_topology(mlv_device *dev, int topology) {
	GROUP_EACH(dev, mlv_ia_set_primitive_topology(c, topology));
	if(!dev) return fail(MLV_ERR_INVALID_ARGUMENT, "null device");
	dev->topology = topology;
	return MLV_OK;
}
int mlv_vs_set_shader(mlv_device *dev, int vs_id) {
	GROUP_EACH(dev, mlv_vs_set_shader(c, vs_id));
	if(!dev) return fail(MLV_ERR_INVALID_ARGUMENT, "null device");
	if(vs_id < 0 || vs_id >= MLV_VS_COUNT) return fail(MLV_ERR_INVALID_ARGUMENT, "unknown vertex shader id %d", vs_id);
	dev->vs_id = vs_id;
	return MLV_OK;
}
int mlv_vs_set_constant_buffer(mlv_device *dev, uint32_t slot, const void *data, size_t bytes) {
	GROUP_EACH(dev, mlv_vs_set_constant_buffer(c, slot, data, bytes));
	if(!dev) return fail(MLV_ERR_INVALID_ARGUMENT, "null device");
	if(slot >= MLV_CONSTANT_BUFFER_SLOT_COUNT || !data || bytes > sizeof(dev->cb[0])) return fail(MLV_ERR_INVALID_ARGUMENT, "bad constant buffer (slot %u, %zu bytes)", slot, bytes);
	memcpy(dev->cb[slot], data, bytes);
	dev->cb_bytes[slot] = bytes;
	return MLV_OK;
}
int mlv_vs_set_shader_resource(mlv_device *dev, uint32_t slot, mlv_texture *tex) {
	GROUP_OWNS(dev, tex);
	GROUP_EACH(dev, mlv_vs_set_shader_resource(c, slot, GROUP_CHILD(tex)));
	if(!dev || slot >= MLV_SHADER_RESOURCE_SLOT_COUNT) return fail(MLV_ERR_INVALID_ARGUMENT, "bad shader resource slot");
	dev->vs_srv[slot] = tex;
	return MLV_OK;
}
int mlv_rs_set_viewport(mlv_device *dev, const mlv_viewport *vp) {
	GROUP_EACH(dev, mlv_rs_set_viewport(c, vp));
	if(!dev || !vp) return fail(MLV_ERR_INVALID_ARGUMENT, "null argument");
	if((int)vp->width != dev->W || (int)vp->height != dev->H)
		return fail(MLV_ERR_INVALID_ARGUMENT, "viewport %gx%g must equal the render target %dx%d (reference tile pitch main.c:582,590)", vp->width, vp->height, dev->W, dev->H);
	dev->viewport = *vp;
	dev->viewport_set = true;
	return MLV_OK;
}
int mlv_ps_set_shader(mlv_device *dev, int ps_id) {
	GROUP_EACH(dev, mlv_ps_set_shader(c, ps_id));
	if(!dev) return fail(MLV_ERR_INVALID_ARGUMENT, "null device");
	if(ps_id < 0 || ps_id >= MLV_PS_COUNT) return fail(MLV_ERR_INVALID_ARGUMENT, "unknown pixel shader id %d", ps_id);
	dev->ps_id = ps_id;
	return MLV_OK;
}
int mlv_ps_set_shader_resource(mlv_device *dev, uint32_t slot, mlv_texture *tex) {
	GROUP_OWNS(dev, tex);
	GROUP_EACH(dev, mlv_ps_set_shader_resource(c, slot, GROUP_CHILD(tex)));
	if(!dev || slot >= MLV_SHADER_RESOURCE_SLOT_COUNT) return fail(MLV_ERR_INVALID_ARGUMENT, "bad shader resource slot");
	dev->ps_srv[slot] = tex;
	return MLV_OK;
}

// ---- clears --------------------------------------------------------------------------------------

// Called before anything on the main stream WRITES the tiled framebuffer (a clear, a draw): an asynchronous broadcast
// may still be reading it on the exchange stream. A frame that starts with a full clear moves to the other framebuffer
// of the pair (nothing of the old contents survives such a clear); anything else waits for the broadcast.
static int wait_framebuffer_readers(mlv_device *dev) { // the main stream waits for whatever still reads fb_pair[fb_sel] on another stream
	if(dev->fb_busy[dev->fb_sel]) {
		CUDA_TRY(cudaStreamWaitEvent(dev->stream, dev->ev_fb_free[dev->fb_sel], 0));
		dev->fb_busy[dev->fb_sel] = false;
	}
	if(dev->fb_reading[dev->fb_sel]) {
		CUDA_TRY(cudaStreamWaitEvent(dev->stream, dev->ev_fb_read[dev->fb_sel], 0));
		dev->fb_reading[dev->fb_sel] = false;
	}
	return MLV_OK;
}

static int acquire_framebuffer(mlv_device *dev, bool full_clear) {
	if(dev->recording) {
		// A recorded list addresses ONE framebuffer of the pair; mlv_execute_command_list waits for it. A recording that OPENS
		// with a full clear keeps nothing of what the framebuffer held, so it may address either one -- and consecutive such
		// recordings alternate: a host that records its frame twice and replays the two lists in turn never makes frame f+1
		// wait for the exchange / read-back that still reads frame f (the deferred-context form of what immediate mode does).
		mlv_command_list *l = dev->recording;
		if(l->launches == 0) {
			if(full_clear && dev->fb_pair[1] && dev->last_recorded_fb == dev->fb_sel) {
				dev->fb_sel ^= 1;
				dev->fb = dev->fb_pair[dev->fb_sel];
				l->fb_sel = dev->fb_sel;
			}
			if(full_clear) {
				dev->last_recorded_fb = l->fb_sel;
				l->opens_with_full_clear = true;
			}
		}
		return MLV_OK;
	}
	if(!(dev->fb_busy[0] || dev->fb_busy[1] || dev->fb_reading[0] || dev->fb_reading[1])) return MLV_OK;
	if(full_clear && (dev->fb_busy[dev->fb_sel] || dev->fb_reading[dev->fb_sel])) {
		dev->fb_sel ^= 1;
		dev->fb = dev->fb_pair[dev->fb_sel];
	}
	return wait_framebuffer_readers(dev);
}

static int flush_clears(mlv_device *dev) {
	if(!dev->pend_color && !dev->pend_depth) return acquire_framebuffer(dev, false);
	const int mode = (dev->pend_color ? 1 : 0) | (dev->pend_depth ? 2 : 0);
	if(int rc = acquire_framebuffer(dev, mode == 3)) return rc;
	const uint32_t n = (dev->bin_end - dev->bin_begin) * 32u; // this rank's band only (nobody reads the tiles of other ranks)
	if(n == 0) { // a rank without tile rows (more ranks than stripes)
		dev->pend_color = dev->pend_depth = false;
		return MLV_OK;
	}
	prof_pre(dev, MLV_STAGE_CLEAR);
	launch_pdl(k_clear, (n + 255) / 256, 256, dev->stream, dev->fb, dev->tile_min, dev->bin_begin, dev->bin_end, dev->clear_color, dev->clear_depth, mode);
	dev->pend_color = dev->pend_depth = false;
	return check_launch(dev, "k_clear");
}

int mlv_clear_render_target_view(mlv_device *dev, const float rgba[4]) {
	GROUP_EACH(dev, mlv_clear_render_target_view(c, rgba));
	NvtxScope nvtx("clear_render_target_view");
	if(!dev || !rgba) return fail(MLV_ERR_INVALID_ARGUMENT, "null argument");
	// encode_color_as_u32 (math.h:322-324): truncation, R in the low byte -- not the PS path's channel order (App. C 4)
	dev->clear_color = (((uint32_t)(rgba[3] * 255.f)) << 24) + (((uint32_t)(rgba[2] * 255.f)) << 16) + (((uint32_t)(rgba[1] * 255.f)) << 8) + (((uint32_t)(rgba[0] * 255.f)));
	dev->pend_color = true;
	return MLV_OK; // the fill itself is fused with the depth clear and issued with the next draw / read-back
}

int mlv_clear_depth_stencil_view(mlv_device *dev, float depth) {
	GROUP_EACH(dev, mlv_clear_depth_stencil_view(c, depth));
	NvtxScope nvtx("clear_depth_stencil_view");
	if(!dev) return fail(MLV_ERR_INVALID_ARGUMENT, "null device");
	dev->clear_depth = depth;
	dev->pend_depth = true;
	return MLV_OK;
}

// ---- draw ----------------------------------------------------------------------------------------

} // extern "C"

// While a command list is being recorded an arena that has to grow is not freed: nodes recorded earlier address it.
// The old allocation goes to the list's graveyard and lives as long as the list.
static thread_local std::vector<void *> *g_graveyard = nullptr;
template <typename T>
static cudaError_t regrow(T **p, size_t count) {
	if(*p) {
		if(g_graveyard) g_graveyard->push_back((void *)*p);
		else cudaFree(*p);
	}
	*p = nullptr;
	return cudaMalloc((void **)p, count * sizeof(T));
}


static TexDesc tex_desc(const mlv_texture *t) {
	TexDesc d;
	d.data = t ? t->d : nullptr;
	d.width = t ? (int)t->width : 0;
	d.height = t ? (int)t->height : 0;
	d.format = t ? t->format : 0;
	d.mips = t ? t->mips : nullptr;
	d.mip_levels = (t && t->mips) ? (int)t->mip_levels : 1;
	return d;
}

static void note_geom_func(mlv_device *dev, const void *func, bool is_vertex) {
	if(!dev->recording) return;
	std::vector<const void *> *v = is_vertex ? dev->recording->vertex_funcs : dev->recording->geom_funcs;
	if(std::find(v->begin(), v->end(), func) == v->end()) v->push_back(func);
}

// The three geometry kernels of a draw for one vertex shader (template dispatch on indexed / debug capture / vertex cache).
template <int VS>
static void launch_front(mlv_device *dev, cudaStream_t fs, const GeomParams &gp_in, uint32_t nblocks, bool indexed, uint32_t vcache_vertices) {
	GeomParams gp = gp_in; // (a copy per launch: each carries its own timeline slot)
	const bool debug = gp.keep_all;
	{ // persistent grid: CTAs stride over the chunks (sort-first: and cull them in place)
		static const uint32_t per_sm = getenv("MLV_FRONT_CTAS_PER_SM") ? (uint32_t)atoi(getenv("MLV_FRONT_CTAS_PER_SM")) : 4u;
		if(nblocks > dev->sm_count * per_sm) nblocks = dev->sm_count * per_sm;
	}
	if(vcache_vertices) {
		prof_pre(dev, MLV_STAGE_VERTEX);
		note_geom_func(dev, (const void *)k_vertex<VS>, true);
		gp.timeline = timeline_slot(dev, MLV_STAGE_VERTEX);
		launch_pdl(k_vertex<VS>, (vcache_vertices + 255u) / 256u, 256, fs, gp, vcache_vertices);
		check_launch(dev, "k_vertex");
	}
	prof_pre(dev, MLV_STAGE_GEOMETRY);
	gp.timeline = timeline_slot(dev, MLV_STAGE_GEOMETRY);
	const void *f;
	if(vcache_vertices) { f = (const void *)k_front<VS, true, false, true>; launch_pdl(k_front<VS, true, false, true>, nblocks, MLV_GEOM_THREADS, fs, gp); }
	else if(indexed && debug) { f = (const void *)k_front<VS, true, true, false>; launch_pdl(k_front<VS, true, true, false>, nblocks, MLV_GEOM_THREADS, fs, gp); }
	else if(indexed) { f = (const void *)k_front<VS, true, false, false>; launch_pdl(k_front<VS, true, false, false>, nblocks, MLV_GEOM_THREADS, fs, gp); }
	else if(debug) { f = (const void *)k_front<VS, false, true, false>; launch_pdl(k_front<VS, false, true, false>, nblocks, MLV_GEOM_THREADS, fs, gp); }
	else { f = (const void *)k_front<VS, false, false, false>; launch_pdl(k_front<VS, false, false, false>, nblocks, MLV_GEOM_THREADS, fs, gp); }
	note_geom_func(dev, f, false);
	check_launch(dev, "k_front");
	{ // clipping pass over the (device-side) queue; most draws queue few or no triangles. One group of MLV_CLIP_SPLIT lanes per
		// queued triangle: the clipper is a long dependent chain per triangle, so the queue is spread over as many warps as it
		// has entries (concentrating it on fewer CTAs was measured slower: 28 vs 20 us)
		uint32_t cb = (gp.tri_count * MLV_CLIP_SPLIT + MLV_CLIP_THREADS - 1u) / MLV_CLIP_THREADS;
		if(cb > dev->sm_count * 4u) cb = dev->sm_count * 4u;
		prof_pre(dev, MLV_STAGE_CLIP);
		gp.timeline = timeline_slot(dev, MLV_STAGE_CLIP);
		if(indexed) { f = (const void *)k_front_clip<VS, true>; launch_pdl(k_front_clip<VS, true>, cb, MLV_CLIP_THREADS, fs, gp); }
		else { f = (const void *)k_front_clip<VS, false>; launch_pdl(k_front_clip<VS, false>, cb, MLV_CLIP_THREADS, fs, gp); }
		note_geom_func(dev, f, false);
		check_launch(dev, "k_front_clip");
	}
}

template <int VS>
static void launch_back(mlv_device *dev, const GeomParams &gp_in, uint32_t nblocks, bool indexed, bool vcache) {
	GeomParams gp = gp_in;
	gp.timeline = timeline_slot(dev, MLV_STAGE_BACK);
	const bool debug = gp.keep_all;
	if(nblocks > dev->sm_count * 4u) nblocks = dev->sm_count * 4u;
	prof_pre(dev, MLV_STAGE_BACK);
	const void *f;
	g_pdl = dev->knob_no_pdl_back ? 0 : 1;
	if(vcache) { f = (const void *)k_back<VS, true, false, true>; launch_pdl(k_back<VS, true, false, true>, nblocks, MLV_GEOM_THREADS, dev->stream, gp); }
	else if(indexed && debug) { f = (const void *)k_back<VS, true, true, false>; launch_pdl(k_back<VS, true, true, false>, nblocks, MLV_GEOM_THREADS, dev->stream, gp); }
	else if(indexed) { f = (const void *)k_back<VS, true, false, false>; launch_pdl(k_back<VS, true, false, false>, nblocks, MLV_GEOM_THREADS, dev->stream, gp); }
	else if(debug) { f = (const void *)k_back<VS, false, true, false>; launch_pdl(k_back<VS, false, true, false>, nblocks, MLV_GEOM_THREADS, dev->stream, gp); }
	else { f = (const void *)k_back<VS, false, false, false>; launch_pdl(k_back<VS, false, false, false>, nblocks, MLV_GEOM_THREADS, dev->stream, gp); }
	note_geom_func(dev, f, false);
	g_pdl = 1;
}

// Both streams of a draw are idle (nothing recorded: nothing is executing): called before an arena is replaced.
static cudaError_t quiesce(mlv_device *dev) {
	if(dev->recording) return cudaSuccess;
	cudaError_t e = cudaStreamSynchronize(dev->stream);
	for(uint32_t i = 0; i < dev->num_front_streams && e == cudaSuccess; ++i) e = cudaStreamSynchronize(dev->front_streams[i]);
	return e;
}

static int draw_common(mlv_device *dev, uint32_t count, bool indexed, uint32_t start_index = 0, int32_t base_vertex = 0) {
	NvtxScope nvtx("draw_indexed");
	if(int rc = use_device(dev)) return rc;
	g_graveyard = (dev->recording || dev->lists_alive) ? dev->graveyard : nullptr;
	if(dev->recording && dev->prof_on) return fail(MLV_ERR_STATE, "per-stage profiling brackets launches with events and cannot be recorded");
	// the reference's asserts (main.c:666,670,1230) become argument errors
	if(dev->topology != MLV_PRIMITIVE_TOPOLOGY_TRIANGLELIST) return fail(MLV_ERR_STATE, "primitive topology must be TRIANGLELIST (main.c:666)");
	if(count & 7u) return fail(MLV_ERR_INVALID_ARGUMENT, "index_count %u is not divisible by 8 (main.c:670)", count);
	if(count % 3u) return fail(MLV_ERR_INVALID_ARGUMENT, "index_count %u is not divisible by 3 (main.c:1230)", count);
	if(dev->input_layout != 32) return fail(MLV_ERR_STATE, "input layout not set");
	if(!dev->vb) return fail(MLV_ERR_STATE, "no vertex buffer bound");
	if(indexed && !dev->ib) return fail(MLV_ERR_STATE, "no index buffer bound");
	const size_t index_bytes = dev->index_format == MLV_INDEX_U16 ? 2 : 4;
	if(indexed && ((size_t)start_index + count) * index_bytes > dev->ib->bytes) return fail(MLV_ERR_INVALID_ARGUMENT, "index range [%u, %u) exceeds the bound index buffer", start_index, start_index + count);
	if(!indexed && (size_t)count * 32 > dev->vb->bytes) return fail(MLV_ERR_INVALID_ARGUMENT, "vertex_count %u exceeds the bound vertex buffer", count);
	if(dev->vs_id < 0 || dev->ps_id < 0) return fail(MLV_ERR_STATE, "vertex and pixel shader must be bound");
	if(!dev->viewport_set) return fail(MLV_ERR_STATE, "viewport not set");
	if(dev->vs_id != MLV_VS_PASSTHROUGH && dev->cb_bytes[0] < 192) return fail(MLV_ERR_STATE, "constant buffer slot 0 must hold the 192-byte PerFrameCB (main.c:169-173)");
	if(dev->vs_id == MLV_VS_VERTEX_LIGHTING && !(dev->vs_srv[0] && dev->vs_srv[0]->format == MLV_FORMAT_R32G32B32A32_FLOAT))
		return fail(MLV_ERR_STATE, "vertex_lighting_vs needs an RGBA32F panorama in VS resource slot 0");
	if((dev->ps_id == MLV_PS_BASIC || dev->ps_id == MLV_PS_BASIC_TRILINEAR) && !(dev->ps_srv[0] && dev->ps_srv[0]->format == MLV_FORMAT_R8G8B8A8_UNORM))
		return fail(MLV_ERR_STATE, "basic_ps needs an R8G8B8A8 texture in PS resource slot 0");
	if(dev->ps_id == MLV_PS_ENV_LIGHTING && !(dev->ps_srv[0] && dev->ps_srv[0]->format == MLV_FORMAT_R32G32B32A32_FLOAT))
		return fail(MLV_ERR_STATE, "env_lighting_ps needs an RGBA32F panorama in PS resource slot 0");
	if(int rc = flush_clears(dev)) return rc;
	dev->last_index_count = count;
	if(count == 0) return MLV_OK;

	const uint32_t T = count / 3u;
	if(T >= (1u << 28)) return fail(MLV_ERR_INVALID_ARGUMENT, "draw too large: triangle keys are (input_triangle << 3 | fan_index) in 32 bits");
	// The reference bounds the triangles that SURVIVE culling by T + max(2T, 512) (main.c:739-740). Here a clipped input
	// triangle leaves its direct slot empty and takes one overflow slot per fan triangle before they are culled, so the
	// overflow arena holds max(3T, 512) slots: whatever fits the reference's bound fits here (e.g. every triangle of the
	// draw clipped into a fan of three).
	const uint32_t ovf_cap = (3u * T > 512u ? 3u * T : 512u);
	const uint32_t need_slots = T + ovf_cap;
	const uint32_t nblocks = (T + MLV_GEOM_THREADS - 1) / MLV_GEOM_THREADS;
	const bool debug = (dev->desc.flags & MLV_DEVICE_DEBUG_CAPTURE) != 0;
	// Stream of the front half. Per-stage profiling brackets every kernel with events on the main stream, so it keeps
	// everything there.
	const uint32_t ctx_index = (uint32_t)(dev->draw_seq % dev->num_ctx), fs_index = ctx_index % dev->num_front_streams;
	cudaStream_t fs = dev->prof_on ? dev->stream : dev->front_streams[fs_index];
	DrawCtx *ctx = &dev->ctxs[ctx_index];
	dev->draw_seq++;

	// ---- arenas shared by all draws (used by the back half / tile kernels, which run one draw at a time)
	if(T > dev->tri_capacity || (debug && (need_slots > dev->dbg_tri_capacity || count > dev->dbg_vertex_capacity || !dev->dbg.infos))) {
		CUDA_TRY(quiesce(dev));
		if(T > dev->tri_capacity) {
			const uint32_t cap = T + T / 4;
			CUDA_TRY(regrow(&dev->tri_cov, (size_t)cap * MLV_TRI_COV_U4));
			CUDA_TRY(regrow(&dev->tri_shade, (size_t)cap * MLV_TRI_SHADE_U4));
			dev->tri_capacity = cap;
		}
		if(debug) {
			if(need_slots > dev->dbg_tri_capacity) {
				const uint32_t cap = need_slots + need_slots / 4;
				CUDA_TRY(regrow(&dev->dbg.tris, (size_t)cap));
				CUDA_TRY(regrow(&dev->dbg.attrs, (size_t)cap * 36));
				CUDA_TRY(regrow(&dev->dbg.slot_key, (size_t)cap));
				dev->dbg_tri_capacity = cap;
			}
			if(count > dev->dbg_vertex_capacity) {
				CUDA_TRY(regrow(&dev->dbg.vs_out, (size_t)count * 12));
				dev->dbg_vertex_capacity = count;
			}
			if(!dev->dbg.infos) CUDA_TRY(regrow(&dev->dbg.infos, (size_t)dev->pair_capacity));
			if(!dev->pair_tmp) CUDA_TRY(regrow(&dev->pair_tmp, (size_t)dev->pair_capacity));
		}
	}
	// ---- the draw context: what the front half hands to the back half
	uint32_t vcache_vertices = 0;
	const bool chunk_cull = dev->part.num_ranks > 1 && !debug && (dev->vs_id == MLV_VS_BASIC || dev->vs_id == MLV_VS_VERTEX_LIGHTING);
	// Post-transform vertex cache: worth it when the index buffer references each vertex of the buffer about twice or
	// more (the bound vertex buffer's size is the only vertex count a D3D11-style draw call has).
	if(indexed && !debug && !chunk_cull) { // (with chunk culling a rank touches ~1/N of the vertices; transforming all of them would cost more)
		const uint64_t vb_vertices = dev->vb->bytes / 32;
		if(vb_vertices > 0 && vb_vertices <= 0x7fffffffull && (uint64_t)count >= 2 * vb_vertices) vcache_vertices = (uint32_t)vb_vertices;
	}
	if(need_slots > ctx->slot_capacity || T > ctx->queue_capacity || ovf_cap > ctx->ovf_capacity || vcache_vertices > ctx->vcache_capacity || (chunk_cull && nblocks > ctx->chunk_capacity)) {
		// every context grows to the same size: draws rotate through the contexts, so each of them meets the largest draw sooner or later
		CUDA_TRY(quiesce(dev));
		for(uint32_t i = 0; i < dev->num_ctx; ++i) {
			DrawCtx *c = &dev->ctxs[i];
			if(need_slots > c->slot_capacity) {
				const uint32_t cap = need_slots + need_slots / 4;
				CUDA_TRY(regrow(&c->tri_bounds, (size_t)cap));
				CUDA_TRY(regrow(&c->big_queue, (size_t)cap));
				CUDA_TRY(regrow(&c->huge_queue, (size_t)cap));
				CUDA_TRY(regrow(&c->warp_sum, (size_t)cap / 32 + 1));
				c->slot_capacity = cap;
			}
			if(T > c->queue_capacity) {
				const uint32_t cap = T + T / 4;
				CUDA_TRY(regrow(&c->clip_queue, (size_t)cap));
				c->queue_capacity = cap;
			}
			if(ovf_cap > c->ovf_capacity) {
				const uint32_t cap = ovf_cap + ovf_cap / 4;
				CUDA_TRY(regrow(&c->ovf_cov, (size_t)cap * MLV_TRI_COV_U4));
				CUDA_TRY(regrow(&c->ovf_shade, (size_t)cap * MLV_TRI_SHADE_U4));
				c->ovf_capacity = cap;
			}
			if(vcache_vertices > c->vcache_capacity) {
				CUDA_TRY(regrow(&c->vcache, (size_t)vcache_vertices * 2));
				c->vcache_capacity = vcache_vertices;
			}
			if(chunk_cull && nblocks > c->chunk_capacity) {
				CUDA_TRY(regrow(&c->chunk_live, (size_t)nblocks));
				c->chunk_capacity = nblocks;
			}
		}
	}
	dev->last_direct_slots = T;
	dev->last_ctx = ctx;

	// ---- stream order. The front half reads the bound buffers and writes its draw context; it waits for (1) uploads of
	// those buffers still in flight, (2) the draw that used this context last (its k_tile re-arms the context), and -- the
	// first time after anything that is not ordered by (2) -- everything issued on the main stream so far.
	for(mlv_buffer *b : { dev->vb, indexed ? dev->ib : (mlv_buffer *)nullptr }) { // uploads still in flight on the copy stream
		if(!b) continue;
		if(dev->recording && std::find(dev->recording->buffers->begin(), dev->recording->buffers->end(), b) == dev->recording->buffers->end()) dev->recording->buffers->push_back(b);
		if(!b->ready_pending) continue;
		if(dev->recording) { // a capturing stream cannot wait for an event recorded outside the capture: the host waits instead
			CUDA_TRY(cudaEventSynchronize(b->ready));
		} else {
			CUDA_TRY(cudaStreamWaitEvent(dev->stream, b->ready, 0));
			for(uint32_t i = 0; i < dev->num_front_streams; ++i) CUDA_TRY(cudaStreamWaitEvent(dev->front_streams[i], b->ready, 0));
		}
		b->ready_pending = false;
	}
	if(!dev->prof_on) {
		if(dev->front_needs_sync[fs_index]) { // fork from the point of the main stream reset_front_order marked
			CUDA_TRY(cudaStreamWaitEvent(fs, dev->ev_main_sync, 0));
			dev->front_needs_sync[fs_index] = false;
			dev->front_touched[fs_index] = true;
		} else if(ctx->free_recorded) {
			CUDA_TRY(cudaStreamWaitEvent(fs, ctx->free_ev, 0));
		}
	}

	GeomParams gp;
	memset(&gp, 0, sizeof(gp));
	gp.ix.ib = indexed ? dev->ib->d : nullptr;
	gp.ix.start_index = start_index;
	gp.ix.base_vertex = base_vertex;
	gp.ix.index16 = dev->index_format == MLV_INDEX_U16 ? 1u : 0u;
	gp.vb = (const float4 *)dev->vb->d;
	gp.tri_count = T;
	gp.ovf_capacity = ovf_cap;
	memcpy(gp.cb, dev->cb[0], 192);
	if(dev->recording && dev->recording->draws < MLV_LIST_CONSTANT_DRAWS) { // replaceable per execution, see mlv_command_list
		gp.cbp = dev->recording->d_constants + (size_t)dev->recording->draws * 64;
		memcpy(dev->recording->h_constants + (size_t)dev->recording->draws * 64, dev->cb[0], 192);
		dev->recording->constants_dirty = true;
	}
	gp.vs_tex = tex_desc(dev->vs_srv[0]);
	gp.rsqrt_lut = dev->rsqrt_lut;
	{ // screen_from_ndc initialiser (main.c:825-830): double arithmetic rounded to f32 per entry
		const mlv_viewport &vp = dev->viewport;
		gp.vp_m00 = (float)((double)vp.width * 0.5);
		gp.vp_m03 = (float)((double)vp.width * 0.5 + (double)vp.top_left_x);
		gp.vp_m11 = (float)((double)(-vp.height) * 0.5);
		gp.vp_m13 = (float)((double)vp.height * 0.5 + (double)vp.top_left_y);
		gp.vp_m22 = vp.max_depth - vp.min_depth;
		gp.vp_m23 = vp.min_depth;
		gp.vp_w = (int)vp.width;
		gp.vp_h = (int)vp.height;
	}
	gp.wt = dev->wt;
	gp.ht = dev->ht;
	{ // v4f32_normalize((+-1,0,0,1)) (math.h:259-270): 1.0 / (f32)sqrt(dot), double divide rounded to f32
		const float len = (float)sqrt((double)2.0f);
		gp.clip_k = (float)(1.0 / (double)len);
	}
	gp.part = dev->part;
	gp.tri_cov = dev->tri_cov;
	gp.tri_shade = dev->tri_shade;
	gp.tri_bounds = ctx->tri_bounds;
	gp.ovf_cov = ctx->ovf_cov;
	gp.ovf_shade = ctx->ovf_shade;
	gp.dctr = ctx->dctr;
	gp.clip_queue = ctx->clip_queue;
	gp.big_queue = ctx->big_queue;
	gp.huge_queue = ctx->huge_queue;
	gp.bin_count = dev->bin_count;
	gp.touch_bits = ctx->touch_bits;
	gp.warp_sum = ctx->warp_sum;
	gp.tile_min = dev->tile_min;
	gp.keep_all = debug;
	if(debug) gp.dbg = dev->dbg;
	gp.ctr = dev->ctr;
	gp.stat_stripes = ctx->stat_stripes;
	gp.index_count = count;
	gp.draw_ordinal = dev->recording ? dev->recording->draws : 0u;
	if(vcache_vertices) gp.vcache = ctx->vcache;

	// Sort-first chunk culling (multi-GPU): object-space chunk bounds cached with the buffer that defines the triangle list.
	if(chunk_cull) {
		mlv_buffer *owner = indexed ? dev->ib : dev->vb;
		const bool valid = owner->chunk_bounds && owner->chunk_count == nblocks && owner->chunk_indexed == (indexed ? 1 : 0) && owner->chunk_self_version == owner->version &&
		                   owner->chunk_vb_uid == dev->vb->uid && owner->chunk_vb_version == dev->vb->version && owner->chunk_start_index == start_index &&
		                   owner->chunk_base_vertex == base_vertex && owner->chunk_index16 == gp.ix.index16;
		if(!valid) {
			if(nblocks > owner->chunk_capacity) {
				CUDA_TRY(quiesce(dev));
				CUDA_TRY(regrow(&owner->chunk_bounds, (size_t)nblocks * 2));
				owner->chunk_capacity = nblocks;
			}
			prof_pre(dev, MLV_STAGE_GEOMETRY);
			if(indexed) launch_pdl(k_chunk_bounds<true>, nblocks, MLV_GEOM_THREADS, fs, gp.ix, gp.vb, T, owner->chunk_bounds);
			else launch_pdl(k_chunk_bounds<false>, nblocks, MLV_GEOM_THREADS, fs, gp.ix, gp.vb, T, owner->chunk_bounds);
			if(int rc = check_launch(dev, "k_chunk_bounds")) return rc;
			// the bounds are cached with the buffer: later draws (on other front streams) use them too
			if(!dev->prof_on) {
				CUDA_TRY(cudaEventRecord(ctx->front_done, fs));
				for(uint32_t i = 0; i < dev->num_front_streams; ++i) {
					if(i == fs_index) continue;
					CUDA_TRY(cudaStreamWaitEvent(dev->front_streams[i], ctx->front_done, 0));
					if(dev->recording) dev->front_touched[i] = true;
				}
			}
			owner->chunk_count = nblocks;
			owner->chunk_indexed = indexed ? 1 : 0;
			owner->chunk_self_version = owner->version;
			owner->chunk_vb_uid = dev->vb->uid;
			owner->chunk_vb_version = dev->vb->version;
			owner->chunk_start_index = start_index;
			owner->chunk_base_vertex = base_vertex;
			owner->chunk_index16 = gp.ix.index16;
		}
		gp.chunk_bounds = owner->chunk_bounds;
		gp.chunk_live = ctx->chunk_live;
	}

	// ---- front half (its own stream): k_vertex, k_front, k_front_clip
	nvtxRangePushA("input_assambler_stage+vertex_shader_stage+primitive_assembly_stage"); // (sic, main.c:663)
	g_launch_priority_set = dev->knob_explicit_priority != 0;
	g_launch_priority = dev->prio_front;
	switch(dev->vs_id) {
		case MLV_VS_PASSTHROUGH: launch_front<0>(dev, fs, gp, nblocks, indexed, vcache_vertices); break;
		case MLV_VS_BASIC: launch_front<1>(dev, fs, gp, nblocks, indexed, vcache_vertices); break;
		case MLV_VS_VERTEX_LIGHTING: launch_front<2>(dev, fs, gp, nblocks, indexed, vcache_vertices); break;
		default: launch_front<3>(dev, fs, gp, nblocks, indexed, vcache_vertices); break;
	}
	nvtxRangePop();
	if(int rc = check_launch_only(dev, "front half")) return rc;
	if(!dev->prof_on) {
		CUDA_TRY(cudaEventRecord(ctx->front_done, fs));
		CUDA_TRY(cudaStreamWaitEvent(dev->stream, ctx->front_done, 0));
	}

	// ---- back half (main stream: ordered after the previous draw's k_tile through the tile minima)
	NvtxScope nvtx_tail("binner+rasterizer_stage+pixel_shader_stage");
	g_launch_priority = dev->prio_chain;
	const uint32_t back_blocks = (need_slots + MLV_GEOM_THREADS - 1) / MLV_GEOM_THREADS;
	switch(dev->vs_id) {
		case MLV_VS_PASSTHROUGH: launch_back<0>(dev, gp, back_blocks, indexed, vcache_vertices != 0); break;
		case MLV_VS_BASIC: launch_back<1>(dev, gp, back_blocks, indexed, vcache_vertices != 0); break;
		case MLV_VS_VERTEX_LIGHTING: launch_back<2>(dev, gp, back_blocks, indexed, vcache_vertices != 0); break;
		default: launch_back<3>(dev, gp, back_blocks, indexed, vcache_vertices != 0); break;
	}
	if(int rc = check_launch(dev, "k_back")) return rc;

	TailParams tp;
	memset(&tp, 0, sizeof(tp));
	tp.tri_bounds = ctx->tri_bounds;
	tp.dctr = ctx->dctr;
	tp.big_queue = ctx->big_queue;
	tp.huge_queue = ctx->huge_queue;
	tp.chunk_live = gp.chunk_live;
	tp.tile_min = dev->tile_min;
	tp.bin_count = dev->bin_count;
	tp.touch_bits = ctx->touch_bits;
	tp.warp_sum = reinterpret_cast<const uint32_t *>(ctx->warp_sum);
	tp.bin_offset = dev->bin_offset;
	tp.pair_ids = dev->pair_ids;
	tp.pair_tmp = dev->pair_tmp;
	tp.cbins = dev->cbins;
	tp.state_sum = dev->scan_state;
	tp.state_nz = dev->scan_state + dev->scan_blocks;
	tp.scan_blocks = dev->scan_blocks;
	tp.ctr = dev->ctr;
	tp.stat_stripes = ctx->stat_stripes;
	tp.direct_slots = T;
	tp.ovf_capacity = ovf_cap;
	tp.pair_capacity = (uint32_t)dev->pair_capacity;
	tp.num_bins = dev->num_bins;
	tp.bin_begin = dev->bin_begin;
	tp.bin_end = dev->bin_end;
	tp.wt = dev->wt;
	tp.ht = dev->ht;
	tp.part = dev->part;
	tp.keep_all = debug;
	tp.sort_lists = debug;
	tp.tri_cov = dev->tri_cov;
	tp.tri_shade = dev->tri_shade;
	tp.ovf_cov = ctx->ovf_cov;
	tp.ovf_shade = ctx->ovf_shade;
	tp.fb = dev->fb;
	tp.ps_tex = tex_desc(dev->ps_srv[0]);
	tp.rsqrt_lut = dev->rsqrt_lut;
	if(debug) tp.dbg = dev->dbg;
	tp.index_count = count;
	tp.key_bits = 3u;
	while(tp.key_bits < 32u && (T >> (tp.key_bits - 3u)) != 0u) tp.key_bits++;
	if(dev->timeline_on && dev->timeline_used + 3u <= MLV_TIMELINE_CAPACITY) { // three consecutive slots
		tp.timeline = timeline_slot(dev, MLV_STAGE_BIN_SCAN);
		timeline_slot(dev, MLV_STAGE_BIN_FILL);
		timeline_slot(dev, MLV_STAGE_TILE);
	}
	dev->timeline_draw++;
	prof_pre(dev, MLV_STAGE_BIN_SCAN);
	launch_pdl(k_bin_scan, dev->scan_blocks, MLV_SCAN_THREADS, dev->stream, tp);
	if(int rc = check_launch(dev, "k_bin_scan")) return rc;
	uint32_t fill_blocks = (need_slots + 255u) / 256u;
	if(fill_blocks > dev->sm_count * 6u) fill_blocks = dev->sm_count * 6u; // persistent: what is resident at once (40 registers)
	prof_pre(dev, MLV_STAGE_BIN_FILL);
	launch_pdl(k_fill, fill_blocks, 256, dev->stream, tp);
	if(int rc = check_launch(dev, "k_fill")) return rc;
	uint32_t tile_blocks = (dev->num_bins + 7u) / 8u;
	{ // persistent: what is resident at once (__launch_bounds__(256, 4)), or fewer to leave room for the front halves of later draws
		static const uint32_t per_sm = getenv("MLV_TILE_CTAS_PER_SM") ? (uint32_t)atoi(getenv("MLV_TILE_CTAS_PER_SM")) : 4u;
		if(tile_blocks > dev->sm_count * per_sm) tile_blocks = dev->sm_count * per_sm;
	}
	prof_pre(dev, MLV_STAGE_TILE);
	g_pdl = dev->knob_no_pdl_tile ? 0 : 1;
	switch(dev->ps_id) {
		case MLV_PS_PASSTHROUGH: launch_pdl(k_tile<0>, tile_blocks, MLV_TILE_THREADS, dev->stream, tp); break;
		case MLV_PS_BASIC: launch_pdl(k_tile<1>, tile_blocks, MLV_TILE_THREADS, dev->stream, tp); break;
		case MLV_PS_BASIC_TRILINEAR: launch_pdl(k_tile<MLV_PS_ID_BASIC_TRILINEAR>, tile_blocks, MLV_TILE_THREADS, dev->stream, tp); break;
		default: launch_pdl(k_tile<2>, tile_blocks, MLV_TILE_THREADS, dev->stream, tp); break;
	}
	g_pdl = 1;
	g_launch_priority_set = false;
	if(int rc = check_launch(dev, "k_tile")) return rc;
	if(!dev->prof_on) { // the context is free again once this draw's k_tile has re-armed it
		CUDA_TRY(cudaEventRecord(ctx->free_ev, dev->stream));
		ctx->free_recorded = true;
	}
	if(dev->recording) {
		dev->recording->draws++;
		return MLV_OK; // (mlv_execute_command_list records ev_last_draw after the whole list)
	}
	CUDA_TRY(cudaEventRecord(dev->ev_last_draw, dev->stream));
	dev->last_draw_recorded = true;
	return MLV_OK;
}

extern "C" {

int mlv_draw_indexed(mlv_device *dev, uint32_t index_count) {
	GROUP_EACH(dev, draw_common(c, index_count, true));
	return draw_common(dev, index_count, true);
}
int mlv_draw(mlv_device *dev, uint32_t vertex_count) {
	GROUP_EACH(dev, draw_common(c, vertex_count, false));
	return draw_common(dev, vertex_count, false);
}
int mlv_draw_indexed_ex(mlv_device *dev, uint32_t index_count, uint32_t start_index_location, int32_t base_vertex_location) {
	GROUP_EACH(dev, draw_common(c, index_count, true, start_index_location, base_vertex_location));
	return draw_common(dev, index_count, true, start_index_location, base_vertex_location);
}

// ---- command lists ---------------------------------------------------------------------------------
// The D3D11 deferred-context pattern (ID3D11DeviceContext::FinishCommandList / ExecuteCommandList) on top of CUDA graphs:
// between mlv_begin_command_list and mlv_finish_command_list the clears, draws and resolves are CAPTURED from the device
// stream instead of executed -- the kernels keep every per-draw quantity on the device (counts, queues, scan epoch), so
// nothing a replay needs comes from the host -- and mlv_execute_command_list replays the frame with ONE graph launch
// (~10 us of host time instead of ~30 us per draw).

static void free_command_list(mlv_command_list *list) {
	if(!list) return;
	if(list->exec) cudaGraphExecDestroy(list->exec);
	if(list->graph) cudaGraphDestroy(list->graph);
	if(list->owner && list->owner->lists_alive) list->owner->lists_alive--;
	if(list->geom_nodes)
		for(GeomNode *n : *list->geom_nodes) delete n;
	if(list->d_constants) cudaFree(list->d_constants);
	free(list->h_constants);
	delete list->buffers;
	delete list->geom_funcs;
	delete list->vertex_funcs;
	delete list->geom_nodes;
	delete list;
}

int mlv_begin_command_list(mlv_device *dev) {
	GROUP_EACH(dev, mlv_begin_command_list(c));
	if(int rc = use_device(dev)) return rc;
	if(dev->recording) return fail(MLV_ERR_STATE, "a command list is already being recorded");
	if(dev->prof_on) return fail(MLV_ERR_STATE, "call mlv_profile_end before recording a command list");
	if(int rc = flush_clears(dev)) return rc; // clears issued before the recording belong to immediate mode
	mlv_command_list *list = new(std::nothrow) mlv_command_list();
	if(!list) return fail(MLV_ERR_OUT_OF_MEMORY, "host allocation failed");
	memset(list, 0, sizeof(*list));
	list->owner = dev;
	list->buffers = new std::vector<mlv_buffer *>();
	list->geom_funcs = new std::vector<const void *>();
	list->vertex_funcs = new std::vector<const void *>();
	list->geom_nodes = new std::vector<GeomNode *>();
	dev->lists_alive++; // from now on a replaced arena is kept until the device is destroyed: recorded nodes may address it
	list->fb_sel = dev->fb_sel;
	list->h_constants = (float *)calloc((size_t)MLV_LIST_CONSTANT_DRAWS * 64, sizeof(float));
	cudaError_t e = list->h_constants ? cudaMalloc(&list->d_constants, (size_t)MLV_LIST_CONSTANT_DRAWS * 64 * sizeof(float)) : cudaErrorMemoryAllocation;
	if(e != cudaSuccess) {
		free_command_list(list);
		return fail(MLV_ERR_OUT_OF_MEMORY, "command-list constants arena: %s", cudaGetErrorString(e));
	}
	e = cudaStreamBeginCapture(dev->stream, cudaStreamCaptureModeRelaxed);
	if(e != cudaSuccess) {
		free_command_list(list);
		return fail(MLV_ERR_CUDA, "cudaStreamBeginCapture: %s", cudaGetErrorString(e));
	}
	dev->recording = list;
	// events recorded outside the capture cannot be waited for inside it: the side stream forks from the main stream again
	reset_front_order(dev);
	return MLV_OK;
}

int mlv_finish_command_list(mlv_device *dev, mlv_command_list **out_list) {
	if(dev && dev->children) {
		if(!out_list) return fail(MLV_ERR_INVALID_ARGUMENT, "null argument");
		mlv_command_list *g = new mlv_command_list();
		memset(g, 0, sizeof(*g));
		g->owner = dev;
		g->children = new std::vector<mlv_command_list *>();
		int first_rc = MLV_OK;
		for(mlv_device *c : *dev->children) { // every rank leaves the recording state, whatever happens to one of them
			mlv_command_list *l = nullptr;
			const int rc = mlv_finish_command_list(c, &l);
			if(rc && !first_rc) first_rc = rc;
			g->children->push_back(l);
		}
		if(first_rc) {
			mlv_release_command_list(dev, g);
			*out_list = nullptr;
			return first_rc;
		}
		g->draws = (*g->children)[0]->draws;
		for(mlv_command_list *l : *g->children) g->launches += l->launches;
		*out_list = g;
		return MLV_OK;
	}
	if(int rc = use_device(dev)) return rc;
	if(!dev->recording) return fail(MLV_ERR_STATE, "no command list is being recorded");
	if(!out_list) return fail(MLV_ERR_INVALID_ARGUMENT, "null argument");
	*out_list = nullptr;
	int rc = flush_clears(dev); // a clear recorded after the last draw
	for(uint32_t i = 0; i < dev->num_front_streams; ++i) { // every stream that joined the capture re-joins the main stream
		if(!dev->front_touched[i]) continue;
		cudaEventRecord(dev->ev_front_join[i], dev->front_streams[i]);
		cudaStreamWaitEvent(dev->stream, dev->ev_front_join[i], 0);
	}
	mlv_command_list *list = dev->recording;
	const uint64_t launches_before = list->launches; // (set by begin: dev->launches at that time)
	(void)launches_before;
	cudaError_t e = cudaStreamEndCapture(dev->stream, &list->graph);
	dev->recording = nullptr;
	reset_front_order(dev); // the captured events mean nothing outside the graph
	if(rc != MLV_OK || e != cudaSuccess || !list->graph) {
		free_command_list(list);
		if(rc != MLV_OK) return rc;
		return fail(MLV_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
	}
	// (per-node priorities: the chain kernels were recorded with the highest launch priority, the front halves with the lowest)
	e = cudaGraphInstantiateWithFlags(&list->exec, list->graph, dev->knob_explicit_priority ? cudaGraphInstantiateFlagUseNodePriority : 0);
	if(e != cudaSuccess) {
		free_command_list(list);
		return fail(MLV_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
	}
	// the geometry kernel nodes, so that mlv_command_list_set_constants can replace their constant buffer
	size_t n = 0;
	cudaGraphGetNodes(list->graph, nullptr, &n);
	std::vector<cudaGraphNode_t> nodes(n);
	if(n) cudaGraphGetNodes(list->graph, nodes.data(), &n);
	for(size_t i = 0; i < n; ++i) {
		cudaGraphNodeType type;
		if(cudaGraphNodeGetType(nodes[i], &type) != cudaSuccess || type != cudaGraphNodeTypeKernel) continue;
		cudaKernelNodeParams kp;
		if(cudaGraphKernelNodeGetParams(nodes[i], &kp) != cudaSuccess || !kp.kernelParams) continue;
		const bool is_vertex = std::find(list->vertex_funcs->begin(), list->vertex_funcs->end(), (const void *)kp.func) != list->vertex_funcs->end();
		const bool is_geom = std::find(list->geom_funcs->begin(), list->geom_funcs->end(), (const void *)kp.func) != list->geom_funcs->end();
		if(!is_vertex && !is_geom) continue;
		if(((const GeomParams *)kp.kernelParams[0])->cbp) continue; // constants in the list's device arena
		GeomNode *g = new GeomNode();
		g->node = nodes[i];
		g->params = kp;
		memcpy(&g->gp, kp.kernelParams[0], sizeof(GeomParams));
		g->extra = is_vertex ? *(const uint32_t *)kp.kernelParams[1] : 0u;
		g->argv[0] = &g->gp;
		g->argv[1] = &g->extra;
		g->params.kernelParams = g->argv;
		g->params.extra = nullptr;
		list->geom_nodes->push_back(g);
	}
	cudaGetLastError();
	list->last_index_count = dev->last_index_count;
	list->last_direct_slots = dev->last_direct_slots;
	*out_list = list;
	return MLV_OK;
}

int mlv_execute_command_list(mlv_device *dev, mlv_command_list *list) {
	GROUP_EACH(dev, (list && list->children && list->owner == dev) ? mlv_execute_command_list(c, (*list->children)[gi]) : fail(MLV_ERR_INVALID_ARGUMENT, "command list does not belong to this device"));
	NvtxScope nvtx("render (recorded command list)");
	if(int rc = use_device(dev)) return rc;
	if(int rc = immediate_only(dev, "mlv_execute_command_list")) return rc;
	if(!list || list->owner != dev) return fail(MLV_ERR_INVALID_ARGUMENT, "command list does not belong to this device");
	// The recorded kernels address ONE framebuffer of the pair. A list that draws over what is there (no full clear of its own)
	// needs those contents in ITS framebuffer: a full clear still pending from immediate mode is simply issued there;
	// otherwise the contents live where immediate mode (or another list) left them.
	if(dev->fb_sel != list->fb_sel) {
		if(!list->opens_with_full_clear && !(dev->pend_color && dev->pend_depth))
			return fail(MLV_ERR_STATE, "the command list draws over framebuffer contents that live in the other tiled framebuffer of the pair (record it with its clears, or re-record it)");
		dev->fb_sel = list->fb_sel;
		dev->fb = dev->fb_pair[dev->fb_sel];
	}
	// what the recorded kernels could not wait for inside the capture: the framebuffer they address (an asynchronous
	// exchange or present may still read it), a read-back of the resolved image in flight, uploads of the buffers they bind
	if(int rc = wait_framebuffer_readers(dev)) return rc;
	if(int rc = flush_clears(dev)) return rc; // clears issued in immediate mode come first (nothing reads this framebuffer any more: no flip)
	if(list->has_resolve && dev->readback_in_flight) CUDA_TRY(cudaStreamWaitEvent(dev->stream, dev->ev_readback_done, 0));
	for(mlv_buffer *b : *list->buffers) {
		if(!b->ready_pending) continue;
		CUDA_TRY(cudaStreamWaitEvent(dev->stream, b->ready, 0));
		b->ready_pending = false;
	}
	if(list->constants_dirty) { // (pageable source: staged by the driver before the call returns, the shadow may be rewritten at once)
		const uint32_t n = list->draws < MLV_LIST_CONSTANT_DRAWS ? list->draws : MLV_LIST_CONSTANT_DRAWS;
		CUDA_TRY(cudaMemcpyAsync(list->d_constants, list->h_constants, (size_t)n * 64 * sizeof(float), cudaMemcpyHostToDevice, dev->stream));
		list->constants_dirty = false;
	}
	CUDA_TRY(cudaGraphLaunch(list->exec, dev->stream));
	dev->launches += list->launches;
	dev->last_index_count = list->last_index_count;
	dev->last_direct_slots = list->last_direct_slots;
	if(list->has_resolve) dev->present_color = dev->resolved_color;
	CUDA_TRY(cudaEventRecord(dev->ev_last_draw, dev->stream));
	dev->last_draw_recorded = true;
	reset_front_order(dev); // the graph ran the front halves on branches of its own; the front stream re-joins the main stream first
	return MLV_OK;
}

int mlv_command_list_set_constants(mlv_device *dev, mlv_command_list *list, uint32_t draw_index, const void *data, size_t bytes) {
	GROUP_EACH(dev, (list && list->children && list->owner == dev) ? mlv_command_list_set_constants(c, (*list->children)[gi], draw_index, data, bytes) : fail(MLV_ERR_INVALID_ARGUMENT, "bad command-list constants"));
	if(int rc = use_device(dev)) return rc;
	if(!list || list->owner != dev || !data || bytes > sizeof(((GeomParams *)0)->cb)) return fail(MLV_ERR_INVALID_ARGUMENT, "bad command-list constants");
	if(draw_index != MLV_ALL_DRAWS && draw_index >= list->draws) return fail(MLV_ERR_INVALID_ARGUMENT, "draw %u out of range (%u draws recorded)", draw_index, list->draws);
	uint32_t touched = 0;
	for(uint32_t d = 0; d < list->draws && d < MLV_LIST_CONSTANT_DRAWS; ++d) {
		if(draw_index != MLV_ALL_DRAWS && d != draw_index) continue;
		memcpy(list->h_constants + (size_t)d * 64, data, bytes);
		list->constants_dirty = true;
		++touched;
	}
	for(GeomNode *g : *list->geom_nodes) { // draws beyond the arena
		if(draw_index != MLV_ALL_DRAWS && g->gp.draw_ordinal != draw_index) continue;
		memcpy(g->gp.cb, data, bytes);
		cudaError_t e = cudaGraphExecKernelNodeSetParams(list->exec, g->node, &g->params);
		if(e != cudaSuccess) return fail(MLV_ERR_CUDA, "cudaGraphExecKernelNodeSetParams: %s", cudaGetErrorString(e));
		++touched;
	}
	if(!touched && list->draws) return fail(MLV_ERR_STATE, "the recorded list exposes no geometry kernel nodes; record it again with the new constants");
	return MLV_OK;
}

int mlv_command_list_info(const mlv_command_list *list, uint32_t *out_draws, uint64_t *out_kernel_launches) {
	if(!list) return fail(MLV_ERR_INVALID_ARGUMENT, "null argument");
	if(out_draws) *out_draws = list->draws;
	if(out_kernel_launches) *out_kernel_launches = list->launches;
	return MLV_OK;
}

void mlv_release_command_list(mlv_device *dev, mlv_command_list *list) {
	if(dev && dev->children) {
		if(!list) return;
		if(list->children) {
			for(size_t gi = 0; gi < list->children->size(); ++gi)
				if((*list->children)[gi]) mlv_release_command_list((*dev->children)[gi], (*list->children)[gi]);
			delete list->children;
		}
		delete list;
		return;
	}
	if(!dev || !list) return;
	cudaSetDevice(dev->cuda_dev);
	cudaStreamSynchronize(dev->stream); // an execution may still be in flight
	free_command_list(list);
}

// ---- results -------------------------------------------------------------------------------------

static int check_flags(mlv_device *dev, const Counters &c) {
	if(c.error_flags & MLV_FLAG_TRI_OVERFLOW) return fail(MLV_ERR_CAPACITY, "a draw clipped its triangles into more than max(3T,512) fan triangles (the reference bounds its output by T + max(2T,512), main.c:739-740); that draw was skipped");
	if(c.error_flags & MLV_FLAG_COMPOSITE_TIMEOUT) return fail(MLV_ERR_STATE, "peer-memory compositing: a rank's stripes did not arrive within 10 s");
	if(c.error_flags & MLV_FLAG_PAIR_OVERFLOW)
		return fail(MLV_ERR_CAPACITY, "a draw produced more (triangle,tile) pairs than max_pairs_per_draw = %llu; that draw was skipped", (unsigned long long)dev->pair_capacity);
	return MLV_OK;
}

int mlv_resolve(mlv_device *dev) {
	if(int rc = use_device(dev)) return rc;
	if(int rc = flush_clears(dev)) return rc;
	if(dev->recording) dev->recording->has_resolve = true; // (mlv_execute_command_list waits for a read-back in flight)
	else if(dev->readback_in_flight) CUDA_TRY(cudaStreamWaitEvent(dev->stream, dev->ev_readback_done, 0));
	const uint32_t items = (uint32_t)(dev->W / 8) * (uint32_t)dev->H; // 8 pixels (4 wide, rows y and y+4) per thread
	prof_pre(dev, MLV_STAGE_RESOLVE);
	launch_pdl(k_resolve, (items + 255) / 256, 256, dev->stream, dev->fb, dev->resolved_color, dev->resolved_depth, dev->W, dev->H);
	dev->present_color = dev->resolved_color;
	return check_launch(dev, "k_resolve");
}

void *mlv_resolved_color_device_ptr(mlv_device *dev) { return dev ? dev->present_color : nullptr; }
void *mlv_resolved_depth_device_ptr(mlv_device *dev) { return dev ? dev->resolved_depth : nullptr; }

int mlv_present_readback_async(mlv_device *dev, uint32_t *colors, float *depths) {
	if(dev && dev->children) return group_present(dev, colors, depths, false);
	NvtxScope nvtx("present");
	if(int rc = immediate_only(dev, "mlv_present_readback_async")) return rc;
	if(int rc = use_device(dev)) return rc;
	if(int rc = flush_clears(dev)) return rc;
	// The resolve runs on the READ-BACK stream, behind the copies of the previous present (the resolved images are reused:
	// stream order) and beside whatever the main stream does next: a frame that opens with a full clear goes to the other
	// framebuffer of the pair, so the next frame's kernels start while this one is still being resolved and copied.
	const uint32_t items = (uint32_t)(dev->W / 8) * (uint32_t)dev->H;
	const bool profiled = dev->prof_on; // (per-stage profiling brackets launches on the main stream)
	cudaStream_t rs = profiled ? dev->stream : dev->readback_stream;
	if(profiled) {
		if(dev->readback_in_flight) CUDA_TRY(cudaStreamWaitEvent(dev->stream, dev->ev_readback_done, 0));
		prof_pre(dev, MLV_STAGE_RESOLVE);
	} else {
		CUDA_TRY(cudaEventRecord(dev->ev_present_src, dev->stream));
		CUDA_TRY(cudaStreamWaitEvent(rs, dev->ev_present_src, 0));
	}
	launch_pdl(k_resolve, (items + 255) / 256, 256, rs, dev->fb, dev->resolved_color, depths ? dev->resolved_depth : nullptr, dev->W, dev->H);
	dev->present_color = dev->resolved_color;
	if(int rc = check_launch(dev, "k_resolve")) return rc;
	if(profiled) {
		CUDA_TRY(cudaEventRecord(dev->ev_resolved, dev->stream));
		CUDA_TRY(cudaStreamWaitEvent(dev->readback_stream, dev->ev_resolved, 0));
	} else {
		CUDA_TRY(cudaEventRecord(dev->ev_fb_read[dev->fb_sel], rs));
		dev->fb_reading[dev->fb_sel] = true;
	}
	const size_t bytes = (size_t)dev->W * dev->H * 4;
	if(colors) CUDA_TRY(cudaMemcpyAsync(colors, dev->resolved_color, bytes, cudaMemcpyDeviceToHost, dev->readback_stream));
	if(depths) CUDA_TRY(cudaMemcpyAsync(depths, dev->resolved_depth, bytes, cudaMemcpyDeviceToHost, dev->readback_stream));
	CUDA_TRY(cudaEventRecord(dev->ev_readback_done, dev->readback_stream));
	dev->readback_in_flight = true;
	return MLV_OK;
}

int mlv_present_wait(mlv_device *dev) {
	if(dev && dev->children) return group_present_wait(dev);
	if(int rc = immediate_only(dev, "mlv_present_wait")) return rc;
	if(int rc = use_device(dev)) return rc;
	if(!dev->readback_in_flight) return MLV_OK;
	CUDA_TRY(cudaEventSynchronize(dev->ev_readback_done));
	dev->readback_in_flight = false;
	return MLV_OK;
}

int mlv_present_readback(mlv_device *dev, uint32_t *colors, float *depths) {
	if(dev && dev->children) return group_present(dev, colors, depths, true);
	NvtxScope nvtx("present");
	if(int rc = immediate_only(dev, "mlv_present_readback")) return rc;
	if(int rc = use_device(dev)) return rc;
	if(int rc = flush_clears(dev)) return rc;
	if(dev->readback_in_flight) CUDA_TRY(cudaStreamWaitEvent(dev->stream, dev->ev_readback_done, 0));
	const uint32_t items = (uint32_t)(dev->W / 8) * (uint32_t)dev->H;
	prof_pre(dev, MLV_STAGE_RESOLVE);
	launch_pdl(k_resolve, (items + 255) / 256, 256, dev->stream, dev->fb, dev->resolved_color, depths ? dev->resolved_depth : nullptr, dev->W, dev->H);
	dev->present_color = dev->resolved_color;
	if(int rc = check_launch(dev, "k_resolve")) return rc;
	const size_t bytes = (size_t)dev->W * dev->H * 4;
	if(colors) CUDA_TRY(cudaMemcpyAsync(colors, dev->resolved_color, bytes, cudaMemcpyDeviceToHost, dev->stream));
	if(depths) CUDA_TRY(cudaMemcpyAsync(depths, dev->resolved_depth, bytes, cudaMemcpyDeviceToHost, dev->stream));
	Counters c;
	CUDA_TRY(cudaMemcpyAsync(&c, dev->ctr, sizeof(c), cudaMemcpyDeviceToHost, dev->stream));
	CUDA_TRY(cudaStreamSynchronize(dev->stream));
	return check_flags(dev, c);
}

int mlv_get_stats(mlv_device *dev, mlv_stats *out) {
	if(dev && dev->children) { // every rank counts its share (a triangle: the owner of its first tile row): the sum is the reference's Stats
		if(!out) return fail(MLV_ERR_INVALID_ARGUMENT, "null argument");
		memset(out, 0, sizeof(*out));
		for(size_t gi = 0; gi < dev->children->size(); ++gi) {
			mlv_stats s;
			if(int rc = mlv_get_stats((*dev->children)[gi], &s)) return rc;
			if(gi == 0) out->vertex_count = s.vertex_count, out->input_triangle_count = s.input_triangle_count; // (replicated inputs)
			out->assembled_triangle_count += s.assembled_triangle_count;
			out->active_bin_count += s.active_bin_count;
			out->total_triangle_count_in_bins += s.total_triangle_count_in_bins;
		}
		return MLV_OK;
	}
	if(int rc = immediate_only(dev, "mlv_get_stats")) return rc;
	if(int rc = use_device(dev)) return rc;
	if(!out) return fail(MLV_ERR_INVALID_ARGUMENT, "null argument");
	Counters c;
	CUDA_TRY(cudaMemcpyAsync(&c, dev->ctr, sizeof(c), cudaMemcpyDeviceToHost, dev->stream));
	CUDA_TRY(cudaStreamSynchronize(dev->stream));
	*out = c.stats;
	return check_flags(dev, c);
}

int mlv_get_work_counters(mlv_device *dev, mlv_work_counters *out) {
	if(dev && dev->children) {
		if(!out) return fail(MLV_ERR_INVALID_ARGUMENT, "null argument");
		memset(out, 0, sizeof(*out));
		for(mlv_device *c : *dev->children) {
			mlv_work_counters w;
			if(int rc = mlv_get_work_counters(c, &w)) return rc;
			out->records_written += w.records_written, out->pairs_listed += w.pairs_listed, out->tiles_visited += w.tiles_visited;
		}
		return MLV_OK;
	}
	if(int rc = immediate_only(dev, "mlv_get_work_counters")) return rc;
	if(int rc = use_device(dev)) return rc;
	if(!out) return fail(MLV_ERR_INVALID_ARGUMENT, "null argument");
	Counters c;
	CUDA_TRY(cudaMemcpyAsync(&c, dev->ctr, sizeof(c), cudaMemcpyDeviceToHost, dev->stream));
	CUDA_TRY(cudaStreamSynchronize(dev->stream));
	*out = c.work;
	return check_flags(dev, c);
}

int mlv_reset_stats(mlv_device *dev) {
	GROUP_EACH(dev, mlv_reset_stats(c));
	if(int rc = use_device(dev)) return rc;
	CUDA_TRY(cudaMemsetAsync((char *)dev->ctr + offsetof(Counters, work), 0, sizeof(mlv_work_counters), dev->stream));
	CUDA_TRY(cudaMemsetAsync((char *)dev->ctr + offsetof(Counters, stats), 0, sizeof(mlv_stats), dev->stream));
	CUDA_TRY(cudaMemsetAsync((char *)dev->ctr + offsetof(Counters, error_flags), 0, sizeof(uint32_t), dev->stream));
	return MLV_OK;
}

int mlv_composite_layout(mlv_device *dev, void **out_gather_device_ptr, size_t *out_chunk_bytes) {
	if(!dev) return fail(MLV_ERR_INVALID_ARGUMENT, "null device");
	if(dev->part.num_ranks <= 1) return fail(MLV_ERR_STATE, "device was created with a single rank");
	if(out_gather_device_ptr) *out_gather_device_ptr = dev->gather;
	if(out_chunk_bytes) *out_chunk_bytes = dev->chunk_bytes;
	return MLV_OK;
}

int mlv_composite_pack(mlv_device *dev) {
	if(int rc = use_device(dev)) return rc;
	if(dev->part.num_ranks <= 1) return fail(MLV_ERR_STATE, "device was created with a single rank");
	if(int rc = flush_clears(dev)) return rc;
	const uint32_t items = (uint32_t)(dev->W / 8) * (uint32_t)dev->H;
	uint4 *chunk = reinterpret_cast<uint4 *>(reinterpret_cast<char *>(dev->gather) + dev->chunk_bytes * (size_t)dev->part.rank);
	prof_pre(dev, MLV_STAGE_COMPOSITE);
	launch_pdl(k_composite_pack, (items + 255) / 256, 256, dev->stream, dev->fb, chunk, dev->W, dev->H, dev->part);
	return check_launch(dev, "k_composite_pack");
}

// COMPOSITE IN HOST MEMORY. A frame whose destination is the host does not need the ranks to exchange anything: every rank
// packs the stripes it owns (k_composite_pack, row-major in stripe order) and copies them straight to their rows of ONE
// host frame -- pinned memory shared by the ranks (one process: any pinned buffer; several processes: a shared mapping
// registered with cudaHostRegister in each) -- over its OWN PCIe link. N links carry the frame in parallel and NVLink is not
// involved; the frame is complete when every rank's mlv_present_wait has returned. Two packed chunks alternate, so the copy
// of frame f is still in flight while frame f+1 is packed.
int mlv_present_owned_rows_async(mlv_device *dev, uint32_t *frame_colors) {
	if(dev && dev->children) { // a device group: every GPU delivers its band
		if(!frame_colors) return fail(MLV_ERR_INVALID_ARGUMENT, "null argument");
		for(mlv_device *c : *dev->children)
			if(int rc = mlv_present_owned_rows_async(c, frame_colors)) return rc;
		return MLV_OK;
	}
	if(int rc = immediate_only(dev, "mlv_present_owned_rows_async")) return rc;
	if(int rc = use_device(dev)) return rc;
	if(!frame_colors) return fail(MLV_ERR_INVALID_ARGUMENT, "null argument");
	if(dev->part.num_ranks <= 1) return mlv_present_readback_async(dev, frame_colors, nullptr);
	if(int rc = flush_clears(dev)) return rc;
	const int n = dev->part.num_ranks, me = dev->part.rank, sh = dev->part.stripe_h;
	dev->owned_parity ^= 1u;
	uint4 *chunk = reinterpret_cast<uint4 *>(reinterpret_cast<char *>(dev->gather) + dev->chunk_bytes * (size_t)((me + (int)dev->owned_parity) % n));
	// Pack and copies run on the READ-BACK stream, behind the copies of the previous call (stream order protects the packed
	// chunks) and beside whatever the main stream does next: a frame that opens with a full clear is drawn into the other
	// framebuffer of the pair while this one is packed and copied.
	const uint32_t items = (uint32_t)(dev->W / 8) * (uint32_t)dev->H;
	const bool profiled = dev->prof_on;
	cudaStream_t rs = profiled ? dev->stream : dev->readback_stream;
	if(profiled) {
		if(dev->readback_in_flight) CUDA_TRY(cudaStreamWaitEvent(dev->stream, dev->ev_readback_done, 0));
		prof_pre(dev, MLV_STAGE_COMPOSITE);
	} else {
		CUDA_TRY(cudaEventRecord(dev->ev_present_src, dev->stream));
		CUDA_TRY(cudaStreamWaitEvent(rs, dev->ev_present_src, 0));
	}
	launch_pdl(k_composite_pack, (items + 255) / 256, 256, rs, dev->fb, chunk, dev->W, dev->H, dev->part);
	if(int rc = check_launch(dev, "k_composite_pack")) return rc;
	if(profiled) {
		CUDA_TRY(cudaEventRecord(dev->ev_resolved, dev->stream));
		CUDA_TRY(cudaStreamWaitEvent(dev->readback_stream, dev->ev_resolved, 0));
	} else {
		CUDA_TRY(cudaEventRecord(dev->ev_fb_read[dev->fb_sel], rs));
		dev->fb_reading[dev->fb_sel] = true;
	}
	const size_t row_bytes = (size_t)dev->W * 4;
	uint32_t local = 0;
	for(int stripe = me; stripe * sh < dev->ht; stripe += n, ++local) { // (ty / sh) % n == me
		const int ty0 = stripe * sh, ty1 = (ty0 + sh < dev->ht) ? ty0 + sh : dev->ht;
		CUDA_TRY(cudaMemcpyAsync((char *)frame_colors + (size_t)ty0 * 8 * row_bytes, (const char *)chunk + (size_t)local * sh * 8 * row_bytes, (size_t)(ty1 - ty0) * 8 * row_bytes,
		                         cudaMemcpyDeviceToHost, dev->readback_stream));
	}
	CUDA_TRY(cudaEventRecord(dev->ev_readback_done, dev->readback_stream));
	dev->readback_in_flight = true;
	return MLV_OK;
}

int mlv_register_host_memory(mlv_device *dev, void *ptr, size_t bytes) {
	if(!dev || !ptr || !bytes) return fail(MLV_ERR_INVALID_ARGUMENT, "null argument");
	mlv_device *d = dev->children ? (*dev->children)[0] : dev;
	if(int rc = use_device(d)) return rc;
	CUDA_TRY(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable)); // portable: every device of a group may copy into it
	return MLV_OK;
}
int mlv_unregister_host_memory(mlv_device *dev, void *ptr) {
	if(!dev || !ptr) return fail(MLV_ERR_INVALID_ARGUMENT, "null argument");
	mlv_device *d = dev->children ? (*dev->children)[0] : dev;
	if(int rc = use_device(d)) return rc;
	CUDA_TRY(cudaHostUnregister(ptr));
	return MLV_OK;
}

int mlv_composite_unpack(mlv_device *dev) {
	if(int rc = immediate_only(dev, "mlv_composite_unpack")) return rc;
	if(int rc = use_device(dev)) return rc;
	if(dev->part.num_ranks <= 1) return fail(MLV_ERR_STATE, "device was created with a single rank");
	const uint32_t quads = (uint32_t)(dev->W / 4) * (uint32_t)dev->H;
	if(dev->readback_in_flight) CUDA_TRY(cudaStreamWaitEvent(dev->stream, dev->ev_readback_done, 0));
	prof_pre(dev, MLV_STAGE_COMPOSITE);
	launch_pdl(k_composite_unpack, (quads + 255) / 256, 256, dev->stream, dev->gather, dev->resolved_color, dev->W, dev->H, dev->part.num_ranks, dev->part.stripe_h, dev->chunk_bytes / 16);
	dev->present_color = dev->resolved_color;
	return check_launch(dev, "k_composite_unpack");
}

// ---- peer-memory compositing ---------------------------------------------------------------------

int mlv_composite_peer_export(mlv_device *dev, mlv_peer_info *out) {
	if(int rc = immediate_only(dev, "mlv_composite_peer_export")) return rc;
	if(int rc = use_device(dev)) return rc;
	if(!out) return fail(MLV_ERR_INVALID_ARGUMENT, "null argument");
	if(dev->part.num_ranks <= 1) return fail(MLV_ERR_STATE, "device was created with a single rank");
	if(dev->part.num_ranks > MLV_MAX_PEERS) return fail(MLV_ERR_INVALID_ARGUMENT, "peer-memory compositing supports up to %d ranks", MLV_MAX_PEERS);
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "mlv_peer_info carries 64-byte IPC handles");
	const size_t image_bytes = (size_t)dev->W * dev->H * 4;
	if(!dev->p2p_flags) {
		CUDA_TRY(cudaMalloc(&dev->p2p_color[0], image_bytes));
		CUDA_TRY(cudaMalloc(&dev->p2p_color[1], image_bytes));
		CUDA_TRY(cudaMalloc(&dev->p2p_flags, MLV_MAX_PEERS * sizeof(uint32_t)));
		CUDA_TRY(cudaMemsetAsync(dev->p2p_color[0], 0, image_bytes, dev->stream));
		CUDA_TRY(cudaMemsetAsync(dev->p2p_color[1], 0, image_bytes, dev->stream));
		CUDA_TRY(cudaMemsetAsync(dev->p2p_flags, 0, MLV_MAX_PEERS * sizeof(uint32_t), dev->stream));
		CUDA_TRY(cudaStreamSynchronize(dev->stream)); // peers may write as soon as they have the handles
	}
	memset(out, 0, sizeof(*out));
	cudaIpcMemHandle_t h;
	for(int k = 0; k < 2; ++k) {
		CUDA_TRY(cudaIpcGetMemHandle(&h, dev->p2p_color[k]));
		memcpy(out->ipc_color[k], &h, 64);
		out->color[k] = dev->p2p_color[k];
	}
	CUDA_TRY(cudaIpcGetMemHandle(&h, dev->p2p_flags));
	memcpy(out->ipc_flags, &h, 64);
	out->flags = dev->p2p_flags;
	out->cuda_device = dev->cuda_dev;
	return MLV_OK;
}

int mlv_composite_peer_attach(mlv_device *dev, const mlv_peer_info *infos, int same_process) {
	if(int rc = immediate_only(dev, "mlv_composite_peer_attach")) return rc;
	if(int rc = use_device(dev)) return rc;
	if(!infos) return fail(MLV_ERR_INVALID_ARGUMENT, "null argument");
	if(!dev->p2p_flags) return fail(MLV_ERR_STATE, "call mlv_composite_peer_export first");
	if(dev->peers_attached) return fail(MLV_ERR_STATE, "peers already attached");
	const int n = dev->part.num_ranks, me = dev->part.rank;
	for(int p = 0; p < n; ++p) {
		if(p == me) {
			dev->peer_color[0][p] = dev->p2p_color[0];
			dev->peer_color[1][p] = dev->p2p_color[1];
			dev->peer_flags[p] = dev->p2p_flags;
			continue;
		}
		if(same_process) {
			if(infos[p].cuda_device != dev->cuda_dev) {
				cudaError_t e = cudaDeviceEnablePeerAccess(infos[p].cuda_device, 0);
				if(e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
				else if(e != cudaSuccess) return fail(MLV_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d): %s", infos[p].cuda_device, cudaGetErrorString(e));
			}
			dev->peer_color[0][p] = (uint4 *)infos[p].color[0];
			dev->peer_color[1][p] = (uint4 *)infos[p].color[1];
			dev->peer_flags[p] = (uint32_t *)infos[p].flags;
			continue;
		}
		void *mapped[3];
		const unsigned char *handles[3] = { infos[p].ipc_color[0], infos[p].ipc_color[1], infos[p].ipc_flags };
		for(int k = 0; k < 3; ++k) {
			cudaIpcMemHandle_t h;
			memcpy(&h, handles[k], 64);
			cudaError_t e = cudaIpcOpenMemHandle(&mapped[k], h, cudaIpcMemLazyEnablePeerAccess);
			if(e != cudaSuccess) return fail(MLV_ERR_CUDA, "cudaIpcOpenMemHandle (rank %d): %s", p, cudaGetErrorString(e));
			dev->ipc_opened[dev->ipc_opened_count++] = mapped[k];
		}
		dev->peer_color[0][p] = (uint4 *)mapped[0];
		dev->peer_color[1][p] = (uint4 *)mapped[1];
		dev->peer_flags[p] = (uint32_t *)mapped[2];
	}
	dev->peers_attached = true;
	return MLV_OK;
}

int mlv_composite_broadcast(mlv_device *dev) {
	if(int rc = immediate_only(dev, "mlv_composite_broadcast")) return rc;
	if(int rc = use_device(dev)) return rc;
	if(!dev->peers_attached) return fail(MLV_ERR_STATE, "call mlv_composite_peer_attach first");
	if(dev->bcast_pending) return fail(MLV_ERR_STATE, "mlv_composite_wait must follow every mlv_composite_broadcast");
	if(dev->xchg_pending) return fail(MLV_ERR_STATE, "mlv_composite_join must follow every mlv_composite_broadcast_async");
	if(int rc = flush_clears(dev)) return rc;
	// peers overwrite the image a read-back may still be copying only after this broadcast has run (they wait for it)
	if(dev->readback_in_flight) CUDA_TRY(cudaStreamWaitEvent(dev->stream, dev->ev_readback_done, 0));
	const uint32_t seq = ++dev->p2p_seq;
	PeerTargets t;
	memset(&t, 0, sizeof(t));
	for(int p = 0; p < dev->part.num_ranks; ++p) {
		t.color[p] = dev->peer_color[seq & 1u][p];
		t.flags[p] = dev->peer_flags[p];
	}
	const uint32_t items = (uint32_t)(dev->W / 8) * (uint32_t)dev->H;
	prof_pre(dev, MLV_STAGE_COMPOSITE);
	launch_pdl(k_composite_broadcast, (items + 255) / 256, 256, dev->stream, dev->fb, t, dev->W, dev->H, dev->part, seq, dev->ctr, 0, dev->part.num_ranks, 1);
	dev->bcast_pending = true;
	return check_launch(dev, "k_composite_broadcast");
}

// The exchange of frame f off the critical path: broadcast + wait go to the exchange stream, ordered after everything
// issued so far on the main stream, and the main stream carries on with frame f+1 at once (in the other tiled
// framebuffer when that frame starts with a full clear, see acquire_framebuffer).
int mlv_composite_broadcast_async(mlv_device *dev) {
	if(int rc = immediate_only(dev, "mlv_composite_broadcast_async")) return rc;
	if(int rc = use_device(dev)) return rc;
	if(!dev->peers_attached) return fail(MLV_ERR_STATE, "call mlv_composite_peer_attach first");
	if(dev->bcast_pending) return fail(MLV_ERR_STATE, "mlv_composite_wait must follow every mlv_composite_broadcast");
	if(dev->xchg_pending) return fail(MLV_ERR_STATE, "mlv_composite_join must follow every mlv_composite_broadcast_async");
	if(int rc = flush_clears(dev)) return rc;
	const uint32_t seq = ++dev->p2p_seq;
	PeerTargets t;
	memset(&t, 0, sizeof(t));
	for(int p = 0; p < dev->part.num_ranks; ++p) {
		t.color[p] = dev->peer_color[seq & 1u][p];
		t.flags[p] = dev->peer_flags[p];
	}
	// after the frame's draws -- and after every consumer of the previously joined image issued on the main stream:
	// peers overwrite that image only once this rank's broadcast of the next frame has run (they wait for it)
	CUDA_TRY(cudaEventRecord(dev->ev_frame_done, dev->stream));
	CUDA_TRY(cudaStreamWaitEvent(dev->xchg_stream, dev->ev_frame_done, 0));
	if(dev->readback_in_flight) CUDA_TRY(cudaStreamWaitEvent(dev->xchg_stream, dev->ev_readback_done, 0));
	const uint32_t items = (uint32_t)(dev->W / 8) * (uint32_t)dev->H;
	const int n = dev->part.num_ranks, me = dev->part.rank;
	const bool band = (uint64_t)dev->part.stripe_h * (uint64_t)n >= (uint64_t)dev->ht; // one contiguous band of rows per rank
	if(band) {
		// Copy-engine form: a short kernel resolves this rank's band into its OWN row-major image, the band (one
		// contiguous range of rows) then travels to every peer by DMA over NVLink -- no SM is busy with the transfer,
		// which therefore really runs underneath the next frame's kernels -- and a one-warp kernel publishes the
		// arrival. Peers are visited in ring order so that every image has one writer at a time.
		launch_pdl(k_composite_broadcast, (items + 255) / 256, 256, dev->xchg_stream, dev->fb, t, dev->W, dev->H, dev->part, seq, dev->ctr, me, me + 1, 0);
		CUDA_TRY(cudaEventRecord(dev->ev_fb_free[dev->fb_sel], dev->xchg_stream));
		const uint32_t row_lo = (uint32_t)dev->part.stripe_h * (uint32_t)me;
		uint32_t row_hi = row_lo + (uint32_t)dev->part.stripe_h;
		if(row_hi > (uint32_t)dev->ht) row_hi = (uint32_t)dev->ht;
		if(row_lo < row_hi) {
			const size_t offset = (size_t)row_lo * 8u * (size_t)dev->W * 4u, bytes = (size_t)(row_hi - row_lo) * 8u * (size_t)dev->W * 4u;
			for(int k = 1; k < n; ++k) {
				const int p = (me + k) % n;
				CUDA_TRY(cudaMemcpyAsync((char *)t.color[p] + offset, (const char *)t.color[me] + offset, bytes, cudaMemcpyDefault, dev->xchg_stream));
			}
		}
		launch_pdl(k_composite_publish, 1, 32, dev->xchg_stream, t, n, me, seq);
		dev->launches += 1;
	} else {
		launch_pdl(k_composite_broadcast, (items + 255) / 256, 256, dev->xchg_stream, dev->fb, t, dev->W, dev->H, dev->part, seq, dev->ctr, 0, n, 1);
		CUDA_TRY(cudaEventRecord(dev->ev_fb_free[dev->fb_sel], dev->xchg_stream));
	}
	dev->fb_busy[dev->fb_sel] = true;
	wait_flags_on_stream(dev, dev->xchg_stream, seq);
	CUDA_TRY(cudaEventRecord(dev->ev_xchg_done, dev->xchg_stream));
	cudaError_t e = cudaGetLastError();
	if(e != cudaSuccess) return fail(MLV_ERR_CUDA, "launch of k_composite_broadcast/k_composite_wait failed: %s", cudaGetErrorString(e));
	dev->launches += 2;
	dev->xchg_pending = true;
	return MLV_OK;
}

int mlv_composite_join(mlv_device *dev) {
	if(int rc = immediate_only(dev, "mlv_composite_join")) return rc;
	if(int rc = use_device(dev)) return rc;
	if(!dev->xchg_pending) return fail(MLV_ERR_STATE, "no asynchronous broadcast to join");
	CUDA_TRY(cudaStreamWaitEvent(dev->stream, dev->ev_xchg_done, 0));
	dev->xchg_pending = false;
	dev->present_color = dev->p2p_color[dev->p2p_seq & 1u];
	return MLV_OK;
}

// Read-back of the composited image (the one mlv_composite_wait / mlv_composite_join handed out) on the read-back stream.
int mlv_composite_readback_async(mlv_device *dev, uint32_t *colors) {
	if(int rc = immediate_only(dev, "mlv_composite_readback_async")) return rc;
	if(int rc = use_device(dev)) return rc;
	if(!colors) return fail(MLV_ERR_INVALID_ARGUMENT, "null argument");
	if(dev->part.num_ranks <= 1) return fail(MLV_ERR_STATE, "device was created with a single rank");
	if(dev->bcast_pending || dev->xchg_pending) return fail(MLV_ERR_STATE, "call mlv_composite_wait / mlv_composite_join first");
	CUDA_TRY(cudaEventRecord(dev->ev_resolved, dev->stream)); // after the wait / join
	CUDA_TRY(cudaStreamWaitEvent(dev->readback_stream, dev->ev_resolved, 0));
	CUDA_TRY(cudaMemcpyAsync(colors, dev->present_color, (size_t)dev->W * dev->H * 4, cudaMemcpyDeviceToHost, dev->readback_stream));
	CUDA_TRY(cudaEventRecord(dev->ev_readback_done, dev->readback_stream));
	dev->readback_in_flight = true;
	return MLV_OK;
}

// (device groups, NCCL exchange) the image mlv_composite_unpack left, to the host on the read-back stream
static int mlv_present_copy_async(mlv_device *dev, uint32_t *colors) {
	if(int rc = use_device(dev)) return rc;
	if(dev->readback_in_flight) CUDA_TRY(cudaEventSynchronize(dev->ev_readback_done));
	CUDA_TRY(cudaEventRecord(dev->ev_resolved, dev->stream));
	CUDA_TRY(cudaStreamWaitEvent(dev->readback_stream, dev->ev_resolved, 0));
	CUDA_TRY(cudaMemcpyAsync(colors, dev->present_color, (size_t)dev->W * dev->H * 4, cudaMemcpyDeviceToHost, dev->readback_stream));
	CUDA_TRY(cudaEventRecord(dev->ev_readback_done, dev->readback_stream));
	dev->readback_in_flight = true;
	return MLV_OK;
}

int mlv_composite_wait(mlv_device *dev) {
	if(int rc = immediate_only(dev, "mlv_composite_wait")) return rc;
	if(int rc = use_device(dev)) return rc;
	if(!dev->bcast_pending) return fail(MLV_ERR_STATE, "no broadcast to wait for");
	prof_pre(dev, MLV_STAGE_COMPOSITE);
	wait_flags_on_stream(dev, dev->stream, dev->p2p_seq);
	dev->bcast_pending = false;
	dev->present_color = dev->p2p_color[dev->p2p_seq & 1u];
	return check_launch(dev, "k_composite_wait");
}

// ---- debug read-back -------------------------------------------------------------------------------

static int debug_counters(mlv_device *dev, Counters *c) {
	if(int rc = use_device(dev)) return rc;
	if(int rc = immediate_only(dev, "mlv_debug_read_*")) return rc;
	if(!(dev->desc.flags & MLV_DEVICE_DEBUG_CAPTURE)) return fail(MLV_ERR_STATE, "device was created without MLV_DEVICE_DEBUG_CAPTURE");
	CUDA_TRY(cudaMemcpyAsync(c, dev->ctr, sizeof(*c), cudaMemcpyDeviceToHost, dev->stream));
	CUDA_TRY(cudaStreamSynchronize(dev->stream));
	return MLV_OK;
}

int mlv_debug_read_vs_out(mlv_device *dev, float *out12_per_vertex, uint32_t *out_vertex_count) {
	Counters c;
	if(int rc = debug_counters(dev, &c)) return rc;
	if(out_vertex_count) *out_vertex_count = dev->last_index_count;
	if(out12_per_vertex && dev->last_index_count) CUDA_TRY(cudaMemcpy(out12_per_vertex, dev->dbg.vs_out, (size_t)dev->last_index_count * 48, cudaMemcpyDeviceToHost));
	return MLV_OK;
}

// The per-tile lists of the last draw, marshalled into the reference's layout: bins in ascending bin index, every list
// segment placed where the reference's exclusive scan (main.c:937-942) puts it. On the device the segments lie in
// allocation order (k_fill) and the work list is unordered.
static int read_sorted_lists(mlv_device *dev, const Counters &c, std::vector<uint32_t> &slots, std::vector<mlv_ref_compacted_bin> &bins, std::vector<uint32_t> *src = nullptr) {
	slots.clear();
	bins.clear();
	if(src) src->clear();
	if(dev->last_index_count == 0 || c.last_pair_total == 0xffffffffu) return MLV_OK;
	const uint32_t pairs = c.last_pair_total, nb = c.last_n_cbins;
	std::vector<uint32_t> raw(pairs);
	std::vector<mlv_ref_compacted_bin> cb(nb);
	if(pairs) CUDA_TRY(cudaMemcpy(raw.data(), dev->pair_ids, (size_t)pairs * 4, cudaMemcpyDeviceToHost));
	if(nb) CUDA_TRY(cudaMemcpy(cb.data(), dev->cbins, (size_t)nb * sizeof(mlv_ref_compacted_bin), cudaMemcpyDeviceToHost));
	std::sort(cb.begin(), cb.end(), [](const mlv_ref_compacted_bin &a, const mlv_ref_compacted_bin &b) { return a.bin_index < b.bin_index; });
	slots.reserve(pairs);
	for(mlv_ref_compacted_bin &b : cb) {
		const uint32_t base = b.num_triangles_upto;
		b.num_triangles_upto = (uint32_t)slots.size();
		for(uint32_t k = 0; k < b.num_triangles_self; ++k) {
			slots.push_back(base + k < pairs ? raw[base + k] : 0xffffffffu);
			if(src) src->push_back(base + k);
		}
	}
	bins.swap(cb);
	return MLV_OK;
}

// The device names triangles by order-preserving keys and stores them in slots (mlv_internal.cuh). The debug
// read-back marshals that into the reference's compact numbering: reference id i == the i-th smallest key.
struct DebugMap {
	std::vector<uint32_t> keys;  // ascending
	std::vector<uint32_t> slots; // slot of keys[i]
	std::vector<uint32_t> slot_rank; // reference id of a slot
	uint32_t rank(uint32_t key) const { return (uint32_t)(std::lower_bound(keys.begin(), keys.end(), key) - keys.begin()); }
};

static int build_debug_map(mlv_device *dev, const Counters &c, DebugMap &m) {
	const uint32_t n_slots = dev->last_direct_slots + c.last_ovf_count;
	std::vector<uint32_t> slot_key(n_slots);
	if(n_slots) CUDA_TRY(cudaMemcpy(slot_key.data(), dev->dbg.slot_key, (size_t)n_slots * 4, cudaMemcpyDeviceToHost));
	std::vector<std::pair<uint32_t, uint32_t>> ks;
	for(uint32_t s = 0; s < n_slots; ++s)
		if(slot_key[s] != 0xffffffffu) ks.emplace_back(slot_key[s], s);
	std::sort(ks.begin(), ks.end());
	m.keys.resize(ks.size());
	m.slots.resize(ks.size());
	m.slot_rank.assign(n_slots, 0xffffffffu);
	for(size_t i = 0; i < ks.size(); ++i) {
		m.keys[i] = ks[i].first;
		m.slots[i] = ks[i].second;
		m.slot_rank[ks[i].second] = (uint32_t)i;
	}
	return MLV_OK;
}

int mlv_debug_read_triangles(mlv_device *dev, mlv_ref_triangle *tris, float *attributes36, uint32_t *out_count) {
	Counters c;
	if(int rc = debug_counters(dev, &c)) return rc;
	if(dev->last_index_count == 0) {
		if(out_count) *out_count = 0;
		return MLV_OK;
	}
	DebugMap m;
	if(int rc = build_debug_map(dev, c, m)) return rc;
	if(out_count) *out_count = (uint32_t)m.keys.size();
	const uint32_t n_slots = dev->last_direct_slots + c.last_ovf_count;
	if(tris && !m.keys.empty()) {
		std::vector<mlv_ref_triangle> all(n_slots);
		CUDA_TRY(cudaMemcpy(all.data(), dev->dbg.tris, (size_t)n_slots * sizeof(mlv_ref_triangle), cudaMemcpyDeviceToHost));
		for(size_t i = 0; i < m.keys.size(); ++i) {
			tris[i] = all[m.slots[i]];
			tris[i].p_attributes = (uint64_t)i * 144ull;
		}
	}
	if(attributes36 && !m.keys.empty()) {
		std::vector<float> all((size_t)n_slots * 36);
		CUDA_TRY(cudaMemcpy(all.data(), dev->dbg.attrs, (size_t)n_slots * 144, cudaMemcpyDeviceToHost));
		for(size_t i = 0; i < m.keys.size(); ++i) memcpy(attributes36 + i * 36, all.data() + (size_t)m.slots[i] * 36, 144);
	}
	return MLV_OK;
}

int mlv_debug_read_bins(mlv_device *dev, uint32_t *triangle_ids, uint32_t *out_pair_count, mlv_ref_compacted_bin *bins, uint32_t *out_bin_count) {
	Counters c;
	if(int rc = debug_counters(dev, &c)) return rc;
	std::vector<uint32_t> slots;
	std::vector<mlv_ref_compacted_bin> cb;
	if(int rc = read_sorted_lists(dev, c, slots, cb)) return rc;
	const uint32_t pairs = (uint32_t)slots.size(), nb = (uint32_t)cb.size();
	if(out_pair_count) *out_pair_count = pairs;
	if(out_bin_count) *out_bin_count = nb;
	if(triangle_ids && pairs) {
		DebugMap m;
		if(int rc = build_debug_map(dev, c, m)) return rc;
		for(uint32_t i = 0; i < pairs; ++i) triangle_ids[i] = slots[i] < m.slot_rank.size() ? m.slot_rank[slots[i]] : 0xffffffffu; // the lists hold slots
	}
	if(bins && nb) memcpy(bins, cb.data(), (size_t)nb * sizeof(mlv_ref_compacted_bin));
	return MLV_OK;
}

int mlv_debug_read_masks(mlv_device *dev, mlv_ref_tile_info *infos, uint32_t *out_pair_count) {
	Counters c;
	if(int rc = debug_counters(dev, &c)) return rc;
	std::vector<uint32_t> slots;
	std::vector<mlv_ref_compacted_bin> cb;
	std::vector<uint32_t> src; // position of every marshalled pair in the device arrays
	if(int rc = read_sorted_lists(dev, c, slots, cb, &src)) return rc;
	const uint32_t pairs = (uint32_t)slots.size();
	if(out_pair_count) *out_pair_count = pairs;
	if(infos && pairs) {
		DebugMap m;
		if(int rc = build_debug_map(dev, c, m)) return rc;
		const uint32_t raw_pairs = c.last_pair_total;
		std::vector<mlv_ref_tile_info> raw(raw_pairs);
		CUDA_TRY(cudaMemcpy(raw.data(), dev->dbg.infos, (size_t)raw_pairs * sizeof(mlv_ref_tile_info), cudaMemcpyDeviceToHost));
		for(uint32_t i = 0; i < pairs; ++i) {
			infos[i] = raw[src[i]];
			infos[i].triangle_id = m.rank(infos[i].triangle_id);
		}
	}
	return MLV_OK;
}

// The key of every assembled triangle of the last draw in ascending order, i.e. keys[i] = device key of the reference's
// triangle id i (mlv_internal.cuh "Triangle identity").
int mlv_debug_read_keys(mlv_device *dev, uint32_t *keys, uint32_t *out_count) {
	Counters c;
	if(int rc = debug_counters(dev, &c)) return rc;
	if(dev->last_index_count == 0) {
		if(out_count) *out_count = 0;
		return MLV_OK;
	}
	DebugMap m;
	if(int rc = build_debug_map(dev, c, m)) return rc;
	if(out_count) *out_count = (uint32_t)m.keys.size();
	if(keys && !m.keys.empty()) memcpy(keys, m.keys.data(), m.keys.size() * sizeof(uint32_t));
	return MLV_OK;
}

// The per-tile lists of the last draw exactly as k_tile consumed them -- device keys, in arrival order, Hi-Z-rejected
// pairs already removed -- and the work list of bins. Works WITHOUT debug capture: this is the production data path.
int mlv_read_bin_lists(mlv_device *dev, uint32_t *keys, uint32_t *out_pair_count, mlv_ref_compacted_bin *bins, uint32_t *out_bin_count) {
	if(int rc = immediate_only(dev, "mlv_read_bin_lists")) return rc;
	if(int rc = use_device(dev)) return rc;
	Counters c;
	CUDA_TRY(cudaMemcpyAsync(&c, dev->ctr, sizeof(c), cudaMemcpyDeviceToHost, dev->stream));
	CUDA_TRY(cudaStreamSynchronize(dev->stream));
	std::vector<uint32_t> slots;
	std::vector<mlv_ref_compacted_bin> cb;
	if(int rc = read_sorted_lists(dev, c, slots, cb)) return rc;
	const uint32_t pairs = (uint32_t)slots.size(), nb = (uint32_t)cb.size();
	if(out_pair_count) *out_pair_count = pairs;
	if(out_bin_count) *out_bin_count = nb;
	if(keys && pairs) { // slots -> keys: direct slot t has key t << 3, an overflow slot's key is word 3 of its bounds entry
		const uint32_t T = dev->last_direct_slots, n_ovf = c.last_ovf_count;
		std::vector<uint4> ovf(n_ovf);
		if(n_ovf && dev->last_ctx) CUDA_TRY(cudaMemcpy(ovf.data(), dev->last_ctx->tri_bounds + T, (size_t)n_ovf * sizeof(uint4), cudaMemcpyDeviceToHost));
		for(uint32_t i = 0; i < pairs; ++i) keys[i] = slots[i] < T ? (slots[i] << 3) : (slots[i] - T < n_ovf ? ovf[slots[i] - T].w : 0xffffffffu);
	}
	if(bins && nb) memcpy(bins, cb.data(), (size_t)nb * sizeof(mlv_ref_compacted_bin));
	return check_flags(dev, c);
}

int mlv_debug_read_tile_min_depths(mlv_device *dev, float *out_bins) {
	if(int rc = immediate_only(dev, "mlv_debug_read_tile_min_depths")) return rc;
	if(int rc = use_device(dev)) return rc;
	if(!out_bins) return fail(MLV_ERR_INVALID_ARGUMENT, "null argument");
	if(int rc = flush_clears(dev)) return rc;
	CUDA_TRY(cudaMemcpyAsync(out_bins, dev->tile_min, (size_t)dev->num_bins * 4, cudaMemcpyDeviceToHost, dev->stream));
	CUDA_TRY(cudaStreamSynchronize(dev->stream));
	return MLV_OK;
}

int mlv_profile_begin(mlv_device *dev) {
	if(int rc = immediate_only(dev, "mlv_profile_begin")) return rc;
	if(int rc = use_device(dev)) return rc;
	CUDA_TRY(cudaStreamSynchronize(dev->stream));
	dev->prof_used = 0;
	dev->prof_stages->clear();
	dev->prof_on = true;
	return MLV_OK;
}

int mlv_profile_end(mlv_device *dev, double *out_ms, uint32_t *out_launches) {
	if(int rc = immediate_only(dev, "mlv_profile_end")) return rc;
	if(int rc = use_device(dev)) return rc;
	if(!out_ms || !out_launches) return fail(MLV_ERR_INVALID_ARGUMENT, "null argument");
	dev->prof_on = false;
	CUDA_TRY(cudaStreamSynchronize(dev->stream));
	for(int i = 0; i < MLV_STAGE_COUNT; ++i) {
		out_ms[i] = 0.0;
		out_launches[i] = 0;
	}
	for(size_t k = 0; k < dev->prof_stages->size(); ++k) {
		float ms = 0.f;
		CUDA_TRY(cudaEventElapsedTime(&ms, (*dev->prof_events)[2 * k], (*dev->prof_events)[2 * k + 1]));
		const int st = (*dev->prof_stages)[k];
		out_ms[st] += ms;
		out_launches[st]++;
	}
	return MLV_OK;
}

// Per-launch records of the last profiled region (valid after mlv_profile_end until the next mlv_profile_begin).
int mlv_profile_read_events(mlv_device *dev, mlv_profile_event *out, uint32_t capacity, uint32_t *out_count) {
	if(int rc = use_device(dev)) return rc;
	if(!out_count) return fail(MLV_ERR_INVALID_ARGUMENT, "null argument");
	if(dev->prof_on) return fail(MLV_ERR_STATE, "call mlv_profile_end first");
	const size_t n = dev->prof_stages->size();
	*out_count = (uint32_t)n;
	if(!out) return MLV_OK;
	for(size_t k = 0; k < n && k < capacity; ++k) {
		float start = 0.f, dur = 0.f;
		CUDA_TRY(cudaEventElapsedTime(&start, (*dev->prof_events)[0], (*dev->prof_events)[2 * k]));
		CUDA_TRY(cudaEventElapsedTime(&dur, (*dev->prof_events)[2 * k], (*dev->prof_events)[2 * k + 1]));
		out[k].stage = (*dev->prof_stages)[k];
		out[k].start_ms = start;
		out[k].duration_ms = dur;
	}
	return MLV_OK;
}

int mlv_timeline_begin(mlv_device *dev) {
	if(int rc = use_device(dev)) return rc;
	if(!dev->timeline) {
		CUDA_TRY(cudaMalloc(&dev->timeline, (size_t)MLV_TIMELINE_CAPACITY * 4 * sizeof(unsigned long long)));
		dev->timeline_stages = new std::vector<int>();
	}
	dev->timeline_stages->clear();
	dev->timeline_used = 0;
	dev->timeline_draw = 0;
	dev->timeline_on = true;
	return mlv_timeline_reset(dev);
}

int mlv_timeline_end(mlv_device *dev) {
	if(!dev) return fail(MLV_ERR_INVALID_ARGUMENT, "null argument");
	dev->timeline_on = false;
	return MLV_OK;
}

int mlv_timeline_reset(mlv_device *dev) {
	if(int rc = use_device(dev)) return rc;
	if(!dev->timeline) return fail(MLV_ERR_STATE, "call mlv_timeline_begin first");
	if(dev->recording) return fail(MLV_ERR_STATE, "mlv_timeline_reset cannot be recorded");
	CUDA_TRY(cudaMemsetAsync(dev->timeline, 0xff, (size_t)MLV_TIMELINE_CAPACITY * 4 * sizeof(unsigned long long), dev->stream));
	return MLV_OK;
}

int mlv_timeline_read(mlv_device *dev, mlv_timeline_event *out, uint32_t capacity, uint32_t *out_count) {
	if(int rc = use_device(dev)) return rc;
	if(!dev->timeline || !out_count) return fail(MLV_ERR_STATE, "call mlv_timeline_begin first");
	*out_count = dev->timeline_used;
	if(!out) return MLV_OK;
	if(capacity < dev->timeline_used) return fail(MLV_ERR_INVALID_ARGUMENT, "capacity %u < %u timeline events", capacity, dev->timeline_used);
	CUDA_TRY(cudaDeviceSynchronize());
	std::vector<unsigned long long> raw((size_t)dev->timeline_used * 4);
	CUDA_TRY(cudaMemcpy(raw.data(), dev->timeline, raw.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
	unsigned long long t0 = ~0ull;
	for(uint32_t i = 0; i < dev->timeline_used; ++i) t0 = std::min(t0, raw[4 * (size_t)i]);
	for(uint32_t i = 0; i < dev->timeline_used; ++i) {
		const unsigned long long *r = &raw[4 * (size_t)i];
		out[i].stage = (*dev->timeline_stages)[i] >> 16;
		out[i].draw = (*dev->timeline_stages)[i] & 0xffff;
		const bool ran = r[0] != ~0ull; // (a launch that did not execute since the last reset)
		out[i].resident_us = ran ? (double)(r[0] - t0) * 1e-3 : -1.0;
		out[i].start_us = ran && r[1] != ~0ull ? (double)(r[1] - t0) * 1e-3 : -1.0;
		out[i].end_us = ran && r[2] != ~0ull ? (double)(~r[2] - t0) * 1e-3 : -1.0;
	}
	return MLV_OK;
}

uint64_t mlv_kernel_launch_count(mlv_device *dev) {
	if(dev && dev->children) {
		uint64_t n = 0;
		for(mlv_device *c : *dev->children) n += c->launches;
		return n;
	}
	return dev ? dev->launches : 0;
}

// Word-wise 64-bit FNV-1a over u32 words (h = 0xcbf29ce484222325; h ^= w; h *= 0x100000001b3): the frame hash of
// tests/golden/golden.json, so that a host can compare a read-back frame with the committed one without Python loops.
uint64_t mlv_fnv64_words(const uint32_t *words, size_t count) {
	uint64_t h = 0xcbf29ce484222325ull;
	for(size_t i = 0; i < count; ++i) h = (h ^ (uint64_t)words[i]) * 0x100000001b3ull;
	return h;
}

// ---- device groups (mlv_device_desc.num_gpus > 1) -----------------------------------------------------
// One host thread, N CUDA devices in one process: rank i of the sort-first split lives on CUDA device cuda_device + i
// (MLV_DEVICE_GROUP_SAME_GPU: all on one device, for tests on a single-GPU box). Geometry, textures and state are
// replicated by the fan-out above; the frame is composed by the asynchronous peer-memory exchange over plain peer pointers
// (mlv_composite_peer_attach, same process: no IPC handles) or, with MLV_DEVICE_GROUP_NCCL, by pack + in-place
// ncclAllGather + unpack inside one NCCL group call. NCCL is loaded on demand (libnccl.so.2): the library has no
// load-time dependency on it and the peer-memory path needs none.
struct GroupNccl {
	void *lib;
	std::vector<void *> comms;
	int (*CommInitAll)(void **, int, const int *);
	int (*CommDestroy)(void *);
	int (*GroupStart)();
	int (*GroupEnd)();
	int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t);
	const char *(*GetErrorString)(int);
};

static int group_nccl_init(mlv_device *dev) {
	GroupNccl *g = new GroupNccl();
	g->lib = nullptr;
	for(const char *name : { "libnccl.so.2", "libnccl.so" })
		if((g->lib = dlopen(name, RTLD_NOW | RTLD_LOCAL))) break;
	if(!g->lib) {
		delete g;
		return fail(MLV_ERR_STATE, "MLV_DEVICE_GROUP_NCCL: cannot load libnccl.so.2 (%s)", dlerror());
	}
	g->CommInitAll = (int (*)(void **, int, const int *))dlsym(g->lib, "ncclCommInitAll");
	g->CommDestroy = (int (*)(void *))dlsym(g->lib, "ncclCommDestroy");
	g->GroupStart = (int (*)())dlsym(g->lib, "ncclGroupStart");
	g->GroupEnd = (int (*)())dlsym(g->lib, "ncclGroupEnd");
	g->AllGather = (int (*)(const void *, void *, size_t, int, void *, cudaStream_t))dlsym(g->lib, "ncclAllGather");
	g->GetErrorString = (const char *(*)(int))dlsym(g->lib, "ncclGetErrorString");
	if(!g->CommInitAll || !g->CommDestroy || !g->GroupStart || !g->GroupEnd || !g->AllGather || !g->GetErrorString) {
		dlclose(g->lib);
		delete g;
		return fail(MLV_ERR_STATE, "MLV_DEVICE_GROUP_NCCL: libnccl lacks an expected symbol");
	}
	std::vector<int> devs;
	for(mlv_device *c : *dev->children) devs.push_back(c->cuda_dev);
	g->comms.assign(devs.size(), nullptr);
	const int rc = g->CommInitAll(g->comms.data(), (int)devs.size(), devs.data());
	if(rc != 0) {
		const int code = fail(MLV_ERR_CUDA, "ncclCommInitAll over %zu devices: %s", devs.size(), g->GetErrorString(rc));
		dlclose(g->lib);
		delete g;
		return code;
	}
	dev->group_nccl = g;
	return MLV_OK;
}

static void group_destroy(mlv_device *dev) {
	if(dev->group_nccl) {
		GroupNccl *g = (GroupNccl *)dev->group_nccl;
		for(size_t i = 0; i < g->comms.size(); ++i) {
			cudaSetDevice((*dev->children)[i]->cuda_dev);
			if(g->comms[i]) g->CommDestroy(g->comms[i]);
		}
		dlclose(g->lib);
		delete g;
	}
	for(mlv_device *c : *dev->children) mlv_finish(c); // nobody writes into a peer's images any more
	for(mlv_device *c : *dev->children) mlv_destroy_device(c);
	delete dev->children;
	free(dev->group_scratch_color);
	delete dev;
}

static int group_create(const mlv_device_desc *desc, mlv_device **out_device) {
	*out_device = nullptr;
	const uint32_t n = desc->num_gpus;
	if(n > MLV_MAX_PEERS) return fail(MLV_ERR_INVALID_ARGUMENT, "num_gpus %u: at most %d", n, MLV_MAX_PEERS);
	if(desc->num_ranks > 1) return fail(MLV_ERR_INVALID_ARGUMENT, "num_gpus and num_ranks exclude each other: a group creates its own ranks");
	if(desc->flags & MLV_DEVICE_DEBUG_CAPTURE) return fail(MLV_ERR_INVALID_ARGUMENT, "debug capture is per device, not per group");
	int first = desc->cuda_device, count = 0;
	if(first < 0 && cudaGetDevice(&first) != cudaSuccess) return fail(MLV_ERR_CUDA, "no CUDA device");
	const bool same = (desc->flags & MLV_DEVICE_GROUP_SAME_GPU) != 0;
	if(cudaGetDeviceCount(&count) != cudaSuccess || (!same && first + (int)n > count)) return fail(MLV_ERR_CUDA, "num_gpus %u from device %d: only %d CUDA devices are visible", n, first, count);
	mlv_device *g = new(std::nothrow) mlv_device();
	if(!g) return fail(MLV_ERR_OUT_OF_MEMORY, "host allocation failed");
	memset((void *)g, 0, sizeof(*g));
	g->desc = *desc;
	g->W = (int)desc->width, g->H = (int)desc->height;
	g->children = new std::vector<mlv_device *>();
	// one contiguous band of tile rows per GPU unless the caller chose a stripe height (interleaved stripes balance scenes
	// whose cost varies down the screen)
	const uint32_t ht = desc->height / 8u;
	mlv_device_desc cd = *desc;
	cd.num_gpus = 0;
	cd.num_ranks = n;
	cd.flags = desc->flags & ~(uint32_t)(MLV_DEVICE_GROUP_SAME_GPU | MLV_DEVICE_GROUP_NCCL | MLV_DEVICE_GROUP_PEER_EXCHANGE);
	cd.stripe_height_tiles = desc->stripe_height_tiles ? desc->stripe_height_tiles : (ht + n - 1u) / n;
	std::vector<mlv_peer_info> infos(n);
	for(uint32_t i = 0; i < n; ++i) {
		cd.rank = i;
		cd.cuda_device = same ? first : first + (int)i;
		mlv_device *c = nullptr;
		int rc = mlv_create_device(&cd, &c);
		if(rc == MLV_OK) {
			g->children->push_back(c);
			rc = mlv_composite_peer_export(c, &infos[i]);
		}
		if(rc != MLV_OK) {
			group_destroy(g);
			return rc;
		}
	}
	for(mlv_device *c : *g->children)
		if(int rc = mlv_composite_peer_attach(c, infos.data(), 1)) {
			group_destroy(g);
			return rc;
		}
	if(desc->flags & MLV_DEVICE_GROUP_NCCL)
		if(int rc = group_nccl_init(g)) {
			group_destroy(g);
			return rc;
		}
	*out_device = g;
	return MLV_OK;
}

// The frame of a group: exchange the ranks' stripes, read rank 0's composed image back. Depth is not exchanged between
// the ranks (nothing on the device needs it): when the caller wants it, every rank resolves its own depth image and the
// host keeps the rows each rank owns.
static int group_present(mlv_device *dev, uint32_t *colors, float *depths, bool wait) {
	std::vector<mlv_device *> &ch = *dev->children;
	if(dev->group_nccl) {
		GroupNccl *g = (GroupNccl *)dev->group_nccl;
		for(mlv_device *c : ch)
			if(int rc = mlv_composite_pack(c)) return rc;
		int rc = g->GroupStart();
		for(size_t i = 0; i < ch.size() && rc == 0; ++i) {
			cudaSetDevice(ch[i]->cuda_dev);
			rc = g->AllGather((const char *)ch[i]->gather + i * ch[i]->chunk_bytes, ch[i]->gather, ch[i]->chunk_bytes, /* ncclChar */ 0, g->comms[i], ch[i]->stream);
		}
		const int rc_end = g->GroupEnd();
		if(rc != 0 || rc_end != 0) return fail(MLV_ERR_CUDA, "ncclAllGather: %s", g->GetErrorString(rc ? rc : rc_end));
		if(int rc2 = mlv_composite_unpack(ch[0])) return rc2; // (only the rank that is read back needs the row-major image)
	} else if(dev->desc.flags & MLV_DEVICE_GROUP_PEER_EXCHANGE) {
		for(mlv_device *c : ch)
			if(int rc = mlv_composite_broadcast_async(c)) return rc;
		for(mlv_device *c : ch)
			if(int rc = mlv_composite_join(c)) return rc;
	}
	if(colors) {
		if(dev->group_nccl) {
			if(int rc = mlv_present_copy_async(ch[0], colors)) return rc;
		} else if(dev->desc.flags & MLV_DEVICE_GROUP_PEER_EXCHANGE) {
			if(int rc = mlv_composite_readback_async(ch[0], colors)) return rc;
		} else { // default: the frame is composed in host memory, every GPU delivers its band over its own PCIe link
			if(int rc = mlv_present_owned_rows_async(dev, colors)) return rc;
		}
	}
	if(depths) {
		const size_t pixels = (size_t)dev->W * dev->H;
		if(!dev->group_scratch_color && !(dev->group_scratch_color = (uint32_t *)malloc(pixels * 4))) return fail(MLV_ERR_OUT_OF_MEMORY, "host allocation failed");
		float *scratch = (float *)dev->group_scratch_color;
		for(mlv_device *c : ch) {
			if(int rc = mlv_resolve(c)) return rc;
			cudaSetDevice(c->cuda_dev);
			CUDA_TRY(cudaMemcpyAsync(scratch, c->resolved_depth, pixels * 4, cudaMemcpyDeviceToHost, c->stream));
			CUDA_TRY(cudaStreamSynchronize(c->stream));
			for(int ty = 0; ty < c->ht; ++ty)
				if(c->part.owns_row(ty)) memcpy(depths + (size_t)ty * 8 * dev->W, scratch + (size_t)ty * 8 * dev->W, (size_t)8 * dev->W * 4);
		}
	}
	return wait ? group_present_wait(dev) : MLV_OK;
}

static int group_present_wait(mlv_device *dev) {
	for(mlv_device *c : *dev->children) {
		if(!c->readback_in_flight) continue;
		cudaSetDevice(c->cuda_dev);
		CUDA_TRY(cudaEventSynchronize(c->ev_readback_done));
		c->readback_in_flight = false;
	}
	return MLV_OK;
}

} // extern "C"
