// C ABI of librad_cuda.so (include/rad_cuda.h): context, memory, the staged S1..S6 calls, the
// device-resident shooting loop (CUDA graph), kernel benchmarks, and the NCCL plumbing of the
// batched multi-GPU mode.  Kernels live in raster.cu / process.cu / select_update.cu.
#include "rad_internal.cuh"
#include "tile_walk.cuh"
#include <vector>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <dlfcn.h>

static std::string g_create_err;

// ---- NCCL through dlopen: the same library instance torch already loaded when there is one -----
namespace {
struct NcclId { char internal[128]; };
typedef int (*fn_ncclGetUniqueId)(NcclId*);
typedef int (*fn_ncclCommInitRank)(void**, int, NcclId, int);
typedef int (*fn_ncclAllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_ncclCommDestroy)(void*);
typedef const char* (*fn_ncclGetErrorString)(int);
struct NcclApi {
	void* lib = nullptr;
	fn_ncclGetUniqueId GetUniqueId = nullptr; fn_ncclCommInitRank CommInitRank = nullptr;
	fn_ncclAllReduce AllReduce = nullptr; fn_ncclCommDestroy CommDestroy = nullptr; fn_ncclGetErrorString GetErrorString = nullptr;
	bool load(std::string& err) {
		if (lib) return true;
		const char* names[] = { "libnccl.so.2", "libnccl.so" };
		for (const char* n : names) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
		if (!lib) { err = std::string("dlopen(libnccl.so.2): ") + dlerror(); return false; }
		GetUniqueId = (fn_ncclGetUniqueId)dlsym(lib, "ncclGetUniqueId");
		CommInitRank = (fn_ncclCommInitRank)dlsym(lib, "ncclCommInitRank");
		AllReduce = (fn_ncclAllReduce)dlsym(lib, "ncclAllReduce");
		CommDestroy = (fn_ncclCommDestroy)dlsym(lib, "ncclCommDestroy");
		GetErrorString = (fn_ncclGetErrorString)dlsym(lib, "ncclGetErrorString");
		if (!GetUniqueId || !CommInitRank || !AllReduce) { err = "libnccl: missing symbols"; lib = nullptr; return false; }
		return true;
	}
};
NcclApi g_nccl;
const int kNcclFloat32 = 7, kNcclSum = 0;

template <typename T> cudaError_t dalloc(T*& p, size_t n) { return cudaMalloc((void**)&p, n * sizeof(T)); }
} // namespace

extern "C" {

const char* rad_version(void) { return "radiosity_b200 0.1 (sm_100a)"; }

const char* rad_last_error(const rad_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

uint32_t rad_patch_count(const rad_ctx* c) { return c ? c->d.P : 0; }
uint32_t rad_atlas_width(const rad_ctx* c) { return c ? c->d.W : 0; }
uint32_t rad_atlas_height(const rad_ctx* c) { return c ? c->d.H : 0; }

int rad_create(rad_ctx** out, const rad_config* cfg) {
	if (!out || !cfg) { g_create_err = "rad_create: null argument"; return RAD_E_ARG; }
	*out = nullptr;
	if (cfg->hemicube_side < 16 || cfg->hemicube_side > 2048 || cfg->hemicube_side % 16) { g_create_err = "rad_create: hemicube_side must be a multiple of 16 in [16, 2048]"; return RAD_E_ARG; }
	if (cfg->hemicubes < 1 || cfg->hemicubes > RAD_MAX_HEMICUBES) { g_create_err = "rad_create: hemicubes must be in [1, 512]"; return RAD_E_ARG; }
	if (cfg->hemicubes > 64 && cfg->select_mode == RAD_SELECT_REFERENCE) { g_create_err = "rad_create: the reference list selection supports at most 64 hemicubes (use RAD_SELECT_TOPK)"; return RAD_E_ARG; }
	if (cfg->max_patches > (1u << 23)) { g_create_err = "rad_create: max_patches must be <= 8388608"; return RAD_E_ARG; }
	if (cfg->max_patches < 1) { g_create_err = "rad_create: max_patches must be >= 1"; return RAD_E_ARG; }
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev == 0) { g_create_err = std::string("rad_create: no CUDA device (") + cudaGetErrorString(e) + ") — there is no CPU fallback"; return RAD_E_CUDA; }
	if (cfg->device < 0 || cfg->device >= ndev) { g_create_err = "rad_create: bad device ordinal"; return RAD_E_ARG; }
	if ((e = cudaSetDevice(cfg->device)) != cudaSuccess) { g_create_err = cudaGetErrorString(e); return RAD_E_CUDA; }

	rad_ctx* c = new rad_ctx();
	c->cfg = *cfg;
	c->have_ff = c->have_scene = c->emitters_ready = c->rendered = c->processed = c->keys_dirty = false;
	c->parity = 0; c->selkey_valid = false; c->cam_valid = false; c->have_nb = false;
	c->graph_exec = nullptr; c->graph_batches = 0; c->graph_keep_items = false;
	c->h_stage = nullptr; c->h_stage_bytes = 0; c->d_stage = nullptr; c->d_stage_bytes = 0; c->saved = nullptr; c->graph_launches = 0; c->graph_parity0 = 0; c->graph_stop = 0;
	c->rank = 0; c->world = 1; c->nccl_comm = nullptr; c->partition_only = false;
	c->peer_mode = false; c->xbuf = nullptr; c->xbuf_bytes = 0; for (int r = 0; r < RAD_MAX_PEERS; r++) c->peer_ptr[r] = nullptr;
	c->launches = 0; c->epoch = 254; c->multi_graph = true;
	if (const char* e = getenv("RAD_MULTI_GRAPH")) c->multi_graph = atoi(e) != 0;
	c->lanes = 8;             // concurrent raster lanes of the fused path (tuning knob RAD_LANES, 1 .. 8)
	if (const char* e = getenv("RAD_LANES")) { const int v = atoi(e); if (v >= 1 && v <= RAD_MAX_LANES) c->lanes = (uint32_t)v; }
	c->ev_fork = nullptr; for (int l = 0; l < RAD_MAX_LANES; l++) { c->lane_stream[l] = nullptr; c->ev_lane[l] = nullptr; }
	c->inline_area_forced = false; c->l2_group_mb = 1u << 20;   // default: the whole batch in one group (measured faster than L2-sized groups)
	if (const char* e = getenv("RAD_L2_GROUP_MB")) { const int v = atoi(e); if (v >= 1) c->l2_group_mb = (uint32_t)v; }   // tuning knob
	c->ring_failed = false; c->lane_delta_done = false;
	c->pdl = false;               // opt-in (RAD_PDL=1, k == 1 only): measured slower, 28.1 k against 30.1 k shots/s — the early CTAs of the next kernels take SM slots from the running one
	if (const char* e = getenv("RAD_PDL")) c->pdl = cfg->hemicubes == 1 && atoi(e) != 0;
	c->ring_mode = false; c->ring_sg = c->ring_rs = c->ring_proc_layers = 0; c->ring_ctas_per_sm = 0;
	if (const char* e = getenv("RAD_RING")) c->ring_mode = atoi(e) != 0;      // opt-in: L2-resident key ring (raster_ring_kernel; measured slower than the raster lanes)
	if (const char* e = getenv("RAD_RING_SG")) { const int v = atoi(e); if (v >= 1 && v <= 16) c->ring_sg = (uint32_t)v; }     // tuning knobs
	if (const char* e = getenv("RAD_RING_RS")) { const int v = atoi(e); if (v >= 2 && v <= 16) c->ring_rs = (uint32_t)v; }
	if (const char* e = getenv("RAD_RING_PROC")) { const int v = atoi(e); if (v >= 1 && v <= 7) c->ring_proc_layers = (uint32_t)v; }
	c->tile_mode = false; memset(&c->tl, 0, sizeof(c->tl));
	if (const char* e = getenv("RAD_RASTER")) c->tile_mode = strcmp(e, "tiles") == 0;   // opt-in: tile-binned rasteriser (raster_tiles.cu)
	c->spec = cfg->hemicubes == 1 && !(cfg->flags & RAD_FLAG_KEEP_ITEMBUFFER);
	if (const char* e = getenv("RAD_SPEC")) c->spec = c->spec && atoi(e) != 0;
	c->spec_slots = c->spec ? RAD_SPEC_SLOTS : 0u;
	c->spec_overlap = false;      // opt-in (RAD_SPEC_OVERLAP=1): the sequential kernel running beside the raster lanes is starved — 14.5 k against 88.4 k shots/s
	c->sel_base = c->sel_count = c->sel_excl = c->sel_excl_n = c->cam_base = c->cam_count = c->last_lanes = 0; c->defer_join = false;
	if (const char* e = getenv("RAD_SPEC_OVERLAP")) c->spec_overlap = c->spec && atoi(e) != 0;
	if (const char* e = getenv("RAD_SPEC_SLOTS")) { const int v = atoi(e); if (c->spec && v >= 8 && v <= RAD_SPEC_SLOTS) c->spec_slots = (uint32_t)v; }   // tuning knob
	c->spec_graph = nullptr; c->spec_graph_batches = 0; c->spec_graph_launches = 0; c->spec_graph_stop = 0; c->spec_graph_epoch_after = 0; c->spec_blocks = 0; c->select_override = 0;
	const uint32_t kslots = c->spec ? (c->spec_overlap ? (uint32_t)RAD_SPEC_POOL : c->spec_slots) : cfg->hemicubes;        // hemicube slots the buffers hold
	c->key_slots = kslots;
	const uint32_t kbufs = kslots > 64u && c->spec ? 64u : kslots;             // key buffers: a speculative context renders 64 slots at a time into buffers 0 .. 63
	c->key_bufs = kbufs;
	RadDev& D = c->d;
	memset(&D, 0, sizeof(D));
	D.N = cfg->hemicube_side; D.W = 2 * D.N; D.H = D.N + D.N / 2; D.RES = D.W * D.H; D.k = cfg->hemicubes;
	D.P = 0; D.h0 = 0; D.h1 = D.k;
	D.reflectivity = cfg->reflectivity;
	// work lists sized for the worst case the grouping in raster.cu plans with — 2 records and 5 (patch, face) pairs per
	// patch and hemicube, for the (at most 64) hemicubes one launch group renders; the real load is a fraction of it.  This
	// is tens of GB for a 1 M-patch scene: HBM capacity is what a B200 has to spare, and a whole batch per launch group is
	// worth it.  Bounded by a third of the free memory (the raster then falls back to smaller hemicube groups).
	{
		const uint64_t kk = kslots < 64 ? kslots : 64;
		size_t free_b = 0, total_b = 0;
		cudaMemGetInfo(&free_b, &total_b);
		const uint64_t budget = (uint64_t)free_b / 3;                         // bytes for the four lists
		uint64_t want = 2ull * cfg->max_patches * kk, wantp = 5ull * cfg->max_patches * kk;
		if (want < (1u << 20)) want = 1u << 20;
		if (wantp < (1u << 20)) wantp = 1u << 20;
		const uint64_t bytes = want * (sizeof(RadBigTri) + sizeof(RadSmallQuad) + 2 * sizeof(RadQueueEntry)) + wantp * 4;
		if (bytes > budget && budget > 0) { const double f = (double)budget / (double)bytes; want = (uint64_t)(want * f); wantp = (uint64_t)(wantp * f); }
		if (want > (1ull << 30)) want = 1ull << 30;
		if (wantp > (1ull << 31)) wantp = 1ull << 31;
		D.q_tri_cap = (uint32_t)want; D.q_sm_cap = (uint32_t)want; D.q_ent_cap = (uint32_t)(2 * want);
		D.pairs_cap = (uint32_t)wantp;
	}
	D.kbase = 0; D.inline_area = 64;
	D.small_steps = D.k == 1 ? 16 : RAD_SMALL_STEPS; D.tile = D.k == 1 ? 16 : RAD_TILE;
	if (const char* e = getenv("RAD_SMALL_STEPS")) { const int v = atoi(e); if (v >= 1 && v <= 64) D.small_steps = (uint32_t)v; }   // tuning knobs
	if (const char* e = getenv("RAD_TILE")) { const int v = atoi(e); if (v == 8 || v == 16 || v == 32 || v == 64) D.tile = (uint32_t)v; }
	if (const char* e = getenv("RAD_INLINE_AREA")) { const int v = atoi(e); if (v >= 1 && v <= 4096) { D.inline_area = (uint32_t)v; c->inline_area_forced = true; } }   // tuning knob
	const size_t Pm = cfg->max_patches;
	float4 *v0, *v1, *v2; float *color, *ff, *proj;
	bool ok = true;
	#define A(expr) do { if (ok && (expr) != cudaSuccess) ok = false; } while (0)
	A(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
	A(cudaEventCreate(&c->ev0)); A(cudaEventCreate(&c->ev1));
	A(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
	for (int l = 0; l < RAD_MAX_LANES; l++) { A(cudaStreamCreateWithFlags(&c->lane_stream[l], cudaStreamNonBlocking)); A(cudaEventCreateWithFlags(&c->ev_lane[l], cudaEventDisableTiming)); }
	A(dalloc(v0, Pm)); A(dalloc(v1, Pm)); A(dalloc(v2, Pm));
	A(dalloc(color, 3 * Pm)); A(dalloc(D.rad, 3 * Pm)); A(dalloc(D.illum, 3 * Pm));
	A(dalloc(ff, (size_t)D.RES));
	A(dalloc(D.keys, (size_t)kbufs * D.RES)); A(dalloc(D.items, (size_t)D.k * D.RES));
	A(dalloc(D.F, (size_t)kslots * Pm)); A(dalloc(D.dB, 3 * Pm));
	A(dalloc(D.mvp, (size_t)kslots * RAD_NFACES * 16)); A(dalloc(D.em, (size_t)kslots)); A(dalloc(D.emlite, 2 * (size_t)kslots)); A(dalloc(D.ctl, 1)); A(dalloc(D.rc, 1));
	A(dalloc(D.spec_cand, (size_t)3 * 256)); A(dalloc(D.spec_steps, (size_t)RAD_SPEC_SLOTS));
	A(dalloc(D.q_tri, (size_t)D.q_tri_cap)); A(dalloc(D.q_ent, (size_t)D.q_ent_cap)); A(dalloc(D.q_sm, (size_t)D.q_sm_cap)); A(dalloc(D.pairs, (size_t)D.pairs_cap)); A(dalloc(D.nb, 8 * Pm)); A(dalloc(D.shade_e, 3 * Pm));
	A(dalloc(D.ework, Pm < RAD_MAX_HEMICUBES ? (size_t)RAD_MAX_HEMICUBES : Pm)); A(dalloc(D.cand0, ((Pm + 2047) / 2048) * (size_t)RAD_MAX_HEMICUBES)); A(dalloc(D.cand1, ((Pm + 2047) / 2048) * (size_t)RAD_MAX_HEMICUBES)); A(dalloc(proj, 16));
	if (c->tile_mode) {
		RadTiles& T = c->tl;
		T.tx = (D.W + RAD_TILE_W - 1) / RAD_TILE_W; T.ty = (D.H + RAD_TILE_H - 1) / RAD_TILE_H; T.T = T.tx * T.ty;
		uint64_t cap = 2ull * ((uint64_t)D.q_sm_cap + D.q_tri_cap);    // a record lies in 1.3 tiles on average
		if (cap > (1ull << 31)) cap = 1ull << 31;
		T.refs_cap = (uint32_t)cap;
		const size_t nl = (size_t)D.k * T.T * 2u;
		A(dalloc(T.cnt, nl)); A(dalloc(T.base, nl + RAD_MAX_LANES + 1)); A(dalloc(T.refs, (size_t)T.refs_cap));
		if (ok) A(cudaMemset(T.cnt, 0, nl * 4));
	}
	#undef A
	if (!ok) {
		g_create_err = std::string("rad_create: allocation failed: ") + cudaGetErrorString(cudaGetLastError());
		delete c; return RAD_E_CUDA;
	}
	D.v0 = v0; D.v1 = v1; D.v2 = v2; D.color = color; D.ff = ff; D.proj = proj;
	D.qc = &D.ctl->lane[0];
	cudaMemcpyAsync(proj, cfg->projection, 64, cudaMemcpyHostToDevice, c->stream);
	cudaMemsetAsync(D.keys, 0xFF, (size_t)kbufs * D.RES * 8, c->stream);
	cudaMemsetAsync(D.items, 0, (size_t)D.k * D.RES * 4, c->stream);
	cudaMemsetAsync(D.F, 0, (size_t)kslots * Pm * 4, c->stream);
	cudaMemsetAsync(D.dB, 0, 3 * Pm * 4, c->stream);
	cudaMemsetAsync(D.em, 0, (size_t)kslots * sizeof(RadEmitter), c->stream);
	cudaMemsetAsync(D.emlite, 0, 2 * (size_t)kslots * sizeof(float4), c->stream);
	cudaMemsetAsync(D.ctl, 0, sizeof(RadControl), c->stream);
	cudaMemsetAsync(D.rc, 0, sizeof(RadRingCtl), c->stream);
	if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess) { g_create_err = cudaGetErrorString(e); delete c; return RAD_E_CUDA; }
	*out = c;
	return RAD_OK;
}

static void drop_graph(rad_ctx* c) {
	if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; c->graph_batches = 0; }
	if (c->spec_graph) { cudaGraphExecDestroy(c->spec_graph); c->spec_graph = nullptr; c->spec_graph_batches = 0; }
}

int rad_destroy(rad_ctx* c) {
	if (!c) return RAD_OK;
	cudaSetDevice(c->cfg.device);
	cudaStreamSynchronize(c->stream);
	drop_graph(c);
	if (c->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->nccl_comm);
	for (int r = 0; r < RAD_MAX_PEERS; r++) if (c->peer_ptr[r]) cudaIpcCloseMemHandle(c->peer_ptr[r]);
	if (c->xbuf) cudaFree(c->xbuf);
	RadDev& D = c->d;
	cudaFree((void*)D.v0); cudaFree((void*)D.v1); cudaFree((void*)D.v2); cudaFree((void*)D.color);
	cudaFree(D.rad); cudaFree(D.illum); cudaFree((void*)D.ff); cudaFree(D.keys); cudaFree(D.items);
	cudaFree(D.F); cudaFree(D.dB); cudaFree(D.mvp); cudaFree(D.em); cudaFree(D.emlite); cudaFree(D.ctl); cudaFree(D.rc); cudaFree(D.spec_cand); cudaFree(D.spec_steps);
	cudaFree(D.q_tri); cudaFree(D.q_ent); cudaFree(D.q_sm); cudaFree(D.pairs); cudaFree(D.nb); cudaFree(D.shade_e); cudaFree(D.ework); cudaFree(D.cand0); cudaFree(D.cand1); cudaFree((void*)D.proj);
	if (c->tl.cnt) cudaFree(c->tl.cnt);
	if (c->tl.base) cudaFree(c->tl.base);
	if (c->tl.refs) cudaFree(c->tl.refs);
	if (c->h_stage) cudaFreeHost(c->h_stage);
	if (c->saved) cudaFree(c->saved);
	if (c->ev_fork) cudaEventDestroy(c->ev_fork);
	for (int l = 0; l < RAD_MAX_LANES; l++) { if (c->ev_lane[l]) cudaEventDestroy(c->ev_lane[l]); if (c->lane_stream[l]) cudaStreamDestroy(c->lane_stream[l]); }
	if (c->d_stage) cudaFree(c->d_stage);
	cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1);
	cudaStreamDestroy(c->stream);
	delete c;
	return RAD_OK;
}

static int stage(rad_ctx* c, size_t bytes) {
	if (c->h_stage_bytes >= bytes) return RAD_OK;
	if (c->h_stage) cudaFreeHost(c->h_stage);
	c->h_stage = nullptr; c->h_stage_bytes = 0;
	RAD_CUDA_TRY(c, cudaMallocHost((void**)&c->h_stage, bytes));
	c->h_stage_bytes = bytes;
	return RAD_OK;
}

int rad_set_formfactors(rad_ctx* c, const float* ff, uint32_t n) {
	if (!c || !ff) return RAD_E_ARG;
	cudaSetDevice(c->cfg.device);
	if (n != c->d.RES) { c->err = "rad_set_formfactors: n must be 3*N*N (one hemicube)"; return RAD_E_ARG; }
	RAD_CUDA_TRY(c, cudaMemcpyAsync((void*)c->d.ff, ff, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
	RAD_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
	c->have_ff = true;
	return RAD_OK;
}

// Host arrays travel in the reference's layouts (AoS float[P*3], float[P*12]); the conversion to / from the device
// layouts (three planes, three float4 streams) runs on the GPU (layout.cu).  Every call packs its arrays back to back
// into the pinned staging buffer, moves them with ONE copy and synchronises ONCE.
static int stage_dev(rad_ctx* c, size_t bytes) {
	int r = stage(c, bytes); if (r) return r;
	if (c->d_stage_bytes >= bytes) return RAD_OK;
	if (c->d_stage) cudaFree(c->d_stage);
	c->d_stage = nullptr; c->d_stage_bytes = 0;
	RAD_CUDA_TRY(c, cudaMalloc((void**)&c->d_stage, bytes));
	c->d_stage_bytes = bytes;
	return RAD_OK;
}
// single-array forms (multi-GPU dB exchange)
static int h2d_aos3(rad_ctx* c, const float* src, float* dst_planes, size_t P) {
	memcpy(c->h_stage, src, 3 * P * 4);
	RAD_CUDA_TRY(c, cudaMemcpyAsync(c->d_stage, c->h_stage, 3 * P * 4, cudaMemcpyHostToDevice, c->stream));
	rad_launch_aos3_to_planes(c, c->d_stage, dst_planes, (uint32_t)P);
	RAD_CUDA_TRY(c, cudaStreamSynchronize(c->stream));       // the staging buffers are reused by the next call
	return RAD_OK;
}
static int d2h_aos3(rad_ctx* c, const float* src_planes, float* dst, size_t P) {
	rad_launch_planes_to_aos3(c, src_planes, c->d_stage, (uint32_t)P);
	RAD_CUDA_TRY(c, cudaMemcpyAsync(c->h_stage, c->d_stage, 3 * P * 4, cudaMemcpyDeviceToHost, c->stream));
	RAD_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
	memcpy(dst, c->h_stage, 3 * P * 4);
	return RAD_OK;
}

// Page-locked caller memory (cudaHostAlloc / cudaHostRegister / torch pin_memory) is copied from / to directly; pageable
// memory goes through the context's pinned staging buffer first.
static bool is_pinned(const void* p) {
	cudaPointerAttributes a;
	if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
	return a.type == cudaMemoryTypeHost;
}
// host floats -> device staging floats [off, off + n): enqueue only
static int enqueue_h2d(rad_ctx* c, const float* src, size_t off_floats, size_t n) {
	const float* from = src;
	if (!is_pinned(src)) { memcpy(c->h_stage + off_floats, src, n * 4); from = c->h_stage + off_floats; }
	RAD_CUDA_TRY(c, cudaMemcpyAsync(c->d_stage + off_floats, from, n * 4, cudaMemcpyHostToDevice, c->stream));
	return RAD_OK;
}

// enqueue only (no synchronisation): B and I through staging floats [off, off + 6P)
static int enqueue_state_upload(rad_ctx* c, const float* rad3, const float* illum3, size_t off_floats) {
	const size_t P = c->d.P;
	int r;
	if ((r = enqueue_h2d(c, rad3, off_floats, 3 * P))) return r;
	if ((r = enqueue_h2d(c, illum3, off_floats + 3 * P, 3 * P))) return r;
	rad_launch_aos3_to_planes(c, c->d_stage + off_floats, c->d.rad, (uint32_t)P);
	rad_launch_aos3_to_planes(c, c->d_stage + off_floats + 3 * P, c->d.illum, (uint32_t)P);
	RAD_CUDA_TRY(c, cudaMemsetAsync(c->d.ctl, 0, sizeof(RadControl), c->stream));
	c->selkey_valid = false; c->emitters_ready = c->rendered = c->processed = false;
	return RAD_OK;
}

int rad_upload_state(rad_ctx* c, const float* rad3, const float* illum3) {
	if (!c || !rad3 || !illum3) return RAD_E_ARG;
	if (!c->have_scene) { c->err = "rad_upload_state: no scene"; return RAD_E_STATE; }
	cudaSetDevice(c->cfg.device);
	const size_t P = c->d.P;
	int r = stage_dev(c, 21 * P * 4); if (r) return r;
	if ((r = enqueue_state_upload(c, rad3, illum3, 0))) return r;
	RAD_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
	return RAD_OK;
}

int rad_upload_scene(rad_ctx* c, const float* verts12, const float* color3, const float* rad3, const float* illum3, uint32_t P) {
	if (!c || !verts12 || !color3 || !rad3 || !illum3) return RAD_E_ARG;
	if (P < 1 || P > c->cfg.max_patches) { c->err = "rad_upload_scene: P out of range (max_patches)"; return RAD_E_ARG; }
	cudaSetDevice(c->cfg.device);
	if (P != c->d.P) { drop_graph(c); c->have_nb = false; }   // the captured launches carry P (nothing else of the scene); the neighbour planes are strided by P
	int r = stage_dev(c, (size_t)P * 21 * 4); if (r) return r;
	// staging layout (floats): [0, 12P) quads | [12P, 15P) colour | [15P, 21P) B, I
	// 48-byte quad records -> three float4 streams (coalesced 16 B loads per lane in the rasteriser)
	if ((r = enqueue_h2d(c, verts12, 0, (size_t)P * 12))) return r;
	if ((r = enqueue_h2d(c, color3, (size_t)P * 12, (size_t)P * 3))) return r;
	rad_launch_split_quads(c, c->d_stage, P);
	rad_launch_aos3_to_planes(c, c->d_stage + (size_t)P * 12, (float*)c->d.color, P);
	RAD_CUDA_TRY(c, cudaMemsetAsync(c->d.F, 0, (size_t)c->d.k * P * 4, c->stream));
	c->d.P = P;
	c->have_scene = true;
	if ((r = enqueue_state_upload(c, rad3, illum3, (size_t)P * 15))) return r;
	RAD_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
	return RAD_OK;
}

int rad_download_state(rad_ctx* c, float* rad3, float* illum3) {
	if (!c || !rad3 || !illum3) return RAD_E_ARG;
	if (!c->have_scene) { c->err = "rad_download_state: no scene"; return RAD_E_STATE; }
	cudaSetDevice(c->cfg.device);
	const size_t P = c->d.P;
	int r = stage_dev(c, 21 * P * 4); if (r) return r;
	rad_launch_planes_to_aos3(c, c->d.rad, c->d_stage, (uint32_t)P);
	rad_launch_planes_to_aos3(c, c->d.illum, c->d_stage + 3 * P, (uint32_t)P);
	const bool pr = is_pinned(rad3), pi = is_pinned(illum3);
	RAD_CUDA_TRY(c, cudaMemcpyAsync(pr ? rad3 : c->h_stage, c->d_stage, 3 * P * 4, cudaMemcpyDeviceToHost, c->stream));
	RAD_CUDA_TRY(c, cudaMemcpyAsync(pi ? illum3 : c->h_stage + 3 * P, c->d_stage + 3 * P, 3 * P * 4, cudaMemcpyDeviceToHost, c->stream));
	RAD_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
	if (!pr) memcpy(rad3, c->h_stage, 3 * P * 4);
	if (!pi) memcpy(illum3, c->h_stage + 3 * P, 3 * P * 4);
	return RAD_OK;
}

static int need_ready(rad_ctx* c, const char* who) {
	if (!c) return RAD_E_ARG;
	if (!c->have_scene || !c->have_ff) { c->err = std::string(who) + ": upload the scene and the form factors first"; return RAD_E_STATE; }
	cudaSetDevice(c->cfg.device);
	return RAD_OK;
}
static int sync_check(rad_ctx* c) {
	RAD_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
	RAD_CUDA_TRY(c, cudaGetLastError());
	return RAD_OK;
}

static int read_emitters(rad_ctx* c, uint32_t* ids_out, uint32_t* valid_out) {
	std::vector<RadEmitter> em(c->d.k);
	RAD_CUDA_TRY(c, cudaMemcpyAsync(em.data(), c->d.em, em.size() * sizeof(RadEmitter), cudaMemcpyDeviceToHost, c->stream));
	RAD_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
	for (uint32_t h = 0; h < c->d.k; h++) { if (ids_out) ids_out[h] = em[h].id; if (valid_out) valid_out[h] = em[h].valid; }
	return RAD_OK;
}

int rad_select(rad_ctx* c, uint32_t* ids_out, uint32_t* valid_out) {
	int r = need_ready(c, "rad_select"); if (r) return r;
	rad_launch_select(c);
	if ((r = sync_check(c))) return r;
	c->emitters_ready = true; c->rendered = c->processed = false;
	if (ids_out || valid_out) return read_emitters(c, ids_out, valid_out);
	return RAD_OK;
}

int rad_set_emitters(rad_ctx* c, const uint32_t* ids, uint32_t n) {
	int r = need_ready(c, "rad_set_emitters"); if (r) return r;
	if (!ids || n > c->d.k) { c->err = "rad_set_emitters: n must be <= hemicubes"; return RAD_E_ARG; }
	RAD_CUDA_TRY(c, cudaMemcpyAsync(c->d.ework, ids, n * 4, cudaMemcpyHostToDevice, c->stream));   // ework doubles as staging
	rad_launch_set_emitters(c, c->d.ework, n);
	if ((r = sync_check(c))) return r;
	c->emitters_ready = true; c->rendered = c->processed = false;
	c->selkey_valid = false;
	return RAD_OK;
}

int rad_render_hemicubes(rad_ctx* c) {
	int r = need_ready(c, "rad_render_hemicubes"); if (r) return r;
	if (!c->emitters_ready) { c->err = "rad_render_hemicubes: call rad_select / rad_set_emitters first"; return RAD_E_STATE; }
	RAD_CUDA_TRY(c, cudaMemsetAsync(c->d.items, 0, (size_t)c->d.k * c->d.RES * 4, c->stream));   // glClear: NULL emitters stay black
	rad_launch_raster(c);
	rad_launch_resolve(c, /*reset=*/false);       // keys stay readable for rad_read_depthbuffer
	if ((r = sync_check(c))) return r;
	c->rendered = true; c->processed = false;
	return RAD_OK;
}

int rad_process_hemicubes(rad_ctx* c) {
	int r = need_ready(c, "rad_process_hemicubes"); if (r) return r;
	if (!c->emitters_ready) { c->err = "rad_process_hemicubes: no emitters"; return RAD_E_STATE; }
	RAD_CUDA_TRY(c, cudaMemsetAsync(c->d.F, 0, (size_t)c->d.k * c->d.P * 4, c->stream));
	rad_launch_process(c);
	if ((r = sync_check(c))) return r;
	c->processed = true;
	return RAD_OK;
}

static int read_ctl(rad_ctx* c, RadControl* out) {
	RAD_CUDA_TRY(c, cudaMemcpyAsync(out, c->d.ctl, sizeof(RadControl), cudaMemcpyDeviceToHost, c->stream));
	RAD_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
	return RAD_OK;
}

int rad_apply(rad_ctx* c, float* last_energy_len) {
	int r = need_ready(c, "rad_apply"); if (r) return r;
	if (!c->processed) { c->err = "rad_apply: call rad_process_hemicubes first"; return RAD_E_STATE; }
	rad_launch_apply(c, false);
	if ((r = sync_check(c))) return r;
	c->processed = c->rendered = c->emitters_ready = false;
	c->selkey_valid = false;
	RadControl ctl; if ((r = read_ctl(c, &ctl))) return r;
	if (last_energy_len) *last_energy_len = ctl.last_energy_len;
	return RAD_OK;
}

// one steady-state batch on the stream (single GPU): select -> camera -> raster -> fused resolve+process -> apply
static void enqueue_batch(rad_ctx* c, bool keep_items) {
	rad_launch_select(c);
	rad_launch_raster_process(c, keep_items);
	const bool fuse = c->d.k == 1;
	rad_launch_apply(c, fuse);
	if (fuse) { c->parity ^= 1; c->selkey_valid = true; c->cam_valid = true; }
}

static int enqueue_batch_multi(rad_ctx* c, bool keep_items) {
	rad_launch_select(c);
	c->lane_delta_done = false;
	rad_launch_raster_process(c, keep_items);
	if (!c->lane_delta_done) rad_launch_delta(c);             // (paths without raster lanes: one kernel for the rank's slots)
	if (c->peer_mode && c->d.xtwo) rad_launch_xreduce(c);
	if (c->nccl_comm && !c->peer_mode) {
		int rc = g_nccl.AllReduce(c->d.dB, c->d.dB, (size_t)3 * c->d.P, kNcclFloat32, kNcclSum, c->nccl_comm, c->stream);
		if (rc != 0) { c->err = std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error"); return RAD_E_NCCL; }
	}
	rad_launch_finish(c, false);
	c->selkey_valid = false;
	return RAD_OK;
}

// ---- speculative strict progressive refinement (k == 1; rad_ctx::spec) ------------------------------------------------------
// The launchers read c->d: the batch view (k = the slots the context holds) is swapped in for the duration of an enqueue.
struct SpecView {
	rad_ctx* c; RadDev saved;
	explicit SpecView(rad_ctx* c_) : c(c_), saved(c_->d) {
		RadDev& S = c->d;
		S.k = c->key_slots; S.spec = 1u; S.stop_gate = 1u; S.deal = 0u; S.small_steps = RAD_SMALL_STEPS; S.tile = RAD_TILE;
	}
	~SpecView() { c->d = saved; c->select_override = 0; c->cam_count = 0; c->sel_count = 0; c->sel_excl_n = 0; c->defer_join = false; }
};
// render ahead: the `n` strongest patches (the argmax's tie rule; patches waiting in em[excl .. +excl_n) left out) go to
// em[base .. +n), cameras, hemicubes + ProcessHemicube through the raster lanes.  defer: the lanes are left running
// (rad_join_lanes) so that the sequential kernel of the OTHER half overlaps them.
static void enqueue_spec_render(rad_ctx* c, uint32_t base, uint32_t n, uint32_t excl, uint32_t excl_n, bool defer) {
	RadDev& S = c->d;
	c->select_override = 2; c->sel_base = base; c->sel_count = n; c->sel_excl = excl; c->sel_excl_n = excl_n; c->cam_base = base; c->cam_count = n;
	S.h0 = base; S.h1 = base + n;
	cudaMemsetAsync(S.F + (size_t)base * S.P, 0, (size_t)n * S.P * 4, c->stream);   // (slots a cut-short batch did not use are not zeroed by anybody else)
	rad_launch_select(c);
	c->defer_join = defer;
	rad_launch_raster_process(c, false);
	c->defer_join = false;
}
static int enqueue_spec_apply(rad_ctx* c, uint32_t base, uint32_t n, int stop_armed) {
	cudaMemsetAsync(&c->d.ctl->spec_key[0], 0, 3 * sizeof(unsigned long long), c->stream);
	return rad_launch_spec_apply(c, c->d, base, n, stop_armed);
}
// one step.  Plain: select + render n slots, then replay the loop over them.  Overlapped (spec_overlap, two halves of
// RAD_SPEC_SLOTS): half a = (t & 1) was rendered during the previous step; the other half is selected from the state at hand
// (without a's patches: they are about to shoot) and rendered WHILE the sequential kernel replays the loop over a.
static int enqueue_spec_step(rad_ctx* c, uint32_t t, uint32_t n, int stop_armed) {
	SpecView view(c);
	if (!c->spec_overlap) {
		enqueue_spec_render(c, 0, n, 0, 0, false);
		return enqueue_spec_apply(c, 0, n, stop_armed);
	}
	const uint32_t H = RAD_SPEC_SLOTS, a = (t & 1u) * H, g = H - a;
	enqueue_spec_render(c, g, H, a, H, true);
	const int r = enqueue_spec_apply(c, a, H, stop_armed);
	rad_join_lanes(c);
	return r;
}

static int shoot_speculative(rad_ctx* c, uint32_t n_shots, int stop_test, const RadControl& ctl0, uint64_t* launches_out, RadControl* ctl_out) {
	int r;
	const uint32_t target = ctl0.shots_done + n_shots;
	RAD_CUDA_TRY(c, cudaMemcpyAsync(&c->d.ctl->spec_target, &target, sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
	RAD_CUDA_TRY(c, cudaMemsetAsync(&c->d.ctl->spec_done, 0, sizeof(uint32_t), c->stream));
	RAD_CUDA_TRY(c, cudaMemsetAsync(&c->d.ctl->gate, 0, sizeof(uint32_t), c->stream));
	RAD_CUDA_TRY(c, cudaStreamSynchronize(c->stream));                            // (target is a stack variable)
	uint64_t launches = 0;
	uint32_t done = 0, t = 0;
	RadControl ct = ctl0;
	const uint32_t GB = 4, M = c->spec_slots;
	if (c->spec_overlap) {                                                        // prologue: half 0 is rendered before the first step
		const uint32_t l0 = c->launches;
		{ SpecView view(c); enqueue_spec_render(c, 0, RAD_SPEC_SLOTS, 0, 0, false); }
		launches += c->launches - l0;
	}
	while (done < n_shots) {
		const uint32_t remaining = n_shots - done;
		if (remaining >= GB * M && (t & 1u) == 0u) {
			// a CUDA graph of GB full steps; steps behind the end of the call are gated on the device
			if (!c->spec_graph || c->spec_graph_stop != (uint32_t)stop_test) {
				if (c->spec_graph) { cudaGraphExecDestroy(c->spec_graph); c->spec_graph = nullptr; }
				cudaGraph_t g = nullptr;
				const uint32_t l0 = c->launches;
				RAD_CUDA_TRY(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
				rad_launch_clear_keys(c);                                        // a replay re-uses the captured epoch tags: every replay starts from cleared keys
				for (uint32_t b = 0; b < GB; b++) if ((r = enqueue_spec_step(c, b, M, stop_test))) { cudaGraph_t junk; cudaStreamEndCapture(c->stream, &junk); return r; }
				c->spec_graph_epoch_after = c->epoch;
				RAD_CUDA_TRY(c, cudaStreamEndCapture(c->stream, &g));
				RAD_CUDA_TRY(c, cudaGraphInstantiate(&c->spec_graph, g, 0));
				cudaGraphDestroy(g);
				c->spec_graph_batches = GB; c->spec_graph_stop = (uint32_t)stop_test;
				c->spec_graph_launches = c->launches - l0; c->launches = l0;
			}
			RAD_CUDA_TRY(c, cudaGraphLaunch(c->spec_graph, c->stream));
			launches += c->spec_graph_launches;
			c->epoch = c->spec_graph_epoch_after;
			t += GB;
		} else {
			uint32_t m = 8; while (m < remaining && m < M) m <<= 1;                 // (plain form: a short tail renders only what it may need)
			if (m > M) m = M;
			const uint32_t l0 = c->launches;
			if ((r = enqueue_spec_step(c, t, m, stop_test))) return r;
			launches += c->launches - l0;
			t++;
		}
		if ((r = read_ctl(c, &ct))) return r;
		done = ct.shots_done - ctl0.shots_done;
		if (ct.spec_done) break;
	}
	c->selkey_valid = false; c->cam_valid = false;
	*launches_out = launches; *ctl_out = ct;
	return RAD_OK;
}

int rad_shoot(rad_ctx* c, uint32_t n_batches, int stop_test, rad_stats* out) {
	int r = need_ready(c, "rad_shoot"); if (r) return r;
	const bool keep = (c->cfg.flags & RAD_FLAG_KEEP_ITEMBUFFER) != 0;
	RadControl ctl0; if ((r = read_ctl(c, &ctl0))) return r;
	c->launches = 0;
	uint64_t launches = 0;
	// stop test: every kernel that changes state checks a device-side gate, so a CUDA-graph replay ends with the batch
	// whose |lastEnergy| < 0.1 fired, exactly like the reference's loop (Main.cpp:1137,1297-1300); the call starts armed and clean
	c->d.stop_gate = stop_test ? 1u : 0u;
	struct Disarm { rad_ctx* c; ~Disarm() { c->d.stop_gate = 0; } } disarm{ c };
	if (stop_test) RAD_CUDA_TRY(c, cudaMemsetAsync(&c->d.ctl->stopped, 0, sizeof(uint32_t), c->stream));
	if (stop_test) RAD_CUDA_TRY(c, cudaMemsetAsync(&c->d.ctl->gate, 0, sizeof(uint32_t), c->stream));
	if (c->keys_dirty) rad_launch_clear_keys(c);
	RAD_CUDA_TRY(c, cudaEventRecord(c->ev0, c->stream));
	uint32_t done = 0; bool stopped = false;
	if (c->world > 1 && !c->nccl_comm && !c->peer_mode) {
		c->err = "rad_shoot: this context only holds a partition (rad_set_partition) and no exchange (rad_comm_init / rad_peer_init): use rad_batch_partial / rad_read_delta / rad_write_delta / rad_batch_finish";
		return RAD_E_STATE;
	}
	if (c->world > 1) {
		// sharded batches: the same CUDA-graph replay as on one GPU, the ncclAllReduce of every batch captured inside it
		// (RAD_MULTI_GRAPH=0 falls back to direct launches)
		const uint32_t GB = 8;
		if (c->multi_graph && (c->nccl_comm || c->peer_mode) && n_batches >= GB) {
			if (!c->graph_exec || c->graph_batches != GB || c->graph_keep_items != keep || c->graph_stop != c->d.stop_gate) {
				drop_graph(c);
				cudaGraph_t g = nullptr;
				const uint32_t l0 = c->launches;
				c->graph_stop = c->d.stop_gate;
				RAD_CUDA_TRY(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
				rad_launch_clear_keys(c);
				for (uint32_t b = 0; b < GB; b++) if ((r = enqueue_batch_multi(c, keep))) { cudaGraph_t junk; cudaStreamEndCapture(c->stream, &junk); return r; }
				c->graph_epoch_after = c->epoch;
				RAD_CUDA_TRY(c, cudaStreamEndCapture(c->stream, &g));
				RAD_CUDA_TRY(c, cudaGraphInstantiate(&c->graph_exec, g, 0));
				cudaGraphDestroy(g);
				c->graph_batches = GB; c->graph_keep_items = keep;
				c->graph_launches = c->launches - l0; c->launches = l0;
				c->graph_parity0 = c->parity;
			}
			while (n_batches - done >= GB && !stopped) {
				RAD_CUDA_TRY(c, cudaGraphLaunch(c->graph_exec, c->stream));
				launches += c->graph_launches;
				done += GB;
				c->epoch = c->graph_epoch_after;
				if (stop_test) { RadControl t; if ((r = read_ctl(c, &t))) return r; stopped = t.stopped != 0; }
			}
		}
		for (; done < n_batches && !stopped; done++) {
			if ((r = enqueue_batch_multi(c, keep))) return r;
			if (stop_test) { RadControl t; if ((r = read_ctl(c, &t))) return r; stopped = t.stopped != 0; }
		}
		launches += c->launches;
	} else if (c->spec && c->d.k == 1 && !keep) {
		RadControl t;
		if ((r = shoot_speculative(c, n_batches, stop_test, ctl0, &launches, &t))) return r;
		c->launches = 0;
	} else {
		if (c->d.k == 1 && !c->selkey_valid) rad_launch_argmax(c);
		// steady state: a CUDA graph of GB batches (even, so that the k==1 key ping-pong returns to its start)
		const uint32_t GB = 16;
		if (n_batches >= GB) {
			if (!c->graph_exec || c->graph_batches != GB || c->graph_keep_items != keep || c->graph_parity0 != c->parity || c->graph_stop != c->d.stop_gate) {
				drop_graph(c);
				c->graph_stop = c->d.stop_gate;
				cudaGraph_t g = nullptr;
				const uint32_t parity0 = c->parity; const bool sk0 = c->selkey_valid, cam0 = c->cam_valid; const uint32_t l0 = c->launches;
				c->cam_valid = false;              // a replay starts with the camera kernel (it cannot know what ran before it)
				RAD_CUDA_TRY(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
				rad_launch_clear_keys(c);          // a replay re-uses the captured epoch tags: start every replay from cleared keys
				for (uint32_t b = 0; b < GB; b++) enqueue_batch(c, keep);
				c->graph_epoch_after = c->epoch;
				RAD_CUDA_TRY(c, cudaStreamEndCapture(c->stream, &g));
				RAD_CUDA_TRY(c, cudaGraphInstantiate(&c->graph_exec, g, 0));
				cudaGraphDestroy(g);
				c->graph_batches = GB; c->graph_keep_items = keep;
				c->graph_launches = c->launches - l0;
				c->parity = parity0; c->selkey_valid = sk0; c->cam_valid = cam0; c->launches = l0;   // capture did not execute anything
				c->graph_parity0 = parity0;
			}
			while (n_batches - done >= GB && !stopped && c->parity == c->graph_parity0) {
				RAD_CUDA_TRY(c, cudaGraphLaunch(c->graph_exec, c->stream));
				launches += c->graph_launches;
				done += GB;
				c->epoch = c->graph_epoch_after;
				if (c->d.k == 1) { c->selkey_valid = true; c->cam_valid = true; }
				if (stop_test) { RadControl t; if ((r = read_ctl(c, &t))) return r; stopped = t.stopped != 0; }
			}
		}
		for (; done < n_batches && !stopped; done++) {
			enqueue_batch(c, keep);
			if (stop_test) { RadControl t; if ((r = read_ctl(c, &t))) return r; stopped = t.stopped != 0; }
		}
		launches += c->launches;
	}
	RAD_CUDA_TRY(c, cudaEventRecord(c->ev1, c->stream));
	if ((r = sync_check(c))) return r;
	c->emitters_ready = c->rendered = c->processed = false;
	RadControl ctl; if ((r = read_ctl(c, &ctl))) return r;
	if (stop_test && ctl.stopped) { c->selkey_valid = false; c->cam_valid = false; }   // gated batches did not prepare the next shooter
	if (out) {
		out->batches_done = ctl.batches_done - ctl0.batches_done;
		out->shots_done = ctl.shots_done - ctl0.shots_done;
		out->stopped = ctl.stopped;
		out->last_energy_len = ctl.last_energy_len;
		cudaEventElapsedTime(&out->gpu_ms, c->ev0, c->ev1);
		out->kernel_launches = (uint32_t)launches;
		out->big_triangles = 0;
		for (int l = 0; l < RAD_MAX_LANES; l++) out->big_triangles += ctl.lane[l].parked;
		out->queue_overflow = ctl.q_overflow;
	}
	if (ctl.q_overflow) { c->err = "rad_shoot: tile queue overflow"; return RAD_E_CUDA; }
	if (ctl.ring_abort) {
		cudaMemsetAsync(&c->d.ctl->ring_abort, 0, sizeof(uint32_t), c->stream);
		c->err = "rad_shoot: the ring pipeline stalled (a stage hand-over inside raster_ring_kernel timed out); set RAD_RING=0 to use the lane path";
		return RAD_E_CUDA;
	}
	if (c->ring_failed) { c->ring_failed = false; return RAD_E_CUDA; }      // the cooperative launch was refused (c->err says why)
	return RAD_OK;
}

int rad_read_itembuffer(rad_ctx* c, uint32_t hi, uint32_t* ids_out) {
	if (!c || !ids_out || hi >= c->d.k) return RAD_E_ARG;
	cudaSetDevice(c->cfg.device);
	RAD_CUDA_TRY(c, cudaMemcpyAsync(ids_out, c->d.items + (size_t)hi * c->d.RES, (size_t)c->d.RES * 4, cudaMemcpyDeviceToHost, c->stream));
	return sync_check(c);
}
int rad_write_itembuffer(rad_ctx* c, uint32_t hi, const uint32_t* ids) {
	if (!c || !ids || hi >= c->d.k) return RAD_E_ARG;
	cudaSetDevice(c->cfg.device);
	RAD_CUDA_TRY(c, cudaMemcpyAsync(c->d.items + (size_t)hi * c->d.RES, ids, (size_t)c->d.RES * 4, cudaMemcpyHostToDevice, c->stream));
	return sync_check(c);
}
int rad_read_depthbuffer(rad_ctx* c, uint32_t hi, uint32_t* depth_out) {
	if (!c || !depth_out || hi >= c->d.k) return RAD_E_ARG;
	if (!c->rendered) { c->err = "rad_read_depthbuffer: only valid right after rad_render_hemicubes"; return RAD_E_STATE; }
	cudaSetDevice(c->cfg.device);
	uint32_t* tmp = nullptr;
	RAD_CUDA_TRY(c, cudaMalloc((void**)&tmp, (size_t)c->d.RES * 4));
	rad_launch_read_depth(c, hi, tmp);
	cudaError_t e = cudaMemcpyAsync(depth_out, tmp, (size_t)c->d.RES * 4, cudaMemcpyDeviceToHost, c->stream);
	cudaStreamSynchronize(c->stream);
	cudaFree(tmp);
	if (e != cudaSuccess) { c->err = cudaGetErrorString(e); return RAD_E_CUDA; }
	return sync_check(c);
}
int rad_read_formfactors(rad_ctx* c, uint32_t hi, float* F) {
	if (!c || !F || hi >= c->d.k) return RAD_E_ARG;
	cudaSetDevice(c->cfg.device);
	RAD_CUDA_TRY(c, cudaMemcpyAsync(F, c->d.F + (size_t)hi * c->d.P, (size_t)c->d.P * 4, cudaMemcpyDeviceToHost, c->stream));
	return sync_check(c);
}
int rad_read_mvp(rad_ctx* c, uint32_t hi, uint32_t face, float* out16) {
	if (!c || !out16 || hi >= c->d.k || face >= RAD_NFACES) return RAD_E_ARG;
	cudaSetDevice(c->cfg.device);
	RAD_CUDA_TRY(c, cudaMemcpyAsync(out16, c->d.mvp + ((size_t)hi * RAD_NFACES + face) * 16, 64, cudaMemcpyDeviceToHost, c->stream));
	return sync_check(c);
}

int rad_bench_process(rad_ctx* c, uint32_t repeat, float* ms_per_launch) {
	int r = need_ready(c, "rad_bench_process"); if (r) return r;
	if (!c->emitters_ready) { c->err = "rad_bench_process: no emitters"; return RAD_E_STATE; }
	if (repeat < 1) repeat = 1;
	rad_launch_process(c);                         // warm-up
	RAD_CUDA_TRY(c, cudaEventRecord(c->ev0, c->stream));
	for (uint32_t i = 0; i < repeat; i++) rad_launch_process(c);
	RAD_CUDA_TRY(c, cudaEventRecord(c->ev1, c->stream));
	RAD_CUDA_TRY(c, cudaMemsetAsync(c->d.F, 0, (size_t)c->d.k * c->d.P * 4, c->stream));
	if ((r = sync_check(c))) return r;
	float ms = 0; cudaEventElapsedTime(&ms, c->ev0, c->ev1);
	if (ms_per_launch) *ms_per_launch = ms / repeat;
	c->processed = false;
	return RAD_OK;
}

int rad_bench_atomics(rad_ctx* c, uint32_t pattern, uint64_t count, float* gops_out) {
	if (!c || !gops_out || pattern > 2) return RAD_E_ARG;
	cudaSetDevice(c->cfg.device);
	const uint32_t blocks = 148 * 16, threads = blocks * 128;
	uint32_t steps = (uint32_t)((count + threads - 1) / threads);
	steps = (steps + 7) / 8 * 8;
	if (steps < 8) steps = 8;
	rad_launch_atomic_bench(c, pattern, 8, blocks);                       // warm-up
	RAD_CUDA_TRY(c, cudaEventRecord(c->ev0, c->stream));
	rad_launch_atomic_bench(c, pattern, steps, blocks);
	RAD_CUDA_TRY(c, cudaEventRecord(c->ev1, c->stream));
	rad_launch_clear_keys(c);                                             // the benchmark scribbles over the key buffers
	int r = sync_check(c); if (r) return r;
	float ms = 0; cudaEventElapsedTime(&ms, c->ev0, c->ev1);
	*gops_out = (float)((double)threads * steps / (ms * 1e-3) / 1e9);
	return RAD_OK;
}

int rad_profile_batch(rad_ctx* c, float* ms6) {
	int r = need_ready(c, "rad_profile_batch"); if (r) return r;
	if (!ms6) return RAD_E_ARG;
	const bool keep = (c->cfg.flags & RAD_FLAG_KEEP_ITEMBUFFER) != 0;
	if (c->d.k == 1 && !c->selkey_valid) rad_launch_argmax(c);
	if (c->keys_dirty) rad_launch_clear_keys(c);
	// the same launch sequence as one steady-state batch of rad_shoot, with an event after every launch
	std::vector<cudaEvent_t> ev; std::vector<int> stage;
	auto mark = [&](int st) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, c->stream); ev.push_back(e); stage.push_back(st); };
	mark(-1);
	rad_launch_select(c); mark(0);
	c->lane_delta_done = false;
	rad_launch_raster_process_marked(c, keep, [&](int st) { mark(st); });
	const bool fuse = c->d.k == 1;
	if (c->world > 1 && (c->nccl_comm || c->peer_mode)) {
		// sharded batch (collective: every rank profiles the same batch): [3] = local dB (+ the reduce-scatter kernel of the
		// two-shot exchange / ncclAllReduce), [5] = the update kernel including its wait for the peers
		if (!c->lane_delta_done) rad_launch_delta(c);
		if (c->peer_mode && c->d.xtwo) rad_launch_xreduce(c);
		if (c->nccl_comm && !c->peer_mode) g_nccl.AllReduce(c->d.dB, c->d.dB, (size_t)3 * c->d.P, kNcclFloat32, kNcclSum, c->nccl_comm, c->stream);
		mark(3);
		rad_launch_finish(c, false); mark(5);
		c->selkey_valid = false;
	} else {
		rad_launch_apply(c, fuse); mark(5);
	}
	if (fuse) { c->parity ^= 1; c->selkey_valid = true; c->cam_valid = true; }
	r = sync_check(c);
	for (int i = 0; i < 6; i++) ms6[i] = 0.0f;
	for (size_t i = 1; i < ev.size(); i++) { float ms = 0; cudaEventElapsedTime(&ms, ev[i - 1], ev[i]); ms6[stage[i]] += ms; }
	for (cudaEvent_t e : ev) cudaEventDestroy(e);
	c->emitters_ready = c->rendered = c->processed = false;
	return r;
}

// ---- display stage ----------------------------------------------------------------------------------
int rad_upload_neighbours(rad_ctx* c, const int32_t* nb8, uint32_t P) {
	if (!c || !nb8) return RAD_E_ARG;
	if (!c->have_scene || P != c->d.P) { c->err = "rad_upload_neighbours: upload the scene first (same P)"; return RAD_E_STATE; }
	cudaSetDevice(c->cfg.device);
	for (size_t i = 0; i < (size_t)P * 8; i++)
		if (nb8[i] < 0 || (uint32_t)nb8[i] >= P) { c->err = "rad_upload_neighbours: neighbour id out of range"; return RAD_E_ARG; }
	int r = stage_dev(c, (size_t)P * 12 * 4); if (r) return r;
	memcpy(c->h_stage, nb8, (size_t)P * 32);
	RAD_CUDA_TRY(c, cudaMemcpyAsync(c->d_stage, c->h_stage, (size_t)P * 32, cudaMemcpyHostToDevice, c->stream));
	rad_launch_nb_to_planes(c, reinterpret_cast<const int32_t*>(c->d_stage), P);
	RAD_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
	c->have_nb = true;
	return RAD_OK;
}

int rad_shade_vertices(rad_ctx* c, float* colors12_out, float* gpu_ms_out) {
	if (!c || !colors12_out) return RAD_E_ARG;
	if (!c->have_scene || !c->have_nb) { c->err = "rad_shade_vertices: upload the scene and the neighbours first"; return RAD_E_STATE; }
	cudaSetDevice(c->cfg.device);
	const size_t P = c->d.P;
	int r = stage_dev(c, P * 12 * 4); if (r) return r;
	RAD_CUDA_TRY(c, cudaEventRecord(c->ev0, c->stream));
	rad_launch_shade(c, c->d_stage);
	RAD_CUDA_TRY(c, cudaEventRecord(c->ev1, c->stream));
	RAD_CUDA_TRY(c, cudaMemcpyAsync(c->h_stage, c->d_stage, P * 48, cudaMemcpyDeviceToHost, c->stream));
	if ((r = sync_check(c))) return r;
	memcpy(colors12_out, c->h_stage, P * 48);
	if (gpu_ms_out) cudaEventElapsedTime(gpu_ms_out, c->ev0, c->ev1);
	return RAD_OK;
}

// ---- device-side state snapshot (benchmarks restart from the same state without host traffic) ----
int rad_save_state(rad_ctx* c) {
	int r = need_ready(c, "rad_save_state"); if (r) return r;
	const size_t P = c->d.P;
	if (!c->saved) RAD_CUDA_TRY(c, cudaMalloc((void**)&c->saved, 6 * (size_t)c->cfg.max_patches * 4));
	RAD_CUDA_TRY(c, cudaMemcpyAsync(c->saved, c->d.rad, 3 * P * 4, cudaMemcpyDeviceToDevice, c->stream));
	RAD_CUDA_TRY(c, cudaMemcpyAsync(c->saved + 3 * P, c->d.illum, 3 * P * 4, cudaMemcpyDeviceToDevice, c->stream));
	return sync_check(c);
}
int rad_restore_state(rad_ctx* c) {
	int r = need_ready(c, "rad_restore_state"); if (r) return r;
	if (!c->saved) { c->err = "rad_restore_state: nothing saved"; return RAD_E_STATE; }
	const size_t P = c->d.P;
	RAD_CUDA_TRY(c, cudaMemcpyAsync(c->d.rad, c->saved, 3 * P * 4, cudaMemcpyDeviceToDevice, c->stream));
	RAD_CUDA_TRY(c, cudaMemcpyAsync(c->d.illum, c->saved + 3 * P, 3 * P * 4, cudaMemcpyDeviceToDevice, c->stream));
	RAD_CUDA_TRY(c, cudaMemsetAsync(c->d.ctl, 0, sizeof(RadControl), c->stream));
	c->selkey_valid = false; c->emitters_ready = c->rendered = c->processed = false;
	return sync_check(c);
}

// ---- multi-GPU ----------------------------------------------------------------------------------
int rad_nccl_unique_id(void* id_out128) {
	if (!id_out128) return RAD_E_ARG;
	if (!g_nccl.load(g_create_err)) return RAD_E_NCCL;
	NcclId id; memset(&id, 0, sizeof(id));
	if (g_nccl.GetUniqueId(&id) != 0) { g_create_err = "ncclGetUniqueId failed"; return RAD_E_NCCL; }
	memcpy(id_out128, &id, 128);
	return RAD_OK;
}

int rad_set_partition(rad_ctx* c, int rank, int world) {
	if (!c || world < 1 || rank < 0 || rank >= world) return RAD_E_ARG;
	drop_graph(c);
	c->rank = rank; c->world = world;
	c->d.h0 = (uint32_t)((uint64_t)c->d.k * rank / world);
	c->d.h1 = (uint32_t)((uint64_t)c->d.k * (rank + 1) / world);
	c->d.deal = (world > 1 && c->d.k % (uint32_t)world == 0) ? (uint32_t)world : 0u;   // top-k lists are dealt out to the ranks (load balance)
	c->partition_only = c->nccl_comm == nullptr;
	return RAD_OK;
}

int rad_comm_init(rad_ctx* c, int rank, int world, const void* id128) {
	if (!c || !id128) return RAD_E_ARG;
	cudaSetDevice(c->cfg.device);
	if (!g_nccl.load(c->err)) return RAD_E_NCCL;
	NcclId id; memcpy(&id, id128, 128);
	int rc = g_nccl.CommInitRank(&c->nccl_comm, world, id, rank);
	if (rc != 0) { c->err = std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error"); c->nccl_comm = nullptr; return RAD_E_NCCL; }
	int r = rad_set_partition(c, rank, world);
	c->partition_only = false;
	return r;
}

// ---- fused exchange over peer memory (NVLink, CUDA IPC): no collective call, the update kernel reads every rank's dB ----
int rad_peer_handle(rad_ctx* c, void* handle64_out) {
	if (!c || !handle64_out) return RAD_E_ARG;
	cudaSetDevice(c->cfg.device);
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
	if (!c->xbuf) {
		c->xbuf_bytes = (size_t)RAD_XB_DATA + 3ull * 3ull * c->cfg.max_patches * 4ull;
		RAD_CUDA_TRY(c, cudaMalloc((void**)&c->xbuf, c->xbuf_bytes));
		RAD_CUDA_TRY(c, cudaMemset(c->xbuf, 0, c->xbuf_bytes));
	}
	cudaIpcMemHandle_t h;
	RAD_CUDA_TRY(c, cudaIpcGetMemHandle(&h, c->xbuf));
	memcpy(handle64_out, &h, 64);
	return RAD_OK;
}
int rad_peer_init(rad_ctx* c, int rank, int world, const void* handles /* world x 64 bytes, rank order */) {
	if (!c || !handles || world < 1 || world > RAD_MAX_PEERS || rank < 0 || rank >= world) return RAD_E_ARG;
	if (!c->xbuf) { c->err = "rad_peer_init: call rad_peer_handle first"; return RAD_E_STATE; }
	cudaSetDevice(c->cfg.device);
	drop_graph(c);
	for (int r = 0; r < world; r++) {
		if (r == rank) { c->d.xb[r] = c->xbuf; continue; }
		cudaIpcMemHandle_t h; memcpy(&h, (const char*)handles + 64 * (size_t)r, 64);
		void* p = nullptr;
		RAD_CUDA_TRY(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
		c->peer_ptr[r] = p; c->d.xb[r] = (char*)p;
	}
	c->d.xrank = (uint32_t)rank; c->d.xworld = (uint32_t)world; c->d.xPmax = c->cfg.max_patches;
	// one shot (every rank reads all peers' planes) while that is a few MB, two shots (reduce-scatter + all-gather) beyond
	c->d.xtwo = (world > 2 && (uint64_t)c->cfg.max_patches * 12ull * (uint64_t)(world - 1) > (4ull << 20)) ? 1u : 0u;
	if (const char* e = getenv("RAD_XTWO")) c->d.xtwo = atoi(e) != 0 ? 1u : 0u;   // tuning knob
	c->d.xnowait = 0u;
	if (const char* e = getenv("RAD_XNOWAIT")) c->d.xnowait = atoi(e) != 0 ? 1u : 0u;   // measurement knob (wrong results)
	c->peer_mode = true;
	int r = rad_set_partition(c, rank, world);
	c->partition_only = false;
	return r;
}

int rad_batch_partial(rad_ctx* c) {
	int r = need_ready(c, "rad_batch_partial"); if (r) return r;
	if (c->peer_mode) { c->err = "rad_batch_partial: the host-mediated exchange is not available after rad_peer_init (dB lives in the exchange buffer)"; return RAD_E_STATE; }
	const bool keep = (c->cfg.flags & RAD_FLAG_KEEP_ITEMBUFFER) != 0;
	rad_launch_select(c);
	c->lane_delta_done = false;
	rad_launch_raster_process(c, keep);
	if (!c->lane_delta_done) rad_launch_delta(c);
	return sync_check(c);
}
int rad_read_delta(rad_ctx* c, float* dB3) {
	if (!c || !dB3) return RAD_E_ARG;
	cudaSetDevice(c->cfg.device);
	int r = stage_dev(c, 12 * (size_t)c->d.P * 4); if (r) return r;
	return d2h_aos3(c, c->d.dB, dB3, c->d.P);
}
int rad_write_delta(rad_ctx* c, const float* dB3) {
	if (!c || !dB3) return RAD_E_ARG;
	cudaSetDevice(c->cfg.device);
	int r = stage_dev(c, 12 * (size_t)c->d.P * 4); if (r) return r;
	return h2d_aos3(c, dB3, c->d.dB, c->d.P);
}
int rad_batch_finish(rad_ctx* c, float* last_energy_len) {
	int r = need_ready(c, "rad_batch_finish"); if (r) return r;
	rad_launch_finish(c, false);
	c->selkey_valid = false;
	if ((r = sync_check(c))) return r;
	RadControl ctl; if ((r = read_ctl(c, &ctl))) return r;
	if (last_energy_len) *last_energy_len = ctl.last_energy_len;
	return RAD_OK;
}

} // extern "C"
