// Internal declarations shared by the CUDA translation units of librad_cuda.so.
// Public surface: include/rad_cuda.h.  All device float math is compiled with --fmad=false and
// IEEE div/sqrt so that it reproduces, operation for operation, the float32 evaluation order of
// the reference's host code (Vector.h / Transform.cpp / Camera.cpp) and of the CPU oracle.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <functional>
#include <utility>
#include "../../include/rad_cuda.h"

#define RAD_NFACES 5
#define RAD_CLEAR_KEY 0xFFFFFFFFFFFFFFFFull
#define RAD_TILE 32               // chunk edge in pixels (chunks are bbox-relative) — batched default; see RadDev::tile
#define RAD_SMALL_STEPS 64         // quarter-warp walk: ceil(bbox pixels / 8) steps at most — batched default; see RadDev::small_steps

struct RadBigTri {                // one screen-space triangle parked for tile processing (64 B)
	int X0, Y0, X1, Y1, X2, Y2;   // snapped window coordinates, 8 sub-pixel bits
	float z0, dz1, dz2, inv_area; // depth plane in barycentric form
	uint32_t id1;                 // patch id + 1
	uint32_t slot;                // hemicube slot (atlas index)
	int px0, py0, px1, py1;       // pixel bbox already clipped to the face scissor
};
struct RadQueueEntry { uint32_t tri; uint16_t tx, ty; };
// One small screen-space QUAD (both triangles of a patch, (0,1,2) and (0,2,3), sharing one bbox walk) parked for the
// quarter-warp walk of raster_queue_kernel; 64 B = four 16-byte loads.  Vertex coordinates are relative to the centre of
// the bbox origin pixel (the set-up kernel guarantees they fit int16 and that every edge function fits int32).
// A lone triangle is stored with v3 = v0 and invB = 0: its second triangle is degenerate and never covers a pixel.
struct __align__(16) RadSmallQuad {
	short x0, y0, x1, y1, x2, y2, x3, y3;   // snapped window coordinates (8 sub-pixel bits) - origin pixel centre
	float Z0, Z1, Z2, Z3;                   // window depth of the four vertices
	float invA, invB;                       // 1 / (2 * area) of triangles (0,1,2) and (0,2,3)
	uint32_t id1;                           // patch id + 1
	uint16_t slot;                          // hemicube slot (atlas index)
	uint16_t rcpw;                          // ceil(32768 / w): (l * rcpw) >> 15 == l / w exactly for l <= 8, w <= 255
	uint16_t px0, py0;                      // bbox origin (already clipped to the face scissor)
	uint8_t w, h;                           // bbox size in pixels (w * h <= 8 * small_steps)
	uint16_t pad0; uint32_t pad1[2];
};

#define RAD_MAX_PEERS 8
#define RAD_XB_DATA 4096           // byte offset of the dB planes inside an exchange buffer
#define RAD_XB_FLAG2 2048          // byte offset of the second flag row (two-shot exchange: reduced slices ready)
#define RAD_MAX_LANES 8            // concurrent raster lanes (streams) a batch can be split into
#define RAD_SPEC_POOL 128          // ... and the slots such a context holds: two halves, one being applied while the other is rendered
#define RAD_SPEC_SLOTS 64          // hemicubes rendered ahead of a strict-progressive (k == 1) run, see rad_ctx::spec
#define RAD_RING_SLOTS 64          // hemicube slots one launch of the ring path renders (per-slot work lists, see RadRing)
struct RadQueueCtl {              // work-list counters of one raster lane
	uint32_t q_tris;              // chunk queue: triangles parked
	uint32_t q_entries;           // chunk queue: (triangle, chunk) entries
	uint32_t q_small;             // small-quad queue: records (one quarter warp each)
	uint32_t n_pairs;             // (patch, face) pairs that survived the conservative culls
	uint32_t parked;              // records parked by the lane's last batch (statistics)
	uint32_t pad[3];
};
struct RadSpecStep {              // one simulated shot of the speculative k == 1 path (spec_sim_kernel -> spec_verify_kernel / spec_commit_kernel)
	unsigned long long key;       // the shooter's argmax key (energy bits << 32 | id) right before the shot
	uint32_t slot, id;            // rendered-ahead slot and patch
	float S[3], c[3];             // its B at that moment, its colour
	float len;                    // |B| of the emitter after the transfer, before the subtraction (the stop test's lastEnergy)
	float pad;
};
struct RadControl {               // small device-resident control block
	unsigned long long selkey[2]; // k==1 selection: (E bits << 32 | id), ping-pong by batch parity
	unsigned long long spec_key[3];   // speculative k == 1 path: the grid's argmax key of shot s in spec_key[s % 3]
	uint32_t q_overflow;
	uint32_t stopped;             // |lastEnergy| < 0.1 seen
	float last_energy_len;
	uint32_t batches_done;
	uint32_t shots_done;
	uint32_t gate;                // stop test armed (RadDev::stop_gate) and `stopped` seen at the start of a batch: the rest of the
	                              // replay is a no-op (the loop of Main.cpp:1137 ends with the batch that stopped)
	RadQueueCtl lane[RAD_MAX_LANES];
	uint32_t ticket;              // blocks finished ("last block merges" pattern of the selection / update kernels)
	uint32_t spec_target;         // speculative k == 1 path: shots_done at which the call ends
	uint32_t spec_done;           // ... reached (or the stop test fired while armed): the batches still enqueued do nothing
	uint32_t spec_count_sim;      // shots spec_sim_kernel could simulate over the rendered-ahead set
	uint32_t spec_jstar;          // ... of which the first spec_jstar are what the strict loop does (spec_verify_kernel: no patch outside overtakes before)
	uint32_t spec_alldark;        // nothing carries energy: the remaining shots are the reference's no-op shots of patch 0
	uint32_t spec_hits, spec_misses;   // statistics: shots served from a rendered-ahead hemicube / batches cut short by a shooter outside the set
	uint32_t ref_fast;            // RAD_SELECT_REFERENCE, k > 1: the tie-free fast path has written the emitter list of this batch
	uint32_t ring_abort;          // ring path watchdog: a wait inside raster_ring_kernel timed out (never expected; the call fails instead of hanging)
};
// Ring path (RadRing): work-list counters per hemicube slot of the launch and the stage hand-over between walk and process
// CTAs.  Zeroed before every launch (one memset node).  A stage's counter (CTAs that have finished it) and its ready flag
// live on different 128-byte lines: the CTAs that wait poll the FLAG, so their loads never queue up in front of the
// atomics of the CTAs that are still signalling.
struct RadRingCtl {
	RadQueueCtl slot[RAD_RING_SLOTS];
	uint32_t walk_cnt[RAD_RING_SLOTS][32];
	uint32_t proc_cnt[RAD_RING_SLOTS][32];
	uint32_t walk_ready[RAD_RING_SLOTS][32];
	uint32_t proc_ready[RAD_RING_SLOTS][32];
	unsigned long long t_start, t_walk_end, t_proc_end, t_walk_wait, t_proc_wait;   // RAD_RING_DEBUG & 16: globaltimer stamps / summed wait time (ns)
};

struct RadEmitter {               // per hemicube slot
	uint32_t id;
	uint32_t valid;               // 0 == the reference's NULL emitter
	float S[3];                   // radiosity snapshot (Main.cpp:1161)
	float color[3];               // emitter colour (Main.cpp:1274)
	float eye[3];                 // patch centre (hemicube eye) and un-normalised patch normal: conservative culling only
	float nrm[3];
	float ax[9];                  // orthonormal shooter frame s, t, f (rows) for the conservative culls
	uint32_t order;               // position in the selection list (== slot unless the list is dealt out to ranks, see RadDev::deal)
	float ctol;                   // squared relative margin of the conservative culls: (2e-3 + 3 * deviation of the faces' real
	                              // float32 view bases from the ideal frame ax)^2, see camera_emitter
};

struct RadDev {                   // device pointers + sizes, passed by value to kernels
	uint32_t P, N, W, H, RES, k;
	uint32_t h0, h1;              // hemicube slots this rank renders/processes (kernels: slots of this launch)
	uint32_t deal;                // G > 1: the top-k list is dealt out to the G ranks like cards (entry j -> rank j % G), so that every
	                              // rank gets the same mix of strong and weak shooters and the same number of non-NULL ones:
	                              // slot(j) = (j % G) * (k / G) + j / G.  0 / 1: slot(j) = j
	uint32_t kbase;               // slot whose keys live in key buffer 0 (fused path recycles L2-resident key buffers per group)
	uint32_t tag;                 // epoch tag (top byte of every key written / accepted by this launch)
	uint32_t inline_area;         // bbox area (px) up to which the owning lane rasterises alone; larger -> chunk queue
	uint32_t small_steps;         // longest quarter-warp walk (8 px per step) accepted by the small-quad queue
	uint32_t spec;                // 1 in the view the speculative k == 1 path launches its batches with: the batch gate follows RadControl::spec_done
	uint32_t stop_gate;           // rad_shoot(stop_test): batches enqueued after the one whose stop test fired do nothing (RadControl::gate)
	uint32_t tile;                // chunk edge (px) of the chunk queue.  k == 1 is latency-bound (one hemicube cannot fill the
	                              // GPU): shorter walks and smaller chunks there, longer ones for batches
	float reflectivity;
	const float4* v0; const float4* v1; const float4* v2;   // verts: (v1.xyz,v2.x) (v2.yz,v3.xy) (v3.z,v4.xyz)
	const float* color;           // [3][P] planes
	float* rad;                   // [3][P] planes  (B, unshot)
	float* illum;                 // [3][P] planes  (I, shot)
	const float* ff;              // [RES]
	unsigned long long* keys;     // [k][RES] (epoch tag << 56 | depth24 << 32 | id+1); empty = any other top byte
	uint32_t* items;              // [k][RES] id+1
	float* F;                     // [k][P]
	float* dB;                    // [3][P] partial received energy (multi-GPU: NCCL / host-mediated exchange)
	// fused peer-memory exchange (multi-GPU, rad_peer_init): every rank owns one exchange buffer, mapped into all its
	// peers over NVLink (CUDA IPC):  [0] uint32 seq | [128 + 128 r] uint32 flag of rank r | [RAD_XB_FLAG2 + 128 r] second flag
	// row | [RAD_XB_DATA] float dB[2][3][Pmax] | float red[3][Pmax] (two-shot: this rank's reduced slice of the patches)
	// xb[r] = rank r's buffer as seen from this GPU (xb[xrank] is the local one); xworld == 0: not in use
	char* xb[RAD_MAX_PEERS];
	uint32_t xrank, xworld, xPmax;
	uint32_t xnowait;             // measurement knob RAD_XNOWAIT=1: do not wait for the peers' flags (results are garbage; shows what the waiting costs)
	uint32_t xtwo;                // two-shot exchange (reduce-scatter kernel + all-gather in the update kernel) for large P
	float* mvp;                   // [k][5][16] column-major
	RadEmitter* em;               // [k]
	float4* emlite;               // [k][2] what the update kernel needs of an emitter: (S, valid | order << 1), (colour, id)
	RadControl* ctl;
	RadSpecStep* spec_steps;      // [RAD_SPEC_SLOTS] the simulated shots of the batch in flight
	float4* spec_cand;            // [3][blocks] (argmax key, B of that patch) of every block of spec_apply_kernel, by shot % 3
	RadRingCtl* rc;               // ring path: per-slot work-list counters + stage hand-over (see RadRingCtl)
	RadQueueCtl* qc;              // this launch's lane counters (= &ctl->lane[lane])
	RadBigTri* q_tri; RadQueueEntry* q_ent;
	uint32_t q_tri_cap, q_ent_cap;
	uint32_t* pairs; uint32_t pairs_cap;  // compacted (patch | face << 23 | local slot << 26) work list of the exact set-up stage
	RadSmallQuad* q_sm; uint32_t q_sm_cap;   // small-quad queue (bbox steps <= small_steps, int32 walk)
	uint32_t* ework;              // [max(P,64)] scratch (emitter id staging)
	unsigned long long* cand0; unsigned long long* cand1;   // top-k tournament candidates, ceil(P/2048) * keep keys each
	const float* proj;            // [16]
	int32_t* nb;                  // [8][P] neighbour ids (display stage), plane j = neighbour j
	float* shade_e;               // [3][P] colour (.) (I + B) scratch of the display stage
};

// Ring path of the steady state (raster.cu, raster_ring_kernel): the slots of a launch are rendered in STAGES of `sg` slots
// through a ring of `rs` stages of key buffers (rs * sg buffers, sized to stay L2-resident).  One persistent kernel holds two
// kinds of warps: walk warps rasterise the parked records of stage after stage into the ring, process warps follow one
// stage behind and turn the finished key buffers into F (ProcessHemicube) — so a key is written by REDs that hit L2 and read
// back from L2, and never crosses HBM.  The work lists are kept per slot (slot-ordered walks need them), the hand-over
// between the two kinds of warps is a pair of counters per stage (RadControl::walk_done / proc_done).
struct RadRing {
	uint32_t nslots;              // slots of this launch: D.h0 .. D.h0 + nslots (<= RAD_RING_SLOTS)
	uint32_t sg, rs, nst;         // slots per stage, stages in the ring, stages of this launch
	uint32_t tag0;                // epoch tag of ring round 0; round r (stage / rs) writes tag0 - r
	uint32_t cap_pairs, cap_sm, cap_tri, cap_ent;   // per-slot capacities of the four work lists
	uint32_t walk_ctas;           // CTAs [0, walk_ctas) walk, the rest process
	uint32_t nw_walk, nw_proc;    // warps of each kind
	uint32_t keep_items;
	uint32_t debug;               // measurement knob RAD_RING_DEBUG: 1 walk CTAs do not wait for the ring (results are garbage), 16 role time stamps (rad_profile_batch prints them)
};

// Tile-binned rasteriser (raster_tiles.cu; opt-in, RAD_RASTER=tiles): the atlas of every hemicube is cut into tiles of
// RAD_TILE_W x RAD_TILE_H pixels (tile_walk.cuh); the parked records of a launch group are binned per (slot, tile) and one
// CTA per tile resolves visibility in shared memory and feeds ProcessHemicube from there — no key buffer, no item buffer.
// Every (slot, tile) has two lists: 0 = small-quad records, 1 = large triangles.
struct RadTiles {
	uint32_t* cnt;                // [slots][T][2] records per list (zero between launches)
	uint32_t* base;               // [slots * T * 2 + 1] exclusive scan of cnt = first reference of every list
	uint32_t* refs;               // record indices (list 0: into q_sm, list 1: into q_tri), grouped by list
	uint32_t refs_cap;
	uint32_t tx, ty, T;           // tiles per atlas row / column / hemicube
};

struct rad_ctx {
	rad_config cfg;
	RadDev d;
	bool tile_mode; RadTiles tl;  // RAD_RASTER=tiles: the fused path (rad_shoot) runs the tile-binned rasteriser
	cudaStream_t stream;
	cudaEvent_t ev0, ev1;
	// raster lanes: a batch's hemicube slots are split into `lanes` groups that run cull -> set-up -> queues -> process
	// concurrently on their own streams (forked from / joined to `stream`), each with its own counters and its own
	// share of the work lists; cur = the stream the launchers enqueue on
	uint32_t lanes; cudaStream_t lane_stream[RAD_MAX_LANES]; cudaEvent_t ev_fork, ev_lane[RAD_MAX_LANES];
	cudaStream_t cur;
	std::string err;
	bool have_nb;
	bool have_ff, have_scene, emitters_ready, rendered, processed, keys_dirty;
	uint32_t parity;              // selkey ping-pong for k==1
	bool selkey_valid;            // selkey[parity] holds the argmax of the current B
	bool cam_valid;               // ... and the fused update's tail has already prepared that shooter's camera (k == 1)
	// CUDA graph of the steady-state loop
	cudaGraphExec_t graph_exec; uint32_t graph_batches; bool graph_keep_items; uint32_t graph_launches, graph_parity0, graph_stop;
	float* saved;                 // device snapshot of (B, I) for rad_save_state / rad_restore_state
	// host staging
	float* h_stage; size_t h_stage_bytes;      // pinned
	float* d_stage; size_t d_stage_bytes;      // device mirror of the staging buffer
	// multi-GPU
	int rank, world; void* nccl_comm; bool partition_only; bool multi_graph;
	bool peer_mode; char* xbuf; size_t xbuf_bytes; void* peer_ptr[RAD_MAX_PEERS];   // fused exchange over peer memory
	uint32_t launches;            // kernels launched since last reset
	uint32_t graph_epoch_after;   // epoch value after one replay of the captured graph
	uint32_t epoch;               // next key epoch tag (254 .. 1, decreasing; 0 = clear the key buffers first)
	bool inline_area_forced;      // RAD_INLINE_AREA set: do not auto-tune the inline tier
	bool ring_mode;               // RAD_RING=1 (opt-in): steady state through the L2-resident key ring instead of raster lanes + whole-batch key buffers
	uint32_t ring_sg, ring_rs, ring_proc_layers;   // tuning knobs (0 = automatic): RAD_RING_SG, RAD_RING_RS, RAD_RING_PROC
	// Speculative strict progressive refinement (k == 1, default; RAD_SPEC=0: one hemicube per launch).  F_h depends on the
	// geometry only, so the hemicubes of the RAD_SPEC_SLOTS currently strongest patches are rendered and processed as ONE
	// batch (raster lanes, full GPU), and spec_apply_kernel then replays the reference's one-shot-at-a-time loop over them:
	// argmax of the CURRENT B, transfer with the S of that moment, emitter update, stop test — exactly the k == 1 semantics.
	// A shot whose shooter is not among the rendered ones ends the batch; the next batch is selected from the state reached.
	uint32_t key_bufs;            // key buffers allocated (<= key_slots)
	uint32_t key_slots;           // hemicube slots the key / F / emitter buffers hold (hemicubes, or RAD_SPEC_SLOTS for a speculative k == 1 context)
	bool spec; uint32_t spec_slots; cudaGraphExec_t spec_graph; uint32_t spec_graph_batches, spec_graph_launches, spec_graph_stop, spec_graph_epoch_after; int spec_blocks;
	// render-ahead overlap of the speculative path: the list of a selection goes to em[sel_base .. +sel_count), patches already
	// waiting in em[sel_excl .. +sel_excl_n) are left out; the camera kernel covers [cam_base, +cam_count); defer_join: the raster
	// lanes are not joined to the context's stream by rad_launch_raster_process (rad_join_lanes does it later)
	bool spec_overlap; uint32_t sel_base, sel_count, sel_excl, sel_excl_n, cam_base, cam_count, last_lanes; bool defer_join;
	int select_override;          // 0: cfg.select_mode; 2: top-k with the argmax's tie rule (higher id first) — the speculative path's candidate set
	bool pdl;                     // RAD_PDL=1 (opt-in, k == 1): the kernels of a shot are chained by programmatic dependent launches
	bool lane_delta_done;         // multi-GPU: the raster lanes of the batch being enqueued have added their dB themselves (no whole-rank kernel needed)
	bool ring_failed;             // a raster_ring_kernel launch was refused since the last check
	int ring_ctas_per_sm;         // resident CTAs per SM of raster_ring_kernel (occupancy query, once)
	uint32_t l2_group_mb;         // key-buffer footprint (MB) of one hemicube group of the fused path
};

// ---- launchers (each enqueues on ctx->stream and bumps ctx->launches) -------------------------
void rad_launch_select(rad_ctx* c);                 // S1 + camera/snapshots (all modes)
void rad_launch_camera(rad_ctx* c, int sel_parity = -1);   // camera/snapshots (+ k==1: emitter from selkey[sel_parity])
void rad_launch_raster(rad_ctx* c);                 // setup + inline/warp raster, then tile queue
void rad_launch_raster_setup_only(rad_ctx* c);
void rad_launch_raster_tiles_only(rad_ctx* c);
void rad_launch_set_emitters(rad_ctx* c, const uint32_t* d_ids, uint32_t n);
void rad_launch_resolve(rad_ctx* c, bool reset);    // keys -> items (+ keys reset)
void rad_launch_clear_keys(rad_ctx* c);
uint32_t rad_next_tag(rad_ctx* c);
void rad_launch_process(rad_ctx* c);                // items -> F
void rad_launch_resolve_process(rad_ctx* c, bool keep_items);   // fused (all slots, keys indexed from kbase = h0)
void rad_launch_raster_process(rad_ctx* c, bool keep_items);    // steady state: per L2-sized hemicube group raster -> fused process
void rad_launch_raster_process_marked(rad_ctx* c, bool keep_items, const std::function<void(int)>& mark);   // same, mark(stage) after each launch (1 set-up, 2 chunks, 4 process)
void rad_launch_apply(rad_ctx* c, bool fuse_select); // S4..S6 (+ argmax of the new B for k==1)
void rad_launch_delta(rad_ctx* c);                  // multi-GPU: local dB (whole rank, one kernel)
void rad_launch_lane_delta(rad_ctx* c, const RadDev& V, cudaStream_t st, uint32_t s0, uint32_t n);   // multi-GPU: local dB of one raster lane's slots, added into the rank's planes
void rad_launch_xreduce(rad_ctx* c);                // multi-GPU, fused two-shot exchange: this rank's slice of the summed dB
void rad_launch_finish(rad_ctx* c, bool fuse_select); // multi-GPU: B += dB, emitter update
int rad_launch_spec_apply(rad_ctx* c, const RadDev& S, uint32_t slot_base, uint32_t nslots, int stop_armed);
void rad_join_lanes(rad_ctx* c);                    // the raster lanes of a deferred-join batch rejoin the context's stream   // speculative k == 1 path: the sequential part (cooperative launch)
void rad_launch_argmax(rad_ctx* c);                 // k==1: prime selkey[parity] from the current B
void rad_launch_read_depth(rad_ctx* c, uint32_t hi, uint32_t* d_out);
void rad_launch_tiles_view(rad_ctx* c, const RadDev& V, const RadTiles& T, cudaStream_t st, uint32_t s0, uint32_t n, bool keep_items,
                           const std::function<void(int)>& mark);   // raster_tiles.cu: bins + tile CTAs (visibility + ProcessHemicube)
void rad_launch_atomic_bench(rad_ctx* c, uint32_t pattern, uint32_t steps, uint32_t blocks);   // raster.cu
void rad_launch_aos3_to_planes(rad_ctx* c, const float* aos, float* planes, uint32_t P);   // layout.cu
void rad_launch_planes_to_aos3(rad_ctx* c, const float* planes, float* aos, uint32_t P);
void rad_launch_split_quads(rad_ctx* c, const float* verts12, uint32_t P);
void rad_launch_nb_to_planes(rad_ctx* c, const int32_t* nb8, uint32_t P);
void rad_launch_shade(rad_ctx* c, float* out12);   // shade.cu

// Programmatic dependent launch (opt-in, RAD_PDL=1; k == 1: one hemicube per launch, the shot is a chain of five latency-bound kernels).  A
// kernel launched through rad_launch_pdl with pdl = true may START while its predecessor in the stream is still running —
// its CTAs get resident and run up to pdl_enter(), which returns once the predecessor has completed and its writes are
// visible — so the launch latency of every link of the chain overlaps the previous kernel.  pdl_enter() must be the first
// statement of such a kernel (it also lets the NEXT kernel start early); it is a no-op in a normal launch.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_enter() { asm volatile("griddepcontrol.launch_dependents;\n\tgriddepcontrol.wait;" ::: "memory"); }
template <typename... KArgs, typename... Args>
static inline cudaError_t rad_launch_pdl(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
	cudaLaunchAttribute at[1];
	at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = at; cfg.numAttrs = pdl ? 1u : 0u;
	return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
#endif

#define RAD_CUDA_TRY(c, expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { \
	(c)->err = std::string(#expr) + ": " + cudaGetErrorString(e_); return RAD_E_CUDA; } } while (0)
