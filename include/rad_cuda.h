/*
 * rad_cuda.h — C ABI of the B200-native radiosity shooting path (librad_cuda.so).
 *
 * This is the drop-in boundary for the hot path of david-sabata/Radiosity: it replaces the raw
 * OpenGL + OpenCL calls that `OnIdle` makes from Main.cpp (the reference has no plugin interface;
 * the seam is that set of calls).  Every entry point cites the reference code it stands in for
 * (file:line under /root/reference/source/).  Plain pointers and sizes only; no C++/torch types.
 *
 * Conventions
 *   - every function returns 0 on success, a negative RAD_E_* code otherwise; rad_last_error()
 *     gives the message.  Nothing throws across the boundary.  (Reference: bool init functions
 *     printing to cerr, Main.cpp:291-294,409-412,846-855.)
 *   - the caller owns all host arrays; the library copies on upload/download and owns all device
 *     memory (reference: GL/CL own device objects, Main.cpp:605-607,665-667).
 *   - a context is bound to one GPU and is NOT thread-safe (reference: one GL context + one in-order
 *     CL queue, Main.cpp:431).  It owns its CUDA streams: one for the calls below, plus the raster
 *     lanes rad_shoot forks from it and joins back into it.
 *   - there is no CPU fallback: without a CUDA device rad_create fails with RAD_E_CUDA.
 *   - patch id = index into the flat scene arrays (ModelContainer.cpp:81-155).  Item buffers hold
 *     `id + 1` per pixel, 0 = nothing rendered (the reference packs the same number into an RGBA8
 *     colour, Colors.cpp:143-174; 0 = black = cleared).
 *   - atlas layout of one hemicube: W = 2N, H = 1.5N, row 0 = bottom (GL), rows [0,N) = LEFT half |
 *     FRONT | RIGHT half, rows [N,1.5N) = UP half | DOWN half (Main.cpp:314-389, FormFactors.cpp:250-272).
 */
#ifndef RAD_CUDA_H
#define RAD_CUDA_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rad_ctx rad_ctx;

enum {
	RAD_OK = 0,
	RAD_E_ARG = -1,      /* bad argument / call order */
	RAD_E_CUDA = -2,     /* CUDA runtime error (no device, OOM, launch failure) */
	RAD_E_STATE = -3,    /* scene / form factors not uploaded yet */
	RAD_E_NCCL = -4      /* NCCL unavailable or failed */
};

/* shooter selection semantics */
enum {
	RAD_SELECT_REFERENCE = 0,  /* exactly ModelContainer::getHighestRadiosityPatchesId (ModelContainer.cpp:259-299),
	                              quirks included: patch 0 seeded, reject-below-min while not full, tie scrambling */
	RAD_SELECT_TOPK = 1        /* clean top-k by |B|^2 (> 0), ordered (energy desc, id asc) — documented divergence */
};

#define RAD_MAX_HEMICUBES 512   /* largest batch (k); RAD_SELECT_REFERENCE: at most 64 */

typedef struct rad_config {
	uint32_t hemicube_side;   /* N  = Config::HEMICUBE_W()  (Config.cpp:34)             */
	uint32_t hemicubes;       /* k  = Config::HEMICUBES_CNT() (Config.cpp:10,150), 1 .. RAD_MAX_HEMICUBES */
	uint32_t max_patches;     /* capacity of the scene arrays                            */
	int32_t  device;          /* CUDA device ordinal                                     */
	uint32_t select_mode;     /* RAD_SELECT_*                                            */
	float    reflectivity;    /* REFLECTIVITY, Patch.h:13 (0.3)                          */
	float    projection[16];  /* column-major f[col][row]: CGLTransform::Perspective(90,1,0.01,1000), Main.cpp:1172 */
	uint32_t flags;           /* RAD_FLAG_*                                              */
} rad_config;

enum {
	RAD_FLAG_KEEP_ITEMBUFFER = 1u  /* rad_shoot also materialises the uint32 item buffers (parity/debug);
	                                  without it the fused path never writes them to HBM */
};

typedef struct rad_stats {
	uint32_t batches_done;     /* passCounter, Main.cpp:1307                              */
	uint32_t shots_done;       /* emitters actually shot (non-NULL)                       */
	uint32_t stopped;          /* 1 if the |lastEnergy| < 0.1 test fired (Main.cpp:1297-1300) */
	float    last_energy_len;  /* |B| of the last non-NULL emitter before subtraction (Main.cpp:1292) */
	float    gpu_ms;           /* CUDA-event time of the whole call on the context's stream */
	uint32_t kernel_launches;  /* kernels launched by this call                           */
	uint32_t big_triangles;    /* triangles that went through the tile queue (last batch) */
	uint32_t queue_overflow;   /* 1 if the tile queue overflowed (results then invalid)   */
} rad_stats;

/* InitGLObjects atlas/FBO/VBO creation (Main.cpp:17-37,271-294) + InitCLObjects (Main.cpp:405-611) */
int rad_create(rad_ctx** out, const rad_config* cfg);
/* CleanupGLObjects / CleanupCLObjects (Main.cpp:616-670) */
int rad_destroy(rad_ctx* ctx);
/* message of the last failing call on ctx (ctx may be NULL for rad_create failures) */
const char* rad_last_error(const rad_ctx* ctx);

/* dFF table upload, clEnqueueWriteBuffer(ocl_arg_ffactors) (Main.cpp:576).  ff = ONE hemicube,
 * n = 3*N*N floats in atlas order (precomputeHemicubeFormFactors, FormFactors.cpp:280-339). */
int rad_set_formfactors(rad_ctx* ctx, const float* ff, uint32_t n);

/* scene upload: glBufferData of scene.getVertices() (Main.cpp:19) + the per-patch state the
 * reference reads through Patch* (Patch.h:46-54).  verts12 = float[P*12] (4 verts x xyz),
 * color3 / radiosity3 / illumination3 = float[P*3].  Page-locked host arrays (cudaHostAlloc, cudaHostRegister, torch
 * pin_memory) are read by the copy engine directly, pageable ones go through the context's pinned staging buffer; the
 * same holds for rad_upload_state and, for the output arrays, rad_download_state.  One synchronisation per call. */
int rad_upload_scene(rad_ctx* ctx, const float* verts12, const float* color3, const float* radiosity3,
                     const float* illumination3, uint32_t P);
/* re-upload only B and I (restart a run on the same geometry) */
int rad_upload_state(rad_ctx* ctx, const float* radiosity3, const float* illumination3);
/* read back Patch::radiosity / Patch::illumination (used by Main.cpp:1323-1366, SaveToFile Main.cpp:1562) */
int rad_download_state(rad_ctx* ctx, float* radiosity3, float* illumination3);

/* S1: scene.getHighestRadiosityPatchesId(k, ...) (Main.cpp:1140) on the device, followed by the
 * per-emitter snapshot (Main.cpp:1161) and the 5 face MVPs (Main.cpp:1172-1183).
 * ids_out[k] (may be NULL); valid_out[k] = 0 for the reference's NULL emitters. */
int rad_select(rad_ctx* ctx, uint32_t* ids_out, uint32_t* valid_out);
/* same, but with a caller-chosen emitter list (n <= k) — fixed schedules for parity tests */
int rad_set_emitters(rad_ctx* ctx, const uint32_t* ids, uint32_t n);

/* S2: glClear + 5*k glDrawElements into the atlas (Main.cpp:1148-1202, DrawPatchLook Main.cpp:689-725) */
int rad_render_hemicubes(rad_ctx* ctx);
/* S3: clEnqueueAcquireGLObjects + ProcessHemicube kernel + read-backs + record gather
 * (Main.cpp:1212-1269, Kernel_ProcessHemicube.h:9-70): F_h[i] = sum of dFF over pixels showing patch i */
int rad_process_hemicubes(rad_ctx* ctx);
/* S4..S6: energy transfer, emitter update, stop test (Main.cpp:1272-1303) */
int rad_apply(rad_ctx* ctx, float* last_energy_len);

/* the whole loop `for shoot < SHOOTS_PER_CYCLE` (Main.cpp:1137-1309), device resident: no host
 * round trip per shot.  stop_test != 0 honours the 0.1 termination exactly like the reference's loop condition
 * (`&& computeRadiosity`, Main.cpp:1137,1297-1300): the batch whose last emitter had |B| < 0.1 is the last one applied —
 * every state-changing kernel checks a device-side gate, so the batches a CUDA-graph replay still holds after it are
 * no-ops; batches_done / shots_done count what was really shot.  The call starts with the test re-armed. */
int rad_shoot(rad_ctx* ctx, uint32_t n_batches, int stop_test, rad_stats* out);

/* device-side snapshot / restore of (B, I): restart a run from the same state with no host traffic */
int rad_save_state(rad_ctx* ctx);
int rad_restore_state(rad_ctx* ctx);

/* ---- display stage (the step after the path) ------------------------------------------------------
 * Colors::smoothShadePatch for every patch (Colors.cpp:198-261, loop at Main.cpp:1323-1341): vertex colour = mean over the
 * patch and three of its 8 neighbours of colour (.) (I + B).  nb8 = int32[P*8] neighbour ids (Patch::neighbours,
 * index 0 = top-left, clockwise; a patch without a neighbour points at itself, Patch.h:50).
 * colors12_out = float[P*12] in the reference's VBO order (lb, rb, rt, lt per patch, Colors.cpp:256-259). */
int rad_upload_neighbours(rad_ctx* ctx, const int32_t* nb8, uint32_t P);
int rad_shade_vertices(rad_ctx* ctx, float* colors12_out, float* gpu_ms_out /* may be NULL */);

/* parity / debugging seams: the `P` preview key and FBO2BMP (FormFactors.cpp:143-171) */
int rad_read_itembuffer(rad_ctx* ctx, uint32_t hi, uint32_t* ids_out /* W*H */);
int rad_read_depthbuffer(rad_ctx* ctx, uint32_t hi, uint32_t* depth24_out /* W*H, valid after rad_render_hemicubes */);
int rad_read_formfactors(rad_ctx* ctx, uint32_t hi, float* F /* P */);
int rad_read_mvp(rad_ctx* ctx, uint32_t hi, uint32_t face /* 0..4 = UP,DOWN,LEFT,RIGHT,FRONT */, float* out16);
/* feed an externally produced item buffer to rad_process_hemicubes (kernel benchmarks, atlas import) */
int rad_write_itembuffer(rad_ctx* ctx, uint32_t hi, const uint32_t* ids /* W*H */);

/* ProcessHemicube in isolation over all k item buffers, `repeat` launches timed with CUDA events on
 * the context's stream; *ms_per_launch = average.  (BASELINE.json: "ProcessHemicube Gpix/s vs HBM") */
int rad_bench_process(rad_ctx* ctx, uint32_t repeat, float* ms_per_launch);
/* 64-bit atomicMin roofline of this GPU for the rasteriser's access pattern (SURVEY.md §8d asks for a measured L2-atomic
 * peak): `count` fire-and-forget RED.MIN.64 over the context's key buffers; pattern 0 = raster-like (each quarter warp
 * walks 8 consecutive pixels x 8 rows of a randomly placed 8x8 box), 1 = fully coalesced (32 consecutive pixels per warp),
 * 2 = fully scattered (every lane its own random pixel).  *gops_out = 1e9 atomics per second. */
int rad_bench_atomics(rad_ctx* ctx, uint32_t pattern, uint64_t count, float* gops_out);
/* per-kernel CUDA-event timing of one un-graphed batch: ms[0]=select+camera ms[1]=raster setup+small
 * ms[2]=raster tiles ms[3]=resolve ms[4]=process ms[5]=apply.  Runs and applies one batch. */
int rad_profile_batch(rad_ctx* ctx, float* ms6);

/* ---- multi-GPU batched shooting (new; the reference is single-GPU) -------------------------
 * Scene + state replicated on every rank; every rank runs the same deterministic selection;
 * rank r renders/processes emitters [r*k/G, (r+1)*k/G); received energy dB[P][3] is combined with
 * ONE ncclAllReduce(sum) per batch over NVLink; every rank applies the identical update. */
int rad_nccl_unique_id(void* id_out128 /* 128 bytes */);
int rad_comm_init(rad_ctx* ctx, int rank, int world, const void* id128);
/* Fused exchange over peer memory (default of bench.py): every rank exposes ONE exchange buffer through CUDA IPC
 * (rad_peer_handle -> 64 opaque bytes, all-gathered by the caller, e.g. torch.distributed), rad_peer_init maps the
 * peers' buffers over NVLink.  The batch then needs no collective call: the local-dB kernel publishes its planes with
 * a release flag on every peer, the update kernel waits for all flags and sums the ranks' planes in rank order while
 * it applies them.  All ranks must have finished shooting before any of them calls rad_destroy. */
int rad_peer_handle(rad_ctx* ctx, void* handle64_out /* 64 bytes */);
int rad_peer_init(rad_ctx* ctx, int rank, int world, const void* handles /* world x 64 bytes, rank order */);
/* With world > 1 (any of the three modes) a RAD_SELECT_TOPK list is DEALT OUT to the ranks when world divides k: list
 * entry j goes to slot (j % world) * (k / world) + j / world, so that every rank renders the same mix of strong and weak
 * shooters (rad_select returns ids in slot order).  The batch's result does not depend on the slot order beyond the
 * order of float additions. */
/* partition-only mode (no NCCL): shard like `world` ranks and expose the partial dB so that a
 * host-side collective (e.g. torch.distributed gloo in CPU tests) can combine it */
int rad_set_partition(rad_ctx* ctx, int rank, int world);
int rad_batch_partial(rad_ctx* ctx);                       /* select + render + process + local dB */
int rad_read_delta(rad_ctx* ctx, float* dB3 /* P*3 */);
int rad_write_delta(rad_ctx* ctx, const float* dB3 /* P*3 */);
int rad_batch_finish(rad_ctx* ctx, float* last_energy_len); /* B += dB, emitter update, stop test */

/* introspection */
uint32_t rad_patch_count(const rad_ctx* ctx);
uint32_t rad_atlas_width(const rad_ctx* ctx);
uint32_t rad_atlas_height(const rad_ctx* ctx);
const char* rad_version(void);

#ifdef __cplusplus
}
#endif
#endif /* RAD_CUDA_H */
