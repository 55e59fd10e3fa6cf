/* swe.h -- C-ABI of the B200-native wind-erosion engine (libswe_b200.so).
 *
 * Drop-in boundary for the reference's erode cycle. The reference has no FFI layer: its
 * "operator API" is the static C++ interface World::erode / World::cascade / World::map
 * accessors / struct Wind (source/world.hpp:9-32, source/wind.hpp:4-26,
 * source/cellpool.h:432-480). Each entry point below names the reference interface it
 * replaces; include/swe_world.hpp re-creates those C++ names on top of this ABI and
 * INTEGRATION.md shows the binding a maintainer adds.
 *
 * Conventions: plain pointers and sizes only; every call returns SWE_OK (0) or a negative
 * swe_status and records a message retrievable with swe_last_error() (per thread). The
 * reference itself has no error convention (oob -> NULL/false/0.0f, cellpool.h:90-94,
 * 432-445,472-476); out-of-bounds cell queries here return SWE_ERR_OOB instead.
 * There is NO CPU fallback: without a CUDA device swe_create fails with SWE_ERR_CUDA.
 *
 * Map layout: nx rows (x, slow index) by nz columns (z, fast index), cell index
 * x*nz + z -- math::flatten (source/include/math.h:11-12). Host-side planes are fp32
 * with an arbitrary byte stride so the engine can read/write the reference's 36-byte AoS
 * quad::cell (cellpool.h:226-240) in place: pass &cell[0].height and stride 36.
 */
#ifndef SWE_H
#define SWE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWE_ABI_VERSION 3

typedef enum swe_status {
  SWE_OK = 0,
  SWE_ERR_INVALID = -1,   /* bad argument */
  SWE_ERR_CUDA = -2,      /* CUDA runtime/driver error (message has the detail) */
  SWE_ERR_OOM = -3,
  SWE_ERR_OOB = -4,       /* cell coordinate outside the map */
  SWE_ERR_STATE = -5      /* call not valid in the engine's current state */
} swe_status;

/* quad::cell members, in declaration order (cellpool.h:226-240). */
typedef enum swe_plane {
  SWE_HEIGHT = 0,
  SWE_MASSFLOW = 1,
  SWE_MOMENTUMX = 2,
  SWE_MOMENTUMY = 3,
  SWE_MOMENTUMZ = 4,
  SWE_MASSFLOW_TRACK = 5,
  SWE_MOMENTUMX_TRACK = 6,
  SWE_MOMENTUMY_TRACK = 7,
  SWE_MOMENTUMZ_TRACK = 8,
  SWE_NUM_PLANES = 9
} swe_plane;

/* How World::erode runs its particles. */
typedef enum swe_mode {
  /* One particle at a time in the reference's order with the reference's libc rand()
   * stream and IEEE fp32 (no FMA contraction): bitwise identical to the reference. */
  SWE_MODE_DETERMINISTIC = 0,
  /* All particles of the call in flight at once, one thread per particle, counter-based
   * RNG, atomic erode/deposit scatter; after every round of steps World::cascade runs once at
   * each touched cell in conflict-free colour order. Same physics, statistical parity, exact
   * mass accounting. */
  SWE_MODE_PARALLEL = 1
} swe_mode;

/* Engine options (swe_set_option). */
typedef enum swe_option {
  /* 0 (default) = fused nine-colour shared-memory tile kernel, tile loaded by TMA when the row
   * pitch allows; 1 = one global-memory launch per colour; 2 = fused tile kernel, plain loads.
   * All are bitwise identical to the reference order; 1 and 2 exist for cross-checking. */
  SWE_OPT_CASCADE_VARIANT = 1,
  /* Parallel mode: Wind::fly steps each particle takes between two cascade passes (default 1:
   * like the reference, a touched cell is relaxed before any particle steps again). */
  SWE_OPT_CASCADE_INTERVAL = 2,
  /* 1 = record CUDA events around every particle-round and cascade launch of swe_erode so that
   * swe_stats.last_round_ms / last_cascade_ms report each kernel's share (default 0). */
  SWE_OPT_KERNEL_TIMING = 3,
  /* 1 (default) = a parallel swe_erode on a single-GPU engine is captured once per (cycles, params,
   * options) as a CUDA graph of ~1550 nodes and replayed; 0 = launch kernel by kernel. */
  SWE_OPT_CUDA_GRAPH = 4,
  /* 1 (default) = the kernels of a parallel swe_erode round (particle step, stamped cascade, commit)
   * are chained with programmatic dependent launches: each may be set up while its predecessor
   * drains and waits on-device (griddepcontrol.wait) before touching its results. Hides part of the
   * launch latency when a call has few particles (the reference's own 8*tilesize frame: 2-5 %);
   * 0 = plain stream order. Results are identical either way. */
  SWE_OPT_PDL = 5,
  /* Parallel mode, single-GPU engines, lock-step schedule (SWE_OPT_SCHEDULE = 1) only: the particles of one swe_erode call fly as B successive batches
   * (default 1 = all at once), each spread over the whole map and seeing the terrain the previous
   * batches left. The reference flies its particles strictly one after another; with many particles
   * per cell in flight at once the parallel mode exports more sand than the reference (measured:
   * +24 % at 0.25 particles per cell, within 1.3 % at 0.03), and batches trade throughput for that
   * fidelity. Track reset and EMA stay once per call. */
  SWE_OPT_CALL_BATCHES = 6,
  /* Parallel mode: how the particles of a call are scheduled. All three keep Wind::fly's arithmetic; they differ
   * in how close the ORDER of operations stays to the reference's one-particle-after-another loop.
   * 2 (default) = tile-resident scheduler (single-GPU engines, closed boundary): the map is cut into tiles, one
   * CTA keeps a tile's heights in shared memory for a super-round of K rounds of nine colour micro-rounds and
   * serves the particles standing in the tile's interior; three shifted tilings take turns (csrc/swe_tile.cu).
   * On stripe engines wired with swe_stripe_connect_ring the tiling is global: a tile belongs to the rank that
   * owns its centre row and works on a neighbour's rows in the neighbour's memory; stripes must be at least two
   * tiles tall, otherwise (and for layered stripes and periodic maps) schedule 0 runs when 2 is selected.
   * 0 = colour-queue scheduler: the cells are coloured 3 x 3, a micro-round serves one colour over the whole map
   * (one persistent cooperative kernel, a grid barrier per micro-round), and every particle standing on a cell
   * of that colour (one per cell, the others wait) flies one Wind::fly step, erodes and runs World::cascade on
   * its 3x3 block at once -- blocks of one colour are disjoint, so each particle sees the reference's order of
   * operations (wind.hpp:108-114) and the height plane needs no atomics. Schedules 0 and 2 serve a particle the
   * same way and are parity-green at the benchmarked densities (tests/test_gpu_density.py, test_gpu_tiles.py).
   * 1 = round 1's lock-step pipeline (all particles step once, then one stamped cascade at every touched cell);
   * it exports +24 % more sand than the reference at 0.25 particles per cell. */
  SWE_OPT_SCHEDULE = 7,
  /* Colour-queue and tile-resident schedulers: the call's particles are injected evenly over the first S rounds (one
   * round = nine micro-rounds) instead of all at once, so that all particle ages are in flight
   * together as in the reference. -1 (default) = auto: S = 8 * (particles per cell) * (maxAge + 2),
   * at most 4 * (maxAge + 2); calls thinner than 0.02 particles per cell are injected at once. */
  SWE_OPT_STAGGER_ROUNDS = 8,
  /* Colour-queue scheduler: 1 (default) = ONE cooperative kernel launch runs every micro-round of the
   * call, the blocks meeting at a grid barrier between micro-rounds (~1.5 us instead of a ~6 us dependent
   * launch; the call ends on the device when the last particle dies); 0 = one launch per micro-round
   * (CUDA graph + programmatic dependent launches), with a host-side completion check. */
  SWE_OPT_PERSISTENT = 9,
  /* 1 = periodic boundary, the variant the reference carries as comments beside its live closed boundary
   * (cellpool.h:96-113 slice::oob/get, :447-465 map::oob/get): nothing is out of bounds, every cell access wraps,
   * particles never leave the map (they die of age). Honoured by the deterministic mode (bitwise against the
   * reference compiled with those blocks swapped in), swe_cascade / swe_height / swe_normal / the vertex and normal
   * read-backs, whole-map sweeps (per-colour kernels) and the colour-queue parallel mode; sweeps and the parallel
   * mode need map edges that are multiples of 3 (the nine-colour order on a torus). Single-GPU engines without a
   * rock plane. Default 0. */
  SWE_OPT_PERIODIC = 10,
  /* Tile-resident schedule (SWE_OPT_SCHEDULE = 2): rounds (nine micro-rounds each) a CTA keeps its tile in shared
   * memory before the particles are re-sorted into the next, shifted tiling. 0 (default) = auto
   * (16 for 60 x 60 tiles, 24 for 120 x 120); range [0, 64]. */
  SWE_OPT_TILE_ROUNDS = 11,
  /* Tile-resident schedule: tile geometry (rows x columns). -1 (default) = by map size and wind: 14,400-cell tiles /
   * 256 threads when that still gives every SM three tiles -- 240 x 60 or 60 x 240, elongated along the prevailing
   * wind, when pspeed points mostly along x or z, else 120 x 120 -- otherwise 60 x 60 / 128 threads.
   * 0 = 60 x 60, 1 = 120 x 120, 2 = 240 x 60, 3 = 60 x 240. */
  SWE_OPT_TILE_SHAPE = 12
} swe_option;

/* Tunables. Defaults (swe_params_default) are the reference's statics and constants:
 * Wind::{maxAge,boundarylayer,suspension,gravity} wind.hpp:23-26,
 * World::{lrate,maxdiff,settling} world.hpp:30-32, quad::mapscale cellpool.h:184,
 * Wind::pspeed wind.hpp:13. */
typedef struct swe_params {
  int32_t max_age;
  float boundarylayer;
  float suspension;
  float gravity;
  float lrate;
  float maxdiff;
  float settling;
  int32_t mapscale;
  float pspeed[3];
  float abrasion;       /* layered extension: fraction of unmet erosion demand ground out of rock (0 = off) */
  int32_t reserved[4];  /* must be zero */
} swe_params;

/* struct Wind (wind.hpp:4-21) as plain data. pos = (x, altitude, z). */
typedef struct swe_wind {
  int32_t age;
  float sediment;
  float pos[3];
  float speed[3];
} swe_wind;

/* Counters accumulated since swe_create / swe_reset_stats. */
typedef struct swe_stats {
  uint64_t particles;      /* particles spawned */
  uint64_t steps;          /* completed Wind::fly calls (returned true) */
  uint64_t exits_oob;      /* particles that left the map (wind.hpp:32-38) */
  uint64_t exits_age;      /* particles that exceeded maxAge (wind.hpp:45-46) */
  double sediment_lost;    /* sum of Wind::sediment at death: mass that left the system */
  uint64_t cascade_cells;  /* World::cascade centres processed by swe_cascade_sweep */
  uint64_t erode_calls;
  double last_erode_ms;    /* device time of the last swe_erode (CUDA events) */
  double last_sweep_ms;    /* device time of the last swe_cascade_sweep call */
  uint64_t kernel_launches;/* kernels launched by the engine so far */
  double last_fly_ms;      /* device time of the particle phase (rounds + cascades) of the last swe_erode */
  double last_round_ms;    /* with SWE_OPT_KERNEL_TIMING: sum over the particle-round kernels */
  double last_cascade_ms;  /* with SWE_OPT_KERNEL_TIMING: sum over the per-round cascade kernels */
  uint64_t stripe_overflow;/* stripe engines: particles dropped for lack of buffer space (must be 0) */
  uint64_t last_micro_rounds;/* colour-queue / tile scheduler: micro-rounds the last swe_erode needed */
  uint64_t queue_dropped;  /* colour-queue / tile scheduler: particle records lost to a full queue (must be 0) */
  uint64_t last_super_rounds;/* tile-resident scheduler: super-rounds (tile_fly_kernel launches) of the last swe_erode; 0 otherwise */
} swe_stats;

/* Threading contract. An engine is not re-entrant: calls on ONE engine must come from one host thread at a time
 * (like the reference, whose World is all statics: world.hpp:12-19,100). Different engines may be driven from
 * different host threads concurrently (one thread per GPU is the natural in-process multi-GPU pattern); the
 * library's one-time set-up is guarded. Work is enqueued on the engine's stream; swe_last_error() is per thread.
 * Logical stripes sharing ONE GPU must be driven through swe_stripe_group_* from one thread (their kernels wait for
 * each other on the device); keep them to at most 8 per device (CUDA_DEVICE_MAX_CONNECTIONS hardware queues). */
typedef struct swe_engine swe_engine;

/* ---- lifecycle -------------------------------------------------------------------- */
int swe_abi_version(void);
/* Tile-resident schedule, introspection (host arithmetic only, no device needed). swe_tile_geometry: rows, columns and
 * threads per CTA of a compiled tile geometry (SWE_OPT_TILE_SHAPE value). swe_tile_owned_rows: on a map of nx rows cut
 * into row stripes, the tile rows [*first, *end) of tiling `tiling` (the three shifted tilings: 0, 1, 2) that the rank
 * owning global rows [own_lo, own_hi) launches -- a tile belongs to the rank that owns the row of its centre --, and
 * the number of tile rows of that tiling. */
int swe_tile_geometry(int shape, int* edge_x, int* edge_z, int* threads);
int swe_tile_owned_rows(int nx, int shape, int tiling, int own_lo, int own_hi, int* first, int* end, int* total);
const char* swe_last_error(void);
int swe_params_default(swe_params* p);

/* Replaces cellpool.reserve(quad::area) + World::map.init's node/slice construction
 * (SimpleWindErosion.cpp:36-38, cellpool.h:139-143,343-352) with device-resident planes;
 * the map size is a run-time value here (quad::tilesize is compile-time, cellpool.h:186).
 * All planes start at zero. `device` = CUDA ordinal. */
int swe_create(swe_engine** out, int nx, int nz, const swe_params* params, int device);
void swe_destroy(swe_engine* e);
int swe_set_params(swe_engine* e, const swe_params* params);   /* Wind::maxAge = ...; etc. */
int swe_get_params(const swe_engine* e, swe_params* out);
int swe_get_size(const swe_engine* e, int* nx, int* nz);
int swe_set_option(swe_engine* e, int option, int value);
int swe_get_option(const swe_engine* e, int option, int* value);

/* srand(seed) (SimpleWindErosion.cpp:29,33): seeds the deterministic mode's libc-rand
 * replica and the parallel mode's counter-based generator. */
int swe_srand(swe_engine* e, uint32_t seed);

/* Run state for checkpoint / resume: the libc-rand replica of the deterministic mode (the
 * reference's stream is process-global and continues across World::erode calls), the seed and the
 * World::erode call index of the parallel mode's counter-based generator. Together with the nine
 * planes (swe_download_cells) this is everything a run needs to continue: a deterministic run
 * resumed from it is bitwise identical to the uninterrupted one. */
#define SWE_STATE_BYTES 256
int swe_export_state(swe_engine* e, void* buf, size_t bytes);
int swe_import_state(swe_engine* e, const void* buf, size_t bytes);

/* ---- map transfer (host buffers) --------------------------------------------------- */
/* Write / read one quad::cell member for every cell. `stride_bytes` is the distance
 * between consecutive cells in the host buffer (4 = packed plane, 36 = reference AoS). */
int swe_upload_plane(swe_engine* e, int plane, const float* src, ptrdiff_t stride_bytes);
int swe_download_plane(swe_engine* e, int plane, float* dst, ptrdiff_t stride_bytes);
/* The per-frame read-back of the reference's render loop (SimpleWindErosion.cpp:219-230: updatenode reads
 * height, the wind map reads the momentum fields): five packed planes height, massflow, momentumx,
 * momentumy, momentumz (5*nx*nz floats) with one pass over the field plane and one synchronisation. */
int swe_download_frame(swe_engine* e, float* dst5);
/* The same read-back, pipelined with the next swe_erode: _begin enqueues the copies and returns; the height plane
 * (dst5[0 .. cells)) travels on the engine's stream, the four flow-field planes are snapshotted on the device and
 * travel on a side stream while later work on the engine's stream proceeds. _wait(e, 0) blocks until the height
 * plane has arrived, _wait(e, 1) until all five planes have (and must be called before the next _begin reuses the
 * snapshot; _begin does so itself if the caller did not). dst5 must stay valid until _wait(e, 1); use pinned host
 * memory (swe_host_alloc) for a truly asynchronous copy. */
int swe_download_frame_begin(swe_engine* e, float* dst5);
int swe_download_frame_wait(swe_engine* e, int all);
/* Whole reference pool in one go: `cells` points at nx*nz 36-byte quad::cell records. */
int swe_upload_cells(swe_engine* e, const void* cells);
int swe_download_cells(swe_engine* e, void* cells);
/* Layered-terrain extension (NOT in the reference snapshot, whose cell has one height; our
 * specification): a bedrock plane under the sand, sand = height - rock >= 0.
 * Parallel-mode erosion takes sand only and grinds rock at swe_params.abrasion efficiency;
 * every cascade transfer is capped by the giving cell's sand. Uploading a rock plane switches
 * the engine to layered mode (single-GPU engines, parallel mode and cascade calls);
 * swe_clear_rock returns to the reference's single layer. */
int swe_upload_rock(swe_engine* e, const float* src, ptrdiff_t stride_bytes);
int swe_download_rock(swe_engine* e, float* dst, ptrdiff_t stride_bytes);
int swe_clear_rock(swe_engine* e);
/* Pinned host memory for fast transfers (optional; any host pointer is accepted). */
int swe_host_alloc(void** out, size_t bytes);
int swe_host_free(void* p);

/* ---- the hot path ------------------------------------------------------------------ */
/* World::erode(cycles) (world.hpp:36-77): zero the track planes, spawn `cycles`
 * particles and fly each to its death, then EMA-update the flow fields. The work is
 * enqueued on the engine's stream; every swe_download_* / swe_get_stats / swe_sync / query
 * call waits for it, so the map a consumer reads afterwards is always the eroded one.
 * Parallel mode on the tile-resident schedule (SWE_OPT_SCHEDULE = 2, the default) enqueues super-rounds until the
 * device reports that no particle is left, so the call returns when the particle phase is over (the EMA pass may
 * still be in flight); the other paths return as soon as the work is enqueued. */
int swe_erode(swe_engine* e, int cycles, int mode);

/* World::cascade(vec2 pos) (world.hpp:79-153) at one cell. */
int swe_cascade(swe_engine* e, int x, int z);

/* Whole-map cascade: `sweeps` passes; each pass applies World::cascade to every cell in
 * the conflict-free 3x3 colour order (colour c = 0..8, cx = c/3, cz = c%3, centres with
 * x%3 == cx and z%3 == cz). Within a colour the 3x3 footprints are disjoint, so the
 * result is bitwise identical to calling the reference's World::cascade in that order.
 * If max_excess_out != NULL it receives max over neighbour pairs of |dh| - d*maxdiff
 * after the last sweep (<= 0 means relaxed to the repose angle). */
int swe_cascade_sweep(swe_engine* e, int sweeps, float* max_excess_out);

/* Iterate sweeps until max excess < tol or max_sweeps reached; returns sweeps done. */
int swe_cascade_converge(swe_engine* e, float tol, int max_sweeps, int* sweeps_done,
                         float* max_excess_out);

/* Wind(vec2 pos) + while(wind.fly()) (wind.hpp:6,28-117) for one particle in
 * deterministic mode, recording after every completed step 8 floats
 * {pos.x,pos.y,pos.z,speed.x,speed.y,speed.z,sediment,age} into `records` (capacity
 * max_records steps; may be NULL). The map is modified exactly as the reference does.
 * *steps_out = completed steps; *final_out = particle state at death. */
int swe_trace(swe_engine* e, float x, float z, int max_records, float* records,
              int* steps_out, swe_wind* final_out);

/* One Wind::fly() call (wind.hpp:28) on a caller-held particle, deterministic mode.
 * *alive_out = the bool fly() returns. */
int swe_wind_init(swe_wind* w, float x, float z);                 /* Wind::Wind(vec2) */
int swe_wind_fly(swe_engine* e, swe_wind* w, int* alive_out);

/* quad::map::normal(p) / height(p) / oob(p) (cellpool.h:432-480). */
int swe_normal(swe_engine* e, int x, int z, float out3[3]);
int swe_height(swe_engine* e, int x, int z, float* out);
/* normals for every cell (what updatenode reads each frame, cellpool.h:302-321):
 * dst = nx*nz*3 floats. */
int swe_download_normals(swe_engine* e, float* dst);

/* The Vertex records updatenode writes for every cell each frame (cellpool.h:302-321,
 * vertexpool.hpp:6-26): 12 floats = 48 bytes per cell, index x*nz + z (math::flatten):
 *   position = (x, mapscale*height, z), normal = map.normal(p),
 *   tangent = T - P, bitangent = B - P   (T, B: the surface at (x+1, z), (x, z+1); outside the map
 *   the height is 0, cellpool.h:257-261).
 * swe_download_vertices fills a host buffer of nx*nz*48 bytes; swe_write_vertices_device writes the
 * same records to DEVICE memory on the engine's stream without a host round trip -- e.g. the
 * CUDA-mapped pointer of the renderer's persistently mapped vertex buffer (vertexpool.hpp:155-172);
 * 16-byte aligned, not synchronised: use swe_sync or the stream. */
int swe_download_vertices(swe_engine* e, void* dst);
int swe_write_vertices_device(swe_engine* e, void* device_dst);

/* ---- generate (host generate of the reference, run on the device) ------------------- */
/* The noise layers of quad::map::init (cellpool.h:388-411; a comment in the mounted snapshot):
 *   for o in 0..layers-1: height += scale_o * noise.GetNoise(x/tile, z/tile, (float)(seed % 10000))
 * with frequency 2^o, scale_0 = 0.6, scale *= 0.6, and `noise` the FastNoiseLite object as map::init
 * configures it (OpenSimplex2, FBm of 3 octaves, seed 1337; cellpool.h:366-368). Bitwise identical
 * to the host library (fp32 + wrapping int arithmetic only). ADDS to the current heights, like the
 * reference adds to its Gaussian hill; tile = coordinate divisor (quad::tilesize), 0 = nz;
 * layers <= 16. Works on stripe engines (global coordinates; ghost rows are generated too). */
int swe_generate_noise(swe_engine* e, int seed, int layers, int tile);
/* The normalisation that ends map::init (cellpool.h:416-428): min and max are folded from 0.0f,
 * then height = 0.5*((height - min)/(max - min)). Split in two calls so that stripes can reduce
 * min/max across ranks in between; a single GPU passes the range straight through. */
int swe_height_range(swe_engine* e, float* min_out, float* max_out);
int swe_normalize_height(swe_engine* e, float min, float max);

/* ---- diagnostics ------------------------------------------------------------------- */
int swe_get_stats(swe_engine* e, swe_stats* out);
int swe_reset_stats(swe_engine* e);
/* checkpoint / resume: make the counters of an earlier swe_get_stats the baseline again (particles, steps, exits,
 * sediment_lost, cascade_cells, erode_calls; the timing fields are ignored) */
int swe_import_stats(swe_engine* e, const swe_stats* in);
int swe_sum_plane(swe_engine* e, int plane, double* out);   /* fp64 sum over the map */
int swe_max_excess(swe_engine* e, float* out);
int swe_hash_height(swe_engine* e, uint64_t* out);          /* FNV-1a-64 of fp32 words */
int swe_sync(swe_engine* e);

/* ---- multi-GPU row stripes --------------------------------------------------------------
 * The map is cut into row stripes (x is the slow index, math.h:11-12), one engine per GPU /
 * process. A stripe engine owns the global rows [own_lo, own_hi) (>= 8 rows) and keeps ghost
 * rows of its neighbours; inside swe_erode / swe_cascade_sweep the neighbours exchange boundary
 * rows and hand over particles by kernels writing peer memory (NVLink P2P), synchronised by
 * flags -- no host round trip per exchange. Every rank must issue the same sequence of
 * swe_erode / swe_cascade_sweep / swe_stripe_sync_halos calls with the same arguments
 * (swe_max_excess on a stripe wants a swe_stripe_sync_halos first: it reads one ghost row);
 * `cycles` of swe_erode is the GLOBAL particle count of the call. With the colour-queue schedule (swe_stripe_connect_ring)
 * a stripe's swe_erode is ONE persistent kernel that holds the whole GPU and waits for the other ranks' kernels on the
 * device: do not enqueue another inter-GPU blocking operation (an NCCL collective, ...) on any stream while such calls
 * are in flight -- synchronise the engine's stream first, or the two can wait for each other (the engine's watchdog then
 * ends the call with SWE_ERR_STATE after a few seconds instead of hanging). Host transfers
 * (swe_upload_plane, swe_download_plane, swe_sum_plane, ...) act on the owned rows only. Parallel mode only.
 *   1. swe_stripe_create on every rank (max_particles >= the largest global `cycles`; the same
 *      value on every rank)
 *   2. exchange links: same process -> swe_stripe_export; other processes -> swe_stripe_export_ipc,
 *      ship the SWE_STRIPE_IPC_BYTES blob (e.g. torch.distributed all_gather), swe_stripe_open_ipc
 *   3. swe_stripe_connect(below, above)  (NULL at the map edges)
 *   4. swe_upload_plane(owned rows) on every rank, then swe_stripe_sync_halos */
#define SWE_STRIPE_IPC_BYTES 512
typedef struct swe_stripe_link {   /* a neighbour's buffers as device pointers valid on THIS GPU */
  void* height;
  void* stamp;
  void* arena;
  int32_t x0;              /* global row of the neighbour's first stored row */
  int32_t inbox_capacity;
  /* colour-queue scheduler (swe_stripe_connect_ring): the rank's particle queues and barrier block */
  void* queue_a;
  void* queue_b;
  void* queue_count;
  void* persist;
  int32_t queue_cap;
  int32_t reserved;
} swe_stripe_link;
int swe_stripe_create(swe_engine** out, int nx, int nz, int own_lo, int own_hi, int max_particles,
                      const swe_params* params, int device);
int swe_stripe_export(swe_engine* e, swe_stripe_link* out);
int swe_stripe_export_ipc(swe_engine* e, void* buf, size_t buf_bytes);
int swe_stripe_open_ipc(swe_engine* e, const void* buf, size_t buf_bytes, swe_stripe_link* out);
int swe_stripe_connect(swe_engine* e, const swe_stripe_link* below, const swe_stripe_link* above);
/* Same wiring plus what the colour-queue scheduler needs on stripes: `links` holds the links of ALL n ranks in
 * row order (the entry of this rank, links[rank], is ignored; n <= 8). Neighbours rank-1 / rank+1 serve the halo
 * and hand-off paths as in swe_stripe_connect; in addition every rank's barrier mailbox is mapped, because the
 * persistent flight kernels of all ranks run the same micro-round sequence and meet at a barrier that spans the
 * GPUs (peer stores over NVLink, system-scope release/acquire). With it a stripe engine runs the tile-resident
 * schedule (SWE_OPT_SCHEDULE = 2, the default: global tiling, a tile belongs to the rank that owns its centre row and
 * works on a neighbour's rows in the neighbour's memory, three cross-GPU barriers per super-round; stripes at least
 * two tiles tall, single layer) or the colour-queue schedule (SWE_OPT_SCHEDULE = 0, and whenever the tile schedule
 * does not apply): a particle in a boundary row reads and writes the neighbour's rows of its 3x3 block directly in
 * the neighbour's memory. Both are parity-green at high particle density. NOTE: with the tile schedule swe_erode on
 * a stripe engine is COLLECTIVE AND BLOCKING: every rank must call it with the same arguments, seed and options (the
 * ranks meet on the device), and it returns when the call's last particle has died. `colocated` = stripe engines
 * sharing this GPU (1 with one process per GPU; logical stripes in one process: their number), which caps the
 * co-resident grid of each kernel. Without it (swe_stripe_connect only) a stripe engine runs the lock-step
 * schedule. */
int swe_stripe_connect_ring(swe_engine* e, int rank, int n, const swe_stripe_link* links, int colocated);
int swe_stripe_sync_halos(swe_engine* e);
int swe_stripe_rows(const swe_engine* e, int* own_lo, int* own_hi);
/* rows a stripe engine stores: [x0, x0 + rows) = its owned rows plus ghost rows. swe_upload_rock on a stripe engine
 * takes exactly these rows (the layered extension needs the bedrock under the ghost rows too). */
int swe_stripe_stored_rows(const swe_engine* e, int* x0, int* rows);
/* One host thread driving several stripes (logical stripes on one GPU, or a process that owns
 * several GPUs): same effect as calling swe_erode(e, cycles, SWE_MODE_PARALLEL) /
 * swe_cascade_sweep(e, sweeps, NULL) / swe_stripe_sync_halos(e) on every engine, with the
 * launches interleaved exchange by exchange so no stripe's wait kernels run ahead of the
 * neighbour launches they wait for. `engines` lists the stripes in row order. */
int swe_stripe_group_erode(swe_engine** engines, int n, int cycles);
int swe_stripe_group_sweep(swe_engine** engines, int n, int sweeps);
int swe_stripe_group_sync_halos(swe_engine** engines, int n);

/* ---- device-side plumbing (for a host that already lives on the GPU) ---------------- */
/* Stream the engine launches on (a cudaStream_t); 0 = the engine's own stream. */
int swe_set_stream(swe_engine* e, void* cuda_stream);
void* swe_get_stream(swe_engine* e);
/* Device pointers of the engine's planes: height = float[nx*nz]; field/track =
 * float4[nx*nz] {massflow, momentumx, momentumy, momentumz}. Valid until swe_destroy. */
void* swe_device_height(swe_engine* e);
void* swe_device_field(swe_engine* e);
void* swe_device_track(swe_engine* e);

#ifdef __cplusplus
}
#endif
#endif /* SWE_H */
