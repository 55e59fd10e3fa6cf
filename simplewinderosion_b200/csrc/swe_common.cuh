// swe_common.cuh -- shared device-side types of the B200 wind-erosion engine.
//
// HBM layout (one GPU, or one row stripe of a multi-GPU map):
//   height : float [rows*nz]           4 B/cell   -- quad::cell::height
//   field  : float4[rows*nz]          16 B/cell   -- {massflow, momentumx, momentumy, momentumz}
//   track  : float4[rows*nz]          16 B/cell   -- the four *_track members
// (reference: 36-byte AoS quad::cell, source/cellpool.h:226-240). The four track values of
// a cell are always updated together (wind.hpp:93-96), so they share one 16-byte vector
// that a single red.global.add.v4.f32 updates; height stays a separate dense plane because
// the 3x3 neighbourhood reads (normal + cascade) touch nothing else.
// Index = x*nz + z (math::flatten, source/include/math.h:11-12): x is the slow index.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/swe.h"

namespace swe {

struct MapView {
  float* height;
  float4* field;
  float4* track;
  float* rock;  // layered-terrain extension: bedrock height (sand = height - rock >= 0); nullptr = single layer
  int nx, nz;   // GLOBAL map extent (bounds tests use these: closed boundary, cellpool.h:432-438)
  int x0;       // global x of local row 0 (0 on a single GPU)
  int rows;     // rows stored locally (== nx on a single GPU)
  int periodic; // 1 = the reference's commented periodic boundary (cellpool.h:96-113,447-465): nothing is out of
                // bounds and every cell access wraps (single-GPU engines only)
  __host__ __device__ __forceinline__ static int wrap(int v, int n) {
    v %= n;
    return v < 0 ? v + n : v;
  }
  __host__ __device__ __forceinline__ size_t idx(int x, int z) const {
    if (periodic) { x = wrap(x, nx); z = wrap(z, nz); }
    return (size_t)(x - x0) * (size_t)nz + (size_t)z;
  }
  __host__ __device__ __forceinline__ bool oob(int x, int z) const {
    if (periodic) return false;
    return x < 0 || z < 0 || x >= nx || z >= nz;
  }
};

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization
// attribute may be set up while its predecessor on the stream is still draining; it must not touch
// the predecessor's results before pdl_sync() (a no-op in an ordinary launch). No early trigger
// (griddepcontrol.launch_dependents) is issued: measured on B200, dependents that become resident
// early and spin at the wait cost the 4096^2 / 4.19 M-particle call 10 %, while the implicit trigger at
// kernel exit still saves 2-5 % of a launch-bound call (512^2 / 4096 particles: 10.15 -> 9.76 ms).
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1u : 0u;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// Device-resident counters (mirrors the public swe_stats).
struct DevStats {
  unsigned long long particles;
  unsigned long long steps;
  unsigned long long exits_oob;
  unsigned long long exits_age;
  double sediment_lost;
  unsigned long long cascade_cells;
};

// glibc TYPE_3 rand() state kept on the device between calls (deterministic mode):
// the reference's libc stream is process-global and continues across World::erode calls.
struct DetRand {
  int32_t r[31];
  int32_t f, b;
};

// ---- launchers implemented in the .cu files -------------------------------------------
// deterministic (bit-exact) path: swe_det.cu
void launch_det_erode(const MapView& m, const swe_params& p, int cycles, DetRand* rng,
                      DevStats* stats, cudaStream_t st);
void launch_det_trace(const MapView& m, const swe_params& p, float x, float z, int max_records,
                      float* d_records, int* d_steps, swe_wind* d_final, DetRand* rng,
                      DevStats* stats, cudaStream_t st);
void launch_det_fly_once(const MapView& m, const swe_params& p, swe_wind* d_wind, int* d_alive,
                         DetRand* rng, DevStats* stats, cudaStream_t st);
void launch_det_cascade_one(const MapView& m, const swe_params& p, int x, int z, cudaStream_t st);
void launch_det_srand(DetRand* rng, uint32_t seed, cudaStream_t st);
void launch_normals_rows(const MapView& m, const swe_params& p, int row_lo, int row_hi, float* d_out3,
                         cudaStream_t st);
void launch_vertices_rows(const MapView& m, const swe_params& p, int row_lo, int row_hi, void* d_out48,
                          cudaStream_t st);
void launch_normal_one(const MapView& m, const swe_params& p, int x, int z, float* d_out3,
                       cudaStream_t st);

// whole-map cascade: swe_cascade.cu
// Which rows a tile sweep stores (and therefore which tiles it launches).
struct RowWindow {
  int store_lo, store_hi;   // global rows written back
  int tile_y0;              // index of the first tile row launched
};

// Scratch of the tile sweep: staging plane for the border cells of each tile's output (same
// indexing as the height plane) and one "changed" byte per tile of the whole-map tile grid.
struct TileScratch {
  float* stage;
  unsigned char* flags;
  unsigned char* hint;   // whole-map tile sweeps: 1 = the tile was mostly moving last sweep (skip the pair scan)
};
size_t tile_flag_count(const MapView& m);

// TMA descriptor of the height plane (CUtensorMap, kept opaque here), built once per engine.
struct alignas(64) HeightTensorMap {
  alignas(64) unsigned char opaque[128];        // box of the whole-map sweep tiles
  alignas(64) unsigned char opaque_small[128];  // box of the half-height stamped-sweep tiles
  bool valid;
};
bool make_height_tensor_map(const MapView& m, HeightTensorMap* out);
// variant 0 = fused 9-colour shared-memory tile kernel (TMA tile load when the descriptor is
// valid), 1 = one global-memory launch per colour, 2 = fused tile kernel with plain loads
int launch_cascade_sweep(const MapView& m, const swe_params& p, const HeightTensorMap* tmap, int sweeps,
                         DevStats* stats, cudaStream_t st, int variant, TileScratch scratch);  // returns #launches
// one fused sweep restricted to cells whose stamp byte equals stamp_now (flight pipeline)
// (gate: device counter; the sweep is skipped when *gate == 0, i.e. no particle flew)
void launch_cascade_stamped(const MapView& m, const swe_params& p, const HeightTensorMap* tmap,
                            const unsigned char* stamp, unsigned char stamp_now, const unsigned* gate,
                            int row_lo, int row_hi, TileScratch scratch, cudaStream_t st, bool pdl = false);
void launch_max_excess(const MapView& m, const swe_params& p, float* d_out, cudaStream_t st);

// Per-call values the particle kernels read from device memory instead of their parameter list, so
// that a captured CUDA graph of a whole World::erode call can be replayed for the next call.
struct CallCtl {
  uint32_t key_x, key_y;   // Philox key = seed
  uint32_t call;           // World::erode call index (Philox counter word)
  uint32_t pad;
};

// parallel flight: swe_par.cu
// Particle records of one World::erode call, 32 B each, SoA of two float4:
//   a = {pos.x, pos.y, pos.z, sediment}   b = {speed.x, speed.y, speed.z, age (int bits; <0 = dead)}
// live[r] = particles alive before round r (device counters, live_slots of them).
struct ParticleBuf {
  float4* a;
  float4* b;
  unsigned* live;
  int live_slots;
  int capacity;
};
// Stripe (multi-GPU) round context. Outboxes are PEER memory: the neighbour's inbox for the
// current round parity (device pointers valid on this GPU through CUDA IPC / peer access).
struct StripeOutbox {
  float4* a;
  float4* b;
  uint32_t* pid;
  unsigned* count;
  int capacity;
};
struct StripeRound {
  int own_lo, own_hi;       // global rows this rank owns
  uint32_t* pid;            // ids of the local particles
  unsigned* n_used;         // local particle slots in use
  unsigned* overflow;       // particles dropped for lack of space (must stay 0)
  StripeOutbox below, above;
};
void launch_par_begin_stripe(const ParticleBuf& buf, const StripeRound& sr, const MapView& m, int n,
                             uint64_t seed, uint32_t call, DevStats* stats, cudaStream_t st);
void launch_par_round_stripe(const ParticleBuf& buf, const StripeRound& sr, const MapView& m,
                             const swe_params& p, int steps_per_round, uint64_t seed, uint32_t call,
                             unsigned char* stamp, unsigned char stamp_now, DevStats* stats,
                             cudaStream_t st);
// ctl != nullptr: seed and call index come from device memory (graph replay), else from the arguments
void launch_par_begin(const ParticleBuf& buf, const MapView& m, int n, uint64_t seed, uint32_t call,
                      const CallCtl* ctl, DevStats* stats, cudaStream_t st);
void launch_par_round(const ParticleBuf& buf, const MapView& m, const swe_params& p, int n, int round,
                      int steps_per_round, uint64_t seed, uint32_t call, const CallCtl* ctl,
                      unsigned char* stamp, unsigned char stamp_now, DevStats* stats, cudaStream_t st,
                      bool pdl = false);

// colour-queue scheduler of the single-GPU parallel erode: swe_queue.cu
// 9 colours x 2 parities of particle queues (32-byte records as in ParticleBuf), one 16-bit claim word
// per cell, 18 counters + ticket + dropped (128 B apart).
struct ColourQueues {
  float4* a;
  float4* b;
  unsigned* count;          // [20 * 32]
  unsigned short* claim;    // [cells]  launch-per-micro-round form: 16-bit tags, cleared per call
  unsigned* claim32;        // [ceil(rows/3) * ceil(nz/3)]  persistent form: one word per 3x3 lattice cell, monotonic
                            // 32-bit tags (atomicMax), shared by the nine colours, cleared when the tags would wrap
  void* persist;            // PersistCtl of the persistent kernel (barrier, live count, claim-tag base)
  int cap;                  // records per queue
};
struct MicroRoundArgs {
  int colour;               // colour served by this micro-round
  unsigned pin;             // bit c = parity of colour c's input side
  unsigned index;           // micro-round number within the call
  unsigned tag;             // claim value (1..65535)
  int inj_first, inj_stride, inj_count, n_total;   // staggered injection: ids inj_first + j*inj_stride, j < inj_count
};
void launch_colour_round(const MapView& m, const swe_params& p, const ColourQueues& cq, const MicroRoundArgs& r,
                         uint64_t seed, uint32_t call, const CallCtl* ctl, DevStats* stats, cudaStream_t st, bool pdl,
                         int grid_blocks);
// Stripe context of the persistent kernel: the neighbours' height planes and colour queues and every rank's
// barrier mailbox, as device pointers valid on this GPU (peer mappings).
struct StripeQueues {
  int own_lo, own_hi;                  // global rows this rank owns
  int rank, nranks;
  float* below_h; int below_x0;        // neighbour's height plane and the global row of its first stored row
  float* above_h; int above_x0;
  float4* below_a; float4* below_b; unsigned* below_count; int below_cap;
  float4* above_a; float4* above_b; unsigned* above_count; int above_cap;
  unsigned long long* mailbox[8];      // [rank] -> that rank's mailbox ring: word (micro-round % 16) * 8 + source rank
};
constexpr int SWE_MAX_STRIPE_RANKS = 8;
constexpr size_t SWE_PERSIST_BYTES = 4096;        // PersistCtl at 0, barrier mailbox at 1024
constexpr size_t SWE_MAILBOX_OFFSET = 1024;
constexpr size_t SWE_TILE_MAILBOX_OFFSET = 3072;  // tile scheduler on stripes: its own barrier mailbox ring (16 slots x 8 ranks x 8 bytes)
constexpr size_t SWE_GROUP_OFFSET = 2048;         // copies 1..15 of the grid barrier's release word, 128 B apart
int colour_persistent_blocks_per_sm(bool layered);
int colour_persistent_blocks_per_sm_stripe();
int colour_persistent_threads();
cudaError_t launch_colour_persistent(const MapView& m, const swe_params& p, const ColourQueues& cq, int n_total,
                                     unsigned stagger, unsigned max_micro, uint64_t seed, uint32_t call,
                                     DevStats* stats, cudaStream_t st, int grid_blocks, const StripeQueues* stripe = nullptr,
                                     bool cooperative = true);
void preload_queue();

// tile-resident scheduler of the single-GPU parallel erode: swe_tile.cu
// Particle records sorted by tile (one contiguous segment per tile of the current tiling), ping-pong.
struct TileSched {
  float4* pa[2];            // records {pos.x, pos.y, pos.z, sediment}
  float4* pb[2];            // records {speed.x, speed.y, speed.z, age (int bits; < 0 = dead)}
  unsigned* count;          // [max_tiles + 1] particles per tile of the NEXT tiling (cleared by the scan)
  unsigned* base[2];        // [max_tiles + 1] segment bases (exclusive prefix sums); base[nt] = records in the segments
  unsigned* cursor;         // [max_tiles] scatter cursors
  unsigned* info;           // [0] records placed for the next super-round, [1] records dropped (must stay 0)
  int cap;                  // records per array
  int max_tiles;
};
// Stripe context of the tile scheduler (host side): this rank's rows and the neighbours' planes, record arrays and
// tile tables as device pointers valid on this GPU (peer mappings).
struct TileStripeHost {
  int enabled;
  int own_lo, own_hi, below_lo;        // below_lo = first row the rank below owns
  int has_below, has_above;
  float* below_h; float4* below_t; int below_x0;
  float* above_h; float4* above_t; int above_x0;
  TileSched below, above;
};
int launch_tile_phase(int phase, const MapView& m, const swe_params& p, const TileSched& S, int shape, int s, int K, int pcap,
                      int inj_lo, int inj_hi, uint64_t seed, uint32_t call, DevStats* stats, cudaStream_t st,
                      const TileStripeHost* H);
void launch_tile_xbar(void* const* mailboxes, int rank, int nranks, unsigned long long seq, unsigned* info, bool want_sum, cudaStream_t st);
void tile_owned_rows(int nx, int shape, int tiling, int own_lo, int own_hi, int* a_lo, int* a_hi, int* ntx);
int tile_shapes();                   // tile geometries compiled in (0 = 60 x 60 cells / 128 threads; 256 threads: 1 = 120 x 120, 2 = 240 x 60, 3 = 60 x 240)
int tile_edge_x(int shape);
int tile_edge_z(int shape);
int tile_threads(int shape);
int tile_pick(int nx, int nz, int sm_count, float px, float pz);
int tile_max_tiles(int nx, int nz);
int tile_max_pcap(int shape, bool layered, int smem_optin);
cudaError_t tile_configure(int shape, bool layered, int pcap);
int launch_tile_super_round(const MapView& m, const swe_params& p, const TileSched& S, int shape, int s, int K, int pcap,
                            int inj_lo, int inj_hi, uint64_t seed, uint32_t call, DevStats* stats, cudaStream_t st);
void preload_tile();

// stripe exchange primitives: swe_stripe.cu
void launch_stripe_push(const float* src_h, float* dst_h, const unsigned char* src_s, unsigned char* dst_s,
                        size_t cells, cudaStream_t st);
void launch_stripe_signal(unsigned* peer_flag_a, unsigned* peer_flag_b, unsigned value, cudaStream_t st);
void launch_stripe_wait(const unsigned* flag_a, const unsigned* flag_b, unsigned value, unsigned* error_flag,
                        cudaStream_t st);
void launch_stripe_merge(const float4* in_a, const float4* in_b, const uint32_t* in_pid, unsigned* in_count,
                         int in_capacity, float4* pa, float4* pb, uint32_t* pid, unsigned* n_used,
                         int capacity, unsigned* overflow, unsigned* scratch_base, cudaStream_t st);

// fused exchange: [wait consumed, push both strips, signal ready] and [wait ready, merge both inboxes]
struct StripePush {
  const float* src_h; float* dst_h;
  const unsigned char* src_s; unsigned char* dst_s;
  size_t cells;
};
void launch_stripe_push_signal(const StripePush& down, const StripePush& up, const unsigned* consumed_a,
                               const unsigned* consumed_b, unsigned* ready_a, unsigned* ready_b, unsigned seq,
                               unsigned* done_counter, unsigned* error_flag, cudaStream_t st);
void launch_stripe_wait_merge(const StripeOutbox& in_a, const StripeOutbox& in_b, const unsigned* ready_a,
                              const unsigned* ready_b, unsigned seq, float4* pa, float4* pb, uint32_t* pid,
                              unsigned* n_used, int capacity, unsigned* overflow, unsigned* done_counter,
                              unsigned* error_flag, cudaStream_t st);

// field passes and utilities: swe_fields.cu
void launch_ema(const MapView& m, float lrate, cudaStream_t st);
void launch_pack_plane(const MapView& m, int plane, const float* d_src, cudaStream_t st);
void launch_unpack_plane(const MapView& m, int plane, float* d_dst, cudaStream_t st);
void launch_unpack_field(const MapView& m, bool track, float* d_dst4, cudaStream_t st);
void launch_cells_to_planes(const MapView& m, const float* d_cells, cudaStream_t st);
void launch_planes_to_cells(const MapView& m, float* d_cells, cudaStream_t st);
void launch_sum_plane(const MapView& m, int plane, double* d_out, cudaStream_t st);

// generate: swe_noise.cu
int noise_max_layers();
void launch_noise_stack(const MapView& m, int seed, int layers, int tile, cudaStream_t st);
void launch_height_range(const float* d_height, size_t n, int* d_out2, cudaStream_t st);
void launch_normalize(float* d_height, size_t n, float mn, float mx, cudaStream_t st);

void preload_det();
void preload_cascade();
void preload_fields();
void preload_par();
void preload_stripe();
void preload_noise();

}  // namespace swe
