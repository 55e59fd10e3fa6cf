// swe_api.cu -- the C-ABI (include/swe.h) over the CUDA kernels: engine object, host<->device
// transfer with arbitrary host stride, stream plumbing, statistics. Host-only logic; all
// arithmetic on map data happens in kernels. No CPU fallback anywhere: every entry point
// that touches the map needs the CUDA device the engine was created on.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "swe_common.cuh"

using namespace swe;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define SWE_CUDA(call)                                                                    \
  do {                                                                                    \
    cudaError_t err__ = (call);                                                           \
    if (err__ != cudaSuccess)                                                             \
      return fail(err__ == cudaErrorMemoryAllocation ? SWE_ERR_OOM : SWE_ERR_CUDA,        \
                  "%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), __FILE__,    \
                  __LINE__);                                                              \
  } while (0)

bool params_ok(const swe_params& p, const char** why) {
  if (p.max_age < 0) { *why = "max_age < 0"; return false; }
  if (p.max_age > (1 << 20)) { *why = "max_age > 2^20"; return false; }
  for (float v : {p.boundarylayer, p.suspension, p.gravity, p.lrate, p.maxdiff, p.settling, p.abrasion, p.pspeed[0], p.pspeed[1],
                  p.pspeed[2]})
    if (!std::isfinite(v)) { *why = "a parameter is NaN or infinite"; return false; }
  if (!(p.boundarylayer > 0.0f)) { *why = "boundarylayer must be > 0"; return false; }
  if (p.mapscale <= 0) { *why = "mapscale must be > 0"; return false; }
  const float l2 = p.pspeed[0] * p.pspeed[0] + p.pspeed[1] * p.pspeed[1] + p.pspeed[2] * p.pspeed[2];
  if (!(l2 > 0.0f)) { *why = "pspeed must be non-zero"; return false; }
  if (!(p.abrasion >= 0.0f) || p.abrasion > 1.0f) { *why = "abrasion must be in [0, 1]"; return false; }
  for (int i = 0; i < 4; i++)
    if (p.reserved[i] != 0) { *why = "reserved fields must be zero"; return false; }
  return true;
}

}  // namespace

struct swe_engine {
  int device = 0;
  MapView m{};
  swe_params p{};
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_c = nullptr, ev_d = nullptr, ev_f0 = nullptr, ev_f1 = nullptr;
  bool erode_timed = false, sweep_timed = false;
  DevStats* d_stats = nullptr;
  DetRand* d_rng = nullptr;
  float* d_small = nullptr;     // 64 floats of scratch for scalar results
  float* d_rock = nullptr;      // layered extension: bedrock plane (m.rock points here when enabled)
  float* d_scratch = nullptr;   // lazily sized staging for plane/AoS conversion
  size_t scratch_bytes = 0;
  void* h_stage = nullptr;      // pinned host staging for strided host buffers
  size_t h_stage_bytes = 0;
  uint64_t seed = 1;
  uint32_t call_index = 0;
  // parallel-mode state
  ParticleBuf pbuf{};
  HeightTensorMap tmap{};
  TileScratch tsc{};                  // staging plane + per-tile flags of the cascade tile kernels
  unsigned char* d_stamp = nullptr;   // per-cell "touched in round" byte
  unsigned stamp_counter = 0;
  int opt_cascade_variant = 0;        // 0 fused tile kernel, 1 per-colour launches
  int opt_cascade_interval = 1;       // Wind::fly steps per particle between cascade passes
  int opt_kernel_timing = 0;          // record an event pair around every round / cascade launch
  int opt_cuda_graph = 1;             // replay the parallel erode of a single-GPU engine as one CUDA graph
  int opt_pdl = 1;                    // programmatic dependent launches between the round kernels
  int opt_call_batches = 1;           // a parallel call's particles fly as this many successive batches
  int opt_schedule = 2;               // 2 = tile-resident scheduler (swe_tile.cu; stripes and periodic maps run the colour queues instead),
                                      // 0 = colour-queue scheduler (swe_queue.cu), 1 = lock-step rounds + stamped cascade
  int opt_stagger = -1;               // colour queues: rounds over which the call's particles are injected (-1 = auto)
  ColourQueues cq{};                  // queues + claim plane of the colour-queue scheduler
  int sm_count = 148;
  int last_micro_rounds = 0;          // micro-rounds the last colour-queue call needed
  int last_super_rounds = 0;          // tile scheduler: super-rounds of the last call
  int opt_persistent = 1;             // colour queues: one cooperative launch per call (0 = one launch per micro-round)
  int coop_ok = 0;                    // device supports cooperative launches
  uint64_t claim_tags_used = 0;       // upper bound of the claim tags handed out since the last clearing
  int persist_blocks_per_sm = 0;
  // pipelined frame read-back (swe_download_frame_begin / _wait): private snapshot of the four flow fields, side stream
  float* d_frame = nullptr;
  size_t frame_bytes = 0;
  cudaStream_t frame_stream = nullptr;
  cudaEvent_t ev_frame_src = nullptr, ev_frame_h = nullptr, ev_frame_f = nullptr;
  bool frame_pending = false;
  TileSched ts{};                     // tile-resident scheduler (swe_tile.cu): records sorted by tile
  int opt_tile_rounds = 0;            // rounds (nine micro-rounds each) a CTA keeps its tile per super-round (0 = auto: edge / 5)
  int opt_tile_shape = -1;            // -1 = by map size (tile_pick), else the index of a compiled tile geometry
  int tile_pcap[8][2] = {};           // particle capacity the fly kernel is configured for [shape][layered]
  int smem_optin = 0;
  CallCtl* d_ctl = nullptr;           // seed / call index of the current call (read by graph-replayed kernels)
  struct CachedGraph { uint64_t key; cudaGraphExec_t exec; };
  std::vector<CachedGraph> graphs;
  std::vector<cudaEvent_t> kt_events; // [3*r] before round, [3*r+1] between, [3*r+2] after cascade
  int kt_rounds = 0;
  int kt_rounds_pending = 0;          // stripe rounds recorded so far in the current call
  double last_round_ms = 0.0, last_cascade_ms = 0.0;
  uint64_t launches = 0, erode_calls = 0;
  double last_erode_ms = 0.0, last_sweep_ms = 0.0, last_fly_ms = 0.0;

  // multi-GPU row stripe; disabled on a single-GPU engine
  struct Stripe {
    bool enabled = false, connected = false;
    int own_lo = 0, own_hi = 0;
    int inbox_capacity = 0, max_particles = 0;
    unsigned char* arena = nullptr;      // flags + inboxes: ONE allocation = one IPC handle
    size_t arena_bytes = 0;
    uint32_t* d_pid = nullptr;
    unsigned* d_vars = nullptr;          // [0] n_used [1] overflow [2] error [3] merge base
    bool has_below = false, has_above = false;
    swe_stripe_link below{}, above{};
    unsigned seq = 0;
    std::vector<void*> ipc_opened;       // peer mappings to close at destroy
    bool ring = false;                   // swe_stripe_connect_ring: colour-queue scheduler available
    bool tiles = false;                  // ... and the tile-resident scheduler (queue buffers big enough for its tables)
    TileStripeHost th{};                 // tile scheduler on stripes: neighbours' planes / records / tables (peer pointers)
    TileSched ts{};                      // this rank's records and tables, carved out of the (exported) queue buffers
    unsigned long long tile_seq = 0;     // cross-GPU barriers passed so far (same on every rank)
    int below_rows = 0, above_rows = 0;  // stored rows of the neighbours (their track plane follows their height plane)
    bool merged_planes = false;          // height and track share one allocation (so that one IPC handle maps both)
    int rank = 0, nranks = 1, colocated = 1;
    void* mailboxes[SWE_MAX_STRIPE_RANKS] = {};
    int persist_blocks_per_sm = 0;
  } st;

  size_t cells() const { return (size_t)m.rows * (size_t)m.nz; }           // stored (owned + ghosts)
  int own_lo() const { return st.enabled ? st.own_lo : 0; }
  int own_hi() const { return st.enabled ? st.own_hi : m.nx; }
  size_t host_cells() const { return (size_t)(own_hi() - own_lo()) * (size_t)m.nz; }   // what the host sees
  // view of the rows the host owns: transfers, sums and hashes never touch ghost rows
  MapView owned(int extra_above = 0) const {
    MapView v = m;
    const size_t off = (size_t)(own_lo() - m.x0) * (size_t)m.nz;
    v.height = m.height + off; v.field = m.field + off; v.track = m.track + off;
    v.x0 = own_lo(); v.rows = own_hi() - own_lo() + extra_above;
    return v;
  }
};

// ---- stripe arena layout (identical on every rank, so a peer base pointer is enough) ----------
namespace {
// Ghost rows of a stripe. A row's value after one nine-colour pass depends on at most 4 rows below
// and 6 rows above it (each of the three x-shifts of the colouring can pull in a whole 3-row block,
// depending on the row's alignment mod 3); erode recomputes one extra row each side instead of
// exchanging it, hence 1+4 below and 1+6 above. The rank above therefore receives my last GHOST_LO
// owned rows, the rank below my first GHOST_HI owned rows.
constexpr int GHOST_LO = 5, GHOST_HI = 7;
constexpr size_t ARENA_FLAGS = 256;   // 4 flags, 64 B apart
enum { FLAG_READY_FROM_BELOW = 0, FLAG_READY_FROM_ABOVE = 1, FLAG_CONSUMED_BY_BELOW = 2, FLAG_CONSUMED_BY_ABOVE = 3 };
enum { FROM_BELOW = 0, FROM_ABOVE = 1 };
inline size_t inbox_bytes(int cap) { return 256 + (size_t)cap * (16 + 16 + 4); }
inline size_t arena_size(int cap) { return ARENA_FLAGS + 4 * inbox_bytes(cap); }
inline unsigned* arena_flag(void* base, int which) {
  return reinterpret_cast<unsigned*>(static_cast<unsigned char*>(base) + 64 * which);
}
inline StripeOutbox arena_inbox(void* base, int cap, int dir, int parity) {
  unsigned char* p = static_cast<unsigned char*>(base) + ARENA_FLAGS + (size_t)(dir * 2 + parity) * inbox_bytes(cap);
  StripeOutbox b;
  b.count = reinterpret_cast<unsigned*>(p);
  b.a = reinterpret_cast<float4*>(p + 256);
  b.b = b.a + cap;
  b.pid = reinterpret_cast<uint32_t*>(b.b + cap);
  b.capacity = cap;
  return b;
}
}  // namespace

namespace {

int bind(swe_engine* e) {
  if (!e) return fail(SWE_ERR_INVALID, "engine is NULL");
  SWE_CUDA(cudaSetDevice(e->device));
  return SWE_OK;
}

int need_scratch(swe_engine* e, size_t bytes) {
  if (e->scratch_bytes >= bytes) return SWE_OK;
  if (e->d_scratch) { SWE_CUDA(cudaStreamSynchronize(e->stream)); SWE_CUDA(cudaFree(e->d_scratch)); }
  e->d_scratch = nullptr; e->scratch_bytes = 0;
  SWE_CUDA(cudaMalloc(&e->d_scratch, bytes));
  e->scratch_bytes = bytes;
  return SWE_OK;
}

int need_stage(swe_engine* e, size_t bytes) {
  if (e->h_stage_bytes >= bytes) return SWE_OK;
  if (e->h_stage) { SWE_CUDA(cudaStreamSynchronize(e->stream)); SWE_CUDA(cudaFreeHost(e->h_stage)); }
  e->h_stage = nullptr; e->h_stage_bytes = 0;
  SWE_CUDA(cudaMallocHost(&e->h_stage, bytes));
  e->h_stage_bytes = bytes;
  return SWE_OK;
}

int need_particles(swe_engine* e, int n) {
  const int slots = e->p.max_age + 8;
  if (e->pbuf.capacity >= n && e->pbuf.live_slots >= slots) return SWE_OK;
  SWE_CUDA(cudaStreamSynchronize(e->stream));
  cudaFree(e->pbuf.a); cudaFree(e->pbuf.b); cudaFree(e->pbuf.live);
  e->pbuf = ParticleBuf{};
  SWE_CUDA(cudaMalloc(&e->pbuf.a, (size_t)n * sizeof(float4)));
  SWE_CUDA(cudaMalloc(&e->pbuf.b, (size_t)n * sizeof(float4)));
  SWE_CUDA(cudaMalloc(&e->pbuf.live, (size_t)slots * sizeof(unsigned)));
  e->pbuf.capacity = n; e->pbuf.live_slots = slots;
  return SWE_OK;
}

int need_stamp(swe_engine* e) {
  if (e->d_stamp) return SWE_OK;
  SWE_CUDA(cudaMalloc(&e->d_stamp, e->cells()));
  SWE_CUDA(cudaMemsetAsync(e->d_stamp, 0, e->cells(), e->stream));
  e->stamp_counter = 0;
  return SWE_OK;
}

int need_queues(swe_engine* e, int cycles) {
  const int cap = ((cycles > 1024 ? cycles : 1024) + 255) / 256 * 256;
  if (e->st.enabled) return SWE_OK;   // a stripe's queues are allocated at create (they are exported to the peers)
  if (!e->cq.claim) {
    SWE_CUDA(cudaMalloc(&e->cq.claim, e->cells() * sizeof(unsigned short)));
    SWE_CUDA(cudaMalloc(&e->cq.count, 20 * 32 * sizeof(unsigned)));
    SWE_CUDA(cudaMemsetAsync(e->cq.count, 0, 20 * 32 * sizeof(unsigned), e->stream));
    const size_t lattice = (size_t)((e->m.rows + 2) / 3 + 1) * (size_t)((e->m.nz + 2) / 3);
    SWE_CUDA(cudaMalloc(&e->cq.claim32, lattice * sizeof(unsigned)));
    SWE_CUDA(cudaMemsetAsync(e->cq.claim32, 0, lattice * sizeof(unsigned), e->stream));
    SWE_CUDA(cudaMalloc(&e->cq.persist, SWE_PERSIST_BYTES));
    SWE_CUDA(cudaMemsetAsync(e->cq.persist, 0, SWE_PERSIST_BYTES, e->stream));
  }
  if (e->cq.cap >= cap) return SWE_OK;
  SWE_CUDA(cudaStreamSynchronize(e->stream));
  cudaFree(e->cq.a); cudaFree(e->cq.b);
  e->cq.a = e->cq.b = nullptr; e->cq.cap = 0;
  SWE_CUDA(cudaMalloc(&e->cq.a, 18 * (size_t)cap * sizeof(float4)));
  SWE_CUDA(cudaMalloc(&e->cq.b, 18 * (size_t)cap * sizeof(float4)));
  e->cq.cap = cap;
  return SWE_OK;
}

int check_launch(swe_engine* e, const char* what, int n = 1) {
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return fail(SWE_ERR_CUDA, "%s launch failed: %s", what, cudaGetErrorString(err));
  e->launches += (uint64_t)n;
  return SWE_OK;
}

// One neighbour exchange of a stripe engine (all on the engine's stream, no host sync):
// wait until the neighbours consumed the previous ghosts, push my boundary rows, signal, wait.
// Returns the sequence number; the caller signals "consumed" with it after using the ghosts.
unsigned stripe_exchange_begin(swe_engine* e, bool with_stamps, int* launches, bool merge_inboxes = false,
                               int parity = 0) {
  auto& S = e->st;
  const unsigned s = ++S.seq;
  unsigned* err = S.d_vars + 2;
  const int nz = e->m.nz;
  StripePush down{}, up{};
  if (S.has_below) {   // my first GHOST_HI owned rows are the below neighbour's ghost rows above its stripe
    const size_t src = (size_t)(S.own_lo - e->m.x0) * nz, dst = (size_t)(S.own_lo - S.below.x0) * nz;
    down.src_h = e->m.height + src; down.dst_h = static_cast<float*>(S.below.height) + dst;
    down.src_s = with_stamps ? e->d_stamp + src : nullptr;
    down.dst_s = with_stamps ? static_cast<unsigned char*>(S.below.stamp) + dst : nullptr;
    down.cells = (size_t)GHOST_HI * nz;
  }
  if (S.has_above) {   // my last GHOST_LO owned rows are the above neighbour's ghost rows below its stripe
    const size_t src = (size_t)(S.own_hi - GHOST_LO - e->m.x0) * nz, dst = (size_t)(S.own_hi - GHOST_LO - S.above.x0) * nz;
    up.src_h = e->m.height + src; up.dst_h = static_cast<float*>(S.above.height) + dst;
    up.src_s = with_stamps ? e->d_stamp + src : nullptr;
    up.dst_s = with_stamps ? static_cast<unsigned char*>(S.above.stamp) + dst : nullptr;
    up.cells = (size_t)GHOST_LO * nz;
  }
  // one kernel: wait until the neighbours consumed the previous ghosts, push, signal "delivered"
  launch_stripe_push_signal(down, up,
                            S.has_below ? arena_flag(S.arena, FLAG_CONSUMED_BY_BELOW) : nullptr,
                            S.has_above ? arena_flag(S.arena, FLAG_CONSUMED_BY_ABOVE) : nullptr,
                            S.has_below ? arena_flag(S.below.arena, FLAG_READY_FROM_ABOVE) : nullptr,
                            S.has_above ? arena_flag(S.above.arena, FLAG_READY_FROM_BELOW) : nullptr,
                            s, S.d_vars + 4, err, e->stream);
  // one kernel: wait for the neighbours' deliveries, merge the particles they handed over
  StripeOutbox in_a{}, in_b{};
  if (merge_inboxes && S.has_below) in_a = arena_inbox(S.arena, S.inbox_capacity, FROM_BELOW, parity);
  if (merge_inboxes && S.has_above) in_b = arena_inbox(S.arena, S.inbox_capacity, FROM_ABOVE, parity);
  launch_stripe_wait_merge(in_a, in_b, S.has_below ? arena_flag(S.arena, FLAG_READY_FROM_BELOW) : nullptr,
                           S.has_above ? arena_flag(S.arena, FLAG_READY_FROM_ABOVE) : nullptr, s, e->pbuf.a,
                           e->pbuf.b, S.d_pid, S.d_vars + 0, e->pbuf.capacity, S.d_vars + 1, S.d_vars + 5, err,
                           e->stream);
  *launches += 2;
  return s;
}

void stripe_exchange_end(swe_engine* e, unsigned s, int* launches) {
  auto& S = e->st;
  launch_stripe_signal(S.has_below ? arena_flag(S.below.arena, FLAG_CONSUMED_BY_ABOVE) : nullptr,
                       S.has_above ? arena_flag(S.above.arena, FLAG_CONSUMED_BY_BELOW) : nullptr, s, e->stream);
  *launches += 1;
}

void finish_timers(swe_engine* e);
int erode_epilogue(swe_engine* e);
int erode_prologue(swe_engine* e);

// ---- striped parallel erode, in phases so that a single host thread can interleave several
// in-process stripes round by round (swe_stripe_group_erode) --------------------------------------
int stripe_erode_check(swe_engine* e, int cycles, int mode) {
  if (mode != SWE_MODE_PARALLEL) return fail(SWE_ERR_STATE, "a stripe engine runs parallel mode only");
  if (!e->st.connected) return fail(SWE_ERR_STATE, "swe_stripe_connect has not been called");
  if (cycles > e->st.max_particles)
    return fail(SWE_ERR_INVALID, "cycles %d exceeds the stripe's max_particles %d", cycles, e->st.max_particles);
  return SWE_OK;
}

StripeRound stripe_round_ctx(swe_engine* e) {
  auto& S = e->st;
  StripeRound sr{};
  sr.own_lo = S.own_lo; sr.own_hi = S.own_hi;
  sr.pid = S.d_pid; sr.n_used = S.d_vars + 0; sr.overflow = S.d_vars + 1;
  return sr;
}

bool stripe_uses_tiles(const swe_engine* e) {   // (a layered stripe keeps the colour queues: the bedrock of a neighbour's rows is not mapped)
  return e->st.enabled && e->st.ring && e->st.tiles && e->opt_schedule == 2 && !e->m.rock;
}
bool stripe_uses_colour_queues(const swe_engine* e) {
  return e->st.enabled && e->st.ring && e->opt_schedule != 1 && e->coop_ok && e->st.persist_blocks_per_sm > 0 && !stripe_uses_tiles(e);
}

int stripe_erode_begin(swe_engine* e, int cycles) {
  e->kt_rounds_pending = 0;
  launch_par_begin_stripe(e->pbuf, stripe_round_ctx(e), e->m, cycles, e->seed, e->call_index, e->d_stats,
                          e->stream);
  return check_launch(e, "stripe begin", 3);
}

int stripe_erode_round(swe_engine* e) {
  auto& S = e->st;
  int nl = 0;
  if (++e->stamp_counter > 255u) {
    SWE_CUDA(cudaMemsetAsync(e->d_stamp, 0, e->cells(), e->stream));
    e->stamp_counter = 1;
  }
  const unsigned char now = (unsigned char)e->stamp_counter;
  const int parity = (int)((S.seq + 1) & 1u);
  StripeRound sr = stripe_round_ctx(e);
  if (S.has_below) sr.below = arena_inbox(S.below.arena, S.inbox_capacity, FROM_ABOVE, parity);
  if (S.has_above) sr.above = arena_inbox(S.above.arena, S.inbox_capacity, FROM_BELOW, parity);
  const bool kt = e->opt_kernel_timing != 0;
  const int r = e->kt_rounds_pending;
  if (kt) {
    while ((int)e->kt_events.size() < 3 * (r + 1)) {
      cudaEvent_t ev; SWE_CUDA(cudaEventCreate(&ev)); e->kt_events.push_back(ev);
    }
    SWE_CUDA(cudaEventRecord(e->kt_events[3 * r], e->stream));
  }
  // particles must be handed over before they step again: one step per round
  launch_par_round_stripe(e->pbuf, sr, e->m, e->p, 1, e->seed, e->call_index, e->d_stamp, now, e->d_stats,
                          e->stream);
  nl += 1;
  if (kt) SWE_CUDA(cudaEventRecord(e->kt_events[3 * r + 1], e->stream));
  const unsigned s = stripe_exchange_begin(e, true, &nl, true, parity);
  // rows own_lo-1 .. own_hi: the ghost row each side is recomputed here instead of exchanged
  launch_cascade_stamped(e->m, e->p, &e->tmap, e->d_stamp, now, nullptr, S.own_lo - 1, S.own_hi + 1, e->tsc,
                         e->stream);
  nl += 2;
  stripe_exchange_end(e, s, &nl);
  if (kt) {   // "cascade" share of a stripe round = exchange + merge + cascade
    SWE_CUDA(cudaEventRecord(e->kt_events[3 * r + 2], e->stream));
    e->kt_rounds_pending = r + 1; e->kt_rounds = r + 1;
  }
  return check_launch(e, "stripe round", nl);
}

// The particle phase of a single-GPU parallel World::erode: spawn, then R rounds of
// [particle step kernel, stamped cascade, commit]. Stamp values are per call: round r uses
// 1 + r % 255 and the plane is cleared whenever the value wraps, so the launch sequence depends only
// on (cycles, params, options) and can be captured once and replayed (ctl != nullptr).
int parallel_rounds(swe_engine* e, int cycles, const CallCtl* ctl) {
  const int k = e->opt_cascade_interval;
  const int rounds = (e->p.max_age + 2 + k - 1) / k;   // every particle is dead after maxAge+2 calls
  const bool kt = e->opt_kernel_timing != 0 && ctl == nullptr;
  const bool pdl = e->opt_pdl != 0 && !kt;             // chain the round kernels with programmatic dependent launches
  // SWE_OPT_CALL_BATCHES: the call's particles fly as B successive batches, each spread over the whole
  // map and seeing the terrain the previous batches left (closer to the reference's one-after-another
  // order when a call has many particles per cell). Batch b draws from the RNG stream call + b * 2^24.
  const int batches = e->opt_call_batches < cycles ? e->opt_call_batches : (cycles > 0 ? cycles : 1);
  int g = 0;                                            // rounds launched so far in this call (stamp values)
  for (int b = 0; b < batches; b++) {
    const int n_b = cycles / batches + (b < cycles % batches ? 1 : 0);
    const uint32_t call_arg = (ctl ? 0u : e->call_index) + (uint32_t)b * 0x01000000u;
    launch_par_begin(e->pbuf, e->m, n_b, e->seed, call_arg, ctl, e->d_stats, e->stream);
    if (int rc = check_launch(e, "par_begin", 2)) return rc;
    for (int r = 0; r < rounds; r++, g++) {
      if (g % 255 == 0) SWE_CUDA(cudaMemsetAsync(e->d_stamp, 0, e->cells(), e->stream));
      const unsigned char now = (unsigned char)(1 + g % 255);
      if (kt) {
        while ((int)e->kt_events.size() < 3 * (g + 1)) {
          cudaEvent_t ev; SWE_CUDA(cudaEventCreate(&ev)); e->kt_events.push_back(ev);
        }
        SWE_CUDA(cudaEventRecord(e->kt_events[3 * g], e->stream));
      }
      launch_par_round(e->pbuf, e->m, e->p, n_b, r, k, e->seed, call_arg, ctl, e->d_stamp, now,
                       e->d_stats, e->stream, pdl);
      if (kt) SWE_CUDA(cudaEventRecord(e->kt_events[3 * g + 1], e->stream));
      // wind.hpp:114, deferred: World::cascade once at every cell touched in this round
      launch_cascade_stamped(e->m, e->p, &e->tmap, e->d_stamp, now, e->pbuf.live + r, 0, e->m.nx, e->tsc, e->stream, pdl);
      if (kt) { SWE_CUDA(cudaEventRecord(e->kt_events[3 * g + 2], e->stream)); e->kt_rounds = g + 1; }
      if (int rc = check_launch(e, "par_round", 3)) return rc;
    }
  }
  return SWE_OK;
}

// Whole parallel World::erode (track reset, particle phase, EMA) as ONE graph launch: ~1550 kernel
// and memset nodes captured once per (cycles, params, options) and replayed; the seed and the call
// index travel through CallCtl in device memory.
uint64_t graph_key(const swe_engine* e, int cycles) {
  uint64_t h = 1469598103934665603ull;
  auto mix = [&h](const void* p, size_t n) {
    const unsigned char* b = static_cast<const unsigned char*>(p);
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
  };
  mix(&cycles, sizeof cycles);
  mix(&e->p, sizeof e->p);
  mix(&e->opt_cascade_interval, sizeof(int));
  mix(&e->opt_pdl, sizeof(int));
  mix(&e->opt_call_batches, sizeof(int));
  mix(&e->opt_schedule, sizeof(int));
  mix(&e->opt_stagger, sizeof(int));
  mix(&e->stream, sizeof e->stream);
  mix(&e->m.rock, sizeof e->m.rock);
  mix(&e->pbuf.a, sizeof e->pbuf.a);
  return h;
}

int erode_parallel_graph(swe_engine* e, int cycles) {
  if (int rc = need_particles(e, cycles)) return rc;
  if (int rc = need_stamp(e)) return rc;
  const uint64_t key = graph_key(e, cycles);
  cudaGraphExec_t exec = nullptr;
  for (auto& g : e->graphs) if (g.key == key) exec = g.exec;
  const CallCtl ctl_host{(uint32_t)e->seed, (uint32_t)(e->seed >> 32), e->call_index, 0u};
  SWE_CUDA(cudaMemcpyAsync(e->d_ctl, &ctl_host, sizeof ctl_host, cudaMemcpyHostToDevice, e->stream));
  if (!exec) {
    cudaGraph_t graph = nullptr;
    SWE_CUDA(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
    int rc = SWE_OK;
    if (cudaMemsetAsync(e->m.track, 0, e->cells() * sizeof(float4), e->stream) != cudaSuccess) rc = SWE_ERR_CUDA;
    if (rc == SWE_OK) rc = parallel_rounds(e, cycles, e->d_ctl);
    if (rc == SWE_OK) { launch_ema(e->m, e->p.lrate, e->stream); rc = check_launch(e, "ema"); }
    const cudaError_t end = cudaStreamEndCapture(e->stream, &graph);
    if (rc != SWE_OK || end != cudaSuccess || !graph) {
      if (graph) cudaGraphDestroy(graph);
      (void)cudaGetLastError();
      return rc != SWE_OK ? rc : fail(SWE_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(end));
    }
    const cudaError_t inst = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (inst != cudaSuccess) return fail(SWE_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(inst));
    if (e->graphs.size() >= 4) { cudaGraphExecDestroy(e->graphs.front().exec); e->graphs.erase(e->graphs.begin()); }
    e->graphs.push_back({key, exec});
    // (the capture pass already counted one call's worth of launches)
  } else {
    const int k = e->opt_cascade_interval;
    const uint64_t b = (uint64_t)(e->opt_call_batches < cycles ? e->opt_call_batches : cycles);
    e->launches += 1ull + b * (2ull + 3ull * (uint64_t)((e->p.max_age + 2 + k - 1) / k));   // kernels inside the replayed graph
  }
  SWE_CUDA(cudaEventRecord(e->ev_a, e->stream));
  SWE_CUDA(cudaEventRecord(e->ev_f0, e->stream));
  SWE_CUDA(cudaGraphLaunch(exec, e->stream));
  SWE_CUDA(cudaEventRecord(e->ev_f1, e->stream));
  SWE_CUDA(cudaEventRecord(e->ev_b, e->stream));
  e->erode_timed = true;
  e->call_index++;
  e->erode_calls++;
  return SWE_OK;
}

// ---- colour-queue scheduler (swe_queue.cu) ---------------------------------------------------------
// Stagger: the call's particles are injected evenly over the first S rounds (9 micro-rounds each).
// auto: S = 8 * density * (maxAge + 2) rounds, i.e. 2 x the particle lifetime at 0.25 particles per
// cell (the setting the statistical-parity tests pass with), capped at 4 lifetimes; calls thinner
// than 0.02 particles per cell (the reference's own frames) are injected at once.
int colour_stagger_micro(const swe_engine* e, int cycles) {
  const int life = e->p.max_age + 2;
  long long rounds = e->opt_stagger;
  if (rounds < 0) {
    const double density = (double)cycles / ((double)e->m.nx * (double)e->m.nz);
    rounds = density < 0.02 ? 0 : (long long)(8.0 * density * life + 0.5);
    if (rounds > 4LL * life) rounds = 4LL * life;
  }
  long long micro = 9 * rounds;
  if (micro > cycles) micro = cycles;
  return micro < 1 ? 1 : (int)micro;
}

int colour_main_micro(const swe_engine* e, int cycles) {
  const int k = colour_stagger_micro(e, cycles) + 9 * (e->p.max_age + 2);
  return (k + 17) / 18 * 18;
}

int colour_grid(const swe_engine* e, int cycles) {
  long long blocks = ((long long)cycles / 3 + 255) / 256 + 1;
  const long long cap = (long long)e->sm_count * 6;
  return (int)(blocks > cap ? cap : blocks);
}

// micro-rounds [k0, k1) of a call: colour k % 9, parity of colour c toggles after each of its turns
int colour_rounds(swe_engine* e, int cycles, int k0, int k1, const CallCtl* ctl) {
  const int S = colour_stagger_micro(e, cycles);
  const bool pdl = e->opt_pdl != 0;
  const int grid = colour_grid(e, cycles);
  for (int k = k0; k < k1; k++) {
    MicroRoundArgs r{};
    r.colour = k % 9;
    // colour c has been served floor((k - c + 8) / 9) times before micro-round k
    unsigned pin = 0u;
    for (int c = 0; c < 9; c++) pin |= (unsigned)(((k - c + 8) / 9) & 1) << c;
    r.pin = pin;
    r.index = (unsigned)k;
    r.tag = 1u + (unsigned)(k / 9) % 65535u;
    r.n_total = cycles;
    if (k < S) { r.inj_first = k; r.inj_stride = S; r.inj_count = (cycles - k + S - 1) / S; }
    launch_colour_round(e->m, e->p, e->cq, r, e->seed, ctl ? 0u : e->call_index, ctl, e->d_stats, e->stream, pdl, grid);
  }
  return check_launch(e, "colour_round", k1 - k0);
}

// host-side completion check: a call is over when every injected particle has died
int colour_finish(swe_engine* e, int cycles, int k_done, uint64_t particles_before) {
  for (int guard = 0; guard < 4096; guard++) {
    DevStats s;
    SWE_CUDA(cudaMemcpyAsync(&s, e->d_stats, sizeof s, cudaMemcpyDeviceToHost, e->stream));
    SWE_CUDA(cudaStreamSynchronize(e->stream));
    if (s.particles - particles_before == (uint64_t)cycles && s.particles == s.exits_oob + s.exits_age) {
      e->last_micro_rounds = k_done;
      return SWE_OK;
    }
    if (int rc = colour_rounds(e, cycles, k_done, k_done + 288, nullptr)) return rc;
    k_done += 288;
  }
  return fail(SWE_ERR_STATE, "colour-queue erode did not finish");
}

// one cooperative launch for the whole particle phase; fully asynchronous (the kernel ends itself)
int colour_launch_persistent(swe_engine* e, int cycles) {
  const int nt = colour_persistent_threads();
  long long blocks = ((long long)cycles / 3 + nt - 1) / nt + 1;
  const bool stripe = e->st.enabled;
  long long cap = stripe ? (long long)e->sm_count * e->st.persist_blocks_per_sm / (e->st.colocated > 0 ? e->st.colocated : 1)
                         : (long long)e->sm_count * e->persist_blocks_per_sm;
  if (cap < 1) cap = 1;
  if (blocks > cap) blocks = cap;
  if (const char* dbg = std::getenv("SWE_DEBUG_GRID")) {   // measurement aid: force the grid (barrier cost vs block count)
    const long long g = std::atoll(dbg);
    if (g >= 1 && g <= cap) blocks = g;
  }
  const unsigned stagger = (unsigned)colour_stagger_micro(e, cycles);
  const unsigned max_micro = 64u * (unsigned)colour_main_micro(e, cycles);
  // claim tags grow by at most max_micro + 2 per call: clear the lattice before they could wrap
  if (e->claim_tags_used + max_micro + 2ull >= 0xffff0000ull) {
    const size_t lattice = (size_t)((e->m.rows + 2) / 3 + 1) * (size_t)((e->m.nz + 2) / 3);
    SWE_CUDA(cudaMemsetAsync(e->cq.claim32, 0, lattice * sizeof(unsigned), e->stream));
    SWE_CUDA(cudaMemsetAsync(static_cast<char*>(e->cq.persist) + 260, 0, sizeof(unsigned), e->stream));   // PersistCtl::tag_base
    e->claim_tags_used = 0;
  }
  e->claim_tags_used += max_micro + 2ull;
  StripeQueues sq{};
  if (stripe) {
    auto& S = e->st;
    sq.own_lo = S.own_lo; sq.own_hi = S.own_hi; sq.rank = S.rank; sq.nranks = S.nranks;
    if (S.has_below) {
      sq.below_h = static_cast<float*>(S.below.height); sq.below_x0 = S.below.x0;
      sq.below_a = static_cast<float4*>(S.below.queue_a); sq.below_b = static_cast<float4*>(S.below.queue_b);
      sq.below_count = static_cast<unsigned*>(S.below.queue_count); sq.below_cap = S.below.queue_cap;
    }
    if (S.has_above) {
      sq.above_h = static_cast<float*>(S.above.height); sq.above_x0 = S.above.x0;
      sq.above_a = static_cast<float4*>(S.above.queue_a); sq.above_b = static_cast<float4*>(S.above.queue_b);
      sq.above_count = static_cast<unsigned*>(S.above.queue_count); sq.above_cap = S.above.queue_cap;
    }
    for (int i = 0; i < S.nranks; i++)
      sq.mailbox[i] = reinterpret_cast<unsigned long long*>(static_cast<char*>(S.mailboxes[i]) + SWE_MAILBOX_OFFSET);
  }
  const cudaError_t err = launch_colour_persistent(e->m, e->p, e->cq, cycles, stagger, max_micro, e->seed, e->call_index,
                                                   e->d_stats, e->stream, (int)blocks, stripe ? &sq : nullptr,
                                                   !(stripe && e->st.colocated > 1));
  if (err != cudaSuccess) return fail(SWE_ERR_CUDA, "cooperative launch failed: %s", cudaGetErrorString(err));
  e->launches++;
  e->last_micro_rounds = -1;   // read back lazily (swe_get_stats)
  e->last_super_rounds = 0;
  return SWE_OK;
}

int erode_colour_persistent(swe_engine* e, int cycles) {
  if (int rc = need_queues(e, cycles)) return rc;
  if (int rc = erode_prologue(e)) return rc;
  if (int rc = colour_launch_persistent(e, cycles)) return rc;
  return erode_epilogue(e);
}

int erode_colour(swe_engine* e, int cycles) {
  if (e->opt_persistent && e->coop_ok && e->persist_blocks_per_sm > 0 && !(e->m.rock && colour_persistent_blocks_per_sm(true) < 1))
    return erode_colour_persistent(e, cycles);
  if (int rc = need_queues(e, cycles)) return rc;
  DevStats s0;
  SWE_CUDA(cudaMemcpyAsync(&s0, e->d_stats, sizeof s0, cudaMemcpyDeviceToHost, e->stream));
  SWE_CUDA(cudaStreamSynchronize(e->stream));
  if (s0.particles != s0.exits_oob + s0.exits_age)
    return fail(SWE_ERR_STATE, "particle counters are inconsistent (an earlier call failed?)");
  const int k_main = colour_main_micro(e, cycles);
  const bool use_graph = e->opt_cuda_graph != 0;
  const CallCtl ctl_host{(uint32_t)e->seed, (uint32_t)(e->seed >> 32), e->call_index, 0u};
  SWE_CUDA(cudaMemcpyAsync(e->d_ctl, &ctl_host, sizeof ctl_host, cudaMemcpyHostToDevice, e->stream));
  SWE_CUDA(cudaEventRecord(e->ev_a, e->stream));
  SWE_CUDA(cudaEventRecord(e->ev_f0, e->stream));
  if (use_graph) {
    const uint64_t key = graph_key(e, cycles);
    cudaGraphExec_t exec = nullptr;
    for (auto& g : e->graphs) if (g.key == key) exec = g.exec;
    if (!exec) {
      cudaGraph_t graph = nullptr;
      SWE_CUDA(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
      int rc = SWE_OK;
      if (cudaMemsetAsync(e->m.track, 0, e->cells() * sizeof(float4), e->stream) != cudaSuccess) rc = SWE_ERR_CUDA;
      if (rc == SWE_OK && cudaMemsetAsync(e->cq.claim, 0, e->cells() * sizeof(unsigned short), e->stream) != cudaSuccess) rc = SWE_ERR_CUDA;
      if (rc == SWE_OK) rc = colour_rounds(e, cycles, 0, k_main, e->d_ctl);
      const cudaError_t end = cudaStreamEndCapture(e->stream, &graph);
      if (rc != SWE_OK || end != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        (void)cudaGetLastError();
        return rc != SWE_OK ? rc : fail(SWE_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(end));
      }
      const cudaError_t inst = cudaGraphInstantiate(&exec, graph, 0);
      cudaGraphDestroy(graph);
      if (inst != cudaSuccess) return fail(SWE_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(inst));
      if (e->graphs.size() >= 4) { cudaGraphExecDestroy(e->graphs.front().exec); e->graphs.erase(e->graphs.begin()); }
      e->graphs.push_back({key, exec});
    } else {
      e->launches += (uint64_t)k_main;
    }
    SWE_CUDA(cudaGraphLaunch(exec, e->stream));
  } else {
    SWE_CUDA(cudaMemsetAsync(e->m.track, 0, e->cells() * sizeof(float4), e->stream));
    SWE_CUDA(cudaMemsetAsync(e->cq.claim, 0, e->cells() * sizeof(unsigned short), e->stream));
    if (int rc = colour_rounds(e, cycles, 0, k_main, nullptr)) return rc;
  }
  if (int rc = colour_finish(e, cycles, k_main, s0.particles)) return rc;
  return erode_epilogue(e);
}

// ---- tile-resident scheduler (swe_tile.cu) -----------------------------------------------------------
int need_tiles(swe_engine* e, int cycles) {
  const int cap = ((cycles > 1024 ? cycles : 1024) + 255) / 256 * 256;
  if (!e->ts.count) {
    const int mt = tile_max_tiles(e->m.nx, e->m.nz);
    const size_t words = (size_t)mt + 1;
    SWE_CUDA(cudaMalloc(&e->ts.count, 4 * words * sizeof(unsigned) + 64));
    SWE_CUDA(cudaMemsetAsync(e->ts.count, 0, 4 * words * sizeof(unsigned) + 64, e->stream));
    e->ts.base[0] = e->ts.count + words; e->ts.base[1] = e->ts.count + 2 * words;
    e->ts.cursor = e->ts.count + 3 * words; e->ts.info = e->ts.count + 4 * words;
    e->ts.max_tiles = mt;
  }
  if (e->ts.cap >= cap) return SWE_OK;
  SWE_CUDA(cudaStreamSynchronize(e->stream));
  cudaFree(e->ts.pa[0]);
  e->ts.pa[0] = nullptr; e->ts.cap = 0;
  SWE_CUDA(cudaMalloc(&e->ts.pa[0], 4 * (size_t)cap * sizeof(float4)));
  e->ts.pa[1] = e->ts.pa[0] + cap; e->ts.pb[0] = e->ts.pa[0] + 2 * (size_t)cap; e->ts.pb[1] = e->ts.pa[0] + 3 * (size_t)cap;
  e->ts.cap = cap;
  return SWE_OK;
}

// Whole particle phase on the tile-resident schedule: particle i becomes active at round i * S / n (S = the
// stagger of the colour-queue scheduler), at the granularity of a super-round; super-rounds are enqueued until
// the device reports that no record is left (checked every few super-rounds once the injection is over).
struct TilePlan {
  int shape, K, pcap, inj_srs, first_check, max_srs;
  long long S, n;
  int started_before(long long round) const {             // particles with start round floor(i S / n) < round
    if (S <= 0) return round > 0 ? (int)n : 0;
    const long long v = (round * n + S - 1) / S;
    return (int)(v > n ? n : v);
  }
};

// Geometry, rounds per super-round, on-chip list capacity and the number of super-rounds to enqueue before the first
// completion check. Depends only on the (global) map, the parameters and the options: every rank of a stripe group
// computes the same plan.
int tile_plan(swe_engine* e, int cycles, TilePlan* P) {
  const bool layered = e->m.rock != nullptr;
  // (a layered engine keeps two planes per tile in shared memory: only ONE 120 x 120 tile would fit an SM)
  int shape = e->opt_tile_shape >= 0 ? e->opt_tile_shape
                                     : (layered ? 0 : tile_pick(e->m.nx, e->m.nz, e->sm_count, e->p.pspeed[0], e->p.pspeed[2]));
  if (e->st.enabled && e->opt_tile_shape < 0) {
    // stripes must be two tiles tall: step down to a shorter tile when they are not (decided from the nominal stripe
    // height nx / ranks, which every rank knows, so that all ranks pick the same geometry)
    const int nominal = e->m.nx / (e->st.nranks > 0 ? e->st.nranks : 1) - 8;
    if (nominal < 2 * tile_edge_x(shape)) shape = nominal >= 2 * tile_edge_x(1) && shape != 0 ? 1 : 0;
  }
  const int tile_cells = tile_edge_x(shape) * tile_edge_z(shape);
  const int K = e->opt_tile_rounds > 0 ? e->opt_tile_rounds : (tile_cells > 3600 ? 24 : 16);   // (A/B on B200; tools/par_emul.c mode D for the parity)
  const int life = e->p.max_age + 2;
  const long long S = colour_stagger_micro(e, cycles) / 9;      // rounds
  // expected particles per tile while the injection runs (Poisson): mean + 5 sigma + slack stays on chip; a tile
  // with more than that parks the excess for a super-round (never lost)
  const double density = (double)cycles / ((double)e->m.nx * (double)e->m.nz);
  const double lam = density * tile_cells * (S > 0 ? (double)life / (double)(S + life) : 1.0);
  int pcap = ((int)(lam + 5.0 * std::sqrt(lam) + 64.0) + 63) / 64 * 64;
  if (const char* dbg = std::getenv("SWE_DEBUG_TILE_PCAP")) { const int v = std::atoi(dbg); if (v >= 64) pcap = v / 64 * 64; }
  const int pmax = tile_max_pcap(shape, layered, e->smem_optin > 0 ? e->smem_optin : 49152);
  if (pmax < 64) return fail(SWE_ERR_STATE, "tile scheduler: not enough shared memory per block");
  if (pcap < 128) pcap = 128;
  if (pcap > pmax) pcap = pmax;
  if (e->tile_pcap[shape][layered] != pcap) {
    const cudaError_t err = tile_configure(shape, layered, pcap);
    if (err != cudaSuccess) return fail(SWE_ERR_CUDA, "tile kernel configuration failed: %s", cudaGetErrorString(err));
    e->tile_pcap[shape][layered] = pcap;
  }
  P->shape = shape; P->K = K; P->pcap = pcap; P->S = S; P->n = cycles;
  P->inj_srs = S > 0 ? (int)((S + K - 1) / K) : 1;         // super-rounds that receive new particles
  P->first_check = P->inj_srs + (life * 6 / 10 + K - 1) / K;
  P->max_srs = P->inj_srs + 64 * ((life + K - 1) / K) + 64;
  return SWE_OK;
}

int erode_tile(swe_engine* e, int cycles) {
  if (int rc = need_tiles(e, cycles)) return rc;
  TilePlan P;
  if (int rc = tile_plan(e, cycles, &P)) return rc;
  const int K = P.K;
  if (int rc = erode_prologue(e)) return rc;
  int nl = launch_tile_super_round(e->m, e->p, e->ts, P.shape, -1, K, P.pcap, 0, P.started_before(K), e->seed, e->call_index,
                                   e->d_stats, e->stream);
  int s = 0;
  for (;; s++) {
    if (s >= P.max_srs) return fail(SWE_ERR_STATE, "tile scheduler: particles still alive after %d super-rounds", s);
    const int lo = P.started_before((long long)(s + 1) * K), hi = P.started_before((long long)(s + 2) * K);
    nl += launch_tile_super_round(e->m, e->p, e->ts, P.shape, s, K, P.pcap, lo, hi, e->seed, e->call_index, e->d_stats, e->stream);
    if (s + 1 >= P.first_check && (s + 1 - P.first_check) % 4 == 0) {
      if (int rc = check_launch(e, "tile super-round", nl)) return rc;
      nl = 0;
      unsigned info[2];
      SWE_CUDA(cudaMemcpyAsync(info, e->ts.info, sizeof info, cudaMemcpyDeviceToHost, e->stream));
      SWE_CUDA(cudaStreamSynchronize(e->stream));
      if (info[1] != 0u) return fail(SWE_ERR_STATE, "tile scheduler dropped %u particle records", info[1]);
      if (info[0] == 0u && s + 1 >= P.inj_srs) { s++; break; }
    }
  }
  e->last_micro_rounds = s * 9 * K;
  e->last_super_rounds = s;
  return erode_epilogue(e);
}

// The same schedule on the stripes of ONE map (n engines driven by this host thread: 1 with one process per GPU,
// several for logical stripes). Every rank launches the tiles it owns (csrc/swe_tile.cu: TileStripe); a cross-GPU
// barrier kernel follows each phase of a super-round; the completion check reads the GLOBAL number of records,
// which the barrier after the scan sums over all ranks, so every rank takes the same decision.
int erode_tile_stripes(swe_engine** engines, int n, int cycles) {
  if (n < 1 || n > SWE_MAX_STRIPE_RANKS) return fail(SWE_ERR_INVALID, "a stripe group on the tile schedule has 1..%d engines", SWE_MAX_STRIPE_RANKS);
  TilePlan P{};
  for (int i = 0; i < n; i++) {
    swe_engine* e = engines[i];
    if (int rc = bind(e)) return rc;
    TilePlan Q;
    if (int rc = tile_plan(e, cycles, &Q)) return rc;
    if (i == 0) P = Q;
    else if (Q.shape != P.shape || Q.K != P.K || Q.S != P.S) return fail(SWE_ERR_STATE, "stripes disagree on the tile plan");
    if (e->cq.cap < 1 || e->st.ts.cap < 1) return fail(SWE_ERR_STATE, "stripe engine has no tile tables");
    if (e->st.own_hi - e->st.own_lo < 2 * tile_edge_x(Q.shape) || e->m.nx / e->st.nranks - 8 < 2 * tile_edge_x(Q.shape))
      return fail(SWE_ERR_STATE, "tile schedule on stripes: every stripe must be at least two tiles (%d rows) tall; "
                                 "use SWE_OPT_TILE_SHAPE = 0 or SWE_OPT_SCHEDULE = 0", 2 * tile_edge_x(Q.shape));
    if (int rc = erode_prologue(e)) return rc;
    const size_t words = (size_t)e->st.ts.max_tiles + 1;
    SWE_CUDA(cudaMemsetAsync(e->st.ts.count, 0, (4 * words + 16) * sizeof(unsigned), e->stream));
  }
  const int K = P.K;
  int nl = 0;
  auto round = [&](int s, int lo, int hi, bool want_sum) -> int {
    for (int phase = 0; phase < 3; phase++)
      for (int i = 0; i < n; i++) {
        swe_engine* e = engines[i];
        if (int rc = bind(e)) return rc;
        nl += launch_tile_phase(phase, e->m, e->p, e->st.ts, P.shape, s, K, P.pcap, lo, hi, e->seed, e->call_index, e->d_stats,
                                e->stream, &e->st.th);
        void* boxes[SWE_MAX_STRIPE_RANKS] = {};
        for (int r = 0; r < e->st.nranks; r++) boxes[r] = static_cast<char*>(e->st.mailboxes[r]) + SWE_TILE_MAILBOX_OFFSET;
        launch_tile_xbar(boxes, e->st.rank, e->st.nranks, ++e->st.tile_seq, e->st.ts.info, want_sum && phase == 1, e->stream);
        nl++;
      }
    return SWE_OK;
  };
  if (int rc = round(-1, 0, P.started_before(K), false)) return rc;
  int s = 0;
  for (;; s++) {
    if (s >= P.max_srs) return fail(SWE_ERR_STATE, "tile scheduler: particles still alive after %d super-rounds", s);
    const int lo = P.started_before((long long)(s + 1) * K), hi = P.started_before((long long)(s + 2) * K);
    const bool check = s + 1 >= P.first_check && (s + 1 - P.first_check) % 4 == 0;
    if (int rc = round(s, lo, hi, check)) return rc;
    if (check) {
      unsigned info[SWE_MAX_STRIPE_RANKS][4];
      for (int i = 0; i < n; i++) {
        if (int rc = bind(engines[i])) return rc;
        if (int rc = check_launch(engines[i], "tile super-round (stripes)", i == 0 ? nl : 0)) return rc;
        SWE_CUDA(cudaMemcpyAsync(info[i], engines[i]->st.ts.info, sizeof info[i], cudaMemcpyDeviceToHost, engines[i]->stream));
      }
      nl = 0;
      for (int i = 0; i < n; i++) {
        if (int rc = bind(engines[i])) return rc;
        SWE_CUDA(cudaStreamSynchronize(engines[i]->stream));
      }
      for (int i = 0; i < n; i++) {
        if (info[i][3] != 0u) return fail(SWE_ERR_STATE, "tile scheduler: cross-GPU barrier watchdog fired on stripe %d", engines[i]->st.rank);
        if (info[i][1] != 0u) return fail(SWE_ERR_STATE, "tile scheduler dropped %u particle records on stripe %d", info[i][1], engines[i]->st.rank);
      }
      if (info[0][2] == 0u && s + 1 >= P.inj_srs) { s++; break; }
    }
  }
  for (int i = 0; i < n; i++) {
    engines[i]->last_micro_rounds = s * 9 * K;
    engines[i]->last_super_rounds = s;
  }
  // the tiles worked on the owners' memory; bring every rank's ghost rows up to date for what follows
  if (n > 1) {
    if (int rc = swe_stripe_group_sync_halos(engines, n)) return rc;
  } else {
    int nh = 0;
    const unsigned q = stripe_exchange_begin(engines[0], false, &nh);
    stripe_exchange_end(engines[0], q, &nh);
    if (int rc = check_launch(engines[0], "stripe halo refresh", nh)) return rc;
  }
  for (int i = 0; i < n; i++) {
    if (int rc = bind(engines[i])) return rc;
    if (int rc = erode_epilogue(engines[i])) return rc;
  }
  return SWE_OK;
}

int erode_prologue(swe_engine* e) {   // (the timers of an earlier call are read at the next synchronising entry point)
  SWE_CUDA(cudaEventRecord(e->ev_a, e->stream));
  SWE_CUDA(cudaMemsetAsync(e->m.track, 0, e->cells() * sizeof(float4), e->stream));   // world.hpp:38-44
  SWE_CUDA(cudaEventRecord(e->ev_f0, e->stream));
  return SWE_OK;
}

int erode_epilogue(swe_engine* e) {
  SWE_CUDA(cudaEventRecord(e->ev_f1, e->stream));
  launch_ema(e->m, e->p.lrate, e->stream);                                              // world.hpp:70-76
  if (int rc = check_launch(e, "ema")) return rc;
  SWE_CUDA(cudaEventRecord(e->ev_b, e->stream));
  e->erode_timed = true;
  e->call_index++;
  e->erode_calls++;
  return SWE_OK;
}

// Reads the event pairs of the last erode / sweep. Called from entry points that may run while that work is still
// in flight (the next swe_erode is enqueued behind it): a pair whose closing event has not completed yet is left
// for a later call instead of asking cudaEventElapsedTime for it (which would answer cudaErrorNotReady).
void finish_timers(swe_engine* e) {
  if (e->erode_timed && cudaEventQuery(e->ev_b) != cudaSuccess) { (void)cudaGetLastError(); return; }
  if (e->sweep_timed && cudaEventQuery(e->ev_d) != cudaSuccess) { (void)cudaGetLastError(); return; }
  float ms = 0.0f;
  if (e->erode_timed && cudaEventElapsedTime(&ms, e->ev_a, e->ev_b) == cudaSuccess) e->last_erode_ms = ms;
  if (e->erode_timed && cudaEventElapsedTime(&ms, e->ev_f0, e->ev_f1) == cudaSuccess) e->last_fly_ms = ms;
  if (e->sweep_timed && cudaEventElapsedTime(&ms, e->ev_c, e->ev_d) == cudaSuccess) e->last_sweep_ms = ms;
  if (e->erode_timed && e->kt_rounds > 0) {
    double r_ms = 0.0, c_ms = 0.0;
    for (int r = 0; r < e->kt_rounds; r++) {
      if (cudaEventElapsedTime(&ms, e->kt_events[3 * r], e->kt_events[3 * r + 1]) == cudaSuccess) r_ms += ms;
      if (cudaEventElapsedTime(&ms, e->kt_events[3 * r + 1], e->kt_events[3 * r + 2]) == cudaSuccess) c_ms += ms;
    }
    e->last_round_ms = r_ms; e->last_cascade_ms = c_ms;
    e->kt_rounds = 0;
  }
  e->erode_timed = e->sweep_timed = false;
}

}  // namespace

extern "C" {

int swe_abi_version(void) { return SWE_ABI_VERSION; }

int swe_tile_geometry(int shape, int* edge_x, int* edge_z, int* threads) {
  if (shape < 0 || shape >= tile_shapes()) return fail(SWE_ERR_INVALID, "tile shape must be in [0, %d)", tile_shapes());
  if (edge_x) *edge_x = tile_edge_x(shape);
  if (edge_z) *edge_z = tile_edge_z(shape);
  if (threads) *threads = tile_threads(shape);
  return SWE_OK;
}

int swe_tile_owned_rows(int nx, int shape, int tiling, int own_lo, int own_hi, int* first, int* end, int* total) {
  if (shape < 0 || shape >= tile_shapes() || tiling < 0 || nx < 1 || own_lo < 0 || own_hi > nx || own_lo >= own_hi || !first || !end || !total)
    return fail(SWE_ERR_INVALID, "bad argument");
  tile_owned_rows(nx, shape, tiling, own_lo, own_hi, first, end, total);
  return SWE_OK;
}
const char* swe_last_error(void) { return g_err.c_str(); }

int swe_params_default(swe_params* p) {
  if (!p) return fail(SWE_ERR_INVALID, "params is NULL");
  std::memset(p, 0, sizeof *p);
  p->max_age = 512;            // wind.hpp:23
  p->boundarylayer = 2.0f;     // wind.hpp:24
  p->suspension = 0.05f;       // wind.hpp:25
  p->gravity = (float)0.1;     // wind.hpp:26
  p->lrate = 0.1f;             // world.hpp:30
  p->maxdiff = (float)0.005;   // world.hpp:31
  p->settling = (float)1.0;    // world.hpp:32
  p->mapscale = 80;            // cellpool.h:184
  p->pspeed[0] = 1.0f; p->pspeed[1] = 0.0f; p->pspeed[2] = 0.0f;  // wind.hpp:13
  return SWE_OK;
}

static int create_engine(swe_engine** out, int nx, int nz, int own_lo, int own_hi, bool stripe,
                         int max_particles, const swe_params* params, int device) {
  if (!out) return fail(SWE_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (nx < 1 || nz < 1) return fail(SWE_ERR_INVALID, "map size must be >= 1x1 (got %dx%d)", nx, nz);
  if ((long long)nx * nz > (1LL << 31) - 1)
    return fail(SWE_ERR_INVALID, "map of %dx%d cells exceeds 2^31-1 cells", nx, nz);
  if (stripe) {
    if (own_lo < 0 || own_hi > nx || own_hi - own_lo < 8)
      return fail(SWE_ERR_INVALID, "stripe rows [%d,%d) must lie in [0,%d) and span >= 8 rows", own_lo, own_hi, nx);
    if (max_particles < 1) return fail(SWE_ERR_INVALID, "max_particles must be >= 1");
  }
  swe_params p;
  if (params) p = *params; else swe_params_default(&p);
  const char* why = nullptr;
  if (!params_ok(p, &why)) return fail(SWE_ERR_INVALID, "bad params: %s", why);
  int ndev = 0;
  cudaError_t derr = cudaGetDeviceCount(&ndev);
  if (derr != cudaSuccess || ndev == 0)
    return fail(SWE_ERR_CUDA, "no CUDA device available (%s); this engine has no CPU fallback",
                derr == cudaSuccess ? "device count is 0" : cudaGetErrorString(derr));
  if (device < 0 || device >= ndev) return fail(SWE_ERR_INVALID, "device %d out of range [0,%d)", device, ndev);
  SWE_CUDA(cudaSetDevice(device));

  swe_engine* e = new (std::nothrow) swe_engine();
  if (!e) return fail(SWE_ERR_OOM, "host allocation failed");
  e->device = device; e->p = p;
  e->m.nx = nx; e->m.nz = nz; e->m.x0 = 0; e->m.rows = nx;
  if (stripe) {   // stored rows = owned + GHOST_LO rows below + GHOST_HI above (clipped to the map)
    e->st.enabled = true; e->st.own_lo = own_lo; e->st.own_hi = own_hi;
    e->m.x0 = own_lo - GHOST_LO < 0 ? 0 : own_lo - GHOST_LO;
    const int hi = own_hi + GHOST_HI > nx ? nx : own_hi + GHOST_HI;
    e->m.rows = hi - e->m.x0;
    e->st.max_particles = max_particles;
    e->st.inbox_capacity = max_particles / 8 + 4096;
    e->st.arena_bytes = arena_size(e->st.inbox_capacity);
  }
  const size_t n = e->cells();
  cudaError_t err;
#define TRY(call) if ((err = (call)) != cudaSuccess) goto cuda_fail
  TRY(cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking));
  e->stream = e->own_stream;
  TRY(cudaEventCreate(&e->ev_a)); TRY(cudaEventCreate(&e->ev_b));
  TRY(cudaEventCreate(&e->ev_c)); TRY(cudaEventCreate(&e->ev_d));
  TRY(cudaEventCreate(&e->ev_f0)); TRY(cudaEventCreate(&e->ev_f1));
  if (stripe) {
    // one allocation for height + track: the neighbours map it through ONE IPC handle and reach the track vectors of
    // the rows next to the stripe boundary (tile scheduler: a tile that straddles the boundary adds them in place)
    const size_t hbytes = (n * sizeof(float) + 255) / 256 * 256;
    TRY(cudaMalloc(&e->m.height, hbytes + n * sizeof(float4)));
    e->m.track = reinterpret_cast<float4*>(reinterpret_cast<char*>(e->m.height) + hbytes);
    e->st.merged_planes = true;
  } else {
    TRY(cudaMalloc(&e->m.height, n * sizeof(float)));
    TRY(cudaMalloc(&e->m.track, n * sizeof(float4)));
  }
  TRY(cudaMalloc(&e->m.field, n * sizeof(float4)));
  TRY(cudaMalloc(&e->d_stats, sizeof(DevStats)));
  TRY(cudaMalloc(&e->d_rng, sizeof(DetRand)));
  TRY(cudaMalloc(&e->d_small, 64 * sizeof(float)));
  TRY(cudaMalloc(&e->d_ctl, sizeof(CallCtl)));
  TRY(cudaMemsetAsync(e->m.height, 0, n * sizeof(float), e->stream));
  TRY(cudaMemsetAsync(e->m.field, 0, n * sizeof(float4), e->stream));
  TRY(cudaMemsetAsync(e->m.track, 0, n * sizeof(float4), e->stream));
  TRY(cudaMalloc(&e->tsc.stage, n * sizeof(float)));
  TRY(cudaMalloc(&e->tsc.flags, 2 * tile_flag_count(e->m)));                    // flags, then hints
  TRY(cudaMemsetAsync(e->tsc.flags, 0, 2 * tile_flag_count(e->m), e->stream));
  e->tsc.hint = e->tsc.flags + tile_flag_count(e->m);
  TRY(cudaMemsetAsync(e->d_stats, 0, sizeof(DevStats), e->stream));
  if (stripe) {
    TRY(cudaMalloc(&e->d_stamp, n));
    TRY(cudaMemsetAsync(e->d_stamp, 0, n, e->stream));
    TRY(cudaMalloc(&e->st.arena, e->st.arena_bytes));
    TRY(cudaMemsetAsync(e->st.arena, 0, e->st.arena_bytes, e->stream));
    {
      // local particle slots: hand-offs append (no slot reuse within a call), so thin stripes,
      // which every particle crosses many times, get the full budget; tall stripes twice their
      // share of the call plus slack. swe_stats.stripe_overflow reports any shortfall.
      const int own_rows = own_hi - own_lo;
      long long cap = max_particles;
      if (own_rows >= 512) {
        cap = 2LL * (long long)((double)max_particles * own_rows / nx) + 262144;
        if (cap > max_particles) cap = max_particles;
      }
      e->pbuf.capacity = (int)cap;
    }
    TRY(cudaMalloc(&e->st.d_pid, (size_t)e->pbuf.capacity * sizeof(uint32_t)));
    TRY(cudaMalloc(&e->st.d_vars, 64));
    TRY(cudaMemsetAsync(e->st.d_vars, 0, 64, e->stream));
    TRY(cudaMalloc(&e->pbuf.a, (size_t)e->pbuf.capacity * sizeof(float4)));
    TRY(cudaMalloc(&e->pbuf.b, (size_t)e->pbuf.capacity * sizeof(float4)));
    e->pbuf.live = nullptr; e->pbuf.live_slots = 0;
    {
      // colour-queue scheduler on stripes: the queues must exist before the links are exported
      const int qcap = ((e->pbuf.capacity > 1024 ? e->pbuf.capacity : 1024) + 255) / 256 * 256;
      const size_t lattice = (size_t)((e->m.rows + 2) / 3 + 1) * (size_t)((e->m.nz + 2) / 3);
      TRY(cudaMalloc(&e->cq.a, 18 * (size_t)qcap * sizeof(float4)));
      TRY(cudaMalloc(&e->cq.b, 18 * (size_t)qcap * sizeof(float4)));
      TRY(cudaMalloc(&e->cq.count, 20 * 32 * sizeof(unsigned)));
      TRY(cudaMemsetAsync(e->cq.count, 0, 20 * 32 * sizeof(unsigned), e->stream));
      TRY(cudaMalloc(&e->cq.claim32, lattice * sizeof(unsigned)));
      TRY(cudaMemsetAsync(e->cq.claim32, 0, lattice * sizeof(unsigned), e->stream));
      TRY(cudaMalloc(&e->cq.persist, SWE_PERSIST_BYTES));
      TRY(cudaMemsetAsync(e->cq.persist, 0, SWE_PERSIST_BYTES, e->stream));
      e->cq.cap = qcap;
    }
  }
  make_height_tensor_map(e->m, &e->tmap);   // optional: plain-load kernels are used when it fails
  preload_det(); preload_cascade(); preload_fields(); preload_par(); preload_stripe(); preload_noise(); preload_queue(); preload_tile();
  if (cudaDeviceGetAttribute(&e->sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || e->sm_count < 1) e->sm_count = 148;
  if (cudaDeviceGetAttribute(&e->coop_ok, cudaDevAttrCooperativeLaunch, device) != cudaSuccess) e->coop_ok = 0;
  e->persist_blocks_per_sm = colour_persistent_blocks_per_sm(false);
  if (cudaDeviceGetAttribute(&e->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) != cudaSuccess) e->smem_optin = 49152;
  (void)cudaGetLastError();
  (void)cudaGetLastError();
  launch_det_srand(e->d_rng, 1u, e->stream);
  TRY(cudaGetLastError());
  e->launches++;
  TRY(cudaStreamSynchronize(e->stream));
#undef TRY
  *out = e;
  return SWE_OK;
cuda_fail:
  {
    const int code = fail(err == cudaErrorMemoryAllocation ? SWE_ERR_OOM : SWE_ERR_CUDA,
                          "swe_create(%d x %d) failed: %s", nx, nz, cudaGetErrorString(err));
    std::string keep = g_err;
    swe_destroy(e);
    g_err = keep;
    return code;
  }
}

int swe_create(swe_engine** out, int nx, int nz, const swe_params* params, int device) {
  return create_engine(out, nx, nz, 0, nx, false, 0, params, device);
}

int swe_stripe_create(swe_engine** out, int nx, int nz, int own_lo, int own_hi, int max_particles,
                      const swe_params* params, int device) {
  return create_engine(out, nx, nz, own_lo, own_hi, true, max_particles, params, device);
}

void swe_destroy(swe_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  cudaFree(e->m.height); cudaFree(e->m.field);
  if (!e->st.merged_planes) cudaFree(e->m.track);
  cudaFree(e->d_stats); cudaFree(e->d_rng); cudaFree(e->d_small); cudaFree(e->d_scratch);
  cudaFree(e->pbuf.a); cudaFree(e->pbuf.b); cudaFree(e->pbuf.live); cudaFree(e->d_stamp);
  cudaFree(e->st.arena); cudaFree(e->st.d_pid); cudaFree(e->st.d_vars);
  cudaFree(e->tsc.stage); cudaFree(e->tsc.flags); cudaFree(e->d_rock); cudaFree(e->d_ctl);
  cudaFree(e->cq.a); cudaFree(e->cq.b); cudaFree(e->cq.count); cudaFree(e->cq.claim);
  cudaFree(e->cq.claim32); cudaFree(e->cq.persist);
  cudaFree(e->ts.count); cudaFree(e->ts.pa[0]);
  if (e->frame_stream) { cudaStreamSynchronize(e->frame_stream); cudaStreamDestroy(e->frame_stream); }
  cudaFree(e->d_frame);
  if (e->ev_frame_src) cudaEventDestroy(e->ev_frame_src);
  if (e->ev_frame_h) cudaEventDestroy(e->ev_frame_h);
  if (e->ev_frame_f) cudaEventDestroy(e->ev_frame_f);
  for (auto& g : e->graphs) cudaGraphExecDestroy(g.exec);
  for (void* p : e->st.ipc_opened) cudaIpcCloseMemHandle(p);
  if (e->h_stage) cudaFreeHost(e->h_stage);
  if (e->ev_a) cudaEventDestroy(e->ev_a);
  if (e->ev_b) cudaEventDestroy(e->ev_b);
  if (e->ev_c) cudaEventDestroy(e->ev_c);
  if (e->ev_d) cudaEventDestroy(e->ev_d);
  if (e->ev_f0) cudaEventDestroy(e->ev_f0);
  if (e->ev_f1) cudaEventDestroy(e->ev_f1);
  for (cudaEvent_t ev : e->kt_events) cudaEventDestroy(ev);
  if (e->own_stream) cudaStreamDestroy(e->own_stream);
  delete e;
}

int swe_set_params(swe_engine* e, const swe_params* params) {
  if (!e || !params) return fail(SWE_ERR_INVALID, "NULL argument");
  const char* why = nullptr;
  if (!params_ok(*params, &why)) return fail(SWE_ERR_INVALID, "bad params: %s", why);
  e->p = *params;
  return SWE_OK;
}

int swe_get_params(const swe_engine* e, swe_params* out) {
  if (!e || !out) return fail(SWE_ERR_INVALID, "NULL argument");
  *out = e->p;
  return SWE_OK;
}

int swe_get_size(const swe_engine* e, int* nx, int* nz) {
  if (!e) return fail(SWE_ERR_INVALID, "engine is NULL");
  if (nx) *nx = e->m.nx;
  if (nz) *nz = e->m.nz;
  return SWE_OK;
}

int swe_srand(swe_engine* e, uint32_t seed) {
  if (int rc = bind(e)) return rc;
  e->seed = seed; e->call_index = 0;
  launch_det_srand(e->d_rng, seed, e->stream);
  return check_launch(e, "srand");
}

// ---- run state (checkpoint / resume) ----------------------------------------------------------
namespace {
struct RunState {                 // the SWE_STATE_BYTES blob
  char magic[8];                  // "SWESTAT1"
  uint64_t seed;
  uint32_t call_index;
  uint32_t pad;
  DetRand rng;                    // glibc TYPE_3 ring of the deterministic mode
};
static_assert(sizeof(RunState) <= SWE_STATE_BYTES, "state blob too small");
}  // namespace

int swe_export_state(swe_engine* e, void* buf, size_t bytes) {
  if (int rc = bind(e)) return rc;
  if (!buf || bytes < SWE_STATE_BYTES) return fail(SWE_ERR_INVALID, "state buffer must hold SWE_STATE_BYTES");
  RunState st{};
  memcpy(st.magic, "SWESTAT1", 8);
  st.seed = e->seed; st.call_index = e->call_index;
  SWE_CUDA(cudaMemcpyAsync(&st.rng, e->d_rng, sizeof(DetRand), cudaMemcpyDeviceToHost, e->stream));
  SWE_CUDA(cudaStreamSynchronize(e->stream));
  memset(buf, 0, SWE_STATE_BYTES);
  memcpy(buf, &st, sizeof st);
  return SWE_OK;
}

int swe_import_state(swe_engine* e, const void* buf, size_t bytes) {
  if (int rc = bind(e)) return rc;
  if (!buf || bytes < SWE_STATE_BYTES) return fail(SWE_ERR_INVALID, "state buffer must hold SWE_STATE_BYTES");
  RunState st;
  memcpy(&st, buf, sizeof st);
  if (memcmp(st.magic, "SWESTAT1", 8) != 0) return fail(SWE_ERR_INVALID, "not a swe state blob");
  if (st.rng.f < 0 || st.rng.f >= 31 || st.rng.b < 0 || st.rng.b >= 31) return fail(SWE_ERR_INVALID, "corrupt state blob");
  e->seed = st.seed; e->call_index = st.call_index;
  SWE_CUDA(cudaStreamSynchronize(e->stream));
  SWE_CUDA(cudaMemcpyAsync(e->d_rng, &st.rng, sizeof(DetRand), cudaMemcpyHostToDevice, e->stream));
  SWE_CUDA(cudaStreamSynchronize(e->stream));
  return SWE_OK;
}

// ---- transfers ----------------------------------------------------------------------------
int swe_upload_plane(swe_engine* e, int plane, const float* src, ptrdiff_t stride_bytes) {
  if (int rc = bind(e)) return rc;
  if (!src) return fail(SWE_ERR_INVALID, "src is NULL");
  if (plane < 0 || plane >= SWE_NUM_PLANES) return fail(SWE_ERR_INVALID, "plane %d out of range", plane);
  if (stride_bytes < 4 || stride_bytes % 4) return fail(SWE_ERR_INVALID, "stride_bytes must be a multiple of 4, >= 4");
  const size_t n = e->host_cells(), bytes = n * sizeof(float);
  const MapView ov = e->owned();
  const float* packed = src;
  if (stride_bytes != 4) {  // gather the member out of the caller's AoS records
    if (int rc = need_stage(e, bytes)) return rc;
    SWE_CUDA(cudaStreamSynchronize(e->stream));
    float* s = static_cast<float*>(e->h_stage);
    const char* b = reinterpret_cast<const char*>(src);
    for (size_t i = 0; i < n; i++) std::memcpy(&s[i], b + i * (size_t)stride_bytes, 4);
    packed = s;
  }
  if (plane == SWE_HEIGHT) {
    SWE_CUDA(cudaMemcpyAsync(ov.height, packed, bytes, cudaMemcpyHostToDevice, e->stream));
  } else {
    if (int rc = need_scratch(e, bytes)) return rc;
    SWE_CUDA(cudaMemcpyAsync(e->d_scratch, packed, bytes, cudaMemcpyHostToDevice, e->stream));
    launch_pack_plane(ov, plane, e->d_scratch, e->stream);
    if (int rc = check_launch(e, "pack_plane")) return rc;
  }
  SWE_CUDA(cudaStreamSynchronize(e->stream));  // the caller may reuse src immediately
  return SWE_OK;
}

int swe_download_plane(swe_engine* e, int plane, float* dst, ptrdiff_t stride_bytes) {
  if (int rc = bind(e)) return rc;
  if (!dst) return fail(SWE_ERR_INVALID, "dst is NULL");
  if (plane < 0 || plane >= SWE_NUM_PLANES) return fail(SWE_ERR_INVALID, "plane %d out of range", plane);
  if (stride_bytes < 4 || stride_bytes % 4) return fail(SWE_ERR_INVALID, "stride_bytes must be a multiple of 4, >= 4");
  const size_t n = e->host_cells(), bytes = n * sizeof(float);
  const MapView ov = e->owned();
  const float* d_src = ov.height;
  if (plane != SWE_HEIGHT) {
    if (int rc = need_scratch(e, bytes)) return rc;
    launch_unpack_plane(ov, plane, e->d_scratch, e->stream);
    if (int rc = check_launch(e, "unpack_plane")) return rc;
    d_src = e->d_scratch;
  }
  if (stride_bytes == 4) {
    SWE_CUDA(cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, e->stream));
    SWE_CUDA(cudaStreamSynchronize(e->stream));
  } else {
    if (int rc = need_stage(e, bytes)) return rc;
    SWE_CUDA(cudaMemcpyAsync(e->h_stage, d_src, bytes, cudaMemcpyDeviceToHost, e->stream));
    SWE_CUDA(cudaStreamSynchronize(e->stream));
    const float* s = static_cast<const float*>(e->h_stage);
    char* b = reinterpret_cast<char*>(dst);
    for (size_t i = 0; i < n; i++) std::memcpy(b + i * (size_t)stride_bytes, &s[i], 4);
  }
  finish_timers(e);
  return SWE_OK;
}

// What the reference's frame loop reads after World::erode (SimpleWindErosion.cpp:219-230): height (updatenode)
// and the flow fields (wind map), as five packed planes in one call: one unpack pass over the field plane, two
// device-to-host copies, one synchronisation (five swe_download_plane calls cost five passes and five syncs).
int swe_download_frame(swe_engine* e, float* dst5) {
  if (int rc = bind(e)) return rc;
  if (!dst5) return fail(SWE_ERR_INVALID, "dst is NULL");
  const size_t n = e->host_cells(), bytes = n * sizeof(float);
  const MapView ov = e->owned();
  if (int rc = need_scratch(e, 4 * bytes)) return rc;
  launch_unpack_field(ov, false, e->d_scratch, e->stream);
  if (int rc = check_launch(e, "unpack_field")) return rc;
  SWE_CUDA(cudaMemcpyAsync(dst5, ov.height, bytes, cudaMemcpyDeviceToHost, e->stream));
  SWE_CUDA(cudaMemcpyAsync(dst5 + n, e->d_scratch, 4 * bytes, cudaMemcpyDeviceToHost, e->stream));
  SWE_CUDA(cudaStreamSynchronize(e->stream));
  finish_timers(e);
  return SWE_OK;
}

// The same read-back, pipelined: the height plane is copied on the engine's stream (the caller needs it before it
// can hand the map back), the four flow-field planes are snapshotted on the device (one unpack pass) and travel to
// the host on a side stream WHILE THE NEXT swe_erode RUNS -- the snapshot is private, so nothing the next call does
// can change what arrives. swe_download_frame_wait(e, 0) returns when the height plane is in dst5[0], (e, 1) when
// all five planes are. dst5 must stay valid (pinned memory for a truly asynchronous copy) until wait(e, 1).
int swe_download_frame_begin(swe_engine* e, float* dst5) {
  if (int rc = bind(e)) return rc;
  if (!dst5) return fail(SWE_ERR_INVALID, "dst is NULL");
  const size_t n = e->host_cells(), bytes = n * sizeof(float);
  const MapView ov = e->owned();
  if (!e->frame_stream) {
    SWE_CUDA(cudaStreamCreateWithFlags(&e->frame_stream, cudaStreamNonBlocking));
    SWE_CUDA(cudaEventCreateWithFlags(&e->ev_frame_src, cudaEventDisableTiming));
    SWE_CUDA(cudaEventCreateWithFlags(&e->ev_frame_h, cudaEventDisableTiming));
    SWE_CUDA(cudaEventCreateWithFlags(&e->ev_frame_f, cudaEventDisableTiming));
  }
  if (e->frame_pending) SWE_CUDA(cudaEventSynchronize(e->ev_frame_f));   // the previous frame still owns the snapshot
  if (e->frame_bytes < 4 * bytes) {
    cudaFree(e->d_frame); e->d_frame = nullptr; e->frame_bytes = 0;
    SWE_CUDA(cudaMalloc(&e->d_frame, 4 * bytes));
    e->frame_bytes = 4 * bytes;
  }
  launch_unpack_field(ov, false, e->d_frame, e->stream);
  if (int rc = check_launch(e, "unpack_field")) return rc;
  SWE_CUDA(cudaEventRecord(e->ev_frame_src, e->stream));
  SWE_CUDA(cudaMemcpyAsync(dst5, ov.height, bytes, cudaMemcpyDeviceToHost, e->stream));
  SWE_CUDA(cudaEventRecord(e->ev_frame_h, e->stream));
  // the flow fields queue up BEHIND the height plane: both copies share the device-to-host engine, and started together
  // they would interleave and the plane the caller is waiting for would take as long as all five (measured: 6.0 vs 1.3 ms)
  SWE_CUDA(cudaStreamWaitEvent(e->frame_stream, e->ev_frame_h, 0));
  SWE_CUDA(cudaMemcpyAsync(dst5 + n, e->d_frame, 4 * bytes, cudaMemcpyDeviceToHost, e->frame_stream));
  SWE_CUDA(cudaEventRecord(e->ev_frame_f, e->frame_stream));
  e->frame_pending = true;
  return SWE_OK;
}

int swe_download_frame_wait(swe_engine* e, int all) {
  if (int rc = bind(e)) return rc;
  if (!e->frame_pending) return fail(SWE_ERR_STATE, "no frame read-back in flight (swe_download_frame_begin)");
  SWE_CUDA(cudaEventSynchronize(e->ev_frame_h));
  if (all) {
    SWE_CUDA(cudaEventSynchronize(e->ev_frame_f));
    e->frame_pending = false;
  }
  return SWE_OK;
}

int swe_upload_rock(swe_engine* e, const float* src, ptrdiff_t stride_bytes) {
  if (int rc = bind(e)) return rc;
  if (e->m.periodic) return fail(SWE_ERR_STATE, "the layered extension is not available with the periodic boundary");
  // a stripe engine takes its STORED rows (swe_stripe_stored_rows: the owned rows plus the ghost rows either side):
  // the bedrock under the ghost rows is needed by the sand-limited cascade at the stripe boundary and, unlike the
  // heights, is not refreshed by the halo exchange (abrasion only ever lowers it, so a ghost copy can only
  // underestimate the neighbour's sand: transfers stay within what the owner really has)
  if (!src) return fail(SWE_ERR_INVALID, "src is NULL");
  if (stride_bytes < 4 || stride_bytes % 4) return fail(SWE_ERR_INVALID, "stride_bytes must be a multiple of 4, >= 4");
  const size_t n = e->cells(), bytes = n * sizeof(float);
  if (!e->d_rock) SWE_CUDA(cudaMalloc(&e->d_rock, bytes));
  const float* packed = src;
  if (stride_bytes != 4) {
    if (int rc = need_stage(e, bytes)) return rc;
    SWE_CUDA(cudaStreamSynchronize(e->stream));
    float* s = static_cast<float*>(e->h_stage);
    const char* b = reinterpret_cast<const char*>(src);
    for (size_t i = 0; i < n; i++) std::memcpy(&s[i], b + i * (size_t)stride_bytes, 4);
    packed = s;
  }
  SWE_CUDA(cudaMemcpyAsync(e->d_rock, packed, bytes, cudaMemcpyHostToDevice, e->stream));
  SWE_CUDA(cudaStreamSynchronize(e->stream));
  e->m.rock = e->d_rock;
  return SWE_OK;
}

int swe_download_rock(swe_engine* e, float* dst, ptrdiff_t stride_bytes) {
  if (int rc = bind(e)) return rc;
  if (!e->m.rock) return fail(SWE_ERR_STATE, "no rock plane: the engine is in single-layer mode");
  if (!dst || stride_bytes != 4) return fail(SWE_ERR_INVALID, "dst must be a packed fp32 plane (stride 4)");
  const size_t off = (size_t)(e->own_lo() - e->m.x0) * (size_t)e->m.nz;     // owned rows only, like every download
  SWE_CUDA(cudaMemcpyAsync(dst, e->m.rock + off, e->host_cells() * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  SWE_CUDA(cudaStreamSynchronize(e->stream));
  return SWE_OK;
}

int swe_clear_rock(swe_engine* e) {
  if (int rc = bind(e)) return rc;
  SWE_CUDA(cudaStreamSynchronize(e->stream));
  e->m.rock = nullptr;
  return SWE_OK;
}

int swe_upload_cells(swe_engine* e, const void* cells) {
  if (int rc = bind(e)) return rc;
  if (!cells) return fail(SWE_ERR_INVALID, "cells is NULL");
  const size_t bytes = e->host_cells() * 9 * sizeof(float);
  if (int rc = need_scratch(e, bytes)) return rc;
  SWE_CUDA(cudaMemcpyAsync(e->d_scratch, cells, bytes, cudaMemcpyHostToDevice, e->stream));
  launch_cells_to_planes(e->owned(), e->d_scratch, e->stream);
  if (int rc = check_launch(e, "cells_to_planes")) return rc;
  SWE_CUDA(cudaStreamSynchronize(e->stream));
  return SWE_OK;
}

int swe_download_cells(swe_engine* e, void* cells) {
  if (int rc = bind(e)) return rc;
  if (!cells) return fail(SWE_ERR_INVALID, "cells is NULL");
  const size_t bytes = e->host_cells() * 9 * sizeof(float);
  if (int rc = need_scratch(e, bytes)) return rc;
  launch_planes_to_cells(e->owned(), e->d_scratch, e->stream);
  if (int rc = check_launch(e, "planes_to_cells")) return rc;
  SWE_CUDA(cudaMemcpyAsync(cells, e->d_scratch, bytes, cudaMemcpyDeviceToHost, e->stream));
  SWE_CUDA(cudaStreamSynchronize(e->stream));
  finish_timers(e);
  return SWE_OK;
}

int swe_host_alloc(void** out, size_t bytes) {
  if (!out) return fail(SWE_ERR_INVALID, "out is NULL");
  SWE_CUDA(cudaMallocHost(out, bytes ? bytes : 1));
  return SWE_OK;
}

int swe_host_free(void* p) {
  if (p) SWE_CUDA(cudaFreeHost(p));
  return SWE_OK;
}

// ---- hot path -------------------------------------------------------------------------------
int swe_erode(swe_engine* e, int cycles, int mode) {
  if (int rc = bind(e)) return rc;
  if (cycles < 0) return fail(SWE_ERR_INVALID, "cycles < 0");
  if (mode != SWE_MODE_DETERMINISTIC && mode != SWE_MODE_PARALLEL)
    return fail(SWE_ERR_INVALID, "unknown mode %d", mode);
  if (e->st.enabled)
    if (int rc = stripe_erode_check(e, cycles, mode)) return rc;
  if (e->st.enabled && e->m.rock && !stripe_uses_colour_queues(e))
    // the lock-step stripe round recomputes its ghost rows instead of exchanging them, which needs bit-identical inputs
    // on both sides; abrasion changes the bedrock under a neighbour's ghost copy
    return fail(SWE_ERR_STATE, "the layered extension on stripes needs the colour-queue schedule (swe_stripe_connect_ring)");
  if (e->m.periodic && mode == SWE_MODE_PARALLEL) {
    // the nine-colour order is conflict-free on a torus only if both edges are multiples of 3
    if (e->m.nx % 3 || e->m.nz % 3 || e->m.nx < 6 || e->m.nz < 6)
      return fail(SWE_ERR_STATE, "periodic parallel mode needs map edges that are multiples of 3 (got %dx%d)", e->m.nx, e->m.nz);
    if (e->opt_schedule == 1 || !e->opt_persistent || !e->coop_ok)
      return fail(SWE_ERR_STATE, "periodic parallel mode runs the colour-queue schedule in its persistent form only");
  }
  if (mode == SWE_MODE_DETERMINISTIC && e->m.rock)   // (checked before anything is enqueued: a failed call leaves the engine untouched)
    return fail(SWE_ERR_STATE, "deterministic mode reproduces the reference, which has no rock layer");
  if (!e->st.enabled && mode == SWE_MODE_PARALLEL && cycles > 0 && e->opt_schedule == 2 && !e->m.periodic)
    return erode_tile(e, cycles);
  if (e->st.enabled && cycles > 0 && stripe_uses_tiles(e)) {   // one process per GPU: this rank's share of the stripe group
    swe_engine* one[1] = {e};
    return erode_tile_stripes(one, 1, cycles);
  }
  if (!e->st.enabled && mode == SWE_MODE_PARALLEL && cycles > 0 && e->opt_schedule != 1)
    return erode_colour(e, cycles);   // (also the tile schedule on a torus: tiles do not wrap)
  if (!e->st.enabled && mode == SWE_MODE_PARALLEL && cycles > 0 && e->opt_cuda_graph && !e->opt_kernel_timing)
    return erode_parallel_graph(e, cycles);
  // world.hpp:38-44: reset the four track members of every cell
  if (int rc = erode_prologue(e)) return rc;
  // world.hpp:47-66: the particles
  if (e->st.enabled && stripe_uses_colour_queues(e)) {
    if (cycles > 0) {
      if (int rc = colour_launch_persistent(e, cycles)) return rc;
      // the kernel works on the owners' memory; bring this rank's ghost rows up to date for what follows
      int nl = 0;
      const unsigned q = stripe_exchange_begin(e, false, &nl);
      stripe_exchange_end(e, q, &nl);
      if (int rc = check_launch(e, "stripe halo refresh", nl)) return rc;
    }
  } else if (e->st.enabled) {
    if (cycles > 0) {
      if (int rc = stripe_erode_begin(e, cycles)) return rc;
      for (int r = 0; r < e->p.max_age + 2; r++)
        if (int rc = stripe_erode_round(e)) return rc;
    }
  } else if (mode == SWE_MODE_DETERMINISTIC) {
    launch_det_erode(e->m, e->p, cycles, e->d_rng, e->d_stats, e->stream);
    if (int rc = check_launch(e, "det_erode")) return rc;
  } else if (cycles > 0) {
    if (int rc = need_particles(e, cycles)) return rc;
    if (int rc = need_stamp(e)) return rc;
    if (int rc = parallel_rounds(e, cycles, nullptr)) return rc;
  }
  // world.hpp:70-76: EMA of the flow fields
  return erode_epilogue(e);
}

// Several stripes driven by ONE host thread (logical stripes on one GPU, or one process owning
// several GPUs): the launches are interleaved round by round, so a stripe's wait kernels never
// queue up ahead of the neighbour launches that satisfy them.
int swe_stripe_group_erode(swe_engine** engines, int n, int cycles) {
  if (!engines || n < 1) return fail(SWE_ERR_INVALID, "no engines");
  if (cycles < 0) return fail(SWE_ERR_INVALID, "cycles < 0");
  for (int i = 0; i < n; i++) {
    if (int rc = bind(engines[i])) return rc;
    if (!engines[i]->st.enabled) return fail(SWE_ERR_STATE, "engine %d is not a stripe engine", i);
    if (int rc = stripe_erode_check(engines[i], cycles, SWE_MODE_PARALLEL)) return rc;
    if (engines[i]->p.max_age != engines[0]->p.max_age) return fail(SWE_ERR_INVALID, "stripes disagree on max_age");
  }
  const bool tiles = stripe_uses_tiles(engines[0]);
  for (int i = 0; i < n; i++)
    if (stripe_uses_tiles(engines[i]) != tiles) return fail(SWE_ERR_STATE, "stripes disagree on the schedule");
  if (tiles && cycles > 0) return erode_tile_stripes(engines, n, cycles);
  const bool colour = stripe_uses_colour_queues(engines[0]);
  for (int i = 0; i < n; i++)
    if (stripe_uses_colour_queues(engines[i]) != colour) return fail(SWE_ERR_STATE, "stripes disagree on the schedule");
  for (int i = 0; i < n; i++) {
    if (int rc = bind(engines[i])) return rc;
    if (int rc = erode_prologue(engines[i])) return rc;
    if (cycles > 0 && !colour)
      if (int rc = stripe_erode_begin(engines[i], cycles)) return rc;
  }
  if (cycles > 0 && colour) {
    // one persistent kernel per stripe, all resident together (they meet at the cross-stripe barrier), then
    // one halo refresh
    for (int i = 0; i < n; i++) {
      if (int rc = bind(engines[i])) return rc;
      if (int rc = colour_launch_persistent(engines[i], cycles)) return rc;
    }
    if (int rc = swe_stripe_group_sync_halos(engines, n)) return rc;
  } else if (cycles > 0)
    for (int r = 0; r < engines[0]->p.max_age + 2; r++)
      for (int i = 0; i < n; i++) {
        if (int rc = bind(engines[i])) return rc;
        if (int rc = stripe_erode_round(engines[i])) return rc;
      }
  for (int i = 0; i < n; i++) {
    if (int rc = bind(engines[i])) return rc;
    if (int rc = erode_epilogue(engines[i])) return rc;
  }
  return SWE_OK;
}

int swe_stripe_group_sweep(swe_engine** engines, int n, int sweeps) {
  if (!engines || n < 1) return fail(SWE_ERR_INVALID, "no engines");
  if (sweeps < 0) return fail(SWE_ERR_INVALID, "sweeps < 0");
  for (int i = 0; i < n; i++)
    if (!engines[i] || !engines[i]->st.enabled || !engines[i]->st.connected)
      return fail(SWE_ERR_STATE, "engine %d is not a connected stripe engine", i);
  for (int s = 0; s < sweeps; s++)
    for (int i = 0; i < n; i++) {
      swe_engine* e = engines[i];
      if (int rc = bind(e)) return rc;
      int nl = 0;
      const unsigned q = stripe_exchange_begin(e, false, &nl);
      launch_cascade_stamped(e->m, e->p, &e->tmap, nullptr, 0, nullptr, e->st.own_lo, e->st.own_hi, e->tsc, e->stream);
      nl += 2;
      stripe_exchange_end(e, q, &nl);
      if (int rc = check_launch(e, "stripe sweep", nl)) return rc;
    }
  return SWE_OK;
}

int swe_stripe_group_sync_halos(swe_engine** engines, int n) {
  if (!engines || n < 1) return fail(SWE_ERR_INVALID, "no engines");
  for (int i = 0; i < n; i++)
    if (int rc = swe_stripe_sync_halos(engines[i])) return rc;
  return SWE_OK;
}

int swe_cascade(swe_engine* e, int x, int z) {
  if (int rc = bind(e)) return rc;
  if (e->st.enabled) return fail(SWE_ERR_STATE, "not available on a stripe engine (deterministic mode is single-GPU)");
  if (e->m.oob(x, z)) return fail(SWE_ERR_OOB, "cell (%d,%d) outside %dx%d", x, z, e->m.nx, e->m.nz);
  launch_det_cascade_one(e->m, e->p, x, z, e->stream);
  return check_launch(e, "cascade_one");
}

int swe_max_excess(swe_engine* e, float* out) {
  if (int rc = bind(e)) return rc;
  if (!out) return fail(SWE_ERR_INVALID, "out is NULL");
  // a stripe also checks the pairs between its last owned row and the ghost row above it: call
  // swe_stripe_sync_halos first so that ghost row is current
  launch_max_excess(e->owned(e->st.enabled && e->st.own_hi < e->m.nx ? 1 : 0), e->p, e->d_small, e->stream);
  if (int rc = check_launch(e, "max_excess", 2)) return rc;
  SWE_CUDA(cudaMemcpyAsync(out, e->d_small, sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  SWE_CUDA(cudaStreamSynchronize(e->stream));
  finish_timers(e);
  return SWE_OK;
}

int swe_cascade_sweep(swe_engine* e, int sweeps, float* max_excess_out) {
  if (int rc = bind(e)) return rc;
  if (sweeps < 0) return fail(SWE_ERR_INVALID, "sweeps < 0");
  if (e->m.periodic && (e->m.nx % 3 || e->m.nz % 3 || e->m.nx < 6 || e->m.nz < 6))
    return fail(SWE_ERR_STATE, "a periodic cascade sweep needs map edges that are multiples of 3 (got %dx%d)", e->m.nx, e->m.nz);
  SWE_CUDA(cudaEventRecord(e->ev_c, e->stream));
  int n = 0;
  if (e->st.enabled) {
    // one exchange of pre-sweep boundary rows per sweep; owned rows come out bitwise equal to the
    // single-GPU sweep (4/6-row dependency cone inside the 5/7-row ghost zone)
    if (!e->st.connected) return fail(SWE_ERR_STATE, "swe_stripe_connect has not been called");
    for (int s = 0; s < sweeps; s++) {
      const unsigned q = stripe_exchange_begin(e, false, &n);
      launch_cascade_stamped(e->m, e->p, &e->tmap, nullptr, 0, nullptr, e->st.own_lo, e->st.own_hi, e->tsc, e->stream);
      n += 2;
      stripe_exchange_end(e, q, &n);
    }
  } else {
    n = launch_cascade_sweep(e->m, e->p, &e->tmap, sweeps, e->d_stats, e->stream, e->opt_cascade_variant, e->tsc);
  }
  if (int rc = check_launch(e, "cascade_sweep", n)) return rc;
  SWE_CUDA(cudaEventRecord(e->ev_d, e->stream));
  e->sweep_timed = true;
  if (max_excess_out) return swe_max_excess(e, max_excess_out);
  return SWE_OK;
}

int swe_cascade_converge(swe_engine* e, float tol, int max_sweeps, int* sweeps_done,
                         float* max_excess_out) {
  if (int rc = bind(e)) return rc;
  if (max_sweeps < 0) return fail(SWE_ERR_INVALID, "max_sweeps < 0");
  int done = 0;
  float ex = 0.0f;
  if (int rc = swe_max_excess(e, &ex)) return rc;
  // check the metric every 4 sweeps: one extra 4 B/cell read pass per check
  while (ex >= tol && done < max_sweeps) {
    const int k = (max_sweeps - done) < 4 ? (max_sweeps - done) : 4;
    if (int rc = swe_cascade_sweep(e, k, &ex)) return rc;
    done += k;
  }
  if (sweeps_done) *sweeps_done = done;
  if (max_excess_out) *max_excess_out = ex;
  return SWE_OK;
}

int swe_trace(swe_engine* e, float x, float z, int max_records, float* records, int* steps_out,
              swe_wind* final_out) {
  if (int rc = bind(e)) return rc;
  if (e->st.enabled) return fail(SWE_ERR_STATE, "not available on a stripe engine (deterministic mode is single-GPU)");
  if (max_records < 0) return fail(SWE_ERR_INVALID, "max_records < 0");
  if (!records) max_records = 0;
  const size_t rec_bytes = (size_t)max_records * 8 * sizeof(float);
  const size_t tail = 64;  // int steps + swe_wind
  if (int rc = need_scratch(e, rec_bytes + tail)) return rc;
  float* d_rec = e->d_scratch;
  char* d_tail = reinterpret_cast<char*>(e->d_scratch) + rec_bytes;
  int* d_steps = reinterpret_cast<int*>(d_tail);
  swe_wind* d_final = reinterpret_cast<swe_wind*>(d_tail + 16);
  launch_det_trace(e->m, e->p, x, z, max_records, max_records ? d_rec : nullptr, d_steps, d_final,
                   e->d_rng, e->d_stats, e->stream);
  if (int rc = check_launch(e, "det_trace")) return rc;
  int steps = 0;
  swe_wind fin;
  SWE_CUDA(cudaMemcpyAsync(&steps, d_steps, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  SWE_CUDA(cudaMemcpyAsync(&fin, d_final, sizeof(swe_wind), cudaMemcpyDeviceToHost, e->stream));
  SWE_CUDA(cudaStreamSynchronize(e->stream));
  if (max_records) {
    const int nrec = steps < max_records ? steps : max_records;
    SWE_CUDA(cudaMemcpy(records, d_rec, (size_t)nrec * 8 * sizeof(float), cudaMemcpyDeviceToHost));
  }
  if (steps_out) *steps_out = steps;
  if (final_out) *final_out = fin;
  return SWE_OK;
}

int swe_wind_init(swe_wind* w, float x, float z) {
  if (!w) return fail(SWE_ERR_INVALID, "wind is NULL");
  w->age = 0; w->sediment = 0.0f;
  w->pos[0] = x; w->pos[1] = 0.0f; w->pos[2] = z;      // Wind(vec2): pos = (x, 0, y)  wind.hpp:6
  w->speed[0] = w->speed[1] = w->speed[2] = 0.0f;
  return SWE_OK;
}

int swe_wind_fly(swe_engine* e, swe_wind* w, int* alive_out) {
  if (int rc = bind(e)) return rc;
  if (e->st.enabled) return fail(SWE_ERR_STATE, "not available on a stripe engine (deterministic mode is single-GPU)");
  if (!w) return fail(SWE_ERR_INVALID, "wind is NULL");
  if (int rc = need_scratch(e, 128)) return rc;
  swe_wind* d_w = reinterpret_cast<swe_wind*>(e->d_scratch);
  int* d_alive = reinterpret_cast<int*>(reinterpret_cast<char*>(e->d_scratch) + 64);
  SWE_CUDA(cudaMemcpyAsync(d_w, w, sizeof(swe_wind), cudaMemcpyHostToDevice, e->stream));
  launch_det_fly_once(e->m, e->p, d_w, d_alive, e->d_rng, e->d_stats, e->stream);
  if (int rc = check_launch(e, "det_fly_once")) return rc;
  int alive = 0;
  SWE_CUDA(cudaMemcpyAsync(w, d_w, sizeof(swe_wind), cudaMemcpyDeviceToHost, e->stream));
  SWE_CUDA(cudaMemcpyAsync(&alive, d_alive, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  SWE_CUDA(cudaStreamSynchronize(e->stream));
  if (alive_out) *alive_out = alive;
  return SWE_OK;
}

int swe_normal(swe_engine* e, int x, int z, float out3[3]) {
  if (int rc = bind(e)) return rc;
  if (!out3) return fail(SWE_ERR_INVALID, "out is NULL");
  if (e->m.oob(x, z) || x < e->own_lo() || x >= e->own_hi())
    return fail(SWE_ERR_OOB, "cell (%d,%d) outside rows [%d,%d) of %dx%d", x, z, e->own_lo(), e->own_hi(), e->m.nx, e->m.nz);
  launch_normal_one(e->m, e->p, x, z, e->d_small, e->stream);
  if (int rc = check_launch(e, "normal_one")) return rc;
  SWE_CUDA(cudaMemcpyAsync(out3, e->d_small, 3 * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  SWE_CUDA(cudaStreamSynchronize(e->stream));
  return SWE_OK;
}

int swe_height(swe_engine* e, int x, int z, float* out) {
  if (int rc = bind(e)) return rc;
  if (!out) return fail(SWE_ERR_INVALID, "out is NULL");
  if (e->m.oob(x, z) || x < e->own_lo() || x >= e->own_hi())
    return fail(SWE_ERR_OOB, "cell (%d,%d) outside rows [%d,%d) of %dx%d", x, z, e->own_lo(), e->own_hi(), e->m.nx, e->m.nz);
  SWE_CUDA(cudaMemcpyAsync(out, e->m.height + e->m.idx(x, z), sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  SWE_CUDA(cudaStreamSynchronize(e->stream));
  return SWE_OK;
}

int swe_download_normals(swe_engine* e, float* dst) {
  if (int rc = bind(e)) return rc;
  if (!dst) return fail(SWE_ERR_INVALID, "dst is NULL");
  const size_t bytes = e->host_cells() * 3 * sizeof(float);
  if (int rc = need_scratch(e, bytes)) return rc;
  {
    MapView nv = e->m;            // normals of the owned rows; neighbours may be ghost rows
    launch_normals_rows(nv, e->p, e->own_lo(), e->own_hi(), e->d_scratch, e->stream);
  }
  if (int rc = check_launch(e, "normals")) return rc;
  SWE_CUDA(cudaMemcpyAsync(dst, e->d_scratch, bytes, cudaMemcpyDeviceToHost, e->stream));
  SWE_CUDA(cudaStreamSynchronize(e->stream));
  return SWE_OK;
}

// updatenode (cellpool.h:302-321): one 48-byte Vertex per cell, to a host buffer or straight into
// device memory (a CUDA-mapped OpenGL vertex buffer, vertexpool.hpp:155-172)
int swe_download_vertices(swe_engine* e, void* dst) {
  if (int rc = bind(e)) return rc;
  if (!dst) return fail(SWE_ERR_INVALID, "dst is NULL");
  const size_t bytes = e->host_cells() * 48;
  if (int rc = need_scratch(e, bytes)) return rc;
  launch_vertices_rows(e->m, e->p, e->own_lo(), e->own_hi(), e->d_scratch, e->stream);
  if (int rc = check_launch(e, "vertices")) return rc;
  SWE_CUDA(cudaMemcpyAsync(dst, e->d_scratch, bytes, cudaMemcpyDeviceToHost, e->stream));
  SWE_CUDA(cudaStreamSynchronize(e->stream));
  return SWE_OK;
}

int swe_write_vertices_device(swe_engine* e, void* device_dst) {
  if (int rc = bind(e)) return rc;
  if (!device_dst) return fail(SWE_ERR_INVALID, "device_dst is NULL");
  if (reinterpret_cast<uintptr_t>(device_dst) & 15u) return fail(SWE_ERR_INVALID, "device_dst must be 16-byte aligned");
  launch_vertices_rows(e->m, e->p, e->own_lo(), e->own_hi(), device_dst, e->stream);
  return check_launch(e, "vertices");
}

// ---- generate --------------------------------------------------------------
int swe_generate_noise(swe_engine* e, int seed, int layers, int tile) {
  if (int rc = bind(e)) return rc;
  if (layers < 0 || layers > noise_max_layers()) return fail(SWE_ERR_INVALID, "layers must be in [0, %d]", noise_max_layers());
  if (tile < 0) return fail(SWE_ERR_INVALID, "tile must be >= 0");
  if (layers == 0) return SWE_OK;
  launch_noise_stack(e->m, seed, layers, tile ? tile : e->m.nz, e->stream);   // ghost rows included
  return check_launch(e, "noise_stack");
}

int swe_height_range(swe_engine* e, float* min_out, float* max_out) {
  if (int rc = bind(e)) return rc;
  if (!min_out || !max_out) return fail(SWE_ERR_INVALID, "output pointer is NULL");
  int* d_out = reinterpret_cast<int*>(e->d_small + 24);
  launch_height_range(e->owned().height, e->host_cells(), d_out, e->stream);
  if (int rc = check_launch(e, "height_range")) return rc;
  float both[2];
  SWE_CUDA(cudaMemcpyAsync(both, d_out, sizeof both, cudaMemcpyDeviceToHost, e->stream));
  SWE_CUDA(cudaStreamSynchronize(e->stream));
  *min_out = both[0]; *max_out = both[1];
  return SWE_OK;
}

int swe_normalize_height(swe_engine* e, float min, float max) {
  if (int rc = bind(e)) return rc;
  if (!(max > min)) return fail(SWE_ERR_INVALID, "normalisation needs max > min (a flat map divides by zero, cellpool.h:427)");
  launch_normalize(e->m.height, e->cells(), min, max, e->stream);              // ghost rows included
  return check_launch(e, "normalize");
}

// ---- diagnostics ----------------------------------------------------------------------------
int swe_sync(swe_engine* e) {
  if (int rc = bind(e)) return rc;
  SWE_CUDA(cudaStreamSynchronize(e->stream));
  finish_timers(e);
  return SWE_OK;
}

int swe_get_stats(swe_engine* e, swe_stats* out) {
  if (int rc = bind(e)) return rc;
  if (!out) return fail(SWE_ERR_INVALID, "out is NULL");
  DevStats s;
  SWE_CUDA(cudaMemcpyAsync(&s, e->d_stats, sizeof s, cudaMemcpyDeviceToHost, e->stream));
  SWE_CUDA(cudaStreamSynchronize(e->stream));
  finish_timers(e);
  std::memset(out, 0, sizeof *out);
  out->particles = s.particles; out->steps = s.steps;
  out->exits_oob = s.exits_oob; out->exits_age = s.exits_age;
  out->sediment_lost = s.sediment_lost; out->cascade_cells = s.cascade_cells;
  out->erode_calls = e->erode_calls;
  out->last_erode_ms = e->last_erode_ms; out->last_sweep_ms = e->last_sweep_ms;
  out->kernel_launches = e->launches;
  if (e->st.enabled) {
    unsigned vars[4] = {0, 0, 0, 0};
    SWE_CUDA(cudaMemcpy(vars, e->st.d_vars, sizeof vars, cudaMemcpyDeviceToHost));
    out->stripe_overflow = vars[1];
    if (vars[2]) return fail(SWE_ERR_STATE, "stripe neighbour exchange timed out (a rank is missing or crashed)");
  }
  if (e->cq.persist) {
    unsigned pc[4] = {0, 0, 0, 0};   // PersistCtl line 2: live, tag_base, micro, error
    SWE_CUDA(cudaMemcpy(pc, static_cast<char*>(e->cq.persist) + 256, sizeof pc, cudaMemcpyDeviceToHost));
    if (e->last_micro_rounds < 0) e->last_micro_rounds = (int)pc[2];
    if (pc[3] == 1u) return fail(SWE_ERR_STATE, "colour-queue kernel: grid barrier watchdog fired");
    if (pc[3] == 2u) return fail(SWE_ERR_STATE, "colour-queue kernel: micro-round cap reached with particles alive");
  }
  out->last_micro_rounds = (uint64_t)(e->last_micro_rounds < 0 ? 0 : e->last_micro_rounds);
  out->last_super_rounds = (uint64_t)e->last_super_rounds;
  if (e->cq.count) {
    unsigned dropped = 0;
    SWE_CUDA(cudaMemcpy(&dropped, e->cq.count + 19 * 32, sizeof dropped, cudaMemcpyDeviceToHost));
    out->queue_dropped = dropped;
  }
  out->last_fly_ms = e->last_fly_ms;
  out->last_round_ms = e->last_round_ms;
  out->last_cascade_ms = e->last_cascade_ms;
  return SWE_OK;
}

// Restore the accumulated counters (checkpoint / resume): what swe_get_stats reported when the checkpoint was
// written becomes the baseline again, so that the mass balance sum(height) + sediment_lost stays continuous.
int swe_import_stats(swe_engine* e, const swe_stats* in) {
  if (int rc = bind(e)) return rc;
  if (!in) return fail(SWE_ERR_INVALID, "stats is NULL");
  DevStats s{};
  s.particles = in->particles; s.steps = in->steps; s.exits_oob = in->exits_oob; s.exits_age = in->exits_age;
  s.sediment_lost = in->sediment_lost; s.cascade_cells = in->cascade_cells;
  SWE_CUDA(cudaStreamSynchronize(e->stream));
  SWE_CUDA(cudaMemcpy(e->d_stats, &s, sizeof s, cudaMemcpyHostToDevice));
  e->erode_calls = in->erode_calls;
  return SWE_OK;
}

int swe_reset_stats(swe_engine* e) {
  if (int rc = bind(e)) return rc;
  SWE_CUDA(cudaMemsetAsync(e->d_stats, 0, sizeof(DevStats), e->stream));
  e->erode_calls = 0; e->launches = 0;
  return SWE_OK;
}

int swe_sum_plane(swe_engine* e, int plane, double* out) {
  if (int rc = bind(e)) return rc;
  if (!out) return fail(SWE_ERR_INVALID, "out is NULL");
  if (plane < 0 || plane >= SWE_NUM_PLANES) return fail(SWE_ERR_INVALID, "plane %d out of range", plane);
  double* d_out = reinterpret_cast<double*>(e->d_small + 16);
  launch_sum_plane(e->owned(), plane, d_out, e->stream);
  if (int rc = check_launch(e, "sum_plane")) return rc;
  SWE_CUDA(cudaMemcpyAsync(out, d_out, sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  SWE_CUDA(cudaStreamSynchronize(e->stream));
  return SWE_OK;
}

int swe_hash_height(swe_engine* e, uint64_t* out) {
  if (int rc = bind(e)) return rc;
  if (!out) return fail(SWE_ERR_INVALID, "out is NULL");
  std::vector<uint32_t> host(e->host_cells());
  SWE_CUDA(cudaMemcpyAsync(host.data(), e->owned().height, host.size() * 4, cudaMemcpyDeviceToHost, e->stream));
  SWE_CUDA(cudaStreamSynchronize(e->stream));
  uint64_t h = 1469598103934665603ull;
  for (uint32_t w : host) { h ^= w; h *= 1099511628211ull; }
  *out = h;
  return SWE_OK;
}

// ---- multi-GPU stripes -----------------------------------------------------------------------
int swe_stripe_export(swe_engine* e, swe_stripe_link* out) {
  if (!e || !out) return fail(SWE_ERR_INVALID, "NULL argument");
  if (!e->st.enabled) return fail(SWE_ERR_STATE, "not a stripe engine");
  std::memset(out, 0, sizeof *out);
  out->height = e->m.height; out->stamp = e->d_stamp; out->arena = e->st.arena;
  out->x0 = e->m.x0; out->inbox_capacity = e->st.inbox_capacity;
  out->queue_a = e->cq.a; out->queue_b = e->cq.b; out->queue_count = e->cq.count; out->persist = e->cq.persist;
  out->queue_cap = e->cq.cap;
  out->reserved = e->m.rows;          // stored rows: the track plane follows the height plane in the same allocation
  return SWE_OK;
}

int swe_stripe_export_ipc(swe_engine* e, void* buf, size_t buf_bytes) {
  if (int rc = bind(e)) return rc;
  if (!e->st.enabled) return fail(SWE_ERR_STATE, "not a stripe engine");
  if (!buf || buf_bytes < SWE_STRIPE_IPC_BYTES) return fail(SWE_ERR_INVALID, "buffer must hold SWE_STRIPE_IPC_BYTES");
  unsigned char* b = static_cast<unsigned char*>(buf);
  std::memset(b, 0, SWE_STRIPE_IPC_BYTES);
  cudaIpcMemHandle_t h[7];
  SWE_CUDA(cudaIpcGetMemHandle(&h[0], e->m.height));
  SWE_CUDA(cudaIpcGetMemHandle(&h[1], e->d_stamp));
  SWE_CUDA(cudaIpcGetMemHandle(&h[2], e->st.arena));
  SWE_CUDA(cudaIpcGetMemHandle(&h[3], e->cq.a));
  SWE_CUDA(cudaIpcGetMemHandle(&h[4], e->cq.b));
  SWE_CUDA(cudaIpcGetMemHandle(&h[5], e->cq.count));
  SWE_CUDA(cudaIpcGetMemHandle(&h[6], e->cq.persist));
  static_assert(7 * sizeof(cudaIpcMemHandle_t) + 4 * sizeof(int) <= SWE_STRIPE_IPC_BYTES, "ipc blob size");
  std::memcpy(b, h, sizeof h);
  const int meta[4] = {e->m.x0, e->st.inbox_capacity, e->cq.cap, e->m.rows};
  std::memcpy(b + sizeof h, meta, sizeof meta);
  return SWE_OK;
}

int swe_stripe_open_ipc(swe_engine* e, const void* buf, size_t buf_bytes, swe_stripe_link* out) {
  if (int rc = bind(e)) return rc;
  if (!buf || !out || buf_bytes < SWE_STRIPE_IPC_BYTES) return fail(SWE_ERR_INVALID, "bad ipc blob");
  const unsigned char* b = static_cast<const unsigned char*>(buf);
  cudaIpcMemHandle_t h[7];
  std::memcpy(h, b, sizeof h);
  int meta[4];
  std::memcpy(meta, b + sizeof h, sizeof meta);
  void* ptr[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  for (int i = 0; i < 7; i++) {
    SWE_CUDA(cudaIpcOpenMemHandle(&ptr[i], h[i], cudaIpcMemLazyEnablePeerAccess));
    e->st.ipc_opened.push_back(ptr[i]);
  }
  std::memset(out, 0, sizeof *out);
  out->height = ptr[0]; out->stamp = ptr[1]; out->arena = ptr[2];
  out->queue_a = ptr[3]; out->queue_b = ptr[4]; out->queue_count = ptr[5]; out->persist = ptr[6];
  out->x0 = meta[0]; out->inbox_capacity = meta[1]; out->queue_cap = meta[2]; out->reserved = meta[3];
  return SWE_OK;
}

int swe_stripe_connect(swe_engine* e, const swe_stripe_link* below, const swe_stripe_link* above) {
  if (int rc = bind(e)) return rc;
  if (!e->st.enabled) return fail(SWE_ERR_STATE, "not a stripe engine");
  auto& S = e->st;
  if ((below == nullptr) != (S.own_lo == 0))
    return fail(SWE_ERR_INVALID, "a stripe needs a neighbour below exactly when it does not start at row 0");
  if ((above == nullptr) != (S.own_hi == e->m.nx))
    return fail(SWE_ERR_INVALID, "a stripe needs a neighbour above exactly when it does not end at the last row");
  for (const swe_stripe_link* l : {below, above})
    if (l && (!l->height || !l->stamp || !l->arena || l->inbox_capacity != S.inbox_capacity))
      return fail(SWE_ERR_INVALID, "neighbour link is incomplete or was created with another max_particles");
  S.has_below = below != nullptr; S.has_above = above != nullptr;
  if (below) S.below = *below;
  if (above) S.above = *above;
  S.connected = true;
  return SWE_OK;
}

int swe_stripe_connect_ring(swe_engine* e, int rank, int n, const swe_stripe_link* links, int colocated) {
  if (int rc = bind(e)) return rc;
  if (!e->st.enabled) return fail(SWE_ERR_STATE, "not a stripe engine");
  if (!links || n < 1 || n > SWE_MAX_STRIPE_RANKS || rank < 0 || rank >= n)
    return fail(SWE_ERR_INVALID, "need the links of all n <= %d ranks and 0 <= rank < n", SWE_MAX_STRIPE_RANKS);
  if (colocated < 1) return fail(SWE_ERR_INVALID, "colocated must be >= 1");
  if (int rc = swe_stripe_connect(e, rank > 0 ? &links[rank - 1] : nullptr, rank + 1 < n ? &links[rank + 1] : nullptr)) return rc;
  auto& S = e->st;
  for (int i = 0; i < n; i++) {
    if (i == rank) { S.mailboxes[i] = e->cq.persist; continue; }
    if (!links[i].persist) return fail(SWE_ERR_INVALID, "link %d carries no barrier block", i);
    S.mailboxes[i] = links[i].persist;
  }
  for (const swe_stripe_link* l : {rank > 0 ? &links[rank - 1] : nullptr, rank + 1 < n ? &links[rank + 1] : nullptr})
    if (l && (!l->queue_a || !l->queue_b || !l->queue_count || l->queue_cap < 1))
      return fail(SWE_ERR_INVALID, "neighbour link carries no particle queues");
  S.rank = rank; S.nranks = n; S.colocated = colocated;
  S.persist_blocks_per_sm = colour_persistent_blocks_per_sm_stripe();
  (void)cudaGetLastError();
  S.ring = true;
  // tile-resident scheduler on stripes: records and per-tile tables live inside the queue buffers (they are what
  // the neighbours have mapped); available when every rank's buffers are big enough and its stripe is at least two
  // tiles tall -- decided from numbers every rank knows, so that all ranks decide alike
  {
    const int words = tile_max_tiles(e->m.nx, e->m.nz) + 1;
    auto layout = [&](void* qa, void* qb, int qcap, TileSched* out) -> bool {
      if (!qa || !qb || qcap < 1024) return false;
      if ((4ull * words + 16ull) * sizeof(unsigned) > 16ull * (unsigned long long)qcap * sizeof(float4)) return false;
      float4* a = static_cast<float4*>(qa);
      float4* b = static_cast<float4*>(qb);
      out->pa[0] = a; out->pa[1] = a + qcap; out->pb[0] = b; out->pb[1] = b + qcap;
      unsigned* tab = reinterpret_cast<unsigned*>(a + 2 * (size_t)qcap);
      out->count = tab; out->base[0] = tab + words; out->base[1] = tab + 2 * (size_t)words;
      out->cursor = tab + 3 * (size_t)words; out->info = tab + 4 * (size_t)words;
      out->cap = qcap; out->max_tiles = words - 1;
      return true;
    };
    bool ok = (e->m.nx / n - 8) >= 2 * tile_edge_x(0) && (S.own_hi - S.own_lo) >= 2 * tile_edge_x(0) && S.merged_planes;   // (the smallest tile)
    ok = ok && layout(e->cq.a, e->cq.b, e->cq.cap, &S.ts);
    TileStripeHost H{};
    H.enabled = 1; H.own_lo = S.own_lo; H.own_hi = S.own_hi;
    H.has_below = rank > 0; H.has_above = rank + 1 < n;
    if (ok && H.has_below) {
      const swe_stripe_link& l = links[rank - 1];
      ok = l.reserved > 0 && layout(l.queue_a, l.queue_b, l.queue_cap, &H.below);
      H.below_h = static_cast<float*>(l.height); H.below_x0 = l.x0;
      H.below_lo = l.x0 == 0 ? 0 : l.x0 + GHOST_LO;
      const size_t hbytes = ((size_t)l.reserved * (size_t)e->m.nz * sizeof(float) + 255) / 256 * 256;
      H.below_t = reinterpret_cast<float4*>(static_cast<char*>(l.height) + hbytes);
    }
    if (ok && H.has_above) {
      const swe_stripe_link& l = links[rank + 1];
      ok = l.reserved > 0 && layout(l.queue_a, l.queue_b, l.queue_cap, &H.above);
      H.above_h = static_cast<float*>(l.height); H.above_x0 = l.x0;
      const size_t hbytes = ((size_t)l.reserved * (size_t)e->m.nz * sizeof(float) + 255) / 256 * 256;
      H.above_t = reinterpret_cast<float4*>(static_cast<char*>(l.height) + hbytes);
    }
    S.tiles = ok;
    S.th = H;
    S.tile_seq = 0;
    if (ok) {
      SWE_CUDA(cudaMemsetAsync(S.ts.count, 0, (4 * (size_t)words + 16) * sizeof(unsigned), e->stream));
      SWE_CUDA(cudaMemsetAsync(static_cast<char*>(e->cq.persist) + SWE_TILE_MAILBOX_OFFSET, 0, 1024, e->stream));
      SWE_CUDA(cudaStreamSynchronize(e->stream));
    }
  }
  return SWE_OK;
}

int swe_stripe_sync_halos(swe_engine* e) {
  if (int rc = bind(e)) return rc;
  if (!e->st.enabled || !e->st.connected) return fail(SWE_ERR_STATE, "stripe engine is not connected");
  int n = 0;
  const unsigned s = stripe_exchange_begin(e, false, &n);
  stripe_exchange_end(e, s, &n);
  return check_launch(e, "stripe halo sync", n);
}

int swe_stripe_stored_rows(const swe_engine* e, int* x0, int* rows) {
  if (!e) return fail(SWE_ERR_INVALID, "engine is NULL");
  if (x0) *x0 = e->m.x0;
  if (rows) *rows = e->m.rows;
  return SWE_OK;
}

int swe_stripe_rows(const swe_engine* e, int* own_lo, int* own_hi) {
  if (!e) return fail(SWE_ERR_INVALID, "engine is NULL");
  if (own_lo) *own_lo = e->own_lo();
  if (own_hi) *own_hi = e->own_hi();
  return SWE_OK;
}

int swe_get_option(const swe_engine* e, int option, int* value) {
  if (!e || !value) return fail(SWE_ERR_INVALID, "NULL argument");
  switch (option) {
    case SWE_OPT_CASCADE_VARIANT: *value = e->opt_cascade_variant; return SWE_OK;
    case SWE_OPT_CASCADE_INTERVAL: *value = e->opt_cascade_interval; return SWE_OK;
    case SWE_OPT_KERNEL_TIMING: *value = e->opt_kernel_timing; return SWE_OK;
    case SWE_OPT_CUDA_GRAPH: *value = e->opt_cuda_graph; return SWE_OK;
    case SWE_OPT_PDL: *value = e->opt_pdl; return SWE_OK;
    case SWE_OPT_CALL_BATCHES: *value = e->opt_call_batches; return SWE_OK;
    case SWE_OPT_SCHEDULE: *value = e->opt_schedule; return SWE_OK;
    case SWE_OPT_STAGGER_ROUNDS: *value = e->opt_stagger; return SWE_OK;
    case SWE_OPT_PERSISTENT: *value = e->opt_persistent; return SWE_OK;
    case SWE_OPT_PERIODIC: *value = e->m.periodic; return SWE_OK;
    case SWE_OPT_TILE_ROUNDS: *value = e->opt_tile_rounds; return SWE_OK;
    case SWE_OPT_TILE_SHAPE: *value = e->opt_tile_shape; return SWE_OK;
    default: return fail(SWE_ERR_INVALID, "unknown option %d", option);
  }
}

int swe_set_option(swe_engine* e, int option, int value) {
  if (!e) return fail(SWE_ERR_INVALID, "engine is NULL");
  switch (option) {
    case SWE_OPT_CASCADE_VARIANT:
      if (value < 0 || value > 2) return fail(SWE_ERR_INVALID, "cascade variant must be 0, 1 or 2");
      e->opt_cascade_variant = value; return SWE_OK;
    case SWE_OPT_CASCADE_INTERVAL:
      if (value < 1) return fail(SWE_ERR_INVALID, "cascade interval must be >= 1");
      e->opt_cascade_interval = value; return SWE_OK;
    case SWE_OPT_KERNEL_TIMING:
      e->opt_kernel_timing = value ? 1 : 0; return SWE_OK;
    case SWE_OPT_CUDA_GRAPH:
      e->opt_cuda_graph = value ? 1 : 0; return SWE_OK;
    case SWE_OPT_PDL:
      e->opt_pdl = value ? 1 : 0; return SWE_OK;
    case SWE_OPT_CALL_BATCHES:
      if (value < 1 || value > 64) return fail(SWE_ERR_INVALID, "call batches must be in [1, 64]");
      e->opt_call_batches = value; return SWE_OK;
    case SWE_OPT_SCHEDULE:
      if (value < 0 || value > 2) return fail(SWE_ERR_INVALID, "schedule must be 0 (colour queues), 1 (lock-step rounds) or 2 (tile-resident)");
      e->opt_schedule = value; return SWE_OK;
    case SWE_OPT_PERSISTENT:
      e->opt_persistent = value ? 1 : 0; return SWE_OK;
    case SWE_OPT_PERIODIC:
      if (value && e->st.enabled) return fail(SWE_ERR_STATE, "the periodic boundary is not available on stripe engines");
      if (value && e->m.rock) return fail(SWE_ERR_STATE, "the periodic boundary is not available with a rock plane");
      if (cudaSetDevice(e->device) == cudaSuccess) cudaStreamSynchronize(e->stream);
      e->m.periodic = value ? 1 : 0;
      for (auto& g : e->graphs) cudaGraphExecDestroy(g.exec);   // captured launches carry the MapView by value
      e->graphs.clear();
      return SWE_OK;
    case SWE_OPT_TILE_ROUNDS:
      if (value < 0 || value > 64) return fail(SWE_ERR_INVALID, "tile rounds must be in [0, 64] (0 = auto)");
      e->opt_tile_rounds = value; return SWE_OK;
    case SWE_OPT_TILE_SHAPE:
      if (value < -1 || value >= tile_shapes()) return fail(SWE_ERR_INVALID, "tile shape must be -1 (auto) or in [0, %d)", tile_shapes());
      e->opt_tile_shape = value; return SWE_OK;
    case SWE_OPT_STAGGER_ROUNDS:
      if (value < -1 || value > (1 << 20)) return fail(SWE_ERR_INVALID, "stagger rounds must be -1 (auto) or in [0, 2^20]");
      e->opt_stagger = value; return SWE_OK;
    default:
      return fail(SWE_ERR_INVALID, "unknown option %d", option);
  }
}

// ---- device-side plumbing -------------------------------------------------------------------
int swe_set_stream(swe_engine* e, void* cuda_stream) {
  if (int rc = bind(e)) return rc;
  SWE_CUDA(cudaStreamSynchronize(e->stream));
  finish_timers(e);
  e->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : e->own_stream;
  return SWE_OK;
}
void* swe_get_stream(swe_engine* e) { return e ? static_cast<void*>(e->stream) : nullptr; }
void* swe_device_height(swe_engine* e) { return e ? e->m.height : nullptr; }
void* swe_device_field(swe_engine* e) { return e ? e->m.field : nullptr; }
void* swe_device_track(swe_engine* e) { return e ? e->m.track : nullptr; }

}  // extern "C"
