// swe_api.cu -- the C-ABI (include/swe.h) over the CUDA kernels: engine object, host<->device
// transfer with arbitrary host stride, stream plumbing, statistics. Host-only logic; all
// arithmetic on map data happens in kernels. No CPU fallback anywhere: every entry point
// that touches the map needs the CUDA device the engine was created on.
#include <cstdarg>
#include <cstdio>
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
  if (!(p.boundarylayer > 0.0f)) { *why = "boundarylayer must be > 0"; return false; }
  if (p.mapscale <= 0) { *why = "mapscale must be > 0"; return false; }
  const float l2 = p.pspeed[0] * p.pspeed[0] + p.pspeed[1] * p.pspeed[1] + p.pspeed[2] * p.pspeed[2];
  if (!(l2 > 0.0f)) { *why = "pspeed must be non-zero"; return false; }
  for (int i = 0; i < 5; i++)
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
  float* d_scratch = nullptr;   // lazily sized staging for plane/AoS conversion
  size_t scratch_bytes = 0;
  void* h_stage = nullptr;      // pinned host staging for strided host buffers
  size_t h_stage_bytes = 0;
  uint64_t seed = 1;
  uint32_t call_index = 0;
  // parallel-mode state
  ParticleBuf pbuf{};
  HeightTensorMap tmap{};
  unsigned char* d_stamp = nullptr;   // per-cell "touched in round" byte
  unsigned stamp_counter = 0;
  int opt_cascade_variant = 0;        // 0 fused tile kernel, 1 per-colour launches
  int opt_cascade_interval = 1;       // Wind::fly steps per particle between cascade passes
  int opt_kernel_timing = 0;          // record an event pair around every round / cascade launch
  std::vector<cudaEvent_t> kt_events; // [3*r] before round, [3*r+1] between, [3*r+2] after cascade
  int kt_rounds = 0;
  double last_round_ms = 0.0, last_cascade_ms = 0.0;
  uint64_t launches = 0, erode_calls = 0;
  double last_erode_ms = 0.0, last_sweep_ms = 0.0, last_fly_ms = 0.0;

  size_t cells() const { return (size_t)m.rows * (size_t)m.nz; }
};

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

int check_launch(swe_engine* e, const char* what, int n = 1) {
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return fail(SWE_ERR_CUDA, "%s launch failed: %s", what, cudaGetErrorString(err));
  e->launches += (uint64_t)n;
  return SWE_OK;
}

void finish_timers(swe_engine* e) {
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

int swe_create(swe_engine** out, int nx, int nz, const swe_params* params, int device) {
  if (!out) return fail(SWE_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (nx < 1 || nz < 1) return fail(SWE_ERR_INVALID, "map size must be >= 1x1 (got %dx%d)", nx, nz);
  if ((long long)nx * nz > (1LL << 31) - 1)
    return fail(SWE_ERR_INVALID, "map of %dx%d cells exceeds 2^31-1 cells", nx, nz);
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
  const size_t n = e->cells();
  cudaError_t err;
#define TRY(call) if ((err = (call)) != cudaSuccess) goto cuda_fail
  TRY(cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking));
  e->stream = e->own_stream;
  TRY(cudaEventCreate(&e->ev_a)); TRY(cudaEventCreate(&e->ev_b));
  TRY(cudaEventCreate(&e->ev_c)); TRY(cudaEventCreate(&e->ev_d));
  TRY(cudaEventCreate(&e->ev_f0)); TRY(cudaEventCreate(&e->ev_f1));
  TRY(cudaMalloc(&e->m.height, n * sizeof(float)));
  TRY(cudaMalloc(&e->m.field, n * sizeof(float4)));
  TRY(cudaMalloc(&e->m.track, n * sizeof(float4)));
  TRY(cudaMalloc(&e->d_stats, sizeof(DevStats)));
  TRY(cudaMalloc(&e->d_rng, sizeof(DetRand)));
  TRY(cudaMalloc(&e->d_small, 64 * sizeof(float)));
  TRY(cudaMemsetAsync(e->m.height, 0, n * sizeof(float), e->stream));
  TRY(cudaMemsetAsync(e->m.field, 0, n * sizeof(float4), e->stream));
  TRY(cudaMemsetAsync(e->m.track, 0, n * sizeof(float4), e->stream));
  TRY(cudaMemsetAsync(e->d_stats, 0, sizeof(DevStats), e->stream));
  make_height_tensor_map(e->m, &e->tmap);   // optional: plain-load kernels are used when it fails
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

void swe_destroy(swe_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  cudaFree(e->m.height); cudaFree(e->m.field); cudaFree(e->m.track);
  cudaFree(e->d_stats); cudaFree(e->d_rng); cudaFree(e->d_small); cudaFree(e->d_scratch);
  cudaFree(e->pbuf.a); cudaFree(e->pbuf.b); cudaFree(e->pbuf.live); cudaFree(e->d_stamp);
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

// ---- transfers ----------------------------------------------------------------------------
int swe_upload_plane(swe_engine* e, int plane, const float* src, ptrdiff_t stride_bytes) {
  if (int rc = bind(e)) return rc;
  if (!src) return fail(SWE_ERR_INVALID, "src is NULL");
  if (plane < 0 || plane >= SWE_NUM_PLANES) return fail(SWE_ERR_INVALID, "plane %d out of range", plane);
  if (stride_bytes < 4 || stride_bytes % 4) return fail(SWE_ERR_INVALID, "stride_bytes must be a multiple of 4, >= 4");
  const size_t n = e->cells(), bytes = n * sizeof(float);
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
    SWE_CUDA(cudaMemcpyAsync(e->m.height, packed, bytes, cudaMemcpyHostToDevice, e->stream));
  } else {
    if (int rc = need_scratch(e, bytes)) return rc;
    SWE_CUDA(cudaMemcpyAsync(e->d_scratch, packed, bytes, cudaMemcpyHostToDevice, e->stream));
    launch_pack_plane(e->m, plane, e->d_scratch, e->stream);
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
  const size_t n = e->cells(), bytes = n * sizeof(float);
  const float* d_src = e->m.height;
  if (plane != SWE_HEIGHT) {
    if (int rc = need_scratch(e, bytes)) return rc;
    launch_unpack_plane(e->m, plane, e->d_scratch, e->stream);
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

int swe_upload_cells(swe_engine* e, const void* cells) {
  if (int rc = bind(e)) return rc;
  if (!cells) return fail(SWE_ERR_INVALID, "cells is NULL");
  const size_t bytes = e->cells() * 9 * sizeof(float);
  if (int rc = need_scratch(e, bytes)) return rc;
  SWE_CUDA(cudaMemcpyAsync(e->d_scratch, cells, bytes, cudaMemcpyHostToDevice, e->stream));
  launch_cells_to_planes(e->m, e->d_scratch, e->stream);
  if (int rc = check_launch(e, "cells_to_planes")) return rc;
  SWE_CUDA(cudaStreamSynchronize(e->stream));
  return SWE_OK;
}

int swe_download_cells(swe_engine* e, void* cells) {
  if (int rc = bind(e)) return rc;
  if (!cells) return fail(SWE_ERR_INVALID, "cells is NULL");
  const size_t bytes = e->cells() * 9 * sizeof(float);
  if (int rc = need_scratch(e, bytes)) return rc;
  launch_planes_to_cells(e->m, e->d_scratch, e->stream);
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
  finish_timers(e);
  SWE_CUDA(cudaEventRecord(e->ev_a, e->stream));
  // world.hpp:38-44: reset the four track members of every cell
  SWE_CUDA(cudaMemsetAsync(e->m.track, 0, e->cells() * sizeof(float4), e->stream));
  // world.hpp:47-66: the particles
  SWE_CUDA(cudaEventRecord(e->ev_f0, e->stream));
  if (mode == SWE_MODE_DETERMINISTIC) {
    launch_det_erode(e->m, e->p, cycles, e->d_rng, e->d_stats, e->stream);
    if (int rc = check_launch(e, "det_erode")) return rc;
  } else if (cycles > 0) {
    if (int rc = need_particles(e, cycles)) return rc;
    if (int rc = need_stamp(e)) return rc;
    const int k = e->opt_cascade_interval;
    const int rounds = (e->p.max_age + 2 + k - 1) / k;   // every particle is dead after maxAge+2 calls
    launch_par_begin(e->pbuf, e->m, cycles, e->seed, e->call_index, e->d_stats, e->stream);
    if (int rc = check_launch(e, "par_begin", 2)) return rc;
    for (int r = 0; r < rounds; r++) {
      if (++e->stamp_counter > 255u) {   // stamp values 1..255; recycle with a clean plane
        SWE_CUDA(cudaMemsetAsync(e->d_stamp, 0, e->cells(), e->stream));
        e->stamp_counter = 1;
      }
      const unsigned char now = (unsigned char)e->stamp_counter;
      const bool kt = e->opt_kernel_timing != 0;
      if (kt) {
        while ((int)e->kt_events.size() < 3 * (r + 1)) {
          cudaEvent_t ev; SWE_CUDA(cudaEventCreate(&ev)); e->kt_events.push_back(ev);
        }
        SWE_CUDA(cudaEventRecord(e->kt_events[3 * r], e->stream));
      }
      launch_par_round(e->pbuf, e->m, e->p, cycles, r, k, e->seed, e->call_index, e->d_stamp, now,
                       e->d_stats, e->stream);
      if (kt) SWE_CUDA(cudaEventRecord(e->kt_events[3 * r + 1], e->stream));
      // wind.hpp:114, deferred: World::cascade once at every cell touched in this round
      launch_cascade_stamped(e->m, e->p, &e->tmap, e->d_stamp, now, e->pbuf.live + r, e->stream);
      if (kt) { SWE_CUDA(cudaEventRecord(e->kt_events[3 * r + 2], e->stream)); e->kt_rounds = r + 1; }
      if (int rc = check_launch(e, "par_round", 2)) return rc;
    }
  }
  SWE_CUDA(cudaEventRecord(e->ev_f1, e->stream));
  // world.hpp:70-76: EMA of the flow fields
  launch_ema(e->m, e->p.lrate, e->stream);
  if (int rc = check_launch(e, "ema")) return rc;
  SWE_CUDA(cudaEventRecord(e->ev_b, e->stream));
  e->erode_timed = true;
  e->call_index++;
  e->erode_calls++;
  return SWE_OK;
}

int swe_cascade(swe_engine* e, int x, int z) {
  if (int rc = bind(e)) return rc;
  if (e->m.oob(x, z)) return fail(SWE_ERR_OOB, "cell (%d,%d) outside %dx%d", x, z, e->m.nx, e->m.nz);
  launch_det_cascade_one(e->m, e->p, x, z, e->stream);
  return check_launch(e, "cascade_one");
}

int swe_max_excess(swe_engine* e, float* out) {
  if (int rc = bind(e)) return rc;
  if (!out) return fail(SWE_ERR_INVALID, "out is NULL");
  launch_max_excess(e->m, e->p, e->d_small, e->stream);
  if (int rc = check_launch(e, "max_excess", 2)) return rc;
  SWE_CUDA(cudaMemcpyAsync(out, e->d_small, sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  SWE_CUDA(cudaStreamSynchronize(e->stream));
  finish_timers(e);
  return SWE_OK;
}

int swe_cascade_sweep(swe_engine* e, int sweeps, float* max_excess_out) {
  if (int rc = bind(e)) return rc;
  if (sweeps < 0) return fail(SWE_ERR_INVALID, "sweeps < 0");
  finish_timers(e);
  SWE_CUDA(cudaEventRecord(e->ev_c, e->stream));
  const int n = launch_cascade_sweep(e->m, e->p, &e->tmap, sweeps, e->d_stats, e->stream, e->opt_cascade_variant);
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
  if (e->m.oob(x, z)) return fail(SWE_ERR_OOB, "cell (%d,%d) outside %dx%d", x, z, e->m.nx, e->m.nz);
  launch_normal_one(e->m, e->p, x, z, e->d_small, e->stream);
  if (int rc = check_launch(e, "normal_one")) return rc;
  SWE_CUDA(cudaMemcpyAsync(out3, e->d_small, 3 * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  SWE_CUDA(cudaStreamSynchronize(e->stream));
  return SWE_OK;
}

int swe_height(swe_engine* e, int x, int z, float* out) {
  if (int rc = bind(e)) return rc;
  if (!out) return fail(SWE_ERR_INVALID, "out is NULL");
  if (e->m.oob(x, z)) return fail(SWE_ERR_OOB, "cell (%d,%d) outside %dx%d", x, z, e->m.nx, e->m.nz);
  SWE_CUDA(cudaMemcpyAsync(out, e->m.height + e->m.idx(x, z), sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  SWE_CUDA(cudaStreamSynchronize(e->stream));
  return SWE_OK;
}

int swe_download_normals(swe_engine* e, float* dst) {
  if (int rc = bind(e)) return rc;
  if (!dst) return fail(SWE_ERR_INVALID, "dst is NULL");
  const size_t bytes = e->cells() * 3 * sizeof(float);
  if (int rc = need_scratch(e, bytes)) return rc;
  launch_normals(e->m, e->p, e->d_scratch, e->stream);
  if (int rc = check_launch(e, "normals")) return rc;
  SWE_CUDA(cudaMemcpyAsync(dst, e->d_scratch, bytes, cudaMemcpyDeviceToHost, e->stream));
  SWE_CUDA(cudaStreamSynchronize(e->stream));
  return SWE_OK;
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
  out->last_fly_ms = e->last_fly_ms;
  out->last_round_ms = e->last_round_ms;
  out->last_cascade_ms = e->last_cascade_ms;
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
  launch_sum_plane(e->m, plane, d_out, e->stream);
  if (int rc = check_launch(e, "sum_plane")) return rc;
  SWE_CUDA(cudaMemcpyAsync(out, d_out, sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  SWE_CUDA(cudaStreamSynchronize(e->stream));
  return SWE_OK;
}

int swe_hash_height(swe_engine* e, uint64_t* out) {
  if (int rc = bind(e)) return rc;
  if (!out) return fail(SWE_ERR_INVALID, "out is NULL");
  std::vector<uint32_t> host(e->cells());
  SWE_CUDA(cudaMemcpyAsync(host.data(), e->m.height, host.size() * 4, cudaMemcpyDeviceToHost, e->stream));
  SWE_CUDA(cudaStreamSynchronize(e->stream));
  uint64_t h = 1469598103934665603ull;
  for (uint32_t w : host) { h ^= w; h *= 1099511628211ull; }
  *out = h;
  return SWE_OK;
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
