// swe_det.cu -- deterministic ("batch-size-1") mode: the reference's exact sequential
// semantics on the device. One thread flies one particle at a time in the reference's
// order, draws from a replica of glibc's rand() stream and rounds every operation like the
// reference binary, so heights, tracks and trajectories are BITWISE identical to
// World::erode / Wind::fly / World::cascade (source/world.hpp:36-153, source/wind.hpp:28-117).
// This mode exists to pin the arithmetic (parity (1) of the north star); throughput comes
// from swe_par.cu.
//
// Compiled with -fmad=false (no FMA contraction), IEEE div/sqrt.
#include "swe_exact.cuh"

namespace swe {
using namespace exact;

// ---- glibc rand(): stdlib/random_r.c TYPE_3 (deg 31, sep 3) --------------------------------
__device__ __forceinline__ int det_rand(DetRand& s) {
  const uint32_t v = (uint32_t)s.r[s.f] + (uint32_t)s.r[s.b];
  s.r[s.f] = (int32_t)v;
  if (++s.f >= 31) s.f = 0;
  if (++s.b >= 31) s.b = 0;
  return (int)(v >> 1);
}

__global__ void det_srand_kernel(DetRand* out, uint32_t seed) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  DetRand s;
  int32_t word = (int32_t)(seed ? seed : 1u);
  s.r[0] = word;
  for (int i = 1; i < 31; i++) {
    const long long hi = word / 127773, lo = word % 127773;
    long long w = 16807 * lo - 2836 * hi;
    if (w < 0) w += 2147483647;
    word = (int32_t)w;
    s.r[i] = word;
  }
  s.f = 3; s.b = 0;
  for (int i = 0; i < 310; i++) (void)det_rand(s);
  *out = s;
}

struct Particle {
  int age;
  float sediment;
  V3 pos, speed;
};

__device__ __forceinline__ void particle_at(Particle& q, float x, float z) {  // wind.hpp:6-12
  q.age = 0; q.sediment = 0.0f;
  q.pos = v3(x, 0.0f, z);
  q.speed = v3(0.0f, 0.0f, 0.0f);
}

// Wind::fly, wind.hpp:28-117. Returns false when the particle dies.
__device__ bool det_fly(const MapView& m, const swe_params& P, const CascadeThr thr, Particle& q,
                        DetRand& rng) {
  const int ix = (int)roundf(q.pos.x), iz = (int)roundf(q.pos.z);          // :30
  if (m.oob(ix, iz)) return false;                                           // :32-38
  const size_t ind = m.idx(ix, iz);
  const float h = m.height[ind];
  const V3 n = normal(m, P, ix, iz);                                         // :40
  const V3 pspeed = v3(P.pspeed[0], P.pspeed[1], P.pspeed[2]);

  if (q.age == 0 || q.pos.y < h) q.pos.y = h;                                // :42-43
  if (q.age++ > P.max_age) return false;                                     // :45-46

  float hfac = expf_glibc(fdiv(-fsub(q.pos.y, h), P.boundarylayer));         // :50
  if (hfac < 0) hfac = 0;

  float shadow = dot(normalize(pspeed), n);                                  // :56-59
  if (shadow < 0) shadow = 0;
  shadow = fsub(1.0f, shadow);

  V3 speed = q.speed;
  speed = add(speed, scale(0.05f, sub(scale(fadd(0.1f, fmul(0.9f, shadow)), pspeed), speed)));  // :61

  if (q.pos.y > h) speed.y = fsub(speed.y, fmul(P.gravity, q.sediment));     // :65-66

  float collision = -dot(normalize(speed), n);                               // :70-71
  if (collision < 0) collision = 0;

  const V3 rspeed = cross(n, cross(scale(fsub(1.0f, collision), speed), n)); // :75

  speed = add(speed, scale(0.9f, sub(scale(shadow, mix(pspeed, rspeed, fmul(shadow, hfac))), speed)));  // :79

  {                                                                          // :83
    // g++ evaluates the vec3 constructor arguments right to left: z draws first
    const int rz = det_rand(rng) % 1001;
    const int ry = det_rand(rng) % 1001;
    const int rx = det_rand(rng) % 1001;
    V3 t = v3(fsub((float)rx, 500.0f), fsub((float)ry, 500.0f), fsub((float)rz, 500.0f));
    t = scale(fmul(fmul(0.1f, hfac), collision), t);
    t = v3(fdiv(t.x, 500.0f), fdiv(t.y, 500.0f), fdiv(t.z, 500.0f));
    speed = add(speed, t);
  }
  {                                                                          // :87 (double factor)
    const float f = __double2float_rn(__dsub_rn(1.0, __dmul_rn(0.3, (double)q.sediment)));
    speed = v3(fmul(speed.x, f), fmul(speed.y, f), fmul(speed.z, f));
  }

  q.pos = add(q.pos, speed);                                                 // :91

  {                                                                          // :93-96
    float4 t = m.track[ind];
    t.y = fadd(t.y, speed.x);
    t.z = fadd(t.z, speed.y);
    t.w = fadd(t.w, speed.z);
    t.x = fadd(t.x, q.sediment);
    m.track[ind] = t;
  }

  float force = fmul(-dot(normalize(speed), n), length(speed));              // :100-102
  if (force < 0) force = 0;
  const float lift = fmul(fsub(1.0f, collision), length(speed));             // :104
  const float capacity = fadd(fmul(force, hfac), fmul(fmul(0.02f, lift), hfac));  // :106

  const float diff = fsub(capacity, q.sediment);                             // :110-112
  m.height[ind] = fsub(m.height[ind], fmul(P.suspension, diff));
  q.sediment = fadd(q.sediment, fmul(P.suspension, diff));
  q.speed = speed;

  cascade_global(m, P, thr, ix, iz);                                         // :114
  return true;
}

// World::erode particle loop, world.hpp:47-66 (track reset and EMA are separate passes).
__global__ void det_erode_kernel(MapView m, swe_params P, int cycles, DetRand* g_rng,
                                 DevStats* stats) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  DetRand rng = *g_rng;
  const CascadeThr thr = cascade_thresholds(P);
  unsigned long long steps = 0, n_oob = 0, n_age = 0;
  double lost = 0.0;
  for (int i = 0; i < cycles; i++) {
    // world.hpp:52-54; constructor arguments are evaluated right to left (z first)
    const int iz = det_rand(rng) % m.nz;
    const int ix = det_rand(rng) % m.nx;
    const int fz = det_rand(rng) % 1000;
    const int fx = det_rand(rng) % 1000;
    const float x = fadd((float)ix, fdiv((float)fx, 1000.0f));
    const float z = fadd((float)iz, fdiv((float)fz, 1000.0f));
    Particle q;
    particle_at(q, x, z);
    while (det_fly(m, P, thr, q, rng)) steps++;
    if (q.age == P.max_age + 2) n_age++; else n_oob++;
    lost += (double)q.sediment;
  }
  *g_rng = rng;
  stats->particles += (unsigned long long)cycles;
  stats->steps += steps;
  stats->exits_oob += n_oob;
  stats->exits_age += n_age;
  stats->sediment_lost += lost;
}

__global__ void det_trace_kernel(MapView m, swe_params P, float x, float z, int max_records,
                                 float* records, int* steps_out, swe_wind* final_out,
                                 DetRand* g_rng, DevStats* stats) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  DetRand rng = *g_rng;
  const CascadeThr thr = cascade_thresholds(P);
  Particle q;
  particle_at(q, x, z);
  int n = 0;
  while (det_fly(m, P, thr, q, rng)) {
    if (records && n < max_records) {
      float* r = records + (size_t)n * 8;
      r[0] = q.pos.x; r[1] = q.pos.y; r[2] = q.pos.z;
      r[3] = q.speed.x; r[4] = q.speed.y; r[5] = q.speed.z;
      r[6] = q.sediment; r[7] = (float)q.age;
    }
    n++;
  }
  *g_rng = rng;
  *steps_out = n;
  final_out->age = q.age; final_out->sediment = q.sediment;
  final_out->pos[0] = q.pos.x; final_out->pos[1] = q.pos.y; final_out->pos[2] = q.pos.z;
  final_out->speed[0] = q.speed.x; final_out->speed[1] = q.speed.y; final_out->speed[2] = q.speed.z;
  stats->particles += 1;
  stats->steps += (unsigned long long)n;
  if (q.age == P.max_age + 2) stats->exits_age += 1; else stats->exits_oob += 1;
  stats->sediment_lost += (double)q.sediment;
}

__global__ void det_fly_once_kernel(MapView m, swe_params P, swe_wind* w, int* alive,
                                    DetRand* g_rng, DevStats* stats) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  DetRand rng = *g_rng;
  const CascadeThr thr = cascade_thresholds(P);
  Particle q;
  q.age = w->age; q.sediment = w->sediment;
  q.pos = v3(w->pos[0], w->pos[1], w->pos[2]);
  q.speed = v3(w->speed[0], w->speed[1], w->speed[2]);
  const bool ok = det_fly(m, P, thr, q, rng);
  *g_rng = rng;
  *alive = ok ? 1 : 0;
  w->age = q.age; w->sediment = q.sediment;
  w->pos[0] = q.pos.x; w->pos[1] = q.pos.y; w->pos[2] = q.pos.z;
  w->speed[0] = q.speed.x; w->speed[1] = q.speed.y; w->speed[2] = q.speed.z;
  if (ok) stats->steps += 1;
}

__global__ void det_cascade_one_kernel(MapView m, swe_params P, int x, int z) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  cascade_global(m, P, cascade_thresholds(P), x, z);
}

// quad::map::normal for every cell (what the renderer's updatenode reads, cellpool.h:302-321)
__global__ void normals_kernel(MapView m, swe_params P, int row_lo, int row_hi, float* out3) {
  const size_t n = (size_t)(row_hi - row_lo) * m.nz;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const int x = row_lo + (int)(i / m.nz), z = (int)(i % m.nz);
    const V3 nn = normal(m, P, x, z);
    out3[3 * i + 0] = nn.x; out3[3 * i + 1] = nn.y; out3[3 * i + 2] = nn.z;
  }
}

// The Vertex record updatenode writes for every cell each frame (cellpool.h:302-321,
// vertexpool.hpp:6-26): position = (x, mapscale*h, z), normal = map.normal(p), tangent = T - P and
// bitangent = B - P with T, B the surface points at (x+1, z) and (x, z+1); a neighbour outside the
// map has height 0 (node::height, cellpool.h:257-261). 12 floats = 48 B per cell, written as three
// 16-byte stores; 4 B read + 48 B written per cell, HBM-bound.
__global__ void vertices_kernel(MapView m, swe_params P, int row_lo, int row_hi, float4* __restrict__ out) {
  const size_t n = (size_t)(row_hi - row_lo) * m.nz;
  const float scale = (float)P.mapscale;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const int x = row_lo + (int)(i / m.nz), z = (int)(i % m.nz);
    const V3 nn = normal(m, P, x, z);
    const float py = fmul(scale, m.height[m.idx(x, z)]);
    const float ty = fmul(scale, m.oob(x + 1, z) ? 0.0f : m.height[m.idx(x + 1, z)]);
    const float by = fmul(scale, m.oob(x, z + 1) ? 0.0f : m.height[m.idx(x, z + 1)]);
    const float fx = (float)x, fz = (float)z;
    const float tx = fsub((float)(x + 1), fx), bz = fsub((float)(z + 1), fz);   // 1 below 2^24
    out[3 * i + 0] = make_float4(fx, py, fz, nn.x);
    out[3 * i + 1] = make_float4(nn.y, nn.z, tx, fsub(ty, py));
    out[3 * i + 2] = make_float4(fsub(fz, fz), fsub(fx, fx), fsub(by, py), bz);
  }
}

__global__ void normal_one_kernel(MapView m, swe_params P, int x, int z, float* out3) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const V3 nn = normal(m, P, x, z);
  out3[0] = nn.x; out3[1] = nn.y; out3[2] = nn.z;
}

// ---- launchers ----------------------------------------------------------------------------
void launch_det_erode(const MapView& m, const swe_params& p, int cycles, DetRand* rng,
                      DevStats* stats, cudaStream_t st) {
  det_erode_kernel<<<1, 32, 0, st>>>(m, p, cycles, rng, stats);
}
void launch_det_trace(const MapView& m, const swe_params& p, float x, float z, int max_records,
                      float* d_records, int* d_steps, swe_wind* d_final, DetRand* rng,
                      DevStats* stats, cudaStream_t st) {
  det_trace_kernel<<<1, 32, 0, st>>>(m, p, x, z, max_records, d_records, d_steps, d_final, rng, stats);
}
void launch_det_fly_once(const MapView& m, const swe_params& p, swe_wind* d_wind, int* d_alive,
                         DetRand* rng, DevStats* stats, cudaStream_t st) {
  det_fly_once_kernel<<<1, 32, 0, st>>>(m, p, d_wind, d_alive, rng, stats);
}
void launch_det_cascade_one(const MapView& m, const swe_params& p, int x, int z, cudaStream_t st) {
  det_cascade_one_kernel<<<1, 32, 0, st>>>(m, p, x, z);
}
void launch_det_srand(DetRand* rng, uint32_t seed, cudaStream_t st) {
  det_srand_kernel<<<1, 32, 0, st>>>(rng, seed);
}
void launch_normals_rows(const MapView& m, const swe_params& p, int row_lo, int row_hi, float* d_out3,
                         cudaStream_t st) {
  const size_t n = (size_t)(row_hi - row_lo) * m.nz;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  normals_kernel<<<blocks, 256, 0, st>>>(m, p, row_lo, row_hi, d_out3);
}
void launch_vertices_rows(const MapView& m, const swe_params& p, int row_lo, int row_hi, void* d_out48,
                          cudaStream_t st) {
  const size_t n = (size_t)(row_hi - row_lo) * m.nz;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  vertices_kernel<<<blocks, 256, 0, st>>>(m, p, row_lo, row_hi, static_cast<float4*>(d_out48));
}
void launch_normal_one(const MapView& m, const swe_params& p, int x, int z, float* d_out3,
                       cudaStream_t st) {
  normal_one_kernel<<<1, 32, 0, st>>>(m, p, x, z, d_out3);
}

// Force-load this translation unit's kernels (CUDA loads modules lazily on first launch, and a
// first launch issued while another stream spins in a stripe wait kernel would stall behind it).
void preload_det() {
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, det_srand_kernel);
  cudaFuncGetAttributes(&a, det_erode_kernel);
  cudaFuncGetAttributes(&a, det_trace_kernel);
  cudaFuncGetAttributes(&a, det_fly_once_kernel);
  cudaFuncGetAttributes(&a, det_cascade_one_kernel);
  cudaFuncGetAttributes(&a, normals_kernel);
  cudaFuncGetAttributes(&a, vertices_kernel);
  cudaFuncGetAttributes(&a, normal_one_kernel);
}

}  // namespace swe
