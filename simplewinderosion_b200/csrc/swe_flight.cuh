// swe_flight.cuh -- pieces shared by the parallel flight kernels (swe_par.cu: round-synchronous
// pipeline of the stripe engines; swe_queue.cu: colour-queue scheduler of the single-GPU engine):
// constants, counter-based RNG, atomic scatter helpers, the stratified spawn.
#pragma once
#include <math.h>

#include "swe_common.cuh"

namespace swe {

struct ParConst {
  float inv_boundarylayer, suspension, gravity, mapscale, abrasion;
  float px, py, pz;      // pspeed
  float pnx, pny, pnz;   // normalize(pspeed)
  int max_age;
};

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
  }
  return c;
}

__device__ __forceinline__ void red_add_f32(float* addr, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void red_add_v4(float4* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c),
               "f"(d)
               : "memory");
}

struct PState {
  float px, py, pz, sx, sy, sz, sed;
  int age;
};

// Wind ctor + spawn (world.hpp:52-54: cell = (rand()%nx, rand()%nz), plus (rand()%1000)/1000 per
// axis). The reference draws iid uniform cells; here the n particles of a call are a STRATIFIED
// uniform sample: the cells are ordered strip-major (strips of 4 rows, z-major inside a strip)
// and particle i takes the cell of rank floor((i + u) * ncells / n), u ~ U[0,1). Every cell is
// equally likely, as in the reference, and particles adjacent in memory start adjacent on the map
// (a warp covers ~4 rows x 32 columns at 0.25 particles/cell), so a warp's neighbourhood reads
// and atomics fall into a handful of 32-byte sectors instead of 32 scattered ones.
__device__ __forceinline__ void spawn_position(int i, int n, int nx, int nz, uint2 key, uint32_t call,
                                               float& px, float& pz) {
  const uint4 r = philox4x32_10(make_uint4((uint32_t)i, call, 0u, 0x5u), key);
  const unsigned long long ncells = (unsigned long long)nx * (unsigned long long)nz;
  unsigned long long rank =
      ((unsigned long long)i * ncells + (((unsigned long long)r.x * ncells) >> 32)) / (unsigned long long)n;
  if (rank >= ncells) rank = ncells - 1;
  const unsigned long long strip_cells = 4ull * (unsigned long long)nz;
  const int strip = (int)(rank / strip_cells);
  const unsigned q = (unsigned)(rank - (unsigned long long)strip * strip_cells);
  const int rows = min(4, nx - 4 * strip);
  const int cz = (int)(q / (unsigned)rows), cx = 4 * strip + (int)(q - (unsigned)cz * (unsigned)rows);
  px = (float)cx + (float)__umulhi(r.z, 1000u) / 1000.0f;
  pz = (float)cz + (float)__umulhi(r.w, 1000u) / 1000.0f;
}

// The reference's own spawn (world.hpp:52-54): an iid uniform cell (rand()%nx, rand()%nz) plus
// (rand()%1000)/1000 per axis. Used by the colour-queue scheduler, whose staggered injection takes
// every S-th particle id per micro-round: with the strip-major stratified order above that would
// sweep a regular lattice of spawn points across the map over the call (measured on B200: the
// along-wind erosion profile then correlates only 0.85 with the sequential oracle).
__device__ __forceinline__ void spawn_position_iid(int i, int nx, int nz, uint2 key, uint32_t call, float& px, float& pz) {
  const uint4 r = philox4x32_10(make_uint4((uint32_t)i, call, 0u, 0x6u), key);
  px = (float)__umulhi(r.x, (unsigned)nx) + (float)__umulhi(r.z, 1000u) / 1000.0f;
  pz = (float)__umulhi(r.y, (unsigned)nz) + (float)__umulhi(r.w, 1000u) / 1000.0f;
}

static inline ParConst make_const(const swe_params& p) {
  ParConst C;
  C.inv_boundarylayer = 1.0f / p.boundarylayer;
  C.suspension = p.suspension; C.gravity = p.gravity;
  C.mapscale = (float)p.mapscale;
  C.abrasion = p.abrasion;
  C.px = p.pspeed[0]; C.py = p.pspeed[1]; C.pz = p.pspeed[2];
  const float pl = sqrtf(C.px * C.px + C.py * C.py + C.pz * C.pz);
  C.pnx = C.px / pl; C.pny = C.py / pl; C.pnz = C.pz / pl;
  C.max_age = p.max_age;
  return C;
}


}  // namespace swe
