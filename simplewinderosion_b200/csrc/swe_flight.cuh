// swe_flight.cuh -- pieces shared by the parallel flight kernels (swe_par.cu: round-synchronous
// pipeline; swe_queue.cu: colour-queue scheduler; swe_tile.cu: tile-resident scheduler):
// constants, counter-based RNG, atomic scatter helpers, the spawn, and Wind::fly for a particle
// that owns its cell (fly_owned).
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

template <int ROUNDS = 10>
__device__ __forceinline__ uint4 philox4x32(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < ROUNDS; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) { return philox4x32<10>(c, k); }

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

__device__ __forceinline__ int cell_colour(int x, int z) { return (x % 3) * 3 + (z % 3); }

// Wind::fly (wind.hpp:28-112) for a particle that owns its cell: heights come in registers, the
// erosion amount goes out. Returns dh: the centre cell loses dh (wind.hpp:110-111).
template <bool ALWAYS_RNG = false, int RNG_ROUNDS = 10>   // ALWAYS_RNG: draw the turbulence numbers unconditionally (no branch in the middle of the step)
__device__ __forceinline__ float fly_owned(const ParConst& C, PState& q, float hc, float hxm, float hxp, float hzm,
                                           float hzp, bool xm, bool xp, bool zm, bool zp, uint4 rng_ctr, uint2 key,
                                           float& collision_out, float& tx, float& ty, float& tz, float& tsed) {
  float nxv, nyv, nzv;   // quad::_normal, closed form (cellpool.h:200-223; see swe_par.cu)
  {
    const float S = C.mapscale;
    const float cz = (zp ? 1.0f : 0.0f) + (zm ? 1.0f : 0.0f);
    const float cx = (xp ? 1.0f : 0.0f) + (xm ? 1.0f : 0.0f);
    nxv = S * cz * ((hxm - hc) - (hxp - hc));
    nyv = cx * cz;
    nzv = S * cx * ((hzm - hc) - (hzp - hc));
    const float l2 = nxv * nxv + nyv * nyv + nzv * nzv;
    if (l2 > 0.0f) { const float il = rsqrtf(l2); nxv *= il; nyv *= il; nzv *= il; }
  }
  if (q.age == 0 || q.py < hc) q.py = hc;                                         // :42-43
  q.age++;                                                                        // :45 (age limit: caller)
  const float hfac = __expf(-(q.py - hc) * C.inv_boundarylayer);                 // :50
  float shadow = C.pnx * nxv + C.pny * nyv + C.pnz * nzv;                         // :56-59
  shadow = 1.0f - fmaxf(shadow, 0.0f);
  float sx = q.sx, sy = q.sy, sz = q.sz;
  {
    const float k = 0.1f + 0.9f * shadow;                                         // :61
    sx += 0.05f * (k * C.px - sx);
    sy += 0.05f * (k * C.py - sy);
    sz += 0.05f * (k * C.pz - sz);
  }
  if (q.py > hc) sy -= C.gravity * q.sed;                                         // :65-66
  float collision;
  {
    const float il = rsqrtf(sx * sx + sy * sy + sz * sz);                         // :70-71
    collision = fmaxf(-(sx * nxv + sy * nyv + sz * nzv) * il, 0.0f);
  }
  {
    const float k = 1.0f - collision;                                             // :75  n x (a x n)
    const float ax = k * sx, ay = k * sy, az = k * sz;
    const float nn = nxv * nxv + nyv * nyv + nzv * nzv;
    const float na = nxv * ax + nyv * ay + nzv * az;
    const float rx = ax * nn - nxv * na, ry = ay * nn - nyv * na, rz = az * nn - nzv * na;
    const float w = shadow * hfac;                                                // :79
    sx += 0.9f * (shadow * (C.px * (1.0f - w) + rx * w) - sx);
    sy += 0.9f * (shadow * (C.py * (1.0f - w) + ry * w) - sy);
    sz += 0.9f * (shadow * (C.pz * (1.0f - w) + rz * w) - sz);
  }
  {
    const float amp = 0.1f * hfac * collision;                                    // :83
    if (ALWAYS_RNG || amp > 0.0f) {
      const uint4 r = philox4x32<RNG_ROUNDS>(rng_ctr, key);
      const float s = amp * (1.0f / 500.0f);
      sx += s * ((float)__umulhi(r.x, 1001u) - 500.0f);
      sy += s * ((float)__umulhi(r.y, 1001u) - 500.0f);
      sz += s * ((float)__umulhi(r.z, 1001u) - 500.0f);
    }
  }
  {
    const float f = 1.0f - 0.3f * q.sed;                                          // :87
    sx *= f; sy *= f; sz *= f;
  }
  q.px += sx; q.py += sy; q.pz += sz;                                             // :91
  q.sx = sx; q.sy = sy; q.sz = sz;
  tx = sx; ty = sy; tz = sz; tsed = q.sed;                                        // :93-96
  const float len = sqrtf(sx * sx + sy * sy + sz * sz);
  const float force = fmaxf(-(sx * nxv + sy * nyv + sz * nzv), 0.0f);             // :100-102
  const float lift = (1.0f - collision) * len;                                    // :104
  const float capacity = hfac * (force + 0.02f * lift);                           // :106
  collision_out = collision;
  return C.suspension * (capacity - q.sed);                                       // :110
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
