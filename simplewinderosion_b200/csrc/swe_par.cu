// swe_par.cu -- parallel flight: all particles of a World::erode call in flight at once.
//
// One thread per particle (persistent, grid-stride respawn), same physics as Wind::fly
// (source/wind.hpp:28-117) and World::cascade (source/world.hpp:79-153) with
//   * a counter-based RNG (Philox4x32-10, counter = particle id / call / age) instead of the
//     sequential libc rand() stream -- same distributions (rand()%1001, rand()%1000, rand()%n),
//   * L2-coherent neighbourhood reads (ld.global.cg: other SMs' atomics are visible),
//   * atomic scatter: one red.global.add.v4.f32 for the four track values (wind.hpp:93-96),
//     red.global.add.f32 for the erosion/deposition of the centre cell (wind.hpp:110-112)
//     and for every cascade transfer (world.hpp:142-149); each transfer is applied as a
//     +t/-t pair so sand mass is conserved to fp32 round-off whatever the interleaving,
//   * fast-math (FMA contraction, rsqrt, ex2-based exp): parity here is statistical.
// Sediment carried by a particle when it dies leaves the system, as in the reference; it is
// accumulated in DevStats::sediment_lost so tests can close the mass balance.
#include "swe_common.cuh"

namespace swe {

struct ParConst {
  float inv_boundarylayer, suspension, gravity, settling_half, thr_ortho, thr_diag, mapscale;
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

// compare-exchange for a descending sort of (height, slot) pairs
__device__ __forceinline__ void cex(float& ha, int& ia, float& hb, int& ib) {
  const bool sw = hb > ha;
  const float th = sw ? hb : ha, tl = sw ? ha : hb;
  const int ti = sw ? ib : ia, tj = sw ? ia : ib;
  ha = th; hb = tl; ia = ti; ib = tj;
}

// One Wind::fly step. Returns 0 = alive, 1 = left the map, 2 = too old.
__device__ __forceinline__ int par_step(const MapView& m, const ParConst& C, PState& q,
                                        uint32_t pid, uint2 key, uint32_t call) {
  const int ix = __float2int_rd(q.px + 0.5f), iz = __float2int_rd(q.pz + 0.5f);  // wind.hpp:30
  if (m.oob(ix, iz)) return 1;                                                    // :32-38
  const int nz = m.nz;
  float* hc_ptr = m.height + m.idx(ix, iz);
  const bool xm = ix > 0, xp = ix + 1 < m.nx, zm = iz > 0, zp = iz + 1 < nz;

  // 3x3 neighbourhood, reference slot order (world.hpp:83-92); oob -> centre value, masked
  const float hc = __ldcg(hc_ptr);
  float h[8];
  h[0] = (xm && zm) ? __ldcg(hc_ptr - nz - 1) : hc;
  h[1] = xm ? __ldcg(hc_ptr - nz) : hc;
  h[2] = (xm && zp) ? __ldcg(hc_ptr - nz + 1) : hc;
  h[3] = zm ? __ldcg(hc_ptr - 1) : hc;
  h[4] = zp ? __ldcg(hc_ptr + 1) : hc;
  h[5] = (xp && zm) ? __ldcg(hc_ptr + nz - 1) : hc;
  h[6] = xp ? __ldcg(hc_ptr + nz) : hc;
  h[7] = (xp && zp) ? __ldcg(hc_ptr + nz + 1) : hc;

  // quad::_normal (cellpool.h:200-223) in closed form: plane (sx,sz) in {+,-}^2 counts iff
  // the diagonal (ix+sx, iz+sz) is in bounds; each plane is (-sx*S*dhx, 1, -sz*S*dhz).
  float nxv, nyv, nzv;
  {
    const float S = C.mapscale;
    const float cz = (zp ? 1.0f : 0.0f) + (zm ? 1.0f : 0.0f);
    const float cx = (xp ? 1.0f : 0.0f) + (xm ? 1.0f : 0.0f);
    const float a = xp ? (h[6] - hc) : 0.0f, b = xm ? (h[1] - hc) : 0.0f;
    const float c = zp ? (h[4] - hc) : 0.0f, d = zm ? (h[3] - hc) : 0.0f;
    nxv = S * cz * (b - a);
    nyv = cx * cz;
    nzv = S * cx * (d - c);
    const float l2 = nxv * nxv + nyv * nyv + nzv * nzv;
    if (l2 > 0.0f) { const float il = rsqrtf(l2); nxv *= il; nyv *= il; nzv *= il; }
  }

  if (q.age == 0 || q.py < hc) q.py = hc;                                         // :42-43
  if (q.age++ > C.max_age) return 2;                                              // :45-46

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
    // rspeed = n x (a x n) = a (n.n) - n (n.a),  a = (1-collision)*speed          :75
    const float k = 1.0f - collision;
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
    if (amp > 0.0f) {  // the draw is multiplied by amp: skip the generator when it is 0
      const uint4 r = philox4x32_10(make_uint4(pid, call, (uint32_t)q.age, 0x7u), key);
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

  red_add_v4(m.track + m.idx(ix, iz), q.sed, sx, sy, sz);                         // :93-96

  float c;  // centre height as this particle sees it
  {
    const float len = sqrtf(sx * sx + sy * sy + sz * sz);
    const float force = fmaxf(-(sx * nxv + sy * nyv + sz * nzv), 0.0f);           // :100-102
    const float lift = (1.0f - collision) * len;                                  // :104
    const float capacity = hfac * (force + 0.02f * lift);                         // :106
    const float dh = C.suspension * (capacity - q.sed);                           // :110-112
    c = hc - dh;
    q.sed += dh;
  }

  // World::cascade (world.hpp:79-153) on the register snapshot.
  {
    const unsigned valid = (xm && zm ? 1u : 0u) | (xm ? 2u : 0u) | (xm && zp ? 4u : 0u) |
                           (zm ? 8u : 0u) | (zp ? 16u : 0u) | (xp && zm ? 32u : 0u) |
                           (xp ? 64u : 0u) | (xp && zp ? 128u : 0u);
    float worst = -1.0f;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const float t = ((0xA5u >> k) & 1u) ? C.thr_diag : C.thr_ortho;
      const float e = fabsf(c - h[k]) - t;
      worst = fmaxf(worst, ((valid >> k) & 1u) ? e : -1.0f);
    }
    if (worst > 0.0f) {
      int s0 = 0, s1 = 1, s2 = 2, s3 = 3, s4 = 4, s5 = 5, s6 = 6, s7 = 7;
      // 19-comparator network for 8 keys, descending (world.hpp:118-120)
      cex(h[0], s0, h[1], s1); cex(h[2], s2, h[3], s3); cex(h[4], s4, h[5], s5); cex(h[6], s6, h[7], s7);
      cex(h[0], s0, h[2], s2); cex(h[1], s1, h[3], s3); cex(h[4], s4, h[6], s6); cex(h[5], s5, h[7], s7);
      cex(h[1], s1, h[2], s2); cex(h[5], s5, h[6], s6); cex(h[0], s0, h[4], s4); cex(h[3], s3, h[7], s7);
      cex(h[1], s1, h[5], s5); cex(h[2], s2, h[6], s6);
      cex(h[1], s1, h[4], s4); cex(h[3], s3, h[6], s6);
      cex(h[2], s2, h[4], s4); cex(h[3], s3, h[5], s5);
      cex(h[3], s3, h[4], s4);
      const int slots[8] = {s0, s1, s2, s3, s4, s5, s6, s7};
#pragma unroll
      for (int i = 0; i < 8; i++) {                                               // :122-151
        const int k = slots[i];
        const bool ok = (valid >> k) & 1u;
        const float t = ((0xA5u >> k) & 1u) ? C.thr_diag : C.thr_ortho;
        const float diff = c - h[i];
        const float excess = fabsf(diff) - t;
        if (ok && excess > 0.0f) {
          const float tr = copysignf(C.settling_half * excess, diff);
          c -= tr;
          const int m9 = k + (k >> 2);            // 0,1,2,3,5,6,7,8: position in the 3x3 block
          const int q3 = (m9 * 11) >> 5;          // m9 / 3
          const int off = (q3 - 1) * nz + (m9 - 3 * q3 - 1);
          red_add_f32(hc_ptr + off, tr);
        }
      }
    }
  }
  const float dc = c - hc;
  if (dc != 0.0f) red_add_f32(hc_ptr, dc);
  return 0;
}

__global__ void __launch_bounds__(256)
par_fly_kernel(MapView m, ParConst C, int n_particles, uint2 key, uint32_t call, DevStats* stats) {
  const uint32_t stride = gridDim.x * blockDim.x;
  uint32_t next = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned steps = 0, n_oob = 0, n_age = 0, n_spawn = 0;
  float lost = 0.0f;
  PState q;
  uint32_t pid = 0;
  bool alive = false;
  while (true) {
    if (!alive) {
      if (next >= (uint32_t)n_particles) break;
      pid = next; next += stride;
      // world.hpp:52-54: cell = (rand()%nx, rand()%nz), plus (rand()%1000)/1000 in each axis
      const uint4 r = philox4x32_10(make_uint4(pid, call, 0u, 0x5u), key);
      q.px = (float)__umulhi(r.x, (uint32_t)m.nx) + (float)__umulhi(r.z, 1000u) / 1000.0f;
      q.pz = (float)__umulhi(r.y, (uint32_t)m.nz) + (float)__umulhi(r.w, 1000u) / 1000.0f;
      q.py = 0.0f; q.sx = q.sy = q.sz = 0.0f; q.sed = 0.0f; q.age = 0;
      alive = true; n_spawn++;
    }
    const int r = par_step(m, C, q, pid, key, call);
    if (r == 0) { steps++; }
    else { alive = false; lost += q.sed; if (r == 1) n_oob++; else n_age++; }
  }
  // warp-aggregate the counters, one atomic per warp
  double lostd = (double)lost;
  for (int o = 16; o > 0; o >>= 1) {
    steps += __shfl_xor_sync(0xffffffffu, steps, o);
    n_oob += __shfl_xor_sync(0xffffffffu, n_oob, o);
    n_age += __shfl_xor_sync(0xffffffffu, n_age, o);
    n_spawn += __shfl_xor_sync(0xffffffffu, n_spawn, o);
    lostd += __shfl_xor_sync(0xffffffffu, lostd, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&stats->steps, (unsigned long long)steps);
    atomicAdd(&stats->exits_oob, (unsigned long long)n_oob);
    atomicAdd(&stats->exits_age, (unsigned long long)n_age);
    atomicAdd(&stats->particles, (unsigned long long)n_spawn);
    atomicAdd(&stats->sediment_lost, lostd);
  }
}

int launch_par_fly(const MapView& m, const swe_params& p, int n_particles, uint64_t seed,
                   uint32_t call_index, DevStats* stats, cudaStream_t st) {
  if (n_particles <= 0) return 0;
  ParConst C;
  C.inv_boundarylayer = 1.0f / p.boundarylayer;
  C.suspension = p.suspension; C.gravity = p.gravity;
  C.settling_half = 0.5f * p.settling;
  C.thr_ortho = p.maxdiff; C.thr_diag = sqrtf(2.0f) * p.maxdiff;
  C.mapscale = (float)p.mapscale;
  C.px = p.pspeed[0]; C.py = p.pspeed[1]; C.pz = p.pspeed[2];
  const float pl = sqrtf(C.px * C.px + C.py * C.py + C.pz * C.pz);
  C.pnx = C.px / pl; C.pny = C.py / pl; C.pnz = C.pz / pl;
  C.max_age = p.max_age;
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  const int block = 256;
  long long blocks = ((long long)n_particles + block - 1) / block;
  const long long cap = 148LL * 8;  // persistent: 8 CTAs of 256 threads per SM
  if (blocks > cap) blocks = cap;
  par_fly_kernel<<<(int)blocks, block, 0, st>>>(m, C, n_particles, key, call_index, stats);
  return 1;
}

}  // namespace swe
