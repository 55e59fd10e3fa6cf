// swe_exact.cuh -- bit-exact device restatement of the reference arithmetic.
//
// Everything here must round exactly like the reference built with g++ -O2 (x86-64 SSE,
// no FMA contraction, reference Makefile:3-4): IEEE fp32 add/mul/div/sqrt, one double
// expression (wind.hpp:87) and glibc's expf (wind.hpp:50). Translation units including
// this header are compiled with -fmad=false -prec-div=true -prec-sqrt=true; the helpers
// additionally use the _rn intrinsics, which nvcc never contracts, wherever a product
// feeds a sum.
//
// glm semantics (third-party, un-vendored, scalar code paths of glm 0.9.9):
//   dot = (x*x' + y*y') + z*z'; length = sqrt(dot); normalize = v * (1/sqrt(dot));
//   cross = (a.y*b.z - b.y*a.z, a.z*b.x - b.z*a.x, a.x*b.y - b.x*a.y); mix = x*(1-a) + y*a.
#pragma once
#include "swe_common.cuh"

namespace swe {
namespace exact {

struct V3 { float x, y, z; };

__device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float fsqrt(float a) { return __fsqrt_rn(a); }

__device__ __forceinline__ V3 add(V3 a, V3 b) { return v3(fadd(a.x, b.x), fadd(a.y, b.y), fadd(a.z, b.z)); }
__device__ __forceinline__ V3 sub(V3 a, V3 b) { return v3(fsub(a.x, b.x), fsub(a.y, b.y), fsub(a.z, b.z)); }
__device__ __forceinline__ V3 mul(V3 a, V3 b) { return v3(fmul(a.x, b.x), fmul(a.y, b.y), fmul(a.z, b.z)); }
__device__ __forceinline__ V3 scale(float s, V3 a) { return v3(fmul(s, a.x), fmul(s, a.y), fmul(s, a.z)); }
__device__ __forceinline__ float dot(V3 a, V3 b) {
  return fadd(fadd(fmul(a.x, b.x), fmul(a.y, b.y)), fmul(a.z, b.z));
}
__device__ __forceinline__ float length(V3 a) { return fsqrt(dot(a, a)); }
__device__ __forceinline__ V3 normalize(V3 a) { return scale(fdiv(1.0f, fsqrt(dot(a, a))), a); }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
  return v3(fsub(fmul(a.y, b.z), fmul(b.y, a.z)),
            fsub(fmul(a.z, b.x), fmul(b.z, a.x)),
            fsub(fmul(a.x, b.y), fmul(b.x, a.y)));
}
__device__ __forceinline__ V3 mix(V3 x, V3 y, float a) { return add(scale(fsub(1.0f, a), x), scale(a, y)); }

// ---- glibc 2.39 expf (sysdeps/ieee754/flt-32/e_expf.c, e_exp2f_data.c: N = 32) ----------
// N = 32 table, degree-3 polynomial evaluated in double. The algorithm was compared with
// glibc 2.39's libm for all 2^32 inputs on the CPU: identical except two inputs, patched
// below. tests/test_gpu_parity.py checks the device version against the host libm.
static __device__ __constant__ unsigned long long c_exp2f_tab[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull};

static __device__ __noinline__ float expf_glibc(float x) {
  const double InvLn2N = 0x1.71547652b82fep+0 * 32;
  const double Shift = 0x1.8p+52;
  const double C0 = 0x1.c6af84b912394p-5 / 32 / 32 / 32;
  const double C1 = 0x1.ebfce50fac4f3p-3 / 32 / 32;
  const double C2 = 0x1.62e42ff0c52d6p-1 / 32;
  const unsigned ux = __float_as_uint(x), abstop = (ux >> 20) & 0x7ffu;
  if (abstop >= (0x42b00000u >> 20)) {
    if (ux == 0xff800000u) return 0.0f;
    if (abstop >= (0x7f800000u >> 20)) return __fadd_rn(x, x);
    if (x > 0x1.62e42ep6f) return __uint_as_float(0x7f800000u);
    if (x < -0x1.9fe368p6f) return 0.0f;
  }
  if (ux == 0x4202422fu) return __uint_as_float(0x56fc9f1cu);
  if (ux == 0xc27c65d9u) return __uint_as_float(0x11fa2993u);
  double z = __dmul_rn(InvLn2N, (double)x);
  double kd = __dadd_rn(z, Shift);
  unsigned long long ki = (unsigned long long)__double_as_longlong(kd);
  kd = __dsub_rn(kd, Shift);
  double r = __dsub_rn(z, kd);
  unsigned long long t = c_exp2f_tab[ki & 31ull] + (ki << (52 - 5));
  double s = __longlong_as_double((long long)t);
  double zz = __dadd_rn(__dmul_rn(C0, r), C1);
  double r2 = __dmul_rn(r, r);
  double y = __dadd_rn(__dmul_rn(C2, r), 1.0);
  y = __dadd_rn(__dmul_rn(zz, r2), y);
  y = __dmul_rn(y, s);
  return __double2float_rn(y);
}

// ---- quad::_normal, cellpool.h:200-223 -------------------------------------------------
// Four triangle normals in the reference's order; a triangle counts iff its DIAGONAL
// neighbour is in bounds. Restated plane by plane (signed zeros included), not in closed
// form, so the sum rounds exactly like the reference.
__device__ __forceinline__ V3 normal(const MapView& m, const swe_params& p, int x, int z) {
  V3 n = v3(0.0f, 0.0f, 0.0f);
  const V3 s = v3(1.0f, (float)p.mapscale, 1.0f);
  const float h = m.height[m.idx(x, z)];
  if (!m.oob(x + 1, z + 1))
    n = add(n, cross(mul(s, v3(0.0f, fsub(m.height[m.idx(x, z + 1)], h), 1.0f)),
                     mul(s, v3(1.0f, fsub(m.height[m.idx(x + 1, z)], h), 0.0f))));
  if (!m.oob(x - 1, z - 1))
    n = add(n, cross(mul(s, v3(0.0f, fsub(m.height[m.idx(x, z - 1)], h), -1.0f)),
                     mul(s, v3(-1.0f, fsub(m.height[m.idx(x - 1, z)], h), 0.0f))));
  if (!m.oob(x + 1, z - 1))
    n = add(n, cross(mul(s, v3(1.0f, fsub(m.height[m.idx(x + 1, z)], h), 0.0f)),
                     mul(s, v3(0.0f, fsub(m.height[m.idx(x, z - 1)], h), -1.0f))));
  if (!m.oob(x - 1, z + 1))
    n = add(n, cross(mul(s, v3(-1.0f, fsub(m.height[m.idx(x - 1, z)], h), 0.0f)),
                     mul(s, v3(0.0f, fsub(m.height[m.idx(x, z + 1)], h), 1.0f))));
  if (length(n) > 0) n = normalize(n);
  return n;
}

// ---- World::cascade core, world.hpp:105-151 ---------------------------------------------
// Given the centre height and the gathered in-bounds neighbours (reference offset order,
// world.hpp:83-92), sort them like libstdc++'s std::sort does for n <= 16 (plain insertion
// sort with comp a.h > b.h: stable, ties keep offset order) and apply the pairwise
// transfers. `nh[k]` is updated in place (k = gather slot), returns the new centre height.
//   thr_o = 1*maxdiff*lodsize, thr_d = sqrtf(2)*maxdiff*lodsize (lodsize = 1).
struct CascadeThr { float ortho, diag; };
__device__ __forceinline__ CascadeThr cascade_thresholds(const swe_params& p) {
  CascadeThr t;
  // length(vec2(nn)) = sqrt(dx*dx + dz*dz): sqrt(1) = 1, sqrt(2) correctly rounded
  t.ortho = fmul(fmul(fsqrt(1.0f), p.maxdiff), 1.0f);
  t.diag = fmul(fmul(fsqrt(2.0f), p.maxdiff), 1.0f);
  return t;
}

// slot k in 0..7 of the reference offset table; diagonal iff both dx,dz non-zero.
__device__ __forceinline__ bool slot_is_diag(int k) { return (0xA5u >> k) & 1u; }  // {0,2,5,7}
__device__ __forceinline__ int slot_dx(int k) { return (k < 3) ? -1 : (k < 5 ? 0 : 1); }
__device__ __forceinline__ int slot_dz(int k) {
  // {-1,0,1, -1,1, -1,0,1}
  return (k == 3) ? -1 : (k == 4) ? 1 : ((k < 3 ? k : k - 5) - 1);
}

// `num` gathered neighbours: h[i] snapshot height, slot[i] offset-table slot (ascending).
// Writes the new neighbour heights back into h[] (same positions as gathered, i.e. slot
// order) and returns the new centre height. `changed` gets a bitmask over gather positions.
// Layered extension (not in the reference snapshot; our specification): with bedrock heights
// rock_c / rk[] given, only sand moves -- every transfer is capped by the giving cell's sand
// max(height - rock, 0). rk == nullptr is the reference behaviour, bit for bit.
__device__ __forceinline__ float cascade_core(float centre, float* h, const int* slot, int num,
                                              const CascadeThr thr, float settling,
                                              unsigned* changed, float rock_c = 0.0f,
                                              const float* rk = nullptr) {
  // insertion sort of indices (libstdc++ __insertion_sort semantics incl. the
  // compare-with-first shortcut, which only matters for NaN heights)
  int ord[8];
  for (int i = 0; i < num; i++) ord[i] = i;
  for (int i = 1; i < num; i++) {
    const int v = ord[i];
    const float vh = h[v];
    if (vh > h[ord[0]]) {
      for (int j = i; j > 0; j--) ord[j] = ord[j - 1];
      ord[0] = v;
    } else {
      int j = i;
      while (vh > h[ord[j - 1]]) { ord[j] = ord[j - 1]; j--; }
      ord[j] = v;
    }
  }
  unsigned ch = 0;
  for (int i = 0; i < num; i++) {
    const int k = ord[i];
    const float diff = fsub(centre, h[k]);                          // world.hpp:127
    if (diff == 0) continue;
    const float t = slot_is_diag(slot[k]) ? thr.diag : thr.ortho;
    const float excess = fsub(fabsf(diff), t);                      // :133
    if (excess <= 0) continue;
    float transfer = fdiv(fmul(settling, excess), 2.0f);            // :139
    if (rk) {
      const float sand = (diff > 0) ? fsub(centre, rock_c) : fsub(h[k], rk[k]);
      transfer = fminf(transfer, fmaxf(sand, 0.0f));
      if (!(transfer > 0.0f)) continue;
    }
    if (diff > 0) { centre = fsub(centre, transfer); h[k] = fadd(h[k], transfer); }
    else          { centre = fadd(centre, transfer); h[k] = fsub(h[k], transfer); }
    ch |= 1u << k;
  }
  *changed = ch;
  return centre;
}

// World::cascade(pos) on global memory (single cell): gather, core, scatter.
__device__ __forceinline__ void cascade_global(const MapView& m, const swe_params& p,
                                               const CascadeThr thr, int x, int z) {
  float h[8], rk[8]; int slot[8]; int num = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const int nx_ = x + slot_dx(k), nz_ = z + slot_dz(k);
    if (m.oob(nx_, nz_)) continue;
    h[num] = m.height[m.idx(nx_, nz_)];
    rk[num] = m.rock ? m.rock[m.idx(nx_, nz_)] : 0.0f;
    slot[num] = k;
    num++;
  }
  unsigned changed;
  const size_t ci = m.idx(x, z);
  const float c0 = m.height[ci];
  const float c1 = cascade_core(c0, h, slot, num, thr, p.settling, &changed, m.rock ? m.rock[ci] : 0.0f,
                                m.rock ? rk : nullptr);
  if (changed) {
    m.height[ci] = c1;
    for (int i = 0; i < num; i++)
      if (changed & (1u << i)) m.height[m.idx(x + slot_dx(slot[i]), z + slot_dz(slot[i]))] = h[i];
  }
}

}  // namespace exact
}  // namespace swe
