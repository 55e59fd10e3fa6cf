// swe_block.cuh -- World::cascade (source/world.hpp:79-153) on a 3x3 block held in registers.
//
// Used by the colour-queue flight kernel (swe_queue.cu): a particle that owns its cell for the
// current micro-round holds the whole 3x3 neighbourhood in registers, so the slope relaxation the
// reference runs after every Wind::fly step (wind.hpp:114) happens in place, right after the
// erosion, with no second pass over memory.
//
// Same result as exact::cascade_core (bit for bit, including the libstdc++ insertion-sort tie
// order): the eight neighbours are sorted with a 19-comparator network under the strict total
// order (height descending, reference slot ascending) -- for distinct slots that is exactly what
// a stable descending sort of the slot-ordered list gives -- then the transfer chain of
// world.hpp:122-151 runs on the sorted registers. Everything is unrolled with constant indices
// (no local memory). The arithmetic goes through SWB_* so that the header also compiles as plain
// C++ (tests/test_block_cascade_cpu.py checks it against the oracle on the host; build that with
// -ffp-contract=off).
#pragma once

#if defined(__CUDACC__)
#define SWB_HD __host__ __device__ __forceinline__
#else
#define SWB_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define SWB_ADD(a, b) __fadd_rn((a), (b))
#define SWB_SUB(a, b) __fsub_rn((a), (b))
#define SWB_MUL(a, b) __fmul_rn((a), (b))
#define SWB_INF __int_as_float(0x7f800000)
#define SWB_F2U(f) __float_as_uint(f)
#define SWB_U2F(u) __uint_as_float(u)
#else
#include <cmath>
#include <cstring>
#define SWB_ADD(a, b) ((a) + (b))
#define SWB_SUB(a, b) ((a) - (b))
#define SWB_MUL(a, b) ((a) * (b))
#define SWB_INF INFINITY
static inline unsigned swb_f2u(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
static inline float swb_u2f(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
#define SWB_F2U(f) swb_f2u(f)
#define SWB_U2F(u) swb_u2f(u)
#endif

namespace swe {
namespace block {

struct Thr { float ortho, diag, settling; };

// reference offset table (world.hpp:83-92): slot k -> (dx, dz)
SWB_HD int slot_dx(int k) { return (k < 3) ? -1 : (k < 5 ? 0 : 1); }
SWB_HD int slot_dz(int k) { return (k == 3) ? -1 : (k == 4) ? 1 : ((k < 3 ? k : k - 5) - 1); }

// compare-exchange: afterwards (va, sa) sorts before (vb, sb)
SWB_HD void cswap(float& va, int& sa, float& ra, float& vb, int& sb, float& rb, bool layered) {
  const bool sw = (vb > va) || (vb == va && sb < sa);
  const float tv = sw ? vb : va; vb = sw ? va : vb; va = tv;
  const int ts = sw ? sb : sa; sb = sw ? sa : sb; sa = ts;
  if (layered) { const float tr = sw ? rb : ra; rb = sw ? ra : rb; ra = tr; }
}

// 64-bit sort key of a neighbour: the order-preserving integer image of its height (descending float order ==
// descending unsigned order; v + 0 folds -0 onto +0, which compare equal in the reference and behave identically in
// the transfer arithmetic) in the high bits, 15 - slot in the low four -- one integer maximum / minimum per
// comparator replaces three float compares and four selects (measured: the comparator was 200 of the kernel's
// 1100 instructions per particle step).
SWB_HD unsigned long long sort_key(float v, int slot) {
  const unsigned u = SWB_F2U(SWB_ADD(v, 0.0f));
  const unsigned k = u ^ ((unsigned)((int)u >> 31) | 0x80000000u);
  return ((unsigned long long)k << 4) | (unsigned long long)(15 - slot);
}
SWB_HD float key_height(unsigned long long key) {
  const unsigned k = (unsigned)(key >> 4);
  return SWB_U2F((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k);
}
SWB_HD void kswap(unsigned long long& a, unsigned long long& b) {
  const unsigned long long hi = a > b ? a : b, lo = a > b ? b : a;
  a = hi; b = lo;
}

// Sorted order of the eight neighbours (single-layer form): key[i] = the i-th neighbour in the reference's order
// (height descending, slot ascending among equal heights), packed by sort_key. `filler` stands in for neighbours
// outside the map: they never move (slot > 7 in the key: infinite gap below), so where they land in the order has
// no effect on the result and any value will do -- which makes the sort independent of the centre and lets a
// caller run it before (or interleaved with) the Wind::fly arithmetic that produces the centre's new height.
SWB_HD void sort_neighbours(const float (&h)[8], unsigned valid, float filler, unsigned long long (&key)[8]) {
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const bool ok = (valid >> k) & 1u;
    key[k] = sort_key(ok ? h[k] : filler, ok ? k : 8 + k);
  }
#define SWB_KS(i, j) kswap(key[i], key[j])
  SWB_KS(0, 2); SWB_KS(1, 3); SWB_KS(4, 6); SWB_KS(5, 7);
  SWB_KS(0, 4); SWB_KS(1, 5); SWB_KS(2, 6); SWB_KS(3, 7);
  SWB_KS(0, 1); SWB_KS(2, 3); SWB_KS(4, 5); SWB_KS(6, 7);
  SWB_KS(2, 4); SWB_KS(3, 5);
  SWB_KS(1, 4); SWB_KS(3, 6);
  SWB_KS(1, 2); SWB_KS(3, 4); SWB_KS(5, 6);
#undef SWB_KS
}

// The transfer chain of world.hpp:122-151 on neighbours already in sorted order (v[i] / s[i] / r[i] = height, slot,
// bedrock of the i-th). Returns the moved mask; c is updated.
template <bool LAYERED>
SWB_HD unsigned transfer_chain(float& c, float (&v)[8], const int (&s)[8], const float (&r)[8], const Thr T, float rock_c) {
  unsigned moved = 0u;
  float cc = c;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const float t = s[i] > 7 ? SWB_INF : (((0xA5u >> s[i]) & 1u) ? T.diag : T.ortho);   // slots {0,2,5,7} are diagonal
    const float diff = SWB_SUB(cc, v[i]);                          // world.hpp:127
    const float excess = SWB_SUB(fabsf(diff), t);                  // :133
    bool go = (diff != 0.0f) && (excess > 0.0f);                   // :128, :135
    float transfer = SWB_MUL(SWB_MUL(T.settling, excess), 0.5f);   // :139 (x/2 == x*0.5 exactly)
    if (LAYERED) {                                                 // extension: only sand moves
      const float sand = (diff > 0.0f) ? SWB_SUB(cc, rock_c) : SWB_SUB(v[i], r[i]);
      transfer = fminf(transfer, fmaxf(sand, 0.0f));
      go = go && (transfer > 0.0f);
    }
    const float sgn = (diff > 0.0f) ? transfer : -transfer;        // :142-149
    if (go) { cc = SWB_SUB(cc, sgn); v[i] = SWB_ADD(v[i], sgn); moved |= 1u << i; }
  }
  c = cc;
  return moved;
}

// In:  c = centre height, h[k] = height of neighbour slot k, valid bit k = neighbour inside the map
//      (closed boundary, cellpool.h:432-438); layered: rock_c / rk[k] = bedrock heights (extension).
// Out: c updated; v[i] / s[i] = height and slot of the i-th neighbour in sorted order, moved bit i =
//      its transfer ran (its cell must be written back). Returns the moved mask (0 = nothing changed,
//      the centre changed iff it is non-zero).
template <bool LAYERED>
SWB_HD unsigned cascade_sorted(float& c, const float (&h)[8], unsigned valid, const Thr T, float (&v)[8], int (&s)[8],
                               float rock_c, const float (&rk)[8]) {
  float r[8];
  if (LAYERED) {
    // (the bedrock heights travel with their neighbours: tuple compare-exchange)
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const bool ok = (valid >> k) & 1u;
      v[k] = ok ? h[k] : c;          // an absent neighbour: never moves (its gap is infinite, see below)
      s[k] = ok ? k : 8 + k;
      r[k] = rk[k];
    }
#define SWB_CS(i, j) cswap(v[i], s[i], r[i], v[j], s[j], r[j], true)
    SWB_CS(0, 2); SWB_CS(1, 3); SWB_CS(4, 6); SWB_CS(5, 7);
    SWB_CS(0, 4); SWB_CS(1, 5); SWB_CS(2, 6); SWB_CS(3, 7);
    SWB_CS(0, 1); SWB_CS(2, 3); SWB_CS(4, 5); SWB_CS(6, 7);
    SWB_CS(2, 4); SWB_CS(3, 5);
    SWB_CS(1, 4); SWB_CS(3, 6);
    SWB_CS(1, 2); SWB_CS(3, 4); SWB_CS(5, 6);
#undef SWB_CS
  } else {
    unsigned long long key[8];
    sort_neighbours(h, valid, c, key);
#pragma unroll
    for (int k = 0; k < 8; k++) { v[k] = key_height(key[k]); s[k] = 15 - (int)(key[k] & 15ull); r[k] = 0.0f; }
  }
  return transfer_chain<LAYERED>(c, v, s, r, T, rock_c);
}

}  // namespace block
}  // namespace swe
