// swe_cascade.cu -- whole-map cascade (slope relaxation to the repose angle).
//
// The reference only has the local World::cascade(pos) (source/world.hpp:79-153), an
// in-place Gauss-Seidel update of one cell's 3x3 block. A whole-map sweep applies it to
// every cell in the conflict-free 3x3 colour order: colour c = 0..8, cx = c/3, cz = c%3,
// centres {x%3 == cx, z%3 == cz}. Two centres of one colour differ by >= 3 in x or z, so
// their 3x3 footprints are disjoint and the colour class can run fully in parallel while
// staying BITWISE identical to calling the reference function in that order
// (oracle: ref_cascade_sweep / swo_cascade_sweep).
//
// Kernels:
//   cascade_colour_kernel   -- v1: one launch per colour, straight from global memory
//   cascade_tile_kernel     -- fused: all 9 colours of a sweep on a shared-memory tile with
//                              a dependency halo ("cascade sweep")
//   max_excess_kernel       -- convergence metric
//
// Compiled with -fmad=false; arithmetic through swe_exact.cuh.
#include <cuda.h>
#include <cstring>
#include <mutex>
#include <set>
#include <utility>

#include "swe_exact.cuh"

namespace swe {
using namespace exact;

__global__ void cascade_colour_kernel(MapView m, swe_params P, int cx, int cz, int ncx, int ncz) {
  const CascadeThr thr = cascade_thresholds(P);
  const long long total = (long long)ncx * ncz;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = cx + 3 * (int)(i / ncz);
    const int z = cz + 3 * (int)(i % ncz);
    cascade_global(m, P, thr, x, z);
  }
}

// =============================================================================================
// Fused sweep: all nine colours of one sweep on a shared-memory tile.
//
// Output tile O = TX x TZ cells whose origin is == 1 (mod 3) in x and z (tile grid origin at
// -2), so the 3x3 blocks of the LAST colour (cx = cz = 2: blocks start at == 1 mod 3) tile O
// exactly. Walking the colours backwards, the cells needed before colour c form the box
//     R_c = O grown by  x: [2-cx_c, 2*(2-cx_c)]   z: [8-c, 2*(8-c)]        (cx_c = c/3)
// (each change of the colour's shift adds one cell on the low side and two on the high side),
// R_c is a union of whole colour-c blocks, and the centres to run at colour c are exactly the
// colour-c centres inside R_c. The tile therefore loads R_0 = O + x:[2,4] z:[8,16], runs the
// nine colour phases with a barrier between them, and stores O. Redundant halo work:
// (TX+6)(TZ+24)/(TX*TZ) = 1.22 for 60 x 216.
//
// Thread layout: blockDim = (TZ/3 + 8, 6): threadIdx.x is the centre's z slot (stride 3 in
// shared memory: bank-conflict free), threadIdx.y strides over the centre rows.
//
// Per centre: (1) optional dirty-stamp test (flight pipeline: only cells a particle touched
// this round cascade); (2) exact early-out: if no in-map neighbour has |c-h|-t > 0 the
// reference's loop transfers nothing; (3) slow path: neighbours sorted descending and stable
// through a 19-comparator network on order-preserving integer keys whose low 4 bits carry
// the neighbour's (dx,dz) code in reference slot order, then the sequential transfer chain.
// Keys that differ only in the 4 bits given up to the code cannot be ordered this way; those
// (rare) centres and NaN-free-ness are delegated to the generic exact core.
// =============================================================================================
constexpr int TILE_TX = 60, TILE_TZ = 216;                  // TZ: multiple of 3 and of 4
constexpr int TILE_LX = TILE_TX + 6, TILE_PITCH = TILE_TZ + 24;
constexpr int TILE_BX = TILE_TZ / 3 + 8, TILE_BY = 6;       // dense scan: 80 x 6 centre slots (threads 0..479)
constexpr int TILE_THREADS = 512;                           // 16 warps: the pair scan wants 8 column strips x 2
constexpr int TILE_OX = -2, TILE_OZ = -8;                   // tile grid origin: == 1 (mod 3); z also == 0 (mod 4)
constexpr int TILE_FLOAT_BYTES = TILE_LX * TILE_PITCH * 4;  // 63,360 B   height tile
constexpr int TILE_QUEUE_BYTES = ((TILE_LX / 3) * (TILE_PITCH / 3) * 2 + 15) / 16 * 16;  // centres of one colour
// pair-bitmap mode (see tile_pair_scan): a second queue and two bit planes (one bit per tile cell)
// (bit planes are COLUMN-major: word [col][row >> 5], bit row & 31 -- a lane of the pair scan owns a column)
constexpr int MASK_WORDS = (TILE_LX + 31) / 32;             // 3 words per tile column
constexpr int TILE_MASK_BYTES = TILE_PITCH * MASK_WORDS * 4;   // 2,880 B per bit plane
constexpr int TILE_SMEM_PLAIN = TILE_FLOAT_BYTES + 2 * TILE_QUEUE_BYTES + 2 * TILE_MASK_BYTES + 256;  // + counters, mbarrier, chain table
static_assert(TILE_PITCH == 8 * 30, "tile_pair_scan walks the tile in 8 groups of 30 columns");

__device__ __forceinline__ void cswap_desc(unsigned& a, unsigned& b) {
  const unsigned hi = max(a, b), lo = min(a, b);
  a = hi; b = lo;
}

__device__ __forceinline__ unsigned smem_u32(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}

// generic exact fallback on the tile (same routine the single-cell path uses)
// rock_c_ptr != nullptr: layered extension, rock_c_ptr points at the centre's bedrock height in
// the global rock plane (row pitch nz); neighbours outside local storage are never gathered
// because they are outside the map.
__device__ __noinline__ bool tile_cascade_generic(float* tile_c, int gx, int gz, int nx, int nz,
                                                  CascadeThr thr, float settling,
                                                  const float* rock_c_ptr = nullptr) {
  float h[8], rk[8]; int slot[8]; int num = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const int dx = slot_dx(k), dz = slot_dz(k);
    const int x_ = gx + dx, z_ = gz + dz;
    if (x_ < 0 || z_ < 0 || x_ >= nx || z_ >= nz) continue;
    h[num] = tile_c[dx * TILE_PITCH + dz];
    rk[num] = rock_c_ptr ? __ldg(rock_c_ptr + dx * nz + dz) : 0.0f;
    slot[num] = k;
    num++;
  }
  unsigned changed;
  const float c1 = cascade_core(tile_c[0], h, slot, num, thr, settling, &changed,
                                rock_c_ptr ? __ldg(rock_c_ptr) : 0.0f, rock_c_ptr ? rk : nullptr);
  if (!changed) return false;
  tile_c[0] = c1;
  for (int i = 0; i < num; i++)
    if (changed & (1u << i)) tile_c[slot_dx(slot[i]) * TILE_PITCH + slot_dz(slot[i])] = h[i];
  return true;
}

// ---- per-centre pieces ----------------------------------------------------------------------
// INTERIOR tiles lie wholly inside the map (all eight neighbours of every centre exist), so the
// validity predicates vanish at compile time; border tiles use the same code with them.
struct Valid8 { bool v0, v1, v2, v3, v4, v5, v6, v7; };

template <bool INTERIOR>
__device__ __forceinline__ Valid8 centre_valid(int gx, int gz, int nx, int nz) {
  Valid8 v;
  if (INTERIOR) { v.v0 = v.v1 = v.v2 = v.v3 = v.v4 = v.v5 = v.v6 = v.v7 = true; return v; }
  const bool xm = gx > 0, xp = gx + 1 < nx, zm = gz > 0, zp = gz + 1 < nz;
  v.v0 = xm && zm; v.v1 = xm; v.v2 = xm && zp; v.v3 = zm; v.v4 = zp; v.v5 = xp && zm; v.v6 = xp; v.v7 = xp && zp;
  return v;
}

// Exact early-out tests (thresholds >= 0). World::cascade leaves the 3x3 block untouched iff no
// transfer of its loop (world.hpp:122-151) changes a value; as long as nothing has changed the
// centre keeps its initial height, so the test can be made against the initial centre:
//   coarse : some in-map neighbour has excess = |c-h| - t > 0        (cheap, a superset)
//   refined: ... and its transfer (settling*excess)/2 actually changes c or h in fp32.
// On a map relaxed to the repose angle most pairs sit within an ulp of the gap: ~87 % of the
// centres pass the coarse test but only a few per cent the refined one (measured, 4096^2).
template <bool REFINED>
__device__ __forceinline__ bool pair_moves(float cen, float h, float t, float half_settling) {
  const float diff = fsub(cen, h);
  const float excess = fsub(fabsf(diff), t);
  if (!REFINED) return !(excess <= 0.0f);
  const float transfer = fmul(excess, half_settling);           // == (settling*excess)/2.0f
  const float sgn = (diff > 0.0f) ? transfer : -transfer;
  return !(excess <= 0.0f) && ((fsub(cen, sgn) != cen) || (fadd(h, sgn) != h));
}

template <bool INTERIOR, bool REFINED>
__device__ __forceinline__ bool centre_needs(const float* tc, const CascadeThr thr, float settling,
                                             const Valid8 v) {   // settling: already halved
  const float cen = tc[0];
  const bool m0 = pair_moves<REFINED>(cen, tc[-TILE_PITCH - 1], thr.diag, settling);
  const bool m1 = pair_moves<REFINED>(cen, tc[-TILE_PITCH], thr.ortho, settling);
  const bool m2 = pair_moves<REFINED>(cen, tc[-TILE_PITCH + 1], thr.diag, settling);
  const bool m3 = pair_moves<REFINED>(cen, tc[-1], thr.ortho, settling);
  const bool m4 = pair_moves<REFINED>(cen, tc[1], thr.ortho, settling);
  const bool m5 = pair_moves<REFINED>(cen, tc[TILE_PITCH - 1], thr.diag, settling);
  const bool m6 = pair_moves<REFINED>(cen, tc[TILE_PITCH], thr.ortho, settling);
  const bool m7 = pair_moves<REFINED>(cen, tc[TILE_PITCH + 1], thr.diag, settling);
  if (INTERIOR) return m0 | m1 | m2 | m3 | m4 | m5 | m6 | m7;
  return (v.v0 & m0) | (v.v1 & m1) | (v.v2 & m2) | (v.v3 & m3) | (v.v4 & m4) | (v.v5 & m5) | (v.v6 & m6) |
         (v.v7 & m7);
}

// order-preserving key of a neighbour height: descending float order == descending unsigned
// order; (h + 0) folds -0 onto +0 (they tie in the reference's comparison); the low 4 bits carry
// 15 - code, code = (dx+1)*4 + (dz+1) = reference slot order, so equal heights keep slot order
// (std::sort on <= 16 elements is a stable insertion sort) and the chain can find the cell.
__device__ __forceinline__ unsigned height_key(float h, bool valid, unsigned code) {
  const unsigned u = __float_as_uint(fadd(h, 0.0f));
  const unsigned k = u ^ ((unsigned)((int)u >> 31) | 0x80000000u);
  return valid ? ((k & ~15u) | (15u - code)) : 0u;
}

// Neighbour table of the transfer chain, in shared memory: entry j = key & 15 = 15 - code gives the
// neighbour's byte offset from the centre in the tile and its repose gap (ortho or diag). Entry 0 is
// the empty slot of a centre at the map edge: it points at the centre itself with an infinite gap, so
// it never moves anything and never reads a cell another thread may be writing.
__device__ __forceinline__ void chain_table_init(float2* lut, int j, const CascadeThr thr) {
  const int code = 15 - j, dxm = (code >> 2) - 1, dzm = (code & 3) - 1;
  const bool real = j != 0 && dxm <= 1 && dzm <= 1 && !(dxm == 0 && dzm == 0);
  lut[j] = real ? make_float2(__int_as_float((dxm * TILE_PITCH + dzm) * 4), (dxm != 0 && dzm != 0) ? thr.diag : thr.ortho)
                : make_float2(__int_as_float(0), INFINITY);
}

// one element of the transfer chain (world.hpp:122-151), branch-free; thresholds >= 0 so that
// diff == 0 implies excess <= 0 and the reference's explicit diff == 0 test is subsumed.
// LAYERED (extension): the transfer is capped by the giving cell's sand, height - rock.
// TRACK: `moved` collects bit `code` of every neighbour whose transfer ran (pair-bitmap mode).
template <bool LAYERED, bool TRACK = false>
__device__ __forceinline__ void chain_step(float* tc, unsigned k, float& cc, const float2* lut,
                                           float settling, const float* rock_c_ptr, int nz, float rock_c,
                                           unsigned* moved = nullptr) {
  const float2 e = lut[k & 15u];
  float* hp = reinterpret_cast<float*>(reinterpret_cast<char*>(tc) + __float_as_int(e.x));
  const float hv = *hp;
  const float diff = fsub(cc, hv);
  const float excess = fsub(fabsf(diff), e.y);
  bool go = !(excess <= 0.0f);
  // (settling*excess)/2.0f == excess*(settling/2): halving is exact and commutes with the
  // rounding of the product (sub-normal products excluded: tiny thresholds use the generic core)
  float transfer = fmul(excess, settling);                       // settling: already halved
  if (LAYERED) {
    const unsigned code = 15u - (k & 15u);
    const int dxm = (int)(code >> 2) - 1, dzm = (int)(code & 3u) - 1;
    const float rk = (k != 0u) ? __ldg(rock_c_ptr + dxm * nz + dzm) : 0.0f;
    const float sand = (diff > 0.0f) ? fsub(cc, rock_c) : fsub(hv, rk);
    transfer = fminf(transfer, fmaxf(sand, 0.0f));
    go = go && (transfer > 0.0f);
  }
  const float sgn = (diff > 0.0f) ? transfer : -transfer;        // diff > 0: centre gives, else takes
  if (go) { cc = fsub(cc, sgn); *hp = fadd(hv, sgn); }
  if (TRACK && go) *moved |= 0x8000u >> (k & 15u);               // 1 << code
}

// the same chain element on register values: returns whether the transfer ran
__device__ __forceinline__ bool chain_reg(float& cc, float& hv, float t, float half_settling) {
  const float diff = fsub(cc, hv);
  const float excess = fsub(fabsf(diff), t);
  const bool go = !(excess <= 0.0f);
  const float transfer = fmul(excess, half_settling);
  const float sgn = (diff > 0.0f) ? transfer : -transfer;        // diff > 0: centre gives, else takes
  if (go) { cc = fsub(cc, sgn); hv = fadd(hv, sgn); }
  return go;
}

// World::cascade at one centre of the tile whose early-out test failed.
// TRACK: *moved gets one bit per 3x3 cell whose value may have changed, bit (dx+1)*4 + (dz+1).
template <bool INTERIOR, bool LAYERED = false, bool TRACK = false>
__device__ __forceinline__ void centre_relax(float* tc, const CascadeThr thr, const float2* lut, float settling, float hs,
                                             const Valid8 v, int gx, int gz, int nx, int nz,
                                             const float* rock_c_ptr = nullptr, unsigned* moved = nullptr) {
  unsigned k0 = height_key(tc[-TILE_PITCH - 1], v.v0, 0), k1 = height_key(tc[-TILE_PITCH], v.v1, 1),
           k2 = height_key(tc[-TILE_PITCH + 1], v.v2, 2), k3 = height_key(tc[-1], v.v3, 4),
           k4 = height_key(tc[1], v.v4, 6), k5 = height_key(tc[TILE_PITCH - 1], v.v5, 8),
           k6 = height_key(tc[TILE_PITCH], v.v6, 9), k7 = height_key(tc[TILE_PITCH + 1], v.v7, 10);
  // 19-comparator network, descending (world.hpp:118-120)
  cswap_desc(k0, k2); cswap_desc(k1, k3); cswap_desc(k4, k6); cswap_desc(k5, k7);
  cswap_desc(k0, k4); cswap_desc(k1, k5); cswap_desc(k2, k6); cswap_desc(k3, k7);
  cswap_desc(k0, k1); cswap_desc(k2, k3); cswap_desc(k4, k5); cswap_desc(k6, k7);
  cswap_desc(k2, k4); cswap_desc(k3, k5);
  cswap_desc(k1, k4); cswap_desc(k3, k6);
  cswap_desc(k1, k2); cswap_desc(k3, k4); cswap_desc(k5, k6);
  // heights that agree in everything but the 4 bits given up to the code cannot be ordered by
  // the keys: hand those (rare) centres to the generic exact core
  unsigned close = min(min(k0 ^ k1, (k2 ? k1 ^ k2 : ~0u)), min((k3 ? k2 ^ k3 : ~0u), (k4 ? k3 ^ k4 : ~0u)));
  close = min(close, min(min((k5 ? k4 ^ k5 : ~0u), (k6 ? k5 ^ k6 : ~0u)), (k7 ? k6 ^ k7 : ~0u)));
  if ((INTERIOR || k1) && close < 16u) {
    const bool ch = tile_cascade_generic(tc, gx, gz, nx, nz, thr, settling, LAYERED ? rock_c_ptr : nullptr);
    if (TRACK) *moved = ch ? 0x777u : 0u;
    return;
  }
  if (!LAYERED) {
    // all eight neighbours are fetched up front (sorted order), the chain runs in registers and the
    // changed cells are stored afterwards: no store -> load round trip through shared memory between
    // the dependent steps (the eight cells are distinct, which the compiler cannot know)
    const float2 e0 = lut[k0 & 15u], e1 = lut[k1 & 15u], e2 = lut[k2 & 15u], e3 = lut[k3 & 15u],
                 e4 = lut[k4 & 15u], e5 = lut[k5 & 15u], e6 = lut[k6 & 15u], e7 = lut[k7 & 15u];
    char* base = reinterpret_cast<char*>(tc);
    float* p0 = reinterpret_cast<float*>(base + __float_as_int(e0.x)); float* p1 = reinterpret_cast<float*>(base + __float_as_int(e1.x));
    float* p2 = reinterpret_cast<float*>(base + __float_as_int(e2.x)); float* p3 = reinterpret_cast<float*>(base + __float_as_int(e3.x));
    float* p4 = reinterpret_cast<float*>(base + __float_as_int(e4.x)); float* p5 = reinterpret_cast<float*>(base + __float_as_int(e5.x));
    float* p6 = reinterpret_cast<float*>(base + __float_as_int(e6.x)); float* p7 = reinterpret_cast<float*>(base + __float_as_int(e7.x));
    float v0 = *p0, v1 = *p1, v2 = *p2, v3 = *p3, v4 = *p4, v5 = *p5, v6 = *p6, v7 = *p7;
    float cc = tc[0];
    const bool g0 = chain_reg(cc, v0, e0.y, hs), g1 = chain_reg(cc, v1, e1.y, hs), g2 = chain_reg(cc, v2, e2.y, hs),
               g3 = chain_reg(cc, v3, e3.y, hs), g4 = chain_reg(cc, v4, e4.y, hs), g5 = chain_reg(cc, v5, e5.y, hs),
               g6 = chain_reg(cc, v6, e6.y, hs), g7 = chain_reg(cc, v7, e7.y, hs);
    if (g0) *p0 = v0; if (g1) *p1 = v1; if (g2) *p2 = v2; if (g3) *p3 = v3;
    if (g4) *p4 = v4; if (g5) *p5 = v5; if (g6) *p6 = v6; if (g7) *p7 = v7;
    tc[0] = cc;                                        // after the neighbour stores: an empty slot aliases the centre
    if (TRACK) {
      unsigned mv = 0u;
      if (g0) mv |= 0x8000u >> (k0 & 15u); if (g1) mv |= 0x8000u >> (k1 & 15u); if (g2) mv |= 0x8000u >> (k2 & 15u);
      if (g3) mv |= 0x8000u >> (k3 & 15u); if (g4) mv |= 0x8000u >> (k4 & 15u); if (g5) mv |= 0x8000u >> (k5 & 15u);
      if (g6) mv |= 0x8000u >> (k6 & 15u); if (g7) mv |= 0x8000u >> (k7 & 15u);
      *moved = mv ? (mv | (1u << 5)) : 0u;             // the centre (code 5) moved iff any transfer ran
    }
    return;
  }
  float cc = tc[0];
  const float rock_c = LAYERED ? __ldg(rock_c_ptr) : 0.0f;
  const int rnz = nz;
  unsigned mv = 0u;
  chain_step<LAYERED, TRACK>(tc, k0, cc, lut, hs, rock_c_ptr, rnz, rock_c, &mv); chain_step<LAYERED, TRACK>(tc, k1, cc, lut, hs, rock_c_ptr, rnz, rock_c, &mv);
  chain_step<LAYERED, TRACK>(tc, k2, cc, lut, hs, rock_c_ptr, rnz, rock_c, &mv); chain_step<LAYERED, TRACK>(tc, k3, cc, lut, hs, rock_c_ptr, rnz, rock_c, &mv);
  chain_step<LAYERED, TRACK>(tc, k4, cc, lut, hs, rock_c_ptr, rnz, rock_c, &mv); chain_step<LAYERED, TRACK>(tc, k5, cc, lut, hs, rock_c_ptr, rnz, rock_c, &mv);
  chain_step<LAYERED, TRACK>(tc, k6, cc, lut, hs, rock_c_ptr, rnz, rock_c, &mv); chain_step<LAYERED, TRACK>(tc, k7, cc, lut, hs, rock_c_ptr, rnz, rock_c, &mv);
  tc[0] = cc;
  if (TRACK) *moved = mv ? (mv | (1u << 5)) : 0u;      // the centre (code 5) moved iff any transfer ran
}

// =============================================================================================
// Pair-bitmap mode of the whole-map tile kernel (interior tiles, reference parameters).
//
// Near the repose angle ~88 % of the centres have a pair whose excess is > 0 but only a few per
// cent have one whose transfer changes a value in fp32, and proving "nothing moves" centre by
// centre costs ~100 instructions (8 pairs, each tested from both of its ends over the sweep).
// Here every neighbour pair of the tile is tested ONCE, before colour 0, with an integer test:
//
//   the transfer excess/2 (settling == 1) leaves a and b unchanged if excess < min(u(a), u(b)),
//   u(v) = distance from |v| to the next float towards zero (= ulp(v), or ulp(v)/2 when |v| is a
//   power of two; 0 for v == 0). Then excess/2 is below half the spacing on either side of both
//   values and round-to-nearest returns them unchanged.
//
// Positive floats order like their bit patterns, so with key(v) = bits(u(v)) the test per pair is
// two subtractions and two integer compares: bits(excess) >= key(a) or >= key(b) (a negative
// excess has its sign bit set and compares below every key). Cells that are an end of such a
// "hot" pair get a bit in `hot`; cells a relaxation changed get a bit in `chg`. A centre needs
// the exact test only if its own hot bit is set or a cell of its 3x3 block changed earlier in this
// sweep -- anything else provably keeps all its pairs cold. The candidates run the exact refined
// test and the survivors the usual relaxation, so the result stays bitwise identical; the work per
// cell drops from ~100 to ~30 instructions when little moves (measured fractions: DESIGN.md).
// =============================================================================================
__device__ __forceinline__ int ulp_key(float v) {
  const int b = __float_as_int(v) & 0x7fffffff;
  return ((max(b, 1) - 1) & 0x7f800000) - (23 << 23);   // bits of u(v); <= 0 for tiny |v| (always hot)
}

__device__ __forceinline__ unsigned pair_bit(float a, int ka, float b, int kb, float t, unsigned bit) {
  const int eb = __float_as_int(fsub(fabsf(fsub(a, b)), t));
  return (eb >= min(ka, kb)) ? bit : 0u;
}

// bit i of `mask` = row row0 + i of column `col`
__device__ __forceinline__ void mask_or(unsigned* plane, int col, int row0, unsigned mask) {
  if (!mask) return;
  const int word = row0 >> 5, sh = row0 & 31;
  atomicOr(&plane[col * MASK_WORDS + word], mask << sh);
  if (sh && word + 1 < MASK_WORDS && (mask >> (32 - sh))) atomicOr(&plane[col * MASK_WORDS + word + 1], mask >> (32 - sh));
}

// Phase A: each warp walks a strip of 30 columns (+ one halo column each side: lanes 0 and 31) down
// a quarter of the tile's rows, the previous row in registers. A lane tests the four pairs that
// need only its west neighbour's values (one shuffle pair per row): W (x,z)-(x,z-1), N (x,z)-(x-1,z),
// NW (x,z)-(x-1,z-1) and the anti-diagonal AD (x-1,z)-(x,z-1). It collects the results as one bit
// per row; the partner cells are credited once per strip with three shuffles, and the lane ORs its
// column's bits into the column-major plane -- no ballots or atomics inside the row loop.
__device__ __forceinline__ void tile_pair_scan(const float* tile, unsigned* hot, const CascadeThr thr,
                                               int warp, int lane, int nwarps) {
  constexpr int CHUNKS = 4;                                       // <= 17 rows + 1 run-in row: fits 32 bits
  for (int item = warp; item < 8 * CHUNKS; item += nwarps) {
    const int g = item & 7, q = item >> 3;
    const int r0 = (TILE_LX * q) / CHUNKS, r1 = (TILE_LX * (q + 1)) / CHUNKS;
    const int z = 30 * g - 1 + lane;
    const bool zin = z >= 0 && z < TILE_PITCH;
    const bool w_ok = zin && lane >= 1 && z >= 1;                 // partner column z-1 is in the strip and the tile
    const int xs = r0 > 0 ? r0 - 1 : 0;                           // one row of run-in above the chunk
    const float* col = tile + (zin ? z : 0);
    float ph, phw;
    int pk, pkw;
    unsigned aW = 0u, aN = 0u, aNW = 0u, aAD = 0u;                // bit i: my pair of that kind at row xs + i is hot
    {                                                             // run-in row (or row 0): state only, W pairs if owned
      ph = col[xs * TILE_PITCH];
      pk = ulp_key(ph);
      phw = __shfl_up_sync(0xffffffffu, ph, 1);
      pkw = __shfl_up_sync(0xffffffffu, pk, 1);
      if (xs >= r0) aW = pair_bit(ph, pk, phw, pkw, thr.ortho, 1u);
    }
    // branch-free row loop (halo lanes compute garbage that is masked out after the loop)
    unsigned bit = 2u;
#pragma unroll 4
    for (int x = xs + 1; x < r1; x++, bit <<= 1) {
      const float h = col[x * TILE_PITCH];
      const int k = ulp_key(h);
      const float hw = __shfl_up_sync(0xffffffffu, h, 1);
      const int kw = __shfl_up_sync(0xffffffffu, k, 1);
      aN |= pair_bit(h, k, ph, pk, thr.ortho, bit);
      aW |= pair_bit(h, k, hw, kw, thr.ortho, bit);
      aNW |= pair_bit(h, k, phw, pkw, thr.diag, bit);
      aAD |= pair_bit(ph, pk, hw, kw, thr.diag, bit);
      ph = h; phw = hw; pk = k; pkw = kw;
    }
    if (!zin) aN = 0u;
    if (!w_ok) aW = aNW = aAD = 0u;
    // credit both ends. Bit i = row xs + i. My cell (x, z): W, N, NW at row i; AD marks (x-1, z): >> 1;
    // N marks (x-1, z) too. The lane to my right credits me with its W and AD (my cell is its
    // (x, z-1): row i) and with its NW (my cell is its (x-1, z-1): row i-1).
    const unsigned eW = __shfl_down_sync(0xffffffffu, aW, 1), eNW = __shfl_down_sync(0xffffffffu, aNW, 1);
    const unsigned eAD = __shfl_down_sync(0xffffffffu, aAD, 1);
    unsigned mine = aW | aN | aNW | ((aN | aAD) >> 1);
    if (lane < 31) mine |= eW | eAD | (eNW >> 1);
    if (zin) mask_or(hot, z, xs, mine);
  }
}

// positions first, first + 3, ..., first + 3 (count - 1) that fall into word j (bit b = position 32 j + b)
__device__ __forceinline__ unsigned stride3_mask(int j, int first, int count) {
  const int sel = (first % 3 + j) % 3;                            // 32 j + b == first (mod 3)  <=>  b == first + j (mod 3)
  unsigned m = sel == 0 ? 0x49249249u : (sel == 1 ? 0x92492492u : 0x24924924u);
  const int lo = first - 32 * j, hi = first + 3 * (count - 1) - 32 * j;
  if (lo > 31 || hi < 0) return 0u;
  if (lo > 0) m &= 0xffffffffu << lo;
  if (hi < 31) m &= 0xffffffffu >> (31 - hi);
  return m;
}

// rows of word j of column `col` that have a changed cell in rows x-1..x+1 of that column
__device__ __forceinline__ unsigned changed_near(const unsigned* chg, int col, int j) {
  const unsigned v = chg[col * MASK_WORDS + j];
  const unsigned l = j > 0 ? chg[col * MASK_WORDS + j - 1] : 0u;
  const unsigned r = j + 1 < MASK_WORDS ? chg[col * MASK_WORDS + j + 1] : 0u;
  return v | (v << 1) | (v >> 1) | (l >> 31) | (r << 31);
}

// moved: bit (dx+1)*4 + (dz+1) per 3x3 cell around (lx, lz)
__device__ __forceinline__ void mark_changed(unsigned* chg, int lx, int lz, unsigned moved) {
#pragma unroll
  for (int d = 0; d < 3; d++) {                                   // column lz + d - 1: rows lx-1, lx, lx+1
    const unsigned b3 = ((moved >> d) & 1u) | (((moved >> (4 + d)) & 1u) << 1) | (((moved >> (8 + d)) & 1u) << 2);
    mask_or(chg, lz + d - 1, lx - 1, b3);
  }
}

// ---- output ---------------------------------------------------------------------------------
// Tiles of one launch run in waves; a tile that starts late must still read the PRE-sweep values
// of its halo, i.e. of the border cells of its neighbours' outputs. Those border cells (the
// first 4 / last 2 rows and first 16 / last 8 columns of O: exactly the cells some other tile's
// R_0 covers, 20 % of the tile) are therefore written to a staging plane and copied into place
// by cascade_commit_kernel after the sweep kernel has finished; the interior of O, which no other
// tile reads, goes straight to the height plane.
template <int TX>
__device__ __forceinline__ bool tile_border(int ox, int oz) {
  return ox < 4 || ox >= TX - 2 || oz < 16 || oz >= TILE_TZ - 8;
}

template <int TX, int NTHREADS>
__device__ __forceinline__ void tile_store(const float* tile, const MapView& m, const RowWindow& win,
                                           float* __restrict__ stage, int ox0, int oz0, int tid, bool vec_ok) {
  const int nz = m.nz;
  if (vec_ok) {
    for (int i = tid; i < TX * (TILE_TZ / 4); i += NTHREADS) {
      const int ox = i / (TILE_TZ / 4), q = i - ox * (TILE_TZ / 4);
      const int gx = ox0 + ox, gz = oz0 + 4 * q;
      if (gx >= win.store_lo && gx < win.store_hi && gz >= 0 && gz < nz) {
        float* dst = tile_border<TX>(ox, 4 * q) ? stage : m.height;
        *reinterpret_cast<float4*>(dst + m.idx(gx, gz)) =
            *reinterpret_cast<const float4*>(tile + (ox + 2) * TILE_PITCH + 8 + 4 * q);
      }
    }
  } else {
    for (int i = tid; i < TX * TILE_TZ; i += NTHREADS) {
      const int ox = i / TILE_TZ, oz = i - ox * TILE_TZ;
      const int gx = ox0 + ox, gz = oz0 + oz;
      if (gx >= win.store_lo && gx < win.store_hi && gz >= 0 && gz < nz) {
        float* dst = tile_border<TX>(ox, oz) ? stage : m.height;
        dst[m.idx(gx, gz)] = tile[(ox + 2) * TILE_PITCH + (oz + 8)];
      }
    }
  }
}

// copies the staged border cells of every tile that changed into the height plane
template <int TX>
__global__ void __launch_bounds__(256)
cascade_commit_kernel(MapView m, RowWindow win, const float* __restrict__ stage, unsigned char* __restrict__ flags) {
  pdl_sync();
  const int tile_id = (win.tile_y0 + (int)blockIdx.y) * (int)gridDim.x + (int)blockIdx.x;
  if (!flags[tile_id]) return;
  const int ox0 = TILE_OX + (win.tile_y0 + (int)blockIdx.y) * TX, oz0 = TILE_OZ + (int)blockIdx.x * TILE_TZ;
  const int nz = m.nz;
  const bool vec_ok = (nz & 3) == 0;
  if (vec_ok) {
    // only the border slots are enumerated: 6 full rows (first 4, last 2) of TZ/4 float4, then the first
    // 4 and last 2 float4 of every row in between
    constexpr int QR = TILE_TZ / 4, N_FULL = 6 * QR, N_ALL = N_FULL + 6 * (TX - 6);
    for (int i = threadIdx.x; i < N_ALL; i += 256) {
      int ox, q;
      if (i < N_FULL) { const int r = i / QR; q = i - r * QR; ox = r < 4 ? r : TX - 6 + r; }
      else { const int j = i - N_FULL, r = j / 6, s_ = j - r * 6; ox = 4 + r; q = s_ < 4 ? s_ : QR - 6 + s_; }
      const int gx = ox0 + ox, gz = oz0 + 4 * q;
      if (gx >= win.store_lo && gx < win.store_hi && gz >= 0 && gz < nz)
        *reinterpret_cast<float4*>(m.height + m.idx(gx, gz)) = *reinterpret_cast<const float4*>(stage + m.idx(gx, gz));
    }
  } else {
    for (int i = threadIdx.x; i < TX * TILE_TZ; i += 256) {
      const int ox = i / TILE_TZ, oz = i - ox * TILE_TZ;
      if (!tile_border<TX>(ox, oz)) continue;
      const int gx = ox0 + ox, gz = oz0 + oz;
      if (gx >= win.store_lo && gx < win.store_hi && gz >= 0 && gz < nz) m.height[m.idx(gx, gz)] = stage[m.idx(gx, gz)];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) flags[tile_id] = 0;
}

// USE_TMA: the height tile arrives by one cp.async.bulk.tensor.2d (UTMALDG) with hardware
// zero-fill outside the map; needs nz % 4 == 0 (16-byte row pitch). Otherwise plain loads.
//
// Each colour phase is split in two so that the expensive part runs on full warps:
//   scan  : every thread tests its centres (stamp, early-out) and appends the ones that need
//           World::cascade to a shared-memory queue (warp-aggregated append);
//   relax : the queue is processed densely, one entry per thread.
// On a map close to the repose angle only a few per cent of the centres reach the queue.
#ifndef SWE_TILE_MIN_BLOCKS
#define SWE_TILE_MIN_BLOCKS 3     // 40 registers, three 512-thread CTAs per SM: the per-colour barriers of one CTA hide behind the others
#endif
#ifndef SWE_TILE_DIRECT
#define SWE_TILE_DIRECT 1         // dense-direct path (see cascade_tile_kernel)
#endif
#ifndef SWE_TILE_DIRECT_PCT
#define SWE_TILE_DIRECT_PCT 45u   // moved centres (per cent) above which the next sweep of the tile runs dense-direct
#endif
template <bool USE_TMA>
__global__ void __launch_bounds__(TILE_THREADS, SWE_TILE_MIN_BLOCKS)
cascade_tile_kernel(const __grid_constant__ CUtensorMap tmap, MapView m, float settling, CascadeThr thr,
                    RowWindow win, TileScratch scratch) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* tile = reinterpret_cast<float*>(smem_raw);
  unsigned short* queue = reinterpret_cast<unsigned short*>(smem_raw + TILE_FLOAT_BYTES);
  unsigned short* queue2 = reinterpret_cast<unsigned short*>(smem_raw + TILE_FLOAT_BYTES + TILE_QUEUE_BYTES);
  unsigned* hot = reinterpret_cast<unsigned*>(smem_raw + TILE_FLOAT_BYTES + 2 * TILE_QUEUE_BYTES);
  unsigned* chg = hot + TILE_PITCH * MASK_WORDS;
  // [0],[1] queue counters of the dense path, [2] hot-cell count, [3] relaxed centres (dense path)
  unsigned* qcount = reinterpret_cast<unsigned*>(smem_raw + TILE_FLOAT_BYTES + 2 * TILE_QUEUE_BYTES + 2 * TILE_MASK_BYTES);
  unsigned* ccount = qcount + 4;                   // pair-bitmap mode: [9] candidates + [9] survivors per colour
  unsigned long long* mbar = reinterpret_cast<unsigned long long*>(qcount + 24);
  float2* lut = reinterpret_cast<float2*>(qcount + 32);   // chain table, 16 entries
  const int ox0 = TILE_OX + (win.tile_y0 + (int)blockIdx.y) * TILE_TX, oz0 = TILE_OZ + (int)blockIdx.x * TILE_TZ;
  const int tile_id = (win.tile_y0 + (int)blockIdx.y) * (int)gridDim.x + (int)blockIdx.x;
  const int sx0 = m.x0, sx1 = m.x0 + m.rows;       // rows present in local storage (a stripe + ghosts)
  const int gx0 = ox0 - 2, gz0 = oz0 - 8;          // global coords of tile[0][0]
  const int tid = (int)threadIdx.x;
  const int tx = tid % TILE_BX, ty = tid / TILE_BX;  // dense scan: centre slot (ty == TILE_BY: the 16th warp sits out)
  const int nx = m.nx, nz = m.nz;
  const float hs = fmul(settling, 0.5f);
  // negative or vanishing thresholds / sub-normal products: the lean paths' shortcuts do not hold
  const bool generic_only = !(thr.ortho >= 0.0f) || !(thr.diag >= 0.0f) || !(fabsf(settling) >= 1e-18f) ||
                            !(fabsf(settling) <= 1e18f) || (thr.ortho != 0.0f && thr.ortho < 1e-18f);
  const bool vec_ok = (nz & 3) == 0;               // rows are 16-byte aligned, tiles 4-aligned in z
  const bool interior = gx0 >= 0 && gz0 >= 0 && gx0 + TILE_LX <= nx && gz0 + TILE_PITCH <= nz;
  // pair-bitmap mode: interior tile, reference parameters (settling == 1: the transfer excess/2 is
  // exact), and the tile was not mostly moving in the previous sweep (hint byte; heuristic only)
  bool sparse = interior && !generic_only && !m.rock && settling == 1.0f && thr.diag < 1e30f &&
                scratch.hint != nullptr && scratch.hint[tile_id] == 0;

  // ---- load R_0 ---------------------------------------------------------------------------
  if (tid < 22) qcount[tid] = 0u;                  // qcount[4] + ccount[18]
  if (tid >= 32 && tid < 48) chain_table_init(lut, tid - 32, thr);
  for (int i = tid; i < 2 * TILE_PITCH * MASK_WORDS; i += TILE_THREADS) hot[i] = 0u;   // hot and chg planes
  if (USE_TMA) {
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)),
                   "r"(TILE_FLOAT_BYTES) : "memory");
      asm volatile(
          "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
          ::"r"(smem_u32(tile)), "l"(&tmap), "r"(gz0), "r"(gx0 - m.x0), "r"(smem_u32(mbar)) : "memory");
    }
  } else {
    for (int i = tid; i < TILE_LX * TILE_PITCH; i += TILE_THREADS) {
      const int lx = i / TILE_PITCH, lz = i - lx * TILE_PITCH;
      const int gx = gx0 + lx, gz = gz0 + lz;
      float v = 0.0f;
      if (gx >= sx0 && gx < sx1 && gz >= 0 && gz < nz) v = __ldg(m.height + m.idx(gx, gz));
      tile[i] = v;
    }
  }
  __syncthreads();   // mbarrier init + counters + bit planes visible; plain loads complete
  if (USE_TMA) {
    unsigned done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done) : "r"(smem_u32(mbar)) : "memory");
    }
  }

  bool any_change = false;
  const unsigned lane = tid & 31;

  // ---- pair-bitmap mode: phase A, then decide ------------------------------------------------
  unsigned hot_cells = 0u;
  if (sparse) {
    tile_pair_scan(tile, hot, thr, tid >> 5, (int)lane, TILE_THREADS / 32);
    __syncthreads();
    unsigned n = 0;
    for (int i = tid; i < TILE_PITCH * MASK_WORDS; i += TILE_THREADS) n += __popc(hot[i]);
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if (lane == 0 && n) atomicAdd(&qcount[2], n);
    __syncthreads();
    hot_cells = qcount[2];
    sparse = hot_cells * 100u < 27u * (unsigned)(TILE_LX * TILE_PITCH);   // CTA-uniform
  }
  if (sparse) {
    const Valid8 all = centre_valid<true>(1, 1, 3, 3);
    // no hot pair anywhere in the tile: no centre can move and nothing can become hot
#pragma unroll 1
    for (int c = 0; c < 9 && hot_cells != 0u; c++) {
      const int cxc = c / 3;
      const int ncx = TILE_TX / 3 + 2 - cxc, ncz = TILE_TZ / 3 + 8 - c;
      // candidates: own hot bit, or a changed cell in the 3x3 block (one queue reservation per warp)
      for (int it0 = 0; it0 < ncz * MASK_WORDS; it0 += TILE_THREADS) {      // one trip: ncz * 3 <= 240
        const int it = it0 + tid;
        unsigned wb = 0u;
        int lz = 0, j = 0;
        if (it < ncz * MASK_WORDS) {
          const int t = it / MASK_WORDS;
          j = it - t * MASK_WORDS;
          lz = c + 1 + 3 * t;
          wb = hot[lz * MASK_WORDS + j] | changed_near(chg, lz - 1, j) | changed_near(chg, lz, j) |
               changed_near(chg, lz + 1, j);
          wb &= stride3_mask(j, cxc + 1, ncx);
        }
        if (__any_sync(0xffffffffu, wb != 0u)) {
          const unsigned n = (unsigned)__popc(wb);
          unsigned incl = n;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const unsigned up = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)lane >= o) incl += up;
          }
          unsigned base = 0u;
          if (lane == 31u) base = atomicAdd(&ccount[c], incl);
          base = __shfl_sync(0xffffffffu, base, 31);
          unsigned slot = base + incl - n;
          while (wb) {
            const int b = __ffs(wb) - 1;
            wb &= wb - 1u;
            queue[slot++] = (unsigned short)((32 * j + b) * TILE_PITCH + lz);
          }
        }
      }
      __syncthreads();
      const unsigned nc = ccount[c];
      if (nc == 0u) continue;                                      // CTA-uniform
      if (nc <= (unsigned)TILE_THREADS) {
        // few candidates: exact test and relaxation by the same thread, one barrier less
        bool did = false;
        if ((unsigned)tid < nc) {
          const int off = queue[tid];
          if (centre_needs<true, true>(tile + off, thr, hs, all)) {
            unsigned moved = 0u;
            centre_relax<true, false, true>(tile + off, thr, lut, settling, hs, all, 1, 1, 3, 3, nullptr, &moved);
            if (moved) { const int lx = off / TILE_PITCH; mark_changed(chg, lx, off - lx * TILE_PITCH, moved); }
            did = true;
          }
        }
        any_change |= did;
        __syncthreads();
        continue;
      }
      // exact early-out test, dense over the candidates
      for (unsigned base_i = 0; base_i < nc; base_i += TILE_THREADS) {
        const unsigned i = base_i + tid;
        int off = 0;
        bool need = false;
        if (i < nc) { off = queue[i]; need = centre_needs<true, true>(tile + off, thr, hs, all); }
        const unsigned mask = __ballot_sync(0xffffffffu, need);
        if (mask) {
          unsigned b = 0;
          if (lane == (unsigned)(__ffs(mask) - 1)) b = atomicAdd(&ccount[9 + c], (unsigned)__popc(mask));
          b = __shfl_sync(0xffffffffu, b, __ffs(mask) - 1);
          if (need) queue2[b + __popc(mask & ((1u << lane) - 1u))] = (unsigned short)off;
        }
      }
      __syncthreads();
      // relax the survivors, remember what they changed
      const unsigned nq = ccount[9 + c];
      any_change |= nq != 0u;
      for (unsigned i = tid; i < nq; i += TILE_THREADS) {
        const int off = queue2[i];
        unsigned moved = 0u;
        centre_relax<true, false, true>(tile + off, thr, lut, settling, hs, all, 1, 1, 3, 3, nullptr, &moved);
        if (moved) { const int lx = off / TILE_PITCH; mark_changed(chg, lx, off - lx * TILE_PITCH, moved); }
      }
      __syncthreads();
    }
    any_change = __syncthreads_or(any_change ? 1 : 0) != 0;        // per-thread flags in the merged path
  } else {
  // which early-out test the scan uses: the refined test costs ~64 more instructions per
  // scanned centre and saves ~250 per centre it keeps out of the queue, so it pays unless more
  // than ~3/4 of the centres pass it. Colour 0 always probes with it; a tile whose refined
  // queue fills beyond 3/4 falls back to the coarse test for its remaining colours.
  bool refined = true;
  unsigned relaxed = 0u, centres = 0u;
  // DENSE-DIRECT: while more than ~45 % of a tile's centres move (the first tens of sweeps of a relaxation from
  // noise) testing every centre first costs more than it saves: every thread relaxes its centres in place in the
  // scan's stride-3 mapping -- no early-out test, no queue, no ballots or atomics, one barrier per colour instead
  // of two, perfectly balanced; a centre that needs nothing runs the chain with every transfer predicated off
  // (same result). The moved centres are counted so that the hint can fall back to the scan path.
  const bool direct = SWE_TILE_DIRECT && interior && !generic_only && !m.rock && scratch.hint != nullptr && scratch.hint[tile_id] == 2;
  if (direct) {
    const Valid8 all = centre_valid<true>(1, 1, 3, 3);
    unsigned mine = 0u;
#pragma unroll 1
    for (int c = 0; c < 9; c++) {
      const int cxc = c / 3;
      const int ncx = TILE_TX / 3 + 2 - cxc, ncz = TILE_TZ / 3 + 8 - c;
      if (ty < TILE_BY && tx < ncz) {
        const int lz = c + 1 + 3 * tx;
#pragma unroll 1
        for (int cxi = ty; cxi < ncx; cxi += TILE_BY) {
          float* tc = tile + (cxc + 1 + 3 * cxi) * TILE_PITCH + lz;
          const float before = tc[0];
          centre_relax<true>(tc, thr, lut, settling, hs, all, 1, 1, 3, 3);
          mine += (tc[0] != before) ? 1u : 0u;    // (a transfer that ran without changing the centre in fp32 does not count)
        }
      }
      centres += (unsigned)(ncx * ncz);
      __syncthreads();
    }
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if (lane == 0 && mine) atomicAdd(&qcount[3], mine);
    __syncthreads();
    relaxed = qcount[3];
    any_change = relaxed != 0u;
  } else {
#pragma unroll 1
  for (int c = 0; c < 9; c++) {
    const int cxc = c / 3;
    const int ncx = TILE_TX / 3 + 2 - cxc, ncz = TILE_TZ / 3 + 8 - c;
    unsigned* qc = qcount + (c & 1);
    // ---- scan --------------------------------------------------------------------------------
    {
      const int lz = c + 1 + 3 * tx, gz = gz0 + lz;
      const bool z_ok = ty < TILE_BY && tx < ncz && (interior || (gz >= 0 && gz < nz));
#pragma unroll 1
      for (int cx_base = 0; cx_base < ncx; cx_base += TILE_BY) {      // same trip count for all threads
        const int cxi = cx_base + ty;
        const int lx = cxc + 1 + 3 * cxi, gx = gx0 + lx;
        bool need = z_ok && cxi < ncx && (interior || (gx >= 0 && gx < nx));
        const int off = lx * TILE_PITCH + lz;
        if (need && !generic_only) {
          const float* tc = tile + off;
          if (interior)
            need = refined ? centre_needs<true, true>(tc, thr, hs, centre_valid<true>(gx, gz, nx, nz))
                           : centre_needs<true, false>(tc, thr, hs, centre_valid<true>(gx, gz, nx, nz));
          else
            need = refined ? centre_needs<false, true>(tc, thr, hs, centre_valid<false>(gx, gz, nx, nz))
                           : centre_needs<false, false>(tc, thr, hs, centre_valid<false>(gx, gz, nx, nz));
        }
        const unsigned mask = __ballot_sync(0xffffffffu, need);
        if (mask) {
          unsigned base = 0;
          if (lane == (unsigned)(__ffs(mask) - 1)) base = atomicAdd(qc, (unsigned)__popc(mask));
          base = __shfl_sync(0xffffffffu, base, __ffs(mask) - 1);
          if (need) queue[base + __popc(mask & ((1u << lane) - 1u))] = (unsigned short)off;
        }
      }
    }
    __syncthreads();
    // ---- relax -------------------------------------------------------------------------------
    const unsigned nq = *qc;
    if (tid == 0) qcount[(c + 1) & 1] = 0u;          // the other counter is idle during this phase
    any_change |= nq != 0u;                          // a queued centre always moves sand
    relaxed += nq; centres += (unsigned)(ncx * ncz);
    refined = refined && (nq * 4u < (unsigned)(ncx * ncz) * 3u);
    for (unsigned i = tid; i < nq; i += TILE_THREADS) {
      const int off = queue[i];
      float* tc = tile + off;
      if (interior && !generic_only && !m.rock) {
        // interior: every neighbour exists; (1,1) in a 3x3 map tells the generic fallback so
        centre_relax<true>(tc, thr, lut, settling, hs, centre_valid<true>(1, 1, 3, 3), 1, 1, 3, 3);
      } else {
        const int lx = off / TILE_PITCH, lz = off - lx * TILE_PITCH;
        const int gx = gx0 + lx, gz = gz0 + lz;
        if (m.rock) {   // layered extension (single-GPU engines): sand-limited transfers
          const float* rp = m.rock + m.idx(gx, gz);
          if (generic_only) tile_cascade_generic(tc, gx, gz, nx, nz, thr, settling, rp);
          else centre_relax<false, true>(tc, thr, lut, settling, hs, centre_valid<false>(gx, gz, nx, nz), gx, gz, nx, nz, rp);
        }
        else if (generic_only) tile_cascade_generic(tc, gx, gz, nx, nz, thr, settling);
        else centre_relax<false>(tc, thr, lut, settling, hs, centre_valid<false>(gx, gz, nx, nz), gx, gz, nx, nz);
      }
    }
    __syncthreads();
  }
  }
  // hint for the next sweep: 0 = pair-bitmap path, 1 = skip the pair scan while most of this tile still moves,
  // 2 = dense-direct while nearly half of the centres move
  if (tid == 0 && scratch.hint != nullptr)
    scratch.hint[tile_id] = (relaxed * 100u > SWE_TILE_DIRECT_PCT * centres) ? 2 : ((relaxed * 100u > 26u * centres) ? 1 : 0);
  }

  // ---- store O if anything in this tile moved -----------------------------------------------
  if (any_change) {                                  // CTA-uniform (derived from shared counters)
    tile_store<TILE_TX, TILE_THREADS>(tile, m, win, scratch.stage, ox0, oz0, tid, vec_ok);
    if (tid == 0) scratch.flags[tile_id] = 1;
  }
}

// =============================================================================================
// Stamped sweep (flight pipeline): World::cascade once at every cell a particle touched in the
// round, same colour order, same tiles -- but the work is driven by the touched cells, not by
// the map: while the stamp bytes of the tile are read, every touched cell is appended to the
// queue of its colour (nine shared-memory queues); a tile without touched cells returns before
// it loads any heights; otherwise each colour phase runs the early-out test densely over its
// queue and relaxes the survivors densely. Touched cells in the halo outside R_c are processed
// too: their blocks do not reach the cells the tile stores (they would have been in R_c), so
// this only costs a few redundant cascades.
// =============================================================================================
// The stamped sweep spends most of its time waiting at the 18 barriers of a tile with only a few
// warps busy (relax of a few hundred centres), so it uses HALF-HEIGHT tiles (30 x 216 outputs,
// 256 threads, 53 KB): four CTAs per SM instead of two hide each other's barrier phases. The
// extra halo (1.33x instead of 1.22x) only costs redundant work where particles were.
constexpr int STAMPED_TX = 30;
template <int TX> struct StampedGeo {
  static constexpr int LX = TX + 6;
  static constexpr int THREADS = (TX <= 30) ? 256 : 480;               // queue-driven: a flat block, whole warps
  static constexpr int FLOAT_BYTES = LX * TILE_PITCH * 4;
  static constexpr int QCAP = ((LX + 2) / 3) * (TILE_PITCH / 3);          // centres of one colour in the tile
  static constexpr int SMEM = FLOAT_BYTES + 10 * QCAP * 2 + 96 + 128;   // + chain table
};

template <int TX, bool USE_TMA>
__global__ void __launch_bounds__(StampedGeo<TX>::THREADS, (TX <= 30) ? 4 : 2)
cascade_tile_stamped_kernel(const __grid_constant__ CUtensorMap tmap, MapView m, float settling, CascadeThr thr,
                            const unsigned char* __restrict__ stamp, unsigned char stamp_now,
                            const unsigned* __restrict__ gate, RowWindow win, TileScratch scratch) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  pdl_sync();
  if (gate && *gate == 0u) return;
  constexpr int TILE_LX = StampedGeo<TX>::LX, TILE_THREADS = StampedGeo<TX>::THREADS;     // shadow the 60-row geometry
  constexpr int TILE_FLOAT_BYTES = StampedGeo<TX>::FLOAT_BYTES, STAMPED_QCAP = StampedGeo<TX>::QCAP;
  constexpr int TILE_TX = TX;
  float* tile = reinterpret_cast<float*>(smem_raw);
  unsigned short* cq = reinterpret_cast<unsigned short*>(smem_raw + TILE_FLOAT_BYTES);   // [9][QCAP] by colour
  unsigned short* rq = cq + 9 * STAMPED_QCAP;                                              // relax queue
  unsigned* cnt = reinterpret_cast<unsigned*>(rq + STAMPED_QCAP);                          // [9] touched + [9] relax
  unsigned long long* mbar = reinterpret_cast<unsigned long long*>(cnt + 18);
  float2* lut = reinterpret_cast<float2*>(cnt + 24);      // chain table, 16 entries
  const int ox0 = TILE_OX + (win.tile_y0 + (int)blockIdx.y) * TILE_TX, oz0 = TILE_OZ + (int)blockIdx.x * TILE_TZ;
  const int gx0 = ox0 - 2, gz0 = oz0 - 8;
  const int sx0 = m.x0, sx1 = m.x0 + m.rows;
  const int tid = threadIdx.y * TILE_BX + threadIdx.x;
  const int nx = m.nx, nz = m.nz;
  const float hs = fmul(settling, 0.5f);
  const bool generic_only = !(thr.ortho >= 0.0f) || !(thr.diag >= 0.0f) || !(fabsf(settling) >= 1e-18f) ||
                            !(fabsf(settling) <= 1e18f) || (thr.ortho != 0.0f && thr.ortho < 1e-18f);
  const bool vec_ok = (nz & 3) == 0;
  const bool interior = gx0 >= 0 && gz0 >= 0 && gx0 + TILE_LX <= nx && gz0 + TILE_PITCH <= nz;

  if (tid < 18) cnt[tid] = 0u;
  if (tid >= 32 && tid < 48) chain_table_init(lut, tid - 32, thr);
  if (USE_TMA && tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  // ---- touched cells -> per-colour queues --------------------------------------------------
  // gx0 == 2 and gz0 == 2 (mod 3): colour of local cell (lx, lz) = ((lx+2)%3)*3 + (lz+2)%3
  auto push = [&](int lx, int lz) {
    if (lx < 1 || lx > TILE_LX - 2 || lz < 1 || lz > TILE_PITCH - 2) return;   // block must fit the tile
    const int c = ((lx + 2) % 3) * 3 + (lz + 2) % 3;
    const unsigned slot = atomicAdd(&cnt[c], 1u);
    cq[c * STAMPED_QCAP + slot] = (unsigned short)(lx * TILE_PITCH + lz);
  };
  if ((nz & 7) == 0) {
    // 8 stamps per 64-bit load (rows are 8-byte aligned and every tile's z origin, -16 + 216 j, is
    // == 0 mod 8); all loads of a thread are issued before any is inspected, and a word is only
    // taken apart if the "has a zero byte" test says one of its stamps matches
    const unsigned pat = 0x01010101u * stamp_now;
    constexpr int QPR = TILE_PITCH / 8;                       // 30 uint2 per tile row
    constexpr int NV = TILE_LX * QPR;
    constexpr int PER = (NV + TILE_THREADS - 1) / TILE_THREADS;
    uint2 v[PER];
#pragma unroll
    for (int k = 0; k < PER; k++) {
      const int i = tid + k * TILE_THREADS;
      const int lx = i / QPR, q = i - lx * QPR;
      const int gx = gx0 + lx, gz = gz0 + 8 * q;
      v[k] = make_uint2(~pat, ~pat);
      if (i < NV && gx >= sx0 && gx < sx1 && gx >= 0 && gx < nx && gz >= 0 && gz < nz)
        v[k] = __ldcg(reinterpret_cast<const uint2*>(stamp + m.idx(gx, gz)));
    }
#pragma unroll
    for (int k = 0; k < PER; k++) {
      const int i = tid + k * TILE_THREADS;
      const int lx = i / QPR, q = i - lx * QPR;
      const unsigned w2[2] = {v[k].x ^ pat, v[k].y ^ pat};
#pragma unroll
      for (int j = 0; j < 2; j++) {
        const unsigned x = w2[j];
        if (((x - 0x01010101u) & ~x & 0x80808080u) == 0u) continue;   // no byte of x is zero
        const int lz = 8 * q + 4 * j;
        if ((x & 0x000000ffu) == 0u) push(lx, lz);
        if ((x & 0x0000ff00u) == 0u) push(lx, lz + 1);
        if ((x & 0x00ff0000u) == 0u) push(lx, lz + 2);
        if ((x & 0xff000000u) == 0u) push(lx, lz + 3);
      }
    }
  } else if (vec_ok) {
    const unsigned pat = 0x01010101u * stamp_now;
    for (int i = tid; i < TILE_LX * (TILE_PITCH / 4); i += TILE_THREADS) {
      const int lx = i / (TILE_PITCH / 4), q = i - lx * (TILE_PITCH / 4);
      const int gx = gx0 + lx, gz = gz0 + 4 * q;
      if (gx < sx0 || gx >= sx1 || gx < 0 || gx >= nx || gz < 0 || gz >= nz) continue;
      const unsigned v = __ldcg(reinterpret_cast<const unsigned*>(stamp + m.idx(gx, gz))) ^ pat;
      if ((v & 0x000000ffu) == 0u) push(lx, 4 * q);
      if ((v & 0x0000ff00u) == 0u) push(lx, 4 * q + 1);
      if ((v & 0x00ff0000u) == 0u) push(lx, 4 * q + 2);
      if ((v & 0xff000000u) == 0u) push(lx, 4 * q + 3);
    }
  } else {
    for (int i = tid; i < TILE_LX * TILE_PITCH; i += TILE_THREADS) {
      const int lx = i / TILE_PITCH, lz = i - lx * TILE_PITCH;
      const int gx = gx0 + lx, gz = gz0 + lz;
      if (gx < sx0 || gx >= sx1 || gx < 0 || gx >= nx || gz < 0 || gz >= nz) continue;
      if (__ldcg(stamp + m.idx(gx, gz)) == stamp_now) push(lx, lz);
    }
  }
  __syncthreads();
  {
    unsigned total = 0;
#pragma unroll
    for (int c = 0; c < 9; c++) total += cnt[c];
    if (total == 0u) return;                       // nothing touched here: no height traffic at all
  }

  // ---- heights -----------------------------------------------------------------------------
  if (USE_TMA) {
    if (tid == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)),
                   "r"(TILE_FLOAT_BYTES) : "memory");
      asm volatile(
          "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
          ::"r"(smem_u32(tile)), "l"(&tmap), "r"(gz0), "r"(gx0 - m.x0), "r"(smem_u32(mbar)) : "memory");
    }
    unsigned done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done) : "r"(smem_u32(mbar)) : "memory");
    }
  } else {
    for (int i = tid; i < TILE_LX * TILE_PITCH; i += TILE_THREADS) {
      const int lx = i / TILE_PITCH, lz = i - lx * TILE_PITCH;
      const int gx = gx0 + lx, gz = gz0 + lz;
      float v = 0.0f;
      if (gx >= sx0 && gx < sx1 && gz >= 0 && gz < nz) v = __ldg(m.height + m.idx(gx, gz));
      tile[i] = v;
    }
    __syncthreads();
  }

  bool any_change = false;
  const unsigned lane = tid & 31;
#pragma unroll 1
  for (int c = 0; c < 9; c++) {
    const unsigned nc = cnt[c];
    if (nc == 0u) continue;                        // CTA-uniform
    unsigned* qc = cnt + 9 + c;                    // one relax counter per colour: no reset races
    const unsigned short* mine = cq + c * STAMPED_QCAP;
    if (nc <= 64u) {
      // sparse tile (few particles per tile, e.g. 16384^2 with 0.06 particles/cell): test and
      // relax in one pass on at most two warps -- one barrier per colour instead of two, no queue
      if ((unsigned)tid < nc) {
        const int off = mine[tid];
        const int lx = off / TILE_PITCH, lz = off - lx * TILE_PITCH;
        const int cxi3 = lx - (c / 3) - 1, czi3 = lz - c - 1;
        const int gx = gx0 + lx, gz = gz0 + lz;
        if (cxi3 >= 0 && cxi3 < 3 * (TILE_TX / 3 + 2 - c / 3) && czi3 >= 0 && czi3 < 3 * (TILE_TZ / 3 + 8 - c) &&
            gx >= 0 && gx < nx && gz >= 0 && gz < nz) {
          float* tc = tile + off;
          if (m.rock && !generic_only) {
            const Valid8 v = centre_valid<false>(gx, gz, nx, nz);
            if (centre_needs<false, false>(tc, thr, hs, v)) {
              centre_relax<false, true>(tc, thr, lut, settling, hs, v, gx, gz, nx, nz, m.rock + m.idx(gx, gz));
              any_change = true;
            }
          } else if (m.rock) {
            any_change |= tile_cascade_generic(tc, gx, gz, nx, nz, thr, settling, m.rock + m.idx(gx, gz));
          } else if (generic_only) {
            any_change |= tile_cascade_generic(tc, gx, gz, nx, nz, thr, settling);
          } else if (interior) {
            if (centre_needs<true, false>(tc, thr, hs, centre_valid<true>(1, 1, 3, 3))) {
              centre_relax<true>(tc, thr, lut, settling, hs, centre_valid<true>(1, 1, 3, 3), 1, 1, 3, 3);
              any_change = true;
            }
          } else {
            const Valid8 v = centre_valid<false>(gx, gz, nx, nz);
            if (centre_needs<false, false>(tc, thr, hs, v)) {
              centre_relax<false>(tc, thr, lut, settling, hs, v, gx, gz, nx, nz);
              any_change = true;
            }
          }
        }
      }
      __syncthreads();
      continue;
    }
    // ---- scan: early-out test, dense over the touched centres of this colour -----------------
    for (unsigned base_i = 0; base_i < nc; base_i += TILE_THREADS) {
      const unsigned i = base_i + tid;
      bool need = i < nc;
      int off = 0;
      if (need) {
        off = mine[i];
        {   // only the colour-c centres inside R_c are part of this tile's dependency cone
          const int lx = off / TILE_PITCH, lz = off - lx * TILE_PITCH;
          const int cxi3 = lx - (c / 3) - 1, czi3 = lz - c - 1;
          need = cxi3 >= 0 && cxi3 < 3 * (TILE_TX / 3 + 2 - c / 3) && czi3 >= 0 && czi3 < 3 * (TILE_TZ / 3 + 8 - c);
        }
        if (need && !generic_only) {
          const float* tc = tile + off;
          if (interior) {
            need = centre_needs<true, false>(tc, thr, hs, centre_valid<true>(1, 1, 3, 3));
          } else {
            const int lx = off / TILE_PITCH, lz = off - lx * TILE_PITCH;
            const int gx = gx0 + lx, gz = gz0 + lz;
            need = gx >= 0 && gx < nx && gz >= 0 && gz < nz &&
                   centre_needs<false, false>(tc, thr, hs, centre_valid<false>(gx, gz, nx, nz));
          }
        }
      }
      const unsigned mask = __ballot_sync(0xffffffffu, need);
      if (mask) {
        unsigned b = 0;
        if (lane == (unsigned)(__ffs(mask) - 1)) b = atomicAdd(qc, (unsigned)__popc(mask));
        b = __shfl_sync(0xffffffffu, b, __ffs(mask) - 1);
        if (need) rq[b + __popc(mask & ((1u << lane) - 1u))] = (unsigned short)off;
      }
    }
    __syncthreads();
    // ---- relax ------------------------------------------------------------------------------
    const unsigned nq = *qc;
    any_change |= nq != 0u;
    for (unsigned i = tid; i < nq; i += TILE_THREADS) {
      const int off = rq[i];
      float* tc = tile + off;
      if (interior && !generic_only && !m.rock) {
        centre_relax<true>(tc, thr, lut, settling, hs, centre_valid<true>(1, 1, 3, 3), 1, 1, 3, 3);
      } else {
        const int lx = off / TILE_PITCH, lz = off - lx * TILE_PITCH;
        const int gx = gx0 + lx, gz = gz0 + lz;
        if (m.rock) {   // layered extension (single-GPU engines): sand-limited transfers
          const float* rp = m.rock + m.idx(gx, gz);
          if (generic_only) tile_cascade_generic(tc, gx, gz, nx, nz, thr, settling, rp);
          else centre_relax<false, true>(tc, thr, lut, settling, hs, centre_valid<false>(gx, gz, nx, nz), gx, gz, nx, nz, rp);
        }
        else if (generic_only) tile_cascade_generic(tc, gx, gz, nx, nz, thr, settling);
        else centre_relax<false>(tc, thr, lut, settling, hs, centre_valid<false>(gx, gz, nx, nz), gx, gz, nx, nz);
      }
    }
    __syncthreads();
  }

  if (__syncthreads_or(any_change ? 1 : 0)) {        // per-thread flags in the sparse path
    tile_store<TX, TILE_THREADS>(tile, m, win, scratch.stage, ox0, oz0, tid, vec_ok);
    if (tid == 0) scratch.flags[(win.tile_y0 + (int)blockIdx.y) * (int)gridDim.x + (int)blockIdx.x] = 1;
  }
}

// max over neighbour pairs of |dh| - d*maxdiff. Each cell checks its 4 "forward"
// neighbours (0,1),(1,-1),(1,0),(1,1) so every pair is visited once.
__device__ __forceinline__ float atomicMaxFloat(float* addr, float v) {
  // order-preserving integer trick; *addr initialised to -inf
  if (v >= 0) return __int_as_float(atomicMax((int*)addr, __float_as_int(v)));
  return __uint_as_float(atomicMin((unsigned*)addr, __float_as_uint(v)));
}

__global__ void max_excess_kernel(MapView m, swe_params P, float* out) {
  const CascadeThr thr = cascade_thresholds(P);
  const size_t n = (size_t)m.rows * m.nz;
  float best = -INFINITY;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const int x = m.x0 + (int)(i / m.nz), z = (int)(i % m.nz);
    const float h = m.height[i];
    if (m.periodic) {   // on a torus every cell has all four forward neighbours
      best = fmaxf(best, fsub(fabsf(fsub(h, m.height[m.idx(x, z + 1)])), thr.ortho));
      best = fmaxf(best, fsub(fabsf(fsub(h, m.height[m.idx(x + 1, z)])), thr.ortho));
      best = fmaxf(best, fsub(fabsf(fsub(h, m.height[m.idx(x + 1, z - 1)])), thr.diag));
      best = fmaxf(best, fsub(fabsf(fsub(h, m.height[m.idx(x + 1, z + 1)])), thr.diag));
      continue;
    }
    if (z + 1 < m.nz) best = fmaxf(best, fsub(fabsf(fsub(h, m.height[i + 1])), thr.ortho));
    if (x + 1 < m.nx && x + 1 < m.x0 + m.rows) {
      const size_t j = i + m.nz;
      best = fmaxf(best, fsub(fabsf(fsub(h, m.height[j])), thr.ortho));
      if (z >= 1) best = fmaxf(best, fsub(fabsf(fsub(h, m.height[j - 1])), thr.diag));
      if (z + 1 < m.nz) best = fmaxf(best, fsub(fabsf(fsub(h, m.height[j + 1])), thr.diag));
    }
  }
  for (int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
  if ((threadIdx.x & 31) == 0 && best > -INFINITY) atomicMaxFloat(out, best);
}

__global__ void set_float_kernel(float* p, float v) { *p = v; }

__global__ void add_cascade_cells_kernel(DevStats* s, unsigned long long n) { s->cascade_cells += n; }

// cuTensorMapEncodeTiled through the runtime's driver entry point (no libcuda link needed)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

bool make_height_tensor_map(const MapView& m, HeightTensorMap* out) {
  out->valid = false;
  if ((m.nz & 3) != 0) return false;                        // row pitch must be a multiple of 16 B
  static EncodeTiledFn encode = nullptr;
  static std::once_flag looked;           // engines may be created from several host threads (one per GPU)
  std::call_once(looked, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      encode = reinterpret_cast<EncodeTiledFn>(fn);
  });
  if (!encode) return false;
  static_assert(sizeof(CUtensorMap) == sizeof(out->opaque), "tensor map size");
  const cuuint64_t dims[2] = {(cuuint64_t)m.nz, (cuuint64_t)m.rows};
  const cuuint64_t strides[1] = {(cuuint64_t)m.nz * sizeof(float)};
  const cuuint32_t estr[2] = {1, 1};
  const cuuint32_t box[2] = {(cuuint32_t)TILE_PITCH, (cuuint32_t)TILE_LX};                       // whole-map sweeps
  const cuuint32_t box_s[2] = {(cuuint32_t)TILE_PITCH, (cuuint32_t)StampedGeo<STAMPED_TX>::LX};  // stamped sweeps
  const CUresult r = encode(reinterpret_cast<CUtensorMap*>(out->opaque), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                            m.height, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  const CUresult r2 = encode(reinterpret_cast<CUtensorMap*>(out->opaque_small), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                             m.height, dims, strides, box_s, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  out->valid = (r == CUDA_SUCCESS && r2 == CUDA_SUCCESS);
  return out->valid;
}

// dynamic shared memory above 48 KB needs an opt-in per function and per device
static void allow_smem(const void* kernel, int bytes) {
  static std::set<std::pair<int, const void*>> done;
  static std::mutex guard;                             // one host thread per GPU is the natural multi-GPU pattern
  std::lock_guard<std::mutex> lock(guard);
  int dev = 0;
  cudaGetDevice(&dev);
  if (done.count({dev, kernel})) return;
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  done.insert({dev, kernel});
}

// rows [row_lo, row_hi) are (re)computed and stored; the tiles covering them are launched. On a
// single GPU that is the whole map; a stripe computes its owned rows plus one ghost row each side
// from a halo of 5 rows below / 7 above (the <= 4/6-row dependency cone of the nine colours).
static void launch_tile_sweep(const MapView& m, const swe_params& p, const HeightTensorMap* tmap,
                              const unsigned char* stamp, unsigned char stamp_now, const unsigned* gate,
                              int row_lo, int row_hi, TileScratch scratch, cudaStream_t st, bool pdl = false) {
  CascadeThr thr;
  thr.ortho = sqrtf(1.0f) * p.maxdiff * 1.0f;      // host IEEE fp32: same values as the device helper
  thr.diag = sqrtf(2.0f) * p.maxdiff * 1.0f;
  if (row_lo < 0) row_lo = 0;
  if (row_hi > m.nx) row_hi = m.nx;
  if (row_hi <= row_lo) return;
  const int tx = stamp ? STAMPED_TX : TILE_TX;     // output rows per tile of the kernel used
  RowWindow win;
  win.store_lo = row_lo; win.store_hi = row_hi;
  win.tile_y0 = (row_lo - TILE_OX) / tx;
  const int tile_y1 = (row_hi - 1 - TILE_OX) / tx;
  const dim3 grid((m.nz - TILE_OZ + TILE_TZ - 1) / TILE_TZ, tile_y1 - win.tile_y0 + 1);
  const bool tma = tmap && tmap->valid;
  CUtensorMap tm;
  if (tma) memcpy(&tm, stamp ? tmap->opaque_small : tmap->opaque, sizeof tm); else memset(&tm, 0, sizeof tm);
  if (stamp) {
    using G = StampedGeo<STAMPED_TX>;
    const dim3 block(G::THREADS, 1);
    if (tma) {
      allow_smem(reinterpret_cast<const void*>(&cascade_tile_stamped_kernel<STAMPED_TX, true>), G::SMEM);
      launch_chain(cascade_tile_stamped_kernel<STAMPED_TX, true>, grid, block, G::SMEM, st, pdl, tm, m, p.settling, thr, stamp,
                   stamp_now, gate, win, scratch);
    } else {
      allow_smem(reinterpret_cast<const void*>(&cascade_tile_stamped_kernel<STAMPED_TX, false>), G::SMEM);
      launch_chain(cascade_tile_stamped_kernel<STAMPED_TX, false>, grid, block, G::SMEM, st, pdl, tm, m, p.settling, thr, stamp,
                   stamp_now, gate, win, scratch);
    }
    launch_chain(cascade_commit_kernel<STAMPED_TX>, grid, dim3(256), 0, st, pdl, m, win, (const float*)scratch.stage, scratch.flags);
  } else {
    const dim3 block(TILE_THREADS);
    if (tma) {
      allow_smem(reinterpret_cast<const void*>(&cascade_tile_kernel<true>), TILE_SMEM_PLAIN);
      cascade_tile_kernel<true><<<grid, block, TILE_SMEM_PLAIN, st>>>(tm, m, p.settling, thr, win, scratch);
    } else {
      allow_smem(reinterpret_cast<const void*>(&cascade_tile_kernel<false>), TILE_SMEM_PLAIN);
      cascade_tile_kernel<false><<<grid, block, TILE_SMEM_PLAIN, st>>>(tm, m, p.settling, thr, win, scratch);
    }
    cascade_commit_kernel<TILE_TX><<<grid, 256, 0, st>>>(m, win, scratch.stage, scratch.flags);
  }
}

void launch_cascade_stamped(const MapView& m, const swe_params& p, const HeightTensorMap* tmap,
                            const unsigned char* stamp, unsigned char stamp_now, const unsigned* gate,
                            int row_lo, int row_hi, TileScratch scratch, cudaStream_t st, bool pdl) {
  launch_tile_sweep(m, p, tmap, stamp, stamp_now, gate, row_lo, row_hi, scratch, st, pdl);
}

size_t tile_flag_count(const MapView& m) {
  // sized for the finer of the two tile grids (stamped sweeps use half-height tiles)
  return (size_t)((m.nz - TILE_OZ + TILE_TZ - 1) / TILE_TZ) * (size_t)((m.nx - TILE_OX + STAMPED_TX - 1) / STAMPED_TX);
}

int launch_cascade_sweep(const MapView& m, const swe_params& p, const HeightTensorMap* tmap, int sweeps,
                         DevStats* stats, cudaStream_t st, int variant, TileScratch scratch) {
  int launches = 0;
  if (m.periodic) variant = 1;   // the tile kernels load their halo with TMA zero-fill / clipped loads: no wrap
  for (int s = 0; s < sweeps; s++) {
    if (variant == 0 || variant == 2) {
      launch_tile_sweep(m, p, variant == 2 ? nullptr : tmap, nullptr, 0, nullptr, 0, m.nx, scratch, st);
      launches += 2;
      continue;
    }
    for (int c = 0; c < 9; c++) {   // v1: one launch per colour straight from global memory
      const int cx = c / 3, cz = c % 3;
      const int ncx = (m.nx - cx + 2) / 3, ncz = (m.nz - cz + 2) / 3;
      if (ncx <= 0 || ncz <= 0) continue;
      const long long total = (long long)ncx * ncz;
      int blocks = (int)((total + 127) / 128);
      if (blocks > 148 * 32) blocks = 148 * 32;
      cascade_colour_kernel<<<blocks, 128, 0, st>>>(m, p, cx, cz, ncx, ncz);
      launches++;
    }
  }
  add_cascade_cells_kernel<<<1, 1, 0, st>>>(stats, (unsigned long long)sweeps * m.nx * m.nz);
  return launches + 1;
}

void launch_max_excess(const MapView& m, const swe_params& p, float* d_out, cudaStream_t st) {
  set_float_kernel<<<1, 1, 0, st>>>(d_out, -INFINITY);
  const size_t n = (size_t)m.rows * m.nz;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  max_excess_kernel<<<blocks, 256, 0, st>>>(m, p, d_out);
}

// Force-load this translation unit's kernels (CUDA loads modules lazily on first launch, and a
// first launch issued while another stream spins in a stripe wait kernel would stall behind it).
void preload_cascade() {
  cudaFuncAttributes a;
  allow_smem(reinterpret_cast<const void*>(&cascade_tile_stamped_kernel<STAMPED_TX, true>), StampedGeo<STAMPED_TX>::SMEM);
  allow_smem(reinterpret_cast<const void*>(&cascade_tile_stamped_kernel<STAMPED_TX, false>), StampedGeo<STAMPED_TX>::SMEM);
  allow_smem(reinterpret_cast<const void*>(&cascade_tile_kernel<true>), TILE_SMEM_PLAIN);
  allow_smem(reinterpret_cast<const void*>(&cascade_tile_kernel<false>), TILE_SMEM_PLAIN);
  cudaFuncGetAttributes(&a, cascade_colour_kernel);
  cudaFuncGetAttributes(&a, cascade_tile_kernel<false>);
  cudaFuncGetAttributes(&a, cascade_tile_kernel<true>);
  cudaFuncGetAttributes(&a, cascade_tile_stamped_kernel<STAMPED_TX, false>);
  cudaFuncGetAttributes(&a, cascade_tile_stamped_kernel<STAMPED_TX, true>);
  cudaFuncGetAttributes(&a, cascade_commit_kernel<TILE_TX>);
  cudaFuncGetAttributes(&a, cascade_commit_kernel<STAMPED_TX>);
  cudaFuncGetAttributes(&a, max_excess_kernel);
  cudaFuncGetAttributes(&a, set_float_kernel);
  cudaFuncGetAttributes(&a, add_cascade_cells_kernel);
}

}  // namespace swe
