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
constexpr int TILE_TX = 60, TILE_TZ = 216;
constexpr int TILE_LX = TILE_TX + 6, TILE_PITCH = TILE_TZ + 24;
constexpr int TILE_BX = TILE_TZ / 3 + 8, TILE_BY = 6;       // 80 x 6 = 480 threads
constexpr int TILE_SMEM_BYTES = TILE_LX * TILE_PITCH * 4;   // 63,360 B

__device__ __forceinline__ void cswap_desc(unsigned& a, unsigned& b) {
  const unsigned hi = max(a, b), lo = min(a, b);
  a = hi; b = lo;
}

// generic exact fallback on the tile (same routine the single-cell path uses)
__device__ __noinline__ bool tile_cascade_generic(float* tile_c, int gx, int gz, int nx, int nz,
                                                  CascadeThr thr, float settling) {
  float h[8]; int slot[8]; int num = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const int dx = slot_dx(k), dz = slot_dz(k);
    const int x_ = gx + dx, z_ = gz + dz;
    if (x_ < 0 || z_ < 0 || x_ >= nx || z_ >= nz) continue;
    h[num] = tile_c[dx * TILE_PITCH + dz];
    slot[num] = k;
    num++;
  }
  unsigned changed;
  const float c1 = cascade_core(tile_c[0], h, slot, num, thr, settling, &changed);
  if (!changed) return false;
  tile_c[0] = c1;
  for (int i = 0; i < num; i++)
    if (changed & (1u << i)) tile_c[slot_dx(slot[i]) * TILE_PITCH + slot_dz(slot[i])] = h[i];
  return true;
}

template <bool USE_STAMP>
__global__ void __launch_bounds__(TILE_BX * TILE_BY, 2)
cascade_tile_kernel(MapView m, float settling, CascadeThr thr, const unsigned char* __restrict__ stamp,
                    unsigned char stamp_now, const unsigned* __restrict__ gate) {
  extern __shared__ float tile[];
  if (USE_STAMP && gate && *gate == 0u) return;
  const int ox0 = -2 + (int)blockIdx.y * TILE_TX, oz0 = -2 + (int)blockIdx.x * TILE_TZ;
  const int gx0 = ox0 - 2, gz0 = oz0 - 8;          // global coords of tile[0][0]
  const int tid = threadIdx.y * TILE_BX + threadIdx.x;
  const int nx = m.nx, nz = m.nz;
  const bool generic_only = !(thr.ortho >= 0.0f) || !(thr.diag >= 0.0f);

  // ---- load R_0 (cells outside the map are never read as neighbours) -------------------------
  for (int i = tid; i < TILE_LX * TILE_PITCH; i += TILE_BX * TILE_BY) {
    const int lx = i / TILE_PITCH, lz = i - lx * TILE_PITCH;
    const int gx = gx0 + lx, gz = gz0 + lz;
    float v = 0.0f;
    if (gx >= 0 && gx < nx && gz >= 0 && gz < nz) v = __ldg(m.height + m.idx(gx, gz));
    tile[i] = v;
  }
  __syncthreads();

  bool any_change = false;
#pragma unroll 1
  for (int c = 0; c < 9; c++) {
    const int cxc = c / 3;
    const int ncx = TILE_TX / 3 + 2 - cxc, ncz = TILE_TZ / 3 + 8 - c;
    if ((int)threadIdx.x < ncz) {
      const int lz = c + 1 + 3 * (int)threadIdx.x, gz = gz0 + lz;
      if (gz >= 0 && gz < nz) {
        const bool zm = gz > 0, zp = gz + 1 < nz;
        for (int cxi = threadIdx.y; cxi < ncx; cxi += TILE_BY) {
          const int lx = cxc + 1 + 3 * cxi, gx = gx0 + lx;
          if (gx < 0 || gx >= nx) continue;
          if (USE_STAMP && stamp[m.idx(gx, gz)] != stamp_now) continue;
          float* tc = tile + lx * TILE_PITCH + lz;
          const bool xm = gx > 0, xp = gx + 1 < nx;
          const float cen = tc[0];
          // reference slot order (world.hpp:83-92) = lexicographic in (dx, dz)
          const float h0 = tc[-TILE_PITCH - 1], h1 = tc[-TILE_PITCH], h2 = tc[-TILE_PITCH + 1];
          const float h3 = tc[-1], h4 = tc[1];
          const float h5 = tc[TILE_PITCH - 1], h6 = tc[TILE_PITCH], h7 = tc[TILE_PITCH + 1];
          const bool v0 = xm && zm, v1 = xm, v2 = xm && zp, v3 = zm, v4 = zp, v5 = xp && zm, v6 = xp,
                     v7 = xp && zp;
          // (2) exact early-out: no neighbour exceeds its repose gap -> nothing moves
          bool need = false;
          need |= v0 && !(fsub(fabsf(fsub(cen, h0)), thr.diag) <= 0.0f);
          need |= v1 && !(fsub(fabsf(fsub(cen, h1)), thr.ortho) <= 0.0f);
          need |= v2 && !(fsub(fabsf(fsub(cen, h2)), thr.diag) <= 0.0f);
          need |= v3 && !(fsub(fabsf(fsub(cen, h3)), thr.ortho) <= 0.0f);
          need |= v4 && !(fsub(fabsf(fsub(cen, h4)), thr.ortho) <= 0.0f);
          need |= v5 && !(fsub(fabsf(fsub(cen, h5)), thr.diag) <= 0.0f);
          need |= v6 && !(fsub(fabsf(fsub(cen, h6)), thr.ortho) <= 0.0f);
          need |= v7 && !(fsub(fabsf(fsub(cen, h7)), thr.diag) <= 0.0f);
          if (!need) continue;
          if (generic_only) { any_change |= tile_cascade_generic(tc, gx, gz, nx, nz, thr, settling); continue; }
          // (3) sort: key = order-preserving bits of (h + 0) [folds -0 onto +0], low 4 bits
          // = 15 - code, code = (dx+1)*4 + (dz+1); out-of-map neighbours get key 0 (last).
          auto mk = [](float h, bool valid, unsigned code) -> unsigned {
            const unsigned u = __float_as_uint(fadd(h, 0.0f));
            const unsigned k = u ^ ((unsigned)((int)u >> 31) | 0x80000000u);
            return valid ? ((k & ~15u) | (15u - code)) : 0u;
          };
          unsigned k0 = mk(h0, v0, 0), k1 = mk(h1, v1, 1), k2 = mk(h2, v2, 2), k3 = mk(h3, v3, 4),
                   k4 = mk(h4, v4, 6), k5 = mk(h5, v5, 8), k6 = mk(h6, v6, 9), k7 = mk(h7, v7, 10);
          cswap_desc(k0, k2); cswap_desc(k1, k3); cswap_desc(k4, k6); cswap_desc(k5, k7);
          cswap_desc(k0, k4); cswap_desc(k1, k5); cswap_desc(k2, k6); cswap_desc(k3, k7);
          cswap_desc(k0, k1); cswap_desc(k2, k3); cswap_desc(k4, k5); cswap_desc(k6, k7);
          cswap_desc(k2, k4); cswap_desc(k3, k5);
          cswap_desc(k1, k4); cswap_desc(k3, k6);
          cswap_desc(k1, k2); cswap_desc(k3, k4); cswap_desc(k5, k6);
          // heights that agree in everything but the 4 bits given to the code: order unknown
          const bool ambiguous =
              (k1 && (k0 ^ k1) < 16u) || (k2 && (k1 ^ k2) < 16u) || (k3 && (k2 ^ k3) < 16u) ||
              (k4 && (k3 ^ k4) < 16u) || (k5 && (k4 ^ k5) < 16u) || (k6 && (k5 ^ k6) < 16u) ||
              (k7 && (k6 ^ k7) < 16u);
          if (ambiguous) { any_change |= tile_cascade_generic(tc, gx, gz, nx, nz, thr, settling); continue; }
          float cc = cen;
          bool moved = false;
          const unsigned ks[8] = {k0, k1, k2, k3, k4, k5, k6, k7};
#pragma unroll
          for (int j = 0; j < 8; j++) {                                    // world.hpp:122-151
            const unsigned k = ks[j];
            if (k == 0u) break;                                            // only out-of-map left
            const unsigned code = 15u - (k & 15u);
            float* hp = tc + ((int)(code >> 2) - 1) * TILE_PITCH + ((int)(code & 3u) - 1);
            const float hv = *hp;
            const float t = (code & 5u) ? thr.ortho : thr.diag;
            const float diff = fsub(cc, hv);
            if (diff == 0.0f) continue;
            const float excess = fsub(fabsf(diff), t);
            if (excess <= 0.0f) continue;
            const float transfer = fmul(fmul(settling, excess), 0.5f);     // (settling*excess)/2
            if (diff > 0.0f) { cc = fsub(cc, transfer); *hp = fadd(hv, transfer); }
            else             { cc = fadd(cc, transfer); *hp = fsub(hv, transfer); }
            moved = true;
          }
          if (moved) { tc[0] = cc; any_change = true; }
        }
      }
    }
    __syncthreads();
  }

  // ---- store O if anything in this tile moved -----------------------------------------------
  if (__syncthreads_or(any_change ? 1 : 0)) {
    for (int i = tid; i < TILE_TX * TILE_TZ; i += TILE_BX * TILE_BY) {
      const int ox = i / TILE_TZ, oz = i - ox * TILE_TZ;
      const int gx = ox0 + ox, gz = oz0 + oz;
      if (gx >= 0 && gx < nx && gz >= 0 && gz < nz)
        m.height[m.idx(gx, gz)] = tile[(ox + 2) * TILE_PITCH + (oz + 8)];
    }
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

static void launch_tile_sweep(const MapView& m, const swe_params& p, const unsigned char* stamp,
                              unsigned char stamp_now, const unsigned* gate, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(cascade_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_SMEM_BYTES);
    cudaFuncSetAttribute(cascade_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_SMEM_BYTES);
    attr_set = true;
  }
  CascadeThr thr;
  thr.ortho = sqrtf(1.0f) * p.maxdiff * 1.0f;      // host IEEE fp32: same values as the device helper
  thr.diag = sqrtf(2.0f) * p.maxdiff * 1.0f;
  const dim3 grid((m.nz + 2 + TILE_TZ - 1) / TILE_TZ, (m.nx + 2 + TILE_TX - 1) / TILE_TX);
  const dim3 block(TILE_BX, TILE_BY);
  if (stamp)
    cascade_tile_kernel<true><<<grid, block, TILE_SMEM_BYTES, st>>>(m, p.settling, thr, stamp, stamp_now, gate);
  else
    cascade_tile_kernel<false><<<grid, block, TILE_SMEM_BYTES, st>>>(m, p.settling, thr, nullptr, 0, nullptr);
}

void launch_cascade_stamped(const MapView& m, const swe_params& p, const unsigned char* stamp,
                            unsigned char stamp_now, const unsigned* gate, cudaStream_t st) {
  launch_tile_sweep(m, p, stamp, stamp_now, gate, st);
}

int launch_cascade_sweep(const MapView& m, const swe_params& p, int sweeps, DevStats* stats,
                         cudaStream_t st, int variant) {
  int launches = 0;
  for (int s = 0; s < sweeps; s++) {
    if (variant == 0) {
      launch_tile_sweep(m, p, nullptr, 0, nullptr, st);
      launches++;
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

}  // namespace swe
