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

int launch_cascade_sweep(const MapView& m, const swe_params& p, int sweeps, DevStats* stats,
                         cudaStream_t st) {
  int launches = 0;
  for (int s = 0; s < sweeps; s++) {
    for (int c = 0; c < 9; c++) {
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
