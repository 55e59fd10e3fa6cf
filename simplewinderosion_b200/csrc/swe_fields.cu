// swe_fields.cu -- bandwidth-bound full-map passes around the particle loop:
//   * EMA field update, World::erode world.hpp:70-76
//       field = (1-lrate)*field + lrate*track          (fp32, same expression order)
//     Algorithmic bytes: 32 B read + 16 B written per cell (+16 B/cell for the track reset
//     of world.hpp:38-44, done with cudaMemsetAsync at the start of swe_erode).
//   * plane <-> reference-AoS conversion for the host boundary (swe_upload_* / swe_download_*)
//   * fp64 plane sums (mass accounting)
// Compiled with -fmad=false so the EMA rounds like the reference.
#include "swe_common.cuh"

namespace swe {

__global__ void ema_kernel(float4* __restrict__ field, const float4* __restrict__ track, size_t n,
                           float lrate) {
  const float keep = __fsub_rn(1.0f, lrate);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    float4 f = field[i];
    const float4 t = __ldcs(track + i);
    f.x = __fadd_rn(__fmul_rn(keep, f.x), __fmul_rn(lrate, t.x));
    f.y = __fadd_rn(__fmul_rn(keep, f.y), __fmul_rn(lrate, t.y));
    f.z = __fadd_rn(__fmul_rn(keep, f.z), __fmul_rn(lrate, t.z));
    f.w = __fadd_rn(__fmul_rn(keep, f.w), __fmul_rn(lrate, t.w));
    field[i] = f;
  }
}

__device__ __forceinline__ float* plane_elem(const MapView& m, int plane, size_t i) {
  if (plane == SWE_HEIGHT) return m.height + i;
  float4* base = (plane <= SWE_MOMENTUMZ) ? m.field : m.track;
  const int comp = (plane <= SWE_MOMENTUMZ) ? plane - SWE_MASSFLOW : plane - SWE_MASSFLOW_TRACK;
  return reinterpret_cast<float*>(base + i) + comp;
}

__global__ void pack_plane_kernel(MapView m, int plane, const float* __restrict__ src) {
  const size_t n = (size_t)m.rows * m.nz;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    *plane_elem(m, plane, i) = src[i];
}

__global__ void unpack_plane_kernel(MapView m, int plane, float* __restrict__ dst) {
  const size_t n = (size_t)m.rows * m.nz;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    dst[i] = *plane_elem(m, plane, i);
}

// the four members of one float4 plane as four packed planes (dst[0..n), dst[n..2n), ...): one pass over the
// plane, float4 loads, coalesced stores -- the per-frame read-back of the flow fields (SimpleWindErosion.cpp:224-230)
__global__ void unpack_vec_plane_kernel(const float4* __restrict__ src, size_t n, float* __restrict__ dst) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = __ldcs(src + i);
    dst[i] = v.x; dst[n + i] = v.y; dst[2 * n + i] = v.z; dst[3 * n + i] = v.w;
  }
}

// reference AoS record = 9 floats in swe_plane order (cellpool.h:226-240)
__global__ void cells_to_planes_kernel(MapView m, const float* __restrict__ cells) {
  const size_t n = (size_t)m.rows * m.nz;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const float* c = cells + 9 * i;
    m.height[i] = c[0];
    m.field[i] = make_float4(c[1], c[2], c[3], c[4]);
    m.track[i] = make_float4(c[5], c[6], c[7], c[8]);
  }
}

__global__ void planes_to_cells_kernel(MapView m, float* __restrict__ cells) {
  const size_t n = (size_t)m.rows * m.nz;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    float* c = cells + 9 * i;
    const float4 f = m.field[i], t = m.track[i];
    c[0] = m.height[i];
    c[1] = f.x; c[2] = f.y; c[3] = f.z; c[4] = f.w;
    c[5] = t.x; c[6] = t.y; c[7] = t.z; c[8] = t.w;
  }
}

__global__ void sum_plane_kernel(MapView m, int plane, double* out) {
  const size_t n = (size_t)m.rows * m.nz;
  double acc = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    acc += (double)*plane_elem(m, plane, i);
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double warp_acc[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) warp_acc[warp] = acc;
  __syncthreads();
  if (warp == 0) {
    acc = (lane < (int)(blockDim.x >> 5)) ? warp_acc[lane] : 0.0;
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) atomicAdd(out, acc);
  }
}

static int grid_for(size_t n, int block, int per_sm) {
  size_t b = (n + block - 1) / block;
  const size_t cap = (size_t)148 * per_sm;
  return (int)(b < cap ? (b ? b : 1) : cap);
}

void launch_ema(const MapView& m, float lrate, cudaStream_t st) {
  const size_t n = (size_t)m.rows * m.nz;
  ema_kernel<<<grid_for(n, 256, 8), 256, 0, st>>>(m.field, m.track, n, lrate);
}
void launch_pack_plane(const MapView& m, int plane, const float* d_src, cudaStream_t st) {
  pack_plane_kernel<<<grid_for((size_t)m.rows * m.nz, 256, 8), 256, 0, st>>>(m, plane, d_src);
}
void launch_unpack_plane(const MapView& m, int plane, float* d_dst, cudaStream_t st) {
  unpack_plane_kernel<<<grid_for((size_t)m.rows * m.nz, 256, 8), 256, 0, st>>>(m, plane, d_dst);
}
void launch_unpack_field(const MapView& m, bool track, float* d_dst4, cudaStream_t st) {
  const size_t n = (size_t)m.rows * m.nz;
  unpack_vec_plane_kernel<<<grid_for(n, 256, 8), 256, 0, st>>>(track ? m.track : m.field, n, d_dst4);
}
void launch_cells_to_planes(const MapView& m, const float* d_cells, cudaStream_t st) {
  cells_to_planes_kernel<<<grid_for((size_t)m.rows * m.nz, 256, 8), 256, 0, st>>>(m, d_cells);
}
void launch_planes_to_cells(const MapView& m, float* d_cells, cudaStream_t st) {
  planes_to_cells_kernel<<<grid_for((size_t)m.rows * m.nz, 256, 8), 256, 0, st>>>(m, d_cells);
}
void launch_sum_plane(const MapView& m, int plane, double* d_out, cudaStream_t st) {
  cudaMemsetAsync(d_out, 0, sizeof(double), st);
  sum_plane_kernel<<<grid_for((size_t)m.rows * m.nz, 256, 4), 256, 0, st>>>(m, plane, d_out);
}

// Force-load this translation unit's kernels (CUDA loads modules lazily on first launch, and a
// first launch issued while another stream spins in a stripe wait kernel would stall behind it).
void preload_fields() {
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, ema_kernel);
  cudaFuncGetAttributes(&a, pack_plane_kernel);
  cudaFuncGetAttributes(&a, unpack_plane_kernel);
  cudaFuncGetAttributes(&a, unpack_vec_plane_kernel);
  cudaFuncGetAttributes(&a, cells_to_planes_kernel);
  cudaFuncGetAttributes(&a, planes_to_cells_kernel);
  cudaFuncGetAttributes(&a, sum_plane_kernel);
}

}  // namespace swe
