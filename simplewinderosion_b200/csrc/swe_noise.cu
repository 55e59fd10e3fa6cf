// swe_noise.cu -- host generate moved onto the device: the noise layers of
// quad::map::init (source/cellpool.h:366-368 configuration, :388-411 layer loop -- a comment in
// the mounted snapshot --, :416-428 normalisation) with FastNoiseLite's OpenSimplex2 FBm
// (source/include/FastNoiseLite.h:322-340 GetNoise, :687-721 coordinate rotation, :866-887
// GenFractalFBm, :1054-1147 SingleOpenSimplex2 3D, :500-506 Hash, :542-553 GradCoord).
//
// The noise is plain fp32 and wrapping int32 arithmetic with no library call, so one thread per
// cell reproduces the host result BITWISE (checked against the unmodified FastNoiseLite class
// through oracle/_ref). Compiled with -fmad=false; products feeding sums go through the _rn
// intrinsics like the rest of the exact path. The Gaussian hill of :378-385 needs libm's
// double exp and stays on the host (include/swe_world.hpp).
//
// Cost: layers x 3 octaves x 2 lattices x ~70 instructions per cell, ALU-bound; 4 B/cell written.
#include "swe_exact.cuh"

namespace swe {
using namespace exact;

namespace {

constexpr int FNL_PX = 501125321, FNL_PY = 1136930381, FNL_PZ = 1720413743;
constexpr int MAX_LAYERS = 16;

struct LayerTable {
  int layers;
  float frequency[MAX_LAYERS];
  float scale[MAX_LAYERS];
};

// Lookup<float>::Gradients3D (FastNoiseLite.h:2529-2547) in closed form: entries 0..59 are five
// repeats of the twelve cube-edge directions (family = which coordinate is zero, then two sign
// bits), entries 60..63 repeat directions 8, 1, 9, 3.
__device__ __forceinline__ float grad_dot(int seed, int xp, int yp, int zp, float xd, float yd, float zd) {
  int h = (int)((unsigned)(seed ^ xp ^ yp ^ zp) * 0x27d4eb2du);
  h ^= h >> 15;
  const int g = (h >> 2) & 63;
  const int e = (g >= 60) ? ((0x3918 >> (4 * (g - 60))) & 15) : g % 12;   // nibbles 8,1,9,3
  const int family = e >> 2;
  const float s0 = (e & 1) ? -1.0f : 1.0f, s1 = (e & 2) ? -1.0f : 1.0f;
  const float gx = (family == 0) ? 0.0f : s0;
  const float gy = (family == 0) ? s0 : ((family == 1) ? 0.0f : s1);
  const float gz = (family == 2) ? 0.0f : s1;
  return fadd(fadd(fmul(xd, gx), fmul(yd, gy)), fmul(zd, gz));
}

// int32 arithmetic that wraps (the reference relies on it; signed overflow is UB in C++)
__device__ __forceinline__ int wmul(int a, int b) { return (int)((unsigned)a * (unsigned)b); }
__device__ __forceinline__ int wadd(int a, int b) { return (int)((unsigned)a + (unsigned)b); }
__device__ __forceinline__ int wsub(int a, int b) { return (int)((unsigned)a - (unsigned)b); }

__device__ __forceinline__ int fnl_round(float f) {
  return f >= 0 ? __float2int_rz(fadd(f, 0.5f)) : __float2int_rz(fsub(f, 0.5f));
}

__device__ __forceinline__ float pow4_times(float a, float g) {     // (a*a)*(a*a)*g
  const float a2 = fmul(a, a);
  return fmul(fmul(a2, a2), g);
}

// SingleOpenSimplex2, 3D: two offset cubic lattices; on each, the nearest lattice point and its
// neighbour along the axis of largest offset contribute falloff^4 * gradient.
__device__ float opensimplex2_3d(int seed, float x, float y, float z) {
  int li = fnl_round(x), lj = fnl_round(y), lk = fnl_round(z);
  float dx = fsub(x, (float)li), dy = fsub(y, (float)lj), dz = fsub(z, (float)lk);
  int sx = __float2int_rz(fsub(-1.0f, dx)) | 1, sy = __float2int_rz(fsub(-1.0f, dy)) | 1,
      sz = __float2int_rz(fsub(-1.0f, dz)) | 1;
  float ax = fmul((float)sx, -dx), ay = fmul((float)sy, -dy), az = fmul((float)sz, -dz);
  li = wmul(li, FNL_PX); lj = wmul(lj, FNL_PY); lk = wmul(lk, FNL_PZ);
  float value = 0.0f;
  float fall = fsub(fsub(0.6f, fmul(dx, dx)), fadd(fmul(dy, dy), fmul(dz, dz)));
#pragma unroll
  for (int lattice = 0; lattice < 2; lattice++) {
    if (fall > 0) value = fadd(value, pow4_times(fall, grad_dot(seed, li, lj, lk, dx, dy, dz)));
    float fall2 = fadd(fall, 1.0f);
    int i2 = li, j2 = lj, k2 = lk;
    float ex = dx, ey = dy, ez = dz;
    if (ax >= ay && ax >= az)     { ex = fadd(ex, (float)sx); fall2 = fsub(fall2, fmul((float)(sx * 2), ex)); i2 = wsub(i2, wmul(sx, FNL_PX)); }
    else if (ay > ax && ay >= az) { ey = fadd(ey, (float)sy); fall2 = fsub(fall2, fmul((float)(sy * 2), ey)); j2 = wsub(j2, wmul(sy, FNL_PY)); }
    else                          { ez = fadd(ez, (float)sz); fall2 = fsub(fall2, fmul((float)(sz * 2), ez)); k2 = wsub(k2, wmul(sz, FNL_PZ)); }
    if (fall2 > 0) value = fadd(value, pow4_times(fall2, grad_dot(seed, i2, j2, k2, ex, ey, ez)));
    if (lattice == 1) break;
    ax = fsub(0.5f, ax); ay = fsub(0.5f, ay); az = fsub(0.5f, az);
    dx = fmul((float)sx, ax); dy = fmul((float)sy, ay); dz = fmul((float)sz, az);
    fall = fadd(fall, fsub(fsub(0.75f, ax), fadd(ay, az)));
    li = wadd(li, (sx >> 1) & FNL_PX); lj = wadd(lj, (sy >> 1) & FNL_PY); lk = wadd(lk, (sz >> 1) & FNL_PZ);
    sx = -sx; sy = -sy; sz = -sz;
    seed = ~seed;
  }
  return fmul(value, 32.69428253173828125f);
}

// GetNoise(x, y, z) of the object map::init sets up: rotate, then three FBm octaves from seed 1337.
__device__ float fnl_getnoise(float frequency, float x, float y, float z) {
  x = fmul(x, frequency); y = fmul(y, frequency); z = fmul(z, frequency);
  const float R3 = (float)(2.0 / 3.0);
  const float r = fmul(fadd(fadd(x, y), z), R3);
  x = fsub(r, x); y = fsub(r, y); z = fsub(r, z);
  int seed = 1337;
  float sum = 0.0f;
  float amp = 1 / 1.75f;
#pragma unroll 1
  for (int o = 0; o < 3; o++) {
    const float noise = opensimplex2_3d(seed++, x, y, z);
    sum = fadd(sum, fmul(noise, amp));
    // amp *= Lerp(1, (noise+1)/2, weightedStrength = 0): exactly 1 for finite noise
    x = fmul(x, 2.0f); y = fmul(y, 2.0f); z = fmul(z, 2.0f);
    amp = fmul(amp, 0.5f);
  }
  return sum;
}

__global__ void __launch_bounds__(256)
noise_stack_kernel(MapView m, LayerTable lt, float tile, float zc) {
  const size_t n = (size_t)m.rows * m.nz;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int x = m.x0 + (int)(i / m.nz), z = (int)(i % m.nz);
    if (x < 0 || x >= m.nx) continue;                 // ghost rows beyond the map edge
    const float px = fdiv((float)x, tile), pz = fdiv((float)z, tile);
    float h = m.height[i];
#pragma unroll 1
    for (int o = 0; o < lt.layers; o++)
      h = fadd(h, fmul(lt.scale[o], fnl_getnoise(lt.frequency[o], px, pz, zc)));
    m.height[i] = h;
  }
}

// min/max as the reference folds them (both start at 0.0f; cellpool.h:416-423); order-free.
__global__ void height_range_kernel(const float* __restrict__ h, size_t n, int* out2) {
  float mn = 0.0f, mx = 0.0f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = h[i];
    mn = (mn < v) ? mn : v;
    mx = (mx > v) ? mx : v;
  }
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) {
    // mn <= 0 <= mx: negative floats order inversely as unsigned, non-negative ones as ints
    atomicMax(reinterpret_cast<unsigned*>(out2), __float_as_uint(mn));
    atomicMax(out2 + 1, __float_as_int(mx));
  }
}

__global__ void normalize_kernel(float* __restrict__ h, size_t n, float mn, float mx) {
  const float span = fsub(mx, mn);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    h[i] = fmul(0.5f, fdiv(fsub(h[i], mn), span));    // 0.5*(double)q rounds to 0.5f*q exactly
}

int grid_for(size_t n, int block, int per_sm) {
  const size_t b = (n + block - 1) / block, cap = (size_t)148 * per_sm;
  return (int)(b < cap ? (b ? b : 1) : cap);
}

}  // namespace

int noise_max_layers() { return MAX_LAYERS; }

void launch_noise_stack(const MapView& m, int seed, int layers, int tile, cudaStream_t st) {
  LayerTable lt;
  lt.layers = layers;
  float frequency = 1.0f, scale = 0.6f;               // cellpool.h:394-395
  for (int o = 0; o < MAX_LAYERS; o++) {
    lt.frequency[o] = frequency; lt.scale[o] = scale;
    frequency *= 2;                                   // :406
    scale = (float)((double)scale * 0.6);             // :407  (scale *= 0.6 with a double literal)
  }
  const size_t n = (size_t)m.rows * m.nz;
  noise_stack_kernel<<<grid_for(n, 256, 8), 256, 0, st>>>(m, lt, (float)tile, (float)(seed % 10000));
}

// d_out2: two ints of scratch; afterwards they hold the bit patterns of {min, max}
void launch_height_range(const float* d_height, size_t n, int* d_out2, cudaStream_t st) {
  cudaMemsetAsync(d_out2, 0, 2 * sizeof(int), st);   // 0.0f, 0.0f
  height_range_kernel<<<grid_for(n, 256, 8), 256, 0, st>>>(d_height, n, d_out2);
}

void launch_normalize(float* d_height, size_t n, float mn, float mx, cudaStream_t st) {
  normalize_kernel<<<grid_for(n, 256, 8), 256, 0, st>>>(d_height, n, mn, mx);
}

void preload_noise() {
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, noise_stack_kernel);
  cudaFuncGetAttributes(&a, height_range_kernel);
  cudaFuncGetAttributes(&a, normalize_kernel);
}

}  // namespace swe
