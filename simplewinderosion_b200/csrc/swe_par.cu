// swe_par.cu -- parallel flight: all particles of a World::erode call in flight at once.
//
// Pipeline of one parallel World::erode(cycles) (the launcher is in swe_api.cu):
//     spawn  ->  [ round kernel (K Wind::fly steps per live particle)  ->  stamped cascade sweep ] x R  -> EMA
// with R = ceil((maxAge+2)/K) rounds. A round kernel thread owns one particle: it loads the
// 32-byte particle record, runs K steps of the Wind::fly physics (source/wind.hpp:28-112)
// against the current map and stores the record back. Map updates are atomic scatter:
//   * one red.global.add.v4.f32 for the four track values (wind.hpp:93-96),
//   * one red.global.add.f32 for the erosion/deposition of the centre cell (wind.hpp:110-112),
//   * a byte stamp marking the cell as touched in this round.
// The slope relaxation the reference runs after every step (World::cascade at the particle's
// cell, wind.hpp:114) is NOT done per particle: concurrent particles applying full pairwise
// corrections from the same stale snapshot overshoot and diverge (measured: heights reach
// 1e10 within 40 steps). Instead, after each round every stamped cell gets World::cascade
// exactly once, in the conflict-free 3x3 colour order, by the shared-memory tile kernel of
// swe_cascade.cu -- exact Gauss-Seidel, unconditionally stable, deterministic given the map.
//
// Differences from the sequential reference (parity here is statistical, by construction):
// counter-based RNG (Philox4x32-10, counter = particle id / call / age; same distributions
// rand()%1001, rand()%1000; spawn cells stratified-uniform, see par_spawn_kernel), particles of
// a call see each other's deposits only through the map, fast-math arithmetic, one cascade per
// touched cell per round.
// Sediment carried by a particle when it dies leaves the system, as in the reference; it is
// accumulated in DevStats::sediment_lost so the mass balance closes exactly.
#include "swe_flight.cuh"

namespace swe {

// One Wind::fly step without the cascade. Returns 0 = alive, 1 = left the map, 2 = too old.
__device__ __forceinline__ int par_step(const MapView& m, const ParConst& C, PState& q, uint32_t pid,
                                        uint2 key, uint32_t call, unsigned char* stamp,
                                        unsigned char stamp_now) {
  const int ix = __float2int_rd(q.px + 0.5f), iz = __float2int_rd(q.pz + 0.5f);  // wind.hpp:30
  if (m.oob(ix, iz)) return 1;                                                    // :32-38
  const int nz = m.nz;
  const size_t ci = m.idx(ix, iz);
  float* hc_ptr = m.height + ci;
  const bool xm = ix > 0, xp = ix + 1 < m.nx, zm = iz > 0, zp = iz + 1 < nz;

  // the five heights quad::_normal reads (cellpool.h:200-223); ld.global.cg = L2-coherent
  const float hc = __ldcg(hc_ptr);
  const float hxm = xm ? __ldcg(hc_ptr - nz) : hc, hxp = xp ? __ldcg(hc_ptr + nz) : hc;
  const float hzm = zm ? __ldcg(hc_ptr - 1) : hc, hzp = zp ? __ldcg(hc_ptr + 1) : hc;

  // closed form of the four-triangle normal: plane (sx,sz) in {+,-}^2 counts iff the diagonal
  // (ix+sx, iz+sz) is in bounds; each plane is (-sx*S*dhx, 1, -sz*S*dhz).
  float nxv, nyv, nzv;
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

  red_add_v4(m.track + ci, q.sed, sx, sy, sz);                                    // :93-96
  {
    const float len = sqrtf(sx * sx + sy * sy + sz * sz);
    const float force = fmaxf(-(sx * nxv + sy * nyv + sz * nzv), 0.0f);           // :100-102
    const float lift = (1.0f - collision) * len;                                  // :104
    const float capacity = hfac * (force + 0.02f * lift);                         // :106
    float dh = C.suspension * (capacity - q.sed);                                 // :110-112
    if (m.rock && dh > 0.0f) {
      // layered extension: only the sand above the bedrock can be picked up; what the wind still
      // demands grinds the rock at `abrasion` efficiency (rock -> suspended sediment). Several
      // particles may stand on one cell in the same round, so where the sand is thin the pick-up
      // is a compare-and-swap loop that never takes the surface below the bedrock.
      const float rk = __ldcg(m.rock + ci);
      float take = dh;
      {
        // always a compare-and-swap loop in layered mode: several particles may stand on one cell in the same
        // round, and the sum of their takes must never bring the surface below the bedrock (sand >= 0)
        unsigned* hp = reinterpret_cast<unsigned*>(hc_ptr);
        unsigned old = __float_as_uint(hc), assumed;
        do {
          assumed = old;
          take = fminf(dh, fmaxf(__uint_as_float(assumed) - rk, 0.0f));
          if (!(take > 0.0f)) { take = 0.0f; break; }
          old = atomicCAS(hp, assumed, __float_as_uint(__uint_as_float(assumed) - take));
        } while (old != assumed);
      }
      const float ground = C.abrasion * (dh - take);
      if (ground > 0.0f) { red_add_f32(m.rock + ci, -ground); red_add_f32(hc_ptr, -ground); }
      dh = take + ground;
      q.sed += dh;
    } else {
      q.sed += dh;
      if (dh != 0.0f) red_add_f32(hc_ptr, -dh);
    }
  }
  stamp[ci] = stamp_now;                                                          // :114 (deferred)
  return 0;
}

__global__ void par_spawn_kernel(float4* __restrict__ pa, float4* __restrict__ pb, int n, int nx,
                                 int nz, uint2 key, uint32_t call, const CallCtl* __restrict__ ctl) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (ctl) { key = make_uint2(ctl->key_x, ctl->key_y); call += ctl->call; }   // graph replay: the argument is the batch salt
  float px, pz;
  spawn_position(i, n, nx, nz, key, call, px, pz);
  pa[i] = make_float4(px, 0.0f, pz, 0.0f);
  pb[i] = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(0));
}

// Stripe variant: the call's n particles are global; this rank materialises those of the
// candidate id range [i0, i1) whose first cell (wind.hpp:30: round(pos)) lies in its owned rows
// (particles spawned beyond the last row belong to the rank owning their spawn cell; they die at
// once, wind.hpp:32-38). Survivors are compacted into the local arrays together with their id.
__global__ void par_spawn_stripe_kernel(float4* __restrict__ pa, float4* __restrict__ pb,
                                        uint32_t* __restrict__ pid, unsigned* __restrict__ n_used,
                                        int capacity, unsigned* __restrict__ overflow, int i0, int i1,
                                        int n, int nx, int nz, int own_lo, int own_hi, uint2 key,
                                        uint32_t call) {
  const int i = i0 + (int)(blockIdx.x * blockDim.x + threadIdx.x);
  bool mine = false;
  float px = 0.0f, pz = 0.0f;
  if (i < i1) {
    spawn_position(i, n, nx, nz, key, call, px, pz);
    int row = __float2int_rd(px + 0.5f);
    if (row >= nx) row = nx - 1;
    mine = row >= own_lo && row < own_hi;
  }
  const unsigned mask = __ballot_sync(0xffffffffu, mine);
  if (!mask) return;
  const unsigned lane = threadIdx.x & 31u;
  unsigned base = 0;
  if (lane == (unsigned)(__ffs(mask) - 1)) base = atomicAdd(n_used, (unsigned)__popc(mask));
  base = __shfl_sync(0xffffffffu, base, __ffs(mask) - 1);
  if (!mine) return;
  const unsigned slot = base + __popc(mask & ((1u << lane) - 1u));
  if (slot >= (unsigned)capacity) { atomicAdd(overflow, 1u); return; }
  pa[slot] = make_float4(px, 0.0f, pz, 0.0f);
  pb[slot] = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(0));
  pid[slot] = (uint32_t)i;
}

// K steps for every live particle. live_in/live_out: device counters of particles alive
// before/after this round (trailing rounds of a call exit at once when nothing is alive).
// STRIPE: the particle arrays hold only this rank's particles (n = capacity, *n_used of them in
// use, ids in pid[]); a particle whose next cell lies in a neighbour's rows is appended to that
// neighbour's inbox through its peer-mapped pointers (NVLink stores + a system-scope atomic for
// the slot) and retired here.
#ifndef SWE_ROUND_MIN_BLOCKS
#define SWE_ROUND_MIN_BLOCKS 4   // resident CTAs of 256 threads per SM the register budget is tuned for
#endif
template <bool STRIPE>
__global__ void __launch_bounds__(256, SWE_ROUND_MIN_BLOCKS)
par_round_kernel(MapView m, ParConst C, float4* __restrict__ pa, float4* __restrict__ pb, int n,
                 int steps_per_round, uint2 key, uint32_t call, unsigned char* stamp,
                 unsigned char stamp_now, const unsigned* live_in, unsigned* live_out,
                 DevStats* stats, StripeRound sr, const CallCtl* __restrict__ ctl) {
  pdl_sync();
  if (live_in && *live_in == 0u) return;
  if (ctl) { key = make_uint2(ctl->key_x, ctl->key_y); call += ctl->call; }   // graph replay: the argument is the batch salt
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned steps = 0, n_oob = 0, n_age = 0, alive = 0;
  float lost = 0.0f;
  const int n_eff = STRIPE ? (int)min(*sr.n_used, (unsigned)n) : n;
  if (i < n_eff) {
    const float4 b = __ldcs(pb + i);   // records stream through once per round: keep L2 for the map
    PState q;
    q.age = __float_as_int(b.w);
    if (q.age >= 0) {
      const float4 a = __ldcs(pa + i);
      const uint32_t id = STRIPE ? sr.pid[i] : (uint32_t)i;
      q.px = a.x; q.py = a.y; q.pz = a.z; q.sed = a.w;
      q.sx = b.x; q.sy = b.y; q.sz = b.z;
      int r = 0;
      for (int s = 0; s < steps_per_round; s++) {
        r = par_step(m, C, q, id, key, call, stamp, stamp_now);
        if (r != 0) break;
        steps++;
      }
      if (r != 0) {
        lost = q.sed;
        if (r == 1) n_oob = 1; else n_age = 1;
        q.age = -1;
      } else {
        alive = 1;
        if (STRIPE) {
          const int row = __float2int_rd(q.px + 0.5f);
          const bool down = row < sr.own_lo && row >= 0, up = row >= sr.own_hi && row < m.nx;
          if (down || up) {   // hand the particle to the owner of its next cell
            const StripeOutbox& ob = down ? sr.below : sr.above;
            const unsigned slot = atomicAdd_system(ob.count, 1u);
            if (slot < (unsigned)ob.capacity) {
              ob.a[slot] = make_float4(q.px, q.py, q.pz, q.sed);
              ob.b[slot] = make_float4(q.sx, q.sy, q.sz, __int_as_float(q.age));
              ob.pid[slot] = id;
            } else {
              atomicAdd(sr.overflow, 1u);   // dropped: reported through swe_stats
              lost = q.sed;
            }
            q.age = -2;     // retired here; it stays counted as alive for the call
          }
        }
      }
      __stcs(pa + i, make_float4(q.px, q.py, q.pz, q.sed));
      __stcs(pb + i, make_float4(q.sx, q.sy, q.sz, __int_as_float(q.age)));
    }
  }
  // block-aggregate the counters: one set of atomics per block
  __shared__ unsigned s_cnt[4];
  __shared__ float s_lost;
  if (threadIdx.x == 0) { s_cnt[0] = s_cnt[1] = s_cnt[2] = s_cnt[3] = 0u; s_lost = 0.0f; }
  __syncthreads();
  for (int o = 16; o > 0; o >>= 1) {
    steps += __shfl_xor_sync(0xffffffffu, steps, o);
    n_oob += __shfl_xor_sync(0xffffffffu, n_oob, o);
    n_age += __shfl_xor_sync(0xffffffffu, n_age, o);
    alive += __shfl_xor_sync(0xffffffffu, alive, o);
    lost += __shfl_xor_sync(0xffffffffu, lost, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (steps) atomicAdd(&s_cnt[0], steps);
    if (n_oob) atomicAdd(&s_cnt[1], n_oob);
    if (n_age) atomicAdd(&s_cnt[2], n_age);
    if (alive) atomicAdd(&s_cnt[3], alive);
    if (lost != 0.0f) atomicAdd(&s_lost, lost);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (s_cnt[0]) atomicAdd(&stats->steps, (unsigned long long)s_cnt[0]);
    if (s_cnt[1]) atomicAdd(&stats->exits_oob, (unsigned long long)s_cnt[1]);
    if (s_cnt[2]) atomicAdd(&stats->exits_age, (unsigned long long)s_cnt[2]);
    if (s_cnt[3] && live_out) atomicAdd(live_out, s_cnt[3]);
    if (s_lost != 0.0f) atomicAdd(&stats->sediment_lost, (double)s_lost);
  }
}

__global__ void par_begin_kernel(DevStats* stats, unsigned* live, int n_live_slots, unsigned n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) { stats->particles += n; live[0] = n; }
  else if (i < n_live_slots) live[i] = 0u;
}

void launch_par_begin(const ParticleBuf& buf, const MapView& m, int n, uint64_t seed, uint32_t call,
                      const CallCtl* ctl, DevStats* stats, cudaStream_t st) {
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  par_begin_kernel<<<(buf.live_slots + 255) / 256, 256, 0, st>>>(stats, buf.live, buf.live_slots, (unsigned)n);
  par_spawn_kernel<<<(n + 255) / 256, 256, 0, st>>>(buf.a, buf.b, n, m.nx, m.nz, key, call, ctl);
}

void launch_par_round(const ParticleBuf& buf, const MapView& m, const swe_params& p, int n, int round,
                      int steps_per_round, uint64_t seed, uint32_t call, const CallCtl* ctl,
                      unsigned char* stamp, unsigned char stamp_now, DevStats* stats, cudaStream_t st,
                      bool pdl) {
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  StripeRound none{};
  launch_chain(par_round_kernel<false>, dim3((n + 255) / 256), dim3(256), 0, st, pdl, m, make_const(p), buf.a, buf.b, n,
               steps_per_round, key, call, stamp, stamp_now, (const unsigned*)(buf.live + round), buf.live + round + 1,
               stats, none, ctl);
}

// ---- stripe (multi-GPU) launchers ---------------------------------------------------------------
__global__ void stripe_begin_kernel(DevStats* stats, unsigned* n_used) { *n_used = 0u; (void)stats; }
__global__ void stripe_count_spawned_kernel(DevStats* stats, const unsigned* n_used, int capacity) {
  stats->particles += min(*n_used, (unsigned)capacity);
}

void launch_par_begin_stripe(const ParticleBuf& buf, const StripeRound& sr, const MapView& m, int n,
                             uint64_t seed, uint32_t call, DevStats* stats, cudaStream_t st) {
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  stripe_begin_kernel<<<1, 1, 0, st>>>(stats, sr.n_used);
  // candidate ids: strip-major rank order is monotone in the row, so the owned rows map to one
  // id interval; +-2 strips of slack cover the jitter u and the round() of the first cell
  const double per_cell = (double)n / ((double)m.nx * (double)m.nz);
  long long i0 = (long long)((double)(sr.own_lo - 8) * m.nz * per_cell) - 64;
  long long i1 = (long long)((double)(sr.own_hi + 8) * m.nz * per_cell) + 64;
  if (i0 < 0) i0 = 0;
  if (i1 > n) i1 = n;
  if (i1 > i0)
    par_spawn_stripe_kernel<<<(unsigned)((i1 - i0 + 255) / 256), 256, 0, st>>>(
        buf.a, buf.b, sr.pid, sr.n_used, buf.capacity, sr.overflow, (int)i0, (int)i1, n, m.nx, m.nz,
        sr.own_lo, sr.own_hi, key, call);
  stripe_count_spawned_kernel<<<1, 1, 0, st>>>(stats, sr.n_used, buf.capacity);
}

void launch_par_round_stripe(const ParticleBuf& buf, const StripeRound& sr, const MapView& m,
                             const swe_params& p, int steps_per_round, uint64_t seed, uint32_t call,
                             unsigned char* stamp, unsigned char stamp_now, DevStats* stats,
                             cudaStream_t st) {
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  par_round_kernel<true><<<(buf.capacity + 255) / 256, 256, 0, st>>>(
      m, make_const(p), buf.a, buf.b, buf.capacity, steps_per_round, key, call, stamp, stamp_now, nullptr,
      nullptr, stats, sr, nullptr);
}

// Force-load this translation unit's kernels (CUDA loads modules lazily on first launch, and a
// first launch issued while another stream spins in a stripe wait kernel would stall behind it).
void preload_par() {
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, par_spawn_kernel);
  cudaFuncGetAttributes(&a, par_spawn_stripe_kernel);
  cudaFuncGetAttributes(&a, par_round_kernel<false>);
  cudaFuncGetAttributes(&a, par_round_kernel<true>);
  cudaFuncGetAttributes(&a, par_begin_kernel);
  cudaFuncGetAttributes(&a, stripe_begin_kernel);
  cudaFuncGetAttributes(&a, stripe_count_spawned_kernel);
}

}  // namespace swe
